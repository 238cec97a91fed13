// The sign classifier behind the key points: the Keras Sequential of demo_isl_translate.py:72-99 applied to a
// 20 x 156 window of feature rows (ISL_Model_parameter.py:322-353), inference only, float32 like Keras.
//   Masking(0) -> BatchNorm -> BiLSTM(32, sequences) -> BiLSTM(32) -> ELU -> Dense(32) -> BN -> ELU -> Dense(32) -> BN -> ELU
//   -> Dense(classes) + softmax
// One CTA per window, 256 threads = 2 directions x 128 gate columns (4 gates x 32 units, Keras order i, f, c, o). Per layer the
// input projections of all T steps are formed first (thread = gate column, weights read once, coalesced; the window sits in
// shared memory), then the recurrence runs with the thread's recurrent-kernel column in registers: two barriers per step.
// Masked steps (all 156 features zero) keep the states and repeat the previous output, as keras' rnn() does. Latency bound by
// construction (20 dependent steps per layer); windows are independent, so a batch of windows fills the device.
#include "prepost.cuh"

namespace islpose {

namespace {

constexpr int kTrUnits = 32;             // LSTM(32)
constexpr int kTrGates = 4 * kTrUnits;   // 128 gate columns per direction
constexpr int kTrMaxT = 32;              // window length (the reference fixes 20)
constexpr float kBnEps = 1e-3f;          // keras BatchNormalization default

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float eluf_(float x) { return x > 0.0f ? x : expf(x) - 1.0f; }

// One bidirectional LSTM layer. xin: shared [T][K]; w6: kernel [K][128], recurrent [32][128], bias [128] of the forward then
// of the backward layer; zin: shared [2][T][128]; gates: shared [2][128]; hc: shared [2][2][32] (h, c per direction).
// seq_out (may be null): shared [T][64], forward outputs at columns 0..31, backward outputs (flipped back) at 32..63.
// On return hc[d][0][*] holds each direction's last output.
template <int K>
__device__ void bilstm_layer(const float* __restrict__ xin, const unsigned mask_bits, int T, const float* __restrict__ kern_f,
                             const float* __restrict__ rec_f, const float* __restrict__ bias_f, const float* __restrict__ kern_b,
                             const float* __restrict__ rec_b, const float* __restrict__ bias_b, float* zin, float* gates, float* hc,
                             float* seq_out) {
  const int d = threadIdx.x >> 7, j = threadIdx.x & 127;
  const float* kern = d ? kern_b : kern_f;
  const float* rec = d ? rec_b : rec_f;
  const float b = (d ? bias_b : bias_f)[j];
  // input projections of every step: acc[t] = bias + sum_k x[t][k] * kernel[k][j]
  float acc[kTrMaxT];
#pragma unroll
  for (int t = 0; t < kTrMaxT; ++t) acc[t] = b;
  for (int k = 0; k < K; ++k) {
    const float w = __ldg(kern + k * kTrGates + j);
#pragma unroll
    for (int t = 0; t < kTrMaxT; ++t) {
      if (t < T) acc[t] = fmaf(xin[t * K + k], w, acc[t]);
    }
  }
#pragma unroll
  for (int t = 0; t < kTrMaxT; ++t) {
    if (t < T) zin[(d * kTrMaxT + t) * kTrGates + j] = acc[t];
  }
  float u[kTrUnits];
#pragma unroll
  for (int k = 0; k < kTrUnits; ++k) u[k] = __ldg(rec + k * kTrGates + j);
  if (j < 2 * kTrUnits) hc[d * 2 * kTrUnits + j] = 0.0f;  // h and c start at zero
  __syncthreads();
  for (int s = 0; s < T; ++s) {
    const int t = d ? T - 1 - s : s;
    const bool live = (mask_bits >> t) & 1u;  // uniform over the CTA's 128 threads of this direction
    if (live) {
      float z = zin[(d * kTrMaxT + t) * kTrGates + j];
      const float* h = hc + d * 2 * kTrUnits;
#pragma unroll
      for (int k = 0; k < kTrUnits; ++k) z = fmaf(h[k], u[k], z);
      gates[d * kTrGates + j] = z;
    }
    __syncthreads();
    if (j < kTrUnits) {
      float* h = hc + d * 2 * kTrUnits;
      float* c = h + kTrUnits;
      if (live) {
        const float* g = gates + d * kTrGates;
        const float gi = sigmoidf_(g[j]), gf = sigmoidf_(g[kTrUnits + j]), gc = tanhf(g[2 * kTrUnits + j]),
                    go = sigmoidf_(g[3 * kTrUnits + j]);
        const float cn = fmaf(gf, c[j], gi * gc);
        c[j] = cn;
        h[j] = go * tanhf(cn);
      }
      if (seq_out != nullptr) seq_out[t * 2 * kTrUnits + d * kTrUnits + j] = h[j];  // a masked step repeats the previous output
    }
    __syncthreads();
  }
}

struct TranslateWeights {
  const float* bn0;   // gamma, beta, mean, var: 4 x F
  const float* l1[6]; // forward kernel / recurrent / bias, backward kernel / recurrent / bias
  const float* l2[6];
  const float* d1;    // [64][32]
  const float* bn1;   // 4 x 32
  const float* d2;    // [32][32]
  const float* bn2;   // 4 x 32
  const float* d3;    // [32][classes]
  const float* b3;    // [classes]
};

constexpr int kTrF = 156;  // features per frame (ISL_Model_parameter.py:376-410)

__global__ void __launch_bounds__(256)
translate_kernel(const double* __restrict__ windows, int T, const TranslateWeights w, int classes, float* __restrict__ probs) {
  extern __shared__ float s_tr[];
  float* x = s_tr;                              // [T][156] normalised window
  float* zin = x + kTrMaxT * kTrF;              // [2][32][128]
  float* seq = zin + 2 * kTrMaxT * kTrGates;    // [T][64]
  float* gates = seq + kTrMaxT * 2 * kTrUnits;  // [2][128]
  float* hc = gates + 2 * kTrGates;             // [2][2][32]
  float* v = hc + 4 * kTrUnits;                 // [64] head scratch
  float* logits = v + 2 * kTrUnits;             // [classes]
  __shared__ unsigned s_mask;
  __shared__ float s_red[2];
  const double* win = windows + static_cast<long long>(blockIdx.x) * T * kTrF;
  if (threadIdx.x == 0) s_mask = 0u;
  __syncthreads();
  // Masking(mask_value=0.): a step is live when any of its features (as float32) is non-zero; then BatchNorm
  for (int e = threadIdx.x; e < T * kTrF; e += blockDim.x) {
    const int t = e / kTrF, k = e - t * kTrF;
    const float val = static_cast<float>(win[e]);
    if (val != 0.0f) atomicOr(&s_mask, 1u << t);
    const float g = __ldg(w.bn0 + k), be = __ldg(w.bn0 + kTrF + k), mu = __ldg(w.bn0 + 2 * kTrF + k), var = __ldg(w.bn0 + 3 * kTrF + k);
    x[e] = (val - mu) / sqrtf(var + kBnEps) * g + be;
  }
  __syncthreads();
  const unsigned mask_bits = s_mask;
  bilstm_layer<kTrF>(x, mask_bits, T, w.l1[0], w.l1[1], w.l1[2], w.l1[3], w.l1[4], w.l1[5], zin, gates, hc, seq);
  bilstm_layer<2 * kTrUnits>(seq, mask_bits, T, w.l2[0], w.l2[1], w.l2[2], w.l2[3], w.l2[4], w.l2[5], zin, gates, hc, nullptr);
  // head: [h_forward | h_backward] -> ELU -> Dense -> BN -> ELU -> Dense -> BN -> ELU -> Dense + bias -> softmax
  const int tid = threadIdx.x;
  if (tid < 2 * kTrUnits) v[tid] = eluf_(hc[(tid >> 5) * 2 * kTrUnits + (tid & 31)]);
  __syncthreads();
  float r = 0.0f;
  if (tid < kTrUnits) {
    for (int k = 0; k < 2 * kTrUnits; ++k) r = fmaf(v[k], __ldg(w.d1 + k * kTrUnits + tid), r);
    r = (r - __ldg(w.bn1 + 2 * kTrUnits + tid)) / sqrtf(__ldg(w.bn1 + 3 * kTrUnits + tid) + kBnEps) * __ldg(w.bn1 + tid) +
        __ldg(w.bn1 + kTrUnits + tid);
    r = eluf_(r);
  }
  __syncthreads();
  if (tid < kTrUnits) v[tid] = r;
  __syncthreads();
  if (tid < kTrUnits) {
    r = 0.0f;
    for (int k = 0; k < kTrUnits; ++k) r = fmaf(v[k], __ldg(w.d2 + k * kTrUnits + tid), r);
    r = (r - __ldg(w.bn2 + 2 * kTrUnits + tid)) / sqrtf(__ldg(w.bn2 + 3 * kTrUnits + tid) + kBnEps) * __ldg(w.bn2 + tid) +
        __ldg(w.bn2 + kTrUnits + tid);
    r = eluf_(r);
  }
  __syncthreads();
  if (tid < kTrUnits) v[tid] = r;
  __syncthreads();
  for (int cidx = tid; cidx < classes; cidx += blockDim.x) {
    float z = __ldg(w.b3 + cidx);
    for (int k = 0; k < kTrUnits; ++k) z = fmaf(v[k], __ldg(w.d3 + k * classes + cidx), z);
    logits[cidx] = z;
  }
  __syncthreads();
  if (tid == 0) {  // classes is a few hundred at most: a serial max and sum keep the summation order fixed
    float m = logits[0];
    for (int cidx = 1; cidx < classes; ++cidx) m = fmaxf(m, logits[cidx]);
    float sum = 0.0f;
    for (int cidx = 0; cidx < classes; ++cidx) sum += expf(logits[cidx] - m);
    s_red[0] = m;
    s_red[1] = sum;
  }
  __syncthreads();
  for (int cidx = tid; cidx < classes; cidx += blockDim.x)
    probs[static_cast<long long>(blockIdx.x) * classes + cidx] = expf(logits[cidx] - s_red[0]) / s_red[1];
}

}  // namespace

long long translate_weight_floats(int classes) {
  const int F = kTrF, U = kTrUnits, G = kTrGates;
  return 4LL * F + 2LL * (F * G + U * G + G) + 2LL * (2 * U * G + U * G + G) + 2LL * U * U + 4LL * U + 1LL * U * U + 4LL * U +
         1LL * U * classes + classes;
}

int launch_translate(const double* windows, int n, int T, const float* weights, long long n_weights, int classes, float* probs,
                     cudaStream_t st) {
  if (T <= 0 || T > kTrMaxT || classes <= 0 || classes > 4096 || n_weights != translate_weight_floats(classes)) return 1;
  const int F = kTrF, U = kTrUnits, G = kTrGates;
  TranslateWeights w;
  const float* p = weights;  // translation_model.get_weights() order (demo_isl_translate.py:72-99)
  w.bn0 = p;
  p += 4 * F;
  for (int d = 0; d < 2; ++d) {
    w.l1[3 * d + 0] = p;
    p += F * G;
    w.l1[3 * d + 1] = p;
    p += U * G;
    w.l1[3 * d + 2] = p;
    p += G;
  }
  for (int d = 0; d < 2; ++d) {
    w.l2[3 * d + 0] = p;
    p += 2 * U * G;
    w.l2[3 * d + 1] = p;
    p += U * G;
    w.l2[3 * d + 2] = p;
    p += G;
  }
  w.d1 = p;
  p += 2 * U * U;
  w.bn1 = p;
  p += 4 * U;
  w.d2 = p;
  p += U * U;
  w.bn2 = p;
  p += 4 * U;
  w.d3 = p;
  p += U * classes;
  w.b3 = p;
  const size_t smem = sizeof(float) * (kTrMaxT * kTrF + 2 * kTrMaxT * kTrGates + kTrMaxT * 2 * kTrUnits + 2 * kTrGates + 4 * kTrUnits +
                                       2 * kTrUnits + classes);
  static bool attr_dev[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_dev[dev & 63]) {
    if (cudaFuncSetAttribute(translate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024) != cudaSuccess) return 1;
    attr_dev[dev & 63] = true;
  }
  if (smem > 96 * 1024) return 1;
  translate_kernel<<<n, 256, smem, st>>>(windows, T, w, classes, probs);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace islpose
