// Memory-bound kernels either side of the networks:
//   resize_pad_norm   body.py:53-56 / hand.py:37-40   uint8 cubic resize + pad(128) + x/256-0.5 -> fp32 NCHW
//   im2col3x3         first layer's 3x3x3 patches -> bf16 [N,h,w,32] so conv1_1 runs as a 1x1 GEMM (K=27->32)
//   maxpool2x2        model.py:30-32, NHWC bf16
//   heat_accumulate   body.py:69-72,80 / hand.py:51-56  two cubic resizes + scale accumulation, float64 planes
//   gauss_nms         body.py:88-107  scipy gaussian_filter(sigma=3) + 4-neighbour NMS + threshold -> peak lists
//   gauss_smooth      hand.py:61      same filter, smoothed plane written out
//   sort_peaks        np.nonzero row-major order for the atomically appended peak lists
#include "prepost.cuh"

#include <cuda_bf16.h>
#include <stdlib.h>

#include "cubic.cuh"

namespace islpose {

// ------------------------------------------------------------------------------------------------ resize
__global__ void resize_pad_norm_kernel(const uint8_t* __restrict__ in, int N, int H, int W, int rh, int rw, int hp,
                                       int wp, double scale_x, double scale_y, float* __restrict__ out_nchw,
                                       uint8_t* __restrict__ out_u8) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int n = blockIdx.z;
  if (x >= wp || y >= hp) return;
  int v[3] = {128, 128, 128};  // padValue (body.py:43)
  if (x < rw && y < rh) {
    int sx, sy;
    float fx, fy;
    float cx[4], cy[4];
    cubic_src(x, scale_x, sx, fx);
    cubic_coeffs(fx, cx);
    cubic_src(y, scale_y, sy, fy);
    cubic_coeffs(fy, cy);
    int ia[4];
    float bf[4];
    int xi[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      // saturate_cast<short>(w * INTER_RESIZE_COEF_SCALE): round half to even, 11 fractional bits
      ia[k] = clampi(__float2int_rn(__fmul_rn(cx[k], 2048.f)), -32768, 32767);
      const int ib = clampi(__float2int_rn(__fmul_rn(cy[k], 2048.f)), -32768, 32767);
      bf[k] = __fmul_rn(static_cast<float>(ib), 1.f / 4194304.f);
      xi[k] = clampi(sx - 1 + k, 0, W - 1);
    }
    const uint8_t* img = in + static_cast<long long>(n) * H * W * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float hs[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint8_t* row = img + static_cast<long long>(clampi(sy - 1 + k, 0, H - 1)) * W * 3 + c;
        const int acc = row[xi[0] * 3] * ia[0] + row[xi[1] * 3] * ia[1] + row[xi[2] * 3] * ia[2] + row[xi[3] * 3] * ia[3];
        hs[k] = static_cast<float>(acc);
      }
      // vertical pass in float, SIMD order, round half to even, saturate to uint8
      v[c] = clampi(__float2int_rn(dot4_rl(hs[0], hs[1], hs[2], hs[3], bf)), 0, 255);
    }
  }
  const long long plane = static_cast<long long>(hp) * wp;
  const long long pix = static_cast<long long>(y) * wp + x;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    out_nchw[(static_cast<long long>(n) * 3 + c) * plane + pix] = __fsub_rn(__fdiv_rn(static_cast<float>(v[c]), 256.f), 0.5f);
    if (out_u8 != nullptr) out_u8[(static_cast<long long>(n) * plane + pix) * 3 + c] = static_cast<uint8_t>(v[c]);
  }
}

// ------------------------------------------------------------------------------------------------ im2col
__global__ void im2col3x3_kernel(const float* __restrict__ in, int N, int h, int w, __nv_bfloat16* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int n = blockIdx.z;
  if (x >= w || y >= h) return;
  const long long plane = static_cast<long long>(h) * w;
  const float* img = in + static_cast<long long>(n) * 3 * plane;
  __align__(16) __nv_bfloat16 v[32];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int yy = y + ky - 1, xx = x + kx - 1;
      const bool in_img = yy >= 0 && yy < h && xx >= 0 && xx < w;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float f = in_img ? __ldg(img + c * plane + static_cast<long long>(yy) * w + xx) : 0.f;
        v[(ky * 3 + kx) * 3 + c] = __float2bfloat16_rn(f);
      }
    }
  }
#pragma unroll
  for (int i = 27; i < 32; ++i) v[i] = __float2bfloat16_rn(0.f);
  uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<long long>(n) * plane + static_cast<long long>(y) * w + x) * 32);
  const uint4* src = reinterpret_cast<const uint4*>(v);
#pragma unroll
  for (int i = 0; i < 4; ++i) dst[i] = src[i];
}

// ------------------------------------------------------------------------------------------------ max pool
__global__ void maxpool2x2_kernel(const __nv_bfloat16* __restrict__ in, int N, int H, int W, int C,
                                  __nv_bfloat16* __restrict__ out) {
  const int c8 = C / 8;
  const int Ho = H / 2, Wo = W / 2;
  const long long total = static_cast<long long>(N) * Ho * Wo * c8;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cg = i % c8;
    long long p = i / c8;
    const int xo = p % Wo;
    p /= Wo;
    const int yo = p % Ho;
    const int n = p / Ho;
    const __nv_bfloat16* base = in + ((static_cast<long long>(n) * H + 2 * yo) * W + 2 * xo) * C + cg * 8;
    const uint4 a = *reinterpret_cast<const uint4*>(base);
    const uint4 b = *reinterpret_cast<const uint4*>(base + C);
    const uint4 c = *reinterpret_cast<const uint4*>(base + static_cast<long long>(W) * C);
    const uint4 d = *reinterpret_cast<const uint4*>(base + static_cast<long long>(W) * C + C);
    uint4 r;
    const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
    const __nv_bfloat162* pc = reinterpret_cast<const __nv_bfloat162*>(&c);
    const __nv_bfloat162* pd = reinterpret_cast<const __nv_bfloat162*>(&d);
    __nv_bfloat162* pr = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int k = 0; k < 4; ++k) pr[k] = __hmax2(__hmax2(pa[k], pb[k]), __hmax2(pc[k], pd[k]));
    *reinterpret_cast<uint4*>(out + ((static_cast<long long>(n) * Ho + yo) * Wo + xo) * C + cg * 8) = r;
  }
}

// ------------------------------------------------------------------------------------------------ accumulate
// One thread = one frame pixel x kChunk channels. For every scale: both cubic stages from the stride-8 map,
// the float32 division by the number of scales, then the float64 accumulation of the reference
// (body.py:80 `avg += avg + m/S` when q1, body.py:81 / hand.py:56 `avg += m/S` otherwise).
constexpr int kChunk = 4;
__global__ void __launch_bounds__(256)
heat_accumulate_kernel(const ScaleSet ss, int N, int H, int W, int parts, int q1, double* __restrict__ out) {
  __shared__ float s_tab[8][4];
  fill_phase_table(s_tab);
  __syncthreads();
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int chunks = (parts + kChunk - 1) / kChunk;
  const int n = blockIdx.z / chunks;
  const int c0 = (blockIdx.z % chunks) * kChunk;
  if (x >= W || y >= H) return;
  double acc[kChunk];
#pragma unroll
  for (int i = 0; i < kChunk; ++i) acc[i] = 0.0;
  const int C = ss.channels;
  const long long tail_start = (static_cast<long long>(W) * C) / 4 * 4;
  const float fS = static_cast<float>(ss.count);
  for (int s = 0; s < ss.count; ++s) {
    const ScaleGeom& g = ss.g[s];
    Axis2 ax, ay;
    make_axis2(x, g.sx, g.wc, g.gw, s_tab, ax);
    make_axis2(y, g.sy, g.hc, g.gh, s_tab, ay);
    const long long plane = static_cast<long long>(g.gh) * g.gw;
#pragma unroll
    for (int i = 0; i < kChunk; ++i) {
      const int c = c0 + i;
      if (c < parts) {
        const bool tail = static_cast<long long>(x) * C + c >= tail_start;
        const float v = sample2(g.low + (static_cast<long long>(n) * C + c) * plane, g.gw, ax, ay, tail);
        const double t = static_cast<double>(__fdiv_rn(v, fS));
        acc[i] = q1 ? __dadd_rn(acc[i], __dadd_rn(acc[i], t)) : __dadd_rn(acc[i], t);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kChunk; ++i) {
    const int c = c0 + i;
    if (c < parts) out[((static_cast<long long>(n) * parts + c) * H + y) * W + x] = acc[i];
  }
}

// Two-pass variant of the same arithmetic (bit-identical results, ~8x fewer multiply-adds): pass 1 materialises
// stage 1 - the x8 up-sampled, cropped map of one scale, float32 [N][parts][hc][wc] - with 20 multiply-adds per
// up-sampled pixel; pass 2 is stage 2 (4x4 taps from that map) plus the division and the float64 accumulation.
__global__ void __launch_bounds__(256)
upsample8_kernel(const float* __restrict__ low, int C, int parts, int gh, int gw, int hc, int wc,
                 float* __restrict__ mid) {
  __shared__ float s_tab[8][4];
  fill_phase_table(s_tab);
  __syncthreads();
  const int u = blockIdx.x * 32 + (threadIdx.x & 31);
  const int v = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int n = blockIdx.z / parts;
  const int c = blockIdx.z - n * parts;
  if (u >= wc || v >= hc) return;
  const int su = ((u + 4) >> 3) - 2, sv = ((v + 4) >> 3) - 2;  // first tap = floor(src) - 1
  const float* wx = s_tab[u & 7];
  const float* wy = s_tab[v & 7];
  const float wxr[4] = {wx[0], wx[1], wx[2], wx[3]};
  const float wyr[4] = {wy[0], wy[1], wy[2], wy[3]};
  const float* plane = low + (static_cast<long long>(n) * C + c) * gh * gw;
  int cx[4];
#pragma unroll
  for (int m = 0; m < 4; ++m) cx[m] = clampi(su + m, 0, gw - 1);
  float t1[4];
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    const float* row = plane + static_cast<long long>(clampi(sv + l, 0, gh - 1)) * gw;
    t1[l] = dot4_lr(__ldg(row + cx[0]), __ldg(row + cx[1]), __ldg(row + cx[2]), __ldg(row + cx[3]), wxr);
  }
  mid[((static_cast<long long>(n) * parts + c) * hc + v) * wc + u] = dot4_rl(t1[0], t1[1], t1[2], t1[3], wyr);
}

struct MidSet {
  const float* mid[kMaxScales];  // [N][parts][hc][wc] per scale
};

__global__ void __launch_bounds__(256)
resize_accumulate_kernel(const ScaleSet ss, const MidSet ms, int N, int H, int W, int parts, int q1,
                         double* __restrict__ out) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int chunks = (parts + kChunk - 1) / kChunk;
  const int n = blockIdx.z / chunks;
  const int c0 = (blockIdx.z % chunks) * kChunk;
  if (x >= W || y >= H) return;
  double acc[kChunk];
#pragma unroll
  for (int i = 0; i < kChunk; ++i) acc[i] = 0.0;
  const int C = ss.channels;
  const long long tail_start = (static_cast<long long>(W) * C) / 4 * 4;
  const float fS = static_cast<float>(ss.count);
  for (int s = 0; s < ss.count; ++s) {
    const ScaleGeom& g = ss.g[s];
    int sx, sy;
    float fx, fy, wx[4], wy[4];
    cubic_src(x, g.sx, sx, fx);
    cubic_coeffs(fx, wx);
    cubic_src(y, g.sy, sy, fy);
    cubic_coeffs(fy, wy);
    int xi[4], yi[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      xi[k] = clampi(sx - 1 + k, 0, g.wc - 1);
      yi[k] = clampi(sy - 1 + k, 0, g.hc - 1);
    }
    const long long plane = static_cast<long long>(g.hc) * g.wc;
#pragma unroll
    for (int i = 0; i < kChunk; ++i) {
      const int c = c0 + i;
      if (c < parts) {
        const float* img = ms.mid[s] + (static_cast<long long>(n) * parts + c) * plane;
        float t2[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float* row = img + static_cast<long long>(yi[j]) * g.wc;
          t2[j] = dot4_lr(__ldg(row + xi[0]), __ldg(row + xi[1]), __ldg(row + xi[2]), __ldg(row + xi[3]), wx);
        }
        const bool tail = static_cast<long long>(x) * C + c >= tail_start;
        const float v = tail ? dot4_lr(t2[0], t2[1], t2[2], t2[3], wy) : dot4_rl(t2[0], t2[1], t2[2], t2[3], wy);
        const double t = static_cast<double>(__fdiv_rn(v, fS));
        acc[i] = q1 ? __dadd_rn(acc[i], __dadd_rn(acc[i], t)) : __dadd_rn(acc[i], t);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kChunk; ++i) {
    const int c = c0 + i;
    if (c < parts) out[((static_cast<long long>(n) * parts + c) * H + y) * W + x] = acc[i];
  }
}

// ------------------------------------------------------------------------------------------------ gaussian
__device__ __forceinline__ int reflect_index(int e, int n) {
  // scipy mode='reflect' (d c b a | a b c d | d c b a), valid for any distance
  const int period = 2 * n;
  int m = e % period;
  if (m < 0) m += period;
  return m >= n ? period - 1 - m : m;
}

constexpr int kGT = 32;              // output tile edge
constexpr int kGS = kGT + 2;         // smoothed tile edge (1-pixel ring for the NMS neighbours)
constexpr int kGR = 12;              // filter radius: int(4.0 * 3 + 0.5)
constexpr int kGI = kGS + 2 * kGR;   // input tile edge

// NI_Correlate1D, symmetric branch: centre first, then the pairs from the outermost inwards.
#define ISL_GAUSS_1D(CENTER, PAIR)                                                    \
  double tmp = __dmul_rn((CENTER), gw.w[kGR]);                                        \
  _Pragma("unroll") for (int jj = -kGR; jj < 0; ++jj) {                               \
    tmp = __dadd_rn(tmp, __dmul_rn(PAIR, gw.w[kGR + jj]));                            \
  }

template <bool kNms>
__global__ void __launch_bounds__(256)
gauss_kernel(const double* __restrict__ heat, int planes, int H, int W, const GaussWeights gw, double thre,
             int cap, int* __restrict__ counts, uint32_t* __restrict__ keys, double* __restrict__ scores,
             double* __restrict__ smoothed) {
  __shared__ double s_in[kGI][kGI + 1];   // later reused for the smoothed tile
  __shared__ double s_v[kGS][kGI + 1];
  const int plane_id = blockIdx.z;  // n * planes + part
  const double* src = heat + static_cast<long long>(plane_id) * H * W;
  const int x0 = blockIdx.x * kGT, y0 = blockIdx.y * kGT;
  for (int i = threadIdx.x; i < kGI * kGI; i += blockDim.x) {
    const int r = i / kGI, c = i - r * kGI;
    s_in[r][c] = src[static_cast<long long>(reflect_index(y0 - 1 - kGR + r, H)) * W + reflect_index(x0 - 1 - kGR + c, W)];
  }
  __syncthreads();
  // axis 0 first (scipy filters the axes in order)
  for (int i = threadIdx.x; i < kGS * kGI; i += blockDim.x) {
    const int r = i / kGI, c = i - r * kGI;
    ISL_GAUSS_1D(s_in[r + kGR][c], __dadd_rn(s_in[r + kGR + jj][c], s_in[r + kGR - jj][c]))
    s_v[r][c] = tmp;
  }
  __syncthreads();
  double (*s_s)[kGI + 1] = s_in;  // smoothed tile [kGS][kGS] overlays the input tile
  for (int i = threadIdx.x; i < kGS * kGS; i += blockDim.x) {
    const int r = i / kGS, c = i - r * kGS;
    const int ys = y0 - 1 + r, xs = x0 - 1 + c;
    double val = 0.0;  // outside the frame the NMS neighbours are zero (body.py:90-97)
    if (ys >= 0 && ys < H && xs >= 0 && xs < W) {
      ISL_GAUSS_1D(s_v[r][c + kGR], __dadd_rn(s_v[r][c + kGR + jj], s_v[r][c + kGR - jj]))
      val = tmp;
    }
    s_s[r][c] = val;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kGT * kGT; i += blockDim.x) {
    const int r = i / kGT, c = i - r * kGT;
    const int y = y0 + r, x = x0 + c;
    if (y >= H || x >= W) continue;
    const double v = s_s[r + 1][c + 1];
    if (kNms) {
      if (v >= s_s[r][c + 1] && v >= s_s[r + 2][c + 1] && v >= s_s[r + 1][c] && v >= s_s[r + 1][c + 2] && v > thre) {
        const int slot = atomicAdd(counts + plane_id, 1);
        if (slot < cap) {
          keys[static_cast<long long>(plane_id) * cap + slot] = static_cast<uint32_t>(y) * W + x;
          scores[static_cast<long long>(plane_id) * cap + slot] = src[static_cast<long long>(y) * W + x];
        }
      }
    } else {
      smoothed[static_cast<long long>(plane_id) * H * W + static_cast<long long>(y) * W + x] = v;
    }
  }
}

// Sliding-window variant of the same filter (bit-identical: the very same __dmul_rn/__dadd_rn sequence per output).
// gauss_kernel above reads 25 shared-memory doubles per output and per axis, which makes it shared-memory-bandwidth
// bound; here a thread produces kR1 (axis 0) / kR2 (axis 1) consecutive outputs from one register window of
// kR + 24 inputs, i.e. 2.5 / 3.2 loads per output, which leaves the FP64 pipe as the limiter.
//   pass 1 (axis 0): lanes = adjacent columns, inputs straight from global memory (coalesced, reflect per index),
//                    results into s_v[32][90]
//   pass 2 (axis 1): lanes = rows (odd pitch: conflict-free), window from s_v, results into s_s[32][66]
//   pass 3: NMS / store of the 30 x 64 interior
constexpr int kG2W = 64;                 // output tile width
constexpr int kG2H = 30;                 // output tile height
constexpr int kG2SW = kG2W + 2;          // smoothed tile (1-pixel ring)
constexpr int kG2SH = kG2H + 2;          // 32
constexpr int kG2IW = kG2SW + 2 * kGR;   // 90 columns enter pass 1
constexpr int kR1 = 16;                  // axis-0 outputs per thread (2 chunks cover 32 rows)
constexpr int kR2 = 11;                  // axis-1 outputs per thread (6 chunks cover 66 columns)
constexpr int kG2Threads = 192;
static_assert(kG2SH % kR1 == 0 && kG2SW % kR2 == 0, "tile / chunk mismatch");
static_assert((kG2SH / kR1) * kG2IW <= kG2Threads && (kG2SW / kR2) * kG2SH <= kG2Threads, "one work item per thread");

template <bool kNms>
__global__ void __launch_bounds__(kG2Threads)
gauss_window_kernel(const double* __restrict__ heat, int H, int W, const GaussWeights gw, double thre, int cap,
                    int* __restrict__ counts, uint32_t* __restrict__ keys, double* __restrict__ scores,
                    double* __restrict__ smoothed) {
  __shared__ double s_v[kG2SH][kG2IW + 1];   // pitch 91 doubles (odd)
  __shared__ double s_s[kG2SH][kG2SW + 1];   // pitch 67 doubles (odd)
  const int plane_id = blockIdx.z;
  const double* src = heat + static_cast<long long>(plane_id) * H * W;
  const int x0 = blockIdx.x * kG2W, y0 = blockIdx.y * kG2H;
  {
    const int item = threadIdx.x;
    if (item < (kG2SH / kR1) * kG2IW) {
      const int chunk = item / kG2IW, c = item - chunk * kG2IW;
      const double* col = src + reflect_index(x0 - 1 - kGR + c, W);
      double win[kR1 + 2 * kGR];
      const int ybase = y0 - 1 - kGR + chunk * kR1;
#pragma unroll
      for (int k = 0; k < kR1 + 2 * kGR; ++k) win[k] = __ldg(col + static_cast<long long>(reflect_index(ybase + k, H)) * W);
#pragma unroll
      for (int o = 0; o < kR1; ++o) {
        double tmp = __dmul_rn(win[o + kGR], gw.w[kGR]);
#pragma unroll
        for (int jj = -kGR; jj < 0; ++jj)
          tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn(win[o + kGR + jj], win[o + kGR - jj]), gw.w[kGR + jj]));
        s_v[chunk * kR1 + o][c] = tmp;
      }
    }
  }
  __syncthreads();
  {
    const int item = threadIdx.x;
    if (item < (kG2SW / kR2) * kG2SH) {
      const int chunk = item / kG2SH, r = item - chunk * kG2SH;
      const int c0 = chunk * kR2;
      double win[kR2 + 2 * kGR];
#pragma unroll
      for (int k = 0; k < kR2 + 2 * kGR; ++k) win[k] = s_v[r][c0 + k];
      const int ys = y0 - 1 + r;
      const bool row_in = ys >= 0 && ys < H;
#pragma unroll
      for (int o = 0; o < kR2; ++o) {
        double tmp = __dmul_rn(win[o + kGR], gw.w[kGR]);
#pragma unroll
        for (int jj = -kGR; jj < 0; ++jj)
          tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn(win[o + kGR + jj], win[o + kGR - jj]), gw.w[kGR + jj]));
        const int xs = x0 - 1 + c0 + o;
        // outside the frame the NMS neighbours are zero (body.py:90-97)
        s_s[r][c0 + o] = (row_in && xs >= 0 && xs < W) ? tmp : 0.0;
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kG2H * kG2W; i += kG2Threads) {
    const int r = i / kG2W, c = i - r * kG2W;
    const int y = y0 + r, x = x0 + c;
    if (y >= H || x >= W) continue;
    const double v = s_s[r + 1][c + 1];
    if (kNms) {
      if (v >= s_s[r][c + 1] && v >= s_s[r + 2][c + 1] && v >= s_s[r + 1][c] && v >= s_s[r + 1][c + 2] && v > thre) {
        const int slot = atomicAdd(counts + plane_id, 1);
        if (slot < cap) {
          keys[static_cast<long long>(plane_id) * cap + slot] = static_cast<uint32_t>(y) * W + x;
          scores[static_cast<long long>(plane_id) * cap + slot] = src[static_cast<long long>(y) * W + x];
        }
      }
    } else {
      smoothed[static_cast<long long>(plane_id) * H * W + static_cast<long long>(y) * W + x] = v;
    }
  }
}

// One CTA per (frame, part): bitonic sort of the appended peaks by y*W+x = np.nonzero order.
constexpr int kSortCap = 1024;
__global__ void __launch_bounds__(512)
sort_peaks_kernel(int cap, int* __restrict__ counts, uint32_t* __restrict__ keys, double* __restrict__ scores,
                  int* __restrict__ overflow) {
  __shared__ uint32_t s_k[kSortCap];
  __shared__ double s_s[kSortCap];
  const int id = blockIdx.x;
  int n = counts[id];
  if (n > cap) {
    if (threadIdx.x == 0) {
      atomicExch(overflow, 1);
      counts[id] = cap;
    }
    n = cap;
  }
  if (n <= 1) return;
  uint32_t* k = keys + static_cast<long long>(id) * cap;
  double* sc = scores + static_cast<long long>(id) * cap;
  int m = 2;
  while (m < n) m <<= 1;
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    s_k[i] = i < n ? k[i] : 0xffffffffu;
    s_s[i] = i < n ? sc[i] : 0.0;
  }
  __syncthreads();
  for (int size = 2; size <= m; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const int j = i ^ stride;
        if (j > i) {
          const bool up = (i & size) == 0;
          const uint32_t a = s_k[i], b = s_k[j];
          if ((a > b) == up) {
            s_k[i] = b;
            s_k[j] = a;
            const double t = s_s[i];
            s_s[i] = s_s[j];
            s_s[j] = t;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    k[i] = s_k[i];
    sc[i] = s_s[i];
  }
}

// ------------------------------------------------------------------------------------------------ launchers
#define ISL_LAUNCH_OK() (cudaGetLastError() == cudaSuccess ? 0 : 1)

// ISLPOSE_GAUSS=1 selects the first-generation tile kernel (kept for A/B measurements; results are identical)
static int gauss_variant() {
  static const int v = [] {
    const char* e = getenv("ISLPOSE_GAUSS");
    return (e != nullptr && e[0] == '1') ? 1 : 2;
  }();
  return v;
}

int launch_resize_pad_norm(const uint8_t* frames, int N, int H, int W, double scale, int rh, int rw, int hp, int wp,
                           float* out_nchw, uint8_t* out_u8, cudaStream_t st) {
  const dim3 block(32, 8);
  const dim3 grid((wp + 31) / 32, (hp + 7) / 8, N);
  const double inv = 1.0 / scale;  // resize(): scale_x = 1. / inv_scale_x with inv_scale_x = fx
  resize_pad_norm_kernel<<<grid, block, 0, st>>>(frames, N, H, W, rh, rw, hp, wp, inv, inv, out_nchw, out_u8);
  return ISL_LAUNCH_OK();
}

int launch_im2col3x3(const float* in, int N, int h, int w, void* out, cudaStream_t st) {
  const dim3 block(32, 8);
  const dim3 grid((w + 31) / 32, (h + 7) / 8, N);
  im2col3x3_kernel<<<grid, block, 0, st>>>(in, N, h, w, static_cast<__nv_bfloat16*>(out));
  return ISL_LAUNCH_OK();
}

int launch_maxpool2x2(const void* in, int N, int H, int W, int C, void* out, cudaStream_t st) {
  const long long total = static_cast<long long>(N) * (H / 2) * (W / 2) * (C / 8);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  maxpool2x2_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(in), N, H, W, C,
                                                                   static_cast<__nv_bfloat16*>(out));
  return ISL_LAUNCH_OK();
}

long long heat_accumulate_workspace_floats(const ScaleSet& ss, int N, int parts) {
  long long total = 0;
  for (int s = 0; s < ss.count; ++s) total += static_cast<long long>(N) * parts * ss.g[s].hc * ss.g[s].wc;
  return total;
}

int launch_heat_accumulate(const ScaleSet& ss, int N, int H, int W, int parts, int q1, double* out, float* workspace,
                           long long workspace_floats, cudaStream_t st) {
  const int chunks = (parts + kChunk - 1) / kChunk;
  const dim3 grid((W + 31) / 32, (H + 7) / 8, N * chunks);
  if (workspace == nullptr || workspace_floats < heat_accumulate_workspace_floats(ss, N, parts)) {
    heat_accumulate_kernel<<<grid, 256, 0, st>>>(ss, N, H, W, parts, q1, out);  // single pass, no scratch needed
    return ISL_LAUNCH_OK();
  }
  MidSet ms;
  float* cursor = workspace;
  for (int s = 0; s < ss.count; ++s) {
    const ScaleGeom& g = ss.g[s];
    ms.mid[s] = cursor;
    const dim3 g1((g.wc + 31) / 32, (g.hc + 7) / 8, N * parts);
    upsample8_kernel<<<g1, 256, 0, st>>>(g.low, ss.channels, parts, g.gh, g.gw, g.hc, g.wc, cursor);
    cursor += static_cast<long long>(N) * parts * g.hc * g.wc;
  }
  resize_accumulate_kernel<<<grid, 256, 0, st>>>(ss, ms, N, H, W, parts, q1, out);
  return ISL_LAUNCH_OK();
}

int launch_gauss_nms(const double* heat, int planes_total, int H, int W, const GaussWeights& gw, double thre, int cap,
                     int* counts, uint32_t* keys, double* scores, int* overflow, cudaStream_t st) {
  if (cap > kSortCap) return 1;
  if (cudaMemsetAsync(counts, 0, sizeof(int) * planes_total, st) != cudaSuccess) return 1;
  if (gauss_variant() == 1) {
    const dim3 grid((W + kGT - 1) / kGT, (H + kGT - 1) / kGT, planes_total);
    gauss_kernel<true><<<grid, 256, 0, st>>>(heat, 0, H, W, gw, thre, cap, counts, keys, scores, nullptr);
  } else {
    const dim3 grid((W + kG2W - 1) / kG2W, (H + kG2H - 1) / kG2H, planes_total);
    gauss_window_kernel<true><<<grid, kG2Threads, 0, st>>>(heat, H, W, gw, thre, cap, counts, keys, scores, nullptr);
  }
  sort_peaks_kernel<<<planes_total, 512, 0, st>>>(cap, counts, keys, scores, overflow);
  return ISL_LAUNCH_OK();
}

int launch_gauss_smooth(const double* heat, int planes_total, int H, int W, const GaussWeights& gw, double* smoothed,
                        cudaStream_t st) {
  if (gauss_variant() == 1) {
    const dim3 grid((W + kGT - 1) / kGT, (H + kGT - 1) / kGT, planes_total);
    gauss_kernel<false><<<grid, 256, 0, st>>>(heat, 0, H, W, gw, 0.0, 0, nullptr, nullptr, nullptr, smoothed);
  } else {
    const dim3 grid((W + kG2W - 1) / kG2W, (H + kG2H - 1) / kG2H, planes_total);
    gauss_window_kernel<false><<<grid, kG2Threads, 0, st>>>(heat, H, W, gw, 0.0, 0, nullptr, nullptr, nullptr, smoothed);
  }
  return ISL_LAUNCH_OK();
}

}  // namespace islpose
