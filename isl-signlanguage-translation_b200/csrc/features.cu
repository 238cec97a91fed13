// Feature-vector formation on the device: the 156 numbers per frame the ISL classifier consumes,
// written straight from the grouping and hand key-point outputs so that they leave the GPU as one small row per frame.
//   body_features   util.get_bodypose's circles (src/util.py:123-130: joints in joint-major, person-minor order, absent
//                   joints skipped) -> the first 15 as x[15], y[15] (populate_features, src/ISL_Model_parameter.py:376-390);
//                   also zero-fills the rest of the row
//   hand_features   util.get_handpose's key points (src/util.py:213-217) of the frame's first two hands, shifted from crop
//                   to frame coordinates where non-zero (demo.py:36-37), as x[21], y[21], index[21] per hand
//                   (ISL_Model_parameter.py:392-409)
// Row layout: [0,15) body x, [15,30) body y, then per hand h: [30+63h, +21) x, [+21, +42) y, [+42, +63) key-point index.
#include "prepost.cuh"

namespace islpose {

constexpr int kFeat = 156;
constexpr int kCircles = 15;

__global__ void __launch_bounds__(32)
body_features_kernel(const double* __restrict__ candidate, const double* __restrict__ subset, const int* __restrict__ n_person,
                     int max_cand, int max_person, int njoint, double* __restrict__ out) {
  const int n = blockIdx.x;
  const int lane = threadIdx.x;
  const int parts = njoint - 1, cols = njoint + 1;
  double* row = out + static_cast<long long>(n) * kFeat;
  for (int i = lane; i < kFeat; i += 32) row[i] = 0.0;
  __syncwarp();
  const double* cand = candidate + static_cast<long long>(n) * max_cand * 4;
  const double* sub = subset + static_cast<long long>(n) * max_person * cols;
  const int P = n_person[n];
  int count = 0;
  for (int i = 0; i < parts && count < kCircles; ++i) {
    for (int base = 0; base < P && count < kCircles; base += 32) {
      const int p = base + lane;
      int index = -1;
      if (p < P) index = static_cast<int>(sub[p * cols + i]);   // int(subset[n][i]), util.py:125
      const bool has = index != -1;
      const unsigned mask = __ballot_sync(0xffffffffu, has);
      const int rank = count + __popc(mask & ((1u << lane) - 1u));
      if (has && rank < kCircles) {
        row[rank] = cand[index * 4 + 0];
        row[kCircles + rank] = cand[index * 4 + 1];
      }
      count += __popc(mask);
    }
  }
}

// table: int32 [n_hands][4] = frame, slot (0 / 1 = the frame's first / second hand), crop x, crop y; xy: int32 [n_hands][21][2]
__global__ void __launch_bounds__(32)
hand_features_kernel(const int* __restrict__ table, const int* __restrict__ xy, int n_frames, double* __restrict__ out) {
  const int h = blockIdx.x;
  const int frame = table[h * 4 + 0], slot = table[h * 4 + 1], x0 = table[h * 4 + 2], y0 = table[h * 4 + 3];
  if (frame < 0 || frame >= n_frames || slot < 0 || slot > 1) return;
  const int k = threadIdx.x;
  if (k >= 21) return;
  const int x = xy[(h * 21 + k) * 2 + 0], y = xy[(h * 21 + k) * 2 + 1];
  double* row = out + static_cast<long long>(frame) * kFeat + 2 * kCircles + slot * 63;
  row[k] = static_cast<double>(x == 0 ? x : x + x0);        // 0 means "not found" (demo.py:36-37)
  row[21 + k] = static_cast<double>(y == 0 ? y : y + y0);
  row[42 + k] = static_cast<double>(k);                     // the key point's number (util.py:217 str(i), float() in populate_features)
}

int launch_body_features(const double* candidate, const double* subset, const int* n_person, int n, int max_cand, int max_person,
                         int njoint, double* out, cudaStream_t st) {
  body_features_kernel<<<n, 32, 0, st>>>(candidate, subset, n_person, max_cand, max_person, njoint, out);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int launch_hand_features(const int* table, const int* xy, int n_hands, int n_frames, double* out, cudaStream_t st) {
  if (n_hands <= 0) return 0;
  hand_features_kernel<<<n_hands, 32, 0, st>>>(table, xy, n_frames, out);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace islpose
