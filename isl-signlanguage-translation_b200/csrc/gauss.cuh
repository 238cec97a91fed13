// scipy.ndimage.gaussian_filter(sigma=3) on float64 planes (body.py:88, hand.py:61): 25 taps, mode='reflect', axis 0 then
// axis 1, NI_Correlate1D's symmetric summation order (centre first, then the pairs from the outermost inwards), every
// operation an explicit __dmul_rn / __dadd_rn so that the result is bit-identical to scipy's.
//
// One CTA (kG2Threads threads) produces a 30 x 64 tile from register sliding windows:
//   pass 1 (axis 0): lanes = adjacent columns, inputs straight from global memory (coalesced, reflect per index); a
//                    thread produces kR1 consecutive outputs of its column from one window of kR1 + 24 inputs
//   pass 2 (axis 1): lanes = rows (odd pitch: conflict-free), window from shared memory, kR2 outputs per thread
//   pass 3: what the caller does with the smoothed 30 x 64 interior (and its 1-pixel ring), selected by kMode:
//           kGaussPeaks  4-neighbour NMS against zero-filled borders + threshold -> appended peak list (body.py:90-107)
//           kGaussLabels smoothed > thre -> initial component labels p+1 / 0 (hand.py:62, first step of the labelling)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace islpose {

struct GaussWeights {
  double w[25];  // scipy _gaussian_kernel1d(sigma=3, radius=12), computed by the host in float64
};

__device__ __forceinline__ int reflect_index(int e, int n) {
  // scipy mode='reflect' (d c b a | a b c d | d c b a), valid for any distance
  const int period = 2 * n;
  int m = e % period;
  if (m < 0) m += period;
  return m >= n ? period - 1 - m : m;
}

constexpr int kGR = 12;                  // filter radius: int(4.0 * 3 + 0.5)
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

enum { kGaussPeaks = 0, kGaussLabels = 1 };

struct GaussSmem {
  double v[kG2SH][kG2IW + 1];   // pitch 91 doubles (odd)
  double s[kG2SH][kG2SW + 1];   // pitch 67 doubles (odd)
};

// src: one H x W float64 plane. Tile origin (x0, y0). kGaussPeaks: counts/keys/scores of this plane (cap entries);
// kGaussLabels: labels of this plane.
template <int kMode>
__device__ __forceinline__ void gauss_window_tile(GaussSmem& sm, const double* __restrict__ src, int H, int W, int x0, int y0,
                                                  const GaussWeights& gw, double thre, int cap, int* __restrict__ count,
                                                  uint32_t* __restrict__ keys, double* __restrict__ scores,
                                                  int* __restrict__ labels) {
  {
    const int item = threadIdx.x;
    if (item < (kG2SH / kR1) * kG2IW) {
      const int chunk = item / kG2IW, c = item - chunk * kG2IW;
      const int xg = x0 - 1 - kGR + c;
      const double* col = src + ((xg >= 0 && xg < W) ? xg : reflect_index(xg, W));
      double win[kR1 + 2 * kGR];
      const int ybase = y0 - 1 - kGR + chunk * kR1;
      if (ybase >= 0 && ybase + kR1 + 2 * kGR <= H) {  // interior: no reflection, no index arithmetic per load
        const double* p = col + static_cast<long long>(ybase) * W;
#pragma unroll
        for (int k = 0; k < kR1 + 2 * kGR; ++k) win[k] = __ldg(p + static_cast<long long>(k) * W);
      } else {
#pragma unroll
        for (int k = 0; k < kR1 + 2 * kGR; ++k) win[k] = __ldg(col + static_cast<long long>(reflect_index(ybase + k, H)) * W);
      }
#pragma unroll
      for (int o = 0; o < kR1; ++o) {
        double tmp = __dmul_rn(win[o + kGR], gw.w[kGR]);
#pragma unroll
        for (int jj = -kGR; jj < 0; ++jj)
          tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn(win[o + kGR + jj], win[o + kGR - jj]), gw.w[kGR + jj]));
        sm.v[chunk * kR1 + o][c] = tmp;
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
      for (int k = 0; k < kR2 + 2 * kGR; ++k) win[k] = sm.v[r][c0 + k];
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
        sm.s[r][c0 + o] = (row_in && xs >= 0 && xs < W) ? tmp : 0.0;
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kG2H * kG2W; i += kG2Threads) {
    const int r = i / kG2W, c = i - r * kG2W;
    const int y = y0 + r, x = x0 + c;
    if (y >= H || x >= W) continue;
    const double v = sm.s[r + 1][c + 1];
    if (kMode == kGaussPeaks) {
      if (v >= sm.s[r][c + 1] && v >= sm.s[r + 2][c + 1] && v >= sm.s[r + 1][c] && v >= sm.s[r + 1][c + 2] && v > thre) {
        const int slot = atomicAdd(count, 1);
        if (slot < cap) {
          keys[slot] = static_cast<uint32_t>(y) * W + x;
          scores[slot] = src[static_cast<long long>(y) * W + x];
        }
      }
    } else {
      const int p = y * W + x;
      labels[p] = v > thre ? p + 1 : 0;
    }
  }
}

}  // namespace islpose
