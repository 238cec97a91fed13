// OpenCV INTER_CUBIC arithmetic, restated for the device with the exact operation order of
// modules/imgproc/src/resize.cpp (generic path): Keys kernel with A = -0.75 evaluated in float32,
// src = (dst + 0.5) * scale - 0.5 evaluated in double and rounded to float, replicate border per tap,
// horizontal pass ((s0*a0 + s1*a1) + s2*a2) + s3*a3, vertical pass in the 4-lane SIMD order
// ((s3*b3 + s2*b2) + s1*b1) + s0*b0 (scalar order for the row tail), no fused multiply-add.
// Every float op goes through __f*_rn so nvcc cannot contract it.
//
// The reference chains two such resizes per network output (body.py:70-72,76-78; hand.py:52-54):
//   stage 1: x8 up-sampling of the stride-8 map [gh,gw] -> [8gh,8gw], cropped to [hc,wc]
//   stage 2: [hc,wc] -> the frame size [H,W]
// Axis2 holds, for one output coordinate along one axis, both stages' taps and weights; sample2() then
// evaluates one output value from a 5x5 window of the stride-8 map (164 multiply-adds), which is what both
// the map accumulation kernel and the lazy PAF sampler use.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace islpose {

__device__ __forceinline__ void cubic_coeffs(float x, float (&c)[4]) {
  const float A = -0.75f;
  const float x1 = __fadd_rn(x, 1.f);
  float t = __fmul_rn(A, x1);
  t = __fsub_rn(t, 5.f * A);
  t = __fmul_rn(t, x1);
  t = __fadd_rn(t, 8.f * A);
  t = __fmul_rn(t, x1);
  c[0] = __fsub_rn(t, 4.f * A);
  t = __fmul_rn(A + 2.f, x);
  t = __fsub_rn(t, A + 3.f);
  t = __fmul_rn(t, x);
  t = __fmul_rn(t, x);
  c[1] = __fadd_rn(t, 1.f);
  const float xm = __fsub_rn(1.f, x);
  t = __fmul_rn(A + 2.f, xm);
  t = __fsub_rn(t, A + 3.f);
  t = __fmul_rn(t, xm);
  t = __fmul_rn(t, xm);
  c[2] = __fadd_rn(t, 1.f);
  c[3] = __fsub_rn(__fsub_rn(__fsub_rn(1.f, c[0]), c[1]), c[2]);
}

// Source position of destination index d: integer part and float fraction, OpenCV style.
__device__ __forceinline__ void cubic_src(int d, double scale, int& s, float& frac) {
  const float f = static_cast<float>(__dsub_rn(__dmul_rn(__dadd_rn(static_cast<double>(d), 0.5), scale), 0.5));
  const float fl = floorf(f);
  s = static_cast<int>(fl);
  frac = __fsub_rn(f, fl);
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__device__ __forceinline__ float dot4_lr(float s0, float s1, float s2, float s3, const float (&w)[4]) {
  return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(s0, w[0]), __fmul_rn(s1, w[1])), __fmul_rn(s2, w[2])),
                   __fmul_rn(s3, w[3]));
}
__device__ __forceinline__ float dot4_rl(float s0, float s1, float s2, float s3, const float (&w)[4]) {
  return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(s3, w[3]), __fmul_rn(s2, w[2])), __fmul_rn(s1, w[1])),
                   __fmul_rn(s0, w[0]));
}

struct Axis2 {
  float w2[4];     // stage-2 weights of this output coordinate
  float w1[4][4];  // stage-1 weights of each of the 4 stage-2 taps
  int off[4];      // 0/1: where each tap's 4-wide stride-8 window starts inside the 5-wide window
  int lo[5];       // clamped stride-8 indices of the 5-wide window
};

// Stage 1 is an exact x8 up-sampling (scale 0.125), so the source fraction of up-sampled index u depends on
// u mod 8 only and the integer part is ((u + 4) >> 3) - 1. The 8 weight sets are computed once per block with the
// very same float code (bit-identical to evaluating them per tap) and kept in shared memory.
__device__ __forceinline__ void fill_phase_table(float (*tab)[4]) {
  if (threadIdx.x < 8) {
    int su;
    float g;
    cubic_src(static_cast<int>(threadIdx.x), 0.125, su, g);
    float c[4];
    cubic_coeffs(g, c);
    tab[threadIdx.x][0] = c[0];
    tab[threadIdx.x][1] = c[1];
    tab[threadIdx.x][2] = c[2];
    tab[threadIdx.x][3] = c[3];
  }
}

// d: output coordinate; scale2: stage-2 src/dst ratio; n_mid: cropped up-sampled extent; n_low: stride-8 extent.
__device__ __forceinline__ void make_axis2(int d, double scale2, int n_mid, int n_low, const float (*tab)[4], Axis2& a) {
  int s;
  float frac;
  cubic_src(d, scale2, s, frac);
  cubic_coeffs(frac, a.w2);
  int first = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int u = clampi(s - 1 + i, 0, n_mid - 1);
    const int su = ((u + 4) >> 3) - 1;
    const float* w = tab[u & 7];
    a.w1[i][0] = w[0];
    a.w1[i][1] = w[1];
    a.w1[i][2] = w[2];
    a.w1[i][3] = w[3];
    if (i == 0) first = su - 1;
    a.off[i] = su - 1 - first;  // 0 or 1: the four taps span at most 3 up-sampled pixels = < 1 cell
  }
#pragma unroll
  for (int t = 0; t < 5; ++t) a.lo[t] = clampi(first + t, 0, n_low - 1);
}

// One output value from the stride-8 plane `low` ([gh][gw] floats). scalar_tail selects the scalar
// summation order OpenCV uses for the last (W*C mod 4) floats of a stage-2 row.
__device__ __forceinline__ float sample2(const float* __restrict__ low, int gw, const Axis2& ax, const Axis2& ay,
                                         bool scalar_tail) {
  float L[5][5];
#pragma unroll
  for (int t = 0; t < 5; ++t) {
    const float* row = low + static_cast<long long>(ay.lo[t]) * gw;
#pragma unroll
    for (int q = 0; q < 5; ++q) L[t][q] = __ldg(row + ax.lo[q]);
  }
  // stage 1, horizontal: T1[t][i] for the 5 stride-8 rows and the 4 stage-2 column taps
  float T1[5][4];
#pragma unroll
  for (int t = 0; t < 5; ++t) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const bool o = ax.off[i] != 0;
      T1[t][i] = dot4_lr(o ? L[t][1] : L[t][0], o ? L[t][2] : L[t][1], o ? L[t][3] : L[t][2],
                         o ? L[t][4] : L[t][3], ax.w1[i]);
    }
  }
  // stage 1, vertical (row length 8*gw*C is a multiple of 4: SIMD order everywhere), then stage 2 horizontal
  float T2[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const bool o = ay.off[j] != 0;
    float I[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      I[i] = dot4_rl(o ? T1[1][i] : T1[0][i], o ? T1[2][i] : T1[1][i], o ? T1[3][i] : T1[2][i],
                     o ? T1[4][i] : T1[3][i], ay.w1[j]);
    }
    T2[j] = dot4_lr(I[0], I[1], I[2], I[3], ax.w2);
  }
  return scalar_tail ? dot4_lr(T2[0], T2[1], T2[2], T2[3], ay.w2) : dot4_rl(T2[0], T2[1], T2[2], T2[3], ay.w2);
}

// Geometry of one scale of one call (frames of a batch share it).
struct ScaleGeom {
  const float* low;  // planar NCHW fp32 [N][C][gh][gw]
  int gh, gw;        // stride-8 grid
  int hc, wc;        // up-sampled extent after cropping the pad (= resized image size before padding)
  double sx, sy;     // stage-2 src/dst ratios: 1.0 / ((double)W / wc), 1.0 / ((double)H / hc)
};

constexpr int kMaxScales = 8;
struct ScaleSet {
  ScaleGeom g[kMaxScales];
  int count;
  int channels;  // C of the network output (the row length OpenCV sees is W*C)
};

}  // namespace islpose
