// Launchers of the pre/post-processing kernels (prepost.cu, group.cu, hand.cu). Host-callable, no torch types.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "cubic.cuh"

namespace islpose {

struct GaussWeights {
  double w[25];  // scipy _gaussian_kernel1d(sigma=3, radius=12), computed by the host in float64
};

// limbSeq / mapIdx of the body model (body.py:109-126)
struct LimbTable {
  int nlimbs;
  int njoint;  // 19 (coco) or 26 (body25); subset rows have njoint+1 columns
  int a[24], b[24];    // joint indices of each limb
  int cx[24], cy[24];  // PAF channels of each limb
};

int launch_resize_pad_norm(const uint8_t* frames, int N, int H, int W, double scale, int rh, int rw, int hp, int wp,
                           float* out_nchw, uint8_t* out_u8, cudaStream_t st);
int launch_im2col3x3(const float* in, int N, int h, int w, void* out, cudaStream_t st);
// conv_first.cu: conv1_1 (3 -> 64, 3x3) straight from the float32 NCHW network input; weights bf16 [64][32] with K index
// (ky*3+kx)*3+c (27 used), out bf16 NHWC with out_cstride channels per pixel
int launch_conv_first(const float* in, int N, int h, int w, const void* weights, const float* bias, const float* slope,
                      void* out, int out_cstride, cudaStream_t st);
int launch_maxpool2x2(const void* in, int N, int H, int W, int C, void* out, cudaStream_t st);
long long heat_accumulate_workspace_floats(const ScaleSet& ss, int N, int parts);
int launch_heat_accumulate(const ScaleSet& ss, int N, int H, int W, int parts, int q1, double* out, float* workspace,
                           long long workspace_floats, cudaStream_t st);
int launch_gauss_nms(const double* heat, int planes_total, int H, int W, const GaussWeights& gw, double thre, int cap,
                     int* counts, uint32_t* keys, double* scores, int* overflow, cudaStream_t st);
int launch_gauss_smooth(const double* heat, int planes_total, int H, int W, const GaussWeights& gw, double* smoothed,
                        cudaStream_t st);

// group.cu
struct GroupBuffers {
  // inputs: sorted peak lists per (frame, part)
  int cap;               // peak capacity per part (<= 1024)
  const int* counts;     // [N*parts]
  const uint32_t* keys;  // [N*parts*cap]  y*W+x
  const double* scores;  // [N*parts*cap]  unsmoothed heat value
  // scratch: dense connection scores per (frame, limb): [i*nB+j] = score, or -1 when the pair is rejected
  long long pair_cap;    // elements available per (frame, limb); nA*nB above this raises overflow code 3
  double* pair_score;    // [N*nlimbs*pair_cap]
  double* end_paf;       // [N*nlimbs*2*cap*2] scratch: PAF vector of each limb at each of its end peaks
  // scratch: chosen connections per (frame, limb)
  int* conn_count;       // [N*nlimbs]
  int* conn_ij;          // [N*nlimbs*cap*2]
  double* conn_score;    // [N*nlimbs*cap]
  int* owner;            // [N*max_cand*2] scratch: row slots holding each candidate id
  // outputs
  int max_cand;          // rows available per frame in candidate
  double* candidate;     // [N*max_cand*4]  x, y, score, id
  int* n_cand;           // [N]
  int max_person;        // row slots available per frame (rows ever created, dead ones included)
  double* subset;        // [N*max_person*(njoint+1)]
  int* n_person;         // [N]
  int* overflow;         // [1] set to non-zero if any capacity was exceeded
};
int launch_paf_score(const ScaleSet& paf, const LimbTable& lt, int N, int H, int W, double thre2, int mid_num,
                     const GroupBuffers& gb, cudaStream_t st);
int launch_group(const LimbTable& lt, int N, int W, const GroupBuffers& gb, cudaStream_t st);

// hand.cu
int launch_hand_peaks(const double* heat, const double* smoothed, int planes_total, int H, int W, double thre,
                      int* labels, double* mass, int32_t* out_xy, cudaStream_t st);

}  // namespace islpose
