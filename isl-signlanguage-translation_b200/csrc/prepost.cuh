// Launchers of the pre/post-processing kernels (prepost.cu, group.cu, hand.cu). Host-callable, no torch types.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "cubic.cuh"
#include "gauss.cuh"

namespace islpose {

// limbSeq / mapIdx of the body model (body.py:109-126)
struct LimbTable {
  int nlimbs;
  int njoint;  // 19 (coco) or 26 (body25); subset rows have njoint+1 columns
  int a[24], b[24];    // joint indices of each limb
  int cx[24], cy[24];  // PAF channels of each limb
};

int launch_resize_pad_norm(const uint8_t* frames, int N, int H, int W, double scale, int rh, int rw, int hp, int wp,
                           float* out_nchw, uint8_t* out_u8, cudaStream_t st);
// conv_first.cu: conv1_1 (3 -> 64, 3x3) straight from the float32 NCHW network input; weights bf16 [64][32] with K index
// (ky*3+kx)*3+c (27 used), out bf16 NHWC with out_cstride channels per pixel
// relu: every slope is 0 (the caller has looked): the epilogue is then one convert-with-ReLU per two channels
int launch_conv_first(const float* in, int N, int h, int w, const void* weights, const float* bias, const float* slope,
                      void* out, int out_cstride, bool relu, cudaStream_t st);
long long heat_accumulate_workspace_floats(const ScaleSet& ss, int N, int parts);
int launch_heat_accumulate(const ScaleSet& ss, int N, int H, int W, int parts, int q1, double* out, float* workspace,
                           long long workspace_floats, cudaStream_t st);
int launch_gauss_nms(const double* heat, int planes_total, int H, int W, const GaussWeights& gw, double thre, int cap,
                     int* counts, uint32_t* keys, double* scores, int* overflow, cudaStream_t st);

// pack.cu: float32 [cout][cin][k][k] -> bf16 [k*k][cout][w_cin] in buffer channel order (conv1_1: [1][cout][32])
int launch_pack_conv_weights(const float* w, int cout, int cin, int ksize, const int* chan_map, int in_c, int w_cin, int first,
                             void* out, cudaStream_t st);

constexpr int kMaxPeakCap = 4096;  // peaks per (frame, part) the sort / matching kernels can hold; powers of two above 1024

// Capacity flags OR-ed into the `overflow` word (one bit each, so that one cannot mask another)
enum { kOverflowPeaks = 1, kOverflowCandidates = 2, kOverflowPairs = 4, kOverflowPersons = 8 };

// group.cu
struct GroupBuffers {
  // inputs: sorted peak lists per (frame, part)
  int cap;               // peak capacity per part (<= kMaxPeakCap)
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

// features.cu: the classifier's 156-number row per frame, float64 [n][156]
int launch_body_features(const double* candidate, const double* subset, const int* n_person, int n, int max_cand, int max_person,
                         int njoint, double* out, cudaStream_t st);
int launch_hand_features(const int* table, const int* xy, int n_hands, int n_frames, double* out, cudaStream_t st);

// translate.cu: the sign classifier (demo_isl_translate.py:72-99) on n windows of T x 156 float64 feature rows; weights = the
// Keras model's get_weights() arrays concatenated, float32; probs float32 [n][classes]
long long translate_weight_floats(int classes);
int launch_translate(const double* windows, int n, int T, const float* weights, long long n_weights, int classes, float* probs,
                     cudaStream_t st);

// hand.cu: key points of up to kHandMaxCrops crops (of any sizes) per launch chain
constexpr int kHandMaxCrops = 32;
constexpr int kHandMaxScales = 4;  // hand.py:25 fixes the list at four scales
struct HandCropDev {
  int H, W;                          // crop size
  const float* low[kHandMaxScales];  // this crop's network output per scale: fp32 [channels][gh][gw]
  int gh[kHandMaxScales], gw[kHandMaxScales];  // stride-8 grid
  int hc[kHandMaxScales], wc[kHandMaxScales];  // up-sampled extent after cropping the pad
  double* heat;                      // [parts][H][W] float64 mean over the scales (hand.py:56)
  int* labels;                       // [parts][H][W] scratch
  double* mass;                      // [parts][H][W] scratch
  int32_t* out_xy;                   // [parts][2]
};
struct HandBatch {
  int n_crops;
  int n_scales;
  int channels;  // 22: the row length OpenCV sees is W * channels
  int parts;     // 21 planes are evaluated (hand.py:58)
  HandCropDev crop[kHandMaxCrops];
};
// compute_heat = false: crop[i].heat already holds the float64 maps (islpose_hand_peaks)
int launch_hand_keypoints(const HandBatch& hb, bool compute_heat, const GaussWeights& gw, double thre, cudaStream_t st);

}  // namespace islpose
