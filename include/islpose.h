/* islpose - C ABI of the B200 (sm_100a) OpenPose keypoint-extraction kernels.
 *
 * The reference (sunilsarolkarcds/ISL-SignLanguage-Translation) is pure Python and has no FFI of its own; its
 * boundary for this path is Body.__call__ (src/body.py:39), Hand.__call__ (src/hand.py:24) and util.handDetect
 * (src/util.py:242). This library is what a Python shim behind those three signatures binds with ctypes
 * (INTEGRATION.md shows the stub). Each entry point below names the reference lines it replaces.
 *
 * Conventions: every function returns 0 on success and non-zero on failure, in which case
 * islpose_last_error() returns a thread-local message. All data pointers are DEVICE pointers unless the
 * name starts with h_. Every launch is asynchronous on the given stream (a cudaStream_t passed as void*).
 * No torch types, no C++ types.
 */
#ifndef ISLPOSE_H_
#define ISLPOSE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ISLPOSE_ABI_VERSION 3
#define ISLPOSE_MAX_SCALES 8

int islpose_abi_version(void);
const char* islpose_last_error(void);
/* Number of kernels this library has launched since it was loaded (all threads); bench.py reports the difference
 * over its timed region as gpu_launches. */
int64_t islpose_launch_count(void);
/* sizeof of the four structs below as this library was compiled: islpose_scale, islpose_conv_desc, islpose_group_buffers,
 * islpose_hand_crop. A binding compares them with its own layout before passing any struct. */
int islpose_struct_sizes(int32_t out[4]);

/* ---------------------------------------------------------------------------------------------------------
 * Network plans: a recorded list of kernel launches (one per layer) over caller-owned device buffers,
 * replayed per forward. Replaces nn.Module.forward of bodypose_model / bodypose_25_model / handpose_model
 * (src/model.py:179-207, 302-329, 394-407). Activations are NHWC bf16; torch.cat disappears because
 * producers write channel slices of shared buffers. */
typedef struct islpose_plan islpose_plan;

typedef struct islpose_conv_desc {
  const void* in;       /* bf16 NHWC, points at the first channel of the input slice at pixel 0 */
  int32_t in_c;         /* channels of the slice (multiple of 8) */
  int32_t in_cstride;   /* channels per pixel of the buffer holding the slice (multiple of 8) */
  int32_t in_c_readable;/* channels readable from the slice start (>= in_c, 0 = in_c); if the next multiple of 64 fits,
                           all TMA boxes stay in bounds (faster); channels past in_c must be finite and meet zero weights */
  int32_t w_cin;        /* Cin stride of the packed weights (>= in_c, multiple of 8; 0 = in_c), zero padded */
  int32_t n, h, w;      /* batch and spatial size (stride 1, "same" padding: output has the same size) */
  const void* weights;  /* bf16 [k*k][cout][w_cin]: nn.Conv2d weight [cout][cin][ky][kx] re-packed tap-major */
  int32_t cout;
  int32_t ksize;        /* 1, 3 or 7 (src/model.py:35-37) */
  const float* bias;    /* fp32, at least 512 entries, zero padded */
  const float* slope;   /* fp32, at least 512 entries: 0 = ReLU, 1 = no activation, else per-channel PReLU */
  void* out_bf16;       /* bf16 NHWC slice (may be NULL): receives cout rounded up to 8 channels, pad = 0 */
  int32_t out_cstride;  /* channels per pixel of that buffer */
  float* out_f32;       /* fp32 planar NCHW (may be NULL): channel 0 of this layer inside [n][out_f32_channels][h][w] */
  int32_t out_f32_channels;
  int32_t n_tile, stages, tile_w, tile_h; /* tuning overrides, 0 = automatic */
  int32_t pool;         /* 1: nn.MaxPool2d(2,2,0) fused behind the activation (src/model.py:30-32); out_bf16 is then the pooled
                           [n][h/2][w/2] buffer. Only for 3x3 / 7x7 layers with >= 64 input and >= 48 output channels */
  int32_t sm_budget;    /* SMs the layer's persistent kernel may occupy, 0 = all. Plans that run side by side on several
                           streams (the scales of one small batch) finish sooner when each keeps to a share of the device
                           than when every layer of every plan asks for all of it */
  /* A second 1x1 layer chained behind this 1x1 layer in the same launch (Mconv6 -> Mconv7, conv5_4 -> conv5_5, conv6_1 ->
   * conv6_2: src/model.py:57-62 builds them as consecutive nn.Conv2d(k=1)). weights2 != NULL selects it: this layer's cout
   * (64..512, a multiple of 64) activations of a pixel tile stay in shared memory, rounded to bf16 as the stored
   * activation would be, and out_bf16 / out_f32 must be NULL; the fields below are the second layer's. */
  const void* weights2; /* bf16 [1][cout2][cout] */
  int32_t cout2;        /* 1..64 */
  const float* bias2;   /* fp32, at least 64 entries, zero padded */
  const float* slope2;
  void* out2_bf16;      /* bf16 NHWC slice (may be NULL): cout2 rounded up to 8 channels, pad = 0 */
  int32_t out2_cstride;
  void* out2b_bf16;     /* the same slice once more in another buffer (may be NULL) */
  int32_t out2b_cstride;
  float* out2_f32;      /* fp32 planar NCHW (may be NULL) */
  int32_t out2_f32_channels;
} islpose_conv_desc;

/* Weight ingestion (src/body.py:35-36, src/util.py:35-44): one nn.Conv2d weight, float32 [cout][cin][k][k] on the device,
 * -> the `weights` operand of islpose_conv_desc: bf16 [k*k][cout][w_cin], input channels in the order of the activation
 * slice the layer reads. chan_map (device int32 [in_c], may be NULL = identity): reference input channel at slice channel i,
 * -1 = pad channel (zero weights). first_layer != 0 packs conv1_1 for islpose_plan_add_first_conv: [1][cout][w_cin = 32],
 * K index (ky*3+kx)*3+c. Round-to-nearest-even. */
int islpose_pack_conv_weights(const float* w, int32_t cout, int32_t cin, int32_t ksize, const int32_t* chan_map, int32_t in_c,
                              int32_t w_cin, int32_t first_layer, void* out_bf16, void* stream);

int islpose_plan_create(islpose_plan** out);
int islpose_plan_destroy(islpose_plan* plan);
int islpose_plan_add_conv(islpose_plan* plan, const islpose_conv_desc* desc);
/* conv1_1 (3 -> 64 channels, 3x3, src/model.py 'conv1_1' of all three networks) in one launch straight from the fp32 NCHW
 * network input: weights bf16 [64][32] with K index (ky*3+kx)*3+c (27 used, 5 zero), bias / slope as in islpose_conv_desc,
 * out bf16 NHWC with out_cstride (>= 64) channels per pixel. relu != 0: the caller states that all 64 slopes are 0 (the
 * layer is followed by nn.ReLU, as conv1_1 is in the coco and hand networks): the epilogue then converts with ReLU in one
 * instruction per two channels instead of reading the slopes. The bias enters the GEMM as two bf16 K columns (hi + lo part,
 * exact to 2^-17 of its value). */
int islpose_plan_add_first_conv(islpose_plan* plan, const float* in_nchw, const void* weights, const float* bias,
                                const float* slope, void* out, int32_t out_cstride, int32_t n, int32_t h, int32_t w,
                                int32_t relu);
/* Replays the plan on `stream`. The first run records the launches into a CUDA graph (programmatic dependent launch edges
 * included); later runs are one graph launch. islpose_plan_set_graph(plan, 0) keeps kernel-by-kernel launches;
 * islpose_plan_graph_state: 0 = not recorded yet, 1 = graph in use, -1 = recording was not possible (direct launches). */
int islpose_plan_run(islpose_plan* plan, void* stream);
int islpose_plan_set_graph(islpose_plan* plan, int32_t enable);
int32_t islpose_plan_graph_state(const islpose_plan* plan);
const char* islpose_plan_graph_note(const islpose_plan* plan); /* why recording was not possible ("" otherwise) */
/* Measurement aid: runs the plan launch by launch, each one `reps` times back to back between two CUDA events (after
 * one untimed run), and blocks until done. h_ms / h_flops / h_variant (HOST arrays of num_launches entries; the last
 * two may be NULL) receive the average milliseconds, the algorithmic FLOPs (0 for non-conv launches) and the conv
 * kernel variant (0 = first-layer kernel) of every launch. */
int islpose_plan_profile(const islpose_plan* plan, void* stream, int32_t reps, float* h_ms, double* h_flops,
                         int32_t* h_variant);
int32_t islpose_plan_num_launches(const islpose_plan* plan);
double islpose_plan_conv_flops(const islpose_plan* plan); /* sum of 2*Cin*Cout*k*k*H*W*N over the slices as given */

/* ---------------------------------------------------------------------------------------------------------
 * Pre-processing: cv2.resize(INTER_CUBIC, fx=fy=scale) + util.padRightDownCorner(8, 128) + x/256-0.5 + HWC->CHW
 * (src/body.py:53-56, src/hand.py:37-40, src/util.py:12-32). frames: uint8 [n,H,W,3] contiguous.
 * rh,rw = round-half-even(H*scale, W*scale); hp,wp = rh,rw rounded up to multiples of 8.
 * out_nchw: fp32 [n,3,hp,wp]; out_u8 (optional, may be NULL): the padded uint8 image [n,hp,wp,3]. */
int islpose_resize_pad_normalize(const uint8_t* frames, int32_t n, int32_t H, int32_t W, double scale, int32_t rh,
                                 int32_t rw, int32_t hp, int32_t wp, float* out_nchw, uint8_t* out_u8, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * One scale of a call: the network output and the geometry of the two cubic resizes that bring it to frame size. */
typedef struct islpose_scale {
  const float* lowres; /* fp32 planar NCHW [n][channels][gh][gw] */
  int32_t gh, gw;      /* stride-8 grid = hp/8, wp/8 */
  int32_t hc, wc;      /* x8 up-sampled map cropped to the un-padded resized image: rh, rw */
} islpose_scale;

/* Map post-processing (src/body.py:69-81, src/hand.py:51-56): x8 cubic, crop, cubic to (W,H), /S, accumulate in
 * float64. Writes channels [0, parts) of the result as planes: out float64 [n][parts][H][W].
 * double_running_sum != 0 reproduces `heatmap_avg += heatmap_avg + heatmap / S` (body.py:80).
 * workspace (optional, may be NULL): float32 scratch of islpose_maps_workspace_floats() elements; with it the
 * up-sampled maps are materialised once per scale (same results, ~8x less arithmetic), without it every output
 * pixel evaluates both resizes from a 5x5 window of the stride-8 map. */
int64_t islpose_maps_workspace_floats(const islpose_scale* scales, int32_t n_scales, int32_t n, int32_t parts);
int islpose_maps_accumulate(const islpose_scale* scales, int32_t n_scales, int32_t channels, int32_t n, int32_t H,
                            int32_t W, int32_t parts, int32_t double_running_sum, double* out, float* workspace,
                            int64_t workspace_floats, void* stream);

/* Peak detection (src/body.py:86-107): scipy gaussian_filter(sigma=3) in float64 with reflect borders, 4-neighbour
 * NMS against zero-filled shifts, threshold; peaks of every (frame, part) plane sorted in row-major order.
 * h_gauss: the 25 float64 filter weights, a HOST pointer. cap <= 1024 peaks per plane, or 2048 / 4096 (the caller grows it
 * when ISLPOSE_OVERFLOW_PEAKS comes back).
 * counts int32 [planes]; keys uint32 [planes][cap] = y*W+x; scores float64 [planes][cap] = unsmoothed value. */
int islpose_body_peaks(const double* heat, int32_t planes, int32_t H, int32_t W, const double* h_gauss, double thre1,
                       int32_t cap, int32_t* counts, uint32_t* keys, double* scores, int32_t* overflow, void* stream);

/* Connection scoring, greedy matching, person assembly and pruning (src/body.py:109-235). PAF values are sampled
 * from the stride-8 PAF maps on demand (both cubic stages and the mean over scales). model_kind: 0 coco, 1 body25.
 * Scratch and outputs are caller-allocated; see islpose_group_buffers. */
typedef struct islpose_group_buffers {
  int32_t cap;             /* peak capacity per part used in islpose_body_peaks */
  const int32_t* counts;
  const uint32_t* keys;
  const double* scores;
  int64_t pair_cap;        /* scratch elements per (frame, limb); needs nA*nB <= pair_cap, else overflow = 3 */
  double* pair_score;      /* [n*nlimbs*pair_cap] dense nA x nB connection scores (-1 = rejected pair) */
  double* end_paf;         /* [n*nlimbs*2*cap*2] scratch: the limb's PAF vector at each of its end peaks */
  int32_t* conn_count;     /* [n*nlimbs] */
  int32_t* conn_ij;        /* [n*nlimbs*cap*2] */
  double* conn_score;      /* [n*nlimbs*cap] */
  int32_t* owner;          /* [n*max_cand*2] scratch */
  int32_t max_cand;
  double* candidate;       /* out [n][max_cand][4]: x, y, score, id */
  int32_t* n_cand;         /* out [n] */
  int32_t max_person;      /* row slots per frame: every row ever created, merged-away ones included (<= 65536) */
  double* subset;          /* out [n][max_person][njoint+1] */
  int32_t* n_person;       /* out [n] */
  int32_t* overflow;       /* out [1]: OR of ISLPOSE_OVERFLOW_* bits, one per capacity that was exceeded (never cleared here) */
} islpose_group_buffers;

#define ISLPOSE_OVERFLOW_PEAKS 1      /* a (frame, part) has more than cap peaks (islpose_body_peaks) */
#define ISLPOSE_OVERFLOW_CANDIDATES 2 /* a frame has more than max_cand peaks in total */
#define ISLPOSE_OVERFLOW_PAIRS 4      /* a (frame, limb) has nA*nB > pair_cap */
#define ISLPOSE_OVERFLOW_PERSONS 8    /* a frame created more than max_person rows */

int islpose_body_group(const islpose_scale* paf_scales, int32_t n_scales, int32_t model_kind, int32_t n, int32_t H,
                       int32_t W, double thre2, int32_t mid_num, const islpose_group_buffers* buffers, void* stream);

/* The classifier's input row per frame, float64 [n][156] (src/util.py:99-151,187-219 get_bodypose / get_handpose,
 * src/ISL_Model_parameter.py:376-410 populate_features): [0,15) x and [15,30) y of the first 15 body joints found (joint-major,
 * person-minor), then per hand (the frame's first two): 21 x, 21 y, 21 key-point numbers.
 * islpose_body_features writes the body part of every row and zero-fills the rest; candidate / subset / n_person are the
 * outputs of islpose_body_group. islpose_hand_features then fills the hand parts: table int32 [n_hands][4] = (frame, slot 0|1,
 * crop x, crop y) on the DEVICE, hand_xy = islpose_hand_keypoints' out_xy of those hands (crop coordinates; non-zero
 * coordinates are shifted by the crop origin as demo.py:36-37 does). */
int islpose_body_features(const double* candidate, const double* subset, const int32_t* n_person, int32_t n, int32_t max_cand,
                          int32_t max_person, int32_t model_kind, double* features, void* stream);
int islpose_hand_features(const int32_t* table, const int32_t* hand_xy, int32_t n_hands, int32_t n_frames, double* features,
                          void* stream);

/* The sign classifier applied to windows of feature rows: the Keras Sequential of demo_isl_translate.py:72-99
 * (Masking(0) -> BatchNorm -> BiLSTM(32, sequences) -> BiLSTM(32) -> ELU -> Dense(32) -> BN -> ELU -> Dense(32) -> BN -> ELU ->
 * Dense(classes, softmax)) as ISLSignPosTranslator.call applies it (src/ISL_Model_parameter.py:353), inference, float32.
 * windows: float64 [n][T][156] (rows as islpose_body_features / islpose_hand_features leave them; all-zero rows are masked
 * steps), T <= 32. weights: `translation_model.get_weights()` concatenated in that order as float32 - BatchNorm (gamma, beta,
 * moving_mean, moving_variance), per LSTM layer forward then backward (kernel [in][128], recurrent_kernel [32][128], bias [128];
 * gate order i, f, c, o), Dense kernels [in][out], the last Dense's bias - islpose_translate_weight_floats(classes) floats.
 * probs: float32 [n][classes]. One launch. */
int64_t islpose_translate_weight_floats(int32_t classes);
int islpose_translate(const double* windows, int32_t n, int32_t T, int32_t n_features, const float* weights, int64_t n_weights,
                      int32_t classes, float* probs, void* stream);

/* Hand key points (src/hand.py:51-74), batched over crops of any sizes: per crop and scale both cubic stages and the
 * float64 mean over the scales (hand.py:51-56), gaussian sigma=3, threshold, 8-connected labelling, the component with the
 * largest mass of the unsmoothed map (numpy's summation order where components tie to within rounding), first arg-max
 * (util.npmax). Three launches per 32 crops. h_crops is a HOST array; scales[s].lowres points at THIS crop's network
 * output, fp32 [22][gh][gw]. workspace: device scratch of islpose_hand_workspace_bytes() bytes, 256-byte aligned.
 * out_xy int32 [n_crops][21][2] (x, y in crop coordinates; 0,0 = not found). n_scales <= 4 (hand.py:25). */
typedef struct islpose_hand_crop {
  int32_t h, w;
  islpose_scale scales[4];
} islpose_hand_crop;

int64_t islpose_hand_workspace_bytes(const islpose_hand_crop* h_crops, int32_t n_crops);
int islpose_hand_keypoints(const islpose_hand_crop* h_crops, int32_t n_crops, int32_t n_scales, const double* h_gauss,
                           double thre, void* workspace, int64_t workspace_bytes, int32_t* out_xy, void* stream);

/* The selection stage of islpose_hand_keypoints alone (hand.py:58-74) on caller-supplied float64 maps:
 * heat [planes][H][W] (all planes of one size); labels / mass: scratch of planes*H*W elements each;
 * out_xy int32 [planes][2]. */
int islpose_hand_peaks(const double* heat, int32_t planes, int32_t H, int32_t W, const double* h_gauss, double thre,
                       int32_t* labels, double* mass, int32_t* out_xy, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ISLPOSE_H_ */
