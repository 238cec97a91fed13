// extern "C" surface declared in include/islpose.h. Plain pointers and sizes only.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <vector>

#include "../../include/islpose.h"
#include "conv_umma.cuh"
#include "prepost.cuh"

using namespace islpose;

namespace {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};  // kernels launched by this library since load (islpose_launch_count)

int set_err(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

int check_cuda(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_err("%s: %s", what, cudaGetErrorString(e));
  return 0;
}

enum OpKind { kConv = 0, kFirst = 3 };

struct Op {
  OpKind kind;
  ConvLaunch conv;  // kConv
  const void* in;   // kFirst
  void* out;
  int n, h, w, c;
  const void* weights;  // kFirst
  const float* bias;
  const float* slope;
  bool relu;            // kFirst: every slope is 0
};

const int kCocoA[19] = {1, 1, 2, 3, 5, 6, 1, 8, 9, 1, 11, 12, 1, 0, 14, 0, 15, 2, 5};
const int kCocoB[19] = {2, 5, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 0, 14, 16, 15, 17, 16, 17};
const int kCocoX[19] = {12, 20, 14, 16, 22, 24, 0, 2, 4, 6, 8, 10, 28, 30, 34, 32, 36, 18, 26};
const int kB25A[24] = {1, 1, 2, 3, 1, 5, 6, 1, 8, 9, 10, 8, 12, 13, 0, 0, 15, 16, 11, 11, 14, 14, 22, 19};
const int kB25B[24] = {0, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 24, 22, 21, 19, 23, 20};
const int kB25X[24] = {30, 14, 16, 18, 22, 24, 26, 0, 6, 2, 4, 8, 10, 12, 32, 34, 36, 38, 50, 46, 44, 40, 48, 42};

LimbTable limb_table(int model_kind) {
  LimbTable lt;
  memset(&lt, 0, sizeof(lt));
  if (model_kind == 1) {
    lt.nlimbs = 24;
    lt.njoint = 26;
    for (int k = 0; k < 24; ++k) {
      lt.a[k] = kB25A[k];
      lt.b[k] = kB25B[k];
      lt.cx[k] = kB25X[k];
      lt.cy[k] = kB25X[k] + 1;
    }
  } else {
    lt.nlimbs = 19;
    lt.njoint = 19;
    for (int k = 0; k < 19; ++k) {
      lt.a[k] = kCocoA[k];
      lt.b[k] = kCocoB[k];
      lt.cx[k] = kCocoX[k];
      lt.cy[k] = kCocoX[k] + 1;
    }
  }
  return lt;
}

int fill_scales(const islpose_scale* scales, int n_scales, int channels, int H, int W, ScaleSet* ss) {
  if (n_scales < 1 || n_scales > kMaxScales) return set_err("between 1 and %d scales are supported, got %d", kMaxScales, n_scales);
  memset(ss, 0, sizeof(*ss));
  ss->count = n_scales;
  ss->channels = channels;
  for (int s = 0; s < n_scales; ++s) {
    const islpose_scale& in = scales[s];
    if (in.lowres == nullptr || in.gh <= 0 || in.gw <= 0 || in.hc <= 0 || in.wc <= 0 || in.hc > in.gh * 8 || in.wc > in.gw * 8)
      return set_err("scale %d: inconsistent geometry (grid %dx%d, crop %dx%d)", s, in.gh, in.gw, in.hc, in.wc);
    ScaleGeom& g = ss->g[s];
    g.low = in.lowres;
    g.gh = in.gh;
    g.gw = in.gw;
    g.hc = in.hc;
    g.wc = in.wc;
    // cv2.resize with an explicit dsize: inv_scale = (double)dst / src; scale = 1. / inv_scale
    g.sx = 1.0 / (static_cast<double>(W) / in.wc);
    g.sy = 1.0 / (static_cast<double>(H) / in.hc);
  }
  return 0;
}

}  // namespace

struct islpose_plan {
  std::vector<Op> ops;
  double flops = 0;
  // The launches of a plan never change (buffers, tensor maps and shapes are fixed when it is built), so the first
  // islpose_plan_run captures them into a CUDA graph and every later run is ONE graph launch: a 93..117-kernel network
  // costs the host ~0.4 ms to launch kernel by kernel, which is the critical path of a single-frame call where eight
  // such networks (4 body + 4 hand scales) are queued one after the other. Programmatic dependent launch edges survive
  // the capture. graph_state: 0 = not tried, 1 = ready, -1 = capture not possible here (direct launches from then on).
  int use_graph = 1;
  int graph_state = 0;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t graph_exec = nullptr;
  char graph_note[160] = "";   // why recording failed (islpose_plan_graph_note)
  ~islpose_plan() {
    if (graph_exec != nullptr) cudaGraphExecDestroy(graph_exec);
    if (graph != nullptr) cudaGraphDestroy(graph);
  }
};

extern "C" {

int islpose_abi_version(void) { return ISLPOSE_ABI_VERSION; }
int64_t islpose_launch_count(void) { return g_launches.load(); }
int islpose_struct_sizes(int32_t out[4]) {
  if (out == nullptr) return set_err("struct_sizes: null pointer");
  out[0] = static_cast<int32_t>(sizeof(islpose_scale));
  out[1] = static_cast<int32_t>(sizeof(islpose_conv_desc));
  out[2] = static_cast<int32_t>(sizeof(islpose_group_buffers));
  out[3] = static_cast<int32_t>(sizeof(islpose_hand_crop));
  return 0;
}
const char* islpose_last_error(void) { return g_err; }

int islpose_plan_create(islpose_plan** out) {
  if (out == nullptr) return set_err("plan_create: null output");
  *out = new islpose_plan();
  return 0;
}

int islpose_plan_destroy(islpose_plan* plan) {
  delete plan;
  return 0;
}

int islpose_plan_add_conv(islpose_plan* plan, const islpose_conv_desc* d) {
  if (plan == nullptr || d == nullptr) return set_err("plan_add_conv: null argument");
  ConvDesc c;
  memset(&c, 0, sizeof(c));
  c.in = static_cast<const __nv_bfloat16*>(d->in);
  c.in_c = d->in_c;
  c.in_cstride = d->in_cstride;
  c.in_c_readable = d->in_c_readable;
  c.w_cin = d->w_cin;
  c.N = d->n;
  c.H = d->h;
  c.W = d->w;
  c.w = static_cast<const __nv_bfloat16*>(d->weights);
  c.cout = d->cout;
  c.ksize = d->ksize;
  c.bias = d->bias;
  c.slope = d->slope;
  c.out_bf16 = static_cast<__nv_bfloat16*>(d->out_bf16);
  c.out_cstride = d->out_cstride;
  c.out_f32 = d->out_f32;
  c.out_f32_channels = d->out_f32_channels;
  c.force_n_tile = d->n_tile;
  c.force_stages = d->stages;
  c.force_bw = d->tile_w;
  c.force_bh = d->tile_h;
  c.pool = d->pool;
  c.sm_budget = d->sm_budget;
  c.w2 = static_cast<const __nv_bfloat16*>(d->weights2);
  c.cout2 = d->cout2;
  c.bias2 = d->bias2;
  c.slope2 = d->slope2;
  c.out2_bf16 = static_cast<__nv_bfloat16*>(d->out2_bf16);
  c.out2_cstride = d->out2_cstride;
  c.out2b_bf16 = static_cast<__nv_bfloat16*>(d->out2b_bf16);
  c.out2b_cstride = d->out2b_cstride;
  c.out2_f32 = d->out2_f32;
  c.out2_f32_channels = d->out2_f32_channels;
  Op op;
  memset(&op, 0, sizeof(op));
  op.kind = kConv;
  char err[256] = "";
  if (conv_prepare(c, &op.conv, err, sizeof(err)) != 0) return set_err("%s", err);
  plan->ops.push_back(op);
  plan->flops += op.conv.flops;
  return 0;
}

int islpose_plan_add_first_conv(islpose_plan* plan, const float* in_nchw, const void* weights, const float* bias,
                                const float* slope, void* out, int32_t out_cstride, int32_t n, int32_t h, int32_t w,
                                int32_t relu) {
  if (plan == nullptr || in_nchw == nullptr || weights == nullptr || bias == nullptr || slope == nullptr || out == nullptr)
    return set_err("plan_add_first_conv: null argument");
  if (out_cstride < 64 || out_cstride % 8 != 0 || (reinterpret_cast<uintptr_t>(out) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(weights) & 15) != 0)
    return set_err("plan_add_first_conv: output needs >= 64 channels per pixel (multiple of 8) and 16-byte alignment");
  Op op;
  memset(&op, 0, sizeof(op));
  op.kind = kFirst;
  op.in = in_nchw;
  op.out = out;
  op.n = n;
  op.h = h;
  op.w = w;
  op.c = out_cstride;
  op.weights = weights;
  op.bias = bias;
  op.slope = slope;
  op.relu = relu != 0;
  plan->ops.push_back(op);
  plan->flops += 2.0 * 32 * 64 * static_cast<double>(n) * h * w;
  return 0;
}

static int run_op(const Op& op, cudaStream_t st) {
  if (op.kind == kConv) return conv_run(op.conv, st);
  return launch_conv_first(static_cast<const float*>(op.in), op.n, op.h, op.w, op.weights, op.bias, op.slope, op.out, op.c, op.relu, st);
}

int islpose_plan_set_graph(islpose_plan* plan, int32_t enable) {
  if (plan == nullptr) return set_err("plan_set_graph: null plan");
  plan->use_graph = enable ? 1 : 0;
  return 0;
}

static int plan_capture(islpose_plan* plan) {
  // Recorded on a private stream (the caller's may be the legacy default stream, which cannot capture; nothing executes
  // while recording, so no ordering with the caller's stream is needed). Thread-local mode: other host threads
  // (allocators, other lanes) keep making CUDA calls while this one records.
  cudaStream_t st = nullptr;
  if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) {
    snprintf(plan->graph_note, sizeof(plan->graph_note), "no capture stream: %s", cudaGetErrorString(cudaGetLastError()));
    return 1;
  }
  cudaError_t e = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
  int rc = 0;
  cudaGraph_t g = nullptr;
  if (e == cudaSuccess) {
    for (size_t i = 0; i < plan->ops.size() && rc == 0; ++i) rc = run_op(plan->ops[i], st);
    e = cudaStreamEndCapture(st, &g);
  }
  cudaGraphExec_t ge = nullptr;
  if (rc == 0 && e == cudaSuccess && g != nullptr) e = cudaGraphInstantiate(&ge, g, 0);
  cudaStreamDestroy(st);
  if (rc != 0 || e != cudaSuccess || ge == nullptr) {
    snprintf(plan->graph_note, sizeof(plan->graph_note), "capture failed (launch rc %d): %s", rc, cudaGetErrorString(e));
    if (g != nullptr) cudaGraphDestroy(g);
    cudaGetLastError();
    return 1;
  }
  plan->graph = g;
  plan->graph_exec = ge;
  return 0;
}

int islpose_plan_run(islpose_plan* plan, void* stream) {
  if (plan == nullptr) return set_err("plan_run: null plan");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (plan->use_graph && plan->graph_state == 0 && plan->ops.size() > 1) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs == cudaStreamCaptureStatusNone)
      plan->graph_state = plan_capture(plan) == 0 ? 1 : -1;
    else
      cudaGetLastError();
  }
  if (plan->use_graph && plan->graph_state == 1) {
    if (cudaGraphLaunch(plan->graph_exec, st) != cudaSuccess) return check_cuda("plan_run/graph") ? 1 : set_err("plan_run: graph launch failed");
    g_launches.fetch_add(static_cast<long long>(plan->ops.size()), std::memory_order_relaxed);
    return 0;
  }
  for (size_t i = 0; i < plan->ops.size(); ++i) {
    const Op& op = plan->ops[i];
    const int rc = run_op(op, st);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (rc != 0) {
      check_cuda("plan_run");
      return set_err("plan_run: launch %d of %d failed: %s", static_cast<int>(i), static_cast<int>(plan->ops.size()), g_err);
    }
  }
  return 0;
}

int islpose_plan_profile(const islpose_plan* plan, void* stream, int32_t reps, float* h_ms, double* h_flops,
                         int32_t* h_variant) {
  if (plan == nullptr || h_ms == nullptr || reps <= 0) return set_err("plan_profile: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaEvent_t e0, e1;
  if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return set_err("plan_profile: no events");
  int rc = 0;
  for (size_t i = 0; i < plan->ops.size() && rc == 0; ++i) {
    const Op& op = plan->ops[i];
    // every launch is timed on its own, after all earlier launches of the plan have run once (valid inputs)
    for (int r = -1; r < reps && rc == 0; ++r) {
      if (r == 0) cudaEventRecord(e0, st);
      rc = run_op(op, st);
    }
    cudaEventRecord(e1, st);
    if (cudaEventSynchronize(e1) != cudaSuccess) rc = 1;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    h_ms[i] = ms / reps;
    if (h_flops != nullptr)
      h_flops[i] = op.kind == kConv ? op.conv.flops : 2.0 * 27 * 64 * static_cast<double>(op.n) * op.h * op.w;
    if (h_variant != nullptr) h_variant[i] = op.kind == kConv ? op.conv.variant : 0;
    g_launches.fetch_add(reps + 1, std::memory_order_relaxed);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (rc != 0) return check_cuda("plan_profile") ? 1 : set_err("plan_profile: a launch failed");
  return 0;
}

int32_t islpose_plan_graph_state(const islpose_plan* plan) { return plan ? plan->graph_state : 0; }
const char* islpose_plan_graph_note(const islpose_plan* plan) { return plan ? plan->graph_note : ""; }
int32_t islpose_plan_num_launches(const islpose_plan* plan) { return plan ? static_cast<int32_t>(plan->ops.size()) : 0; }
double islpose_plan_conv_flops(const islpose_plan* plan) { return plan ? plan->flops : 0.0; }

int islpose_pack_conv_weights(const float* w, int32_t cout, int32_t cin, int32_t ksize, const int32_t* chan_map, int32_t in_c,
                              int32_t w_cin, int32_t first_layer, void* out_bf16, void* stream) {
  if (w == nullptr || out_bf16 == nullptr) return set_err("pack_conv_weights: null pointer");
  if (cout <= 0 || cin <= 0 || (ksize != 1 && ksize != 3 && ksize != 7) || in_c <= 0 || w_cin < in_c || w_cin % 8 != 0)
    return set_err("pack_conv_weights: bad shape (cout %d, cin %d, k %d, slice %d, stride %d)", cout, cin, ksize, in_c, w_cin);
  if (first_layer && (ksize * ksize * cin > w_cin)) return set_err("pack_conv_weights: first-layer patch of %d values exceeds %d", ksize * ksize * cin, w_cin);
  if (!first_layer && chan_map == nullptr && cin != in_c) return set_err("pack_conv_weights: slice width %d differs from Cin %d and no channel map was given", in_c, cin);
  if (launch_pack_conv_weights(w, cout, cin, ksize, chan_map, in_c, w_cin, first_layer, out_bf16, static_cast<cudaStream_t>(stream)) != 0)
    return check_cuda("pack_conv_weights") ? 1 : set_err("pack_conv_weights: launch failed");
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int islpose_resize_pad_normalize(const uint8_t* frames, int32_t n, int32_t H, int32_t W, double scale, int32_t rh,
                                 int32_t rw, int32_t hp, int32_t wp, float* out_nchw, uint8_t* out_u8, void* stream) {
  if (frames == nullptr || out_nchw == nullptr) return set_err("resize_pad_normalize: null pointer");
  if (n <= 0 || H <= 0 || W <= 0 || rh <= 0 || rw <= 0 || hp < rh || wp < rw || hp % 8 != 0 || wp % 8 != 0 || !(scale > 0))
    return set_err("resize_pad_normalize: bad geometry (%dx%d -> %dx%d padded %dx%d)", H, W, rh, rw, hp, wp);
  if (launch_resize_pad_norm(frames, n, H, W, scale, rh, rw, hp, wp, out_nchw, out_u8, static_cast<cudaStream_t>(stream)) != 0)
    return check_cuda("resize_pad_normalize") ? 1 : set_err("resize_pad_normalize: launch failed");
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int64_t islpose_maps_workspace_floats(const islpose_scale* scales, int32_t n_scales, int32_t n, int32_t parts) {
  int64_t total = 0;
  if (scales == nullptr) return 0;
  for (int s = 0; s < n_scales; ++s) total += static_cast<int64_t>(n) * parts * scales[s].hc * ((scales[s].wc + 3) / 4 * 4);
  return total;
}

int islpose_maps_accumulate(const islpose_scale* scales, int32_t n_scales, int32_t channels, int32_t n, int32_t H,
                            int32_t W, int32_t parts, int32_t double_running_sum, double* out, float* workspace,
                            int64_t workspace_floats, void* stream) {
  if (scales == nullptr || out == nullptr) return set_err("maps_accumulate: null pointer");
  if (parts <= 0 || parts > channels || n <= 0 || H <= 0 || W <= 0) return set_err("maps_accumulate: bad sizes");
  ScaleSet ss;
  if (fill_scales(scales, n_scales, channels, H, W, &ss) != 0) return 1;
  if (launch_heat_accumulate(ss, n, H, W, parts, double_running_sum, out, workspace, workspace_floats,
                             static_cast<cudaStream_t>(stream)) != 0)
    return check_cuda("maps_accumulate") ? 1 : set_err("maps_accumulate: launch failed");
  const bool two_pass = workspace != nullptr && workspace_floats >= heat_accumulate_workspace_floats(ss, n, parts);
  g_launches.fetch_add(two_pass ? n_scales + 1 : 1, std::memory_order_relaxed);
  return 0;
}

int islpose_body_peaks(const double* heat, int32_t planes, int32_t H, int32_t W, const double* h_gauss, double thre1,
                       int32_t cap, int32_t* counts, uint32_t* keys, double* scores, int32_t* overflow, void* stream) {
  if (heat == nullptr || h_gauss == nullptr || counts == nullptr || keys == nullptr || scores == nullptr || overflow == nullptr)
    return set_err("body_peaks: null pointer");
  if (cap <= 0 || cap > kMaxPeakCap || (cap > 1024 && (cap & (cap - 1)) != 0))
    return set_err("body_peaks: cap must be in 1..1024 or a power of two up to %d, got %d", kMaxPeakCap, cap);
  GaussWeights gw;
  memcpy(gw.w, h_gauss, sizeof(gw.w));
  if (launch_gauss_nms(heat, planes, H, W, gw, thre1, cap, counts, keys, scores, overflow, static_cast<cudaStream_t>(stream)) != 0)
    return check_cuda("body_peaks") ? 1 : set_err("body_peaks: launch failed");
  g_launches.fetch_add(2, std::memory_order_relaxed);
  return 0;
}

int islpose_body_group(const islpose_scale* paf_scales, int32_t n_scales, int32_t model_kind, int32_t n, int32_t H,
                       int32_t W, double thre2, int32_t mid_num, const islpose_group_buffers* b, void* stream) {
  if (paf_scales == nullptr || b == nullptr) return set_err("body_group: null pointer");
  if (model_kind != 0 && model_kind != 1) return set_err("body_group: model_kind must be 0 (coco) or 1 (body25)");
  if (mid_num != 10) return set_err("body_group: mid_num is fixed at 10 (body.py:130), got %d", mid_num);
  const LimbTable lt = limb_table(model_kind);
  ScaleSet ss;
  if (fill_scales(paf_scales, n_scales, model_kind == 1 ? 52 : 38, H, W, &ss) != 0) return 1;
  GroupBuffers gb;
  gb.cap = b->cap;
  gb.counts = b->counts;
  gb.keys = b->keys;
  gb.scores = b->scores;
  gb.pair_cap = b->pair_cap;
  gb.pair_score = b->pair_score;
  gb.end_paf = b->end_paf;
  gb.conn_count = b->conn_count;
  gb.conn_ij = b->conn_ij;
  gb.conn_score = b->conn_score;
  gb.owner = b->owner;
  gb.max_cand = b->max_cand;
  gb.candidate = b->candidate;
  gb.n_cand = b->n_cand;
  gb.max_person = b->max_person;
  gb.subset = b->subset;
  gb.n_person = b->n_person;
  gb.overflow = b->overflow;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (launch_paf_score(ss, lt, n, H, W, thre2, mid_num, gb, st) != 0)
    return check_cuda("body_group/paf_score") ? 1 : set_err("body_group: peak capacity out of range (cap %d)", gb.cap);
  if (launch_group(lt, n, W, gb, st) != 0) return check_cuda("body_group/group") ? 1 : set_err("body_group: launch failed");
  g_launches.fetch_add(4, std::memory_order_relaxed);
  return 0;
}

int islpose_body_features(const double* candidate, const double* subset, const int32_t* n_person, int32_t n, int32_t max_cand,
                          int32_t max_person, int32_t model_kind, double* features, void* stream) {
  if (candidate == nullptr || subset == nullptr || n_person == nullptr || features == nullptr) return set_err("body_features: null pointer");
  if (model_kind != 0 && model_kind != 1) return set_err("body_features: model_kind must be 0 (coco) or 1 (body25)");
  if (n <= 0 || max_cand <= 0 || max_person <= 0) return set_err("body_features: bad sizes");
  if (launch_body_features(candidate, subset, n_person, n, max_cand, max_person, model_kind == 1 ? 26 : 19, features,
                           static_cast<cudaStream_t>(stream)) != 0)
    return check_cuda("body_features") ? 1 : set_err("body_features: launch failed");
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int islpose_hand_features(const int32_t* table, const int32_t* hand_xy, int32_t n_hands, int32_t n_frames, double* features,
                          void* stream) {
  if (n_hands == 0) return 0;
  if (table == nullptr || hand_xy == nullptr || features == nullptr || n_hands < 0 || n_frames <= 0)
    return set_err("hand_features: bad argument");
  if (launch_hand_features(table, hand_xy, n_hands, n_frames, features, static_cast<cudaStream_t>(stream)) != 0)
    return check_cuda("hand_features") ? 1 : set_err("hand_features: launch failed");
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int64_t islpose_translate_weight_floats(int32_t classes) { return classes > 0 ? translate_weight_floats(classes) : 0; }

int islpose_translate(const double* windows, int32_t n, int32_t T, int32_t n_features, const float* weights, int64_t n_weights,
                      int32_t classes, float* probs, void* stream) {
  if (n == 0) return 0;
  if (windows == nullptr || weights == nullptr || probs == nullptr || n < 0) return set_err("translate: bad argument");
  if (n_features != 156) return set_err("translate: rows have 156 features (ISL_Model_parameter.py:376-410), got %d", n_features);
  if (T <= 0 || T > 32) return set_err("translate: window length must be in 1..32, got %d", T);
  if (classes <= 0 || classes > 4096) return set_err("translate: classes must be in 1..4096, got %d", classes);
  if (n_weights != translate_weight_floats(classes))
    return set_err("translate: %d classes need %lld weight floats", classes, static_cast<long long>(translate_weight_floats(classes)));
  if (launch_translate(windows, n, T, weights, n_weights, classes, probs, static_cast<cudaStream_t>(stream)) != 0)
    return check_cuda("translate") ? 1 : set_err("translate: launch failed");
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

static int fill_gauss(const double* h_gauss, GaussWeights* gw) {
  if (h_gauss == nullptr) return set_err("hand: null gaussian weights");
  memcpy(gw->w, h_gauss, sizeof(gw->w));
  return 0;
}

int islpose_hand_peaks(const double* heat, int32_t planes, int32_t H, int32_t W, const double* h_gauss, double thre,
                       int32_t* labels, double* mass, int32_t* out_xy, void* stream) {
  if (heat == nullptr || labels == nullptr || mass == nullptr || out_xy == nullptr) return set_err("hand_peaks: null pointer");
  if (planes <= 0 || H <= 0 || W <= 0 || static_cast<int64_t>(H) * W > (1 << 30)) return set_err("hand_peaks: bad sizes");
  GaussWeights gw;
  if (fill_gauss(h_gauss, &gw) != 0) return 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // the planes are handed to the batched kernels as "crops" of up to 21 planes each
  const int64_t plane = static_cast<int64_t>(H) * W;
  for (int p0 = 0; p0 < planes;) {
    HandBatch hb;
    memset(&hb, 0, sizeof(hb));
    hb.parts = planes - p0 < 21 ? planes - p0 : 21;
    hb.channels = 22;
    while (hb.n_crops < kHandMaxCrops && p0 + hb.parts <= planes) {
      HandCropDev& c = hb.crop[hb.n_crops++];
      c.H = H;
      c.W = W;
      c.heat = const_cast<double*>(heat) + p0 * plane;
      c.labels = labels + p0 * plane;
      c.mass = mass + p0 * plane;
      c.out_xy = out_xy + p0 * 2;
      p0 += hb.parts;
    }
    if (launch_hand_keypoints(hb, false, gw, thre, st) != 0)
      return check_cuda("hand_peaks") ? 1 : set_err("hand_peaks: launch failed");
    g_launches.fetch_add(2, std::memory_order_relaxed);
  }
  return 0;
}

int64_t islpose_hand_workspace_bytes(const islpose_hand_crop* h_crops, int32_t n_crops) {
  int64_t total = 0;
  if (h_crops == nullptr) return 0;
  for (int i = 0; i < n_crops; ++i) {
    const int64_t elems = 21LL * h_crops[i].h * h_crops[i].w;
    total += ((elems * 8 + 255) / 256 * 256) * 2 + (elems * 4 + 255) / 256 * 256;  // heat, mass (float64), labels (int32)
  }
  return total;
}

int islpose_hand_keypoints(const islpose_hand_crop* h_crops, int32_t n_crops, int32_t n_scales, const double* h_gauss,
                           double thre, void* workspace, int64_t workspace_bytes, int32_t* out_xy, void* stream) {
  if (h_crops == nullptr || workspace == nullptr || out_xy == nullptr) return set_err("hand_keypoints: null pointer");
  if (n_crops <= 0) return set_err("hand_keypoints: no crops");
  if (n_scales < 1 || n_scales > kHandMaxScales)
    return set_err("hand_keypoints: between 1 and %d scales are supported (hand.py:25 uses 4), got %d", kHandMaxScales, n_scales);
  if (workspace_bytes < islpose_hand_workspace_bytes(h_crops, n_crops))
    return set_err("hand_keypoints: workspace of %lld bytes is too small (need %lld)", static_cast<long long>(workspace_bytes),
                   static_cast<long long>(islpose_hand_workspace_bytes(h_crops, n_crops)));
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return set_err("hand_keypoints: workspace must be 256-byte aligned");
  GaussWeights gw;
  if (fill_gauss(h_gauss, &gw) != 0) return 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* cursor = static_cast<uint8_t*>(workspace);
  for (int c0 = 0; c0 < n_crops; c0 += kHandMaxCrops) {
    HandBatch hb;
    memset(&hb, 0, sizeof(hb));
    hb.n_crops = n_crops - c0 < kHandMaxCrops ? n_crops - c0 : kHandMaxCrops;
    hb.n_scales = n_scales;
    hb.channels = 22;
    hb.parts = 21;
    for (int i = 0; i < hb.n_crops; ++i) {
      const islpose_hand_crop& in = h_crops[c0 + i];
      HandCropDev& c = hb.crop[i];
      if (in.h <= 0 || in.w <= 0 || static_cast<int64_t>(in.h) * in.w > (1 << 26))
        return set_err("hand_keypoints: crop %d has a bad size %dx%d", c0 + i, in.h, in.w);
      c.H = in.h;
      c.W = in.w;
      for (int s = 0; s < n_scales; ++s) {
        const islpose_scale& sc = in.scales[s];
        if (sc.lowres == nullptr || sc.gh <= 0 || sc.gw <= 0 || sc.hc <= 0 || sc.wc <= 0 || sc.hc > sc.gh * 8 || sc.wc > sc.gw * 8)
          return set_err("hand_keypoints: crop %d scale %d: inconsistent geometry (grid %dx%d, crop %dx%d)", c0 + i, s, sc.gh,
                         sc.gw, sc.hc, sc.wc);
        c.low[s] = sc.lowres;
        c.gh[s] = sc.gh;
        c.gw[s] = sc.gw;
        c.hc[s] = sc.hc;
        c.wc[s] = sc.wc;
      }
      const int64_t elems = 21LL * in.h * in.w;
      c.heat = reinterpret_cast<double*>(cursor);
      cursor += (elems * 8 + 255) / 256 * 256;
      c.mass = reinterpret_cast<double*>(cursor);
      cursor += (elems * 8 + 255) / 256 * 256;
      c.labels = reinterpret_cast<int*>(cursor);
      cursor += (elems * 4 + 255) / 256 * 256;
      c.out_xy = out_xy + static_cast<int64_t>(c0 + i) * 21 * 2;
    }
    if (launch_hand_keypoints(hb, true, gw, thre, st) != 0)
      return check_cuda("hand_keypoints") ? 1 : set_err("hand_keypoints: launch failed");
    g_launches.fetch_add(3, std::memory_order_relaxed);
  }
  return 0;
}

}  // extern "C"
