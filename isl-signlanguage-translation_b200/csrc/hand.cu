// Hand key points (reference: src/hand.py:51-74), batched over all crops of a network replay: three launches per
// batch of up to kHandMaxCrops crops of any sizes (a per-crop table travels as a kernel parameter).
//   hand_heat_kernel    hand.py:51-56  both cubic stages per scale + the float64 mean over the scales -> heat [21][h][w]
//   hand_gauss_kernel   hand.py:61-62  gaussian sigma=3 (gauss.cuh), smoothed > thre -> initial labels
//   hand_select_kernel  hand.py:63-73  one CTA per (crop, part) plane: 8-connected labelling, heaviest component of the
//                       *unsmoothed* map, everything else zeroed, first arg-max (util.npmax, util.py:394-399)
// Labels are (smallest flat pixel index in the component)+1, i.e. the raster order skimage.measure.label numbers
// components in, so "first maximum" ties resolve the same way.
//
// Component masses, deterministically: the reference takes np.argmax over np.sum(map_ori[label_img == i]) (hand.py:68),
// i.e. numpy's pairwise summation over the component's pixels in raster order. Masses are first accumulated with
// float64 atomics (any order) only to find which components can be the maximum at all: a component whose approximate
// mass is more than kMassTol * sum|values| below the largest cannot win (the atomic sum of n <= 2^16 terms is within
// 7.3e-12 * sum|values| of the exact one, kMassTol is 1e-10). If one component is left - the normal case - it is kept;
// otherwise the masses of the contenders are re-computed exactly as numpy does (np_pairwise_sum below walks the
// component in raster order through numpy's 8-accumulator / 128-element-block / halving recursion) and the first
// largest wins. The kept component and hence the key point never depend on the order of the atomics.
#include "gauss.cuh"
#include "prepost.cuh"

namespace islpose {

namespace {

constexpr double kMassTol = 1e-10;

struct Best {
  double v;
  int idx;
};
// larger value wins, ties go to the smaller index (np.argmax returns the first maximum)
__device__ __forceinline__ Best better(Best a, Best b) {
  if (b.idx < 0) return a;
  if (a.idx < 0) return b;
  if (b.v > a.v || (b.v == a.v && b.idx < a.idx)) return b;
  return a;
}
__device__ Best block_best(Best mine, Best* s_tmp) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Best other;
    other.v = __shfl_xor_sync(0xffffffffu, mine.v, o);
    other.idx = __shfl_xor_sync(0xffffffffu, mine.idx, o);
    mine = better(mine, other);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) s_tmp[warp] = mine;
  __syncthreads();
  Best r = s_tmp[0];
  for (int w = 1; w < static_cast<int>(blockDim.x >> 5); ++w) r = better(r, s_tmp[w]);
  __syncthreads();
  return r;
}
__device__ double block_sum_any_order(double mine, double* s_tmp) {  // used for a bound only
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) s_tmp[warp] = mine;
  __syncthreads();
  double r = 0.0;
  for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) r += s_tmp[w];
  __syncthreads();
  return r;
}

// numpy's DOUBLE_pairwise_sum (numpy/_core/src/umath/loops_utils.h.src) over the pixels of one component in raster
// order, executed by ONE thread. The elements are consumed strictly left to right (the recursion visits its leaves in
// order and a leaf reads its elements in order), so the component needs no compaction: `next` scans the label plane.
struct ComponentCursor {
  const int* lab;
  const double* hm;
  int label;
  int p;
  __device__ double next() {
    while (lab[p] != label) ++p;
    return hm[p++];
  }
};
__device__ double np_pairwise_leaf(ComponentCursor& cur, int n) {
  if (n < 8) {
    double res = 0.;
    for (int i = 0; i < n; ++i) res = __dadd_rn(res, cur.next());
    return res;
  }
  double r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = cur.next();
  int i = 8;
  for (; i < n - (n % 8); i += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], cur.next());
  }
  double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                         __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
  for (; i < n; ++i) res = __dadd_rn(res, cur.next());
  return res;
}
__device__ double np_pairwise_sum(ComponentCursor& cur, int n) {
  // post-order walk of sum(n) = n <= 128 ? leaf(n) : sum(n2) + sum(n - n2), n2 = n/2 rounded down to a multiple of 8
  int len[32];
  unsigned char state[32];
  double left[32];
  int sp = 1;
  len[0] = n;
  state[0] = 0;
  double ret = 0.0;
  while (sp > 0) {
    const int t = sp - 1;
    if (state[t] == 0) {
      if (len[t] <= 128) {
        ret = np_pairwise_leaf(cur, len[t]);
        --sp;
        continue;
      }
      int n2 = len[t] / 2;
      n2 -= n2 % 8;
      state[t] = 1;
      len[sp] = n2;
      state[sp] = 0;
      ++sp;
    } else if (state[t] == 1) {
      left[t] = ret;
      state[t] = 2;
      int n2 = len[t] / 2;
      n2 -= n2 % 8;
      len[sp] = len[t] - n2;
      state[sp] = 0;
      ++sp;
    } else {
      ret = __dadd_rn(left[t], ret);
      --sp;
    }
  }
  return ret;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ heat maps
// One thread = one crop pixel x kHChunk parts: per scale both cubic stages from the stride-8 map (sample2), the float32
// division by the number of scales and the float64 accumulation of hand.py:56 (`avg += m / S`).
constexpr int kHChunk = 4;
__global__ void __launch_bounds__(256)
hand_heat_kernel(const HandBatch hb) {
  __shared__ float s_tab[8][4];
  fill_phase_table(s_tab);
  __syncthreads();
  const int chunks = (hb.parts + kHChunk - 1) / kHChunk;
  const int ci = blockIdx.z / chunks;
  const int c0 = (blockIdx.z - ci * chunks) * kHChunk;
  const HandCropDev& cr = hb.crop[ci];
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int H = cr.H, W = cr.W;
  if (x >= W || y >= H) return;
  double acc[kHChunk];
#pragma unroll
  for (int i = 0; i < kHChunk; ++i) acc[i] = 0.0;
  const int C = hb.channels;
  const long long tail_start = (static_cast<long long>(W) * C) / 4 * 4;
  const float fS = static_cast<float>(hb.n_scales);
  for (int s = 0; s < hb.n_scales; ++s) {
    const int gh = cr.gh[s], gw = cr.gw[s], hc = cr.hc[s], wc = cr.wc[s];
    // cv2.resize with an explicit dsize: inv_scale = (double)dst / src; scale = 1. / inv_scale
    const double sx = __ddiv_rn(1.0, __ddiv_rn(static_cast<double>(W), static_cast<double>(wc)));
    const double sy = __ddiv_rn(1.0, __ddiv_rn(static_cast<double>(H), static_cast<double>(hc)));
    Axis2 ax, ay;
    make_axis2(x, sx, wc, gw, s_tab, ax);
    make_axis2(y, sy, hc, gh, s_tab, ay);
    const long long plane = static_cast<long long>(gh) * gw;
#pragma unroll
    for (int i = 0; i < kHChunk; ++i) {
      const int c = c0 + i;
      if (c < hb.parts) {
        const bool tail = static_cast<long long>(x) * C + c >= tail_start;
        const float v = sample2(cr.low[s] + c * plane, gw, ax, ay, tail);
        acc[i] = __dadd_rn(acc[i], static_cast<double>(__fdiv_rn(v, fS)));
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kHChunk; ++i) {
    const int c = c0 + i;
    if (c < hb.parts) cr.heat[(static_cast<long long>(c) * H + y) * W + x] = acc[i];
  }
}

// ------------------------------------------------------------------------------------------------ gaussian + threshold
__global__ void __launch_bounds__(kG2Threads)
hand_gauss_kernel(const HandBatch hb, const GaussWeights gw, double thre) {
  __shared__ GaussSmem sm;
  const int ci = blockIdx.z / hb.parts;
  const int part = blockIdx.z - ci * hb.parts;
  const HandCropDev& cr = hb.crop[ci];
  const int x0 = blockIdx.x * kG2W, y0 = blockIdx.y * kG2H;
  if (x0 >= cr.W || y0 >= cr.H) return;  // the grid is sized for the largest crop of the batch
  const long long off = static_cast<long long>(part) * cr.H * cr.W;
  gauss_window_tile<kGaussLabels>(sm, cr.heat + off, cr.H, cr.W, x0, y0, gw, thre, 0, nullptr, nullptr, nullptr,
                                  cr.labels + off);
}

// ------------------------------------------------------------------------------------------------ component selection
__global__ void __launch_bounds__(1024)
hand_select_kernel(const HandBatch hb) {
  __shared__ int s_flag;
  __shared__ Best s_tmp[32];
  __shared__ double s_dtmp[32];
  __shared__ double s_exact;
  const int ci = blockIdx.x / hb.parts;
  const int part = blockIdx.x - ci * hb.parts;
  const HandCropDev& cr = hb.crop[ci];
  const int H = cr.H, W = cr.W;
  const int total = H * W;
  const double* hm = cr.heat + static_cast<long long>(part) * total;
  int* lab = cr.labels + static_cast<long long>(part) * total;
  double* ms = cr.mass + static_cast<long long>(part) * total;
  int32_t* out_xy = cr.out_xy + part * 2;

  if (threadIdx.x == 0) s_flag = 0;
  __syncthreads();
  int any = 0;
  for (int p = threadIdx.x; p < total; p += blockDim.x) {
    ms[p] = 0.0;
    any |= lab[p] != 0;
  }
  if (any) s_flag = 1;
  __syncthreads();
  if (s_flag == 0) {  // hand.py:64-66: nothing above the threshold
    if (threadIdx.x == 0) {
      out_xy[0] = 0;
      out_xy[1] = 0;
    }
    return;
  }
  // label propagation: neighbour minimum + pointer jumping until nothing changes
  for (int iter = 0; iter < total + 2; ++iter) {
    __syncthreads();
    if (threadIdx.x == 0) s_flag = 0;
    __syncthreads();
    int changed = 0;
    for (int p = threadIdx.x; p < total; p += blockDim.x) {
      int l = lab[p];
      if (l == 0) continue;
      const int y = p / W, x = p - y * W;
      int m = l;
      for (int dy = -1; dy <= 1; ++dy) {
        const int yy = y + dy;
        if (yy < 0 || yy >= H) continue;
        for (int dx = -1; dx <= 1; ++dx) {
          const int xx = x + dx;
          if (xx < 0 || xx >= W) continue;
          const int o = lab[yy * W + xx];
          if (o != 0 && o < m) m = o;
        }
      }
      // follow the chain of representatives a few steps
      for (int hop = 0; hop < 8; ++hop) {
        const int up = lab[m - 1];
        if (up >= m || up == 0) break;
        m = up;
      }
      if (m < l) {
        atomicMin(lab + p, m);
        changed = 1;
      }
    }
    if (changed) s_flag = 1;
    __syncthreads();
    if (s_flag == 0) break;
  }
  __syncthreads();
  // approximate component masses of the unsmoothed map (order-dependent in the last bits: used for pruning only), and
  // the bound sum|values| the pruning tolerance scales with
  double absum = 0.0;
  for (int p = threadIdx.x; p < total; p += blockDim.x) {
    const int l = lab[p];
    if (l != 0) {
      const double v = hm[p];
      atomicAdd(ms + (l - 1), v);
      absum += fabs(v);
    }
  }
  absum = block_sum_any_order(absum, s_dtmp);
  __syncthreads();
  Best mine;
  mine.v = 0.0;
  mine.idx = -1;
  for (int p = threadIdx.x; p < total; p += blockDim.x) {
    if (lab[p] == p + 1) {
      Best c;
      c.v = ms[p];
      c.idx = p;
      mine = better(mine, c);
    }
  }
  Best comp = block_best(mine, s_tmp);
  // contenders: roots whose approximate mass is within the tolerance of the largest
  const double floor_v = comp.v - kMassTol * absum;
  int contenders = 0;
  for (int p = threadIdx.x; p < total; p += blockDim.x) contenders += (lab[p] == p + 1 && ms[p] >= floor_v) ? 1 : 0;
  contenders = __syncthreads_count(contenders > 1) ? 2 : __syncthreads_count(contenders == 1);
  if (contenders > 1) {
    // rare: exact numpy-order masses of the contenders, visited in label (= raster) order; first largest wins
    Best exact_best;
    exact_best.v = 0.0;
    exact_best.idx = -1;
    int after = -1;
    while (true) {
      Best nxt;  // the contender root with the smallest index > after (block_best prefers the larger v: use -index)
      nxt.v = 0.0;
      nxt.idx = -1;
      for (int p = threadIdx.x; p < total; p += blockDim.x) {
        if (p > after && lab[p] == p + 1 && ms[p] >= floor_v) {
          Best c;
          c.v = -static_cast<double>(p);
          c.idx = p;
          nxt = better(nxt, c);
          break;  // this thread's later candidates have larger indices
        }
      }
      nxt = block_best(nxt, s_tmp);
      if (nxt.idx < 0) break;
      int members = 0;
      for (int p = threadIdx.x; p < total; p += blockDim.x) members += lab[p] == nxt.idx + 1 ? 1 : 0;
      // exact count of the component's pixels (integer block sum)
      __shared__ int s_count;
      if (threadIdx.x == 0) s_count = 0;
      __syncthreads();
      if (members) atomicAdd(&s_count, members);
      __syncthreads();
      if (threadIdx.x == 0) {
        ComponentCursor cur;
        cur.lab = lab;
        cur.hm = hm;
        cur.label = nxt.idx + 1;
        cur.p = nxt.idx;
        s_exact = np_pairwise_sum(cur, s_count);
      }
      __syncthreads();
      if (exact_best.idx < 0 || s_exact > exact_best.v) {  // np.argmax: the first maximum
        exact_best.v = s_exact;
        exact_best.idx = nxt.idx;
      }
      after = nxt.idx;
      __syncthreads();
    }
    comp = exact_best;
  }
  const int keep = comp.idx + 1;
  mine.v = 0.0;
  mine.idx = -1;
  for (int p = threadIdx.x; p < total; p += blockDim.x) {
    Best c;
    c.v = lab[p] == keep ? hm[p] : 0.0;  // hand.py:69-70 zeroes everything outside the kept component
    c.idx = p;
    mine = better(mine, c);
  }
  const Best top = block_best(mine, s_tmp);
  if (threadIdx.x == 0) {
    out_xy[0] = top.idx % W;
    out_xy[1] = top.idx / W;
  }
}

int launch_hand_keypoints(const HandBatch& hb, bool compute_heat, const GaussWeights& gw, double thre, cudaStream_t st) {
  int maxH = 0, maxW = 0;
  for (int i = 0; i < hb.n_crops; ++i) {
    maxH = hb.crop[i].H > maxH ? hb.crop[i].H : maxH;
    maxW = hb.crop[i].W > maxW ? hb.crop[i].W : maxW;
  }
  if (hb.n_crops <= 0 || maxH <= 0 || maxW <= 0) return 1;
  if (compute_heat) {
    const int chunks = (hb.parts + kHChunk - 1) / kHChunk;
    hand_heat_kernel<<<dim3((maxW + 31) / 32, (maxH + 7) / 8, hb.n_crops * chunks), 256, 0, st>>>(hb);
  }
  hand_gauss_kernel<<<dim3((maxW + kG2W - 1) / kG2W, (maxH + kG2H - 1) / kG2H, hb.n_crops * hb.parts), kG2Threads, 0, st>>>(
      hb, gw, thre);
  hand_select_kernel<<<hb.n_crops * hb.parts, 1024, 0, st>>>(hb);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace islpose
