// Body grouping on the device (reference: src/body.py:128-235), deterministic by construction:
//   paf_score   body.py:142-164  one warp per (limb, candA i, candB j): 10 line samples x 2 PAF channels, each
//               sampled lazily from the stride-8 PAF maps through both cubic stages and the scale mean
//               (paf_avg is never materialised), float64 scalar math in the reference's operation order
//               and writes the score (or -1 when the pair fails criterion1/criterion2) into a dense nA x nB matrix
//   match       body.py:166-173  one CTA per (frame, limb). Walking the candidates in stable descending score
//               order and taking a pair when both ends are free is the same as repeatedly taking the arg-max over
//               the still-free pairs (ties: lowest i*nB+j = the order the reference enumerated them in), so no
//               sort is needed: per-row best partners are cached and only rows that lose their partner rescan
//   assemble    body.py:180-231  one CTA per frame: candidate table, then the serial person assembly and pruning
//               carried out by one warp (row scans ballot-parallel, order preserved)
#include "prepost.cuh"

namespace islpose {

// paf_avg (never materialised) at one integer frame position for the two PAF channels of a limb: both cubic stages per
// scale, float32 division by the number of scales, float64 accumulation (body.py:81) - the reference's values exactly.
__device__ __forceinline__ void paf_at(const ScaleSet& ss, const float (*s_tab)[4], int n, int rx, int ry, int chx, int chy,
                                       int W, double& vx, double& vy) {
  const int C = ss.channels;
  const long long tail_start = (static_cast<long long>(W) * C) / 4 * 4;
  const float fS = static_cast<float>(ss.count);
  const bool tailx = static_cast<long long>(rx) * C + chx >= tail_start;
  const bool taily = static_cast<long long>(rx) * C + chy >= tail_start;
  vx = 0.0;
  vy = 0.0;
  for (int s = 0; s < ss.count; ++s) {
    const ScaleGeom& g = ss.g[s];
    Axis2 sx, sy;
    make_axis2(rx, g.sx, g.wc, g.gw, s_tab, sx);
    make_axis2(ry, g.sy, g.hc, g.gh, s_tab, sy);
    const float* img = g.low + static_cast<long long>(n) * C * g.gh * g.gw;
    const long long plane = static_cast<long long>(g.gh) * g.gw;
    vx = __dadd_rn(vx, static_cast<double>(__fdiv_rn(sample2(img + chx * plane, g.gw, sx, sy, tailx), fS)));
    vy = __dadd_rn(vy, static_cast<double>(__fdiv_rn(sample2(img + chy * plane, g.gw, sx, sy, taily), fS)));
  }
}

// The first and the last of the 10 line samples of a candidate pair (body.py:148-155) sit ON the two peaks, so their
// PAF vectors depend on the peak and the limb only, not on the pair: nA + nB evaluations per limb instead of
// 2 * nA * nB. end_paf[((n * nlimbs + k) * 2 + side) * cap + peak] = (x component, y component).
__global__ void __launch_bounds__(128)
paf_endpoints_kernel(const ScaleSet ss, const LimbTable lt, int W, int parts, const GroupBuffers gb) {
  __shared__ float s_tab[8][4];
  fill_phase_table(s_tab);
  __syncthreads();
  const int n = blockIdx.z;
  const int k = blockIdx.y >> 1, side = blockIdx.y & 1;
  const int part = side ? lt.b[k] : lt.a[k];
  const int cnt = min(gb.counts[n * parts + part], gb.cap);
  const uint32_t* keys = gb.keys + static_cast<long long>(n * parts + part) * gb.cap;
  double2* out = reinterpret_cast<double2*>(gb.end_paf) + (static_cast<long long>(n * lt.nlimbs + k) * 2 + side) * gb.cap;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
    const uint32_t key = keys[i];
    double vx, vy;
    paf_at(ss, s_tab, n, static_cast<int>(key % W), static_cast<int>(key / W), lt.cx[k], lt.cy[k], W, vx, vy);
    out[i] = make_double2(vx, vy);
  }
}

// One group of 8 lanes per candidate pair: the 8 interior samples (t = 1..8) in parallel, the two end samples from
// end_paf. A pair whose two end samples both fail criterion1's threshold cannot reach the required 9 of 10
// (body.py:158) and is rejected before any interior sample is evaluated - on noisy maps that is most pairs. Every
// group walks its own sequence of pairs and skips the rejected ones, so the four groups of a warp always sample
// four surviving pairs together. Scores are written into the dense nA x nB matrix exactly as before.
__global__ void __launch_bounds__(256)
paf_score_kernel(const ScaleSet ss, const LimbTable lt, int H, int W, int parts, double thre2, const GroupBuffers gb) {
  __shared__ float s_tab[8][4];
  fill_phase_table(s_tab);
  __syncthreads();
  const int n = blockIdx.z;
  const int k = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int pa = lt.a[k], pb = lt.b[k];
  const int nA = gb.counts[n * parts + pa];
  const int nB = gb.counts[n * parts + pb];
  const long long total = static_cast<long long>(nA) * nB;
  if (total > gb.pair_cap) {
    if (threadIdx.x == 0 && blockIdx.x == 0) atomicOr(gb.overflow, kOverflowPairs);
    return;
  }
  if (nA > gb.cap || nB > gb.cap) return;  // peak overflow is flagged by the peak kernel
  const uint32_t* keyA = gb.keys + static_cast<long long>(n * parts + pa) * gb.cap;
  const uint32_t* keyB = gb.keys + static_cast<long long>(n * parts + pb) * gb.cap;
  const int slot_base = n * lt.nlimbs + k;
  const double2* endA = reinterpret_cast<const double2*>(gb.end_paf) + static_cast<long long>(slot_base) * 2 * gb.cap;
  const double2* endB = endA + gb.cap;
  double* scores = gb.pair_score + static_cast<long long>(slot_base) * gb.pair_cap;
  const int chx = lt.cx[k], chy = lt.cy[k];

  const int grp = lane >> 3;        // 4 pairs per warp
  const int t = (lane & 7) + 1;     // this lane's interior sample
  const unsigned gmask = 0xffu << (grp * 8);
  const long long stride = static_cast<long long>(gridDim.x) * 32;
  long long pair = (static_cast<long long>(blockIdx.x) * 8 + warp) * 4 + grp;
  while (true) {
    // ---- next pair of this group that survives the end-sample test
    bool have = false;
    double mid0 = 0.0, mid9 = 0.0, ux = 0.0, uy = 0.0, norm = 1.0;
    int ax = 0, ay = 0, bx = 0, by = 0;
    long long dxi = 0, dyi = 0;
    while (pair < total) {
      const int i = static_cast<int>(pair / nB);
      const int j = static_cast<int>(pair - static_cast<long long>(i) * nB);
      const uint32_t ka = keyA[i], kb = keyB[j];
      ax = ka % W, ay = ka / W, bx = kb % W, by = kb / W;
      dxi = bx - ax, dyi = by - ay;
      norm = fmax(0.001, sqrt(static_cast<double>(dxi * dxi + dyi * dyi)));
      ux = __ddiv_rn(static_cast<double>(dxi), norm);
      uy = __ddiv_rn(static_cast<double>(dyi), norm);
      const double2 ea = endA[i], eb = endB[j];
      mid0 = __dadd_rn(__dmul_rn(ea.x, ux), __dmul_rn(ea.y, uy));
      mid9 = __dadd_rn(__dmul_rn(eb.x, ux), __dmul_rn(eb.y, uy));
      if (mid0 > thre2 || mid9 > thre2) {
        have = true;
        break;
      }
      if (t == 1) scores[pair] = -1.0;  // at most 8 of 10 samples can pass: not a candidate
      pair += stride;
    }
    if (!__any_sync(0xffffffffu, have)) break;
    double mid = 0.0;
    if (have) {
      // np.linspace(a, b, 10): t * ((b - a) / 9) + a for the interior samples
      const double stepx = __ddiv_rn(static_cast<double>(dxi), 9.0);
      const double stepy = __ddiv_rn(static_cast<double>(dyi), 9.0);
      const double xs = __dadd_rn(__dmul_rn(static_cast<double>(t), stepx), static_cast<double>(ax));
      const double ys = __dadd_rn(__dmul_rn(static_cast<double>(t), stepy), static_cast<double>(ay));
      double vx, vy;
      paf_at(ss, s_tab, n, static_cast<int>(rint(xs)), static_cast<int>(rint(ys)), chx, chy, W, vx, vy);  // int(round())
      mid = __dadd_rn(__dmul_rn(vx, ux), __dmul_rn(vy, uy));
    }
    const unsigned above = __ballot_sync(0xffffffffu, have && mid > thre2) & gmask;
    double sum = __dadd_rn(0.0, mid0);  // Python sum(): left to right from 0
#pragma unroll
    for (int q = 0; q < 8; ++q) sum = __dadd_rn(sum, __shfl_sync(0xffffffffu, mid, grp * 8 + q));
    sum = __dadd_rn(sum, mid9);
    if (have && t == 1) {
      const double prior = __dadd_rn(__ddiv_rn(sum, 10.0),
                                     fmin(__dsub_rn(__ddiv_rn(__dmul_rn(0.5, static_cast<double>(H)), norm), 1.0), 0.0));
      const int cnt = __popc(above) + (mid0 > thre2 ? 1 : 0) + (mid9 > thre2 ? 1 : 0);
      // > 0.8 * mid_num samples above thre2 and a positive score, else the pair is not a candidate
      scores[pair] = (cnt > 8 && prior > 0.0) ? prior : -1.0;
    }
    if (have) pair += stride;
  }
}


// Greedy matching in parallel rounds. With a strict total order on the candidate pairs (score descending, then
// i*nB+j ascending = the reference's enumeration order, which Python's stable sort preserves for equal scores),
// "walk the sorted list and take a pair when both ends are free" (body.py:166-173) yields the same set as:
// repeat { accept every free pair that is the best of its row AND the best of its column among the free pairs }.
// (The best free pair overall is always such a pair, an accepted pair can never be blocked by a better one, and
// accepting it blocks exactly the pairs the sequential walk would skip.) The walk stops at min(nA, nB) connections,
// which is also when no free row or no free column is left. Accepted connections are finally sorted into the
// order the walk would have produced them in, because the person assembly consumes them in that order.
// Dynamic shared memory, sized by the peak capacity (match_smem_bytes): 32.25 bytes per peak slot, plus the work area.
// Array lengths are the capacity rounded up to a power of two (the final bitonic sort pads to one).
__host__ __device__ inline int pow2_at_least(int v) {
  int p = 32;
  while (p < v) p <<= 1;
  return p;
}
__host__ __device__ inline size_t match_smem_bytes(int cap) {
  const size_t c = static_cast<size_t>(pow2_at_least(cap));
  return c * 32 + (c / 32 + 1) * 8;
}

// Behind those arrays sits the work area (`work_bytes`): on noisy maps a limb has a few ten thousand candidate pairs of
// which a few thousand carry a positive score, and only those can ever be accepted. They are compacted once into a list
// (score, i, j) in shared memory and the rounds run over the list - three passes of a few entries per thread (row / column
// maxima by 64-bit atomicMax on the score bits, ties to the smallest index by atomicMin, then the mutual-best test) instead
// of two scans of the whole nA x nB matrix in L2 per round. When the list does not fit, the dense rounds below run.
__global__ void __launch_bounds__(256)
match_kernel(const LimbTable lt, const GroupBuffers gb, int work_bytes) {
  extern __shared__ double s_match[];
  const int kcap = pow2_at_least(gb.cap);
  double* const s_rowv = s_match;                 // [cap] best free partner's score per row
  double* const s_cv = s_rowv + kcap;             // [cap] accepted connections: score, i, j
  int* const s_rowj = reinterpret_cast<int*>(s_cv + kcap);
  int* const s_coli = s_rowj + kcap;
  int* const s_ci = s_coli + kcap;
  int* const s_cj = s_ci + kcap;
  uint32_t* const s_usedA = reinterpret_cast<uint32_t*>(s_cj + kcap);
  uint32_t* const s_usedB = s_usedA + (kcap / 32 + 1);
  __shared__ int s_made, s_round, s_nent;
  const int k = blockIdx.x, n = blockIdx.y;
  const int parts = lt.njoint - 1;
  const int slot = n * lt.nlimbs + k;
  const int nA = gb.counts[n * parts + lt.a[k]];
  const int nB = gb.counts[n * parts + lt.b[k]];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (nA == 0 || nB == 0) {
    if (threadIdx.x == 0) gb.conn_count[slot] = -1;  // the limb is in special_k (body.py:174-176)
    return;
  }
  if (static_cast<long long>(nA) * nB > gb.pair_cap) {
    if (threadIdx.x == 0) gb.conn_count[slot] = 0;
    return;
  }
  const double* sc = gb.pair_score + static_cast<long long>(slot) * gb.pair_cap;
  for (int i = threadIdx.x; i < kcap / 32 + 1; i += blockDim.x) {
    s_usedA[i] = 0;
    s_usedB[i] = 0;
  }
  if (threadIdx.x == 0) {
    s_made = 0;
    s_nent = 0;
  }
  __syncthreads();

  // ---- work area: column maxima [kcap] (64-bit), then the list: score bits [list_cap], (i << 16 | j) [list_cap]
  unsigned long long* const s_colmax = reinterpret_cast<unsigned long long*>(s_match + match_smem_bytes(gb.cap) / sizeof(double));
  unsigned long long* const s_rowmax = reinterpret_cast<unsigned long long*>(s_rowv);
  const int list_cap = (work_bytes - kcap * 8) / 12;
  unsigned long long* const s_ev = s_colmax + kcap;
  uint32_t* const s_eij = reinterpret_cast<uint32_t*>(s_ev + (list_cap > 0 ? list_cap : 0));
  bool listed = list_cap > 0 && nA <= 65535 && nB <= 65535;
  if (listed) {
    const int total = nA * nB;
    for (int base = 0; base < total; base += blockDim.x) {
      const int idx = base + threadIdx.x;
      const double v = idx < total ? sc[idx] : -1.0;
      const bool pos = v > 0.0;
      const unsigned m = __ballot_sync(0xffffffffu, pos);
      int wbase = 0;
      if (lane == 0 && m != 0) wbase = atomicAdd(&s_nent, __popc(m));
      wbase = __shfl_sync(0xffffffffu, wbase, 0);
      if (pos) {
        const int at = wbase + __popc(m & ((1u << lane) - 1u));
        if (at < list_cap) {
          const int i = idx / nB;
          s_ev[at] = static_cast<unsigned long long>(__double_as_longlong(v));  // positive doubles order like their bits
          s_eij[at] = (static_cast<uint32_t>(i) << 16) | static_cast<uint32_t>(idx - i * nB);
        }
      }
    }
    __syncthreads();
    listed = s_nent <= list_cap;
  }
  if (listed) {
    const int nent = s_nent;
    while (true) {
      for (int i = threadIdx.x; i < nA; i += blockDim.x) {
        s_rowmax[i] = 0ull;
        s_rowj[i] = 0x7fffffff;
      }
      for (int j = threadIdx.x; j < nB; j += blockDim.x) {
        s_colmax[j] = 0ull;
        s_coli[j] = 0x7fffffff;
      }
      if (threadIdx.x == 0) s_round = 0;
      __syncthreads();
      for (int e = threadIdx.x; e < nent; e += blockDim.x) {  // best free score of every row and column
        const int i = s_eij[e] >> 16, j = s_eij[e] & 0xffffu;
        if (((s_usedA[i >> 5] >> (i & 31)) | (s_usedB[j >> 5] >> (j & 31))) & 1u) continue;
        atomicMax(&s_rowmax[i], s_ev[e]);
        atomicMax(&s_colmax[j], s_ev[e]);
      }
      __syncthreads();
      for (int e = threadIdx.x; e < nent; e += blockDim.x) {  // ties: the smallest j of a row, the smallest i of a column
        const int i = s_eij[e] >> 16, j = s_eij[e] & 0xffffu;
        if (((s_usedA[i >> 5] >> (i & 31)) | (s_usedB[j >> 5] >> (j & 31))) & 1u) continue;
        if (s_ev[e] == s_rowmax[i]) atomicMin(&s_rowj[i], j);
        if (s_ev[e] == s_colmax[j]) atomicMin(&s_coli[j], i);
      }
      __syncthreads();
      for (int e = threadIdx.x; e < nent; e += blockDim.x) {  // a pair that is the best of its row and of its column
        const int i = s_eij[e] >> 16, j = s_eij[e] & 0xffffu;
        if (s_ev[e] == s_rowmax[i] && s_rowj[i] == j && s_ev[e] == s_colmax[j] && s_coli[j] == i &&
            !(((s_usedA[i >> 5] >> (i & 31)) | (s_usedB[j >> 5] >> (j & 31))) & 1u)) {
          const int pos = atomicAdd(&s_made, 1);
          s_ci[pos] = i;
          s_cj[pos] = j;
          s_round = 1;
        }
      }
      __syncthreads();
      // the accepted pairs of this round: mark their ends used (after the pass, so that it saw one state of the flags)
      const int made_now = s_made;
      const bool any = s_round != 0;
      __syncthreads();
      if (!any) break;
      for (int c = threadIdx.x; c < made_now; c += blockDim.x) {
        atomicOr(&s_usedA[s_ci[c] >> 5], 1u << (s_ci[c] & 31));
        atomicOr(&s_usedB[s_cj[c] >> 5], 1u << (s_cj[c] & 31));
      }
      __syncthreads();
    }
    // scores of the accepted pairs (the row maxima array doubles as s_rowv and is dead now)
    for (int c = threadIdx.x; c < s_made; c += blockDim.x) s_cv[c] = sc[static_cast<long long>(s_ci[c]) * nB + s_cj[c]];
    __syncthreads();
  } else
  while (true) {
    // best free column of every free row (one warp per row, lanes across columns: coalesced)
    for (int i = warp; i < nA; i += 8) {
      if ((s_usedA[i >> 5] >> (i & 31)) & 1u) continue;
      double bv = -1.0;
      int bj = 0x7fffffff;
      const double* row = sc + static_cast<long long>(i) * nB;
      for (int j = lane; j < nB; j += 32) {
        if ((s_usedB[j >> 5] >> (j & 31)) & 1u) continue;
        const double v = row[j];
        if (v > bv) {  // ascending j inside a lane: strict > keeps the smallest j among equals
          bv = v;
          bj = j;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
        if (ov > bv || (ov == bv && oj < bj)) {
          bv = ov;
          bj = oj;
        }
      }
      if (lane == 0) {
        s_rowv[i] = bv;
        s_rowj[i] = bj;
      }
    }
    // best free row of every free column (one thread per column, adjacent lanes read adjacent addresses)
    for (int j = threadIdx.x; j < nB; j += blockDim.x) {
      int bi = -1;
      if (!((s_usedB[j >> 5] >> (j & 31)) & 1u)) {
        double bv = -1.0;
        for (int i = 0; i < nA; ++i) {
          if ((s_usedA[i >> 5] >> (i & 31)) & 1u) continue;
          const double v = sc[static_cast<long long>(i) * nB + j];
          if (v > bv) {  // ascending i: strict > keeps the smallest i among equals
            bv = v;
            bi = i;
          }
        }
        if (!(bv > 0.0)) bi = -1;
      }
      s_coli[j] = bi;
    }
    if (threadIdx.x == 0) s_round = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < nA; i += blockDim.x) {
      if ((s_usedA[i >> 5] >> (i & 31)) & 1u) continue;
      const double v = s_rowv[i];
      const int j = s_rowj[i];
      if (v > 0.0 && s_coli[j] == i) {
        const int pos = atomicAdd(&s_made, 1);
        s_cv[pos] = v;
        s_ci[pos] = i;
        s_cj[pos] = j;
        atomicOr(&s_usedA[i >> 5], 1u << (i & 31));
        atomicOr(&s_usedB[j >> 5], 1u << (j & 31));
        s_round = 1;
      }
    }
    __syncthreads();
    if (s_round == 0) break;
    __syncthreads();
  }
  // order of the sequential walk: score descending, pair index ascending
  const int made = s_made;
  int m2 = 1;
  while (m2 < made) m2 <<= 1;
  for (int i = made + threadIdx.x; i < m2; i += blockDim.x) {
    s_cv[i] = -1.0;
    s_ci[i] = 0x7fffffff;
    s_cj[i] = 0;
  }
  __syncthreads();
  for (int size = 2; size <= m2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < m2; i += blockDim.x) {
        const int j = i ^ stride;
        if (j > i) {
          const bool up = (i & size) == 0;
          const double a = s_cv[i], b = s_cv[j];
          const long long pa = static_cast<long long>(s_ci[i]) * nB + s_cj[i];
          const long long pb = static_cast<long long>(s_ci[j]) * nB + s_cj[j];
          const bool j_first = (b > a) || (b == a && pb < pa);
          if (j_first == up) {
            s_cv[i] = b;
            s_cv[j] = a;
            const int ti = s_ci[i], tj = s_cj[i];
            s_ci[i] = s_ci[j];
            s_cj[i] = s_cj[j];
            s_ci[j] = ti;
            s_cj[j] = tj;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int c = threadIdx.x; c < made; c += blockDim.x) {
    gb.conn_ij[(static_cast<long long>(slot) * gb.cap + c) * 2 + 0] = s_ci[c];
    gb.conn_ij[(static_cast<long long>(slot) * gb.cap + c) * 2 + 1] = s_cj[c];
    gb.conn_score[static_cast<long long>(slot) * gb.cap + c] = s_cv[c];
  }
  if (threadIdx.x == 0) gb.conn_count[slot] = made;
}

constexpr int kMaxSlots = 65536;  // row slots an assembly may create (dead ones included)

__global__ void __launch_bounds__(256)
assemble_kernel(const LimbTable lt, int W, const GroupBuffers gb) {
  __shared__ int s_off[32];
  __shared__ uint32_t s_alive[kMaxSlots / 32];
  const int n = blockIdx.x;
  const int parts = lt.njoint - 1;
  const int cols = lt.njoint + 1;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;

  // ---- candidate table: all peaks, ids run across parts (body.py:101-107,183)
  if (threadIdx.x == 0) {
    int run = 0;
    for (int p = 0; p < parts; ++p) {
      s_off[p] = run;
      run += gb.counts[n * parts + p];
    }
    s_off[parts] = run;
    gb.n_cand[n] = run < gb.max_cand ? run : gb.max_cand;
    if (run > gb.max_cand) atomicOr(gb.overflow, kOverflowCandidates);
  }
  __syncthreads();
  double* cand = gb.candidate + static_cast<long long>(n) * gb.max_cand * 4;
  for (int p = 0; p < parts; ++p) {
    const int cnt = gb.counts[n * parts + p];
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
      const int row = s_off[p] + i;
      if (row < gb.max_cand) {
        const uint32_t key = gb.keys[static_cast<long long>(n * parts + p) * gb.cap + i];
        cand[row * 4 + 0] = static_cast<double>(key % W);
        cand[row * 4 + 1] = static_cast<double>(key / W);
        cand[row * 4 + 2] = gb.scores[static_cast<long long>(n * parts + p) * gb.cap + i];
        cand[row * 4 + 3] = static_cast<double>(row);
      }
    }
  }
  // owner[id] = the (at most two) row slots that currently hold candidate `id`. Every id belongs to one part,
  // i.e. one column, and the reference's own two-slot subset_idx (body.py:193-197) implies an id never sits in
  // more than two rows, so "rows with partA at indexA or partB at indexB" is a lookup instead of a scan.
  int* owner = gb.owner + static_cast<long long>(n) * gb.max_cand * 2;
  const int ncand = s_off[parts] < gb.max_cand ? s_off[parts] : gb.max_cand;
  for (int i = threadIdx.x; i < ncand * 2; i += blockDim.x) owner[i] = -1;
  for (int i = threadIdx.x; i < kMaxSlots / 32; i += blockDim.x) s_alive[i] = 0;
  __threadfence_block();
  __syncthreads();
  if (warp != 0) return;

  // ---- person assembly (body.py:180-225), one warp. Rows live in stable slots of the output buffer in creation
  // order; np.delete marks a slot dead, so "row order" is slot order and nothing is shifted until the end.
  double* sub = gb.subset + static_cast<long long>(n) * gb.max_person * cols;
  int slots = 0;
  auto owner_add = [&](int id, int slot) {
    if (owner[id * 2] == slot || owner[id * 2 + 1] == slot) return;
    if (owner[id * 2] < 0) owner[id * 2] = slot; else if (owner[id * 2 + 1] < 0) owner[id * 2 + 1] = slot;
  };
  auto owner_remove = [&](int id, int slot) {
    if (owner[id * 2] == slot) { owner[id * 2] = owner[id * 2 + 1]; owner[id * 2 + 1] = -1; }
    else if (owner[id * 2 + 1] == slot) owner[id * 2 + 1] = -1;
  };
  for (int k = 0; k < lt.nlimbs; ++k) {
    const int slot = n * lt.nlimbs + k;
    const int cc = gb.conn_count[slot];
    if (cc < 0) continue;
    const int ia = lt.a[k], ib = lt.b[k];
    for (int c = 0; c < cc; ++c) {
      const int ci = gb.conn_ij[(static_cast<long long>(slot) * gb.cap + c) * 2 + 0];
      const int cj = gb.conn_ij[(static_cast<long long>(slot) * gb.cap + c) * 2 + 1];
      const double cs = gb.conn_score[static_cast<long long>(slot) * gb.cap + c];
      const int idA = s_off[ia] + ci, idB = s_off[ib] + cj;
      if (idA >= ncand || idB >= ncand) continue;  // only after a candidate-table overflow (already flagged)
      const double partA = static_cast<double>(idA), partB = static_cast<double>(idB);
      // rows holding partA at indexA or partB at indexB: the two smallest slots (a third would make the reference
      // raise IndexError, body.py:193-197; the first two are used here)
      int o[4] = {owner[idA * 2], owner[idA * 2 + 1], owner[idB * 2], owner[idB * 2 + 1]};
      int j1 = 0x7fffffff, j2 = 0x7fffffff;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int v = o[q];
        if (v < 0 || v == j1 || v == j2) continue;
        if (v < j1) { j2 = j1; j1 = v; } else if (v < j2) { j2 = v; }
      }
      const int found = (j1 != 0x7fffffff) + (j2 != 0x7fffffff);
      const double sB = cand[idB * 4 + 2];
      const double sA = cand[idA * 4 + 2];
      if (found == 1) {
        if (lane == 0 && sub[j1 * cols + ib] != partB) {
          const double old = sub[j1 * cols + ib];
          if (old >= 0.0) owner_remove(static_cast<int>(old), j1);
          owner_add(idB, j1);
          sub[j1 * cols + ib] = partB;
          sub[j1 * cols + cols - 1] = __dadd_rn(sub[j1 * cols + cols - 1], 1.0);
          sub[j1 * cols + cols - 2] = __dadd_rn(sub[j1 * cols + cols - 2], __dadd_rn(sB, cs));
        }
      } else if (found == 2) {
        const bool both = lane < parts && sub[j1 * cols + lane] >= 0.0 && sub[j2 * cols + lane] >= 0.0;
        const unsigned overlap = __ballot_sync(0xffffffffu, both);
        if (overlap == 0) {
          if (lane < parts) {
            const double moved = sub[j2 * cols + lane];
            if (moved >= 0.0) {  // this id now lives in j1 instead of j2
              const int id = static_cast<int>(moved);
              if (owner[id * 2] == j2) owner[id * 2] = j1; else if (owner[id * 2 + 1] == j2) owner[id * 2 + 1] = j1;
            }
            sub[j1 * cols + lane] = __dadd_rn(sub[j1 * cols + lane], __dadd_rn(moved, 1.0));
          } else if (lane == parts) {
            sub[j1 * cols + cols - 2] = __dadd_rn(__dadd_rn(sub[j1 * cols + cols - 2], sub[j2 * cols + cols - 2]), cs);
          } else if (lane == parts + 1) {
            sub[j1 * cols + cols - 1] = __dadd_rn(sub[j1 * cols + cols - 1], sub[j2 * cols + cols - 1]);
          }
          if (lane == 0) s_alive[j2 >> 5] &= ~(1u << (j2 & 31));  // np.delete(subset, j2, 0)
        } else if (lane == 0) {
          const double old = sub[j1 * cols + ib];
          if (old != partB) {
            if (old >= 0.0) owner_remove(static_cast<int>(old), j1);
            owner_add(idB, j1);
          }
          sub[j1 * cols + ib] = partB;
          sub[j1 * cols + cols - 1] = __dadd_rn(sub[j1 * cols + cols - 1], 1.0);
          sub[j1 * cols + cols - 2] = __dadd_rn(sub[j1 * cols + cols - 2], __dadd_rn(sB, cs));
        }
      } else if (k < lt.njoint - 2) {
        if (slots < gb.max_person && slots < kMaxSlots) {
          if (lane < cols) {
            double v = -1.0;
            if (lane == ia) v = partA;
            if (lane == ib) v = partB;
            if (lane == cols - 1) v = 2.0;
            if (lane == cols - 2) v = __dadd_rn(__dadd_rn(__dadd_rn(0.0, sA), sB), cs);
            sub[slots * cols + lane] = v;
          }
          if (lane == 0) {
            owner_add(idA, slots);
            owner_add(idB, slots);
            s_alive[slots >> 5] |= 1u << (slots & 31);
          }
          ++slots;
        } else if (lane == 0) {
          atomicOr(gb.overflow, kOverflowPersons);
        }
      }
      __syncwarp();
    }
  }
  // ---- np.delete + pruning (body.py:227-231): compact the live rows that pass, order preserved
  int kept = 0;
  for (int r = 0; r < slots; ++r) {
    const bool alive = (s_alive[r >> 5] >> (r & 31)) & 1u;
    const double cnt = sub[r * cols + cols - 1];
    const double sc = sub[r * cols + cols - 2];
    const bool keep = alive && !(cnt < 4.0 || __ddiv_rn(sc, cnt) < 0.4);
    if (keep) {
      if (kept != r && lane < cols) sub[kept * cols + lane] = sub[r * cols + lane];
      ++kept;
    }
    __syncwarp();
  }
  if (lane == 0) gb.n_person[n] = kept;
}

int launch_paf_score(const ScaleSet& paf, const LimbTable& lt, int N, int H, int W, double thre2, int mid_num,
                     const GroupBuffers& gb, cudaStream_t st) {
  if (mid_num != 10 || gb.cap > kMaxPeakCap) return 1;
  if (gb.end_paf == nullptr) return 1;
  paf_endpoints_kernel<<<dim3(8, lt.nlimbs * 2, N), 128, 0, st>>>(paf, lt, W, lt.njoint - 1, gb);
  // CTAs per (limb, frame): on a single frame the pairs of a few limbs are all the work there is - spread them wider
  const dim3 grid(N <= 2 ? 144 : (N <= 4 ? 96 : 48), lt.nlimbs, N);
  paf_score_kernel<<<grid, 256, 0, st>>>(paf, lt, H, W, lt.njoint - 1, thre2, gb);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int launch_group(const LimbTable& lt, int N, int W, const GroupBuffers& gb, cudaStream_t st) {
  if (gb.cap > kMaxPeakCap || lt.njoint + 1 > 32) return 1;
  const size_t smem = match_smem_bytes(gb.cap);
  // behind the per-peak arrays: the work area of the list-based rounds (column maxima + up to 12288 listed pairs), as far
  // as the 227 KB of an SM allow
  const size_t limit = 220 * 1024;
  if (smem > limit) return 1;
  const size_t kcap = static_cast<size_t>(pow2_at_least(gb.cap));
  size_t work = kcap * 8 + 12 * 12288;
  if (work > limit - smem) work = limit - smem;
  const int work_bytes = static_cast<int>(work / 8 * 8);
  {  // opt in to large dynamic shared memory once per device
    static bool done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!done[dev & 63]) {
      if (cudaFuncSetAttribute(match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(limit)) != cudaSuccess)
        return 1;
      done[dev & 63] = true;
    }
  }
  match_kernel<<<dim3(lt.nlimbs, N), 256, smem + static_cast<size_t>(work_bytes), st>>>(lt, gb, work_bytes);
  assemble_kernel<<<N, 256, 0, st>>>(lt, W, gb);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace islpose
