// Body grouping on the device (reference: src/body.py:128-235), deterministic by construction:
//   paf_score   body.py:142-164  one warp per (limb, candA i, candB j): 10 line samples x 2 PAF channels, each
//               sampled lazily from the stride-8 PAF maps through both cubic stages and the scale mean
//               (paf_avg is never materialised), float64 scalar math in the reference's operation order
//   group       body.py:166-231  one CTA per frame: per limb a bitonic sort by (score desc, pair index asc)
//               = Python's stable sorted(reverse=True), greedy one-to-one matching, then the serial person
//               assembly and pruning carried out by one warp (row scans ballot-parallel, order preserved)
#include "prepost.cuh"

namespace islpose {

__global__ void __launch_bounds__(256)
paf_score_kernel(const ScaleSet ss, const LimbTable lt, int H, int W, int parts, double thre2, const GroupBuffers gb) {
  const int n = blockIdx.z;
  const int k = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int pa = lt.a[k], pb = lt.b[k];
  const int nA = gb.counts[n * parts + pa];
  const int nB = gb.counts[n * parts + pb];
  const long long total = static_cast<long long>(nA) * nB;
  const uint32_t* keyA = gb.keys + static_cast<long long>(n * parts + pa) * gb.cap;
  const uint32_t* keyB = gb.keys + static_cast<long long>(n * parts + pb) * gb.cap;
  const int C = ss.channels;
  const long long tail_start = (static_cast<long long>(W) * C) / 4 * 4;
  const float fS = static_cast<float>(ss.count);
  const int slot_base = n * lt.nlimbs + k;

  for (long long pair = blockIdx.x * 8 + warp; pair < total; pair += static_cast<long long>(gridDim.x) * 8) {
    const int i = static_cast<int>(pair / nB);
    const int j = static_cast<int>(pair - static_cast<long long>(i) * nB);
    const uint32_t ka = keyA[i], kb = keyB[j];
    const int ax = ka % W, ay = ka / W, bx = kb % W, by = kb / W;
    const long long dxi = bx - ax, dyi = by - ay;
    double norm = sqrt(static_cast<double>(dxi * dxi + dyi * dyi));
    norm = fmax(0.001, norm);
    const double ux = __ddiv_rn(static_cast<double>(dxi), norm);
    const double uy = __ddiv_rn(static_cast<double>(dyi), norm);

    // lanes 0..9: x channel of sample t, lanes 10..19: y channel
    double val = 0.0;
    if (lane < 20) {
      const int t = lane % 10;
      const int ch = lane < 10 ? lt.cx[k] : lt.cy[k];
      // np.linspace(a, b, 10): t * ((b - a) / 9) + a, last sample forced to b
      const double stepx = __ddiv_rn(static_cast<double>(dxi), 9.0);
      const double stepy = __ddiv_rn(static_cast<double>(dyi), 9.0);
      const double xs = t == 9 ? static_cast<double>(bx) : __dadd_rn(__dmul_rn(static_cast<double>(t), stepx), static_cast<double>(ax));
      const double ys = t == 9 ? static_cast<double>(by) : __dadd_rn(__dmul_rn(static_cast<double>(t), stepy), static_cast<double>(ay));
      const int rx = static_cast<int>(rint(xs));  // int(round()) = round half to even
      const int ry = static_cast<int>(rint(ys));
      const bool tail = static_cast<long long>(rx) * C + ch >= tail_start;
      for (int s = 0; s < ss.count; ++s) {
        const ScaleGeom& g = ss.g[s];
        Axis2 sx, sy;
        make_axis2(rx, g.sx, g.wc, g.gw, sx);
        make_axis2(ry, g.sy, g.hc, g.gh, sy);
        const float v = sample2(g.low + (static_cast<long long>(n) * C + ch) * g.gh * g.gw, g.gw, sx, sy, tail);
        val = __dadd_rn(val, static_cast<double>(__fdiv_rn(v, fS)));  // paf_avg += paf / S  (body.py:81)
      }
    }
    const double vx = __shfl_sync(0xffffffffu, val, lane % 10);
    const double vy = __shfl_sync(0xffffffffu, val, lane % 10 + 10);
    const double mid = __dadd_rn(__dmul_rn(vx, ux), __dmul_rn(vy, uy));
    const unsigned above = __ballot_sync(0xffffffffu, lane < 10 && mid > thre2);
    double sum = 0.0;  // Python sum(): left to right from 0
#pragma unroll
    for (int t = 0; t < 10; ++t) sum = __dadd_rn(sum, __shfl_sync(0xffffffffu, mid, t));
    if (lane == 0) {
      const double prior = __dadd_rn(__ddiv_rn(sum, 10.0),
                                     fmin(__dsub_rn(__ddiv_rn(__dmul_rn(0.5, static_cast<double>(H)), norm), 1.0), 0.0));
      if (__popc(above) > 8 && prior > 0.0) {  // > 0.8 * mid_num samples above thre2, positive score
        const int slot = atomicAdd(gb.cand_count + slot_base, 1);
        if (slot < gb.cand_cap) {
          gb.cand_pair[static_cast<long long>(slot_base) * gb.cand_cap + slot] = static_cast<uint32_t>(pair);
          gb.cand_score[static_cast<long long>(slot_base) * gb.cand_cap + slot] = prior;
        }
      }
    }
  }
}

constexpr int kCandCap = 2048;
constexpr int kPeakCap = 1024;

__global__ void __launch_bounds__(256)
group_kernel(const LimbTable lt, int W, const GroupBuffers gb) {
  __shared__ double s_score[kCandCap];
  __shared__ uint32_t s_pair[kCandCap];
  __shared__ uint32_t s_usedA[kPeakCap / 32], s_usedB[kPeakCap / 32];
  __shared__ int s_off[32];
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
    if (run > gb.max_cand) atomicExch(gb.overflow, 2);
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

  // ---- per limb: stable descending sort + greedy one-to-one matching (body.py:166-173)
  for (int k = 0; k < lt.nlimbs; ++k) {
    const int slot = n * lt.nlimbs + k;
    const int nA = gb.counts[n * parts + lt.a[k]];
    const int nB = gb.counts[n * parts + lt.b[k]];
    int m = gb.cand_count[slot];
    if (m > gb.cand_cap) {
      if (threadIdx.x == 0) atomicExch(gb.overflow, 3);
      m = gb.cand_cap;
    }
    __syncthreads();
    if (nA == 0 || nB == 0 || m == 0) {
      if (threadIdx.x == 0) gb.conn_count[slot] = (nA == 0 || nB == 0) ? -1 : 0;  // -1: limb is in special_k
      continue;
    }
    int m2 = 1;
    while (m2 < m) m2 <<= 1;
    for (int i = threadIdx.x; i < m2; i += blockDim.x) {
      s_score[i] = i < m ? gb.cand_score[static_cast<long long>(slot) * gb.cand_cap + i] : -1.0;  // scores are > 0
      s_pair[i] = i < m ? gb.cand_pair[static_cast<long long>(slot) * gb.cand_cap + i] : 0xffffffffu;
    }
    for (int i = threadIdx.x; i < kPeakCap / 32; i += blockDim.x) {
      s_usedA[i] = 0;
      s_usedB[i] = 0;
    }
    __syncthreads();
    for (int size = 2; size <= m2; size <<= 1) {
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int i = threadIdx.x; i < m2; i += blockDim.x) {
          const int j = i ^ stride;
          if (j > i) {
            const bool up = (i & size) == 0;
            const double a = s_score[i], b = s_score[j];
            const uint32_t pa = s_pair[i], pb = s_pair[j];
            // "i before j" order: higher score first, ties by original (i, j) enumeration order
            const bool j_first = (b > a) || (b == a && pb < pa);
            if (j_first == up) {
              s_score[i] = b;
              s_score[j] = a;
              s_pair[i] = pb;
              s_pair[j] = pa;
            }
          }
        }
        __syncthreads();
      }
    }
    if (threadIdx.x == 0) {
      const int limit = nA < nB ? nA : nB;
      int made = 0;
      for (int c = 0; c < m && made < limit; ++c) {
        const uint32_t pr = s_pair[c];
        const int i = pr / nB, j = pr % nB;
        if ((s_usedA[i >> 5] >> (i & 31)) & 1u) continue;
        if ((s_usedB[j >> 5] >> (j & 31)) & 1u) continue;
        s_usedA[i >> 5] |= 1u << (i & 31);
        s_usedB[j >> 5] |= 1u << (j & 31);
        gb.conn_ij[(static_cast<long long>(slot) * gb.cap + made) * 2 + 0] = i;
        gb.conn_ij[(static_cast<long long>(slot) * gb.cap + made) * 2 + 1] = j;
        gb.conn_score[static_cast<long long>(slot) * gb.cap + made] = s_score[c];
        ++made;
      }
      gb.conn_count[slot] = made;
    }
    __syncthreads();
  }
  __threadfence_block();
  __syncthreads();
  if (warp != 0) return;

  // ---- person assembly (body.py:180-225), one warp; `sub` rows live in the output buffer
  double* sub = gb.subset + static_cast<long long>(n) * gb.max_person * cols;
  int rows = 0;
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
      const double partA = static_cast<double>(idA), partB = static_cast<double>(idB);
      // rows that already hold partA at indexA or partB at indexB, first two in row order
      int found = 0, j1 = -1, j2 = -1;
      for (int base = 0; base < rows && found < 2; base += 32) {
        const int r = base + lane;
        const bool hit = r < rows && (sub[r * cols + ia] == partA || sub[r * cols + ib] == partB);
        unsigned mask = __ballot_sync(0xffffffffu, hit);
        while (mask != 0 && found < 2) {
          const int b = __ffs(mask) - 1;
          mask &= mask - 1;
          if (found == 0) j1 = base + b; else j2 = base + b;
          ++found;
        }
      }
      // (a third match would make the reference raise IndexError, body.py:193-197; the first two are used)
      const double sB = cand[(idB < gb.max_cand ? idB : 0) * 4 + 2];
      const double sA = cand[(idA < gb.max_cand ? idA : 0) * 4 + 2];
      if (found == 1) {
        if (lane == 0 && sub[j1 * cols + ib] != partB) {
          sub[j1 * cols + ib] = partB;
          sub[j1 * cols + cols - 1] = __dadd_rn(sub[j1 * cols + cols - 1], 1.0);
          sub[j1 * cols + cols - 2] = __dadd_rn(sub[j1 * cols + cols - 2], __dadd_rn(sB, cs));
        }
      } else if (found == 2) {
        const bool both = lane < parts && sub[j1 * cols + lane] >= 0.0 && sub[j2 * cols + lane] >= 0.0;
        const unsigned overlap = __ballot_sync(0xffffffffu, both);
        if (overlap == 0) {
          if (lane < parts) {
            sub[j1 * cols + lane] = __dadd_rn(sub[j1 * cols + lane], __dadd_rn(sub[j2 * cols + lane], 1.0));
          } else if (lane == parts) {
            sub[j1 * cols + cols - 2] = __dadd_rn(__dadd_rn(sub[j1 * cols + cols - 2], sub[j2 * cols + cols - 2]), cs);
          } else if (lane == parts + 1) {
            sub[j1 * cols + cols - 1] = __dadd_rn(sub[j1 * cols + cols - 1], sub[j2 * cols + cols - 1]);
          }
          __syncwarp();
          if (lane < cols) {  // np.delete(subset, j2, 0): every lane shifts its own column upwards
            for (int r = j2; r < rows - 1; ++r) sub[r * cols + lane] = sub[(r + 1) * cols + lane];
          }
          --rows;
        } else if (lane == 0) {
          sub[j1 * cols + ib] = partB;
          sub[j1 * cols + cols - 1] = __dadd_rn(sub[j1 * cols + cols - 1], 1.0);
          sub[j1 * cols + cols - 2] = __dadd_rn(sub[j1 * cols + cols - 2], __dadd_rn(sB, cs));
        }
      } else if (k < lt.njoint - 2) {
        if (rows < gb.max_person) {
          if (lane < cols) {
            double v = -1.0;
            if (lane == ia) v = partA;
            if (lane == ib) v = partB;
            if (lane == cols - 1) v = 2.0;
            if (lane == cols - 2) v = __dadd_rn(__dadd_rn(__dadd_rn(0.0, sA), sB), cs);
            sub[rows * cols + lane] = v;
          }
          ++rows;
        } else if (lane == 0) {
          atomicExch(gb.overflow, 4);
        }
      }
      __syncwarp();
    }
  }
  // ---- pruning (body.py:227-231), order preserved
  int kept = 0;
  for (int r = 0; r < rows; ++r) {
    const double cnt = sub[r * cols + cols - 1];
    const double sc = sub[r * cols + cols - 2];
    const bool drop = cnt < 4.0 || __ddiv_rn(sc, cnt) < 0.4;
    if (!drop) {
      if (kept != r && lane < cols) sub[kept * cols + lane] = sub[r * cols + lane];
      ++kept;
    }
    __syncwarp();
  }
  if (lane == 0) gb.n_person[n] = kept;
}

int launch_paf_score(const ScaleSet& paf, const LimbTable& lt, int N, int H, int W, double thre2, int mid_num,
                     const GroupBuffers& gb, cudaStream_t st) {
  if (mid_num != 10 || gb.cap > kPeakCap || gb.cand_cap > kCandCap) return 1;
  if (cudaMemsetAsync(gb.cand_count, 0, sizeof(int) * N * lt.nlimbs, st) != cudaSuccess) return 1;
  const dim3 grid(32, lt.nlimbs, N);
  paf_score_kernel<<<grid, 256, 0, st>>>(paf, lt, H, W, lt.njoint - 1, thre2, gb);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int launch_group(const LimbTable& lt, int N, int W, const GroupBuffers& gb, cudaStream_t st) {
  if (gb.cap > kPeakCap || gb.cand_cap > kCandCap || lt.njoint + 1 > 32) return 1;
  group_kernel<<<N, 256, 0, st>>>(lt, W, gb);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace islpose
