// Hand key-point selection (reference: src/hand.py:58-74): per part, threshold the smoothed map, label the
// 8-connected components, keep the component with the largest mass of the *unsmoothed* map, zero the rest
// and take the first arg-max (util.npmax, util.py:394-399). One 1024-thread CTA per (hand, part) plane.
// Labels are (smallest flat pixel index in the component)+1, i.e. the raster order skimage.measure.label
// numbers components in, so "first maximum" ties resolve the same way.
#include "prepost.cuh"

namespace islpose {

namespace {

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

}  // namespace

__global__ void __launch_bounds__(1024)
hand_peaks_kernel(const double* __restrict__ heat, const double* __restrict__ smoothed, int H, int W, double thre,
                  int* __restrict__ labels, double* __restrict__ mass, int32_t* __restrict__ out_xy) {
  __shared__ int s_flag;
  __shared__ Best s_tmp[32];
  const int plane = blockIdx.x;
  const int total = H * W;
  const double* hm = heat + static_cast<long long>(plane) * total;
  const double* sm = smoothed + static_cast<long long>(plane) * total;
  int* lab = labels + static_cast<long long>(plane) * total;
  double* ms = mass + static_cast<long long>(plane) * total;

  if (threadIdx.x == 0) s_flag = 0;
  __syncthreads();
  int any = 0;
  for (int p = threadIdx.x; p < total; p += blockDim.x) {
    const int fg = sm[p] > thre;
    lab[p] = fg ? p + 1 : 0;
    ms[p] = 0.0;
    any |= fg;
  }
  if (any) s_flag = 1;
  __syncthreads();
  if (s_flag == 0) {  // hand.py:64-66: nothing above the threshold
    if (threadIdx.x == 0) {
      out_xy[plane * 2 + 0] = 0;
      out_xy[plane * 2 + 1] = 0;
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
  // component mass of the unsmoothed map
  for (int p = threadIdx.x; p < total; p += blockDim.x) {
    const int l = lab[p];
    if (l != 0) atomicAdd(ms + (l - 1), hm[p]);
  }
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
  const Best comp = block_best(mine, s_tmp);
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
    out_xy[plane * 2 + 0] = top.idx % W;
    out_xy[plane * 2 + 1] = top.idx / W;
  }
}

int launch_hand_peaks(const double* heat, const double* smoothed, int planes_total, int H, int W, double thre,
                      int* labels, double* mass, int32_t* out_xy, cudaStream_t st) {
  hand_peaks_kernel<<<planes_total, 1024, 0, st>>>(heat, smoothed, H, W, thre, labels, mass, out_xy);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace islpose
