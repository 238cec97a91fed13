// Memory-bound kernels either side of the networks:
//   resize_pad_norm   body.py:53-56 / hand.py:37-40   uint8 cubic resize + pad(128) + x/256-0.5 -> fp32 NCHW
//   heat_accumulate   body.py:69-72,80 / hand.py:51-56  two cubic resizes + scale accumulation, float64 planes
//   gauss_nms         body.py:88-107  scipy gaussian_filter(sigma=3) (gauss.cuh) + 4-neighbour NMS + threshold -> peak lists
//   sort_peaks        np.nonzero row-major order for the atomically appended peak lists
#include "prepost.cuh"

#include <cuda.h>


#include "cubic.cuh"
#include "ptx.cuh"

namespace islpose {

// ------------------------------------------------------------------------------------------------ resize
__global__ void resize_pad_norm_kernel(const uint8_t* __restrict__ in, int N, int H, int W, int rh, int rw, int hp,
                                       int wp, double scale_x, double scale_y, float* __restrict__ out_nchw,
                                       uint8_t* __restrict__ out_u8) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int n = blockIdx.z;
  if (x >= wp || y >= hp) return;
  int v[3] = {128, 128, 128};  // padValue (body.py:43)
  if (x < rw && y < rh) {
    int sx, sy;
    float fx, fy;
    float cx[4], cy[4];
    cubic_src(x, scale_x, sx, fx);
    cubic_coeffs(fx, cx);
    cubic_src(y, scale_y, sy, fy);
    cubic_coeffs(fy, cy);
    int ia[4];
    float bf[4];
    int xi[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      // saturate_cast<short>(w * INTER_RESIZE_COEF_SCALE): round half to even, 11 fractional bits
      ia[k] = clampi(__float2int_rn(__fmul_rn(cx[k], 2048.f)), -32768, 32767);
      const int ib = clampi(__float2int_rn(__fmul_rn(cy[k], 2048.f)), -32768, 32767);
      bf[k] = __fmul_rn(static_cast<float>(ib), 1.f / 4194304.f);
      xi[k] = clampi(sx - 1 + k, 0, W - 1);
    }
    const uint8_t* img = in + static_cast<long long>(n) * H * W * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float hs[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint8_t* row = img + static_cast<long long>(clampi(sy - 1 + k, 0, H - 1)) * W * 3 + c;
        const int acc = row[xi[0] * 3] * ia[0] + row[xi[1] * 3] * ia[1] + row[xi[2] * 3] * ia[2] + row[xi[3] * 3] * ia[3];
        hs[k] = static_cast<float>(acc);
      }
      // vertical pass in float, SIMD order, round half to even, saturate to uint8
      v[c] = clampi(__float2int_rn(dot4_rl(hs[0], hs[1], hs[2], hs[3], bf)), 0, 255);
    }
  }
  const long long plane = static_cast<long long>(hp) * wp;
  const long long pix = static_cast<long long>(y) * wp + x;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    out_nchw[(static_cast<long long>(n) * 3 + c) * plane + pix] = __fsub_rn(__fdiv_rn(static_cast<float>(v[c]), 256.f), 0.5f);
    if (out_u8 != nullptr) out_u8[(static_cast<long long>(n) * plane + pix) * 3 + c] = static_cast<uint8_t>(v[c]);
  }
}

// ------------------------------------------------------------------------------------------------ accumulate
// One thread = one frame pixel x kChunk channels. For every scale: both cubic stages from the stride-8 map,
// the float32 division by the number of scales, then the float64 accumulation of the reference
// (body.py:80 `avg += avg + m/S` when q1, body.py:81 / hand.py:56 `avg += m/S` otherwise).
constexpr int kChunk = 4;
__global__ void __launch_bounds__(256)
heat_accumulate_kernel(const ScaleSet ss, int N, int H, int W, int parts, int q1, double* __restrict__ out) {
  __shared__ float s_tab[8][4];
  fill_phase_table(s_tab);
  __syncthreads();
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int chunks = (parts + kChunk - 1) / kChunk;
  const int n = blockIdx.z / chunks;
  const int c0 = (blockIdx.z % chunks) * kChunk;
  if (x >= W || y >= H) return;
  double acc[kChunk];
#pragma unroll
  for (int i = 0; i < kChunk; ++i) acc[i] = 0.0;
  const int C = ss.channels;
  const long long tail_start = (static_cast<long long>(W) * C) / 4 * 4;
  const float fS = static_cast<float>(ss.count);
  for (int s = 0; s < ss.count; ++s) {
    const ScaleGeom& g = ss.g[s];
    Axis2 ax, ay;
    make_axis2(x, g.sx, g.wc, g.gw, s_tab, ax);
    make_axis2(y, g.sy, g.hc, g.gh, s_tab, ay);
    const long long plane = static_cast<long long>(g.gh) * g.gw;
#pragma unroll
    for (int i = 0; i < kChunk; ++i) {
      const int c = c0 + i;
      if (c < parts) {
        const bool tail = static_cast<long long>(x) * C + c >= tail_start;
        const float v = sample2(g.low + (static_cast<long long>(n) * C + c) * plane, g.gw, ax, ay, tail);
        const double t = static_cast<double>(__fdiv_rn(v, fS));
        acc[i] = q1 ? __dadd_rn(acc[i], __dadd_rn(acc[i], t)) : __dadd_rn(acc[i], t);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kChunk; ++i) {
    const int c = c0 + i;
    if (c < parts) out[((static_cast<long long>(n) * parts + c) * H + y) * W + x] = acc[i];
  }
}

// Two-pass variant of the same arithmetic (bit-identical results): pass 1 materialises stage 1 - the x8 up-sampled,
// cropped map of one scale, float32 [N][parts][hc][wc]; pass 2 is stage 2 (4x4 taps from that map) plus the division
// and the float64 accumulation. Both passes were instruction-issue bound in their first form (one thread per value:
// ~130 / ~174 warp instructions per value, mostly index arithmetic, profiles/r1_ncu_full_post.txt), so both now share
// work between neighbouring values instead:
//   pass 1: one thread = one 8 x 8 block of up-sampled values. Up-sampled columns 8b-4 .. 8b+3 read the same four
//           source columns (and rows likewise), so the block needs ONE 4 x 4 source window, 32 horizontal dot
//           products (4 source rows x 8 columns) and 64 vertical ones: 10.5 multiply-adds per value instead of 35,
//           and no per-value index arithmetic. The weights of column/row i of a block are phase (i+4)&7, a constant.
//   pass 2: one CTA = one 32 x 32 tile of the frame. Per (scale, channel) the horizontal dot products of every
//           source row the tile needs (<= 37) go through shared memory and are shared by the rows that use them;
//           cubic coefficients and tap indices are computed once per thread and scale, not per value.
__global__ void __launch_bounds__(256)
upsample8_kernel(const float* __restrict__ low, int C, int parts, int gh, int gw, int hc, int wc, int pitch,
                 float* __restrict__ mid) {
  __shared__ float s_tab[8][4];
  fill_phase_table(s_tab);
  __syncthreads();
  const int bx = blockIdx.x * 32 + (threadIdx.x & 31);
  const int by = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int n = blockIdx.z / parts;
  const int c = blockIdx.z - n * parts;
  const int u0 = 8 * bx - 4, v0 = 8 * by - 4;
  if (u0 >= wc || v0 >= hc) return;
  const float* plane = low + (static_cast<long long>(n) * C + c) * gh * gw;
  // the 4 x 4 source window: first tap = floor(src) - 1 = b - 2, replicate border
  float L[4][4];
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    const float* row = plane + clampi(by - 2 + l, 0, gh - 1) * gw;
#pragma unroll
    for (int m = 0; m < 4; ++m) L[l][m] = __ldg(row + clampi(bx - 2 + m, 0, gw - 1));
  }
  float w[8][4];  // weights of block column / row i: phase (i + 4) & 7
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int k = 0; k < 4; ++k) w[i][k] = s_tab[(i + 4) & 7][k];
  }
  float T1[4][8];
#pragma unroll
  for (int l = 0; l < 4; ++l) {
#pragma unroll
    for (int i = 0; i < 8; ++i) T1[l][i] = dot4_lr(L[l][0], L[l][1], L[l][2], L[l][3], w[i]);
  }
  // rows of `mid` are `pitch` floats apart (wc rounded up to 4), so both halves of a block row are 16-byte aligned;
  // the pad columns may receive values, nobody reads them
  float* out = mid + (static_cast<long long>(n) * parts + c) * hc * pitch;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int v = v0 + j;
    if (v < 0 || v >= hc) continue;
    float* orow = out + static_cast<long long>(v) * pitch;
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = dot4_rl(T1[0][i], T1[1][i], T1[2][i], T1[3][i], w[j]);
    if (u0 >= 0) *reinterpret_cast<float4*>(orow + u0) = make_float4(o[0], o[1], o[2], o[3]);  // u0 + 3 < wc holds: u0 < wc, both = 0 mod 4 ... see launcher
    if (u0 + 7 < pitch) {
      *reinterpret_cast<float4*>(orow + u0 + 4) = make_float4(o[4], o[5], o[6], o[7]);
    } else {
#pragma unroll
      for (int i = 4; i < 8; ++i) {
        if (u0 + i < wc) orow[u0 + i] = o[i];
      }
    }
  }
}

struct MidSet {
  const float* mid[kMaxScales];  // [N][parts][hc][pitch] per scale
  int pitch[kMaxScales];         // wc rounded up to a multiple of 4 floats
};

constexpr int kRT = 32;        // frame tile edge of pass 2
constexpr int kRMaxRows = 180; // source rows a tile may need: 32 * (hc / H) + 4; 2 x 180 x 33 floats = 47.5 KB of shared memory
constexpr int kRChunk = 5;     // channels per CTA (accumulators stay in registers)

__global__ void __launch_bounds__(256, 2)
resize_accumulate_kernel(const ScaleSet ss, const MidSet ms, int N, int H, int W, int parts, int q1, int rows_cap,
                         double* __restrict__ out) {
  extern __shared__ float s_dyn[];  // [2][rows_cap][kRT + 1]
  float (*s_t0)[kRT + 1] = reinterpret_cast<float (*)[kRT + 1]>(s_dyn);
  float (*s_t1)[kRT + 1] = s_t0 + rows_cap;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int x = blockIdx.x * kRT + tx;
  const int xc = x < W ? x : W - 1;  // threads beyond the frame compute a valid column and store nothing
  const int y0 = blockIdx.y * kRT;
  const int chunks = (parts + kRChunk - 1) / kRChunk;
  const int n = blockIdx.z / chunks;
  const int c0 = (blockIdx.z - n * chunks) * kRChunk;
  const int nch = parts - c0 < kRChunk ? parts - c0 : kRChunk;
  double acc[4][kRChunk];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int i = 0; i < kRChunk; ++i) acc[k][i] = 0.0;
  }
  const int C = ss.channels;
  const long long tail_start = (static_cast<long long>(W) * C) / 4 * 4;
  const float fS = static_cast<float>(ss.count);
  int buf = 0;
  for (int s = 0; s < ss.count; ++s) {
    const ScaleGeom& g = ss.g[s];
    // this thread's column: stage-2 taps and weights (once per scale)
    int sx;
    float fx, wx[4];
    cubic_src(xc, g.sx, sx, fx);
    cubic_coeffs(fx, wx);
    int xi[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) xi[k] = clampi(sx - 1 + k, 0, g.wc - 1);
    // source rows of the tile: first tap of the first row .. last tap of the last row (clamped, monotonic in y)
    int sy_first, sy_last;
    float fdummy;
    cubic_src(y0, g.sy, sy_first, fdummy);
    const int y_last = y0 + kRT - 1 < H ? y0 + kRT - 1 : H - 1;
    cubic_src(y_last, g.sy, sy_last, fdummy);
    const int r_lo = clampi(sy_first - 1, 0, g.hc - 1);
    const int r_hi = clampi(sy_last + 2, 0, g.hc - 1);
    const int nrows = r_hi - r_lo + 1;  // <= rows_cap (sized by the launcher)
    // this thread's four rows: weights and offsets into the shared rows
    float wy[4][4];
    unsigned ro4[4];  // shared-row index of each of the four taps of this thread's four rows, one byte each (< kRMaxRows)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int y = y0 + ty + 8 * k;
      const int yc = y < H ? y : H - 1;
      int sy;
      float fy;
      cubic_src(yc, g.sy, sy, fy);
      cubic_coeffs(fy, wy[k]);
      ro4[k] = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) ro4[k] |= static_cast<unsigned>(clampi(sy - 1 + j, 0, g.hc - 1) - r_lo) << (8 * j);
    }
    const bool x_interior = xi[3] == xi[0] + 3;
    const int pitch = ms.pitch[s];
    const long long plane = static_cast<long long>(g.hc) * pitch;
    // v / S in float32 (body.py:80-81): for a power of two the multiplication by 1/S gives the identical result
    const bool pow2 = (ss.count & (ss.count - 1)) == 0;
    const float rS = 1.0f / fS;
#pragma unroll
    for (int i = 0; i < kRChunk; ++i) {
      if (i < nch) {
        const int c = c0 + i;
        const float* img = ms.mid[s] + (static_cast<long long>(n) * parts + c) * plane + static_cast<long long>(r_lo) * pitch;
        float (*s_t)[kRT + 1] = buf ? s_t1 : s_t0;
        // horizontal pass: one dot product per (source row, column); rows strided over the 8 warps, five rows (20
        // independent loads) in flight per thread. 32-bit offsets from the (CTA-uniform) plane pointer; away from the
        // left / right border the four taps are consecutive floats.
        for (int rb = ty; rb < nrows; rb += 40) {
          float v[5][4];
#pragma unroll
          for (int m = 0; m < 5; ++m) {
            const unsigned r = rb + 8 * m < nrows ? rb + 8 * m : nrows - 1;
            const unsigned off = r * static_cast<unsigned>(pitch);
            if (x_interior) {
              const float* p4 = img + (off + static_cast<unsigned>(xi[0]));
#pragma unroll
              for (int k = 0; k < 4; ++k) v[m][k] = __ldg(p4 + k);
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k) v[m][k] = __ldg(img + (off + static_cast<unsigned>(xi[k])));
            }
          }
#pragma unroll
          for (int m = 0; m < 5; ++m) {
            if (rb + 8 * m < nrows) s_t[rb + 8 * m][tx] = dot4_lr(v[m][0], v[m][1], v[m][2], v[m][3], wx);
          }
        }
        __syncthreads();
        const bool tail = static_cast<long long>(xc) * C + c >= tail_start;
        const float* col = &s_t[0][tx];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // shared rows of this output row's four taps: byte offsets packed per scale (ro4), pitch kRT + 1 floats
          const float t0 = col[(ro4[k] & 0xffu) * (kRT + 1)], t1 = col[((ro4[k] >> 8) & 0xffu) * (kRT + 1)],
                      t2 = col[((ro4[k] >> 16) & 0xffu) * (kRT + 1)], t3 = col[(ro4[k] >> 24) * (kRT + 1)];
          const float v = tail ? dot4_lr(t0, t1, t2, t3, wy[k]) : dot4_rl(t0, t1, t2, t3, wy[k]);
          const double t = static_cast<double>(pow2 ? __fmul_rn(v, rS) : __fdiv_rn(v, fS));
          acc[k][i] = q1 ? __dadd_rn(acc[k][i], __dadd_rn(acc[k][i], t)) : __dadd_rn(acc[k][i], t);
        }
        buf ^= 1;  // the next horizontal pass writes the other buffer: one barrier per (scale, channel)
      }
    }
  }
  if (x < W) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int y = y0 + ty + 8 * k;
      if (y < H) {
#pragma unroll
        for (int i = 0; i < kRChunk; ++i) {
          if (i < nch) out[((static_cast<long long>(n) * parts + c0 + i) * H + y) * W + x] = acc[k][i];
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ accumulate, TMA-fed
// The same second stage (bit-identical arithmetic), with the source rows of every (scale, channel) brought into shared
// memory by TMA instead of per-thread loads. ncu on the kernel above (profiles/r2_ncu_full_other_kernels.txt):
// `long_scoreboard` is the top stall, 124 registers hold 20 loads in flight per thread, occupancy 25 %, issue-active 52 %,
// DRAM 11 % - memory latency that a CTA serialised by one barrier per iteration cannot hide. Here one thread issues, two
// iterations ahead, a cp.async.bulk.tensor of the (rows_cap x cols_cap) window of the up-sampled map that the tile's
// 32 x 32 outputs can touch (replicate borders = clamped indices, which stay inside the window; what TMA zero-fills beyond
// the map is never read); the horizontal pass reads its four taps from that window, everything else is as above.
constexpr int kTmaRing = 3;
struct ResizeMaps {
  CUtensorMap m[kMaxScales];  // per scale: (pitch, hc, N * parts) float32, box (cols_cap, rows_cap, 1)
};
struct TileWindow {
  int r_lo, c_lo, nrows;
};

// three CTAs per SM (80 registers, 64 bytes spilled) measured 3 % faster than two at 1280 x 720 and equal at 640 x 480
__global__ void __launch_bounds__(256, 3)
resize_accumulate_tma_kernel(const __grid_constant__ ResizeMaps tms, const ScaleSet ss, int N, int H, int W, int parts, int q1,
                             int rows_cap, int cols_cap, double* __restrict__ out) {
  extern __shared__ __align__(128) uint8_t s_dyn_raw[];
  __shared__ TileWindow s_win[kMaxScales];
  __shared__ __align__(8) unsigned long long s_bars[kTmaRing];
  const uint32_t buf_bytes = (static_cast<uint32_t>(rows_cap) * cols_cap * 4u + 127u) & ~127u;
  const uint32_t box_bytes = static_cast<uint32_t>(rows_cap) * cols_cap * 4u;
  // TMA destinations must be 128-byte aligned: align by hand (the launcher adds 128 bytes), as the conv kernels do
  uint8_t* const s_dyn = s_dyn_raw + ((128u - (ptx::smem_u32(s_dyn_raw) & 127u)) & 127u);
  const float* const s_src = reinterpret_cast<const float*>(s_dyn);
  float (*s_t0)[kRT + 1] = reinterpret_cast<float (*)[kRT + 1]>(s_dyn + kTmaRing * buf_bytes);
  float (*s_t1)[kRT + 1] = s_t0 + rows_cap;
  const uint32_t src0 = ptx::smem_u32(s_dyn);
  const uint32_t bar0 = ptx::smem_u32(s_bars);

  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int x0 = blockIdx.x * kRT;
  const int x = x0 + tx;
  const int xc = x < W ? x : W - 1;  // threads beyond the frame compute a valid column and store nothing
  const int y0 = blockIdx.y * kRT;
  const int chunks = (parts + kRChunk - 1) / kRChunk;
  const int n = blockIdx.z / chunks;
  const int c0 = (blockIdx.z - n * chunks) * kRChunk;
  const int nch = parts - c0 < kRChunk ? parts - c0 : kRChunk;
  const int total = ss.count * nch;

  if (threadIdx.x < ss.count) {
    // the window of this tile in scale s: first tap of the first row / column .. last tap of the last (clamped, monotonic)
    const ScaleGeom& g = ss.g[threadIdx.x];
    int a, b;
    float fdummy;
    const int y_last = y0 + kRT - 1 < H ? y0 + kRT - 1 : H - 1;
    cubic_src(y0, g.sy, a, fdummy);
    cubic_src(y_last, g.sy, b, fdummy);
    TileWindow w;
    w.r_lo = clampi(a - 1, 0, g.hc - 1);
    w.nrows = clampi(b + 2, 0, g.hc - 1) - w.r_lo + 1;
    cubic_src(x0, g.sx, a, fdummy);
    // the innermost TMA coordinate must be 16-byte aligned: an unaligned one is an illegal-instruction fault on sm_100a
    // (measured with build/tma_probe, profiles/r2_tma_probe.txt), so the window starts at a multiple of four floats
    w.c_lo = clampi(a - 1, 0, g.wc - 1) & ~3;
    s_win[threadIdx.x] = w;
  }
  if (threadIdx.x == 0) {
    for (int b = 0; b < kTmaRing; ++b) ptx::mbar_init(bar0 + 8 * b, 1);
    ptx::mbar_fence_init();
  }
  __syncthreads();
  auto issue = [&](int itn) {  // thread 0: the window of iteration itn into ring slot itn % kTmaRing
    const int s2 = itn / nch, i2 = itn - s2 * nch;
    const uint32_t b = static_cast<uint32_t>(itn % kTmaRing);
    ptx::mbar_arrive_expect_tx(bar0 + 8 * b, box_bytes);
    ptx::tma_load_3d(src0 + b * buf_bytes, &tms.m[s2], bar0 + 8 * b, s_win[s2].c_lo, s_win[s2].r_lo, n * parts + c0 + i2);
  };
  if (threadIdx.x == 0) {
    issue(0);
    if (total > 1) issue(1);
  }

  double acc[4][kRChunk];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int i = 0; i < kRChunk; ++i) acc[k][i] = 0.0;
  }
  const int C = ss.channels;
  const long long tail_start = (static_cast<long long>(W) * C) / 4 * 4;
  const float fS = static_cast<float>(ss.count);
  const bool pow2 = (ss.count & (ss.count - 1)) == 0;  // v / S in float32 (body.py:80-81): for a power of two v * (1/S) is identical
  const float rS = 1.0f / fS;
  int buf = 0;
  int it = 0;
  for (int s = 0; s < ss.count; ++s) {
    const ScaleGeom& g = ss.g[s];
    const TileWindow win = s_win[s];
    // this thread's column: stage-2 taps (as offsets into the window) and weights (once per scale)
    int sx;
    float fx, wx[4];
    cubic_src(xc, g.sx, sx, fx);
    cubic_coeffs(fx, wx);
    int xo[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) xo[k] = clampi(sx - 1 + k, 0, g.wc - 1) - win.c_lo;
    // this thread's four rows: weights and offsets into the shared rows
    float wy[4][4];
    unsigned ro4[4];  // shared-row index of each of the four taps of this thread's four rows, one byte each (< kRMaxRows)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int y = y0 + ty + 8 * k;
      const int yc = y < H ? y : H - 1;
      int sy;
      float fy;
      cubic_src(yc, g.sy, sy, fy);
      cubic_coeffs(fy, wy[k]);
      ro4[k] = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) ro4[k] |= static_cast<unsigned>(clampi(sy - 1 + j, 0, g.hc - 1) - win.r_lo) << (8 * j);
    }
    for (int i = 0; i < nch; ++i, ++it) {
      const int c = c0 + i;
      const uint32_t slot = static_cast<uint32_t>(it % kTmaRing);
      ptx::mbar_wait(bar0 + 8 * slot, static_cast<uint32_t>(it / kTmaRing) & 1u);
      const float* src = s_src + slot * (buf_bytes / 4);
      float (*s_t)[kRT + 1] = buf ? s_t1 : s_t0;
      // horizontal pass: one dot product per (source row, column); rows strided over the 8 warps
      for (int rb = ty; rb < win.nrows; rb += 8) {
        const float* row = src + rb * cols_cap;
        s_t[rb][tx] = dot4_lr(row[xo[0]], row[xo[1]], row[xo[2]], row[xo[3]], wx);
      }
      __syncthreads();
      // every thread is past its reads of the slot iteration it - 1 used: refill it with the window of iteration it + 2
      if (threadIdx.x == 0 && it + 2 < total) issue(it + 2);
      const bool tail = static_cast<long long>(xc) * C + c >= tail_start;
      const float* col = &s_t[0][tx];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float t0 = col[(ro4[k] & 0xffu) * (kRT + 1)], t1 = col[((ro4[k] >> 8) & 0xffu) * (kRT + 1)],
                    t2 = col[((ro4[k] >> 16) & 0xffu) * (kRT + 1)], t3 = col[(ro4[k] >> 24) * (kRT + 1)];
        const float v = tail ? dot4_lr(t0, t1, t2, t3, wy[k]) : dot4_rl(t0, t1, t2, t3, wy[k]);
        const double t = static_cast<double>(pow2 ? __fmul_rn(v, rS) : __fdiv_rn(v, fS));
        // acc is indexed by a loop variable that is not unrolled: select the accumulator without dynamic indexing
#pragma unroll
        for (int q = 0; q < kRChunk; ++q) {
          if (q == i) acc[k][q] = q1 ? __dadd_rn(acc[k][q], __dadd_rn(acc[k][q], t)) : __dadd_rn(acc[k][q], t);
        }
      }
      buf ^= 1;  // the next horizontal pass writes the other buffer: one barrier per (scale, channel)
    }
  }
  if (x < W) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int y = y0 + ty + 8 * k;
      if (y < H) {
#pragma unroll
        for (int i = 0; i < kRChunk; ++i) {
          if (i < nch) out[((static_cast<long long>(n) * parts + c0 + i) * H + y) * W + x] = acc[k][i];
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ gaussian + NMS
// body.py:88-107 on every (frame, part) plane: gauss.cuh's sliding-window filter, then the 4-neighbour NMS against
// zero-filled borders and the threshold; peaks are appended to the plane's list (sorted afterwards).
__global__ void __launch_bounds__(kG2Threads)
gauss_nms_kernel(const double* __restrict__ heat, int H, int W, const GaussWeights gw, double thre, int cap,
                 int* __restrict__ counts, uint32_t* __restrict__ keys, double* __restrict__ scores) {
  __shared__ GaussSmem sm;
  const int plane_id = blockIdx.z;  // n * parts + part
  gauss_window_tile<kGaussPeaks>(sm, heat + static_cast<long long>(plane_id) * H * W, H, W, blockIdx.x * kG2W, blockIdx.y * kG2H,
                                 gw, thre, cap, counts + plane_id, keys + static_cast<long long>(plane_id) * cap,
                                 scores + static_cast<long long>(plane_id) * cap, nullptr);
}

// One CTA per (frame, part): bitonic sort of the appended peaks by y*W+x = np.nonzero order. Lists of up to 1024
// peaks are sorted in shared memory; longer ones (caps up to kMaxPeakCap, after the host grew the capacity) in place in
// global memory, padded with sentinels up to the next power of two (cap itself is a power of two >= that).
constexpr int kSortSmem = 1024;
template <typename KeyPtr, typename ValPtr>
__device__ __forceinline__ void bitonic_by_key(KeyPtr k, ValPtr v, int m) {
  for (int size = 2; size <= m; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const int j = i ^ stride;
        if (j > i) {
          const bool up = (i & size) == 0;
          const uint32_t a = k[i], b = k[j];
          if ((a > b) == up) {
            k[i] = b;
            k[j] = a;
            const double t = v[i];
            v[i] = v[j];
            v[j] = t;
          }
        }
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(512)
sort_peaks_kernel(int cap, int* __restrict__ counts, uint32_t* __restrict__ keys, double* __restrict__ scores,
                  int* __restrict__ overflow) {
  __shared__ uint32_t s_k[kSortSmem];
  __shared__ double s_s[kSortSmem];
  const int id = blockIdx.x;
  int n = counts[id];
  if (n > cap) {
    if (threadIdx.x == 0) {
      atomicOr(overflow, kOverflowPeaks);
      counts[id] = cap;
    }
    n = cap;
  }
  if (n <= 1) return;
  uint32_t* k = keys + static_cast<long long>(id) * cap;
  double* sc = scores + static_cast<long long>(id) * cap;
  int m = 2;
  while (m < n) m <<= 1;
  if (m <= kSortSmem) {
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
      s_k[i] = i < n ? k[i] : 0xffffffffu;
      s_s[i] = i < n ? sc[i] : 0.0;
    }
    __syncthreads();
    bitonic_by_key(s_k, s_s, m);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      k[i] = s_k[i];
      sc[i] = s_s[i];
    }
  } else {  // m <= cap: the sentinels sort to the end and are never read (counts stays n)
    for (int i = n + threadIdx.x; i < m; i += blockDim.x) {
      k[i] = 0xffffffffu;
      sc[i] = 0.0;
    }
    __syncthreads();
    bitonic_by_key(k, sc, m);
  }
}

// ------------------------------------------------------------------------------------------------ launchers
#define ISL_LAUNCH_OK() (cudaGetLastError() == cudaSuccess ? 0 : 1)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int launch_resize_pad_norm(const uint8_t* frames, int N, int H, int W, double scale, int rh, int rw, int hp, int wp,
                           float* out_nchw, uint8_t* out_u8, cudaStream_t st) {
  const dim3 block(32, 8);
  const dim3 grid((wp + 31) / 32, (hp + 7) / 8, N);
  const double inv = 1.0 / scale;  // resize(): scale_x = 1. / inv_scale_x with inv_scale_x = fx
  resize_pad_norm_kernel<<<grid, block, 0, st>>>(frames, N, H, W, rh, rw, hp, wp, inv, inv, out_nchw, out_u8);
  return ISL_LAUNCH_OK();
}

long long heat_accumulate_workspace_floats(const ScaleSet& ss, int N, int parts) {
  long long total = 0;
  for (int s = 0; s < ss.count; ++s) total += static_cast<long long>(N) * parts * ss.g[s].hc * ((ss.g[s].wc + 3) / 4 * 4);
  return total;
}

int launch_heat_accumulate(const ScaleSet& ss, int N, int H, int W, int parts, int q1, double* out, float* workspace,
                           long long workspace_floats, cudaStream_t st) {
  const int chunks = (parts + kChunk - 1) / kChunk;
  const dim3 grid((W + 31) / 32, (H + 7) / 8, N * chunks);
  if (workspace == nullptr || workspace_floats < heat_accumulate_workspace_floats(ss, N, parts)) {
    heat_accumulate_kernel<<<grid, 256, 0, st>>>(ss, N, H, W, parts, q1, out);  // single pass, no scratch needed
    return ISL_LAUNCH_OK();
  }
  MidSet ms;
  float* cursor = workspace;
  int rows_cap = 8;
  for (int s = 0; s < ss.count; ++s) {
    const ScaleGeom& g = ss.g[s];
    // a 32-row frame tile must find its source rows in the shared rows: 32 * hc / H + 4 (+1 for rounding)
    const int need = static_cast<int>(32.0 * g.hc / H) + 6;
    if (need > rows_cap) rows_cap = need;
  }
  if (rows_cap > kRMaxRows) {
    heat_accumulate_kernel<<<grid, 256, 0, st>>>(ss, N, H, W, parts, q1, out);  // very strong down-scaling: single pass
    return ISL_LAUNCH_OK();
  }
  for (int s = 0; s < ss.count; ++s) {
    const ScaleGeom& g = ss.g[s];
    ms.mid[s] = cursor;
    ms.pitch[s] = (g.wc + 3) / 4 * 4;
    const int nbx = (g.wc + 4 + 7) / 8, nby = (g.hc + 4 + 7) / 8;  // 8 x 8 blocks with origin (8b - 4, 8b - 4)
    const dim3 g1((nbx + 31) / 32, (nby + 7) / 8, N * parts);
    upsample8_kernel<<<g1, 256, 0, st>>>(g.low, ss.channels, parts, g.gh, g.gw, g.hc, g.wc, ms.pitch[s], cursor);
    cursor += static_cast<long long>(N) * parts * g.hc * ms.pitch[s];
  }
  const int rchunks = (parts + kRChunk - 1) / kRChunk;
  const dim3 g2((W + kRT - 1) / kRT, (H + kRT - 1) / kRT, N * rchunks);
  const size_t smem = sizeof(float) * 2 * rows_cap * (kRT + 1);
  // TMA-fed second stage when the windows fit a TMA box (<= 256 per dimension) and the driver entry point is there
  int cols_cap = 8;
  for (int s = 0; s < ss.count; ++s) {
    // 32 * wc / W + 4 taps (+ rounding), + 3 for the window start rounded down to 16 bytes, as a multiple of 4 floats
    const int need = (static_cast<int>(32.0 * ss.g[s].wc / W) + 6 + 3 + 3) / 4 * 4;
    if (need > cols_cap) cols_cap = need;
  }
  EncodeTiledFn encode = get_encode_tiled();
  const size_t ring = kTmaRing * ((static_cast<size_t>(rows_cap) * cols_cap * 4 + 127) & ~static_cast<size_t>(127));
  if (encode != nullptr && cols_cap <= 256 && rows_cap <= 255 && ring + smem + 128 <= 160 * 1024) {
    ResizeMaps tms;
    bool ok = true;
    for (int s = 0; s < ss.count && ok; ++s) {
      const ScaleGeom& g = ss.g[s];
      cuuint64_t gdim[3] = {static_cast<cuuint64_t>(ms.pitch[s]), static_cast<cuuint64_t>(g.hc), static_cast<cuuint64_t>(N) * parts};
      cuuint64_t gstr[2] = {static_cast<cuuint64_t>(ms.pitch[s]) * 4, static_cast<cuuint64_t>(ms.pitch[s]) * 4 * g.hc};
      cuuint32_t box[3] = {static_cast<cuuint32_t>(cols_cap), static_cast<cuuint32_t>(rows_cap), 1};
      cuuint32_t estr[3] = {1, 1, 1};
      ok = encode(&tms.m[s], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ms.mid[s]), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    }
    if (ok) {
      static bool attr_dev[64] = {};
      int dev = 0;
      cudaGetDevice(&dev);
      if (!attr_dev[dev & 63]) {
        ok = cudaFuncSetAttribute(resize_accumulate_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024) == cudaSuccess;
        attr_dev[dev & 63] = ok;
      }
    }
    if (ok) {
      resize_accumulate_tma_kernel<<<g2, 256, ring + smem + 128, st>>>(tms, ss, N, H, W, parts, q1, rows_cap, cols_cap, out);
      return ISL_LAUNCH_OK();
    }
  }
  resize_accumulate_kernel<<<g2, 256, smem, st>>>(ss, ms, N, H, W, parts, q1, rows_cap, out);
  return ISL_LAUNCH_OK();
}

int launch_gauss_nms(const double* heat, int planes_total, int H, int W, const GaussWeights& gw, double thre, int cap,
                     int* counts, uint32_t* keys, double* scores, int* overflow, cudaStream_t st) {
  if (cap > kMaxPeakCap || (cap > kSortSmem && (cap & (cap - 1)) != 0)) return 1;
  if (cudaMemsetAsync(counts, 0, sizeof(int) * planes_total, st) != cudaSuccess) return 1;
  const dim3 grid((W + kG2W - 1) / kG2W, (H + kG2H - 1) / kG2H, planes_total);
  gauss_nms_kernel<<<grid, kG2Threads, 0, st>>>(heat, H, W, gw, thre, cap, counts, keys, scores);
  sort_peaks_kernel<<<planes_total, 512, 0, st>>>(cap, counts, keys, scores, overflow);
  return ISL_LAUNCH_OK();
}

}  // namespace islpose
