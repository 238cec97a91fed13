// tcgen05 / TMEM / TMA implicit-GEMM convolution (see conv_umma.cuh for the data layout, DESIGN.md section 4.1 for
// the measurements behind every choice).
//
// Variants (conv_prepare picks one per layer):
//   v5  swapped operands + resident activation halo, persistent: every 3x3 / 7x7 layer with >= 64 input and >= 48
//       output channels (~97 % of the FLOPs); bias / ReLU / PReLU / bf16 / slice write / 2x2 max-pool in the epilogue
//   v1  one pixel tile per CTA, two CTAs per SM: 1x1 layers and the float32 network heads
//   v2  persistent, double-buffered TMEM, two pixel sub-tiles per weight stage: narrow 1x1 heads on large grids
//   v3, v4  earlier halo / swapped-operand kernels: compiled only into build/conv_test (-DISLPOSE_BRINGUP_VARIANTS) for
//           A/B measurements; libislpose.so holds v1, v2, v5 and has no environment switches
// Warp roles in every variant:
//   warp 0      : TMA producer - the whole warp runs the loop with warp-uniform state, one elected lane issues
//   warp 1      : TMEM owner + MMA issuer - same pattern; tcgen05.commit releases ring slots and signals the epilogue
//   warps 2..   : epilogue (4 or 8 warps) - tcgen05.ld the accumulator (warp w may only touch TMEM lanes
//                 32*(w%4)..+31), bias + ReLU/PReLU, bf16, stores into the NHWC slice; network heads store fp32.
#include "conv_umma.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <utility>

#include "ptx.cuh"

namespace islpose {

namespace {

#ifdef ISLPOSE_BRINGUP_VARIANTS
constexpr int kThreads = 192;  // v3 / v4
#endif
constexpr int kThreadsV1 = 320;  // v1: eight epilogue warps (two per TMEM lane quarter, alternating 32-column chunks)
constexpr uint32_t kASlotBytes = 128 * 128;  // 128 pixel rows x 64 bf16
constexpr int kMaxStages = 8;
constexpr uint32_t kCtrlBytes = 256 + 2 * 256 * 4;  // barriers + bias + slope

// The four K=16 MMAs of one 64-channel block, fully unrolled with constant descriptor offsets (+32 bytes along K
// inside the 128-byte swizzle row = +2 in the >>4 address field). Every instruction of an issue loop is on the
// critical path of the tensor pipe (build/mma_rate: a lean loop issues an MMA every ~70 cycles; runtime divisions and
// per-MMA operand broadcasts from one lane's registers had pushed that to ~270). No division, no variable trip
// count, no per-MMA predicate in the common case.
__device__ __forceinline__ void issue_kblock4(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, bool accumulate_first) {
  ptx::umma_bf16(d, da, db, idesc, accumulate_first ? 1u : 0u);
  ptx::umma_bf16_acc(d, da + 2, db + 2, idesc);
  ptx::umma_bf16_acc(d, da + 4, db + 4, idesc);
  ptx::umma_bf16_acc(d, da + 6, db + 6, idesc);
}
__device__ __forceinline__ void issue_kblock(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, bool accumulate_first,
                                             int ksteps) {
  if (ksteps == 4) {
    issue_kblock4(d, da, db, idesc, accumulate_first);
  } else {  // the slice ends inside this block (only the 32-channel first layer)
    for (int k = 0; k < ksteps; ++k) ptx::umma_bf16(d, da + 2 * k, db + 2 * k, idesc, (accumulate_first || k != 0) ? 1u : 0u);
  }
}

__global__ void __launch_bounds__(kThreadsV1, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const ConvArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* const gbase = smem_raw + (base - raw);

  const uint32_t sA0 = base;
  const uint32_t sB0 = base + a.stages * kASlotBytes;
  const uint32_t ctrl = sB0 + a.stages * a.b_stage_bytes;
  uint8_t* const gctrl = gbase + (ctrl - base);
  const uint32_t bar_full = ctrl;         // kMaxStages x 8 B
  const uint32_t bar_empty = ctrl + 64;   // kMaxStages x 8 B
  const uint32_t bar_accum = ctrl + 128;  // accumulator ready
  const uint32_t tmem_slot = ctrl + 136;  // TMEM base address written by tcgen05.alloc
  volatile uint32_t* const tmem_slot_g = reinterpret_cast<volatile uint32_t*>(gctrl + 136);
  float* const s_bias = reinterpret_cast<float*>(gctrl + 256);
  float* const s_slope = s_bias + 256;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // tile coordinates
  const int t = blockIdx.x;
  const int tx = t % a.tiles_x;
  const int ty = (t / a.tiles_x) % a.tiles_y;
  const int img = t / (a.tiles_x * a.tiles_y);
  const int x0 = tx * a.bw;
  const int y0 = ty * a.bh;
  const int n0 = blockIdx.y * a.n_tile;

  const int cblocks = (a.cin_k16 + 3) >> 2;
  const int taps = a.ksize * a.ksize;

  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) {
      ptx::mbar_init(bar_full + 8 * s, 1);
      ptx::mbar_init(bar_empty + 8 * s, 1);
    }
    ptx::mbar_init(bar_accum, 1);
    ptx::mbar_fence_init();
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, a.tmem_cols);
    ptx::tmem_relinquish();
  }
  if (warp >= 2) {
    for (int i = threadIdx.x - 64; i < a.n_tile; i += kThreadsV1 - 64) {
      s_bias[i] = a.bias[n0 + i];
      s_slope[i] = a.slope[n0 + i];
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_g;
  ptx::pdl_launch_dependents();  // programmatic dependent launch: see the v5 kernel

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (whole warp, one elected lane issues)
    {
      ptx::pdl_wait();  // the previous layer's output is first touched by the loads below
      const uint32_t tx_bytes = 128u * a.bw * a.bh + 128u * a.n_tile;
      const int total = taps * cblocks;
      const int xb = x0 - a.pad, yb = y0 - a.pad;
      ptx::RingPos r(bar_full, bar_empty, a.stages);
      uint32_t sa = sA0, sb = sB0;
      int kx = 0, ky = 0, cb = 0, tap = 0;
      for (int it = 0; it < total; ++it) {
        ptx::mbar_wait(r.empty, r.ph ^ 1);
        if (ptx::elect_one()) {
          if (a.debug_no_loads && it >= a.stages) {  // measurement aid: the MMAs re-read what the ring already holds
            ptx::mbar_arrive(r.full);
          } else {
            ptx::mbar_arrive_expect_tx(r.full, tx_bytes);
            ptx::tma_load_4d(sa, &tmA, r.full, cb * 64, xb + kx, yb + ky, img);
            ptx::tma_load_3d(sb, &tmB, r.full, cb * 64, n0, tap);
          }
        }
        __syncwarp();
        if (++cb == cblocks) {
          cb = 0;
          ++tap;
          if (++kx == a.ksize) {
            kx = 0;
            ++ky;
          }
        }
        sa += kASlotBytes;
        sb += a.b_stage_bytes;
        if (r.advance()) {
          sa = sA0;
          sb = sB0;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (whole warp, one elected lane issues)
    {
      const uint32_t idesc = ptx::umma_idesc_bf16(128, a.n_tile);
      const int total = taps * cblocks;
      const int tail_ksteps = a.cin_k16 - (cblocks - 1) * 4;  // K=16 steps of the last channel block: 1..4
      const uint64_t da0 = ptx::umma_desc_sw128(sA0), db0 = ptx::umma_desc_sw128(sB0);
      const uint32_t a_step = kASlotBytes >> 4, b_step = a.b_stage_bytes >> 4;
      uint64_t da = da0, db = db0;
      ptx::RingPos r(bar_full, bar_empty, a.stages);
      int cb = 0;
      for (int it = 0; it < total; ++it) {
        ptx::mbar_wait(r.full, r.ph);
        ptx::tc_fence_after();
        int ksteps = 4;
        if (tail_ksteps != 4 && ++cb == cblocks) {
          cb = 0;
          ksteps = tail_ksteps;
        }
        if (ptx::elect_one()) {
          issue_kblock(tmem_base, da, db, idesc, it != 0, ksteps);
          ptx::umma_commit(r.empty);  // slot reusable once these MMAs have read it
        }
        __syncwarp();
        da += a_step;
        db += b_step;
        if (r.advance()) {
          da = da0;
          db = db0;
        }
      }
      if (ptx::elect_one()) ptx::umma_commit(bar_accum);
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------ epilogue
    const int q = warp & 3;          // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;   // accumulator row = pixel index inside the tile box
    const int py = row / a.bw;
    const int px = row - py * a.bw;
    const int x = x0 + px;
    const int y = y0 + py;
    const bool valid = (row < a.bw * a.bh) && (x < a.W) && (y < a.H);
    const long long pix = (static_cast<long long>(img) * a.H + y) * a.W + x;

    ptx::mbar_wait(bar_accum, 0);
    ptx::tc_fence_after();

    // Layers with a short K loop (1x1) spend longer here than in the MMAs, and one warp per scheduler issues an
    // instruction every ~4 cycles at best: two warps share each TMEM lane quarter and take alternate 32-column chunks.
    for (int c = ((warp - 2) >> 2) * 32; c < a.n_tile; c += 64) {
      uint32_t r[32];
      ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c, r);
      ptx::tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float acc = __uint_as_float(r[i]) + s_bias[c + i];
        v[i] = acc > 0.f ? acc : acc * s_slope[c + i];
      }
      if (a.tma_store) {
        // The bf16 tile leaves through the (now idle) operand ring: per 64 channels one [pixel][64 ch] block in the
        // 128-byte-swizzled layout a TMA store expects. A thread owns one pixel row, so direct stores would put 16
        // bytes into each of 32 different lines per instruction (measured: 1x1 layers with wide outputs stuck at
        // ~1.1 TB/s of output, 150-450 TFLOP/s); the bulk store writes whole lines and clips pixels outside the
        // frame and channels beyond the slice.
        const uint32_t blk = sA0 + static_cast<uint32_t>(c >> 6) * kASlotBytes + static_cast<uint32_t>(row) * 128u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          __nv_bfloat162 h0 = __floats2bfloat162_rn(v[8 * j + 0], v[8 * j + 1]);
          __nv_bfloat162 h1 = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]);
          __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]);
          __nv_bfloat162 h3 = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]);
          const uint32_t chunk = static_cast<uint32_t>(((c & 63) >> 3) + j);
          const uint32_t addr = blk + ((chunk ^ (static_cast<uint32_t>(row) & 7u)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(*reinterpret_cast<uint32_t*>(&h0)),
                       "r"(*reinterpret_cast<uint32_t*>(&h1)), "r"(*reinterpret_cast<uint32_t*>(&h2)),
                       "r"(*reinterpret_cast<uint32_t*>(&h3))
                       : "memory");
        }
      }
      if (valid) {
        if (a.out_bf16 != nullptr && !a.tma_store) {
          __nv_bfloat16* dst = a.out_bf16 + pix * a.out_pix_stride + n0 + c;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (n0 + c + 8 * j < a.cout_store) {
              uint4 pk;
              __nv_bfloat162 h;
              h = __floats2bfloat162_rn(v[8 * j + 0], v[8 * j + 1]);
              pk.x = *reinterpret_cast<uint32_t*>(&h);
              h = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]);
              pk.y = *reinterpret_cast<uint32_t*>(&h);
              h = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]);
              pk.z = *reinterpret_cast<uint32_t*>(&h);
              h = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]);
              pk.w = *reinterpret_cast<uint32_t*>(&h);
              *reinterpret_cast<uint4*>(dst + 8 * j) = pk;
            }
          }
        }
        if (a.out_f32 != nullptr) {
          // planar NCHW: for a fixed channel the 32 lanes of a warp write neighbouring pixels
          const long long plane = static_cast<long long>(a.H) * a.W;
          float* dst = a.out_f32 + (static_cast<long long>(img) * a.out_f32_channels + n0 + c) * plane +
                       static_cast<long long>(y) * a.W + x;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (n0 + c + i < a.cout) dst[i * plane] = v[i];
          }
        }
      }
    }
    if (a.tma_store) {
      ptx::fence_proxy_async_smem();
      asm volatile("bar.sync 1, 256;" ::: "memory");  // the eight epilogue warps
      if (threadIdx.x == 64) {
        for (int b = 0; b * 64 < a.n_tile; ++b) {
          if (n0 + b * 64 < a.cout_store) ptx::tma_store_4d(&tmC, sA0 + b * kASlotBytes, n0 + b * 64, x0, y0, img);
        }
        ptx::tma_store_commit();
        ptx::tma_store_wait_all();
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tmem_dealloc(tmem_base, a.tmem_cols);
  }
}


// ------------------------------------------------------------------------------------------------ v2
// Persistent variant: one CTA per SM walks work items (msub pixel tiles x one channel tile) round-robin.
//   * the accumulator is double buffered in TMEM (2 x msub x n_tile columns), so the epilogue of work item w
//     overlaps the main loop of w+1 and the per-tile set-up (barrier init, TMEM allocation) is paid once per CTA;
//   * with msub = 2 one weight stage feeds two 128-pixel MMAs: half the weight traffic from L2 per FLOP.
__global__ void __launch_bounds__(kThreadsV1, 2)
conv_umma_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                            const ConvArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* const gbase = smem_raw + (base - raw);

  const uint32_t a_stage_bytes = a.msub * kASlotBytes;
  const uint32_t sA0 = base;
  const uint32_t sB0 = base + a.stages * a_stage_bytes;
  const uint32_t ctrl = sB0 + a.stages * a.b_stage_bytes;
  uint8_t* const gctrl = gbase + (ctrl - base);
  const uint32_t bar_full = ctrl;            // kMaxStages x 8 B
  const uint32_t bar_empty = ctrl + 64;      // kMaxStages x 8 B
  const uint32_t bar_acc_full = ctrl + 128;  // 2 x 8 B: accumulator buffer written
  const uint32_t bar_acc_empty = ctrl + 144; // 2 x 8 B: accumulator buffer drained
  const uint32_t tmem_slot = ctrl + 160;
  volatile uint32_t* const tmem_slot_g = reinterpret_cast<volatile uint32_t*>(gctrl + 160);
  float* const s_bias = reinterpret_cast<float*>(gctrl + 256);
  float* const s_slope = s_bias + 256;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cblocks = (a.cin_k16 + 3) >> 2;
  const int taps = a.ksize * a.ksize;
  const int tiles_per_img = a.tiles_x * a.tiles_y;
  const int m_groups = (a.m_tiles + a.msub - 1) / a.msub;

  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) {
      ptx::mbar_init(bar_full + 8 * s, 1);
      ptx::mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(bar_acc_full + 8 * b, 1);
      ptx::mbar_init(bar_acc_empty + 8 * b, 8);  // one arrival per epilogue warp
    }
    ptx::mbar_fence_init();
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, a.tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_g;
  const uint32_t acc_cols = a.msub * a.n_tile;  // columns of one accumulator buffer
  ptx::pdl_launch_dependents();  // programmatic dependent launch: see the v5 kernel

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (whole warp, one elected lane issues)
    {
      ptx::pdl_wait();  // the previous layer's output is first touched by the loads below
      ptx::RingPos r(bar_full, bar_empty, a.stages);
      uint32_t sa = sA0, sb = sB0;
      for (int w = blockIdx.x; w < a.work_items; w += gridDim.x) {
        const int mg = w % m_groups;
        const int n0 = (w / m_groups) * a.n_tile;
        int x0[2], y0[2], img[2];
        int live = 0;
        for (int sub = 0; sub < a.msub; ++sub) {
          const int t = mg * a.msub + sub;
          if (t < a.m_tiles) {
            img[sub] = t / tiles_per_img;
            const int rr = t - img[sub] * tiles_per_img;
            y0[sub] = (rr / a.tiles_x) * a.bh - a.pad;
            x0[sub] = (rr % a.tiles_x) * a.bw - a.pad;
            ++live;
          }
        }
        const uint32_t tx_bytes = 128u * a.bw * a.bh * live + 128u * a.n_tile;
        int tap = 0;
        for (int ky = 0; ky < a.ksize; ++ky) {
          for (int kx = 0; kx < a.ksize; ++kx, ++tap) {
            for (int cb = 0; cb < cblocks; ++cb) {
              ptx::mbar_wait(r.empty, r.ph ^ 1);
              if (ptx::elect_one()) {
                ptx::mbar_arrive_expect_tx(r.full, tx_bytes);
                ptx::tma_load_4d(sa, &tmA, r.full, cb * 64, x0[0] + kx, y0[0] + ky, img[0]);
                if (live > 1) ptx::tma_load_4d(sa + kASlotBytes, &tmA, r.full, cb * 64, x0[1] + kx, y0[1] + ky, img[1]);
                ptx::tma_load_3d(sb, &tmB, r.full, cb * 64, n0, tap);
              }
              __syncwarp();
              sa += a_stage_bytes;
              sb += a.b_stage_bytes;
              if (r.advance()) {
                sa = sA0;
                sb = sB0;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (whole warp, one elected lane issues)
    {
      const uint32_t idesc = ptx::umma_idesc_bf16(128, a.n_tile);
      const int total = taps * cblocks;
      const int tail_ksteps = a.cin_k16 - (cblocks - 1) * 4;
      const uint64_t da0 = ptx::umma_desc_sw128(sA0), db0 = ptx::umma_desc_sw128(sB0);
      const uint32_t a_step = a_stage_bytes >> 4, b_step = a.b_stage_bytes >> 4, sub_step = kASlotBytes >> 4;
      uint64_t da = da0, db = db0;
      ptx::RingPos r(bar_full, bar_empty, a.stages);
      uint32_t buf = 0, buf_ph = 0;  // accumulator buffer of this work item and the phase of its current use
      for (int w = blockIdx.x; w < a.work_items; w += gridDim.x) {
        const int mg = w % m_groups;
        const int live = min(a.msub, a.m_tiles - mg * a.msub);
        ptx::mbar_wait(bar_acc_empty + 8 * buf, buf_ph ^ 1);  // epilogue has drained this buffer
        ptx::tc_fence_after();
        const uint32_t acc = tmem_base + buf * acc_cols;
        int cb = 0;
        for (int kb = 0; kb < total; ++kb) {
          ptx::mbar_wait(r.full, r.ph);
          ptx::tc_fence_after();
          int ksteps = 4;
          if (tail_ksteps != 4) {
            if (++cb == cblocks) {
              cb = 0;
              ksteps = tail_ksteps;
            }
          }
          if (ptx::elect_one()) {
            issue_kblock(acc, da, db, idesc, kb != 0, ksteps);
            if (live > 1) issue_kblock(acc + a.n_tile, da + sub_step, db, idesc, kb != 0, ksteps);
            ptx::umma_commit(r.empty);
          }
          __syncwarp();
          da += a_step;
          db += b_step;
          if (r.advance()) {
            da = da0;
            db = db0;
          }
        }
        if (ptx::elect_one()) ptx::umma_commit(bar_acc_full + 8 * buf);
        __syncwarp();
        if (++buf == static_cast<uint32_t>(a.acc_bufs)) {
          buf = 0;
          buf_ph ^= 1;
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..9 = 256 threads; the two warps of a
    // TMEM lane quarter take alternate 32-column chunks)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int py = row / a.bw;
    const int px = row - py * a.bw;
    const int et = threadIdx.x - 64;
    int local = 0;
    int cur_n0 = -1;
    for (int w = blockIdx.x; w < a.work_items; w += gridDim.x, ++local) {
      const int mg = w % m_groups;
      const int n0 = (w / m_groups) * a.n_tile;
      const int live = min(a.msub, a.m_tiles - mg * a.msub);
      const int buf = local % a.acc_bufs;
      if (n0 != cur_n0) {  // (re)load this channel tile's bias / slope; named barrier 1 = the 4 epilogue warps
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int i = et; i < a.n_tile; i += 256) {
          s_bias[i] = a.bias[n0 + i];
          s_slope[i] = a.slope[n0 + i];
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        cur_n0 = n0;
      }
      ptx::mbar_wait(bar_acc_full + 8 * buf, (local / a.acc_bufs) & 1);
      ptx::tc_fence_after();
      for (int sub = 0; sub < live; ++sub) {
        const int t = mg * a.msub + sub;
        const int img = t / tiles_per_img;
        const int r = t - img * tiles_per_img;
        const int y = (r / a.tiles_x) * a.bh + py;
        const int x = (r % a.tiles_x) * a.bw + px;
        const bool valid = (row < a.bw * a.bh) && (x < a.W) && (y < a.H);
        const long long pix = (static_cast<long long>(img) * a.H + y) * a.W + x;
        const uint32_t acc = tmem_base + buf * acc_cols + sub * a.n_tile + (static_cast<uint32_t>(q * 32) << 16);
        for (int c = ((warp - 2) >> 2) * 32; c < a.n_tile; c += 64) {
          uint32_t rr[32];
          ptx::tmem_ld_32x32(acc + c, rr);
          ptx::tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float accv = __uint_as_float(rr[i]) + s_bias[c + i];
            v[i] = accv > 0.f ? accv : accv * s_slope[c + i];
          }
          if (valid) {
            if (a.out_bf16 != nullptr) {
              __nv_bfloat16* dst = a.out_bf16 + pix * a.out_pix_stride + n0 + c;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (n0 + c + 8 * j < a.cout_store) {
                  uint4 pk;
                  __nv_bfloat162 h;
                  h = __floats2bfloat162_rn(v[8 * j + 0], v[8 * j + 1]);
                  pk.x = *reinterpret_cast<uint32_t*>(&h);
                  h = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]);
                  pk.y = *reinterpret_cast<uint32_t*>(&h);
                  h = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]);
                  pk.z = *reinterpret_cast<uint32_t*>(&h);
                  h = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]);
                  pk.w = *reinterpret_cast<uint32_t*>(&h);
                  *reinterpret_cast<uint4*>(dst + 8 * j) = pk;
                }
              }
            }
            if (a.out_f32 != nullptr) {
              const long long plane = static_cast<long long>(a.H) * a.W;
              float* dst = a.out_f32 + (static_cast<long long>(img) * a.out_f32_channels + n0 + c) * plane +
                           static_cast<long long>(y) * a.W + x;
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                if (n0 + c + i < a.cout) dst[i * plane] = v[i];
              }
            }
          }
        }
      }
      // this warp has finished reading the buffer: hand it back to the MMA issuer
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_acc_empty + 8 * buf);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tmem_dealloc(tmem_base, a.tmem_cols);
  }
}


#ifdef ISLPOSE_BRINGUP_VARIANTS  // v3 / v4: earlier kernels kept for A/B measurements in build/conv_test only
// ------------------------------------------------------------------------------------------------ v3
// Halo variant for k > 1: the CTA's 8 x 16 pixel tile needs, per 64-channel block, the (8+2p) x (16+2p) input
// pixels around it. They are fetched ONCE as one TMA box of 16 x (16+2p) pixel rows (pitch 16 keeps every 8-row
// group of the operand on the same swizzle phase) and all k*k taps read shifted windows of that box straight from
// shared memory: the A operand of tap (ky, kx) starts at halo row ky*16 + kx, its 8-row groups (one image row of
// the tile each) are 16 rows = 2048 B apart, and the swizzle is a function of the absolute shared-memory
// address (TMA wrote it that way, the MMA reads it that way), so the descriptor's base offset stays 0. Only the weights stream from L2 per tap, which cuts the bytes delivered to the SM per FLOP by ~45 %
// (7x7) / ~30 % (3x3). Superseded by v5 (same idea with the weights as the M operand and a persistent CTA).
__global__ void __launch_bounds__(kThreads, 2)
conv_umma_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const ConvArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* const gbase = smem_raw + (base - raw);

  const uint32_t halo_bytes = 128u * a.halo_w * a.halo_h;
  const uint32_t halo_slot = (halo_bytes + 1023u) & ~1023u;
  const uint32_t sH = base;                       // one halo buffer
  const uint32_t sB0 = base + halo_slot;          // weight ring
  const uint32_t ctrl = sB0 + a.stages * a.b_stage_bytes;
  uint8_t* const gctrl = gbase + (ctrl - base);
  const uint32_t bar_full = ctrl;         // weight stage filled
  const uint32_t bar_empty = ctrl + 64;   // weight stage consumed
  const uint32_t bar_accum = ctrl + 128;
  const uint32_t tmem_slot = ctrl + 136;
  const uint32_t bar_hfull = ctrl + 144;  // halo tile landed
  const uint32_t bar_hempty = ctrl + 152; // all taps of this channel block have read the halo tile
  volatile uint32_t* const tmem_slot_g = reinterpret_cast<volatile uint32_t*>(gctrl + 136);
  float* const s_bias = reinterpret_cast<float*>(gctrl + 256);
  float* const s_slope = s_bias + 256;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int t = blockIdx.x;
  const int tx = t % a.tiles_x;
  const int ty = (t / a.tiles_x) % a.tiles_y;
  const int img = t / (a.tiles_x * a.tiles_y);
  const int x0 = tx * a.bw;
  const int y0 = ty * a.bh;
  const int n0 = blockIdx.y * a.n_tile;
  const int cblocks = (a.cin_k16 + 3) >> 2;
  const int taps = a.ksize * a.ksize;

  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) {
      ptx::mbar_init(bar_full + 8 * s, 1);
      ptx::mbar_init(bar_empty + 8 * s, 1);
    }
    ptx::mbar_init(bar_accum, 1);
    ptx::mbar_init(bar_hfull, 1);
    ptx::mbar_init(bar_hempty, 1);
    ptx::mbar_fence_init();
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, a.tmem_cols);
    ptx::tmem_relinquish();
  }
  if (warp >= 2) {
    for (int i = threadIdx.x - 64; i < a.n_tile; i += kThreads - 64) {
      s_bias[i] = a.bias[n0 + i];
      s_slope[i] = a.slope[n0 + i];
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_g;

  if (warp == 0) {
    if (lane == 0) {
      ptx::RingPos r(bar_full, bar_empty, a.stages);
      uint32_t sb = sB0;
      for (int cb = 0; cb < cblocks; ++cb) {
        ptx::mbar_wait(bar_hempty, (cb & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(bar_hfull, halo_bytes);
        ptx::tma_load_4d(sH, &tmA, bar_hfull, cb * 64, x0 - a.pad, y0 - a.pad, img);
        for (int tap = 0; tap < taps; ++tap) {
          ptx::mbar_wait(r.empty, r.ph ^ 1);
          ptx::mbar_arrive_expect_tx(r.full, 128u * a.n_tile);
          ptx::tma_load_3d(sb, &tmB, r.full, cb * 64, n0, tap);
          sb += a.b_stage_bytes;
          if (r.advance()) sb = sB0;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = ptx::umma_idesc_bf16(128, a.n_tile);
      const uint32_t sbo = 128u * a.halo_w;  // one image row of the tile further down = halo_w pixel rows
      const uint64_t db0 = ptx::umma_desc_sw128(sB0);
      const uint32_t b_step = a.b_stage_bytes >> 4;
      uint64_t db = db0;
      ptx::RingPos r(bar_full, bar_empty, a.stages);
      for (int cb = 0; cb < cblocks; ++cb) {
        ptx::mbar_wait(bar_hfull, cb & 1);
        ptx::tc_fence_after();
        const int ksteps = min(4, a.cin_k16 - cb * 4);
        for (int ky = 0; ky < a.ksize; ++ky) {
          for (int kx = 0; kx < a.ksize; ++kx) {
            ptx::mbar_wait(r.full, r.ph);
            ptx::tc_fence_after();
            const uint32_t arow = sH + 128u * (ky * a.halo_w + kx);
            // measured on B200: the swizzle XOR is taken from the absolute shared-memory address, so a window that
            // starts on any 128-byte row needs base offset 0 (mode 1 = the start row's phase, kept for the bring-up test)
            const uint32_t bo = a.halo_bo_mode == 1 ? ((arow >> 7) & 7u) : 0u;
            const uint64_t da = ptx::umma_desc_sw128_strided(arow, sbo, bo);
            issue_kblock(tmem_base, da, db, idesc, (cb | ky | kx) != 0, ksteps);
            ptx::umma_commit(r.empty);
            db += b_step;
            if (r.advance()) db = db0;
          }
        }
        ptx::umma_commit(bar_hempty);  // the halo tile may be overwritten once these MMAs have retired
      }
      ptx::umma_commit(bar_accum);
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int py = row / a.bw;
    const int px = row - py * a.bw;
    const int x = x0 + px;
    const int y = y0 + py;
    const bool valid = (x < a.W) && (y < a.H);
    const long long pix = (static_cast<long long>(img) * a.H + y) * a.W + x;
    ptx::mbar_wait(bar_accum, 0);
    ptx::tc_fence_after();
    for (int c = 0; c < a.n_tile; c += 32) {
      uint32_t r[32];
      ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c, r);
      ptx::tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float acc = __uint_as_float(r[i]) + s_bias[c + i];
        v[i] = acc > 0.f ? acc : acc * s_slope[c + i];
      }
      if (valid) {
        if (a.out_bf16 != nullptr) {
          __nv_bfloat16* dst = a.out_bf16 + pix * a.out_pix_stride + n0 + c;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (n0 + c + 8 * j < a.cout_store) {
              uint4 pk;
              __nv_bfloat162 h;
              h = __floats2bfloat162_rn(v[8 * j + 0], v[8 * j + 1]);
              pk.x = *reinterpret_cast<uint32_t*>(&h);
              h = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]);
              pk.y = *reinterpret_cast<uint32_t*>(&h);
              h = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]);
              pk.z = *reinterpret_cast<uint32_t*>(&h);
              h = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]);
              pk.w = *reinterpret_cast<uint32_t*>(&h);
              *reinterpret_cast<uint4*>(dst + 8 * j) = pk;
            }
          }
        }
        if (a.out_f32 != nullptr) {
          const long long plane = static_cast<long long>(a.H) * a.W;
          float* dst = a.out_f32 + (static_cast<long long>(img) * a.out_f32_channels + n0 + c) * plane +
                       static_cast<long long>(y) * a.W + x;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (n0 + c + i < a.cout) dst[i * plane] = v[i];
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tmem_dealloc(tmem_base, a.tmem_cols);
  }
}


// ------------------------------------------------------------------------------------------------ v4
// Swapped operands for 65..128 output channels. Measured on B200 (profiles/conv_limits_r1.log): the MMA loop runs
// at the same speed with and without TMA traffic, and throughput follows the instruction shape (N=64: 558, N=128:
// 910, N=256: 1340 TFLOP/s) - every tcgen05.mma carries a fixed cost that only a larger instruction amortises. A
// 128-channel layer cannot offer N=256 channels, but it can offer 256 pixels: D^T[channels, pixels] = W * X^T with
// the weights as the M=128 operand (rows beyond Cout are zero-filled by TMA) and a 256-pixel activation box as N.
// The accumulator is then [128 lanes = channels] x [256 columns = pixels]; epilogue lane = channel.
__global__ void __launch_bounds__(kThreads, 2)
conv_umma_swapped_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const __grid_constant__ CUtensorMap tmC, const ConvArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* const gbase = smem_raw + (base - raw);

  constexpr uint32_t kWBytes = 128 * 128;   // 128 output channels x 64 input channels
  constexpr uint32_t kXBytes = 256 * 128;   // up to 256 pixels x 64 input channels
  const uint32_t sW0 = base;
  const uint32_t sX0 = base + a.stages * kWBytes;
  const uint32_t ctrl = sX0 + a.stages * kXBytes;
  uint8_t* const gctrl = gbase + (ctrl - base);
  const uint32_t bar_full = ctrl;
  const uint32_t bar_empty = ctrl + 64;
  const uint32_t bar_accum = ctrl + 128;
  const uint32_t tmem_slot = ctrl + 136;
  volatile uint32_t* const tmem_slot_g = reinterpret_cast<volatile uint32_t*>(gctrl + 136);
  long long* const s_pix = reinterpret_cast<long long*>(gctrl + 256);  // 256 pixel offsets (-1 = outside)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int t = blockIdx.x;
  const int tx = t % a.tiles_x;
  const int ty = (t / a.tiles_x) % a.tiles_y;
  const int img = t / (a.tiles_x * a.tiles_y);
  const int x0 = tx * a.bw;
  const int y0 = ty * a.bh;
  const int n0 = blockIdx.y * 128;
  const int cblocks = (a.cin_k16 + 3) >> 2;
  const int taps = a.ksize * a.ksize;

  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) {
      ptx::mbar_init(bar_full + 8 * s, 1);
      ptx::mbar_init(bar_empty + 8 * s, 1);
    }
    ptx::mbar_init(bar_accum, 1);
    ptx::mbar_fence_init();
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, a.tmem_cols);
    ptx::tmem_relinquish();
  }
  if (warp >= 2) {
    for (int j = threadIdx.x - 64; j < 256; j += kThreads - 64) {
      const int py = j / a.bw;
      const int px = j - py * a.bw;
      const int x = x0 + px, y = y0 + py;
      const bool ok = j < a.bw * a.bh && x < a.W && y < a.H;
      s_pix[j] = ok ? (static_cast<long long>(img) * a.H + y) * a.W + x : -1;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_g;

  if (warp == 0) {
    {
      const uint32_t tx_bytes = kWBytes + 128u * a.bw * a.bh;
      const int total = taps * cblocks;
      const int xb = x0 - a.pad, yb = y0 - a.pad;
      ptx::RingPos r(bar_full, bar_empty, a.stages);
      uint32_t sw = sW0, sx = sX0;
      int kx = 0, ky = 0, cb = 0, tap = 0;
      for (int it = 0; it < total; ++it) {
        ptx::mbar_wait(r.empty, r.ph ^ 1);
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(r.full, tx_bytes);
          ptx::tma_load_3d(sw, &tmB, r.full, cb * 64, n0, tap);
          ptx::tma_load_4d(sx, &tmA, r.full, cb * 64, xb + kx, yb + ky, img);
        }
        __syncwarp();
        if (++cb == cblocks) {
          cb = 0;
          ++tap;
          if (++kx == a.ksize) {
            kx = 0;
            ++ky;
          }
        }
        sw += kWBytes;
        sx += kXBytes;
        if (r.advance()) {
          sw = sW0;
          sx = sX0;
        }
      }
    }
  } else if (warp == 1) {
    {
      const uint32_t idesc = ptx::umma_idesc_bf16(128, a.n_pix);
      const int total = taps * cblocks;
      const int tail_ksteps = a.cin_k16 - (cblocks - 1) * 4;
      const uint64_t dw0 = ptx::umma_desc_sw128(sW0), dx0 = ptx::umma_desc_sw128(sX0);
      uint64_t dw = dw0, dx = dx0;
      ptx::RingPos r(bar_full, bar_empty, a.stages);
      int cb = 0;
      for (int it = 0; it < total; ++it) {
        ptx::mbar_wait(r.full, r.ph);
        ptx::tc_fence_after();
        int ksteps = 4;
        if (tail_ksteps != 4 && ++cb == cblocks) {
          cb = 0;
          ksteps = tail_ksteps;
        }
        if (ptx::elect_one()) {
          issue_kblock(tmem_base, dw, dx, idesc, it != 0, ksteps);
          ptx::umma_commit(r.empty);
        }
        __syncwarp();
        dw += kWBytes >> 4;
        dx += kXBytes >> 4;
        if (r.advance()) {
          dw = dw0;
          dx = dx0;
        }
      }
      if (ptx::elect_one()) ptx::umma_commit(bar_accum);
      __syncwarp();
    }
  } else {
    // epilogue: this thread owns output channel n0 + 32*q + lane for all pixels of the tile
    const int q = warp & 3;
    const int chl = q * 32 + lane;  // channel inside the 128-channel tile
    const int ch = n0 + chl;
    const float bias = a.bias[ch];
    const float slope = a.slope[ch];
    const bool st32 = a.out_f32 != nullptr && ch < a.cout;
    const long long plane = static_cast<long long>(a.H) * a.W;
    ptx::mbar_wait(bar_accum, 0);
    ptx::tc_fence_after();
    const int npix = a.bw * a.bh;
    if (a.tma_store) {
      // The bf16 tile goes out as [pixel][channel] rows through the (now idle) activation ring: two swizzled blocks
      // of 64 channels, exactly the layout a TMA store with SWIZZLE_128B expects, then one bulk store per block.
      // Pixels outside the frame and channels beyond the slice are clipped by the tensor map.
      const uint32_t blk = sX0 + (chl >> 6) * kXBytes;       // channels 0..63 -> stage 0, 64..127 -> stage 1
      const uint32_t chunk = (chl & 63) >> 3;                 // 16-byte chunk of the 128-byte pixel row
      const uint32_t inner = (chl & 7) * 2;
      for (int c = 0; c < npix; c += 32) {
        uint32_t r[32];
        ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c, r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float acc = __uint_as_float(r[i]) + bias;
          const float v = acc > 0.f ? acc : acc * slope;
          const uint32_t row = c + i;
          const uint32_t addr = blk + row * 128u + ((chunk ^ (row & 7u)) << 4) + inner;
          const __nv_bfloat16 h = __float2bfloat16_rn(v);
          asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(*reinterpret_cast<const uint16_t*>(&h)) : "memory");
          if (st32) {
            const long long pix = s_pix[row];
            if (pix >= 0) {
              const long long im = pix / plane;
              a.out_f32[(im * a.out_f32_channels + ch) * plane + (pix - im * plane)] = v;
            }
          }
        }
      }
      ptx::fence_proxy_async_smem();
      asm volatile("bar.sync 1, 128;" ::: "memory");  // the four epilogue warps
      if (threadIdx.x == 64) {
        ptx::tma_store_4d(&tmC, sX0, n0, x0, y0, img);
        if (n0 + 64 < a.cout_store) ptx::tma_store_4d(&tmC, sX0 + kXBytes, n0 + 64, x0, y0, img);
        ptx::tma_store_commit();
        ptx::tma_store_wait_all();
      }
    } else {
      const bool st16 = a.out_bf16 != nullptr && ch < a.cout_store;
      for (int c = 0; c < npix; c += 32) {
        uint32_t r[32];
        ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c, r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const long long pix = s_pix[c + i];  // same for the whole warp: no divergence
          if (pix < 0) continue;
          const float acc = __uint_as_float(r[i]) + bias;
          const float v = acc > 0.f ? acc : acc * slope;
          // 32 lanes = 32 neighbouring channels of one pixel: one 64-byte segment per warp store
          if (st16) a.out_bf16[pix * a.out_pix_stride + ch] = __float2bfloat16_rn(v);
          if (st32) {
            const long long im = pix / plane;
            a.out_f32[(im * a.out_f32_channels + ch) * plane + (pix - im * plane)] = v;
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tmem_dealloc(tmem_base, a.tmem_cols);
  }
}

#endif  // ISLPOSE_BRINGUP_VARIANTS

// ------------------------------------------------------------------------------------------------ v5
// Swapped operands + resident activation halo, persistent. With lean issue loops the v1 / v4 kernels are bound by
// L2 -> SM operand delivery (~60 B/clk/SM: 32 KB per 128x128x64 K-block, 48 KB per 128x256x64), not by the tensor
// pipe (build/mma_rate, profiles/). For k > 1 almost all of those bytes are the same activations fetched again
// for every tap. Here a work item is 128 output channels x (8 x th) pixels, th <= 32:
//   * per 64-channel block the (8+2p) x (th+2p) input window is fetched ONCE as a TMA box of 16 x (th+2p) pixel
//     rows (pitch 16 rows = 2048 B keeps every 8-row operand group on one swizzle phase); it is the B operand
//     (N = 8*th pixels): tap (ky,kx) reads the window that starts at halo row ky*16+kx, its 8-row groups (one image
//     row of the tile each) 2048 B apart. The UMMA swizzle is a function of the absolute shared-memory address
//     (established with v3), so a window may start on any 128-byte row with base offset 0;
//   * only the 16 KB weight tile of each (tap, channel block) streams through a ring: 16 KB per 128 x 256 x 64
//     K-block = 32 B/clk/SM at the tensor pipe's full rate;
//   * halo buffers and TMEM accumulators are double buffered and the CTA is persistent (one per SM), so the halo
//     fetch of the next channel block, the epilogue of the previous work item and the weight stream all overlap
//     the MMAs. Epilogue lane = output channel; 32 lanes store 32 neighbouring channels of one pixel.
__global__ void __launch_bounds__(kThreadsV1, 1)
conv_umma_halo_swapped_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                              const ConvArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* const gbase = smem_raw + (base - raw);

  constexpr uint32_t kWBytes = 128 * 128;
  const uint32_t halo_bytes = 2048u * a.halo_h;  // 16 pixel rows x 128 B per image row of the window
  const uint32_t sH0 = base;
  const uint32_t sW0 = base + 2 * halo_bytes;
  const uint32_t ctrl = sW0 + a.stages * kWBytes;
  uint8_t* const gctrl = gbase + (ctrl - base);
  const uint32_t bar_wfull = ctrl;            // kMaxStages x 8 B
  const uint32_t bar_wempty = ctrl + 64;      // kMaxStages x 8 B
  const uint32_t bar_hfull = ctrl + 128;      // 2 x 8 B: halo buffer landed
  const uint32_t bar_hempty = ctrl + 144;     // 2 x 8 B: every tap of the channel block has read it
  const uint32_t bar_acc_full = ctrl + 160;   // 2 x 8 B
  const uint32_t bar_acc_empty = ctrl + 176;  // 2 x 8 B
  const uint32_t tmem_slot = ctrl + 192;
  volatile uint32_t* const tmem_slot_g = reinterpret_cast<volatile uint32_t*>(gctrl + 192);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cblocks = (a.cin_k16 + 3) >> 2;
  const int tiles_per_img = a.tiles_x * a.tiles_y;
  // Every CTA walks the same k*k weight tiles per channel block. Starting each CTA at a different tap keeps the
  // 148 CTAs from requesting the same 16 KB tile from the same L2 slices at the same moment (fp32 accumulation
  // order differs per CTA, deterministically).
  const int rot = a.tap_rot ? static_cast<int>(blockIdx.x % (a.ksize * a.ksize)) : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) {
      ptx::mbar_init(bar_wfull + 8 * s, 1);
      ptx::mbar_init(bar_wempty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(bar_hfull + 8 * b, 1);
      ptx::mbar_init(bar_hempty + 8 * b, 1);
      ptx::mbar_init(bar_acc_full + 8 * b, 1);
      ptx::mbar_init(bar_acc_empty + 8 * b, 8);  // one arrival per epilogue warp
    }
    ptx::mbar_fence_init();
    ptx::prefetch_tensormap(&tmX);
    ptx::prefetch_tensormap(&tmW);
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_g;
  // Programmatic dependent launch: the set-up above (barriers, TMEM, descriptor prefetch) overlapped the tail of the
  // previous layer; its output is first touched by the halo loads below, and this layer's stores come after MMAs that
  // consumed those loads, so one wait in the producer warp orders everything.
  ptx::pdl_launch_dependents();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (whole warp, one elected lane issues)
    {
      ptx::pdl_wait();
      ptx::RingPos r(bar_wfull, bar_wempty, a.stages);
      uint32_t sw = sW0;
      uint32_t hb = 0, hph = 0;
      const int taps = a.ksize * a.ksize;
      for (int w = blockIdx.x; w < a.work_items; w += gridDim.x) {
        const int ct = w % a.n_tiles;
        const int pt = w / a.n_tiles;
        const int img = pt / tiles_per_img;
        const int rr = pt - img * tiles_per_img;
        const int ty = rr / a.tiles_x;
        const int xh = (rr - ty * a.tiles_x) * 8 - a.pad;
        const int yh = ty * a.bh - a.pad;
        const int c0 = ct * 128;
        for (int cb = 0; cb < cblocks; ++cb) {
          ptx::mbar_wait(bar_hempty + 8 * hb, hph ^ 1);
          if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx(bar_hfull + 8 * hb, halo_bytes);
            ptx::tma_load_4d(sH0 + hb * halo_bytes, &tmX, bar_hfull + 8 * hb, cb * 64, xh, yh, img);
          }
          __syncwarp();
          int tap = rot;
          for (int t = 0; t < taps; ++t) {
            ptx::mbar_wait(r.empty, r.ph ^ 1);
            if (ptx::elect_one()) {
              ptx::mbar_arrive_expect_tx(r.full, kWBytes);
              ptx::tma_load_3d(sw, &tmW, r.full, cb * 64, c0, tap);
            }
            __syncwarp();
            if (++tap == taps) tap = 0;
            sw += kWBytes;
            if (r.advance()) sw = sW0;
          }
          hb ^= 1;
          if (hb == 0) hph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (whole warp, one elected lane issues)
    {
      const uint32_t idesc = ptx::umma_idesc_bf16(128, a.n_pix);
      const uint64_t dw0 = ptx::umma_desc_sw128(sW0);
      uint64_t dw = dw0;
      const uint64_t dh0 = ptx::umma_desc_sw128_strided(sH0, 2048u, 0u);
      const uint64_t dh1 = ptx::umma_desc_sw128_strided(sH0 + halo_bytes, 2048u, 0u);
      ptx::RingPos r(bar_wfull, bar_wempty, a.stages);
      uint32_t hb = 0, hph = 0;
      uint32_t buf = 0, buf_ph = 0;
      const uint32_t row_skip = static_cast<uint32_t>(16 - a.ksize) * 8u;  // descriptor units (16 B) to the next window row
      const int taps = a.ksize * a.ksize;
      const int rot_ky = rot / a.ksize, rot_kx = rot - rot_ky * a.ksize;
      const uint32_t rot_off = static_cast<uint32_t>(rot_ky * 16 + rot_kx) * 8u;
      for (int w = blockIdx.x; w < a.work_items; w += gridDim.x) {
        ptx::mbar_wait(bar_acc_empty + 8 * buf, buf_ph ^ 1);  // the epilogue has drained this accumulator
        ptx::tc_fence_after();
        const uint32_t acc = tmem_base + buf * 256u;
        uint32_t accumulate = 0;
        for (int cb = 0; cb < cblocks; ++cb) {
          ptx::mbar_wait(bar_hfull + 8 * hb, hph);
          ptx::tc_fence_after();
          const int ksteps = min(4, a.cin_k16 - cb * 4);
          const uint64_t dx0 = hb ? dh1 : dh0;
          uint64_t dx = dx0 + rot_off;
          int kx = rot_kx, tap = rot;
          for (int t = 0; t < taps; ++t) {
            ptx::mbar_wait(r.full, r.ph);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
              if (ksteps == 4) {
                ptx::umma_bf16(acc, dw, dx, idesc, accumulate);
                ptx::umma_bf16_acc(acc, dw + 2, dx + 2, idesc);
                ptx::umma_bf16_acc(acc, dw + 4, dx + 4, idesc);
                ptx::umma_bf16_acc(acc, dw + 6, dx + 6, idesc);
              } else {
                for (int k = 0; k < ksteps; ++k) ptx::umma_bf16(acc, dw + 2 * k, dx + 2 * k, idesc, accumulate | k);
              }
              ptx::umma_commit(r.empty);
            }
            __syncwarp();
            accumulate = 1;
            dx += 8;  // next tap to the right: one 128-byte halo row further
            if (++kx == a.ksize) {
              kx = 0;
              dx += row_skip;
            }
            if (++tap == taps) {  // rotated tap order wraps to the first tap
              tap = 0;
              dx = dx0;
            }
            dw += kWBytes >> 4;
            if (r.advance()) dw = dw0;
          }
          if (ptx::elect_one()) ptx::umma_commit(bar_hempty + 8 * hb);  // refill allowed once these MMAs have retired
          __syncwarp();
          hb ^= 1;
          if (hb == 0) hph ^= 1;
        }
        if (ptx::elect_one()) ptx::umma_commit(bar_acc_full + 8 * buf);
        __syncwarp();
        buf ^= 1;
        if (buf == 0) buf_ph ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..9): lane = output channel; the two
    // warps of a TMEM lane quarter take alternate 32-pixel chunks (short-K layers are bound by this loop)
    const int q = warp & 3;
    const int chl = q * 32 + lane;
    uint32_t buf = 0, buf_ph = 0;
    for (int w = blockIdx.x; w < a.work_items; w += gridDim.x) {
      const int ct = w % a.n_tiles;
      const int pt = w / a.n_tiles;
      const int img = pt / tiles_per_img;
      const int rr = pt - img * tiles_per_img;
      const int ty = rr / a.tiles_x;
      const int x0 = (rr - ty * a.tiles_x) * 8;
      const int y0 = ty * a.bh;
      const int ch = ct * 128 + chl;
      const float bias = a.bias[ch];
      const float slope = a.slope[ch];
      const bool ch_ok = ch < a.cout_store;
      __nv_bfloat16* const out = a.out_bf16 + ch;
      ptx::mbar_wait(bar_acc_full + 8 * buf, buf_ph);
      ptx::tc_fence_after();
      const uint32_t acc = tmem_base + buf * 256u + (static_cast<uint32_t>(q * 32) << 16);
      // Short-K layers (3x3 with 64..128 input channels) are bound by this loop, not by the MMAs, so it is written for
      // instruction count: one 64-bit row pointer per image row of the tile, pixel offsets i * stride that are
      // warp-uniform constants, and no per-pixel bounds checks on tiles that lie inside the frame.
      const long long pstride = a.out_pix_stride;
      const bool full_w = x0 + 8 <= a.W;
      if (a.pool) {
        // nn.MaxPool2d(2, 2) behind the activation (model.py:30-32 after conv1_2 / conv2_2 / conv3_4): a chunk holds 4
        // image rows x 8 pixels of this thread's channel, i.e. 2 x 4 complete pooling windows (tile origin and height
        // are even). max() commutes with the monotonic bf16 rounding, so the result equals pooling the stored layer.
        const int Hp = a.H >> 1, Wp = a.W >> 1;
        for (int c = ((warp - 2) >> 2) * 32; c < a.n_pix; c += 64) {
          uint32_t r[32];
          ptx::tmem_ld_32x32(acc + c, r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 4; g += 2) {
            const int yp = (y0 + (c >> 3) + g) >> 1;
            if (c + 8 * g < a.n_pix && yp < Hp && ch_ok) {
              __nv_bfloat16* const row = out + (static_cast<long long>(img * Hp + yp) * Wp + (x0 >> 1)) * pstride;
#pragma unroll
              for (int i = 0; i < 8; i += 2) {
                if ((x0 + i) >> 1 < Wp) {
                  float m = -3.0e38f;
#pragma unroll
                  for (int dy = 0; dy < 2; ++dy) {
#pragma unroll
                    for (int dx = 0; dx < 2; ++dx) {
                      const float v0 = __uint_as_float(r[8 * (g + dy) + i + dx]) + bias;
                      m = fmaxf(m, v0 > 0.f ? v0 : v0 * slope);
                    }
                  }
                  row[(i >> 1) * pstride] = __float2bfloat16_rn(m);
                }
              }
            }
          }
        }
      } else
      for (int c = ((warp - 2) >> 2) * 32; c < a.n_pix; c += 64) {
        uint32_t r[32];
        ptx::tmem_ld_32x32(acc + c, r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {  // 4 image rows of 8 pixels per 32-column chunk
          const int y = y0 + (c >> 3) + g;
          if (c + 8 * g < a.n_pix && y < a.H && ch_ok) {
            __nv_bfloat16* const row = out + (static_cast<long long>(img * a.H + y) * a.W + x0) * pstride;
            if (full_w) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float v0 = __uint_as_float(r[8 * g + i]) + bias;
                row[i * pstride] = __float2bfloat16_rn(v0 > 0.f ? v0 : v0 * slope);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                if (x0 + i < a.W) {
                  const float v0 = __uint_as_float(r[8 * g + i]) + bias;
                  row[i * pstride] = __float2bfloat16_rn(v0 > 0.f ? v0 : v0 * slope);
                }
              }
            }
          }
        }
      }
      // this warp has finished reading the accumulator: hand it back to the MMA issuer
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_acc_empty + 8 * buf);
      buf ^= 1;
      if (buf == 0) buf_ph ^= 1;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------ v6: 1x1 -> 1x1 pair
// Two chained 1x1 layers in one launch (Mconv6 -> Mconv7, conv5_4 -> conv5_5, conv6_1 -> conv6_2 of the reference's stages,
// src/model.py:57-62): a 1x1 layer is bound by moving its operands, and the pair moved the 128..512-channel intermediate
// out to L2 / HBM and back in again. Here a CTA (persistent, up to three per SM) walks 128-pixel tiles; per tile:
//   1. first layer as in v1 (TMA ring, M = 128 pixels, N = all `mid` output channels: one or two N <= 256 MMAs per K step,
//      accumulators in `mid` TMEM columns);
//   2. its epilogue (bias, ReLU / PReLU, bf16) writes the tile into the now idle operand ring as mid/64 blocks of
//      [128 pixels][64 channels] in the 128-byte-swizzled K-major layout - exactly the A operand of the second layer;
//      meanwhile the second layer's weights (64 rows, zero-filled beyond cout2) arrive behind them by TMA;
//   3. second layer: mid/64 x 4 MMAs of N = 64 into TMEM columns 0..63 (drained in step 2), then its own epilogue: bias,
//      activation, bf16 slice(s) and / or float32 planar output.
// Between tiles: the next tile's operands may enter the ring only after the second layer's MMAs have read the intermediate
// and the weights that overlay it (the producer waits for their commit), and the next first-layer MMAs overwrite TMEM
// columns 0..63 only after the eight epilogue warps have the second accumulator in registers (bar_done).
// The intermediate is rounded to bf16 exactly as the stored activation was, so results equal the two separate launches.
constexpr uint32_t kPairW2Block = 64 * 128;   // 64 weight rows x 64 channels
constexpr uint32_t kPairCtrlBytes = 256 + 2 * 512 * 4 + 2 * 64 * 4;  // barriers, bias / slope of layer 1 (<= 512), of layer 2

__global__ void __launch_bounds__(kThreadsV1, 3)
conv_umma_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmB2, const ConvArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* const gbase = smem_raw + (base - raw);

  const uint32_t sA0 = base;
  const uint32_t sB0 = base + a.stages * kASlotBytes;
  const uint32_t sI0 = base;                                   // intermediate blocks (over the idle ring)
  const uint32_t sW2 = base + a.mid_blocks * kASlotBytes;      // second layer's weight blocks
  const uint32_t ctrl = base + a.pair_ctrl_off;
  uint8_t* const gctrl = gbase + a.pair_ctrl_off;
  const uint32_t bar_full = ctrl;          // kMaxStages x 8 B
  const uint32_t bar_empty = ctrl + 64;    // kMaxStages x 8 B
  const uint32_t bar_accum = ctrl + 128;   // first layer's accumulators complete (and the ring idle)
  const uint32_t tmem_slot = ctrl + 136;
  const uint32_t bar_mid = ctrl + 144;     // intermediate written by the eight epilogue warps
  const uint32_t bar_w2 = ctrl + 152;      // second layer's weights landed
  const uint32_t bar_accum2 = ctrl + 160;  // second layer's accumulator complete
  const uint32_t bar_done = ctrl + 168;    // the eight epilogue warps have drained the second accumulator
  volatile uint32_t* const tmem_slot_g = reinterpret_cast<volatile uint32_t*>(gctrl + 136);
  float* const s_bias = reinterpret_cast<float*>(gctrl + 256);
  float* const s_slope = s_bias + 512;
  float* const s_bias2 = s_slope + 512;
  float* const s_slope2 = s_bias2 + 64;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_per_img = a.tiles_x * a.tiles_y;
  const int cblocks = (a.cin_k16 + 3) >> 2;
  const int mid = a.mid_blocks * 64;

  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) {
      ptx::mbar_init(bar_full + 8 * s, 1);
      ptx::mbar_init(bar_empty + 8 * s, 1);
    }
    ptx::mbar_init(bar_accum, 1);
    ptx::mbar_init(bar_mid, 8);  // one arrival per epilogue warp
    ptx::mbar_init(bar_w2, 1);
    ptx::mbar_init(bar_accum2, 1);
    ptx::mbar_init(bar_done, 8);
    ptx::mbar_fence_init();
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
    ptx::prefetch_tensormap(&tmB2);
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, a.tmem_cols);
    ptx::tmem_relinquish();
  }
  if (warp >= 2) {
    for (int i = threadIdx.x - 64; i < mid; i += kThreadsV1 - 64) {
      s_bias[i] = a.bias[i];
      s_slope[i] = a.slope[i];
    }
    if (threadIdx.x - 64 < 64) {
      s_bias2[threadIdx.x - 64] = a.bias2[threadIdx.x - 64];
      s_slope2[threadIdx.x - 64] = a.slope2[threadIdx.x - 64];
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_g;
  ptx::pdl_launch_dependents();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    ptx::pdl_wait();
    const uint32_t tx_bytes = 128u * a.bw * a.bh + 128u * static_cast<uint32_t>(mid);
    ptx::RingPos r(bar_full, bar_empty, a.stages);
    uint32_t sa = sA0, sb = sB0;
    uint32_t ph = 0;
    for (int t = blockIdx.x; t < a.m_tiles; t += gridDim.x, ph ^= 1) {
    const int img = t / tiles_per_img;
    const int rr = t - img * tiles_per_img;
    const int ty = rr / a.tiles_x;
    const int x0 = (rr - ty * a.tiles_x) * a.bw, y0 = ty * a.bh;
    for (int cb = 0; cb < cblocks; ++cb) {
      ptx::mbar_wait(r.empty, r.ph ^ 1);
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(r.full, tx_bytes);
        ptx::tma_load_4d(sa, &tmA, r.full, cb * 64, x0, y0, img);
        for (int p = 0; p < a.n1_parts; ++p)
          ptx::tma_load_3d(sb + static_cast<uint32_t>(p * a.n1_mma) * 128u, &tmB, r.full, cb * 64, p * a.n1_mma, 0);
      }
      __syncwarp();
      sa += kASlotBytes;
      sb += a.b_stage_bytes;
      if (r.advance()) {
        sa = sA0;
        sb = sB0;
      }
    }
    // the ring is idle once every MMA of the first layer has retired: the second layer's weights go behind the
    // intermediate blocks (they may overlap ring slots)
    ptx::mbar_wait(bar_accum, ph);
    if (ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx(bar_w2, static_cast<uint32_t>(a.mid_blocks) * kPairW2Block);
      for (int b = 0; b < a.mid_blocks; ++b) ptx::tma_load_3d(sW2 + b * kPairW2Block, &tmB2, bar_w2, b * 64, 0, 0);
    }
    __syncwarp();
    // the next tile's operands go into ring slots the intermediate and these weights occupy: not before the second
    // layer's MMAs have read them
    ptx::mbar_wait(bar_accum2, ph);
    }  // tile loop
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    {
      const uint32_t idesc = ptx::umma_idesc_bf16(128, a.n1_mma);
      const int tail_ksteps = a.cin_k16 - (cblocks - 1) * 4;
      const uint64_t da0 = ptx::umma_desc_sw128(sA0), db0 = ptx::umma_desc_sw128(sB0);
      const uint32_t a_step = kASlotBytes >> 4, b_step = a.b_stage_bytes >> 4;
      const uint32_t part_step = (static_cast<uint32_t>(a.n1_mma) * 128u) >> 4;
      uint64_t da = da0, db = db0;
      ptx::RingPos r(bar_full, bar_empty, a.stages);
      const uint32_t idesc2 = ptx::umma_idesc_bf16(128, 64);
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < a.m_tiles; t += gridDim.x, ph ^= 1) {
      if (t != static_cast<int>(blockIdx.x)) {  // the previous tile's second accumulator (columns 0..63) has been read
        ptx::mbar_wait(bar_done, ph ^ 1);
        ptx::tc_fence_after();
      }
      for (int cb = 0; cb < cblocks; ++cb) {
        ptx::mbar_wait(r.full, r.ph);
        ptx::tc_fence_after();
        const int ksteps = cb == cblocks - 1 ? tail_ksteps : 4;
        if (ptx::elect_one()) {
          for (int p = 0; p < a.n1_parts; ++p)
            issue_kblock(tmem_base + static_cast<uint32_t>(p * a.n1_mma), da, db + p * part_step, idesc, cb != 0, ksteps);
          ptx::umma_commit(r.empty);
        }
        __syncwarp();
        da += a_step;
        db += b_step;
        if (r.advance()) {
          da = da0;
          db = db0;
        }
      }
      if (ptx::elect_one()) ptx::umma_commit(bar_accum);
      __syncwarp();
    // second layer: A = the intermediate blocks, B = its weight blocks, D = TMEM columns 0..63
    ptx::mbar_wait(bar_mid, ph);
    ptx::mbar_wait(bar_w2, ph);
    ptx::tc_fence_after();
    if (ptx::elect_one()) {
      uint64_t di = ptx::umma_desc_sw128(sI0), dw = ptx::umma_desc_sw128(sW2);
      for (int b = 0; b < a.mid_blocks; ++b) {
        issue_kblock4(tmem_base, di, dw, idesc2, b != 0);
        di += kASlotBytes >> 4;
        dw += kPairW2Block >> 4;
      }
      ptx::umma_commit(bar_accum2);
    }
    __syncwarp();
      }  // tile loop
    }
  } else {
    // ------------------------------------------------------------ epilogues
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int py = row / a.bw;
    const int px = row - py * a.bw;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint32_t ph = 0;
    for (int t = blockIdx.x; t < a.m_tiles; t += gridDim.x, ph ^= 1) {
    const int img = t / tiles_per_img;
    const int rr = t - img * tiles_per_img;
    const int ty = rr / a.tiles_x;
    const int x = (rr - ty * a.tiles_x) * a.bw + px;
    const int y = ty * a.bh + py;
    const bool valid = (row < a.bw * a.bh) && (x < a.W) && (y < a.H);
    const long long pix = (static_cast<long long>(img) * a.H + y) * a.W + x;

    ptx::mbar_wait(bar_accum, ph);
    ptx::tc_fence_after();
    for (int c = half * 32; c < mid; c += 64) {
      uint32_t r[32];
      ptx::tmem_ld_32x32(lane_base + c, r);
      ptx::tmem_ld_wait();
      const uint32_t blk = sI0 + static_cast<uint32_t>(c >> 6) * kASlotBytes + static_cast<uint32_t>(row) * 128u;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t pk[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = 8 * j + 2 * e;
          float v0 = __uint_as_float(r[i]) + s_bias[c + i];
          float v1 = __uint_as_float(r[i + 1]) + s_bias[c + i + 1];
          v0 = v0 > 0.f ? v0 : v0 * s_slope[c + i];
          v1 = v1 > 0.f ? v1 : v1 * s_slope[c + i + 1];
          const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
          pk[e] = *reinterpret_cast<const uint32_t*>(&h);
        }
        const uint32_t chunk = static_cast<uint32_t>(((c & 63) >> 3) + j);
        const uint32_t addr = blk + ((chunk ^ (static_cast<uint32_t>(row) & 7u)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
      }
    }
    ptx::fence_proxy_async_smem();  // generic-proxy writes -> the tensor core's async proxy
    ptx::tc_fence_before();
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(bar_mid);

    ptx::mbar_wait(bar_accum2, ph);
    ptx::tc_fence_after();
    {
      const int c = half * 32;
      uint32_t r[32];
      ptx::tmem_ld_32x32(lane_base + c, r);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_done);  // the accumulator is in registers: the next tile may overwrite it
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float acc = __uint_as_float(r[i]) + s_bias2[c + i];
        v[i] = acc > 0.f ? acc : acc * s_slope2[c + i];
      }
      if (valid) {
#pragma unroll
        for (int dsti = 0; dsti < 2; ++dsti) {
          __nv_bfloat16* const ob = dsti ? a.out2b_bf16 : a.out2_bf16;
          if (ob == nullptr) continue;
          __nv_bfloat16* dst = ob + pix * (dsti ? a.out2b_pix_stride : a.out2_pix_stride) + c;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (c + 8 * j < a.cout2_store) {
              uint4 pk;
              __nv_bfloat162 h;
              h = __floats2bfloat162_rn(v[8 * j + 0], v[8 * j + 1]);
              pk.x = *reinterpret_cast<uint32_t*>(&h);
              h = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]);
              pk.y = *reinterpret_cast<uint32_t*>(&h);
              h = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]);
              pk.z = *reinterpret_cast<uint32_t*>(&h);
              h = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]);
              pk.w = *reinterpret_cast<uint32_t*>(&h);
              *reinterpret_cast<uint4*>(dst + 8 * j) = pk;
            }
          }
        }
        if (a.out2_f32 != nullptr) {
          const long long plane = static_cast<long long>(a.H) * a.W;
          float* dst = a.out2_f32 + (static_cast<long long>(img) * a.out2_f32_channels + c) * plane + static_cast<long long>(y) * a.W + x;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (c + i < a.cout2) dst[i * plane] = v[i];
          }
        }
      }
    }
    }  // tile loop
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tmem_dealloc(tmem_base, a.tmem_cols);
  }
}

// ------------------------------------------------------------------ host side

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  }
  return fn;
}

int fail(char* err, int errlen, const char* fmt, long long a = 0, long long b = 0, long long c = 0) {
  if (err != nullptr && errlen > 0) snprintf(err, errlen, fmt, a, b, c);
  return 1;
}

// SM count and opt-in shared memory per block of the current device (queried once per device ordinal): grids, wave models
// and ring depths follow the part the library runs on (a MIG slice or a part with fewer enabled SMs included).
struct DeviceLimits {
  int sms;
  uint32_t smem_optin;
};
DeviceLimits device_limits() {
  static DeviceLimits cache[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  DeviceLimits& l = cache[dev & 63];
  if (l.sms == 0) {
    int sms = 0, smem = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    if (cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess || smem <= 0) smem = 227 * 1024;
    l.sms = sms;
    l.smem_optin = static_cast<uint32_t>(smem);
  }
  return l;
}

}  // namespace

int conv_prepare(const ConvDesc& d, ConvLaunch* out, char* err, int errlen) {
  if (d.ksize != 1 && d.ksize != 3 && d.ksize != 7)
    return fail(err, errlen, "conv: unsupported kernel size %lld", d.ksize);
  if (d.in_c % 8 != 0 || d.in_cstride % 8 != 0 || d.in_c <= 0)
    return fail(err, errlen, "conv: input slice channels (%lld) and stride (%lld) must be multiples of 8",
                d.in_c, d.in_cstride);
  if ((reinterpret_cast<uintptr_t>(d.in) & 15) != 0 || (reinterpret_cast<uintptr_t>(d.w) & 15) != 0)
    return fail(err, errlen, "conv: input / weight pointers must be 16-byte aligned");
  if (d.out_bf16 != nullptr &&
      ((reinterpret_cast<uintptr_t>(d.out_bf16) & 15) != 0 || d.out_cstride % 8 != 0))
    return fail(err, errlen, "conv: bf16 output slice must be 16-byte aligned (stride %lld)", d.out_cstride);
  if (d.out_bf16 == nullptr && d.out_f32 == nullptr && d.w2 == nullptr) return fail(err, errlen, "conv: no output given");
  if (d.N <= 0 || d.H <= 0 || d.W <= 0 || d.cout <= 0) return fail(err, errlen, "conv: empty problem");
  EncodeTiledFn encode = get_encode_tiled();
  if (encode == nullptr) return fail(err, errlen, "conv: cuTensorMapEncodeTiled entry point not found");

  ConvArgs& a = out->args;
  memset(out, 0, sizeof(*out));
  const DeviceLimits lim = device_limits();
  // SMs this layer may occupy: all of them, or the caller's share when several plans run side by side (latency mode)
  const int nsm = (d.sm_budget > 0 && d.sm_budget < lim.sms) ? d.sm_budget : lim.sms;
  a.ksize = d.ksize;
  a.pad = (d.ksize - 1) / 2;
  a.cin_k16 = (d.in_c + 15) / 16;
  a.H = d.H;
  a.W = d.W;

  // Automatic choice (measured, profiles/conv_test_v5_r1.log): every 3x3 / 7x7 layer with at least 64 input channels and
  // more than 64 output channels runs fastest on v5 (7x7 128->128: 1419 TFLOP/s against 1277 for v4 and 1085 for v1 on a
  // 92x164x8 grid, 1245 against 938 on 69x92x8; 3x3 256->256: 1358 against 1047). 1x1 layers, the float32 network heads,
  // the 32-channel first layer stay on v1 / v2. A 64-output-channel layer (conv1_2) wastes half of the M=128 rows and still
  // beats the persistent v2 kernel, which is L2-bound on the re-fetched activations (617 against 495 TFLOP/s at 736x736x2).
  const bool auto_v5 = d.variant <= 0 && d.ksize >= 3 && d.in_c >= 64 && d.cout >= 48 && d.out_bf16 != nullptr &&
                       d.out_f32 == nullptr && d.force_n_tile <= 0 && d.force_bw <= 0;
  if (d.variant == 5 || auto_v5) {
    // swapped operands + resident halo, persistent (see the kernel): 8 x th pixel tiles, th even, <= 32
    if (d.ksize < 3) return fail(err, errlen, "conv: the halo variants need k > 1");
    if (d.out_bf16 == nullptr || d.out_f32 != nullptr) return fail(err, errlen, "conv: v5 writes bf16 slices only");
    if (d.pool && (d.H % 2 != 0 || d.W % 2 != 0)) return fail(err, errlen, "conv: fused 2x2 pooling needs even H, W (%lld x %lld)", d.H, d.W);
    a.pool = d.pool ? 1 : 0;
    const int kb5 = d.ksize * d.ksize * ((d.in_c + 63) / 64);
    const int n_ct = (d.cout + 127) / 128;
    int th = d.force_bh;
    if (th <= 0) {
      double best = -1;
      for (int h = 2; h <= 32; h += 2) {
        const long long tiles = static_cast<long long>((d.W + 7) / 8) * ((d.H + h - 1) / h) * d.N * n_ct;
        // fitted to profiles/conv_test_v5_r1.log: 121 / 138 / 166 cycles per MMA at N = 128 / 192 / 256 (tensor pipe
        // N/2 plus shared-memory contention with the weight stream), never below ~100 (issue rate of one warp)
        const double fit = 76.0 + 2.8 * h;
        const double per_mma = fit > 100.0 ? fit : 100.0;
        const double tile_const = 6000.0;  // per-tile overhead (cycles)
        const double cost = static_cast<double>((tiles + nsm - 1) / nsm) * (kb5 * 4.0 * per_mma + tile_const);
        if (best < 0 || cost < best) {
          best = cost;
          th = h;
        }
      }
    }
    if (th < 2 || th > 32 || th % 2 != 0) return fail(err, errlen, "conv: v5 tile height must be even and <= 32, got %lld", th);
    a.bw = 8;
    a.bh = th;
    a.tiles_x = (d.W + 7) / 8;
    a.tiles_y = (d.H + th - 1) / th;
    a.n_pix = 8 * th;
    a.n_tile = 128;
    a.n_tiles = n_ct;
    a.m_tiles = a.tiles_x * a.tiles_y * d.N;
    a.work_items = a.m_tiles * n_ct;
    a.halo_w = 16;
    a.halo_h = th + 2 * a.pad;
    a.cout = d.cout;
    a.cout_store = (d.cout + 7) / 8 * 8;
    a.tmem_cols = 512;
    a.b_stage_bytes = 128 * 128;
    const uint32_t fixed = 2u * 2048u * a.halo_h + 1024u + 512u;
    int stages = (d.force_stages > 0 && d.variant == 5) ? d.force_stages : static_cast<int>((lim.smem_optin - fixed) / (128u * 128u));
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages < 2) return fail(err, errlen, "conv: v5 does not fit shared memory (halo %lld rows)", a.halo_h);
    a.stages = stages;
    a.tap_rot = 0;  // measured: no effect (the weight tiles are not an L2 hot spot); kept as a kernel option
    out->smem_bytes = fixed + stages * 128u * 128u;
    out->variant = 5;
    a.out_bf16 = d.out_bf16;
    a.out_pix_stride = d.out_cstride;
    a.bias = d.bias;
    a.slope = d.slope;
    {
      int c_dim = d.in_c;
      const int c64 = (d.in_c + 63) / 64 * 64;
      if (d.in_c_readable >= c64) c_dim = c64;
      cuuint64_t gdim[4] = {static_cast<cuuint64_t>(c_dim), static_cast<cuuint64_t>(d.W), static_cast<cuuint64_t>(d.H),
                            static_cast<cuuint64_t>(d.N)};
      cuuint64_t gstr[3] = {static_cast<cuuint64_t>(d.in_cstride) * 2, static_cast<cuuint64_t>(d.in_cstride) * 2 * d.W,
                            static_cast<cuuint64_t>(d.in_cstride) * 2 * d.W * d.H};
      cuuint32_t box[4] = {64, 16, static_cast<cuuint32_t>(a.halo_h), 1};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      CUresult r = encode(&out->tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(d.in), gdim, gstr,
                          box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail(err, errlen, "conv: activation tensor map rejected (CUresult %lld)", r);
    }
    {
      const int w_cin = d.w_cin > 0 ? d.w_cin : d.in_c;
      if (w_cin < d.in_c || w_cin % 8 != 0) return fail(err, errlen, "conv: bad weight Cin stride %lld", w_cin);
      cuuint64_t gdim[3] = {static_cast<cuuint64_t>(w_cin), static_cast<cuuint64_t>(d.cout),
                            static_cast<cuuint64_t>(d.ksize * d.ksize)};
      cuuint64_t gstr[2] = {static_cast<cuuint64_t>(w_cin) * 2, static_cast<cuuint64_t>(w_cin) * 2 * d.cout};
      cuuint32_t box[3] = {64, 128, 1};
      cuuint32_t estr[3] = {1, 1, 1};
      CUresult r = encode(&out->tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<__nv_bfloat16*>(d.w), gdim, gstr,
                          box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail(err, errlen, "conv: weight tensor map rejected (CUresult %lld)", r);
    }
    out->grid = dim3(static_cast<unsigned>(a.work_items < nsm ? a.work_items : nsm), 1, 1);
    out->flops = 2.0 * d.in_c * d.cout * d.ksize * d.ksize * static_cast<double>(d.H) * d.W * d.N;
    // the opt-in to large dynamic shared memory is per device: one flag per device ordinal
    static bool attr5_dev[64] = {};
    int dev5 = 0;
    cudaGetDevice(&dev5);
    bool& attr5 = attr5_dev[dev5 & 63];
    if (!attr5) {
      if (cudaFuncSetAttribute(conv_umma_halo_swapped_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(lim.smem_optin)) !=
          cudaSuccess)
        return fail(err, errlen, "conv: cannot raise dynamic shared memory limit");
      attr5 = true;
    }
    return 0;
  }

  if (d.pool) return fail(err, errlen, "conv: fused pooling is only available in the halo variant (k >= 3, Cin >= 64, Cout >= 48, bf16 output)");
  // Legacy v1-versus-v4 model (only reached when v5 is switched off or not applicable): fitted to the first-generation
  // issue loops, where an MMA cost about 207 + N/2 cycles alone and 192 + N with two CTAs per SM; `epi` is the epilogue.
  const int kblocks = d.ksize * d.ksize * ((d.in_c + 63) / 64);
  auto estimate = [&](long long tiles, int n, double epi) {
    const double per_mma = tiles <= lim.sms ? 207.0 + n / 2.0 : 192.0 + n;
    return static_cast<double>((tiles + 295) / 296) * (kblocks * 4.0 * per_mma + epi);
  };
  auto best_box = [&](int cap, bool wave_model, int* obw, int* obh) {
    double best = -1;
    for (int w = 1; w <= cap && w <= d.W && w <= 256; ++w) {
      const int hmax = cap / w < d.H ? cap / w : d.H;
      for (int h = wave_model ? 1 : hmax; h <= hmax && h <= 256; ++h) {
        const long long tiles = static_cast<long long>((d.W + w - 1) / w) * ((d.H + h - 1) / h) * d.N;
        const double cost = wave_model ? estimate(tiles, (w * h + 15) / 16 * 16, 6000.0) + 1e-6 * tiles
                                       : static_cast<double>(tiles) + 1e-3 * (w * h);
        if (best < 0 || cost < best) {
          best = cost;
          *obw = w;
          *obh = h;
        }
      }
    }
    return best;
  };
  int want_variant = d.variant;
  int bw = d.force_bw, bh = d.force_bh;
#ifndef ISLPOSE_BRINGUP_VARIANTS
  if (want_variant == 3 || want_variant == 4)
    return fail(err, errlen, "conv: variant %lld is a bring-up kernel, built into build/conv_test only", want_variant);
#endif
  if (want_variant == 3) {
    if (d.ksize == 1) return fail(err, errlen, "conv: the halo variant needs k > 1");
    bw = 8;  // one 8-row operand group = 8 neighbouring pixels of one image row
    bh = 16;
  }
  const int pix_cap = want_variant == 4 ? 256 : 128;  // v4: the pixels are the N operand (up to 256)
  if (bw <= 0 || bh <= 0) best_box(pix_cap, want_variant == 4, &bw, &bh);
  if (bw * bh > pix_cap || bw > 256 || bh > 256) return fail(err, errlen, "conv: bad pixel tile %lldx%lld", bw, bh);
  a.bw = bw;
  a.bh = bh;
  a.tiles_x = (d.W + bw - 1) / bw;
  a.tiles_y = (d.H + bh - 1) / bh;
  const long long m_tiles = static_cast<long long>(a.tiles_x) * a.tiles_y * d.N;

  if (d.w2 != nullptr) {
    // ---- variant 6: this 1x1 layer and the 1x1 layer behind it in one launch (see the kernel)
    if (d.ksize != 1) return fail(err, errlen, "conv: a chained pair needs two 1x1 layers");
    if (d.cout % 64 != 0 || d.cout < 64 || d.cout > 512) return fail(err, errlen, "conv: pair: first layer must have 64..512 output channels in 64s, got %lld", d.cout);
    if (d.cout2 < 1 || d.cout2 > 64) return fail(err, errlen, "conv: pair: second layer must have 1..64 output channels, got %lld", d.cout2);
    if (d.out_bf16 != nullptr || d.out_f32 != nullptr) return fail(err, errlen, "conv: pair: the intermediate is not stored");
    if (d.out2_bf16 == nullptr && d.out2_f32 == nullptr) return fail(err, errlen, "conv: pair: no output given");
    if (d.bias2 == nullptr || d.slope2 == nullptr) return fail(err, errlen, "conv: pair: null bias / slope");
    if ((reinterpret_cast<uintptr_t>(d.w2) & 15) != 0) return fail(err, errlen, "conv: pair: weights must be 16-byte aligned");
    if (d.out2_bf16 != nullptr && ((reinterpret_cast<uintptr_t>(d.out2_bf16) & 15) != 0 || d.out2_cstride % 8 != 0))
      return fail(err, errlen, "conv: pair: bf16 output slice must be 16-byte aligned (stride %lld)", d.out2_cstride);
    if (d.out2b_bf16 != nullptr && ((reinterpret_cast<uintptr_t>(d.out2b_bf16) & 15) != 0 || d.out2b_cstride % 8 != 0))
      return fail(err, errlen, "conv: pair: second bf16 output slice must be 16-byte aligned (stride %lld)", d.out2b_cstride);
    const int cblocks6 = (a.cin_k16 + 3) / 4;
    a.n_tile = d.cout;
    a.n_tiles = 1;
    a.cout = d.cout;
    a.cout_store = d.cout;
    a.n1_mma = d.cout < 256 ? d.cout : 256;
    a.n1_parts = d.cout / a.n1_mma;
    a.mid_blocks = d.cout / 64;
    a.b_stage_bytes = static_cast<uint32_t>(d.cout) * 128u;
    int cols6 = 64;
    while (cols6 < d.cout) cols6 <<= 1;
    a.tmem_cols = cols6;
    a.stages = d.force_stages > 0 ? d.force_stages : (cblocks6 < 2 ? 1 : 2);
    if (a.stages > kMaxStages) a.stages = kMaxStages;
    const uint32_t ring = static_cast<uint32_t>(a.stages) * (kASlotBytes + a.b_stage_bytes);
    const uint32_t chain = static_cast<uint32_t>(a.mid_blocks) * (kASlotBytes + kPairW2Block);
    a.pair_ctrl_off = ring > chain ? ring : chain;
    out->smem_bytes = a.pair_ctrl_off + kPairCtrlBytes + 1024u;
    if (out->smem_bytes > lim.smem_optin) return fail(err, errlen, "conv: pair: shared memory budget exceeded (%lld B)", out->smem_bytes);
    a.m_tiles = static_cast<int>(m_tiles);
    a.work_items = a.m_tiles;
    a.bias = d.bias;
    a.slope = d.slope;
    a.cout2 = d.cout2;
    a.cout2_store = (d.cout2 + 7) / 8 * 8;
    a.bias2 = d.bias2;
    a.slope2 = d.slope2;
    a.out2_bf16 = d.out2_bf16;
    a.out2_pix_stride = d.out2_cstride;
    a.out2b_bf16 = d.out2b_bf16;
    a.out2b_pix_stride = d.out2b_cstride;
    a.out2_f32 = d.out2_f32;
    a.out2_f32_channels = d.out2_f32_channels;
    {
      int c_dim = d.in_c;
      const int c64 = (d.in_c + 63) / 64 * 64;
      if (d.in_c_readable >= c64) c_dim = c64;
      cuuint64_t gdim[4] = {static_cast<cuuint64_t>(c_dim), static_cast<cuuint64_t>(d.W), static_cast<cuuint64_t>(d.H),
                            static_cast<cuuint64_t>(d.N)};
      cuuint64_t gstr[3] = {static_cast<cuuint64_t>(d.in_cstride) * 2, static_cast<cuuint64_t>(d.in_cstride) * 2 * d.W,
                            static_cast<cuuint64_t>(d.in_cstride) * 2 * d.W * d.H};
      cuuint32_t box[4] = {64, static_cast<cuuint32_t>(bw), static_cast<cuuint32_t>(bh), 1};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      if (encode(&out->tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(d.in), gdim, gstr, box, estr,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return fail(err, errlen, "conv: pair: activation tensor map rejected");
    }
    {
      const int w_cin = d.w_cin > 0 ? d.w_cin : d.in_c;
      if (w_cin < d.in_c || w_cin % 8 != 0) return fail(err, errlen, "conv: bad weight Cin stride %lld", w_cin);
      cuuint64_t gdim[3] = {static_cast<cuuint64_t>(w_cin), static_cast<cuuint64_t>(d.cout), 1};
      cuuint64_t gstr[2] = {static_cast<cuuint64_t>(w_cin) * 2, static_cast<cuuint64_t>(w_cin) * 2 * d.cout};
      cuuint32_t box[3] = {64, static_cast<cuuint32_t>(a.n1_mma), 1};
      cuuint32_t estr[3] = {1, 1, 1};
      if (encode(&out->tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<__nv_bfloat16*>(d.w), gdim, gstr, box, estr,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return fail(err, errlen, "conv: pair: weight tensor map rejected");
      cuuint64_t gdim2[3] = {static_cast<cuuint64_t>(d.cout), static_cast<cuuint64_t>(d.cout2), 1};
      cuuint64_t gstr2[2] = {static_cast<cuuint64_t>(d.cout) * 2, static_cast<cuuint64_t>(d.cout) * 2 * d.cout2};
      cuuint32_t box2[3] = {64, 64, 1};
      if (encode(&out->tmB2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<__nv_bfloat16*>(d.w2), gdim2, gstr2, box2, estr,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return fail(err, errlen, "conv: pair: second weight tensor map rejected");
    }
    out->variant = 6;
    {
      // persistent: as many CTAs as fit side by side (shared memory, TMEM columns, 320 threads x 63 registers: at most 3),
      // each walking its share of the tiles - barriers, TMEM and biases are set up once per CTA instead of once per 128
      // pixels. Measured against one tile per CTA (launch by launch, tools/gpu_call24.sh): coco Mconv6+7 291 -> 254 us per
      // 10, conv5_4+5 229 -> 179 us per 2, body25 Mconv6+7 1065 -> 889 us per 6, hand Mconv6+7 206 -> 177 us per 5.
      int per_sm = static_cast<int>(lim.smem_optin / (out->smem_bytes + 1024u));
      if (per_sm * a.tmem_cols > 512) per_sm = 512 / a.tmem_cols;
      if (per_sm > 3) per_sm = 3;
      if (per_sm < 1) per_sm = 1;
      const long long slots = static_cast<long long>(nsm) * per_sm;
      out->grid = dim3(static_cast<unsigned>(m_tiles < slots ? m_tiles : slots), 1, 1);
    }
    const double px = static_cast<double>(d.H) * d.W * d.N;
    out->flops = 2.0 * d.in_c * d.cout * px + 2.0 * d.cout * d.cout2 * px;
    static bool attr6_dev[64] = {};
    int dev6 = 0;
    cudaGetDevice(&dev6);
    if (!attr6_dev[dev6 & 63]) {
      if (cudaFuncSetAttribute(conv_umma_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(lim.smem_optin)) != cudaSuccess)
        return fail(err, errlen, "conv: cannot raise dynamic shared memory limit");
      attr6_dev[dev6 & 63] = true;
    }
    return 0;
  }

  // Channel tile (UMMA N): multiple of 16, at most 256. 128 keeps a stage at 32 KB so that two CTAs fit on one SM
  // with a 3-deep ring each: the second CTA's main loop hides the first one's epilogue and pipeline fill.
  const int cout16 = (d.cout + 15) / 16 * 16;
  int n_tile = want_variant == 4 ? 128 : d.force_n_tile;
  if (n_tile <= 0) {
    // An M=128 x N=128 tile is shared-memory-bandwidth bound (each K=16 MMA reads 8 KB in 64 cycles while TMA writes
    // as much): measured ~900 TFLOP/s. N=256 halves the activation traffic per FLOP: measured 1243 TFLOP/s on
    // 3x3 512->512 with a 2-deep ring and two CTAs per SM. Use it whenever the grid still fills the machine.
    if (cout16 <= 128) n_tile = cout16;
    else n_tile = (cout16 % 256 == 0 && m_tiles * (cout16 / 256) >= lim.sms) ? 256 : 128;
  }
  if (n_tile % 16 != 0 || n_tile > 256 || n_tile < 16) return fail(err, errlen, "conv: bad channel tile %lld", n_tile);
  const int n_tiles = (cout16 + n_tile - 1) / n_tile;
  a.n_tile = n_tile;
  a.cout = d.cout;
  a.cout_store = (d.cout + 7) / 8 * 8;
  a.b_stage_bytes = (static_cast<uint32_t>(n_tile) * 128u + 1023u) & ~1023u;
  int cols = 32;
  while (cols < (n_tile + 31) / 32 * 32) cols <<= 1;
  a.tmem_cols = cols;

  // Variant: persistent CTAs pay off once every SM gets at least a couple of tiles.
  // Measured on B200 (profiles/conv_test_v2_r1.log): 128-wide channel tiles run best as one tile per CTA with two
  // CTAs per SM (v1); narrower tiles (64, 96) gain 20-65 % from the persistent variant with two pixel sub-tiles per
  // weight stage, as long as the grid is large enough to give every persistent CTA several work items.
  int variant = want_variant;
  if (variant <= 0) variant = (n_tile <= 96 && m_tiles * n_tiles >= 2 * 296) ? 2 : 1;
  int msub = 1;
  int auto_stages = 0;
  if (variant == 2) {
    msub = d.msub > 0 ? d.msub : 2;
    if (d.variant <= 0) auto_stages = 2;
    int bufs = d.acc_bufs > 0 ? d.acc_bufs : (2 * msub * n_tile <= 256 ? 2 : 1);  // keep <= 256 columns: two CTAs per SM
    if (msub < 1 || msub > 2 || bufs < 1 || bufs > 2 || bufs * msub * n_tile > 512)
      return fail(err, errlen, "conv: bad sub-tile / accumulator buffer count %lld x %lld", msub, bufs);
    int cols2 = 32;
    // the epilogue reads 32-column chunks: the last chunk of the last accumulator may overhang n_tile
    while (cols2 < (bufs * msub - 1) * n_tile + (n_tile + 31) / 32 * 32) cols2 <<= 1;
    if (cols2 > 512) return fail(err, errlen, "conv: accumulator does not fit TMEM (%lld columns)", cols2);
    a.tmem_cols = cols2;
    a.acc_bufs = bufs;
  }
  a.msub = msub;
  a.m_tiles = static_cast<int>(m_tiles);
  a.n_tiles = n_tiles;
  a.work_items = static_cast<int>((m_tiles + msub - 1) / msub) * n_tiles;
  out->variant = variant;

  if (variant == 3) {
    a.halo_w = 16;
    a.halo_h = 16 + 2 * a.pad;
    a.halo_bo_mode = d.halo_base_offset_mode;
  }
  a.debug_no_loads = d.debug_no_loads;
  if (variant == 4) {
    if (d.force_n_tile > 0 && d.force_n_tile != 128) return fail(err, errlen, "conv: the swapped variant uses 128-channel tiles");
    a.n_tile = 128;
    a.n_pix = (bw * bh + 15) / 16 * 16;
    a.tmem_cols = 256;
    a.b_stage_bytes = 128 * 128;
  }
  const uint32_t per_stage = variant == 4 ? (128u * 128u + 256u * 128u)
                                          : (variant == 3 ? 0u : msub * kASlotBytes) + a.b_stage_bytes;
  int stages = d.force_stages > 0 ? d.force_stages : auto_stages;
  if (stages <= 0) {
    uint32_t budget = variant == 2 ? 200u * 1024u : 110u * 1024u - kCtrlBytes - 1024u;  // v1: two CTAs per SM
    if (variant == 3) budget -= (128u * a.halo_w * a.halo_h + 1023u) & ~1023u;
    stages = static_cast<int>(budget / per_stage);
    if (stages < 2) stages = 2;
    const int iters = d.ksize * d.ksize * ((a.cin_k16 + 3) / 4);
    if (stages > iters && variant == 1) stages = iters;
  }
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 1) stages = 1;
  a.stages = stages;
  out->smem_bytes = stages * per_stage + kCtrlBytes + 1024;
  if (variant == 3) out->smem_bytes += (128u * a.halo_w * a.halo_h + 1023u) & ~1023u;
  if (out->smem_bytes > lim.smem_optin) return fail(err, errlen, "conv: shared memory budget exceeded (%lld B)", out->smem_bytes);

  a.out_bf16 = d.out_bf16;
  a.out_pix_stride = d.out_cstride;
  a.out_f32 = d.out_f32;
  a.out_f32_channels = d.out_f32_channels;
  a.bias = d.bias;
  a.slope = d.slope;

  // Tensor maps. A: (C, W, H, N) over the NHWC slice; B: (Cin, Cout, taps) over the packed weights.
  {
    int c_dim = d.in_c;
    const int c64 = (d.in_c + 63) / 64 * 64;
    if (d.in_c_readable >= c64) c_dim = c64;
    cuuint64_t gdim[4] = {static_cast<cuuint64_t>(c_dim), static_cast<cuuint64_t>(d.W),
                          static_cast<cuuint64_t>(d.H), static_cast<cuuint64_t>(d.N)};
    cuuint64_t gstr[3] = {static_cast<cuuint64_t>(d.in_cstride) * 2,
                          static_cast<cuuint64_t>(d.in_cstride) * 2 * d.W,
                          static_cast<cuuint64_t>(d.in_cstride) * 2 * d.W * d.H};
    cuuint32_t box[4] = {64, static_cast<cuuint32_t>(bw), static_cast<cuuint32_t>(bh), 1};
    if (variant == 3) {
      box[1] = static_cast<cuuint32_t>(a.halo_w);
      box[2] = static_cast<cuuint32_t>(a.halo_h);
    }
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&out->tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(d.in), gdim,
                        gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(err, errlen, "conv: activation tensor map rejected (CUresult %lld)", r);
  }
  {
    const int w_cin = d.w_cin > 0 ? d.w_cin : d.in_c;
    if (w_cin < d.in_c || w_cin % 8 != 0) return fail(err, errlen, "conv: bad weight Cin stride %lld", w_cin);
    cuuint64_t gdim[3] = {static_cast<cuuint64_t>(w_cin), static_cast<cuuint64_t>(d.cout),
                          static_cast<cuuint64_t>(d.ksize * d.ksize)};
    cuuint64_t gstr[2] = {static_cast<cuuint64_t>(w_cin) * 2, static_cast<cuuint64_t>(w_cin) * 2 * d.cout};
    cuuint32_t box[3] = {64, static_cast<cuuint32_t>(variant == 4 ? 128 : n_tile), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(&out->tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<__nv_bfloat16*>(d.w), gdim, gstr,
                        box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(err, errlen, "conv: weight tensor map rejected (CUresult %lld)", r);
  }

  // v1: the staging blocks (16 KB per 64 channels) must fit into the operand ring
  const bool v1_store = variant == 1 && d.out_bf16 != nullptr &&
                        static_cast<uint32_t>((n_tile + 63) / 64) * kASlotBytes <= stages * per_stage;
  if ((variant == 4 && d.out_bf16 != nullptr && stages >= 2) || v1_store) {
    cuuint64_t gdim[4] = {static_cast<cuuint64_t>(a.cout_store), static_cast<cuuint64_t>(d.W),
                          static_cast<cuuint64_t>(d.H), static_cast<cuuint64_t>(d.N)};
    cuuint64_t gstr[3] = {static_cast<cuuint64_t>(d.out_cstride) * 2, static_cast<cuuint64_t>(d.out_cstride) * 2 * d.W,
                          static_cast<cuuint64_t>(d.out_cstride) * 2 * d.W * d.H};
    cuuint32_t box[4] = {64, static_cast<cuuint32_t>(bw), static_cast<cuuint32_t>(bh), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&out->tmC, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d.out_bf16, gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(err, errlen, "conv: output tensor map rejected (CUresult %lld)", r);
    a.tma_store = d.variant == 4 && d.msub == 99 ? 0 : 1;  // msub = 99 with an explicit v4 request: scalar-store epilogue
  }

  if (variant == 2) {
    int per_sm = static_cast<int>((lim.smem_optin - 1024u) / out->smem_bytes);  // persistent CTAs that fit on one SM
    if (per_sm * a.tmem_cols > 512) per_sm = 512 / a.tmem_cols;
    if (per_sm > 2) per_sm = 2;  // 320 threads x ~100 registers: two CTAs per SM
    if (per_sm < 1) per_sm = 1;
    const int slots = nsm * per_sm;
    const int ctas = a.work_items < slots ? a.work_items : slots;
    out->grid = dim3(static_cast<unsigned>(ctas), 1, 1);
  } else {
    out->grid = dim3(static_cast<unsigned>(m_tiles), static_cast<unsigned>(n_tiles), 1);
  }
  out->flops = 2.0 * d.in_c * d.cout * d.ksize * d.ksize * static_cast<double>(d.H) * d.W * d.N;

  static bool attr_set_dev[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  bool& attr_set = attr_set_dev[dev & 63];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(lim.smem_optin));
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_umma_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(lim.smem_optin));
#ifdef ISLPOSE_BRINGUP_VARIANTS
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_umma_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(lim.smem_optin));
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_umma_swapped_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(lim.smem_optin));
#endif
    if (e != cudaSuccess) return fail(err, errlen, "conv: cannot raise dynamic shared memory limit (%lld)", e);
    attr_set = true;
  }
  return 0;
}

// v1, v2 and v5 are launched with programmatic stream serialization: their set-up (barriers, TMEM allocation,
// descriptor prefetch, bias loads) may overlap the tail of the previous launch in the stream; each of them executes
// griddepcontrol.wait before its first access to the previous layer's output.
template <typename... KArgs, typename... Args>
static int launch_pdl(void (*kernel)(KArgs...), dim3 grid, int threads, uint32_t smem, cudaStream_t stream, Args&&... args) {
  const bool no_pdl = false;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = no_pdl ? 0 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...) == cudaSuccess ? 0 : 1;
}

int conv_run(const ConvLaunch& l, cudaStream_t stream) {
  if (l.variant == 2) return launch_pdl(conv_umma_persistent_kernel, l.grid, kThreadsV1, l.smem_bytes, stream, l.tmA, l.tmB, l.args);
  if (l.variant == 5) return launch_pdl(conv_umma_halo_swapped_kernel, l.grid, kThreadsV1, l.smem_bytes, stream, l.tmA, l.tmB, l.args);
  if (l.variant == 6) return launch_pdl(conv_umma_pair_kernel, l.grid, kThreadsV1, l.smem_bytes, stream, l.tmA, l.tmB, l.tmB2, l.args);
#ifdef ISLPOSE_BRINGUP_VARIANTS
  if (l.variant == 3) {
    conv_umma_halo_kernel<<<l.grid, kThreads, l.smem_bytes, stream>>>(l.tmA, l.tmB, l.args);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
  }
  if (l.variant == 4) {
    conv_umma_swapped_kernel<<<l.grid, kThreads, l.smem_bytes, stream>>>(l.tmA, l.tmB, l.tmC, l.args);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
  }
#endif
  return launch_pdl(conv_umma_kernel, l.grid, kThreadsV1, l.smem_bytes, stream, l.tmA, l.tmB, l.tmC, l.args);
}

}  // namespace islpose
