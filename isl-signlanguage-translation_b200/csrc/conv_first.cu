// First layer of the three networks (conv1_1: 3 -> 64 channels, 3x3, ReLU; reference src/model.py:25-45 via the
// 'conv1_1' block entries) as ONE memory-bound launch: float32 NCHW network input in, bf16 NHWC activations out.
//
// The layer has K = 27, far too little arithmetic to matter (3.5 kFLOP per pixel), and was 5-6 % of a plan's time
// as two launches (a gather pass writing 64 B per pixel + a 1x1 GEMM reading them back through a persistent
// TMA pipeline whose per-tile latency chain dominated). Here a CTA owns 128 pixels (32 x 4):
//   * every thread gathers its pixel's 3x3x3 patch straight from the network input (27 coalesced loads, zero
//     padding by predicate), converts to bf16 and writes the 32-wide K row (27 + 5 zeros) into shared memory in the
//     128-byte-swizzled K-major layout tcgen05 reads (generic-proxy writes, then fence.proxy.async);
//   * the 64 x 32 weight tile is written the same way, one thread issues the two K=16 tcgen05.mma (M=128, N=64),
//     the accumulator comes back through tcgen05.ld, bias + ReLU, bf16;
//   * the output tile is staged in shared memory (XOR-swizzled 16-byte chunks, conflict free) and leaves as fully
//     coalesced 16-byte stores: 128 B per pixel, 4 KB contiguous per tile row.
// Bytes per pixel: 12 read + 128 written, nothing else. The CTAs are persistent (several per SM, each walking tiles
// round-robin): barrier, TMEM allocation and the 4 KB weight tile are set up once per CTA instead of once per 128 pixels
// (at 16 x 736 x 984 that was 90 k allocations and 370 MB of weight re-reads next to 1.6 GB of real traffic); the phases
// of neighbouring CTAs on an SM overlap each other's latencies.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "prepost.cuh"
#include "ptx.cuh"

namespace islpose {

namespace {

constexpr int kFTW = 32, kFTH = 4;  // pixel tile
constexpr int kFHW = 36;            // row pitch of the staged input halo (34 columns used)
constexpr int kFHalo = 3 * (kFTH + 2) * (kFTW + 2);  // 612 input values per tile
constexpr int kFStage = (kFHalo + 127) / 128;         // values each thread stages

// Instruction count is what bounds this layer, not bytes (ncu, profiles/r2_ncu_full_other_kernels.txt: issue-active 64 %,
// L1 80 %, DRAM 14 % with one thread gathering its 27 patch values from global memory and a scalar bias / ReLU epilogue).
// So: the tile's input halo is staged once in shared memory as bf16 (5 loads per thread instead of 27, converted once
// instead of 9 times), the patch rows are assembled from it with constant offsets, the bias rides in the GEMM as two extra
// K columns (K = 27 data + bf16 hi and lo parts of the bias against a constant 1.0, exact to 2^-17 of the bias), and ReLU
// layers (conv1_1 of the coco and hand networks; body25's has a PReLU) leave through cvt.rn.relu.bf16x2, one instruction
// per two channels.
template <bool kRelu>
__global__ void __launch_bounds__(128, 8)
conv_first_kernel(const float* __restrict__ in, int N, int h, int w, const __nv_bfloat16* __restrict__ wt /*[64][32]*/,
                  const float* __restrict__ bias, const float* __restrict__ slope, __nv_bfloat16* __restrict__ out,
                  int out_cstride) {
  __shared__ __align__(1024) uint8_t s_a[128 * 128];   // A operand, later the output staging tile
  __shared__ __align__(1024) uint8_t s_b[64 * 128];    // B operand: 64 output channels x 32 K
  __shared__ __align__(16) uint16_t s_in[3 * (kFTH + 2) * kFHW];  // input halo, bf16 bits, [c][row][col]
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_tmem;
  __shared__ float s_slope[64];
  const int t = threadIdx.x;
  const int warp = t >> 5;
  const int tiles_x = (w + kFTW - 1) / kFTW;
  const int tiles_y = (h + kFTH - 1) / kFTH;
  const int tiles = tiles_x * tiles_y * N;

  if (t == 0) {
    ptx::mbar_init(ptx::smem_u32(&s_bar), 1);
    ptx::mbar_fence_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(ptx::smem_u32(&s_tmem), 64);
    ptx::tmem_relinquish();
  }
  if (!kRelu && t < 64) s_slope[t] = slope[t];
  // ---- weights: row = output channel, 4 chunks of 8 bf16 (K = 32), chunk j of row r at ((j ^ (r & 7)) << 4);
  // K columns 27 and 28 receive the bias (hi and lo bf16 parts): the patch rows hold 1.0 there
  {
    const int r = t >> 1, half = t & 1;  // 128 threads x 32 B
    const uint4* src = reinterpret_cast<const uint4*>(wt + r * 32 + half * 16);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int chunk = half * 2 + j;
      uint4 v = __ldg(src + j);
      if (chunk == 3) {
        const float bf = __ldg(bias + r);
        const __nv_bfloat16 hi = __float2bfloat16_rn(bf);
        const __nv_bfloat16 lo = __float2bfloat16_rn(bf - __bfloat162float(hi));
        v.y = (v.y & 0x0000ffffu) | (static_cast<uint32_t>(__bfloat16_as_ushort(hi)) << 16);  // K index 27
        v.z = (v.z & 0xffff0000u) | static_cast<uint32_t>(__bfloat16_as_ushort(lo));          // K index 28
      }
      *reinterpret_cast<uint4*>(s_b + r * 128 + ((chunk ^ (r & 7)) << 4)) = v;
    }
  }
  __syncthreads();
  // ---- what this thread stages of every tile's halo: element e = t + 128 i -> (channel, halo row, halo column)
  int st_off[kFStage];
  uint32_t st_pos[kFStage];  // shared-memory index [0,10) | halo row [10,13) | halo column [13,19) | inside the halo [19]
  const long long plane = static_cast<long long>(h) * w;
#pragma unroll
  for (int i = 0; i < kFStage; ++i) {
    const int e = t + 128 * i;
    const int seg = e / (kFTW + 2), col = e - seg * (kFTW + 2);
    const int c = seg / (kFTH + 2), row = seg - c * (kFTH + 2);
    st_off[i] = static_cast<int>(c * plane) + (row - 1) * w + (col - 1);
    st_pos[i] = e < kFHalo ? static_cast<uint32_t>((c * (kFTH + 2) + row) * kFHW + col) | (row << 10) | (col << 13) | (1u << 19) : 0u;
  }
  const int tx = t & 31, ty = t >> 5;
  uint32_t phase = 0;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int n = tile / (tiles_x * tiles_y);
    const int rr = tile - n * tiles_x * tiles_y;
    const int y0 = (rr / tiles_x) * kFTH, x0 = (rr % tiles_x) * kFTW;
    // ---- stage the halo: zero outside the image (the convolution's padding)
    {
      const float* img = in + static_cast<long long>(n) * 3 * plane + static_cast<long long>(y0) * w + x0;
#pragma unroll
      for (int i = 0; i < kFStage; ++i) {
        const int yy = y0 - 1 + static_cast<int>((st_pos[i] >> 10) & 7u), xx = x0 - 1 + static_cast<int>((st_pos[i] >> 13) & 63u);
        const bool in_halo = (st_pos[i] >> 19) & 1u;
        const bool ok = in_halo && static_cast<unsigned>(yy) < static_cast<unsigned>(h) && static_cast<unsigned>(xx) < static_cast<unsigned>(w);
        const float v = ok ? __ldg(img + st_off[i]) : 0.f;
        if (in_halo) s_in[st_pos[i] & 1023u] = __bfloat16_as_ushort(__float2bfloat16_rn(v));
      }
    }
    __syncthreads();
    // ---- this thread's pixel: 3x3x3 patch, K index (ky*3+kx)*3 + c, then 1.0, 1.0 (bias columns), zeros
    {
      uint32_t k32[16];
      uint16_t v[27];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
          for (int c = 0; c < 3; ++c) v[(ky * 3 + kx) * 3 + c] = s_in[(c * (kFTH + 2) + ty + ky) * kFHW + tx + kx];
        }
      }
#pragma unroll
      for (int i = 0; i < 13; ++i) k32[i] = static_cast<uint32_t>(v[2 * i]) | (static_cast<uint32_t>(v[2 * i + 1]) << 16);
      k32[13] = static_cast<uint32_t>(v[26]) | (0x3f80u << 16);  // K 26, K 27 = 1.0
      k32[14] = 0x3f80u;                                         // K 28 = 1.0, K 29 = 0
      k32[15] = 0u;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(s_a + t * 128 + ((j ^ (t & 7)) << 4)) = make_uint4(k32[4 * j], k32[4 * j + 1], k32[4 * j + 2], k32[4 * j + 3]);
    }
    ptx::fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async proxy
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = s_tmem;
    if (t == 0) {
      const uint32_t idesc = ptx::umma_idesc_bf16(128, 64);
      const uint64_t da = ptx::umma_desc_sw128(ptx::smem_u32(s_a)), db = ptx::umma_desc_sw128(ptx::smem_u32(s_b));
      ptx::umma_bf16(tmem, da, db, idesc, 0u);
      ptx::umma_bf16(tmem, da + 2, db + 2, idesc, 1u);
      ptx::umma_commit(ptx::smem_u32(&s_bar));
    }
    ptx::mbar_wait(ptx::smem_u32(&s_bar), phase);
    phase ^= 1;
    ptx::tc_fence_after();
    // ---- epilogue: thread = pixel row t of the tile (TMEM lane t), 64 channels -> 8 chunks of 8 bf16 into the
    // staging tile (the A operand is dead: the commit above covers the reads of both MMAs)
#pragma unroll
    for (int c = 0; c < 64; c += 32) {
      uint32_t r[32];
      ptx::tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c, r);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t pk[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (kRelu) {
            asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(pk[q]) : "f"(__uint_as_float(r[8 * j + 2 * q + 1])), "f"(__uint_as_float(r[8 * j + 2 * q])));
          } else {
            const int ch = c + 8 * j + 2 * q;
            float a0 = __uint_as_float(r[8 * j + 2 * q]);
            float a1 = __uint_as_float(r[8 * j + 2 * q + 1]);
            a0 = a0 > 0.f ? a0 : a0 * s_slope[ch];
            a1 = a1 > 0.f ? a1 : a1 * s_slope[ch + 1];
            const __nv_bfloat162 hh = __floats2bfloat162_rn(a0, a1);
            pk[q] = *reinterpret_cast<const uint32_t*>(&hh);
          }
        }
        const int chunk = (c >> 3) + j;
        *reinterpret_cast<uint4*>(s_a + t * 128 + ((chunk ^ (t & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
    }
    ptx::tc_fence_before();
    __syncthreads();
    // ---- coalesced copy out: 1024 chunks of 16 B, consecutive threads = consecutive chunks of consecutive pixels
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = t + 128 * i;
      const int row = idx >> 3, chunk = idx & 7;
      const int px = x0 + (row & 31), py = y0 + (row >> 5);
      if (px < w && py < h) {
        const uint4 val = *reinterpret_cast<const uint4*>(s_a + row * 128 + ((chunk ^ (row & 7)) << 4));
        *reinterpret_cast<uint4*>(out + ((static_cast<long long>(n) * h + py) * w + px) * out_cstride + chunk * 8) = val;
      }
    }
    __syncthreads();  // the staging tile is free again: the next tile's patch rows go into the same shared memory
  }  // tile loop
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(s_tmem, 64);
}

}  // namespace

int launch_conv_first(const float* in, int N, int h, int w, const void* weights, const float* bias, const float* slope,
                      void* out, int out_cstride, bool relu, cudaStream_t st) {
  const long long tiles = static_cast<long long>((w + kFTW - 1) / kFTW) * ((h + kFTH - 1) / kFTH) * N;
  if (tiles <= 0 || tiles > 0x7fffffffLL || out_cstride < 64 || out_cstride % 8 != 0) return 1;
  static int sms_of[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  int& sms = sms_of[dev & 63];
  if (sms == 0 && (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)) sms = 148;
  const long long slots = static_cast<long long>(sms) * 8;  // 25.7 KB of shared memory and 64 TMEM columns per CTA: 8 fit
  const unsigned grid = static_cast<unsigned>(tiles < slots ? tiles : slots);
  if (relu)
    conv_first_kernel<true><<<grid, 128, 0, st>>>(in, N, h, w, static_cast<const __nv_bfloat16*>(weights), bias, slope,
                                                  static_cast<__nv_bfloat16*>(out), out_cstride);
  else
    conv_first_kernel<false><<<grid, 128, 0, st>>>(in, N, h, w, static_cast<const __nv_bfloat16*>(weights), bias, slope,
                                                   static_cast<__nv_bfloat16*>(out), out_cstride);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace islpose
