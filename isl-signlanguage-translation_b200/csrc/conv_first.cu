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

__global__ void __launch_bounds__(128)
conv_first_kernel(const float* __restrict__ in, int N, int h, int w, const __nv_bfloat16* __restrict__ wt /*[64][32]*/,
                  const float* __restrict__ bias, const float* __restrict__ slope, __nv_bfloat16* __restrict__ out,
                  int out_cstride) {
  __shared__ __align__(1024) uint8_t s_a[128 * 128];   // A operand, later the output staging tile
  __shared__ __align__(1024) uint8_t s_b[64 * 128];    // B operand: 64 output channels x 32 K
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_tmem;
  __shared__ float s_bias[64], s_slope[64];
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
  if (t < 64) {
    s_bias[t] = bias[t];
    s_slope[t] = slope[t];
  }
  // ---- weights: row = output channel, 4 chunks of 8 bf16 (K = 32), chunk j of row r at ((j ^ (r & 7)) << 4)
  {
    const int r = t >> 1, half = t & 1;  // 128 threads x 32 B
    const uint4* src = reinterpret_cast<const uint4*>(wt + r * 32 + half * 16);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int chunk = half * 2 + j;
      *reinterpret_cast<uint4*>(s_b + r * 128 + ((chunk ^ (r & 7)) << 4)) = __ldg(src + j);
    }
  }
  uint32_t phase = 0;
  uint32_t tmem = 0;
  (void)tmem;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
  const int n = tile / (tiles_x * tiles_y);
  const int rr = tile - n * tiles_x * tiles_y;
  const int y0 = (rr / tiles_x) * kFTH, x0 = (rr % tiles_x) * kFTW;
  const int x = x0 + (t & 31), y = y0 + (t >> 5);
  // ---- this thread's pixel: 3x3x3 patch, K index (ky*3+kx)*3 + c, zero outside the image
  {
    const long long plane = static_cast<long long>(h) * w;
    const float* img = in + static_cast<long long>(n) * 3 * plane;
    __align__(16) __nv_bfloat16 v[32];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int yy = y + ky - 1, xx = x + kx - 1;
        const bool ok = yy >= 0 && yy < h && xx >= 0 && xx < w;
        const long long off = static_cast<long long>(yy) * w + xx;
#pragma unroll
        for (int c = 0; c < 3; ++c) v[(ky * 3 + kx) * 3 + c] = __float2bfloat16_rn(ok ? __ldg(img + c * plane + off) : 0.f);
      }
    }
#pragma unroll
    for (int i = 27; i < 32; ++i) v[i] = __float2bfloat16_rn(0.f);
    const uint4* src = reinterpret_cast<const uint4*>(v);
#pragma unroll
    for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(s_a + t * 128 + ((j ^ (t & 7)) << 4)) = src[j];
  }
  ptx::fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async proxy
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  tmem = s_tmem;
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
        const int ch = c + 8 * j + 2 * q;
        float a0 = __uint_as_float(r[8 * j + 2 * q]) + s_bias[ch];
        float a1 = __uint_as_float(r[8 * j + 2 * q + 1]) + s_bias[ch + 1];
        a0 = a0 > 0.f ? a0 : a0 * s_slope[ch];
        a1 = a1 > 0.f ? a1 : a1 * s_slope[ch + 1];
        const __nv_bfloat162 hh = __floats2bfloat162_rn(a0, a1);
        pk[q] = *reinterpret_cast<const uint32_t*>(&hh);
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
                      void* out, int out_cstride, cudaStream_t st) {
  const long long tiles = static_cast<long long>((w + kFTW - 1) / kFTW) * ((h + kFTH - 1) / kFTH) * N;
  if (tiles <= 0 || tiles > 0x7fffffffLL || out_cstride < 64 || out_cstride % 8 != 0) return 1;
  static int sms_of[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  int& sms = sms_of[dev & 63];
  if (sms == 0 && (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)) sms = 148;
  const long long slots = static_cast<long long>(sms) * 8;  // 24.8 KB of shared memory and 64 TMEM columns per CTA: 8 fit
  conv_first_kernel<<<static_cast<unsigned>(tiles < slots ? tiles : slots), 128, 0, st>>>(in, N, h, w, static_cast<const __nv_bfloat16*>(weights), bias,
                                                                   slope, static_cast<__nv_bfloat16*>(out), out_cstride);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace islpose
