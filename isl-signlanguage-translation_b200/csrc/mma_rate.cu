// Issue-rate microbenchmark for tcgen05.mma (bf16 -> fp32, K = 16 per instruction, operands in 128-byte-swizzled
// shared memory that is never refilled, so nothing but the tensor pipe and its operand fetch is on the path).
// It answers the design questions of the conv kernels: what one instruction costs as a function of its shape,
// whether independent accumulators overlap, what two co-resident CTAs gain, and what cta_group::2 (M = 256 over
// an SM pair, each SM fetching half of B) gains.   build/mma_rate   prints one line per configuration.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "ptx.cuh"

using namespace islpose::ptx;

namespace {

constexpr int kStages = 2;
constexpr int kABytes = 128 * 128;        // 128 rows x 64 bf16
constexpr int kBBytesMax = 256 * 128;     // up to 256 rows x 64 bf16

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_ts_bf16(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}

// mode 0: cta_group::1, A and B from shared memory; mode 1: cta_group::1, A from TMEM; mode 2: cta_group::2
// (a template parameter: a kernel that merely contains cta_group::2 instructions cannot be launched without a cluster)
template <int mode>
__global__ void __launch_bounds__(128)
mma_rate_kernel(int m, int n, int kblocks, int n_acc, int tmem_cols, unsigned long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_tmem;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int warp = threadIdx.x >> 5;
  const bool pair = mode == 2;
  uint32_t rank = 0;
  if constexpr (mode == 2) rank = cluster_ctarank();
  for (int i = threadIdx.x; i < kStages * (kABytes + kBBytesMax) / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&s_bar), 1);
    mbar_fence_init();
  }
  fence_proxy_async_smem();
  __syncthreads();
  if constexpr (mode == 2) cluster_sync_all();
  if (warp == 1) {
    if constexpr (mode == 2) {
      tmem_alloc2(smem_u32(&s_tmem), tmem_cols);
      tmem_relinquish2();
    } else {
      tmem_alloc(smem_u32(&s_tmem), tmem_cols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  unsigned long long dt = 0;
  if (warp == 0 && (threadIdx.x & 31) == 0 && rank == 0) {
    const uint32_t idesc = umma_idesc_bf16(m, n);
    const uint32_t acc_cols = n;  // fp32 accumulator: one column per N
    const long long t0 = clock64();
    for (int kb = 0; kb < kblocks; ++kb) {
      const int stage = kb % kStages;
      const uint32_t a_addr = smem_u32(smem + stage * (kABytes + kBBytesMax));
      const uint32_t b_addr = a_addr + kABytes;
      const uint64_t da = umma_desc_sw128(a_addr);
      const uint64_t db = umma_desc_sw128(b_addr);
      const uint32_t d = tmem + static_cast<uint32_t>(kb % n_acc) * acc_cols;
      const uint32_t accumulate = kb >= n_acc ? 1u : 0u;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if constexpr (mode == 2) {
          umma2_bf16(d, da + k * 2, db + k * 2, idesc, accumulate | k);
        } else if constexpr (mode == 1) {
          umma_ts_bf16(d, tmem + n_acc * acc_cols + k * 8, db + k * 2, idesc, accumulate | k);
        } else {
          umma_bf16(d, da + k * 2, db + k * 2, idesc, accumulate | k);
        }
      }
    }
    if constexpr (mode == 2) umma2_commit_mc(smem_u32(&s_bar), 3); else umma_commit(smem_u32(&s_bar));
    mbar_wait(smem_u32(&s_bar), 0);
    dt = clock64() - t0;
    cycles[blockIdx.x] = dt;
  } else if (pair && rank == 1 && threadIdx.x == 0) {
    mbar_wait(smem_u32(&s_bar), 0);
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (mode == 2) cluster_sync_all();
  if (warp == 1) {
    if constexpr (mode == 2) tmem_dealloc2(tmem, tmem_cols); else tmem_dealloc(tmem, tmem_cols);
  }
}


// The conv kernels' issue protocol without any data movement: a producer lane waits for a ring slot to be released
// (tcgen05.commit -> empty barrier) and immediately marks it full; the MMA lane waits for the full barrier, issues
// the K-block and commits the slot. sync = 0: no barriers at all (4 MMAs back to back, reference);
// sync = 1: commit per K-block but nobody waits; sync = 2: the full ring protocol.
__global__ void __launch_bounds__(128)
mma_protocol_kernel(int n, int kblocks, int stages, int sync, int tmem_cols, int pattern, unsigned long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_full[8], s_empty[8], s_done;
  __shared__ uint32_t s_tmem;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kStages * (kABytes + kBBytesMax) / 16; i += blockDim.x) {
    // pattern 1: pseudo-random bf16 values in (-2, 2) (sign + exponent 0x3f + random mantissa): realistic switching activity
    uint32_t h = (i * 2654435761u) ^ (blockIdx.x * 40503u);
    uint4 v;
    h = h * 1664525u + 1013904223u; v.x = (h & 0x807f807fu) | 0x3f003f00u;
    h = h * 1664525u + 1013904223u; v.y = (h & 0x807f807fu) | 0x3f003f00u;
    h = h * 1664525u + 1013904223u; v.z = (h & 0x807f807fu) | 0x3f003f00u;
    h = h * 1664525u + 1013904223u; v.w = (h & 0x807f807fu) | 0x3f003f00u;
    reinterpret_cast<uint4*>(smem)[i] = pattern ? v : make_uint4(0, 0, 0, 0);
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < 8; ++s) {
      mbar_init(smem_u32(&s_full[s]), 1);
      mbar_init(smem_u32(&s_empty[s]), 1);
    }
    mbar_init(smem_u32(&s_done), 1);
    mbar_fence_init();
  }
  fence_proxy_async_smem();
  __syncthreads();
  if (warp == 1) {
    tmem_alloc(smem_u32(&s_tmem), tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  if (warp == 0 && lane == 0 && sync == 2) {
    RingPos r(smem_u32(&s_full[0]), smem_u32(&s_empty[0]), stages);
    for (int it = 0; it < kblocks; ++it) {
      mbar_wait(r.empty, r.ph ^ 1);
      mbar_arrive(r.full);
      r.advance();
    }
  } else if (warp == 1 && lane == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, n);
    const uint64_t da0 = umma_desc_sw128(smem_u32(smem)), db0 = umma_desc_sw128(smem_u32(smem) + kABytes);
    uint64_t da = da0, db = db0;
    RingPos r(smem_u32(&s_full[0]), smem_u32(&s_empty[0]), stages);
    const long long t0 = clock64();
    for (int it = 0; it < kblocks; ++it) {
      if (sync == 2) {
        mbar_wait(r.full, r.ph);
        tc_fence_after();
      }
      umma_bf16(tmem, da, db, idesc, it != 0 ? 1u : 0u);
      umma_bf16_acc(tmem, da + 2, db + 2, idesc);
      umma_bf16_acc(tmem, da + 4, db + 4, idesc);
      umma_bf16_acc(tmem, da + 6, db + 6, idesc);
      if (sync >= 1) umma_commit(r.empty);
      // the data ring has kStages slots whatever the barrier ring depth is (nothing is ever loaded)
      da = (r.s & 1) ? da0 : da0 + ((kABytes + kBBytesMax) >> 4);
      db = da + (kABytes >> 4);
      r.advance();
    }
    umma_commit(smem_u32(&s_done));
    mbar_wait(smem_u32(&s_done), 0);
    cycles[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, tmem_cols);
}

struct Cfg {
  const char* name;
  int mode, m, n, n_acc, ctas_per_sm;
};

}  // namespace

int main() {
  int dev = 0, sms = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const int smem_full = kStages * (kABytes + kBBytesMax) + 1024;
  cudaFuncSetAttribute(mma_rate_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
  cudaFuncSetAttribute(mma_rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
  cudaFuncSetAttribute(mma_rate_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
  unsigned long long* d_cycles = nullptr;
  cudaMalloc(&d_cycles, sizeof(unsigned long long) * 1024);
  const Cfg cfgs[] = {
      {"cta1 SS M128 N64  acc1 1cta/SM", 0, 128, 64, 1, 1},   {"cta1 SS M128 N128 acc1 1cta/SM", 0, 128, 128, 1, 1},
      {"cta1 SS M128 N256 acc1 1cta/SM", 0, 128, 256, 1, 1},  {"cta1 SS M128 N128 acc2 1cta/SM", 0, 128, 128, 2, 1},
      {"cta1 SS M128 N256 acc2 1cta/SM", 0, 128, 256, 2, 1},  {"cta1 SS M128 N128 acc4 1cta/SM", 0, 128, 128, 4, 1},
      {"cta1 SS M128 N128 acc1 2cta/SM", 0, 128, 128, 1, 2},  {"cta1 SS M128 N256 acc1 2cta/SM", 0, 128, 256, 1, 2},
      {"cta1 SS M128 N128 acc2 2cta/SM", 0, 128, 128, 2, 2},  {"cta1 TS M128 N128 acc1 1cta/SM", 1, 128, 128, 1, 1},
      {"cta1 TS M128 N256 acc1 1cta/SM", 1, 128, 256, 1, 1},  {"cta1 TS M128 N128 acc2 1cta/SM", 1, 128, 128, 2, 1},
      {"cta1 TS M128 N256 acc1 2cta/SM", 1, 128, 256, 1, 2},  {"cta2 SS M256 N128 acc1 1cta/SM", 2, 256, 128, 1, 1},
      {"cta2 SS M256 N256 acc1 1cta/SM", 2, 256, 256, 1, 1},  {"cta2 SS M256 N128 acc2 1cta/SM", 2, 256, 128, 2, 1},
      {"cta2 SS M256 N256 acc2 1cta/SM", 2, 256, 256, 2, 1},  {"cta2 SS M256 N64  acc1 1cta/SM", 2, 256, 64, 1, 1},
      {"cta2 SS M256 N128 acc1 2cta/SM", 2, 256, 128, 1, 2},  {"cta2 SS M256 N256 acc1 2cta/SM", 2, 256, 256, 1, 2},
  };
  const int kblocks = 600;
  printf("clock %.0f MHz, %d SMs; one K-block = 4 MMAs of K=16; %d K-blocks per CTA\n", khz / 1e3, sms, kblocks);
  for (const Cfg& c : cfgs) {
    int tmem_cols = 32;
    int need = c.n * c.n_acc + (c.mode == 1 ? 32 : 0);
    while (tmem_cols < need) tmem_cols <<= 1;
    if (tmem_cols > 512 || (c.ctas_per_sm == 2 && tmem_cols > 256)) {
      printf("%-34s skipped (TMEM)\n", c.name);
      continue;
    }
    const int grid = sms * c.ctas_per_sm;
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(grid);
    lc.blockDim = dim3(128);
    // co-residency is controlled through the dynamic allocation: the ring itself (97 KB) fits twice per SM,
    // 120 KB does not
    lc.dynamicSmemBytes = c.ctas_per_sm == 2 ? smem_full : 120 * 1024;
    cudaLaunchAttribute at[1];
    lc.numAttrs = 0;
    if (c.mode == 2) {
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      lc.attrs = at;
      lc.numAttrs = 1;
    }
    cudaMemset(d_cycles, 0, sizeof(unsigned long long) * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    cudaError_t err;
    if (c.mode == 2) err = cudaLaunchKernelEx(&lc, mma_rate_kernel<2>, c.m, c.n, kblocks, c.n_acc, tmem_cols, d_cycles);
    else if (c.mode == 1) err = cudaLaunchKernelEx(&lc, mma_rate_kernel<1>, c.m, c.n, kblocks, c.n_acc, tmem_cols, d_cycles);
    else err = cudaLaunchKernelEx(&lc, mma_rate_kernel<0>, c.m, c.n, kblocks, c.n_acc, tmem_cols, d_cycles);
    cudaEventRecord(e1);
    if (err == cudaSuccess) err = cudaDeviceSynchronize();
    if (err != cudaSuccess) {
      printf("%-34s FAILED: %s\n", c.name, cudaGetErrorString(err));
      return 1;
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long h[1024];
    cudaMemcpy(h, d_cycles, sizeof(h), cudaMemcpyDeviceToHost);
    double sum = 0;
    int cnt = 0;
    for (int i = 0; i < grid && i < 1024; ++i)
      if (h[i]) {
        sum += static_cast<double>(h[i]);
        ++cnt;
      }
    const double per_mma = sum / cnt / (kblocks * 4.0);
    const double flop_per_mma = 2.0 * c.m * c.n * 16;
    const int issuers = c.mode == 2 ? grid / 2 : grid;
    const double tflops = flop_per_mma * kblocks * 4.0 * issuers / (ms * 1e-3) / 1e12;
    printf("%-34s %7.1f cyc/MMA  (ideal %4d)  %7.1f TFLOP/s chip (event-timed, %.3f ms)\n", c.name, per_mma,
           c.m * c.n / (256 * (c.mode == 2 ? 2 : 1)), tflops, ms);
  }
  cudaFuncSetAttribute(mma_protocol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
  printf("-- issue protocol (cta_group::1, M128): sync 0 = none, 1 = commit per K-block, 2 = full/empty ring with a producer lane\n");
  const int ns[2] = {128, 256};
  for (int pattern = 0; pattern <= 1; ++pattern)
  for (int ni = 0; ni < 2; ++ni)
    for (int per_sm = 1; per_sm <= 2; ++per_sm)
      for (int sync = 0; sync <= 2; ++sync)
        for (int stages = 3; stages <= 6; stages += 3) {
          if (sync < 2 && stages != 3) continue;
          if (pattern == 1 && sync == 1) continue;
          const int n = ns[ni];
          const int grid = sms * per_sm;
          cudaMemset(d_cycles, 0, sizeof(unsigned long long) * 1024);
          cudaEvent_t e0, e1;
          cudaEventCreate(&e0);
          cudaEventCreate(&e1);
          cudaEventRecord(e0);
          mma_protocol_kernel<<<grid, 128, per_sm == 2 ? smem_full : 120 * 1024>>>(n, kblocks * (pattern ? 40 : 1), stages, sync, 256, pattern, d_cycles);
          cudaEventRecord(e1);
          cudaError_t err = cudaDeviceSynchronize();
          if (err != cudaSuccess) {
            printf("protocol n %d sync %d FAILED: %s\n", n, sync, cudaGetErrorString(err));
            return 1;
          }
          float ms = 0;
          cudaEventElapsedTime(&ms, e0, e1);
          unsigned long long h[1024];
          cudaMemcpy(h, d_cycles, sizeof(h), cudaMemcpyDeviceToHost);
          double sum = 0;
          for (int i = 0; i < grid; ++i) sum += static_cast<double>(h[i]);
          const double kb = kblocks * (pattern ? 40.0 : 1.0);
          printf("%s N%-3d %dcta/SM sync %d stages %d: %7.1f cyc/MMA (ideal %3d)  %7.1f TFLOP/s chip  (%.2f ms)\n",
                 pattern ? "random" : "zeros ", n, per_sm, sync, stages, sum / grid / (kb * 4.0), n / 2,
                 2.0 * 128 * n * 16 * kb * 4.0 * grid / (ms * 1e-3) / 1e12, ms);
        }
  return 0;
}
