// Bring-up probe (build/tma_probe, not part of the library): one unswizzled float32 TMA box load from a pitched 2-D / 3-D
// tensor at an arbitrary inner coordinate, checked element by element against the source. Answers what the map
// accumulation's TMA-fed second stage needs to know: which (box, coordinate, rank) combinations the hardware accepts.
//   tma_probe rank box0 box1 c0 c1 pitch rows planes promo struct_param
// A timeout is reported as "timeout" (no trap), a device fault as the CUDA error string. One configuration per process.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "ptx.cuh"

using namespace islpose;

struct Maps {
  CUtensorMap m[8];
};

__device__ int g_timeout;

__device__ bool wait_flag(uint32_t bar, uint32_t parity) {
  const uint64_t t0 = ptx::globaltimer_ns();
  while (!ptx::mbar_try_wait(bar, parity)) {
    if (ptx::globaltimer_ns() - t0 > 200000000ull) return false;  // 0.2 s
  }
  return true;
}

template <bool kStruct>
__global__ void probe_kernel(const __grid_constant__ Maps maps, const __grid_constant__ CUtensorMap single, int which, int rank,
                             int box0, int box1, int c0, int c1, int c2, float* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) unsigned long long bar;
  const uint32_t b = ptx::smem_u32(&bar);
  const uint32_t dst = ptx::smem_u32(smem);
  if (threadIdx.x == 0) {
    ptx::mbar_init(b, 1);
    ptx::mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const CUtensorMap* tm = kStruct ? &maps.m[which] : &single;
    ptx::mbar_arrive_expect_tx(b, static_cast<uint32_t>(box0) * box1 * 4u);
    if (rank == 3) {
      ptx::tma_load_3d(dst, tm, b, c0, c1, c2);
    } else {
      asm volatile(
          "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
          "l"(reinterpret_cast<uint64_t>(tm)), "r"(b), "r"(c0), "r"(c1)
          : "memory");
    }
  }
  if (!wait_flag(b, 0)) {
    if (threadIdx.x == 0) g_timeout = 1;
    return;
  }
  const float* s = reinterpret_cast<const float*>(smem);
  for (int i = threadIdx.x; i < box0 * box1; i += blockDim.x) out[i] = s[i];
}

int main(int argc, char** argv) {
  if (argc < 11) {
    fprintf(stderr, "usage: tma_probe rank box0 box1 c0 c1 pitch rows planes promo struct_param\n");
    return 2;
  }
  const int rank = atoi(argv[1]), box0 = atoi(argv[2]), box1 = atoi(argv[3]), c0 = atoi(argv[4]), c1 = atoi(argv[5]);
  const int pitch = atoi(argv[6]), rows = atoi(argv[7]), planes = atoi(argv[8]), promo = atoi(argv[9]), use_struct = atoi(argv[10]);
  const int c2 = planes > 1 ? planes - 1 : 0;
  printf("rank %d box %dx%d at (%d,%d,%d) of %d x %d x %d promo %d %s: ", rank, box0, box1, c0, c1, c2, pitch, rows, planes, promo,
         use_struct ? "struct[] param" : "single param");
  fflush(stdout);
  std::vector<float> h(static_cast<size_t>(pitch) * rows * planes);
  for (size_t i = 0; i < h.size(); ++i) h[i] = static_cast<float>(i % 1000003);
  float *d = nullptr, *out = nullptr;
  cudaMalloc(&d, h.size() * 4);
  cudaMalloc(&out, static_cast<size_t>(box0) * box1 * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                         const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  Fn encode = reinterpret_cast<Fn>(fnp);
  Maps maps;
  CUtensorMap single;
  cuuint64_t gdim[3] = {static_cast<cuuint64_t>(pitch), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(planes)};
  cuuint64_t gstr[2] = {static_cast<cuuint64_t>(pitch) * 4, static_cast<cuuint64_t>(pitch) * 4 * rows};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(box0), static_cast<cuuint32_t>(box1), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapL2promotion pr = promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                               : (promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
  const CUresult r = encode(&single, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("encode failed (%d)\n", static_cast<int>(r));
    return 0;
  }
  for (int i = 0; i < 8; ++i) maps.m[i] = single;
  const size_t smem = static_cast<size_t>(box0) * box1 * 4 + 128;
  cudaFuncSetAttribute(probe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(probe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (use_struct)
    probe_kernel<true><<<1, 256, smem>>>(maps, single, 3, rank, box0, box1, c0, c1, c2, out);
  else
    probe_kernel<false><<<1, 256, smem>>>(maps, single, 3, rank, box0, box1, c0, c1, c2, out);
  const cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("FAULT: %s\n", cudaGetErrorString(e));
    return 0;
  }
  int to = 0;
  cudaMemcpyFromSymbol(&to, g_timeout, sizeof(int));
  if (to) {
    printf("timeout\n");
    return 0;
  }
  std::vector<float> o(static_cast<size_t>(box0) * box1);
  cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost);
  long long bad = 0;
  for (int y = 0; y < box1; ++y) {
    for (int x = 0; x < box0; ++x) {
      const int gx = c0 + x, gy = c1 + y;
      const float want = (gx >= 0 && gx < pitch && gy >= 0 && gy < rows)
                             ? h[(static_cast<size_t>(c2) * rows + gy) * pitch + gx]
                             : 0.0f;
      if (o[static_cast<size_t>(y) * box0 + x] != want) ++bad;
    }
  }
  printf("%s (%lld mismatches)\n", bad == 0 ? "ok" : "WRONG DATA", bad);
  return 0;
}
