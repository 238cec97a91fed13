// Weight ingestion on the device: nn.Conv2d weights float32 [cout][cin][k][k] (the flat Caffe-named tensors of the
// reference's weight files, body.py:35-36 / util.py:35-44) -> the conv kernels' operand layout, bf16
// [tap = ky*k+kx][cout][w_cin] with the input channels in the order of the activation buffer slice the layer reads
// (chan_map[i] = reference input channel at slice channel i, -1 = a pad channel: zero weights) and zero padding up to
// w_cin. conv1_1 instead packs [1][cout][32] with K index (ky*3+kx)*3 + c (27 used), what conv_first.cu expects.
// fp32 -> bf16 is round-to-nearest-even, the rounding torch's .to(torch.bfloat16) applies.
#include <cuda_bf16.h>

#include "prepost.cuh"

namespace islpose {

__global__ void pack_conv_weights_kernel(const float* __restrict__ w, int cout, int cin, int kk, const int* __restrict__ chan_map,
                                         int in_c, int w_cin, int first, __nv_bfloat16* __restrict__ out) {
  const long long total = static_cast<long long>(first ? 1 : kk) * cout * w_cin;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(i % w_cin);
    const int co = static_cast<int>((i / w_cin) % cout);
    const int tap = static_cast<int>(i / (static_cast<long long>(w_cin) * cout));
    float v = 0.f;
    if (first) {
      if (ci < kk * cin) v = w[(static_cast<long long>(co) * cin + ci % cin) * kk + ci / cin];
    } else if (ci < in_c) {
      const int src = chan_map != nullptr ? chan_map[ci] : ci;
      if (src >= 0 && src < cin) v = w[(static_cast<long long>(co) * cin + src) * kk + tap];
    }
    out[i] = __float2bfloat16_rn(v);
  }
}

int launch_pack_conv_weights(const float* w, int cout, int cin, int ksize, const int* chan_map, int in_c, int w_cin, int first,
                             void* out, cudaStream_t st) {
  const long long total = static_cast<long long>(first ? 1 : ksize * ksize) * cout * w_cin;
  long long blocks = (total + 255) / 256;
  if (blocks > 4096) blocks = 4096;
  pack_conv_weights_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(w, cout, cin, ksize * ksize, chan_map, in_c, w_cin, first,
                                                                           static_cast<__nv_bfloat16*>(out));
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace islpose
