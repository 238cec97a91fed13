// Implicit-GEMM convolution for the OpenPose CPM stacks (reference: src/model.py:25-64 builds every
// layer as nn.Conv2d(k in {1,3,7}, stride 1, pad (k-1)/2) + ReLU | PReLU | nothing).
//
// Data layout in HBM
//   activations : NHWC bf16, one buffer may hold several concatenated tensors; a layer reads a channel
//                 *slice* [c0, c0+Cin) of a buffer with `cstride` channels per pixel (this is how
//                 torch.cat in model.py:177,190,199,308-324,397-405 disappears: producers write slices).
//   weights     : bf16 [tap = ky*k+kx][Cout][Cin8]   (Cin8 = Cin rounded up to 8, zero padded)
//   outputs     : bf16 slice of an NHWC buffer and/or fp32 planar NCHW (network heads; the layout
//                 model.forward returns, model.py:207,329,407)
//
// GEMM view: D[128 pixels, Cout-tile] = sum over (tap, 64-channel block) A[128 pixels, 64] * B[Cout-tile, 64]^T
//   A tile = one TMA box (64 ch, bw, bh, 1 image) fetched at pixel offset (kx-pad, ky-pad): TMA zero-fills
//            out-of-image pixels, which *is* the convolution's zero padding, and out-of-slice channels.
//   B tile = one TMA box (64 ch, Cout-tile, 1 tap) of the packed weights.
//   Both land in 128B-swizzled K-major shared memory and feed tcgen05.mma (M=128, N=Cout-tile, K=16)
//   with the fp32 accumulator in TMEM. bw*bh may be < 128: the unused accumulator lanes are never read.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace islpose {

// One convolution layer as the host describes it.
struct ConvDesc {
  // input slice
  const __nv_bfloat16* in;  // points at channel c0 of pixel (0,0,0)
  int in_c;                 // channels in the slice (multiple of 8)
  int in_cstride;           // channels per pixel of the underlying buffer (multiple of 8)
  int in_c_readable;        // channels that may be read from the slice start (>= in_c; 0 = in_c). When the next
                            // multiple of 64 fits, every TMA box is fully in bounds (partially out-of-bounds boxes
                            // in the innermost dimension were measured ~2x slower); the extra channels meet zero weights
  int w_cin;                // Cin stride of the packed weights (>= in_c, multiple of 8; 0 = in_c)
  int N, H, W;
  // weights
  const __nv_bfloat16* w;  // [k*k][cout][in_c]
  int cout;
  int ksize;  // 1, 3 or 7
  const float* bias;   // [>= n_tiles*n_tile], zero padded
  const float* slope;  // negative-side slope per channel: 0 = ReLU, 1 = identity, else PReLU weight
  // outputs (either may be null)
  __nv_bfloat16* out_bf16;  // points at channel offset of pixel 0
  int out_cstride;          // channels per pixel of the destination buffer (multiple of 8)
  float* out_f32;           // planar NCHW [N,out_f32_channels,H,W]; channel 0 of this layer's outputs
  int out_f32_channels;
  // tuning overrides (0 = auto)
  int force_n_tile;
  int force_stages;
  int force_bw, force_bh;
  int variant;  // 0 = automatic, 1 = one tile per CTA (v1), 2 = persistent CTAs with double-buffered TMEM (v2),
                // 3 = one 8x16 tile per CTA, activation halo tile resident in shared memory (v3),
                // 4 = swapped operands: weights are the M=128 operand, up to 256 pixels are N (v4)
  int debug_no_loads;  // measurement aid (v1): after the first ring fill the producer only signals, no TMA traffic
  int halo_base_offset_mode;  // v3 bring-up switch: 0 = base offset 0 (correct on B200), 1 = start row's swizzle phase
  int msub;     // v2: pixel sub-tiles per work item sharing one weight stage (0 = automatic, 1 or 2)
  int acc_bufs; // v2: TMEM accumulator buffers (0 = automatic, 1 or 2)
  int pool;     // 1: nn.MaxPool2d(2,2) fused behind the activation (v5 only): out_bf16 is the [N,H/2,W/2] pooled buffer
  int sm_budget;  // SMs the persistent variants may occupy (0 = all): plans that run side by side each get a share
  // A second 1x1 layer chained behind this 1x1 layer in the same launch (variant 6; w2 != nullptr): this layer's cout
  // (64..512, multiple of 64) activations of a pixel tile stay in shared memory as the second layer's A operand and are
  // never written (out_bf16 / out_f32 must be null); the outputs below are the second layer's.
  const __nv_bfloat16* w2;  // [1][cout2][cout]
  int cout2;                // 1..64
  const float* bias2;
  const float* slope2;
  __nv_bfloat16* out2_bf16;   // may be null
  int out2_cstride;
  __nv_bfloat16* out2b_bf16;  // a second copy of the same slice in another buffer (may be null)
  int out2b_cstride;
  float* out2_f32;            // planar NCHW, may be null
  int out2_f32_channels;
};

// A fully resolved launch (tensor maps built once, reusable for every replay).
struct ConvArgs {
  int ksize, pad;
  int cin_k16;  // number of K=16 steps that cover the input slice
  int bw, bh;
  int tiles_x, tiles_y;
  int H, W;
  int n_tile;
  int cout;        // fp32 channels stored
  int cout_store;  // bf16 channels stored (cout rounded up to 8)
  int stages;
  int tmem_cols;
  uint32_t b_stage_bytes;
  // persistent variant
  int msub;        // pixel sub-tiles per work item
  int acc_bufs;    // accumulator buffers in TMEM (2 = epilogue overlaps the next main loop inside the CTA)
  int m_tiles;     // pixel tiles in total (tiles_x * tiles_y * N)
  int n_tiles;     // channel tiles
  int work_items;  // ceil(m_tiles / msub) * n_tiles
  // halo variant
  int halo_w, halo_h;     // halo box in pixels: 16 x (16 + 2 pad)
  int halo_bo_mode;
  int debug_no_loads;
  int n_pix;       // v4: UMMA N = pixels per tile rounded up to 16
  int tma_store;   // v4: the bf16 tile leaves through shared memory + TMA store
  int tap_rot;     // v5: CTAs start the tap loop at different taps
  int pool;        // v5: 2x2 max-pool in the epilogue
  __nv_bfloat16* out_bf16;
  long long out_pix_stride;
  float* out_f32;
  int out_f32_channels;
  const float* bias;
  const float* slope;
  // 1x1 pair (variant 6)
  int n1_mma, n1_parts;   // the first layer's N per MMA (<= 256) and MMAs per K step
  int mid_blocks;         // 64-channel blocks of the intermediate
  uint32_t pair_ctrl_off; // control block offset from the 1024-aligned base
  int cout2, cout2_store;
  const float* bias2;
  const float* slope2;
  __nv_bfloat16* out2_bf16;
  long long out2_pix_stride;
  __nv_bfloat16* out2b_bf16;
  long long out2b_pix_stride;
  float* out2_f32;
  int out2_f32_channels;
};

struct ConvLaunch {
  alignas(64) CUtensorMap tmA;
  alignas(64) CUtensorMap tmB;
  alignas(64) CUtensorMap tmC;  // v4: bf16 output slice, written with TMA stores
  alignas(64) CUtensorMap tmB2; // v6: the second layer's weights
  ConvArgs args;
  dim3 grid;
  uint32_t smem_bytes;
  int variant;
  double flops;  // algorithmic: 2*Cin*Cout*k*k*H*W*N on the unpadded slice width the caller states
};

// Returns 0 on success; on failure writes a message into err (if non-null).
int conv_prepare(const ConvDesc& d, ConvLaunch* out, char* err, int errlen);
int conv_run(const ConvLaunch& l, cudaStream_t stream);

}  // namespace islpose
