// Stand-alone bring-up harness for the tcgen05 convolution: random bf16 NHWC input + packed weights,
// compared on the GPU against a naive direct convolution that reads the same bf16 values and
// accumulates in fp32. Prints one PASS/FAIL line per configuration and a timing.
//   build/conv_test            -> all configurations
//   build/conv_test <index>    -> one configuration
#include <stdio.h>
#include <stdlib.h>

#include <string>
#include <vector>

#include "conv_umma.cuh"

using namespace islpose;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

__global__ void ref_conv(const __nv_bfloat16* in, int in_c, int in_cstride, int N, int H, int W,
                         const __nv_bfloat16* w, int cout, int k, const float* bias, const float* slope,
                         float* out /*[N,H,W,cout]*/) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(N) * H * W * cout;
  if (idx >= total) return;
  const int co = idx % cout;
  const long long p = idx / cout;
  const int x = p % W;
  const int y = (p / W) % H;
  const int n = p / (static_cast<long long>(W) * H);
  const int pad = (k - 1) / 2;
  float acc = 0.f;
  for (int ky = 0; ky < k; ++ky) {
    const int yy = y + ky - pad;
    if (yy < 0 || yy >= H) continue;
    for (int kx = 0; kx < k; ++kx) {
      const int xx = x + kx - pad;
      if (xx < 0 || xx >= W) continue;
      const __nv_bfloat16* ip = in + ((static_cast<long long>(n) * H + yy) * W + xx) * in_cstride;
      const __nv_bfloat16* wp = w + (static_cast<long long>(ky * k + kx) * cout + co) * in_c;
      for (int c = 0; c < in_c; ++c) acc += __bfloat162float(ip[c]) * __bfloat162float(wp[c]);
    }
  }
  acc += bias[co];
  out[idx] = acc > 0.f ? acc : acc * slope[co];
}

static uint32_t rng_state = 12345u;
static float frand() {
  rng_state = rng_state * 1664525u + 1013904223u;
  return ((rng_state >> 8) & 0xffff) / 65536.0f - 0.5f;
}

struct Cfg {
  const char* name;
  int N, H, W, in_c, in_cstride, coff, cout, k;
  bool f32_out, bf16_out;
  int force_n_tile, force_stages;
  int variant, msub, acc_bufs, bo_mode, no_loads;
  int pool;  // 2x2 max-pool fused into the layer: the bf16 output is compared with the pooled reference
};

static int run_cfg(const Cfg& c, bool timing) {
  const long long npix = static_cast<long long>(c.N) * c.H * c.W;
  std::vector<__nv_bfloat16> h_in(npix * c.in_cstride);
  for (auto& v : h_in) v = __float2bfloat16(frand());
  const int taps = c.k * c.k;
  std::vector<__nv_bfloat16> h_w(static_cast<size_t>(taps) * c.cout * c.in_c);
  const float wscale = 2.0f / sqrtf(static_cast<float>(taps * c.in_c));
  for (auto& v : h_w) v = __float2bfloat16(frand() * wscale);
  std::vector<float> h_bias(512, 0.f), h_slope(512, 0.f);
  for (int i = 0; i < c.cout; ++i) {
    h_bias[i] = frand() * 0.2f;
    h_slope[i] = (i % 3 == 0) ? 0.f : ((i % 3 == 1) ? 1.f : 0.25f);
  }
  const int out_cstride = (c.cout + 7) / 8 * 8 + 16;  // destination buffer wider than the slice
  const int out_coff = 8;

  __nv_bfloat16 *d_in, *d_w, *d_out16;
  float *d_bias, *d_slope, *d_out32, *d_ref;
  CK(cudaMalloc(&d_in, h_in.size() * 2));
  CK(cudaMalloc(&d_w, h_w.size() * 2));
  CK(cudaMalloc(&d_out16, npix * out_cstride * 2));
  CK(cudaMalloc(&d_out32, npix * c.cout * 4));
  CK(cudaMalloc(&d_ref, npix * c.cout * 4));
  CK(cudaMalloc(&d_bias, 512 * 4));
  CK(cudaMalloc(&d_slope, 512 * 4));
  CK(cudaMemcpy(d_in, h_in.data(), h_in.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_w, h_w.data(), h_w.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_bias, h_bias.data(), 512 * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_slope, h_slope.data(), 512 * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(d_out16, 0x7f, npix * out_cstride * 2));
  CK(cudaMemset(d_out32, 0x7f, npix * c.cout * 4));

  ConvDesc d;
  memset(&d, 0, sizeof(d));
  d.in = d_in + c.coff;
  d.in_c = c.in_c;
  d.in_cstride = c.in_cstride;
  d.N = c.N;
  d.H = c.H;
  d.W = c.W;
  d.w = d_w;
  d.cout = c.cout;
  d.ksize = c.k;
  d.bias = d_bias;
  d.slope = d_slope;
  d.out_bf16 = c.bf16_out ? d_out16 + out_coff : nullptr;
  d.out_cstride = out_cstride;
  d.out_f32 = c.f32_out ? d_out32 : nullptr;
  d.out_f32_channels = c.cout;
  d.force_n_tile = c.force_n_tile;
  d.force_stages = c.force_stages;
  d.variant = c.variant;
  d.msub = c.msub;
  if (c.variant == 5) d.force_bh = c.msub;  // v5: the msub column carries the tile height
  d.acc_bufs = c.acc_bufs;
  d.halo_base_offset_mode = c.bo_mode;
  d.debug_no_loads = c.no_loads;
  d.pool = c.pool;

  ConvLaunch L;
  char err[256];
  if (conv_prepare(d, &L, err, sizeof(err)) != 0) {
    printf("FAIL %-28s prepare: %s\n", c.name, err);
    return 1;
  }
  if (conv_run(L, 0) != 0) {
    printf("FAIL %-28s launch: %s\n", c.name, cudaGetErrorString(cudaGetLastError()));
    return 1;
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("FAIL %-28s run: %s\n", c.name, cudaGetErrorString(e));
    return 3;  // context is dead
  }
  {
    const long long total = npix * c.cout;
    ref_conv<<<static_cast<unsigned>((total + 255) / 256), 256>>>(d_in + c.coff, c.in_c, c.in_cstride, c.N, c.H, c.W,
                                                                   d_w, c.cout, c.k, d_bias, d_slope, d_ref);
    CK(cudaDeviceSynchronize());
  }
  std::vector<float> h_ref(npix * c.cout), h_o32(npix * c.cout);
  std::vector<__nv_bfloat16> h_o16(npix * out_cstride);
  CK(cudaMemcpy(h_ref.data(), d_ref, h_ref.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(h_o32.data(), d_out32, h_o32.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(h_o16.data(), d_out16, h_o16.size() * 2, cudaMemcpyDeviceToHost));
  double max_ref = 0, err32 = 0, err16 = 0;
  long long bad_pad = 0;
  if (c.pool) {
    // the device wrote [N][H/2][W/2][out_cstride]; compare with the 2x2 max of the reference layer
    const int Hp = c.H / 2, Wp = c.W / 2;
    for (int n = 0; n < c.N; ++n)
      for (int y = 0; y < Hp; ++y)
        for (int x = 0; x < Wp; ++x) {
          const long long q = (static_cast<long long>(n) * Hp + y) * Wp + x;
          for (int co = 0; co < c.cout; ++co) {
            double m = -1e30;
            for (int dy = 0; dy < 2; ++dy)
              for (int dx = 0; dx < 2; ++dx) {
                const long long p = (static_cast<long long>(n) * c.H + 2 * y + dy) * c.W + 2 * x + dx;
                m = fmax(m, static_cast<double>(h_ref[p * c.cout + co]));
              }
            if (fabs(m) > max_ref) max_ref = fabs(m);
            const double o = __bfloat162float(h_o16[q * out_cstride + out_coff + co]);
            const double dd = fabs(o - m) / (1.0 + fabs(m));
            if (!(dd <= err16)) err16 = dd;
          }
          const uint16_t* raw = reinterpret_cast<const uint16_t*>(h_o16.data());
          for (int co = 0; co < out_coff; ++co)
            if (raw[q * out_cstride + co] != 0x7f7f) ++bad_pad;
          for (int co = out_coff + (c.cout + 7) / 8 * 8; co < out_cstride; ++co)
            if (raw[q * out_cstride + co] != 0x7f7f) ++bad_pad;
        }
  } else
  for (long long p = 0; p < npix; ++p) {
    for (int co = 0; co < c.cout; ++co) {
      const double r = h_ref[p * c.cout + co];
      if (fabs(r) > max_ref) max_ref = fabs(r);
      if (c.f32_out) {
        const long long hw = static_cast<long long>(c.H) * c.W;
        const double dd = fabs(h_o32[((p / hw) * c.cout + co) * hw + p % hw] - r);
        if (!(dd <= err32)) err32 = dd;
      }
      if (c.bf16_out) {
        const double o = __bfloat162float(h_o16[p * out_cstride + out_coff + co]);
        const double dd = fabs(o - r) / (1.0 + fabs(r));
        if (!(dd <= err16)) err16 = dd;
      }
    }
    if (c.bf16_out) {
      // pad channels of the slice must be exact zeros; channels outside the slice must be untouched
      for (int co = c.cout; co < (c.cout + 7) / 8 * 8; ++co)
        if (__bfloat162float(h_o16[p * out_cstride + out_coff + co]) != 0.f) ++bad_pad;
      const uint16_t* raw = reinterpret_cast<const uint16_t*>(h_o16.data());
      for (int co = 0; co < out_coff; ++co)
        if (raw[p * out_cstride + co] != 0x7f7f) ++bad_pad;
      for (int co = out_coff + (c.cout + 7) / 8 * 8; co < out_cstride; ++co)
        if (raw[p * out_cstride + co] != 0x7f7f) ++bad_pad;
    }
  }
  const bool ok = c.no_loads || (!c.f32_out || err32 <= 2e-3 * (1.0 + max_ref)) && (!c.bf16_out || err16 <= 1e-2) && bad_pad == 0;
  float ms = 0.f;
  if (timing && ok) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) conv_run(L, 0);
    CK(cudaEventRecord(e0));
    const int reps = 20;
    for (int i = 0; i < reps; ++i) conv_run(L, 0);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= reps;
  }
  printf("%s %-34s v%d msub %d acc %d tile %dx%d n_tile %d stages %d grid %ux%u  max|ref| %.3f  err32 %.2e  err16 %.2e  badpad %lld",
         ok ? "PASS" : "FAIL", c.name, L.variant, L.args.msub, L.args.acc_bufs, L.args.bw, L.args.bh, L.args.n_tile, L.args.stages, L.grid.x, L.grid.y,
         max_ref, err32, err16, bad_pad);
  if (ms > 0.f) printf("  %.3f ms  %.1f TFLOP/s", ms, L.flops / (ms * 1e-3) / 1e12);
  printf("\n");
  fflush(stdout);
  cudaFree(d_in);
  cudaFree(d_w);
  cudaFree(d_out16);
  cudaFree(d_out32);
  cudaFree(d_ref);
  cudaFree(d_bias);
  cudaFree(d_slope);
  return ok ? 0 : 1;
}

static int run_v2_suite() {
  // persistent variant: correctness on awkward shapes, then head-to-head timings against v1
  // name, N, H, W, in_c, cstride, coff, cout, k, f32, bf16, n_tile, stages, variant, msub, acc_bufs
  const Cfg cfgs[] = {
      {"v2 1x1 64->64 16x8", 1, 8, 16, 64, 64, 0, 64, 1, true, true, 0, 0, 2, 1, 0},
      {"v2 3x3 128->128 23x41 b3 m2", 3, 23, 41, 128, 128, 0, 128, 3, true, true, 0, 0, 2, 2, 0},
      {"v2 7x7 192->128 23x41 b2 m2 a1", 2, 23, 41, 192, 192, 0, 128, 7, true, true, 0, 2, 2, 2, 1},
      {"v2 1x1 128->512 46x62 m2", 1, 46, 62, 128, 128, 0, 512, 1, true, true, 0, 0, 2, 2, 0},
      {"v2 1x1 512->38 head f32", 2, 23, 41, 512, 512, 0, 38, 1, true, true, 0, 0, 2, 1, 0},
      {"v2 3x3 slice 96/288->96 m2", 2, 23, 41, 96, 288, 96, 96, 3, true, true, 0, 0, 2, 2, 0},
      {"v1 7x7 128->128 92x164 b8", 8, 92, 164, 128, 128, 0, 128, 7, false, true, 0, 0, 1, 0, 0},
      {"v1 7x7 128->128 92x164 b8 s2", 8, 92, 164, 128, 128, 0, 128, 7, false, true, 0, 2, 1, 0, 0},
      {"v2 7x7 .. b8 m1 s3 a2 (2/SM)", 8, 92, 164, 128, 128, 0, 128, 7, false, true, 0, 3, 2, 1, 2},
      {"v2 7x7 .. b8 m1 s2 a1 (3/SM)", 8, 92, 164, 128, 128, 0, 128, 7, false, true, 0, 2, 2, 1, 1},
      {"v2 7x7 .. b8 m2 s2 a1 (2/SM)", 8, 92, 164, 128, 128, 0, 128, 7, false, true, 0, 2, 2, 2, 1},
      {"v1 3x3 512->512 92x164 b2", 2, 92, 164, 512, 512, 0, 512, 3, false, true, 0, 0, 1, 0, 0},
      {"v2 3x3 512 .. m1 s2 a1 (3/SM)", 2, 92, 164, 512, 512, 0, 512, 3, false, true, 0, 2, 2, 1, 1},
      {"v2 3x3 512 .. m2 s2 a1 (2/SM)", 2, 92, 164, 512, 512, 0, 512, 3, false, true, 0, 2, 2, 2, 1},
      {"v1 3x3 256->256 184x328 b2", 2, 184, 328, 256, 256, 0, 256, 3, false, true, 0, 0, 1, 0, 0},
      {"v2 3x3 256 .. m2 s2 a1 (2/SM)", 2, 184, 328, 256, 256, 0, 256, 3, false, true, 0, 2, 2, 2, 1},
      {"v2 3x3 256 .. m1 s2 a1 (3/SM)", 2, 184, 328, 256, 256, 0, 256, 3, false, true, 0, 2, 2, 1, 1},
      {"v1 3x3 128->128 368x496 b2", 2, 368, 496, 128, 128, 0, 128, 3, false, true, 0, 0, 1, 0, 0},
      {"v2 3x3 128 .. m2 s2 a1 (2/SM)", 2, 368, 496, 128, 128, 0, 128, 3, false, true, 0, 2, 2, 2, 1},
      {"v2 3x3 128 .. m1 s2 a1 (3/SM)", 2, 368, 496, 128, 128, 0, 128, 3, false, true, 0, 2, 2, 1, 1},
      {"v1 3x3 64->64 736x984 b2", 2, 736, 984, 64, 64, 0, 64, 3, false, true, 0, 0, 1, 0, 0},
      {"v2 3x3 64 .. m2 s2 a2 (2/SM)", 2, 736, 984, 64, 64, 0, 64, 3, false, true, 0, 2, 2, 2, 2},
      {"v2 3x3 64 .. m2 s2 a1 (2/SM)", 2, 736, 984, 64, 64, 0, 64, 3, false, true, 0, 2, 2, 2, 1},
      {"v2 3x3 64 .. m1 s2 a2 (4/SM)", 2, 736, 984, 64, 64, 0, 64, 3, false, true, 0, 2, 2, 1, 2},
      {"v2 3x3 64 .. m2 s1 a2 (4/SM?)", 2, 736, 984, 64, 64, 0, 64, 3, false, true, 0, 1, 2, 2, 2},
      {"v1 1x1 32->64 736x984 b2", 2, 736, 984, 32, 32, 0, 64, 1, false, true, 0, 0, 1, 0, 0},
      {"v2 1x1 32->64 .. m2 s2 a2", 2, 736, 984, 32, 32, 0, 64, 1, false, true, 0, 2, 2, 2, 2},
      {"v2 1x1 32->64 .. m1 s2 a2 (4/SM)", 2, 736, 984, 32, 32, 0, 64, 1, false, true, 0, 2, 2, 1, 2},
      {"v1 7x7 192->128 60x80 b8", 8, 60, 80, 192, 192, 0, 128, 7, false, true, 0, 0, 1, 0, 0},
      {"v2 7x7 192 60x80 m1 s2 a1 (3/SM)", 8, 60, 80, 192, 192, 0, 128, 7, false, true, 0, 2, 2, 1, 1},
      {"v1 3x3 288->96 92x164 b8", 8, 92, 164, 288, 288, 0, 96, 3, false, true, 0, 0, 1, 0, 0},
      {"v2 3x3 288->96 m2 s2 a1", 8, 92, 164, 288, 288, 0, 96, 3, false, true, 0, 2, 2, 2, 1},
      {"v2 3x3 288->96 m1 s2 a1 (3/SM)", 8, 92, 164, 288, 288, 0, 96, 3, false, true, 0, 2, 2, 1, 1},
      {"v1 3x3 512->512 92x164 nt256 s2", 2, 92, 164, 512, 512, 0, 512, 3, false, true, 256, 2, 1, 0, 0},
      {"v1 3x3 256->256 184x328 nt256s2", 2, 184, 328, 256, 256, 0, 256, 3, false, true, 256, 2, 1, 0, 0},
      {"auto 3x3 64->64 736x984 b2", 2, 736, 984, 64, 64, 0, 64, 3, false, true, 0, 0, 0, 0, 0},
      {"auto 3x3 288->96 92x164 b8", 8, 92, 164, 288, 288, 0, 96, 3, false, true, 0, 0, 0, 0, 0},
      {"auto 3x3 96->96 slice 92x164 b8", 8, 92, 164, 96, 288, 96, 96, 3, false, true, 0, 0, 0, 0, 0},
      {"v1 3x3 96->96 slice 92x164 b8", 8, 92, 164, 96, 288, 96, 96, 3, false, true, 0, 0, 1, 0, 0},
      {"auto 1x1 32->64 736x984 b2", 2, 736, 984, 32, 32, 0, 64, 1, false, true, 0, 0, 0, 0, 0},
      {"auto 3x3 64->128 368x496 b2", 2, 368, 496, 64, 64, 0, 128, 3, false, true, 0, 0, 0, 0, 0},
      {"v1 1x1 128->512 92x124 b8", 8, 92, 124, 128, 128, 0, 512, 1, false, true, 0, 0, 1, 0, 0},
      {"v2 1x1 128->512 92x124 b8 m2", 8, 92, 124, 128, 128, 0, 512, 1, false, true, 0, 2, 2, 2, 1},
  };
  int fails = 0;
  for (const Cfg& c : cfgs) {
    const int r = run_cfg(c, true);
    if (r == 3) {
      printf("context lost, stopping\n");
      return 3;
    }
    fails += r;
  }
  printf("%s: %d failing configuration(s)\n", fails ? "FAILED" : "ALL PASS", fails);
  return fails ? 1 : 0;
}

static int run_v3_suite() {
  // halo variant: both base-offset conventions on a small case first (one of them is wrong by construction),
  // then correctness on awkward shapes and timings against v1
  const Cfg cfgs[] = {
      {"v3 3x3 64->64 16x8", 1, 16, 8, 64, 64, 0, 64, 3, true, true, 0, 0, 3, 0, 0, 0},
      {"v3 7x7 128->128 23x41 b2", 2, 23, 41, 128, 128, 0, 128, 7, true, true, 0, 0, 3, 0, 0, 0},
      {"v3 7x7 192->128 23x41 b2", 2, 23, 41, 192, 192, 0, 128, 7, true, true, 0, 0, 3, 0, 0, 0},
      {"v3 3x3 slice 96/288->96", 2, 23, 41, 96, 288, 96, 96, 3, true, true, 0, 0, 3, 0, 0, 0},
      {"v3 3x3 512->38 f32 46x62", 1, 46, 62, 512, 512, 0, 38, 3, true, true, 0, 0, 3, 0, 0, 0},
      {"v1 7x7 128->128 92x164 b8", 8, 92, 164, 128, 128, 0, 128, 7, false, true, 0, 0, 1, 0, 0, 0},
      {"v3 7x7 128->128 92x164 b8", 8, 92, 164, 128, 128, 0, 128, 7, false, true, 0, 0, 3, 0, 0, 0},
      {"v3 7x7 128->128 92x164 b8 s2", 8, 92, 164, 128, 128, 0, 128, 7, false, true, 0, 2, 3, 0, 0, 0},
      {"v3 7x7 128->128 92x164 b8 s6", 8, 92, 164, 128, 128, 0, 128, 7, false, true, 0, 6, 3, 0, 0, 0},
      {"v1 7x7 192->128 60x80 b8", 8, 60, 80, 192, 192, 0, 128, 7, false, true, 0, 0, 1, 0, 0, 0},
      {"v3 7x7 192->128 60x80 b8", 8, 60, 80, 192, 192, 0, 128, 7, false, true, 0, 0, 3, 0, 0, 0},
      {"v1 3x3 128->128 368x496 b2", 2, 368, 496, 128, 128, 0, 128, 3, false, true, 0, 0, 1, 0, 0, 0},
      {"v3 3x3 128->128 368x496 b2", 2, 368, 496, 128, 128, 0, 128, 3, false, true, 0, 0, 3, 0, 0, 0},
      {"v1 3x3 512->512 92x164 b2 (nt256)", 2, 92, 164, 512, 512, 0, 512, 3, false, true, 0, 0, 1, 0, 0, 0},
      {"v3 3x3 512->512 92x164 b2 nt128", 2, 92, 164, 512, 512, 0, 512, 3, false, true, 128, 0, 3, 0, 0, 0},
      {"v3 3x3 512->512 92x164 b2 nt256", 2, 92, 164, 512, 512, 0, 512, 3, false, true, 256, 0, 3, 0, 0, 0},
      {"v1 3x3 256->256 184x328 b2", 2, 184, 328, 256, 256, 0, 256, 3, false, true, 0, 0, 1, 0, 0, 0},
      {"v3 3x3 256->256 184x328 b2 nt256", 2, 184, 328, 256, 256, 0, 256, 3, false, true, 256, 0, 3, 0, 0, 0},
      {"auto 3x3 64->64 736x984 b2", 2, 736, 984, 64, 64, 0, 64, 3, false, true, 0, 0, 0, 0, 0, 0},
      {"v3 3x3 64->64 736x984 b2", 2, 736, 984, 64, 64, 0, 64, 3, false, true, 0, 0, 3, 0, 0, 0},
      {"v1 3x3 288->96 92x164 b8", 8, 92, 164, 288, 288, 0, 96, 3, false, true, 0, 0, 1, 0, 0, 0},
      {"v3 3x3 288->96 92x164 b8", 8, 92, 164, 288, 288, 0, 96, 3, false, true, 0, 0, 3, 0, 0, 0},
      {"v1 7x7 128->128 23x31 b8", 8, 23, 31, 128, 128, 0, 128, 7, false, true, 0, 0, 1, 0, 0, 0},
      {"v3 7x7 128->128 23x31 b8", 8, 23, 31, 128, 128, 0, 128, 7, false, true, 0, 0, 3, 0, 0, 0},
  };
  int fails = 0;
  for (const Cfg& c : cfgs) {
    const int r = run_cfg(c, true);
    if (r == 3) {
      printf("context lost, stopping\n");
      return 3;
    }
    fails += r;
  }
  printf("%s: %d failing configuration(s)\n", fails ? "FAILED" : "DONE", fails);
  return 0;
}

static int run_v4_suite() {
  // swapped operands (weights = M, pixels = N): correctness first, then timings against the best other variant
  const Cfg cfgs[] = {
      {"v4 1x1 64->128 16x8", 1, 8, 16, 64, 64, 0, 128, 1, true, true, 0, 0, 4, 0, 0, 0, 0},
      {"v4 3x3 128->128 23x41 b2", 2, 23, 41, 128, 128, 0, 128, 3, true, true, 0, 0, 4, 0, 0, 0, 0},
      {"v4 7x7 192->128 23x41 b2", 2, 23, 41, 192, 192, 0, 128, 7, true, true, 0, 0, 4, 0, 0, 0, 0},
      {"v4 3x3 slice 96/288->96", 2, 23, 41, 96, 288, 96, 96, 3, true, true, 0, 0, 4, 0, 0, 0, 0},
      {"v4 1x1 128->512 46x62", 1, 46, 62, 128, 128, 0, 512, 1, true, true, 0, 0, 4, 0, 0, 0, 0},
      {"v4 1x1 512->38 head f32", 2, 23, 41, 512, 512, 0, 38, 1, true, true, 0, 0, 4, 0, 0, 0, 0},
      {"v4 7x7 160->128 23x23 b3", 3, 23, 23, 160, 160, 0, 128, 7, true, true, 0, 0, 4, 0, 0, 0, 0},
      {"v1 7x7 128->128 92x164 b8", 8, 92, 164, 128, 128, 0, 128, 7, false, true, 0, 0, 1, 0, 0, 0, 0},
      {"v4 7x7 128->128 92x164 b8", 8, 92, 164, 128, 128, 0, 128, 7, false, true, 0, 0, 4, 0, 0, 0, 0},
      {"v1 7x7 192->128 60x80 b8", 8, 60, 80, 192, 192, 0, 128, 7, false, true, 0, 0, 1, 0, 0, 0, 0},
      {"v4 7x7 192->128 60x80 b8", 8, 60, 80, 192, 192, 0, 128, 7, false, true, 0, 0, 4, 0, 0, 0, 0},
      {"v1 3x3 128->128 368x496 b2", 2, 368, 496, 128, 128, 0, 128, 3, false, true, 0, 0, 1, 0, 0, 0, 0},
      {"v4 3x3 128->128 368x496 b2", 2, 368, 496, 128, 128, 0, 128, 3, false, true, 0, 0, 4, 0, 0, 0, 0},
      {"v1 3x3 64->128 368x496 b2", 2, 368, 496, 64, 64, 0, 128, 3, false, true, 0, 0, 1, 0, 0, 0, 0},
      {"v4 3x3 64->128 368x496 b2", 2, 368, 496, 64, 64, 0, 128, 3, false, true, 0, 0, 4, 0, 0, 0, 0},
      {"v1 3x3 288->96 92x164 b8", 8, 92, 164, 288, 288, 0, 96, 3, false, true, 0, 0, 1, 0, 0, 0, 0},
      {"v4 3x3 288->96 92x164 b8", 8, 92, 164, 288, 288, 0, 96, 3, false, true, 0, 0, 4, 0, 0, 0, 0},
      {"v1 3x3 96->96 slice 92x164 b8", 8, 92, 164, 96, 288, 96, 96, 3, false, true, 0, 0, 1, 0, 0, 0, 0},
      {"v4 3x3 96->96 slice 92x164 b8", 8, 92, 164, 96, 288, 96, 96, 3, false, true, 0, 0, 4, 0, 0, 0, 0},
      {"v1 3x3 512->512 92x164 b2 nt256", 2, 92, 164, 512, 512, 0, 512, 3, false, true, 0, 0, 1, 0, 0, 0, 0},
      {"v4 3x3 512->512 92x164 b2", 2, 92, 164, 512, 512, 0, 512, 3, false, true, 0, 0, 4, 0, 0, 0, 0},
      {"v1 3x3 512->128 92x92 b8", 8, 92, 92, 512, 512, 0, 128, 3, false, true, 0, 0, 1, 0, 0, 0, 0},
      {"v4 3x3 512->128 92x92 b8", 8, 92, 92, 512, 512, 0, 128, 3, false, true, 0, 0, 4, 0, 0, 0, 0},
      {"v1 1x1 128->128 92x124 b8", 8, 92, 124, 128, 128, 0, 128, 1, false, true, 0, 0, 1, 0, 0, 0, 0},
      {"v4 1x1 128->128 92x124 b8", 8, 92, 124, 128, 128, 0, 128, 1, false, true, 0, 0, 4, 0, 0, 0, 0},
      {"v1 7x7 128->128 23x31 b8", 8, 23, 31, 128, 128, 0, 128, 7, false, true, 0, 0, 1, 0, 0, 0, 0},
      {"v4 7x7 128->128 23x31 b8", 8, 23, 31, 128, 128, 0, 128, 7, false, true, 0, 0, 4, 0, 0, 0, 0},
      {"v1 7x7 128->128 46x62 b8", 8, 46, 62, 128, 128, 0, 128, 7, false, true, 0, 0, 1, 0, 0, 0, 0},
      {"v4 7x7 128->128 46x62 b8", 8, 46, 62, 128, 128, 0, 128, 7, false, true, 0, 0, 4, 0, 0, 0, 0},
  };
  int fails = 0;
  for (const Cfg& c : cfgs) {
    const int r = run_cfg(c, true);
    if (r == 3) {
      printf("context lost, stopping\n");
      return 3;
    }
    fails += r;
  }
  printf("%s: %d failing configuration(s)\n", fails ? "FAILED" : "DONE", fails);
  return 0;
}

static int g_only = -1;  // conv_test <suite> <index>: run one configuration of a suite (profiling)
static int run_v5_suite() {
  // swapped operands + resident halo, persistent: correctness on awkward shapes first, then timings against v1 / v4
  const Cfg cfgs[] = {
      {"v5 3x3 64->128 16x8", 1, 8, 16, 64, 64, 0, 128, 3, false, true, 0, 0, 5, 0, 0, 0, 0},
      {"v5 3x3 128->128 23x41 b2", 2, 23, 41, 128, 128, 0, 128, 3, false, true, 0, 0, 5, 0, 0, 0, 0},
      {"v5 7x7 192->128 23x41 b2", 2, 23, 41, 192, 192, 0, 128, 7, false, true, 0, 0, 5, 0, 0, 0, 0},
      {"v5 3x3 slice 96/288->96", 2, 23, 41, 96, 288, 96, 96, 3, false, true, 0, 0, 5, 0, 0, 0, 0},
      {"v5 7x7 160->128 23x23 b3", 3, 23, 23, 160, 160, 0, 128, 7, false, true, 0, 0, 5, 0, 0, 0, 0},
      {"v5 3x3 256->256 46x82 b2", 2, 46, 82, 256, 256, 0, 256, 3, false, true, 0, 0, 5, 0, 0, 0, 0},
      {"v5 7x7 128->38 33x17 b1", 1, 33, 17, 128, 128, 0, 38, 7, false, true, 0, 0, 5, 0, 0, 0, 0},
      {"v4 7x7 128->128 92x164 b8", 8, 92, 164, 128, 128, 0, 128, 7, false, true, 0, 0, 4, 0, 0, 0, 0},
      {"v5 7x7 128->128 92x164 b8", 8, 92, 164, 128, 128, 0, 128, 7, false, true, 0, 0, 5, 0, 0, 0, 0},
      {"v5 7x7 128 92x164 b8 th32", 8, 92, 164, 128, 128, 0, 128, 7, false, true, 0, 0, 5, 32, 0, 0, 0},
      {"v5 7x7 128 92x164 b8 th32 norot", 8, 92, 164, 128, 128, 0, 128, 7, false, true, 0, 0, 5, 32, 1, 0, 0},
      {"v5 7x7 128 92x164 b8 th24 norot", 8, 92, 164, 128, 128, 0, 128, 7, false, true, 0, 0, 5, 24, 1, 0, 0},
      {"v5 7x7 128 92x164 b8 th24 s3", 8, 92, 164, 128, 128, 0, 128, 7, false, true, 0, 3, 5, 24, 0, 0, 0},
      {"v5 7x7 128 92x164 b8 th16", 8, 92, 164, 128, 128, 0, 128, 7, false, true, 0, 0, 5, 16, 0, 0, 0},
      {"v5 7x7 128 92x164 b8 th16 s4", 8, 92, 164, 128, 128, 0, 128, 7, false, true, 0, 4, 5, 16, 0, 0, 0},
      {"v4 7x7 192->128 60x80 b8", 8, 60, 80, 192, 192, 0, 128, 7, false, true, 0, 0, 4, 0, 0, 0, 0},
      {"v5 7x7 192->128 60x80 b8", 8, 60, 80, 192, 192, 0, 128, 7, false, true, 0, 0, 5, 0, 0, 0, 0},
      {"v1 3x3 128->128 368x496 b2", 2, 368, 496, 128, 128, 0, 128, 3, false, true, 0, 0, 1, 0, 0, 0, 0},
      {"v5 3x3 128->128 368x496 b2", 2, 368, 496, 128, 128, 0, 128, 3, false, true, 0, 0, 5, 0, 0, 0, 0},
      {"v1 3x3 64->128 368x496 b2", 2, 368, 496, 64, 64, 0, 128, 3, false, true, 0, 0, 1, 0, 0, 0, 0},
      {"v5 3x3 64->128 368x496 b2", 2, 368, 496, 64, 64, 0, 128, 3, false, true, 0, 0, 5, 0, 0, 0, 0},
      {"v1 3x3 288->96 92x164 b8", 8, 92, 164, 288, 288, 0, 96, 3, false, true, 0, 0, 1, 0, 0, 0, 0},
      {"v5 3x3 288->96 92x164 b8", 8, 92, 164, 288, 288, 0, 96, 3, false, true, 0, 0, 5, 0, 0, 0, 0},
      {"v1 3x3 96->96 slice 92x164 b8", 8, 92, 164, 96, 288, 96, 96, 3, false, true, 0, 0, 1, 0, 0, 0, 0},
      {"v5 3x3 96->96 slice 92x164 b8", 8, 92, 164, 96, 288, 96, 96, 3, false, true, 0, 0, 5, 0, 0, 0, 0},
      {"v1 3x3 128->128 92x164 b8", 8, 92, 164, 128, 128, 0, 128, 3, false, true, 0, 0, 1, 0, 0, 0, 0},
      {"v5 3x3 128->128 92x164 b8", 8, 92, 164, 128, 128, 0, 128, 3, false, true, 0, 0, 5, 0, 0, 0, 0},
      {"v1 3x3 512->512 92x164 b2 nt256", 2, 92, 164, 512, 512, 0, 512, 3, false, true, 0, 0, 1, 0, 0, 0, 0},
      {"v5 3x3 512->512 92x164 b2", 2, 92, 164, 512, 512, 0, 512, 3, false, true, 0, 0, 5, 0, 0, 0, 0},
      {"v1 3x3 256->256 184x328 b2", 2, 184, 328, 256, 256, 0, 256, 3, false, true, 0, 0, 1, 0, 0, 0, 0},
      {"v5 3x3 256->256 184x328 b2", 2, 184, 328, 256, 256, 0, 256, 3, false, true, 0, 0, 5, 0, 0, 0, 0},
      {"v4 3x3 512->128 92x92 b8", 8, 92, 92, 512, 512, 0, 128, 3, false, true, 0, 0, 4, 0, 0, 0, 0},
      {"v5 3x3 512->128 92x92 b8", 8, 92, 92, 512, 512, 0, 128, 3, false, true, 0, 0, 5, 0, 0, 0, 0},
      {"v1 7x7 128->128 23x31 b8", 8, 23, 31, 128, 128, 0, 128, 7, false, true, 0, 0, 1, 0, 0, 0, 0},
      {"v5 7x7 128->128 23x31 b8", 8, 23, 31, 128, 128, 0, 128, 7, false, true, 0, 0, 5, 0, 0, 0, 0},
      {"v1 7x7 128->128 46x62 b8", 8, 46, 62, 128, 128, 0, 128, 7, false, true, 0, 0, 1, 0, 0, 0, 0},
      {"v5 7x7 128->128 46x62 b8", 8, 46, 62, 128, 128, 0, 128, 7, false, true, 0, 0, 5, 0, 0, 0, 0},
      {"v1 7x7 128->128 69x92 b8", 8, 69, 92, 128, 128, 0, 128, 7, false, true, 0, 0, 1, 0, 0, 0, 0},
      {"v5 7x7 128->128 69x92 b8", 8, 69, 92, 128, 128, 0, 128, 7, false, true, 0, 0, 5, 0, 0, 0, 0},
      {"v5 7x7 128->128 23x31 b2", 2, 23, 31, 128, 128, 0, 128, 7, false, true, 0, 0, 5, 0, 0, 0, 0},
      {"v5 pool 3x3 64->64 46x62 b2", 2, 46, 62, 64, 64, 0, 64, 3, false, true, 0, 0, 5, 0, 0, 0, 0, 1},
      {"v5 pool 3x3 128->128 40x24 b3", 3, 40, 24, 128, 128, 0, 128, 3, false, true, 0, 0, 5, 0, 0, 0, 0, 1},
      {"v5 pool 3x3 256->256 92x124 b2", 2, 92, 124, 256, 256, 0, 256, 3, false, true, 0, 0, 5, 0, 0, 0, 0, 1},
      {"auto 3x3 64->64 368x496 b8", 8, 368, 496, 64, 64, 0, 64, 3, false, true, 0, 0, 0, 0, 0, 0, 0},
      {"v5 3x3 64->64 368x496 b8", 8, 368, 496, 64, 64, 0, 64, 3, false, true, 0, 0, 5, 0, 0, 0, 0},
      {"auto 3x3 64->64 736x736 b2", 2, 736, 736, 64, 64, 0, 64, 3, false, true, 0, 0, 0, 0, 0, 0, 0},
      {"v5 3x3 64->64 736x736 b2", 2, 736, 736, 64, 64, 0, 64, 3, false, true, 0, 0, 5, 0, 0, 0, 0},
      {"v1 7x7 128->128 23x31 b2", 2, 23, 31, 128, 128, 0, 128, 7, false, true, 0, 0, 1, 0, 0, 0, 0},
  };
  int fails = 0;
  int index = -1;
  for (const Cfg& c : cfgs) {
    if (g_only >= 0 && ++index != g_only) continue;
    const int r = run_cfg(c, true);
    if (r == 3) {
      printf("context lost, stopping\n");
      return 3;
    }
    fails += r;
  }
  printf("%s: %d failing configuration(s)\n", fails ? "FAILED" : "DONE", fails);
  return fails ? 1 : 0;
}

static int run_gemm_suite() {
  // 1x1 layers that are real GEMMs (body25 Mconv6: 384 -> 512 and 288 -> 256 on the stride-8 grid): which structure wins?
  const Cfg cfgs[] = {
      {"auto 1x1 384->512 92x164 b8", 8, 92, 164, 384, 384, 0, 512, 1, false, true, 0, 0, 0, 0, 0},
      {"v1 nt256 s2", 8, 92, 164, 384, 384, 0, 512, 1, false, true, 256, 2, 1, 0, 0},
      {"v1 nt128 s3", 8, 92, 164, 384, 384, 0, 512, 1, false, true, 128, 3, 1, 0, 0},
      {"v2 nt256 m1 a2 s4", 8, 92, 164, 384, 384, 0, 512, 1, false, true, 256, 4, 2, 1, 2},
      {"v2 nt256 m1 a2 s3", 8, 92, 164, 384, 384, 0, 512, 1, false, true, 256, 3, 2, 1, 2},
      {"v2 nt128 m2 a2 s4", 8, 92, 164, 384, 384, 0, 512, 1, false, true, 128, 4, 2, 2, 2},
      {"v2 nt128 m2 a2 s3", 8, 92, 164, 384, 384, 0, 512, 1, false, true, 128, 3, 2, 2, 2},
      {"v2 nt256 m2 a1 s3", 8, 92, 164, 384, 384, 0, 512, 1, false, true, 256, 3, 2, 2, 1},
      {"v2 nt128 m1 a2 s4 (2/SM)", 8, 92, 164, 384, 384, 0, 512, 1, false, true, 128, 4, 2, 1, 2},
      {"auto 1x1 288->256 92x164 b8", 8, 92, 164, 288, 288, 0, 256, 1, false, true, 0, 0, 0, 0, 0},
      {"v2 nt256 m1 a2 s4 288->256", 8, 92, 164, 288, 288, 0, 256, 1, false, true, 256, 4, 2, 1, 2},
      {"v2 nt128 m2 a2 s4 288->256", 8, 92, 164, 288, 288, 0, 256, 1, false, true, 128, 4, 2, 2, 2},
      {"auto 1x1 128->512 92x124 b16", 16, 92, 124, 128, 128, 0, 512, 1, false, true, 0, 0, 0, 0, 0},
      {"v2 nt128 m2 a2 s4 128->512", 16, 92, 124, 128, 128, 0, 512, 1, false, true, 128, 4, 2, 2, 2},
      {"auto 1x1 128->128 92x124 b16", 16, 92, 124, 128, 128, 0, 128, 1, false, true, 0, 0, 0, 0, 0},
      {"v2 nt128 m2 a2 s4 128->128", 16, 92, 124, 128, 128, 0, 128, 1, false, true, 128, 4, 2, 2, 2},
      {"auto 1x1 512->512 92x164 b8 f32", 8, 92, 164, 512, 512, 0, 52, 1, true, true, 0, 0, 0, 0, 0},
      {"v2 512->52 m2 a2 s4 f32", 8, 92, 164, 512, 512, 0, 52, 1, true, true, 0, 4, 2, 2, 2},
  };
  int fails = 0;
  for (const Cfg& c : cfgs) {
    const int r = run_cfg(c, true);
    if (r == 3) return 3;
    fails += r;
  }
  printf("%s: %d failing configuration(s)\n", fails ? "FAILED" : "DONE", fails);
  return fails ? 1 : 0;
}

static int run_limits_suite() {
  // where is the ceiling? same layers with and without TMA traffic (no_loads: results are garbage by design)
  const Cfg cfgs[] = {
      {"7x7 128 nt128 s3 loads", 8, 92, 164, 128, 128, 0, 128, 7, false, true, 0, 3, 1, 0, 0, 0, 0},
      {"7x7 128 nt128 s3 NO loads", 8, 92, 164, 128, 128, 0, 128, 7, false, true, 0, 3, 1, 0, 0, 0, 1},
      {"7x7 128 nt128 s6 NO loads (1/SM)", 8, 92, 164, 128, 128, 0, 128, 7, false, true, 0, 6, 1, 0, 0, 0, 1},
      {"7x7 128 nt64 s3 NO loads", 8, 92, 164, 128, 128, 0, 128, 7, false, true, 64, 3, 1, 0, 0, 0, 1},
      {"3x3 512 nt256 s2 loads", 2, 92, 164, 512, 512, 0, 512, 3, false, true, 256, 2, 1, 0, 0, 0, 0},
      {"3x3 512 nt256 s2 NO loads", 2, 92, 164, 512, 512, 0, 512, 3, false, true, 256, 2, 1, 0, 0, 0, 1},
      {"3x3 512 nt128 s3 NO loads", 2, 92, 164, 512, 512, 0, 512, 3, false, true, 128, 3, 1, 0, 0, 0, 1},
      {"3x3 512 nt256 s4 NO loads (1/SM)", 2, 92, 164, 512, 512, 0, 512, 3, false, true, 256, 4, 1, 0, 0, 0, 1},
  };
  for (const Cfg& c : cfgs) {
    if (run_cfg(c, true) == 3) return 3;
  }
  return 0;
}

int main(int argc, char** argv) {
  const Cfg cfgs[] = {
      // name                       N   H    W  in_c cstr coff cout k  f32   bf16  ntile stages
      {"1x1 64->64 16x8", 1, 8, 16, 64, 64, 0, 64, 1, true, true, 0, 0},
      {"1x1 64->64 k-slices", 1, 8, 16, 48, 64, 0, 64, 1, true, true, 0, 0},
      {"3x3 64->64 16x8", 1, 8, 16, 64, 64, 0, 64, 3, true, true, 0, 0},
      {"3x3 128->128 23x41 b2", 2, 23, 41, 128, 128, 0, 128, 3, true, true, 0, 0},
      {"3x3 slice 96/288 ->96", 2, 23, 41, 96, 288, 96, 96, 3, true, true, 0, 0},
      {"7x7 192->128 23x41 b2", 2, 23, 41, 192, 192, 0, 128, 7, true, true, 0, 0},
      {"7x7 160->128 23x23 b3", 3, 23, 23, 160, 160, 0, 128, 7, true, true, 0, 0},
      {"1x1 128->512 46x62", 1, 46, 62, 128, 128, 0, 512, 1, true, true, 0, 0},
      {"1x1 128->512 nt128", 1, 46, 62, 128, 128, 0, 512, 1, true, true, 128, 0},
      {"1x1 512->38 head f32", 2, 23, 41, 512, 512, 0, 38, 1, true, true, 0, 0},
      {"1x1 128->19 head f32", 2, 23, 41, 128, 128, 0, 19, 1, true, false, 0, 0},
      {"1x1 32->64 im2col 184x328", 1, 184, 328, 32, 32, 0, 64, 1, false, true, 0, 0},
      {"3x3 256->256 46x82 b4", 4, 46, 82, 256, 256, 0, 256, 3, false, true, 0, 0},
      {"3x3 512->512 92x164", 1, 92, 164, 512, 512, 0, 512, 3, false, true, 0, 0},
      {"7x7 128->128 92x164 b2", 2, 92, 164, 128, 128, 0, 128, 7, false, true, 0, 0},
      {"7x7 128->128 92x164 s3", 2, 92, 164, 128, 128, 0, 128, 7, false, true, 0, 3},
      {"3x3 64->64 368x656", 1, 368, 656, 64, 64, 0, 64, 3, false, true, 0, 0},
      // tiling comparisons on the heavy layers (auto = n_tile 128, 3 stages, two CTAs per SM)
      {"3x3 256->256 46x82 b4 nt256s4", 4, 46, 82, 256, 256, 0, 256, 3, false, true, 256, 4},
      {"3x3 256->256 46x82 b4 nt256s2", 4, 46, 82, 256, 256, 0, 256, 3, false, true, 256, 2},
      {"3x3 512->512 92x164 nt256s4", 1, 92, 164, 512, 512, 0, 512, 3, false, true, 256, 4},
      {"3x3 512->512 92x164 nt128s6", 1, 92, 164, 512, 512, 0, 512, 3, false, true, 128, 6},
      {"7x7 128->128 92x164 b2 s6", 2, 92, 164, 128, 128, 0, 128, 7, false, true, 0, 6},
      {"7x7 128->128 92x164 b2 s2", 2, 92, 164, 128, 128, 0, 128, 7, false, true, 0, 2},
      {"7x7 128->128 23x41 b8", 8, 23, 41, 128, 128, 0, 128, 7, false, true, 0, 0},
      {"7x7 128->128 23x41 b8 nt64", 8, 23, 41, 128, 128, 0, 128, 7, false, true, 64, 0},
      {"7x7 192->128 60x80 b8", 8, 60, 80, 192, 192, 0, 128, 7, false, true, 0, 0},
      {"3x3 128->128 184x248 b8", 8, 184, 248, 128, 128, 0, 128, 3, false, true, 0, 0},
      {"3x3 64->64 368x496 b8", 8, 368, 496, 64, 64, 0, 64, 3, false, true, 0, 0},
      {"1x1 32->64 368x496 b8", 8, 368, 496, 32, 32, 0, 64, 1, false, true, 0, 0},
      {"3x3 96->96 slice 92x164 b2", 2, 92, 164, 96, 288, 96, 96, 3, false, true, 0, 0},
      {"3x3 288->96 92x164 b2", 2, 92, 164, 288, 288, 0, 96, 3, false, true, 0, 0},
  };
  const int n = sizeof(cfgs) / sizeof(cfgs[0]);
  if (argc > 1 && std::string(argv[1]) == "v2") return run_v2_suite();
  if (argc > 1 && std::string(argv[1]) == "v3") return run_v3_suite();
  if (argc > 1 && std::string(argv[1]) == "limits") return run_limits_suite();
  if (argc > 1 && std::string(argv[1]) == "v4") return run_v4_suite();
  if (argc > 2) g_only = atoi(argv[2]);
  if (argc > 1 && std::string(argv[1]) == "v5") return run_v5_suite();
  if (argc > 1 && std::string(argv[1]) == "gemm") return run_gemm_suite();
  int fails = 0;
  for (int i = 0; i < n; ++i) {
    if (argc > 1 && atoi(argv[1]) != i) continue;
    const int r = run_cfg(cfgs[i], true);
    if (r == 3) {
      printf("context lost, stopping\n");
      return 3;
    }
    fails += r;
  }
  printf("%s: %d failing configuration(s)\n", fails ? "FAILED" : "ALL PASS", fails);
  return fails ? 1 : 0;
}
