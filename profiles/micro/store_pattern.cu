// Microbenchmark: what the fused kernel's STORE PATTERN can reach by itself.  Every CTA (256 threads, 100 KB of
// dynamic shared memory so that two fit an SM, like ms_fused_kernel) writes one tile of the [N][8][D][h][w] volume --
// 8 x D row segments of 128 bytes, each in a different 2 MB plane -- and nothing else.
//   mode 0: thread = (pixel quad, d), 128-bit stores, 4 row segments per warp instruction (the kernel's sweeps)
//   mode 1: lane = pixel, 32-bit stores, one row segment per warp instruction
//   mode 2: as 0, but each tile writes ONE contiguous 8 x D x 128 B block
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_pattern store_pattern.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256, 2) k(float* out, int N, int D, int h, int w, int n_tiles) {
  extern __shared__ float sm[];
  const int tiles_x = w / 32;
  for (int tile0 = blockIdx.x; tile0 < n_tiles; tile0 += gridDim.x) {   // (one tile per CTA when the grid covers them all)
  int tile = tile0;
  const int xt = tile % tiles_x; tile /= tiles_x;
  const int y = tile % h, n = tile / h;
  const size_t plane = (size_t)h * w, chan = plane * D;
  const int tid = threadIdx.x;
  if (tid == 0) sm[0] = 1.0f;
  __syncthreads();
  const float v = sm[0] + tid;
  if (MODE == 0 || MODE == 2) {
    const int q4 = (tid & 7) * 4;
    float* base = (MODE == 2) ? out + (size_t)tile0 * 8 * D * 32 + q4
                              : out + (size_t)n * 8 * chan + (size_t)y * w + xt * 32 + q4;
    const size_t pl = (MODE == 2) ? 32 : plane, ch = (MODE == 2) ? (size_t)D * 32 : chan;
    for (int pass = 0; pass < 2; ++pass)
      for (int d = tid >> 3; d < D; d += 32)
#pragma unroll
        for (int c = 0; c < 4; ++c)
          __stcs(reinterpret_cast<float4*>(base + (pass * 4 + c) * ch + (size_t)d * pl), make_float4(v, v + 1, v + 2, v + d));
  } else if (MODE == 4) {
    const int o8 = (tid & 3) * 8;
    float* base = out + (size_t)n * 8 * chan + (size_t)y * w + xt * 32 + o8;
    for (int c = 0; c < 8; ++c)
      for (int d = tid >> 2; d < D; d += 64) {
        float* p = base + c * chan + (size_t)d * plane;
        asm volatile("st.global.cs.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v), "f"(v + 1), "f"(v + 2),
                     "f"(v + 3), "f"(v + 4), "f"(v + 5), "f"(v + 6), "f"(v + d) : "memory");
      }
  } else {
    const int lane = tid & 31, warp = tid >> 5;
    float* base = out + (size_t)n * 8 * chan + (size_t)y * w + xt * 32 + lane;
    for (int c = 0; c < 8; ++c)
      for (int d = warp; d < D; d += 8) __stcs(base + c * chan + (size_t)d * plane, v + d);
  }
  }
}
// ctas > 0: that many persistent CTAs walk the tiles (few CTAs = the store rate ONE SM reaches with DRAM far from busy)
template <int MODE> void run(const char* name, float* out, int N, int D, int h, int w, int ctas = 0) {
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int n_tiles = ctas > 0 ? ctas * 64 : N * h * (w / 32);
  const int grid = ctas > 0 ? ctas : n_tiles;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 2; ++i) k<MODE><<<grid, 256, 100 * 1024>>>(out, N, D, h, w, n_tiles);
  cudaEventRecord(e0);
  for (int i = 0; i < 5; ++i) k<MODE><<<grid, 256, 100 * 1024>>>(out, N, D, h, w, n_tiles);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  const double bytes = (double)n_tiles * 8 * D * 32 * 4;
  if (ctas > 0) printf("%-44s %d CTAs (one per SM): %.1f GB/s per CTA = %.1f B/clk at 1.93 GHz\n", name, ctas, bytes / ms * 1e-6 / ctas, bytes / ms * 1e-6 / ctas / 1.93);
  else printf("%-44s %.3f ms per %d pairs = %.3f ms/pair  %.0f GB/s\n", name, ms, N, ms / N, bytes / ms * 1e-6);
}
int main() {
  const int N = 4, D = 192, h = 540, w = 960;
  float* out; cudaMalloc(&out, (size_t)N * 8 * D * h * w * 4);
  run<0>("128-bit stores, 4 row segments per instruction", out, N, D, h, w);
  run<1>("32-bit stores, 1 row segment per instruction", out, N, D, h, w);
  run<2>("128-bit stores, contiguous block per tile", out, N, D, h, w);
  run<0>("128-bit stores, 4 row segments per instruction", out, N, D, h, w, 16);
  run<0>("128-bit stores, 4 row segments per instruction", out, N, D, h, w, 148);
  run<0>("128-bit stores, 4 row segments per instruction", out, N, D, h, w, 296);
  run<2>("128-bit stores, contiguous block per tile", out, N, D, h, w, 16);
  run<4>("256-bit stores, 8 row segments per instruction", out, N, D, h, w);
  run<4>("256-bit stores, 8 row segments per instruction", out, N, D, h, w, 16);
  run<1>("32-bit stores, 1 row segment per instruction", out, N, D, h, w, 16);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
