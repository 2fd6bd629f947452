// gather.cu -- what a B200 sustains on the learner kernel's access pattern: bursts of independent 8-byte loads at
// random offsets inside a per-warp window (one env's weight table) of a multi-GB array.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather gather.cu && ./gather
// Prints G loads/s (= 32-byte sectors/s) for several (footprint, window, loads-in-flight, width) points, a
// prefetch.global.L2 variant, and a coalesced 32 KB bulk read per warp for comparison.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

template <int K, int MODE>  // MODE 0: ld.cg 8 B, 1: prefetch L2, 2: ld.cg 16 B
__global__ void k_gather(const double* __restrict__ base, size_t n_windows, size_t window_doubles, int rounds, double* out, uint32_t salt) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  double acc = 0.0;
  for (int r = 0; r < rounds; ++r) {
    const size_t w = (size_t)(mix(warp * 977u + r * 131071u + salt) % (uint32_t)n_windows);
    const double* tab = base + w * window_doubles;
    uint32_t h = mix((warp * 32u + lane) * 2654435761u + r + salt);
    double v[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      h = h * 1664525u + 1013904223u;
      size_t off = (size_t)(mix(h) % (uint32_t)window_doubles);
      if (MODE == 1) { asm volatile("prefetch.global.L2 [%0];" ::"l"(tab + off)); v[k] = 0.0; }
      else if (MODE == 2) { off &= ~(size_t)1; double2 t = __ldcg((const double2*)(tab + off)); v[k] = t.x + t.y; }
      else v[k] = __ldcg(tab + off);
    }
#pragma unroll
    for (int k = 0; k < K; ++k) acc += v[k];
  }
  if (acc == 123.456) out[0] = acc;
}

// coalesced bulk read of one window per warp (what TMA staging of a small table costs)
__global__ void k_bulk(const double* __restrict__ base, size_t n_windows, size_t window_doubles, int rounds, double* out, uint32_t salt) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  double acc = 0.0;
  for (int r = 0; r < rounds; ++r) {
    const size_t w = (size_t)(mix(warp * 977u + r * 131071u + salt) % (uint32_t)n_windows);
    const double2* tab = (const double2*)(base + w * window_doubles);
    for (size_t i = lane; i < window_doubles / 2; i += 32 * 8) {
      double2 t[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) t[k] = (i + 32 * k < window_doubles / 2) ? __ldcg(tab + i + 32 * k) : make_double2(0, 0);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc += t[k].x + t[k].y;
    }
  }
  if (acc == 123.456) out[0] = acc;
}

template <int K, int MODE>
static int run(const char* name, const double* d, size_t foot_bytes, size_t window_bytes, int warps, int rounds, double* d_out) {
  const size_t wd = window_bytes / 8, nw = foot_bytes / window_bytes;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int threads = 128, blocks = warps * 32 / threads;
  float best = 1e30f;
  for (int it = 0; it < 5; ++it) {
    cudaEventRecord(a);
    k_gather<K, MODE><<<blocks, threads>>>(d, nw, wd, rounds, d_out, 1234u + it * 77u);
    cudaEventRecord(b);
    CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (it > 0 && ms < best) best = ms;
  }
  const double loads = (double)warps * 32 * K * rounds;
  printf("%-34s footprint %6.0f MB window %7.0f KB warps %6d x %2d in flight x %3d rounds: %8.1f us  %7.2f G loads/s  (%6.1f GB/s of %d-byte sectors)\n", name,
         foot_bytes / 1e6, window_bytes / 1e3, warps, K, rounds, best * 1e3, loads / (best * 1e-3) / 1e9, loads * 32 / (best * 1e-3) / 1e9, 32);
  return 0;
}

int main() {
  const size_t total = (size_t)8 << 30;
  double* d; double* d_out;
  CK(cudaMalloc(&d, total)); CK(cudaMalloc(&d_out, 64));
  CK(cudaMemset(d, 0, total));
  cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32);
  // the learner's burst at C1: 1205 warps x 27 loads per lane, 512 KB windows inside 2 GB
  run<27, 0>("C1 burst (one step per warp)", d, (size_t)2 << 30, 512 << 10, 1204, 1, d_out);
  run<27, 0>("same, 8 steps per warp", d, (size_t)2 << 30, 512 << 10, 1204, 8, d_out);
  run<27, 0>("4736 warps x 8 steps", d, (size_t)2 << 30, 512 << 10, 4736, 8, d_out);
  run<27, 0>("9472 warps x 8 steps", d, (size_t)2 << 30, 512 << 10, 9472, 8, d_out);
  run<9, 0>("9472 warps, 9 in flight x 24", d, (size_t)2 << 30, 512 << 10, 9472, 24, d_out);
  run<27, 0>("L2-resident footprint (64 MB)", d, (size_t)64 << 20, 512 << 10, 9472, 8, d_out);
  run<27, 0>("TLB-reach footprint (256 MB)", d, (size_t)256 << 20, 512 << 10, 9472, 8, d_out);
  run<27, 0>("8 GB footprint", d, (size_t)8 << 30, 512 << 10, 9472, 8, d_out);
  run<27, 0>("8 GB, windows of 128 KB", d, (size_t)8 << 30, 128 << 10, 9472, 8, d_out);
  run<27, 0>("4 GB, windows of 32 KB", d, (size_t)4 << 30, 32 << 10, 9472, 8, d_out);
  run<27, 0>("2 GB, no window (whole array)", d, (size_t)2 << 30, (size_t)2 << 30, 9472, 8, d_out);
  run<27, 1>("prefetch.global.L2, 2 GB", d, (size_t)2 << 30, 512 << 10, 9472, 8, d_out);
  run<27, 2>("16-byte loads, 2 GB", d, (size_t)2 << 30, 512 << 10, 9472, 8, d_out);
  cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 128);
  run<27, 0>("2 GB, L2 fetch granularity 128", d, (size_t)2 << 30, 512 << 10, 9472, 8, d_out);
  cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32);
  {
    const size_t win = 32 << 10, foot = (size_t)4 << 30;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int warps : {1184, 4736, 9472}) {
      float best = 1e30f;
      for (int it = 0; it < 4; ++it) {
        cudaEventRecord(a);
        k_bulk<<<warps * 32 / 128, 128>>>(d, foot / win, win / 8, 8, d_out, 99u + it);
        cudaEventRecord(b);
        CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (it > 0 && ms < best) best = ms;
      }
      printf("bulk 32 KB window per warp-step      warps %6d x 8 rounds: %8.1f us  %7.1f GB/s  %7.2f M windows/s\n", warps, best * 1e3,
             (double)warps * 8 * win / (best * 1e-3) / 1e9, (double)warps * 8 / (best * 1e-3) / 1e6);
    }
  }
  return 0;
}
