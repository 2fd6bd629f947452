// gather_tma.cu -- can the TMA engine's own request path take (part of) the learner's gather burst?
// Each lane fetches K random 8-byte weights of a 512 KB window, KT of them as 16-byte cp.async.bulk copies into shared
// memory (completion on an mbarrier), the other K - KT as ld.global.cg.  Same launch shape as rlm_learn_kernel: 96
// threads per CTA.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_tma gather_tma.cu && ./gather_tma
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int K, int KT>
__global__ void __launch_bounds__(96) k_mix(const double* __restrict__ base, size_t n_windows, size_t window_doubles, int rounds, double* out, uint32_t salt) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp = blockIdx.x * 3 + wib;
  unsigned long long* mbar = (unsigned long long*)smem + wib;
  double2* slots = (double2*)(smem + 64 + (size_t)wib * (KT > 0 ? KT : 1) * 32 * 16);
  const unsigned mbar_a = smem_u32(mbar);
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_a) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  unsigned parity = 0;
  double acc = 0.0;
  for (int r = 0; r < rounds; ++r) {
    const size_t w = (size_t)(mix(warp * 977u + r * 131071u + salt) % (uint32_t)n_windows);
    const double* tab = base + w * window_doubles;
    uint32_t h = mix((warp * 32u + lane) * 2654435761u + r + salt);
    if (KT > 0) {
      if (lane == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_a), "r"(KT * 32 * 16) : "memory");
      }
      __syncwarp();
    }
    double v[K - KT > 0 ? K - KT : 1];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      h = h * 1664525u + 1013904223u;
      size_t off = (size_t)(mix(h) % (uint32_t)window_doubles);
      if (k < KT) {
        off &= ~(size_t)1;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];" ::"r"(smem_u32(slots + k * 32 + lane)),
                     "l"(tab + off), "r"(mbar_a)
                     : "memory");
      } else v[k - KT] = __ldcg(tab + off);
    }
#pragma unroll
    for (int k = KT; k < K; ++k) acc += v[k - KT];
    if (KT > 0) {
      unsigned done = 0;
      while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(mbar_a), "r"(parity) : "memory");
      parity ^= 1;
#pragma unroll
      for (int k = 0; k < KT; ++k) { const double2 t = slots[k * 32 + lane]; acc += t.x + t.y; }
      __syncwarp();
    }
  }
  if (acc == 123.456) out[0] = acc;
}

template <int K, int KT>
static int run(const char* name, const double* d, size_t foot_bytes, size_t window_bytes, int warps, int rounds, double* d_out) {
  const size_t wd = window_bytes / 8, nw = foot_bytes / window_bytes;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int blocks = warps / 3;
  const size_t smem = 64 + (size_t)3 * (KT > 0 ? KT : 1) * 32 * 16;
  CK(cudaFuncSetAttribute(k_mix<K, KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  float best = 1e30f;
  for (int it = 0; it < 6; ++it) {
    cudaEventRecord(a);
    k_mix<K, KT><<<blocks, 96, smem>>>(d, nw, wd, rounds, d_out, 1234u + it * 77u);
    cudaEventRecord(b);
    CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (it > 0 && ms < best) best = ms;
  }
  const double loads = (double)blocks * 3 * 32 * K * rounds;
  printf("%-44s warps %5d x %2d per lane (%2d by TMA) x %2d rounds: %8.1f us  %7.2f G fetches/s\n", name, blocks * 3, K, KT, rounds, best * 1e3,
         loads / (best * 1e-3) / 1e9);
  return 0;
}

int main() {
  const size_t total = (size_t)2 << 30;
  double* d; double* d_out;
  CK(cudaMalloc(&d, total)); CK(cudaMalloc(&d_out, 64));
  CK(cudaMemset(d, 0, total));
  cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32);
  const size_t W = 512 << 10;
  run<27, 0>("C1 burst, ld.global.cg", d, total, W, 1206, 1, d_out);
  run<27, 27>("C1 burst, cp.async.bulk 16 B", d, total, W, 1206, 1, d_out);
  run<27, 9>("C1 burst, 1/3 TMA + 2/3 ld", d, total, W, 1206, 1, d_out);
  run<27, 14>("C1 burst, 1/2 TMA + 1/2 ld", d, total, W, 1206, 1, d_out);
  run<27, 0>("8 steps per warp, ld.global.cg", d, total, W, 1206, 8, d_out);
  run<27, 27>("8 steps per warp, cp.async.bulk 16 B", d, total, W, 1206, 8, d_out);
  run<27, 14>("8 steps per warp, 1/2 TMA + 1/2 ld", d, total, W, 1206, 8, d_out);
  run<27, 0>("3996 warps x 8, ld.global.cg", d, total, W, 3996, 8, d_out);
  run<27, 27>("3996 warps x 8, cp.async.bulk 16 B", d, total, W, 3996, 8, d_out);
  run<27, 14>("3996 warps x 8, 1/2 TMA + 1/2 ld", d, total, W, 3996, 8, d_out);
  return 0;
}
