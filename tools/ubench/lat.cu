// Latency micro-benchmarks used to reason about the learner kernel's serial chains (B200, sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o lat lat.cu && ./lat
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dadd(double* out, double x, int n, long long* cyc) {
  double a = out[0];
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int j = 0; j < 32; ++j) a += x;
  }
  long long t1 = clock64();
  out[1] = a; cyc[0] = t1 - t0;
}
__global__ void k_dmuladd(double* out, double w, const double* v, int n, long long* cyc) {
  __shared__ double s[32];
  if (threadIdx.x < 32) s[threadIdx.x] = v[threadIdx.x];
  __syncthreads();
  double a = out[0];
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
#pragma unroll 8
    for (int j = 0; j < 32; ++j) a += w * s[j];
  }
  long long t1 = clock64();
  out[2] = a; cyc[1] = t1 - t0;
}
__global__ void k_fadd(float* out, float x, int n, long long* cyc) {
  float a = out[0];
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int j = 0; j < 32; ++j) a += x;
  }
  long long t1 = clock64();
  out[1] = a; cyc[2] = t1 - t0;
}
__global__ void k_chase(const int* p, int n, int* out, long long* cyc, int slot) {
  int i = 0;
  long long t0 = clock64();
#pragma unroll 1
  for (int k = 0; k < n; ++k) i = __ldcg(p + i);
  long long t1 = clock64();
  out[0] = i; cyc[slot] = t1 - t0;
}
__global__ void k_lds(int* out, int n, long long* cyc) {
  __shared__ int s[64];
  if (threadIdx.x < 64) s[threadIdx.x] = (threadIdx.x + 1) & 63;
  __syncthreads();
  int i = 0;
  long long t0 = clock64();
#pragma unroll 1
  for (int k = 0; k < n; ++k) i = s[i];
  long long t1 = clock64();
  out[1] = i; cyc[5] = t1 - t0;
}
__global__ void k_imad(unsigned long long* out, unsigned long long m, int n, long long* cyc) {
  unsigned long long a = out[0];
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int j = 0; j < 32; ++j) a = a * m + 12345ull;
  }
  long long t1 = clock64();
  out[1] = a; cyc[6] = t1 - t0;
}
int main() {
  double* d; float* f; long long* c; int *p, *o; unsigned long long* u;
  cudaMalloc(&d, 64 * 8); cudaMalloc(&f, 256); cudaMalloc(&c, 256); cudaMalloc(&o, 256); cudaMalloc(&u, 256);
  cudaMemset(c, 0, 256);
  cudaMemset(d, 0, 64 * 8); cudaMemset(f, 0, 64); cudaMemset(u, 0, 64);
  double hv[32]; for (int i = 0; i < 32; ++i) hv[i] = 1.0 / (i + 3);
  double* v; cudaMalloc(&v, 256); cudaMemcpy(v, hv, 256, cudaMemcpyHostToDevice);
  // pointer chase: small ring (L2 hit) and a 1 GiB ring with a large stride (DRAM)
  const size_t NB = 256u << 20;  // ints = 1 GiB
  cudaMalloc(&p, NB * 4);
  {
    int* h = (int*)malloc(NB * 4);
    const size_t stride = 1u << 18;  // 1 MiB
    for (size_t i = 0; i < NB; ++i) h[i] = 0;
    size_t cur = 0;
    for (int k = 0; k < 1000; ++k) { size_t nx = (cur + stride * 37 + 32 * (k % 7)) % NB; h[cur] = (int)nx; cur = nx; }
    h[cur] = 0;
    cudaMemcpy(p, h, NB * 4, cudaMemcpyHostToDevice);
    free(h);
  }
  const int n = 64;
  for (int rep = 0; rep < 2; ++rep) {
    k_dadd<<<1, 32>>>(d, 1e-9, n, c);
    k_dmuladd<<<1, 32>>>(d, 0.65, v, n, c);
    k_fadd<<<1, 32>>>(f, 1e-9f, n, c);
    k_chase<<<1, 1>>>(p, 1000, o, c, 3);   // first pass: DRAM (cold), second pass: L2 hits
    k_chase<<<1, 1>>>(p, 1000, o, c, 4);
    k_lds<<<1, 64>>>(o, 2048, c);
    k_imad<<<1, 32>>>(u, 6364136223846793005ull, n, c);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  }
  long long h[8];
  cudaMemcpy(h, c, 64, cudaMemcpyDeviceToHost);
  printf("dependent DADD          %.1f cycles\n", h[0] / (double)(n * 32));
  printf("dependent a += w*s[j]   %.1f cycles (smem load + DMUL off the chain, unroll 8)\n", h[1] / (double)(n * 32));
  printf("dependent FADD          %.1f cycles\n", h[2] / (double)(n * 32));
  printf("ld.cg chase pass 1      %.1f cycles/load\n", h[3] / 1000.0);
  printf("ld.cg chase pass 2 (L2) %.1f cycles/load\n", h[4] / 1000.0);
  printf("dependent LDS           %.1f cycles\n", h[5] / 2048.0);
  printf("dependent IMAD.64       %.1f cycles\n", h[6] / (double)(n * 32));
  return 0;
}
