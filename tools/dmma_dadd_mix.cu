// Microbenchmark behind the 3M kernel design: how much FP64-pipe time do the operand sums (DADD) of the 3-multiplication
// complex product cost when they share the pipe with DMMA.8x8x4, depending on how many there are and where they sit?
// One CTA of 8 warps per SM, 48 independent DMMA accumulators per warp and iteration (= one k-step of the 32 x 32 warp tile).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_dadd_mix dmma_dadd_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void xadd(double& x, double y) { asm volatile("add.f64 %0, %0, %1;" : "+d"(x) : "d"(y)); }

// MODE 0: no sums.  1: NADD sums spread between the DMMAs (one after every 32/NADD DMMAs of the first 32).
// 2: NADD sums in one block after the first 16 DMMAs, pinned there by a one-trip loop with a run-time bound.
// 3: like 2 without the pinning loop (ptxas is free to move them).
template <int MODE, int NADD>
__global__ void __launch_bounds__(256) k_mix(double* out, const double* __restrict__ in, int iters, int one) {
  double c[48][2]; double a[4][2], b[4][2];
#pragma unroll
  for (int i = 0; i < 48; ++i) { c[i][0] = 0; c[i][1] = 0; }
#pragma unroll
  for (int i = 0; i < 4; ++i) { a[i][0] = in[threadIdx.x * 16 + i]; a[i][1] = in[threadIdx.x * 16 + 4 + i]; b[i][0] = in[threadIdx.x * 16 + 8 + i]; b[i][1] = in[threadIdx.x * 16 + 12 + i]; }
  for (int it = 0; it < iters; ++it) {
    // P1: 16 DMMAs on the "real" parts
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        dmma(c[i * 4 + j][0], c[i * 4 + j][1], a[i][0], b[j][0]);
        if (MODE == 1 && NADD > 0 && ((i * 4 + j) % (16 / (NADD / 2 > 0 ? NADD / 2 : 1))) == 0 && (i * 4 + j) / (16 / (NADD / 2 > 0 ? NADD / 2 : 1)) < NADD / 2) {
          const int q = (i * 4 + j) / (16 / (NADD / 2)); xadd(a[q % 4][0], a[q % 4][1]);
        }
      }
    if (MODE == 2) {
#pragma unroll 1
      for (int r = 0; r < one; ++r) {
#pragma unroll
        for (int q = 0; q < NADD; ++q) { if (q < 4) xadd(a[q][0], a[q][1]); else xadd(b[q - 4][0], b[q - 4][1]); }
      }
    }
    if (MODE == 3) {
#pragma unroll
      for (int q = 0; q < NADD; ++q) { if (q < 4) xadd(a[q][0], a[q][1]); else xadd(b[q - 4][0], b[q - 4][1]); }
    }
    // P2: 16 DMMAs on the "imaginary" parts
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        dmma(c[16 + i * 4 + j][0], c[16 + i * 4 + j][1], a[i][1], b[j][1]);
        if (MODE == 1 && NADD > 4 && ((i * 4 + j) % (16 / (NADD - NADD / 2))) == 0 && (i * 4 + j) / (16 / (NADD - NADD / 2)) < NADD - NADD / 2) {
          const int q = (i * 4 + j) / (16 / (NADD - NADD / 2)); xadd(b[q % 4][0], b[q % 4][1]);
        }
      }
    // P3: 16 DMMAs on the sums
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dmma(c[32 + i * 4 + j][0], c[32 + i * 4 + j][1], a[i][0], b[j][0]);
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < 48; ++i) r += c[i][0] + c[i][1];
  if (r == 123.456) out[0] = r;
}

template <typename F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; r++) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
  return best;
}
#define RUN(MODE, NADD) { float ms = timeit([&] { k_mix<MODE, NADD><<<sms, 256>>>(out, in, iters, 1); }); \
  double fl = 2.0 * 256 * 48 * (double)iters * 8 * sms; printf(" \"mode%d_nadd%d_dmma_tflops\":%.2f,\n", MODE, NADD, fl / ms / 1e9); }
int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount; printf("{\"gpu\":\"%s\",\"sms\":%d,\n", p.name, sms);
  double* out; CK(cudaMalloc(&out, 8));
  double* in; CK(cudaMalloc(&in, 256 * 16 * 8)); CK(cudaMemset(in, 0, 256 * 16 * 8));
  int iters = 4000;
  RUN(0, 0) RUN(1, 8) RUN(1, 4) RUN(1, 2) RUN(2, 8) RUN(2, 4) RUN(3, 8)
  printf(" \"note\":\"DMMA TFLOP/s (executed) with NADD DADDs per 48 DMMAs\"}\n");
  return 0;
}
