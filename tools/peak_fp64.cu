// Microbenchmark: FP64 peak on this GPU via DMMA.8x8x4 (mma.sync m8n8k4 / m16n8k8 f64) and DFMA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o peak_fp64 peak_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

template<int CH>
__global__ void __launch_bounds__(256) k_dmma884(double* out, int iters, double s){
  double c[CH][2]; double a = s*threadIdx.x, b = s*(threadIdx.x+1);
  #pragma unroll
  for(int i=0;i<CH;i++){c[i][0]=0;c[i][1]=0;}
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<CH;i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};"
        :"+d"(c[i][0]),"+d"(c[i][1]):"d"(a),"d"(b));
  }
  double r=0;
  #pragma unroll
  for(int i=0;i<CH;i++) r+=c[i][0]+c[i][1];
  if(r==123.456) out[0]=r;
}
template<int CH>
__global__ void __launch_bounds__(256) k_dmma1688(double* out, int iters, double s){
  double c[CH][4]; double a[4], b[2];
  for(int i=0;i<4;i++)a[i]=s*(threadIdx.x+i); b[0]=s; b[1]=2*s;
  #pragma unroll
  for(int i=0;i<CH;i++){c[i][0]=0;c[i][1]=0;c[i][2]=0;c[i][3]=0;}
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<CH;i++)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3},{%4,%5,%6,%7},{%8,%9},{%0,%1,%2,%3};"
       :"+d"(c[i][0]),"+d"(c[i][1]),"+d"(c[i][2]),"+d"(c[i][3]):"d"(a[0]),"d"(a[1]),"d"(a[2]),"d"(a[3]),"d"(b[0]),"d"(b[1]));
  }
  double r=0;
  #pragma unroll
  for(int i=0;i<CH;i++) r+=c[i][0]+c[i][1]+c[i][2]+c[i][3];
  if(r==123.456) out[0]=r;
}
template<int CH>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double s){
  double c[CH]; double a = s*threadIdx.x, b = s;
  #pragma unroll
  for(int i=0;i<CH;i++) c[i]=i;
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<CH;i++) c[i]=fma(a,c[i],b);
  }
  double r=0;
  #pragma unroll
  for(int i=0;i<CH;i++) r+=c[i];
  if(r==123.456) out[0]=r;
}
template<typename F> float timeit(F f){
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  float best=1e30f;
  for(int r=0;r<5;r++){ cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms;}
  return best;
}
int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
  int sms=p.multiProcessorCount; printf("{\"gpu\":\"%s\",\"sms\":%d,\n",p.name,sms);
  double* out; CK(cudaMalloc(&out,8));
  int iters=20000;
  for(int bps=1;bps<=4;bps*=2){
    int grid=sms*bps;
    { float ms=timeit([&]{k_dmma884<8><<<grid,256>>>(out,iters,1e-30);});
      double fl=2.0*256*8*(double)iters*8*grid; printf(" \"dmma884_ch8_bps%d_tflops\":%.2f,\n",bps,fl/ms/1e9); }
    { float ms=timeit([&]{k_dmma884<16><<<grid,256>>>(out,iters,1e-30);});
      double fl=2.0*256*16*(double)iters*8*grid; printf(" \"dmma884_ch16_bps%d_tflops\":%.2f,\n",bps,fl/ms/1e9); }
    { float ms=timeit([&]{k_dmma1688<8><<<grid,256>>>(out,iters,1e-30);});
      double fl=2.0*1024*8*(double)iters*8*grid; printf(" \"dmma1688_ch8_bps%d_tflops\":%.2f,\n",bps,fl/ms/1e9); }
    { float ms=timeit([&]{k_dfma<16><<<grid,256>>>(out,iters,1e-30);});
      double fl=2.0*16*(double)iters*256*grid; printf(" \"dfma_ch16_bps%d_tflops\":%.2f,\n",bps,fl/ms/1e9); }
  }
  // single warp per SMSP latency probe: 1 block of 128 threads per SM, 1 chain
  { float ms=timeit([&]{k_dmma884<1><<<sms,128>>>(out,iters,1e-30);});
    printf(" \"dmma884_dep_chain_ns_per_instr\":%.2f,\n", ms*1e6/iters); }
  { float ms=timeit([&]{k_dfma<1><<<sms,128>>>(out,iters,1e-30);});
    printf(" \"dfma_dep_chain_ns_per_instr\":%.2f,\n", ms*1e6/iters); }
  printf(" \"clock_khz\":%d}\n",p.clockRate);
  return 0;
}
