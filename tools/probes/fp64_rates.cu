// Probe: raw FP64 issue rates on the target GPU (DMMA.8x8x4 vs DFMA), used to pick the
// arithmetic path for the batched QP operator apply.  Not part of the product path.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dmma_rate(double* out, int iters){
  double a0=1.0+threadIdx.x*1e-9, b0=1.0-threadIdx.x*1e-9;
  double c[16]; for(int i=0;i<16;i++) c[i]=0;
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int i=0;i<8;i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};":"+d"(c[2*i]),"+d"(c[2*i+1]):"d"(a0),"d"(b0));
  }
  double s=0; for(int i=0;i<16;i++) s+=c[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
__global__ void dfma_rate(double* out, int iters){
  double a0=1.0+threadIdx.x*1e-9, b0=1e-9*threadIdx.x;
  double c[16]; for(int i=0;i<16;i++) c[i]=i;
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int i=0;i<16;i++) c[i]=fma(c[i],a0,b0);
  }
  double s=0; for(int i=0;i<16;i++) s+=c[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
int main(){
  cudaDeviceProp p; cudaGetDeviceProperties(&p,0);
  printf("device %s sms %d clock %d kHz\n",p.name,p.multiProcessorCount,p.clockRate);
  double* out; cudaMalloc(&out, 148*8*1024*8);
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for(int warps=4; warps<=32; warps*=2){
    int iters=20000; float ms;
    dmma_rate<<<148*2,warps*32>>>(out,100);
    cudaEventRecord(e0); dmma_rate<<<148*2,warps*32>>>(out,iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms,e0,e1);
    double fl=2.0*256*8*iters*(double)warps*148*2;
    printf("DMMA warps/CTA %d (2 CTA/SM): %.2f TFLOP/s\n",warps,fl/ms*1e-9);
    dfma_rate<<<148*2,warps*32>>>(out,100);
    cudaEventRecord(e0); dfma_rate<<<148*2,warps*32>>>(out,iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms,e0,e1);
    fl=2.0*16*32*iters*(double)warps*148*2;
    printf("DFMA warps/CTA %d (2 CTA/SM): %.2f TFLOP/s\n",warps,fl/ms*1e-9);
  }
  printf("err %s\n",cudaGetErrorString(cudaGetLastError()));
  return 0;
}
