// Probe: raw issue rate of the FP64 mma.sync shapes on this GPU.  Not part of the product path.
#include <cstdio>
#include <cuda_runtime.h>
template <int SHAPE> __global__ void rate(double* out, int iters) {
  double a[8], b[4], c[16][4];
  for (int i = 0; i < 8; i++) a[i] = 1.0 + threadIdx.x * 1e-9 + i;
  for (int i = 0; i < 4; i++) b[i] = 1.0 - threadIdx.x * 1e-9 - i;
  for (int i = 0; i < 16; i++) for (int j = 0; j < 4; j++) c[i][j] = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (SHAPE == 0)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a[0]), "d"(b[0]));
      else if (SHAPE == 1)
        asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3},{%4,%5},{%6},{%0,%1,%2,%3};" : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
      else if (SHAPE == 2)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3},{%4,%5,%6,%7},{%8,%9},{%0,%1,%2,%3};" : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
      else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3},{%4,%5,%6,%7,%8,%9,%10,%11},{%12,%13,%14,%15},{%0,%1,%2,%3};" : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
    }
  }
  double s = 0;
  for (int i = 0; i < 16; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int SHAPE> void run(const char* name, double macs, double* out) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int ctas = 1; ctas <= 2; ctas++)
    for (int warps = 4; warps <= 16; warps *= 2) {
      int iters = 20000; float ms;
      rate<SHAPE><<<148 * ctas, warps * 32>>>(out, 100);
      cudaEventRecord(e0); rate<SHAPE><<<148 * ctas, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
      double fl = 2.0 * macs * 8 * iters * (double)warps * 148 * ctas;
      printf("%s warps/CTA %2d x %d CTA/SM: %.2f TFLOP/s\n", name, warps, ctas, fl / ms * 1e-9);
    }
}
int main() {
  double* out; cudaMalloc(&out, 148 * 2 * 1024 * 8);
  run<0>("m8n8k4  ", 8 * 8 * 4, out);
  run<1>("m16n8k4 ", 16 * 8 * 4, out);
  run<2>("m16n8k8 ", 16 * 8 * 8, out);
  run<3>("m16n8k16", 16 * 8 * 16, out);
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
