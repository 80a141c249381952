// tools/ulat.cu -- dependent-issue latency of scalar vs packed fp32 instructions on B200 (one warp, one chain).
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
template <int MODE> __global__ void k(float *out, float a, float b, long long *cyc) {
  float x = a; u64 p; asm("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(a), "f"(b));
  u64 q; asm("mov.b64 %0, {%1, %2};" : "=l"(q) : "f"(b), "f"(a));
  long long t0 = clock64();
  for (int it = 0; it < 1024; ++it) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      if (MODE == 0) x = __fadd_rn(x, b);
      if (MODE == 1) x = __fmaf_rn(x, a, b);
      if (MODE == 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p) : "l"(q));
      if (MODE == 3) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p) : "l"(q));
      if (MODE == 4) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p) : "l"(q));
      if (MODE == 5) { asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p) : "l"(q)); x = __fadd_rn(x, b); }   // two independent chains
    }
  }
  long long t1 = clock64();
  float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p));
  out[threadIdx.x] = x + lo + hi;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE> void run(const char *name) {
  float *out; long long *cyc, h;
  cudaMalloc(&out, 4 * 32); cudaMalloc(&cyc, 8);
  k<MODE><<<1, 32>>>(out, 1.0001f, 0.5f, cyc);
  k<MODE><<<1, 32>>>(out, 1.0001f, 0.5f, cyc);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-34s %6.2f cycles per dependent step\n", name, (double)h / (1024.0 * 32));
}
int main() {
  run<0>("FADD chain"); run<1>("FFMA chain"); run<2>("FADD2 chain"); run<3>("FFMA2 chain"); run<4>("FMUL2 chain");
  run<5>("FADD2 chain + FADD chain");
  return 0;
}
