// tools/ubench.cu -- issue-rate microbenchmarks that size the GN kernel's instruction budget on
// B200: scalar FFMA vs packed fma.rn.f32x2, unfused FMUL+FADD, PRMT, I2F, LDS.
// Prints warp-instructions / clk / SM for each.
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define UNROLL 16

__device__ __forceinline__ void fadd2(float2 &d, float2 a) {
  unsigned long long da = *reinterpret_cast<unsigned long long *>(&d), aa = *reinterpret_cast<unsigned long long *>(&a);
  asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(da) : "l"(aa));
  d = *reinterpret_cast<float2 *>(&da);
}
__device__ __forceinline__ void fmul2(float2 &d, float2 a) {
  unsigned long long da = *reinterpret_cast<unsigned long long *>(&d), aa = *reinterpret_cast<unsigned long long *>(&a);
  asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(da) : "l"(aa));
  d = *reinterpret_cast<float2 *>(&da);
}
__device__ __forceinline__ void ffma2(float2 &d, float2 a, float2 b) {
  unsigned long long da, aa, bb;
  aa = *reinterpret_cast<unsigned long long *>(&a);
  bb = *reinterpret_cast<unsigned long long *>(&b);
  da = *reinterpret_cast<unsigned long long *>(&d);
  asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(da) : "l"(aa), "l"(bb));
  d = *reinterpret_cast<float2 *>(&da);
}

template <int MODE> __global__ void k(float *out, float a, float b, long long *cyc) {
  float acc[UNROLL];
  float2 acc2[UNROLL];
  unsigned int iacc[UNROLL];
  __shared__ float sm[1024];
  sm[threadIdx.x] = a;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < UNROLL; ++i) { acc[i] = a + i; acc2[i] = make_float2(a + i, b + i); iacc[i] = threadIdx.x + i; }
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) {
      if (MODE == 0) acc[i] = fmaf(acc[i], a, b);
      if (MODE == 1) ffma2(acc2[i], make_float2(a, b), make_float2(b, a));
      if (MODE == 2) acc[i] = __fadd_rn(__fmul_rn(acc[i], a), b);
      if (MODE == 3) iacc[i] = __byte_perm(iacc[i], 0x4B000000u, 0x7650u | (iacc[i] & 3));
      if (MODE == 4) acc[i] += (float)(__float_as_uint(acc[i]) & 0xff);
      if (MODE == 5) acc[i] += sm[(threadIdx.x + i * 32 + __float_as_uint(acc[i]) ) & 1023];
      if (MODE == 6) acc[i] = __fadd_rn(acc[i], a);
      if (MODE == 7) fadd2(acc2[i], make_float2(a, b));
      if (MODE == 8) fmul2(acc2[i], make_float2(a, b));
      if (MODE == 9) { if (i & 1) fadd2(acc2[i], make_float2(a, b)); else acc[i] = __fadd_rn(acc[i], a); }          // 1 packed : 1 scalar
      if (MODE == 10) { fadd2(acc2[i], make_float2(a, b)); iacc[i] = iacc[i] + (iacc[i] >> 3) + 7u; }                // packed + ALU (IADD3 / SHF)
      if (MODE == 11) { if ((i & 3) == 3) acc[i] = __fadd_rn(acc[i], a); else fmul2(acc2[i], make_float2(a, b)); }  // 3 packed : 1 scalar
      if (MODE == 12) { fmul2(acc2[i], make_float2(a, b)); iacc[i] = __byte_perm(iacc[i], 0x4B000000u, 0x7650u | (iacc[i] & 3)); }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < UNROLL; ++i) s += acc[i] + acc2[i].x + acc2[i].y + iacc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE> void run(const char *name, int inst_per_iter) {
  float *out; long long *cyc, h;
  int threads = 1024, blocks = 148;
  cudaMalloc(&out, sizeof(float) * threads * blocks);
  cudaMalloc(&cyc, 8);
  k<MODE><<<blocks, threads>>>(out, 1.0001f, 0.5f, cyc);
  k<MODE><<<blocks, threads>>>(out, 1.0001f, 0.5f, cyc);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  double warp_inst = (double)ITERS * UNROLL * inst_per_iter * (threads / 32);
  printf("%-28s %8.3f warp-inst/clk/SM  (%lld cycles)\n", name, warp_inst / (double)h, h);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("FFMA (3 reg)", 1);
  run<1>("FFMA2 (fma.rn.f32x2)", 1);
  run<2>("FMUL+FADD unfused", 2);
  run<3>("PRMT + LOP", 2);
  run<4>("I2F + LOP + FADD", 3);
  run<5>("LDS.32 + addr + FADD", 4);
  run<6>("FADD", 1);
  run<7>("FADD2 (add.rn.f32x2)", 1);
  run<8>("FMUL2 (mul.rn.f32x2)", 1);
  run<9>("FADD2 : FADD 1:1", 1);
  run<10>("FADD2 + SHF + IADD3", 3);
  run<11>("FMUL2 : FADD 3:1", 1);
  run<12>("FMUL2 + LOP + PRMT", 3);
  cudaError_t e = cudaGetLastError();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
