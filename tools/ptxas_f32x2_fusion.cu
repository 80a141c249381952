// tools/ptxas_f32x2_fusion.cu -- shows that ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 (and fma(a,b,-0) + add, mul + fma(m,1,c)) into
// one FFMA2 for sm_100a, with or without -fmad=false:  nvcc -gencode arch=compute_100a,code=sm_100a -cubin ... && cuobjdump -sass
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b){ u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b){ u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 sub2(u64 a, u64 b){ u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b){ u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c){ u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__global__ void k(const u64* in, u64* out){
  u64 a=in[0], b=in[1], c=in[2];
  out[0] = add2(c, fma2(a, b, pk(-0.0f, -0.0f)));          // V1
  out[1] = fma2(mul2(a, b), pk(1.0f, 1.0f), c);            // V2
  out[2] = sub2(c, mul2(a, b));                            // does sub fuse too?
  out[3] = add2(mul2(a, b), mul2(b, c));                   // two products
}
