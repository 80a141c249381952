// dic_f32x2.cuh -- packed fp32 pairs (Blackwell FADD2 / FMUL2 / FFMA2) and the two-pixel form of the reference's
// bicubic arithmetic.
//
// Why it was tried: parity mode replays the reference's unfused fp32 operations (interpolation_class.cpp:79-138),
// ~210 of them per pixel, and the scalar loop is bound by ISSUE slots (ncu, profiles/r2_c4_batch_tiles_parity_ncu_full.txt:
// issue 83 %, FMA pipe 67 %). sm_100 has packed `*.f32x2` instructions that apply ONE IEEE operation to the two halves
// of a 64-bit register pair for one issue slot (tools/ubench.cu: FADD2 / FMUL2 1.95, FFMA2 1.79 warp-instructions per
// clock per SM = the scalar lane rate, i.e. two FMA-pipe passes each; tools/ulat.cu: the same 4-cycle dependent
// latency as the scalar forms). Each half is rounded exactly like the scalar instruction (`.rn`, or `.rm` for the
// floor trick), so evaluating TWO pixels of a lane's column side by side -- lo half = pixel A, hi half = pixel B --
// leaves every per-pixel bit unchanged and halves the instruction count of the arithmetic.
// What was measured: the loop then becomes bound by the FMA pipe, whose passes packing does not
// reduce, and it loses to the scalar loop by 9 % on c4. The code is the DIC_PARITY_LOOP == 3 option of dic_tiles.cuh.
//
// ptxas encodes a broadcast operand {v, v} and literal constants inside the instruction (R.F32 / UR.F32 / immediate),
// so `bc(v)` costs no register pair and no MOV.
#pragma once
#include <cstdint>

namespace dic {

typedef unsigned long long f2; // {lo, hi} fp32 pair in one 64-bit register pair

__device__ __forceinline__ f2 pk(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f2 bc(float v) { return pk(v, v); }
__device__ __forceinline__ float lo(f2 v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float hi(f2 v) { return __uint_as_float((uint32_t)(v >> 32)); }
__device__ __forceinline__ uint32_t lo_bits(f2 v) { return (uint32_t)v; }
__device__ __forceinline__ uint32_t hi_bits(f2 v) { return (uint32_t)(v >> 32); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f2 add2_rd(f2 a, f2 b) { f2 r; asm("add.rm.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

// A product that FEEDS AN ADDITION the reference rounds separately. ptxas (12.9) contracts `mul.rn.f32x2` followed by
// `add.rn.f32x2` into one FFMA2 -- the explicit `.rn` that keeps the scalar pair apart is not honoured for the packed
// forms, nor is -fmad=false, nor do fma(a, b, -0) / fma(m, 1, c) survive its simplifier (tools/ptxas_f32x2_fusion.cu
// shows all four). One half of the product therefore passes through an XOR with a kernel parameter that is always
// zero: an integer instruction on the ALU pipe whose result ptxas cannot prove equal to its input, so the multiply
// stays a multiply and is rounded on its own, exactly like __fmul_rn.
__device__ __forceinline__ f2 mul2_sep(f2 a, f2 b, uint32_t zero) {
  f2 r = mul2(a, b);
  uint32_t l = lo_bits(r), h = hi_bits(r);
  asm("xor.b32 %0, %0, %1;" : "+r"(h) : "r"(zero));
  f2 o; asm("mov.b64 %0, {%1, %2};" : "=l"(o) : "r"(l), "r"(h));
  return o;
}

// x-direction cubic of one window row for BOTH pixels (see row_coeffs_u8<PARITY>): every value is a small multiple
// of 1/2, so any association is exact; written in 12 operations (c0 = p0 - 2 c3 and c2 = c3 + d - c1 follow from
// the polynomial's value and slope at s = 1).
__device__ __forceinline__ void row_coeffs_u8_x2(uint32_t win_a, uint32_t win_b, f2 c[4]) {
  const f2 f0 = pk(__uint_as_float(__byte_perm(win_a, 0x4B000000u, 0x7650u)), __uint_as_float(__byte_perm(win_b, 0x4B000000u, 0x7650u)));
  const f2 f1 = pk(__uint_as_float(__byte_perm(win_a, 0x4B000000u, 0x7651u)), __uint_as_float(__byte_perm(win_b, 0x4B000000u, 0x7651u)));
  const f2 f2_ = pk(__uint_as_float(__byte_perm(win_a, 0x4B000000u, 0x7652u)), __uint_as_float(__byte_perm(win_b, 0x4B000000u, 0x7652u)));
  const f2 f3 = pk(__uint_as_float(__byte_perm(win_a, 0x4B000000u, 0x7653u)), __uint_as_float(__byte_perm(win_b, 0x4B000000u, 0x7653u)));
  const f2 a = sub2(f1, f2_), b = sub2(f3, f0), d = sub2(f1, f0), p0 = sub2(f0, bc(8388608.0f));
  c[3] = fma2(bc(1.5f), a, mul2(bc(0.5f), b));
  c[0] = fma2(bc(-2.f), c[3], p0);
  c[1] = fma2(bc(1.5f), d, fma2(bc(8.f), a, mul2(bc(2.5f), b)));
  c[2] = sub2(add2(c[3], d), c[1]);
}

// y pass of the coefficient stage for one column ik (monomial_from_samples on the four window rows), 11 exact
// operations: out[jk] = a[jk][ik].
__device__ __forceinline__ void monomial_from_rows_x2(f2 r0, f2 r1, f2 r2, f2 r3, f2 out[4]) {
  const f2 a = sub2(r1, r2), b = sub2(r3, r0), d = sub2(r1, r0);
  out[3] = fma2(bc(1.5f), a, mul2(bc(0.5f), b));
  out[0] = fma2(bc(-2.f), out[3], r0);
  out[1] = fma2(bc(1.5f), d, fma2(bc(8.f), a, mul2(bc(2.5f), b)));
  out[2] = sub2(add2(out[3], d), out[1]);
}

// parity_eval_f for two pixels: the 40 terms of interpolation_class.cpp:108-126 in the reference's order, each
// half rounded exactly like the scalar code (same exact shortcuts: x * 1 skipped, power-of-two scaling folded into
// an FMA, 3 a exact; `0 + first term` is the term).
__device__ __forceinline__ void parity_eval_x2(const f2 a[4][4], f2 dx, f2 dy, uint32_t zero, f2 &w, f2 &wx, f2 &wy) {
  f2 px[4], py[4];
  px[1] = dx; px[2] = mul2(dx, dx); px[3] = mul2(px[2], dx);
  py[1] = dy; py[2] = mul2(dy, dy); py[3] = mul2(py[2], dy);
  f2 rw = 0, rx = 0, ry = 0;
#pragma unroll
  for (int jk = 0; jk < 4; ++jk) {
#pragma unroll
    for (int ik = 0; ik < 4; ++ik) {
      const f2 c = a[jk][ik];
      // u = a * py[jk]: added as it is for ik == 0 (w) and ik == 1 (dw/dx), multiplied further otherwise
      const f2 u = jk == 0 ? c : (ik <= 1 ? mul2_sep(c, py[jk], zero) : mul2(c, py[jk]));
      const f2 term = ik == 0 ? u : mul2_sep(u, px[ik], zero);
      rw = (jk == 0 && ik == 0) ? term : add2(rw, term);
      if (ik == 1) rx = jk == 0 ? u : add2(rx, u);
      if (ik == 2) rx = fma2(bc(2.f), mul2(u, px[1]), rx);
      if (ik == 3) {
        f2 t = mul2(bc(3.f), c);
        t = jk == 0 ? t : mul2(t, py[jk]);
        rx = add2(rx, mul2_sep(t, px[2], zero));
      }
      if (jk == 1) ry = ik == 0 ? c : add2(ry, mul2_sep(c, px[ik], zero));
      if (jk == 2) {
        const f2 v = mul2(c, py[1]);
        ry = fma2(bc(2.f), ik == 0 ? v : mul2(v, px[ik]), ry);
      }
      if (jk == 3) {
        const f2 t3 = mul2(bc(3.f), c); // exact
        ry = add2(ry, ik == 0 ? mul2_sep(t3, py[2], zero) : mul2_sep(mul2(t3, py[2]), px[ik], zero));
      }
    }
  }
  w = rw; wx = rx; wy = ry;
}

} // namespace dic
