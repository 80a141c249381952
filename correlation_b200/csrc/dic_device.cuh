// dic_device.cuh -- device-side building blocks of the B200 DIC engine (sm_100a).
//
// Everything on the Gauss-Newton hot path of the reference CPU engine, fused per pixel:
//   warp      model_class.cpp:150-202 (+ U/UV/UVQ :48-148, + 12-parameter extension)
//   bicubic   interpolation_class.cpp:79-138 / :243-336 (bilinear :140-195, nearest :197-226)
//   residual  interpolation_class.cpp:671-764  (V, H = grad w . dT/dp, A += H H^T, b += H V, chi += V^2)
//   solve     correlation_class.cpp:642-768    (scale, damp, solve) -- single-warp Cholesky here
//   LM loop   correlation_class.cpp:349-640    -- on-device state machine, never returns to host
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dic_b200.h"

namespace dic {

constexpr int kMaxLevels = DIC_MAX_LEVELS;
constexpr int kMaxParams = DIC_MAX_PARAMS;
constexpr int kThreads = 256; // threads per CTA of the GN kernels
constexpr int kMaxMarks = 128;
constexpr int kMaxRanks = 8;
constexpr int kMaxCtaMarks = 1024;

// ------------------------------------------------------------------ descriptors

struct LevelImage {
  const uint8_t *ptr; // u8, row-major, `pitch` bytes per row (multiple of 128, >= cols * colors + 16)
  int rows, cols, pitch;
  int colors;         // 1, or 3 interleaved channels (the reference's color_color mode, enums.hpp:37)
};

// Per-sector device record (one per domain / subdivision subset).
struct SectorDev {
  const float2 *xy[kMaxLevels]; // per-level pixel list (x, y) in level units
  int n[kMaxLevels];       // pixels this GPU owns at each level
  int n_total[kMaxLevels]; // pixels of the whole domain (== n unless the domain is row-split over GPUs)
  float cx, cy; // level-0 centre
};

enum Phase { PH_INIT = 0, PH_REDO = 1, PH_TENT = 2 };

// Levenberg-Marquardt state of one sector (global memory in grid mode, shared in batch mode).
struct LMState {
  float p[kMaxParams];         // parameters the next evaluation uses (level units)
  float mp[kMaxParams];        // the reference's `model_parameters` (what it would return)
  float last_good[kMaxParams]; // correlation_class.cpp:369-371
  float tentative[kMaxParams];
  float saved[kMaxParams];
  float lambda, last_good_chi, scaling;
  int level, level_old, iteration, use_saved, phase, done;
  int error_code, reached_iterations;
  int evals[kMaxLevels], iters[kMaxLevels];
};

// One rank's mailbox: slot [parity][sender rank] holds the sender's sums of evaluation `seq`.
struct Mailbox {
  float sums[2][kMaxRanks][96];
  unsigned int seq[2][kMaxRanks]; // written last (release): evaluation number + 1
};

// Grid-wide reduction / barrier scratch (grid mode). Every evaluation ends with an all-reduce: each CTA
// adds its sums to acc[e % 3] with fp64 atomics and bumps `arrive` (release); every CTA then waits for
// the count, reads the same totals and runs the SAME LM step + solve on its own copy of the state in
// shared memory -- bitwise identical decisions everywhere, no master, no publish / re-read hop.
// `arrive` and `acc` are zero between launches: CTA 0 clears acc[(e + 2) % 3] after barrier e and the last CTA to
// leave the kernel clears the rest (see `departed`).
struct GridWork {
  unsigned int arrive;     // CTAs that have delivered their sums, cumulative over the launch
  unsigned int departed;   // CTAs that have left the kernel: the last one re-zeroes arrive / abort / acc,
                           // so a launch needs no memset in front of it (nothing of a correlate touches a copy engine)
  int abort;               // set when a grid-barrier wait timed out (never expected)
  unsigned int slow_units; // units of the last launch that took the per-pixel (non-staged) path
  int img_error;           // set by the pyramid kernel when a TMA transfer never landed (sticky until read by the host)
  unsigned int next_sector;    // batch form: ticket counter of the sector queue ...
  unsigned int batch_departed; // ... and the CTAs that have left (the last one zeroes both)
  double acc[3][96];       // grid-wide sums of evaluations e % 3 (fp64 atomics)
  // row-split of one domain over several GPUs: CTA 0 of every rank adds the rank's sums to
  // every peer's mailbox over NVLink, then all ranks add the rows in rank order (bitwise identical)
  int rs_rank, rs_world;
  unsigned int rs_seq;             // evaluations exchanged so far (same on all ranks); carried from launch to launch
  struct Mailbox *rs_local;        // this rank's mailbox (peers write into it)
  struct Mailbox *rs_peer[kMaxRanks]; // peer-mapped mailboxes, index = rank
  // CTA 0's timeline of the last launch (ns, %globaltimer): per evaluation
  // [0] pass started, [1] own pass done, [2] all CTAs arrived, [3] LM step done
  int n_marks;
  unsigned long long marks[kMaxMarks][4];
  unsigned long long cta_done[kMaxCtaMarks]; // per CTA: when its pass of the LAST evaluation ended (load-balance probe)
  unsigned int cta_smid[kMaxCtaMarks];       // per CTA: the SM it runs on (same probe)
};

struct SolveSettings {
  LevelImage und[kMaxLevels];
  LevelImage def[kMaxLevels];
  int start, step, stop;
  int max_iters;
  float precision;
  unsigned int opaque_zero; // always 0; a value the compiler cannot know (dic_f32x2.cuh, mul2_sep)
};

// ------------------------------------------------------------------ small helpers

template <int NP> struct Acc {
  static constexpr int kA = NP * (NP + 1) / 2;
  static constexpr int kB = kA;          // offset of b
  static constexpr int kChi = kA + NP;   // offset of chi
  static constexpr int kOob = kChi + 1;  // offset of the out-of-image counter
  static constexpr int kN = kOob + 1;
};

__host__ __device__ constexpr int model_nparams(int model) {
  return model == DIC_FM_U ? 1 : model == DIC_FM_UV ? 2 : model == DIC_FM_UVQ ? 3
       : model == DIC_FM_UVUxUyVxVy ? 6 : 12;
}

__device__ __forceinline__ float u8_to_float(uint32_t word, int byte) {
  // PRMT builds 0x4B0000bb = 2^23 + bb, one FADD removes the bias: no I2F on the hot path
  uint32_t r = __byte_perm(word, 0x4B000000u, 0x7650u | byte);
  return __uint_as_float(r) - 8388608.0f;
}

// 4 consecutive pixels starting at byte address (row + x), any alignment: two aligned 32-bit
// loads + funnel shift. Rows are padded so the second word is always readable.
__device__ __forceinline__ uint32_t load4_u8(const uint8_t *row, int x) {
  const uint32_t *w = reinterpret_cast<const uint32_t *>(row + (x & ~3));
  uint32_t lo = __ldg(w), hi = __ldg(w + 1);
  return __funnelshift_r(lo, hi, (x & 3) * 8);
}

// ------------------------------------------------------------------ TMA / mbarrier (sm_90+ PTX)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// bounded by TIME (2 s on %globaltimer, looked at every 4096 polls): a transfer that never lands must end the
// launch with an error code, not hang the GPU and not poison the context with a trap. Returns false on timeout;
// the caller raises its kernel's error flag (DIC_ERROR_CUDA) and stops consuming staged data.
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return true;
  unsigned int spins = 0;
  unsigned long long t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0xfffu) == 0) {
      const unsigned long long now = global_timer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 2000000000ull) return false;
    }
  }
  return true;
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int x, int y, uint64_t *bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}

// ------------------------------------------------------------------ interpolation

// Reference arithmetic (parity mode): monomial bicubic in dx = 1 + frac, coefficients
// a = (B (x) B) v with B the integer inverse of the 1-D Hermite constraint matrix
// (M of interpolation_class.cpp:539-558 equals that Kronecker product exactly; every
// intermediate is a multiple of 1/4 below 2^22, so the coefficient stage is exact in fp32 in any
// order), then the 40 terms of :108-126 in the reference's order with unfused mul/add.
__device__ __forceinline__ void hermite_to_monomial(float f1, float f2, float d1, float d2,
                                                    float c[4]) {
  c[0] = -4.f * f1 + 5.f * f2 - 4.f * d1 - 2.f * d2;
  c[1] = 12.f * f1 - 12.f * f2 + 8.f * d1 + 5.f * d2;
  c[2] = -9.f * f1 + 9.f * f2 - 5.f * d1 - 4.f * d2;
  c[3] = 2.f * f1 - 2.f * f2 + d1 + d2;
}

// hermite_to_monomial(p1, p2, (p2 - p0) / 2, (p3 - p1) / 2) written with differences: 11 operations
// instead of 20 (c0 and c2 from the cubic's value and slope at s = 1), identical results because every intermediate is exact (see above).
__device__ __forceinline__ void monomial_from_samples(float p0, float p1, float p2, float p3, float c[4]) {
  const float a = p1 - p2, b = p3 - p0, d = p1 - p0;
  c[3] = fmaf(1.5f, a, 0.5f * b);
  c[0] = fmaf(-2.f, c[3], p0);                  // p0 - b - 3 a
  c[1] = fmaf(1.5f, d, fmaf(8.f, a, 2.5f * b));
  c[2] = (c[3] + d) - c[1];                     // -2 b - 6.5 a - 0.5 d
}

// The 40 terms of interpolation_class.cpp:108-126 in the reference's order, unfused mul / add.
// Exact shortcuts only: x * 1 is skipped, (2 a) * y == 2 (a * y) and acc + 2 t == fma(2, t, acc)
// bit for bit (power-of-two scaling commutes with rounding), 3 a is exact (|a| < 2^20, quarter units).
// fix, fiy = (float)ix, (float)iy (the caller may have them already: floor_magic returns exactly that value).
__device__ __forceinline__ void parity_eval_f(const float a[4][4], float xdef, float ydef, float fix, float fiy,
                                              float &w, float &wx, float &wy) {
  const float dx = __fadd_rn(__fsub_rn(xdef, fix), 1.f);
  const float dy = __fadd_rn(__fsub_rn(ydef, fiy), 1.f);
  float px[4], py[4];
  px[0] = 1.f; px[1] = dx; px[2] = __fmul_rn(dx, dx); px[3] = __fmul_rn(px[2], dx);
  py[0] = 1.f; py[1] = dy; py[2] = __fmul_rn(dy, dy); py[3] = __fmul_rn(py[2], dy);
  float rw = 0.f, rx = 0.f, ry = 0.f;
#pragma unroll
  for (int jk = 0; jk < 4; ++jk) {
#pragma unroll
    for (int ik = 0; ik < 4; ++ik) {
      const float c = a[jk][ik];
      const float u = jk == 0 ? c : __fmul_rn(c, py[jk]);            // a * py[jk]
      rw = __fadd_rn(rw, ik == 0 ? u : __fmul_rn(u, px[ik]));        // w += (a * py[jk]) * px[ik]
      // wx += ((ik * a) * py[jk]) * px[ik - 1]
      if (ik == 1) rx = __fadd_rn(rx, u);
      if (ik == 2) rx = __fmaf_rn(2.f, __fmul_rn(u, px[1]), rx);
      if (ik == 3) {
        float t = __fmul_rn(3.f, c);
        t = jk == 0 ? t : __fmul_rn(t, py[jk]);
        rx = __fadd_rn(rx, __fmul_rn(t, px[2]));
      }
      // wy += ((jk * a) * py[jk - 1]) * px[ik]
      if (jk == 1) ry = __fadd_rn(ry, ik == 0 ? c : __fmul_rn(c, px[ik]));
      if (jk == 2) {
        const float v = __fmul_rn(c, py[1]);
        ry = __fmaf_rn(2.f, ik == 0 ? v : __fmul_rn(v, px[ik]), ry);
      }
      if (jk == 3) {
        const float t = __fmul_rn(__fmul_rn(3.f, c), py[2]);
        ry = __fadd_rn(ry, ik == 0 ? t : __fmul_rn(t, px[ik]));
      }
    }
  }
  w = rw; wx = rx; wy = ry;
}
__device__ __forceinline__ void parity_eval(const float a[4][4], float xdef, float ydef, int ix, int iy,
                                            float &w, float &wx, float &wy) {
  parity_eval_f(a, xdef, ydef, (float)ix, (float)iy, w, wx, wy);
}

// y pass of the separable coefficient stage: rows r0..r3 hold the x-direction monomial
// coefficients (in s = 1 + t) of image rows iy-1 .. iy+2.
__device__ __forceinline__ void bicubic_parity_rows(const float *r0, const float *r1, const float *r2,
                                                    const float *r3, float xdef, float ydef, int ix, int iy,
                                                    float &w, float &wx, float &wy) {
  float a[4][4];
#pragma unroll
  for (int ik = 0; ik < 4; ++ik) {
    float c[4];
    monomial_from_samples(r0[ik], r1[ik], r2[ik], r3[ik], c);
#pragma unroll
    for (int jk = 0; jk < 4; ++jk) a[jk][ik] = c[jk];
  }
  parity_eval(a, xdef, ydef, ix, iy, w, wx, wy);
}
__device__ __forceinline__ void bicubic_parity_rows_f(const float *r0, const float *r1, const float *r2,
                                                      const float *r3, float xdef, float ydef, float fix, float fiy,
                                                      float &w, float &wx, float &wy) {
  float a[4][4];
#pragma unroll
  for (int ik = 0; ik < 4; ++ik) {
    float c[4];
    monomial_from_samples(r0[ik], r1[ik], r2[ik], r3[ik], c);
#pragma unroll
    for (int jk = 0; jk < 4; ++jk) a[jk][ik] = c[jk];
  }
  parity_eval_f(a, xdef, ydef, fix, fiy, w, wx, wy);
}

__device__ __forceinline__ void bicubic_parity(const float p[4][4], float xdef, float ydef, int ix,
                                               int iy, float &w, float &wx, float &wy) {
  // x pass: for each image row r, cubic in x through columns 1,2 with central-difference slopes
  float cx[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
    hermite_to_monomial(p[r][1], p[r][2], (p[r][2] - p[r][0]) * 0.5f, (p[r][3] - p[r][1]) * 0.5f,
                        cx[r]);
  bicubic_parity_rows(cx[0], cx[1], cx[2], cx[3], xdef, ydef, ix, iy, w, wx, wy);
}

// Fast mode on row coefficients: value and x-derivative of each row by Horner, then Catmull-Rom
// weights in y.
__device__ __forceinline__ void cr_weights(float t, float w[4], float d[4]);
__device__ __forceinline__ void bicubic_fast_rows(const float *r0, const float *r1, const float *r2,
                                                  const float *r3, float tx, float ty, float &w, float &wx,
                                                  float &wy) {
  const float t2 = tx + tx, t3 = 3.f * tx * tx;
  float wyw[4], wyd[4];
  cr_weights(ty, wyw, wyd);
  const float *rows[4] = {r0, r1, r2, r3};
  float rw = 0.f, rx = 0.f, ry = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float *c = rows[k];
    const float v = fmaf(fmaf(fmaf(c[3], tx, c[2]), tx, c[1]), tx, c[0]);
    const float dv = fmaf(c[3], t3, fmaf(c[2], t2, c[1]));
    rw = fmaf(wyw[k], v, rw);
    rx = fmaf(wyw[k], dv, rx);
    ry = fmaf(wyd[k], v, ry);
  }
  w = rw; wx = rx; wy = ry;
}

// Fast mode: the same interpolant (Keys a = -1/2, Catmull-Rom) in weight form.
__device__ __forceinline__ void cr_weights(float t, float w[4], float d[4]) {
  float t2 = t * t;
  w[0] = ((-0.5f * t + 1.0f) * t - 0.5f) * t;
  w[1] = (1.5f * t - 2.5f) * t2 + 1.0f;
  w[2] = ((-1.5f * t + 2.0f) * t + 0.5f) * t;
  w[3] = (0.5f * t - 0.5f) * t2;
  d[0] = (-1.5f * t + 2.0f) * t - 0.5f;
  d[1] = (4.5f * t - 5.0f) * t;
  d[2] = (-4.5f * t + 4.0f) * t + 0.5f;
  d[3] = (1.5f * t - 1.0f) * t;
}

__device__ __forceinline__ void bicubic_fast(const float p[4][4], float tx, float ty, float &w,
                                             float &wx, float &wy) {
  float wxw[4], wxd[4], wyw[4], wyd[4];
  cr_weights(tx, wxw, wxd);
  cr_weights(ty, wyw, wyd);
  float rw = 0.f, rx = 0.f, ry = 0.f;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    float s = p[r][0] * wxw[0] + p[r][1] * wxw[1] + p[r][2] * wxw[2] + p[r][3] * wxw[3];
    float sd = p[r][0] * wxd[0] + p[r][1] * wxd[1] + p[r][2] * wxd[2] + p[r][3] * wxd[3];
    rw += wyw[r] * s;
    rx += wyw[r] * sd;
    ry += wyd[r] * s;
  }
  w = rw; wx = rx; wy = ry;
}

// One deformed-image sample. Returns false (and w = wx = wy = 0) when out of image, with the
// reference's own bounds tests.
template <int INTERP, int MODE>
__device__ __forceinline__ bool sample_def(const LevelImage &img, float xdef, float ydef, float &w,
                                           float &wx, float &wy) {
  if (INTERP == DIC_IM_BICUBIC) {
    // interpolation_class.cpp:82-83
    if (!(xdef > 1.f && ydef > 1.f && xdef < (float)img.cols - 2.f && ydef < (float)img.rows - 2.f)) {
      w = wx = wy = 0.f;
      return false;
    }
    int ix = (int)xdef, iy = (int)ydef;
    float p[4][4];
    const uint8_t *row = img.ptr + (size_t)(iy - 1) * img.pitch;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      uint32_t v = load4_u8(row + (size_t)r * img.pitch, ix - 1);
#pragma unroll
      for (int c = 0; c < 4; ++c) p[r][c] = u8_to_float(v, c);
    }
    if (MODE == DIC_MODE_PARITY) bicubic_parity(p, xdef, ydef, ix, iy, w, wx, wy);
    else bicubic_fast(p, xdef - (float)ix, ydef - (float)iy, w, wx, wy);
    return true;
  } else if (INTERP == DIC_IM_BILINEAR) {
    // interpolation_class.cpp:143-144, :338-374
    if (!(xdef > 0.f && ydef > 0.f && xdef < (float)(img.cols - 1) && ydef < (float)(img.rows - 1))) {
      w = wx = wy = 0.f;
      return false;
    }
    int ix = (int)xdef, iy = (int)ydef;
    const uint8_t *r0 = img.ptr + (size_t)iy * img.pitch + ix;
    float w00 = (float)__ldg(r0), w10 = (float)__ldg(r0 + 1);
    float w01 = (float)__ldg(r0 + img.pitch), w11 = (float)__ldg(r0 + img.pitch + 1);
    float a0 = w00, a1 = __fsub_rn(w10, w00), a2 = __fsub_rn(w01, w00);
    float a3 = __fadd_rn(__fsub_rn(__fsub_rn(w11, w10), w01), w00);
    float dx = __fsub_rn(xdef, (float)ix), dy = __fsub_rn(ydef, (float)iy);
    // reference order: (jk,ik) = (0,0),(0,1),(1,0),(1,1); w += a*py*px
    float rw = a0;
    rw = __fadd_rn(rw, __fmul_rn(a1, dx));
    rw = __fadd_rn(rw, __fmul_rn(a2, dy));
    rw = __fadd_rn(rw, __fmul_rn(__fmul_rn(a3, dy), dx));
    float rx = __fadd_rn(a1, __fmul_rn(a3, dy));
    float ry = __fadd_rn(a2, __fmul_rn(a3, dx));
    w = rw; wx = rx; wy = ry;
    return true;
  } else {
    // interpolation_class.cpp:200-201, :376-406
    if (!(xdef > 0.f && ydef > 0.f && xdef < (float)(img.cols - 1) && ydef < (float)(img.rows - 1))) {
      w = wx = wy = 0.f;
      return false;
    }
    int ix = (int)(xdef + 0.5f), iy = (int)(ydef + 0.5f);
    const uint8_t *r0 = img.ptr + (size_t)iy * img.pitch + ix;
    float w00 = (float)__ldg(r0), w10 = (float)__ldg(r0 + 1), w01 = (float)__ldg(r0 + img.pitch);
    w = w00; wx = __fsub_rn(w10, w00); wy = __fsub_rn(w01, w00);
    return true;
  }
}

// ------------------------------------------------------------------ warp + residual row

// Deformed position and steepest-descent row H = wx * dTx/dp + wy * dTy/dp for one pixel.
// Parity mode keeps the reference's left-to-right fp32 evaluation of the warp.
template <int MODEL, int MODE>
__device__ __forceinline__ void warp_point(const float *p, float x, float y, float cx, float cy,
                                           float &xd, float &yd, float &dx, float &dy) {
  dx = __fsub_rn(x, cx);
  dy = __fsub_rn(y, cy);
  if (MODEL == DIC_FM_U) {
    xd = __fadd_rn(x, p[0]); yd = y;
  } else if (MODEL == DIC_FM_UV) {
    xd = __fadd_rn(x, p[0]); yd = __fadd_rn(y, p[1]);
  } else if (MODEL == DIC_FM_UVQ) {
    xd = __fsub_rn(__fadd_rn(x, p[0]), __fmul_rn(p[2], dy));
    yd = __fadd_rn(__fadd_rn(y, p[1]), __fmul_rn(p[2], dx));
  } else if (MODEL == DIC_FM_UVUxUyVxVy) {
    if (MODE == DIC_MODE_PARITY) {
      xd = __fadd_rn(__fadd_rn(__fadd_rn(x, p[0]), __fmul_rn(p[2], dx)), __fmul_rn(p[3], dy));
      yd = __fadd_rn(__fadd_rn(__fadd_rn(y, p[1]), __fmul_rn(p[4], dx)), __fmul_rn(p[5], dy));
    } else {
      xd = fmaf(p[3], dy, fmaf(p[2], dx, x + p[0]));
      yd = fmaf(p[5], dy, fmaf(p[4], dx, y + p[1]));
    }
  } else {
    if (MODE == DIC_MODE_PARITY) {
      float t = __fadd_rn(__fadd_rn(__fadd_rn(x, p[0]), __fmul_rn(p[2], dx)), __fmul_rn(p[3], dy));
      t = __fadd_rn(t, __fmul_rn(__fmul_rn(__fmul_rn(0.5f, p[6]), dx), dx));
      t = __fadd_rn(t, __fmul_rn(__fmul_rn(p[7], dx), dy));
      xd = __fadd_rn(t, __fmul_rn(__fmul_rn(__fmul_rn(0.5f, p[8]), dy), dy));
      t = __fadd_rn(__fadd_rn(__fadd_rn(y, p[1]), __fmul_rn(p[4], dx)), __fmul_rn(p[5], dy));
      t = __fadd_rn(t, __fmul_rn(__fmul_rn(__fmul_rn(0.5f, p[9]), dx), dx));
      t = __fadd_rn(t, __fmul_rn(__fmul_rn(p[10], dx), dy));
      yd = __fadd_rn(t, __fmul_rn(__fmul_rn(__fmul_rn(0.5f, p[11]), dy), dy));
    } else {
      float hxx = 0.5f * dx * dx, hxy = dx * dy, hyy = 0.5f * dy * dy;
      xd = x + p[0] + p[2] * dx + p[3] * dy + p[6] * hxx + p[7] * hxy + p[8] * hyy;
      yd = y + p[1] + p[4] * dx + p[5] * dy + p[9] * hxx + p[10] * hxy + p[11] * hyy;
    }
  }
}

template <int MODEL>
__device__ __forceinline__ void descent_row(float wx, float wy, float dx, float dy, float *H) {
  if (MODEL == DIC_FM_U) {
    H[0] = wx;
  } else if (MODEL == DIC_FM_UV) {
    H[0] = wx; H[1] = wy;
  } else if (MODEL == DIC_FM_UVQ) {
    H[0] = wx; H[1] = wy; H[2] = wy * dx - wx * dy;
  } else {
    H[0] = wx; H[1] = wy; H[2] = wx * dx; H[3] = wx * dy; H[4] = wy * dx; H[5] = wy * dy;
    if (MODEL == DIC_FM_QUADRATIC) {
      float hxx = 0.5f * dx * dx, hxy = dx * dy, hyy = 0.5f * dy * dy;
      H[6] = wx * hxx; H[7] = wx * hxy; H[8] = wx * hyy;
      H[9] = wy * hxx; H[10] = wy * hxy; H[11] = wy * hyy;
    }
  }
}

// One pixel of one evaluation, accumulated into acc[] (upper-triangular A, b, chi, oob count).
template <int MODEL, int INTERP, int MODE>
__device__ __forceinline__ void accumulate_pixel(const LevelImage &und, const LevelImage &def,
                                                 const float *p, float cx, float cy, float x, float y,
                                                 float *acc) {
  constexpr int NP = model_nparams(MODEL);
  using L = Acc<NP>;
  float xd, yd, dx, dy, w, wx, wy;
  warp_point<MODEL, MODE>(p, x, y, cx, cy, xd, yd, dx, dy);
  bool inside = sample_def<INTERP, MODE>(def, xd, yd, w, wx, wy);
  // nearest reference-image pixel, interpolation_class.cpp:701-714
  int uix = (int)(x + 0.5f), uiy = (int)(y + 0.5f);
  float und_w = (float)__ldg(und.ptr + (size_t)uiy * und.pitch + uix);
  float V = und_w - w;
  acc[L::kChi] = fmaf(V, V, acc[L::kChi]);
  if (!inside) acc[L::kOob] += 1.f;
  float H[NP];
  descent_row<MODEL>(wx, wy, dx, dy, H);
  int k = 0;
#pragma unroll
  for (int p1 = 0; p1 < NP; ++p1) {
    acc[L::kB + p1] = fmaf(H[p1], V, acc[L::kB + p1]);
#pragma unroll
    for (int p2 = p1; p2 < NP; ++p2) {
      acc[k] = fmaf(H[p1], H[p2], acc[k]);
      ++k;
    }
  }
}

// ---- three-channel colour images (interpolation_class.cpp:712-750: the per-colour loop of one evaluation).
// Column x of channel c sits at byte x * mult + add of its row. The reference's bicubic and bilinear coefficient
// builders index the deformed image with `color = number_of_colors + color_in; index_ix = ix * color`
// (interpolation_class.cpp:268-273, :356-359): mult = 3 + c, add = 0 -- right for channel 0 only, and what the CPU
// engine EXECUTES for channels 1 and 2; parity means the same bytes here. Nearest uses ix * 3 + c (:391-398), and so
// does the reference-image pixel (:701-704). Rows are padded and zero-filled, so the far reads of channel 2
// (5 bytes per column) stay inside the level's allocation.
template <int INTERP, int MODE>
__device__ __forceinline__ bool sample_def_color(const LevelImage &img, float xdef, float ydef, int c, float &w,
                                                 float &wx, float &wy) {
  if (INTERP == DIC_IM_BICUBIC) {
    if (!(xdef > 1.f && ydef > 1.f && xdef < (float)img.cols - 2.f && ydef < (float)img.rows - 2.f)) {
      w = wx = wy = 0.f;
      return false;
    }
    const int ix = (int)xdef, iy = (int)ydef, mult = 3 + c;
    float p[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const uint8_t *row = img.ptr + (size_t)(iy - 1 + r) * img.pitch;
#pragma unroll
      for (int k = 0; k < 4; ++k) p[r][k] = (float)__ldg(row + (ix - 1 + k) * mult);
    }
    if (MODE == DIC_MODE_PARITY) bicubic_parity(p, xdef, ydef, ix, iy, w, wx, wy);
    else bicubic_fast(p, xdef - (float)ix, ydef - (float)iy, w, wx, wy);
    return true;
  }
  if (!(xdef > 0.f && ydef > 0.f && xdef < (float)(img.cols - 1) && ydef < (float)(img.rows - 1))) {
    w = wx = wy = 0.f;
    return false;
  }
  if (INTERP == DIC_IM_BILINEAR) {
    const int ix = (int)xdef, iy = (int)ydef, mult = 3 + c;
    const uint8_t *r0 = img.ptr + (size_t)iy * img.pitch, *r1 = r0 + img.pitch;
    const float w00 = (float)__ldg(r0 + ix * mult), w10 = (float)__ldg(r0 + (ix + 1) * mult);
    const float w01 = (float)__ldg(r1 + ix * mult), w11 = (float)__ldg(r1 + (ix + 1) * mult);
    const float a0 = w00, a1 = __fsub_rn(w10, w00), a2 = __fsub_rn(w01, w00);
    const float a3 = __fadd_rn(__fsub_rn(__fsub_rn(w11, w10), w01), w00);
    const float dx = __fsub_rn(xdef, (float)ix), dy = __fsub_rn(ydef, (float)iy);
    float rw = a0;
    rw = __fadd_rn(rw, __fmul_rn(a1, dx));
    rw = __fadd_rn(rw, __fmul_rn(a2, dy));
    rw = __fadd_rn(rw, __fmul_rn(__fmul_rn(a3, dy), dx));
    w = rw; wx = __fadd_rn(a1, __fmul_rn(a3, dy)); wy = __fadd_rn(a2, __fmul_rn(a3, dx));
    return true;
  }
  const int ix = (int)(xdef + 0.5f), iy = (int)(ydef + 0.5f);
  const uint8_t *r0 = img.ptr + (size_t)iy * img.pitch + ix * 3 + c;
  const float w00 = (float)__ldg(r0), w10 = (float)__ldg(r0 + 3), w01 = (float)__ldg(r0 + img.pitch);
  w = w00; wx = __fsub_rn(w10, w00); wy = __fsub_rn(w01, w00);
  return true;
}

template <int MODEL, int INTERP, int MODE>
__device__ __forceinline__ void accumulate_pixel_color(const LevelImage &und, const LevelImage &def,
                                                       const float *p, float cx, float cy, float x, float y,
                                                       float *acc) {
  constexpr int NP = model_nparams(MODEL);
  using L = Acc<NP>;
  float xd, yd, dx, dy;
  warp_point<MODEL, MODE>(p, x, y, cx, cy, xd, yd, dx, dy);
  const int uix = (int)(x + 0.5f), uiy = (int)(y + 0.5f);
  const uint8_t *upx = und.ptr + (size_t)uiy * und.pitch + uix * 3;
#pragma unroll 1
  for (int c = 0; c < 3; ++c) {
    float w, wx, wy;
    const bool inside = sample_def_color<INTERP, MODE>(def, xd, yd, c, w, wx, wy);
    const float V = (float)__ldg(upx + c) - w;
    acc[L::kChi] = fmaf(V, V, acc[L::kChi]);
    if (!inside) acc[L::kOob] += 1.f;
    float H[NP];
    descent_row<MODEL>(wx, wy, dx, dy, H);
    int k = 0;
#pragma unroll
    for (int p1 = 0; p1 < NP; ++p1) {
      acc[L::kB + p1] = fmaf(H[p1], V, acc[L::kB + p1]);
#pragma unroll
      for (int p2 = p1; p2 < NP; ++p2) {
        acc[k] = fmaf(H[p1], H[p2], acc[k]);
        ++k;
      }
    }
  }
}

// ------------------------------------------------------------------ reductions

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block reduction of N per-thread accumulators into out[0..N) (shared). red: [warps][N] scratch.
template <int N>
__device__ __forceinline__ void block_reduce(float *acc, float *red, float *out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kWarps = kThreads / 32;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    float v = warp_sum(acc[k]);
    if (lane == 0) red[warp * N + k] = v;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < N; k += kThreads) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += red[w * N + k];
    out[k] = s;
  }
  __syncthreads();
}

// ------------------------------------------------------------------ solve (one warp)

// Damped normal equations of correlation_class.cpp:642-688: A, b scaled by 1/N, diagonal times
// (1 + lambda). Solved by a Jacobi-equilibrated Cholesky factorisation held in registers, one
// matrix row per lane, columns exchanged by warp shuffles (replaces the reference's cuSOLVER
// potrf/potrs, cuda_solver.cu:119-149, and the CPU's Eigen QR). tot: packed upper A, then b
// (shared memory). smem: NP*NP floats of scratch for the transposed back-substitution.
// Rank deficiency follows the CPU engine's column-pivoted QR (Eigen 3.4.0 ColPivHouseholderQR behind
// correlation_class.cpp:742-747, restated in oracle/qr_colpiv.h):
//   * a direction whose diagonal or pivot is not positive while others are (one zero row of A: no gradient along
//     that parameter) is rank-truncated there -- it gets a ZERO step and the others are solved;
//   * an identically zero matrix (textureless / saturated subset) is NOT truncated by Eigen (its threshold is
//     relative to the largest column norm, 0 < 0 is false), the triangular solve divides 0 by 0 and the step is
//     NaN: the next evaluation is out of the image and the pyramid ends with error 2 and NaN parameters. The same
//     happens here (dp = NaN), so that the report matches the reference's even on such subsets.
// Always returns true (kept bool for the call sites).
template <int NP>
__device__ bool warp_solve(const float *tot, float scaling, float lambda, float *smem, float *dp) {
  const int lane = threadIdx.x & 31;
  const int i = lane < NP ? lane : NP - 1;
  const unsigned full = 0xffffffffu;
  float a[NP];
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    int r = i < j ? i : j, c = i < j ? j : i;
    float v = tot[r * NP - r * (r - 1) / 2 + (c - r)] * scaling;
    a[j] = (i == j) ? v * (1.f + lambda) : v;
  }
  float rhs = tot[NP * (NP + 1) / 2 + i] * scaling;
  float dii = 0.f;
#pragma unroll
  for (int j = 0; j < NP; ++j) dii = (j == i) ? a[j] : dii;
  const bool all_zero = __all_sync(full, lane >= NP || !(dii > 0.f));
  const float sc = dii > 0.f ? 1.0f / sqrtf(dii) : 0.f; // zero row and column: the direction drops out
#pragma unroll
  for (int j = 0; j < NP; ++j) a[j] *= sc * __shfl_sync(full, sc, j);
  rhs *= sc;
  // right-looking Cholesky: after step k, a[k] of lane i >= k holds L[i][k]
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    const float dk = __shfl_sync(full, a[k], k);
    const float inv = dk > 1e-12f ? rsqrtf(dk) : 0.f; // 2 ulp is far below the fp32 conditioning noise of the system
    const float lik = a[k] * inv;
    a[k] = lik;
#pragma unroll
    for (int j = k + 1; j < NP; ++j) {
      const float ljk = __shfl_sync(full, lik, j);
      a[j] = (i >= j) ? fmaf(-lik, ljk, a[j]) : a[j];
    }
  }
  float lii = 1.f;
#pragma unroll
  for (int j = 0; j < NP; ++j) lii = (j == i) ? a[j] : lii;
  const float rinv = lii > 0.f ? __frcp_rn(lii) : 0.f;
  // forward substitution L y = rhs (column oriented)
  float y = rhs;
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    const float yk = __shfl_sync(full, y * rinv, k);
    y = (i == k) ? yk : ((i > k) ? fmaf(-a[k], yk, y) : y);
  }
  // backward substitution L^T x = y needs column access: rows go through shared memory once
  if (lane < NP) {
#pragma unroll
    for (int j = 0; j < NP; ++j) smem[lane * NP + j] = a[j];
  }
  __syncwarp();
  float x = y;
#pragma unroll
  for (int k = NP - 1; k >= 0; --k) {
    const float xk = __shfl_sync(full, x * rinv, k);
    const float lki = smem[k * NP + i]; // L[k][i], used by lanes i < k
    x = (i == k) ? xk : ((i < k) ? fmaf(-lki, xk, x) : x);
  }
  if (lane < NP) dp[lane] = all_zero ? __int_as_float(0x7fc00000) : x * sc;
  __syncwarp();
  return true;
}

// pyramid_class.cpp:260-287: only u, v scale between levels; the quadratic extension scales its
// second-order terms by the inverse factor.
template <int MODEL>
__device__ __forceinline__ float translate_param(float v, int idx, int src, int dst) {
  float mag = dst > src ? 1.f / (float)(1 << (dst - src)) : (float)(1 << (src - dst));
  if (idx < 2) return v * mag;
  if (MODEL == DIC_FM_QUADRATIC && idx >= 6) return v * (1.f / mag);
  return v;
}

} // namespace dic
