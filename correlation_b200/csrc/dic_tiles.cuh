// dic_tiles.cuh -- the structured (tile) form of the fused Gauss-Newton evaluation.
//
// Integer-grid domains (rectangles, annuli, blobs: everything the reference's builders produce,
// manager_class.cpp:1596-1614 / :816-940, polygon_class.cpp) are stored per pyramid level as
// 32 x 32 pixel tiles with one 32-bit membership mask per row, ordered COLUMN-major so that a warp
// walks down one 32-pixel-wide strip: lane <-> x, loop <-> y. That layout buys three things the
// pixel-list kernel cannot have:
//   1. coalesced u8 traffic: a warp-row of the reference image is one 32-byte sector, and the
//      deformed-image footprint of a whole unit (about 40 x 40 pixels) is staged ONCE into shared
//      memory by TMA; a window row is converted once per pixel of the column, not four times;
//   2. dx = x - cx is a per-lane constant, so J^T J / J^T r are accumulated as moments in dy only
//      (sum g g' dy^b, sum V g dy^b): 15 + 6 + 2 registers for the 12-parameter model instead of
//      92, 9 + 4 + 2 instead of 29 for the affine one, and 3-4x fewer FMAs per pixel;
//   3. the in-image test (interpolation_class.cpp:82-83) is hoisted to one test per tile; tiles
//      whose warped footprint leaves the image or the staging buffer take the per-pixel path.
// The moments are expanded to the full upper-triangular A, b with the lane's dx powers only when
// the warp changes strip or the evaluation ends.
#pragma once
#include "dic_kernels.cuh"
#include "dic_f32x2.cuh"

#ifndef DIC_PAIR_LATE_BRANCH
#define DIC_PAIR_LATE_BRANCH 0 // measured: the late branch costs registers and is slower (c2 +9 %, c5 +10 %)
#endif
#ifndef DIC_BATCH_QUEUE
#define DIC_BATCH_QUEUE 1
#endif
#ifndef DIC_FAST_UNROLL_BATCH
#define DIC_FAST_UNROLL_BATCH 1 // fast mode, batch form: pixel steps per loop trip (4: static window rotation; 1: one copy + register moves)
#endif
#ifndef DIC_YSTAGE_PACKED
#define DIC_YSTAGE_PACKED 1
#endif
#ifndef DIC_BATCH_TIMELINE
#define DIC_BATCH_TIMELINE 0 // diagnostics only
#endif

namespace dic {

// Resident CTAs per SM the tile kernel is compiled for. 2 CTAs = 16 warps per SM at 128 registers.
// Measured on B200 (c4 / c5, affine): 3 CTAs at 80 registers (20-170 bytes of spill) is 10 % slower in
// parity mode and within +-5 % in fast mode; the quadratic model spills ~1 KB per thread at 80.
#ifndef DIC_TILE_CTAS_AFFINE
#define DIC_TILE_CTAS_AFFINE 2
#endif
__host__ __device__ constexpr int tile_ctas_per_sm(int model) { return model == DIC_FM_QUADRATIC ? 2 : DIC_TILE_CTAS_AFFINE; }
#ifndef DIC_TILE_H
#define DIC_TILE_H 32
#endif
// 32 x kTileH pixels per tile. 32 rows (16 in round 1): a unit -- the rows of one tile a warp walks between two
// plan / stage / window-open sequences -- is then up to 32 rows, i.e. 16 steps of the two-stream parity loop instead
// of 8; the per-unit work (~330 instructions and one L2 round trip for the tile record) was 15 % of the kernel's
// instructions with 16-row tiles.
constexpr int kTileW = 32, kTileH = DIC_TILE_H;
constexpr int kQuadsPerTile = kTileH / 4;
static_assert(kTileH == 16 || kTileH == 32, "a column of the membership mask is one 32-bit word");
__host__ __device__ constexpr uint32_t low_bits(int n) { return n >= 32 ? 0xffffffffu : ((1u << n) - 1u); }
// Per-warp staging, filled by TMA (cp.async.bulk.tensor.2d) one unit ahead of the arithmetic:
//   the deformed-image footprint of a unit as u8, kPatchW x kPatchH bytes,
//   the reference-image pixels of the unit's tile, kUndW x kTileH bytes.
// The innermost TMA coordinate must be a multiple of 16 bytes (measured on B200: any other value
// faults with "illegal instruction"), so both boxes start at x & ~15 and carry 15 spare columns:
// 64 = 15 + 32 + 3 halo + 1 + 13 for strain / rotation across the unit; 48 = 15 + 32 + 1. Rows: kTileH + 3 halo + 5 (9)
// for strain / rotation.
// Two buffers of each per warp and one mbarrier per buffer.
constexpr int kPatchW = 64, kPatchH = kTileH == 32 ? 44 : 24, kUndW = 48;
constexpr int kPatchBytes = kPatchW * kPatchH, kUndBytes = kUndW * kTileH;
constexpr int kStageBytes = kPatchBytes + kUndBytes;          // one buffer: 2304 B = 18 x 128
constexpr int kWarpStageBytes = 2 * kStageBytes;              // per warp
constexpr int kWarpsPerCta = kThreads / 32;
static_assert(kStageBytes % 128 == 0 && kPatchBytes % 128 == 0, "TMA destinations must be 128-byte aligned");

// Tensor maps of the pyramid levels the launch reads (kernel parameter, __grid_constant__).
struct alignas(64) TileMaps {
  CUtensorMap def[kMaxLevels]; // deformed image, box kPatchW x kPatchH
  CUtensorMap und[kMaxLevels]; // reference image, box kUndW x kTileH
};

// A warp's staging state: buffer k & 1 serves the k-th staged unit, its mbarrier phase is (k >> 1) & 1.
struct WarpStage {
  uint8_t *buf;   // [2][kStageBytes]: patch then und tile
  uint64_t *bar;  // [2]
  uint32_t issued, consumed;
};

struct Tile {
  int x0, y0;             // level coordinates of the tile's first pixel
  uint32_t rows[kTileH];  // bit l of rows[r] <=> pixel (x0 + l, y0 + r) belongs to the domain
  uint32_t cols[kTileW];  // the same membership transposed: bit r of cols[l]
  uint32_t full_rows;     // bit r set <=> rows[r] == 0xffffffff
  uint32_t pad;
};

// fills the derived fields of a tile from rows[]
__device__ __forceinline__ void tile_finish(Tile &t) {
  uint32_t full = 0;
#pragma unroll
  for (int r = 0; r < kTileH; ++r) full |= (t.rows[r] == 0xffffffffu ? 1u : 0u) << r;
  t.full_rows = full;
  t.pad = 0;
  for (int l = 0; l < kTileW; ++l) {
    uint32_t c = 0;
#pragma unroll
    for (int r = 0; r < kTileH; ++r) c |= ((t.rows[r] >> l) & 1u) << r;
    t.cols[l] = c;
  }
}

struct TileLevel {
  const Tile *tiles;
  int n_tiles;
  const float2 *extra; // pixels that occur more than once in the reference list (blob edges)
  int n_extra;
};

struct SectorTiles {
  TileLevel lev[kMaxLevels];
};

// ---- monomial bookkeeping: parameter k <-> (gradient component g, X^a Y^b, coefficient)
// affine:    u v ux uy vx vy                     (model_class.cpp:150-202)
// quadratic: ... uxx uxy uyy vxx vxy vyy         (extension; 1/2 on the pure second-order terms)
template <int NP> struct Mono {
  static __host__ __device__ constexpr int g(int k) { return k < 2 ? k : (k < 6 ? (k - 2) / 2 : (k - 6) / 3); }
  static __host__ __device__ constexpr int a(int k) {
    return k < 2 ? 0 : k < 6 ? ((k - 2) % 2 == 0 ? 1 : 0) : ((k - 6) % 3 == 0 ? 2 : (k - 6) % 3 == 1 ? 1 : 0);
  }
  static __host__ __device__ constexpr int b(int k) {
    return k < 2 ? 0 : k < 6 ? ((k - 2) % 2 == 1 ? 1 : 0) : ((k - 6) % 3 == 2 ? 2 : (k - 6) % 3 == 1 ? 1 : 0);
  }
  static __host__ __device__ constexpr float c(int k) { return (k >= 6 && (k - 6) % 3 != 1) ? 0.5f : 1.f; }
};

template <int NP> struct Mom {
  static constexpr int kDeg = NP == 12 ? 2 : 1;    // degree of the warp in Y
  static constexpr int kNB = 2 * kDeg + 1;          // Y powers in A: 0 .. 2 deg
  static constexpr int kNBV = kDeg + 1;             // Y powers in b
  static constexpr int kGG = 0;                     // [3][kNB]  xx, xy, yy
  static constexpr int kVG = 3 * kNB;               // [2][kNBV]
  static constexpr int kChi = kVG + 2 * kNBV;
  static constexpr int kOob = kChi + 1;
  static constexpr int kN = kOob + 1;
};

// One pixel into the lane's moment accumulators.
template <int NP>
__device__ __forceinline__ void accumulate_moments(float *mom, float V, float wx, float wy, float Y) {
  using M = Mom<NP>;
  float gxx = wx * wx, gxy = wx * wy, gyy = wy * wy, vx = V * wx, vy = V * wy;
  mom[M::kChi] = fmaf(V, V, mom[M::kChi]);
  float yp = 1.f;
#pragma unroll
  for (int b = 0; b < M::kNB; ++b) {
    if (b == 0) {
      mom[M::kGG + 0] += gxx; mom[M::kGG + M::kNB] += gxy; mom[M::kGG + 2 * M::kNB] += gyy;
      mom[M::kVG + 0] += vx; mom[M::kVG + M::kNBV] += vy;
    } else {
      yp = b == 1 ? Y : yp * Y;
      mom[M::kGG + b] = fmaf(gxx, yp, mom[M::kGG + b]);
      mom[M::kGG + M::kNB + b] = fmaf(gxy, yp, mom[M::kGG + M::kNB + b]);
      mom[M::kGG + 2 * M::kNB + b] = fmaf(gyy, yp, mom[M::kGG + 2 * M::kNB + b]);
      if (b < M::kNBV) {
        mom[M::kVG + b] = fmaf(vx, yp, mom[M::kVG + b]);
        mom[M::kVG + M::kNBV + b] = fmaf(vy, yp, mom[M::kVG + M::kNBV + b]);
      }
    }
  }
}

// Warp transpose-reduction of N per-lane values (N padded to a multiple of 32): five butterfly
// stages, each halving the values a lane still carries, N - N/32 shuffles in total instead of 5 N.
// On return lane l holds in v[j] (j < NPAD/32) the warp total of value index (j * 32 + perm(l)),
// where perm(l) reverses nothing: index = j * 32 + l with the bit order produced below.
template <int NPAD>
__device__ __forceinline__ void warp_transpose_sum(float *v) {
  const int lane = threadIdx.x & 31;
  int cnt = NPAD;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const int half = cnt >> 1;
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < NPAD / 2; ++i) {
      if (i < half) {
        const float keep = up ? v[i + half] : v[i];
        const float give = up ? v[i] : v[i + half];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, give, o);
      }
    }
    cnt = half;
  }
}
// value index held by `lane` in slot j after warp_transpose_sum<NPAD>
template <int NPAD>
__device__ __forceinline__ int transpose_index(int lane, int j) {
  // stage o keeps the upper half of the current range when (lane & o): the final index is
  // sum over stages of (bit ? half_size_at_that_stage : 0) + j
  int idx = 0, half = NPAD >> 1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { idx += (lane & o) ? half : 0; half >>= 1; }
  return idx + j;
}

// Expand the lane's moments with its X = dx into the packed upper-triangular A, b, chi, oob,
// warp-reduce, and add into the warp's shared accumulator row. Clears the moments.
template <int NP>
__device__ __forceinline__ void flush_moments(float *mom, float X, float *warp_acc) {
  using M = Mom<NP>;
  using L = Acc<NP>;
  using MO = Mono<NP>;
  constexpr int NPAD = (L::kN + 31) / 32 * 32;
  const int lane = threadIdx.x & 31;
  float xp[5];
  xp[0] = 1.f; xp[1] = X; xp[2] = X * X; xp[3] = xp[2] * X; xp[4] = xp[2] * xp[2];
  float v[NPAD];
  int k = 0;
#pragma unroll
  for (int p1 = 0; p1 < NP; ++p1) {
#pragma unroll
    for (int p2 = p1; p2 < NP; ++p2) {
      const int gg = MO::g(p1) + MO::g(p2); // 0 xx, 1 xy, 2 yy
      const int a = MO::a(p1) + MO::a(p2), b = MO::b(p1) + MO::b(p2);
      v[k] = MO::c(p1) * MO::c(p2) * xp[a] * mom[M::kGG + gg * M::kNB + b];
      ++k;
    }
  }
#pragma unroll
  for (int p = 0; p < NP; ++p)
    v[L::kB + p] = MO::c(p) * xp[MO::a(p)] * mom[M::kVG + MO::g(p) * M::kNBV + MO::b(p)];
  v[L::kChi] = mom[M::kChi];
  v[L::kOob] = mom[M::kOob];
#pragma unroll
  for (int i = L::kN; i < NPAD; ++i) v[i] = 0.f;
  warp_transpose_sum<NPAD>(v);
#pragma unroll
  for (int j = 0; j < NPAD / 32; ++j) {
    const int idx = transpose_index<NPAD>(lane, j);
    if (idx < L::kN) warp_acc[idx] += v[j];
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < M::kN; ++i) mom[i] = 0.f;
}

__device__ __forceinline__ float floor_magic(float x, int &i) {
  // floor for 0 <= x < 2^22 without the conversion pipe: round-down add of 2^23
  float m = __fadd_rd(x, 8388608.0f);
  i = __float_as_int(m) & 0x7fffff;
  return m - 8388608.0f;
}

// Deformed position of pixel (X = x - cx, Y = y - cy): per-lane constants in X, polynomial in Y.
// Parity mode evaluates the reference's left-to-right fp32 expression instead (warp_point).
template <int NP> struct LaneWarp {
  float cx0, cx1, cx2, cy0, cy1, cy2;
  __device__ __forceinline__ void set(const float *p, float x, float X) {
    // x' = x + u + ux X + uy Y (+ 1/2 uxx X^2 + uxy X Y + 1/2 uyy Y^2)
    cx0 = x + p[0] + p[2] * X; cx1 = p[3]; cx2 = 0.f;
    cy0 = p[1] + p[4] * X; cy1 = p[5]; cy2 = 0.f;       // y' = y + v + vx X + vy Y (+ ...)
    if (NP == 12) {
      cx0 += 0.5f * p[6] * X * X; cx1 += p[7] * X; cx2 = 0.5f * p[8];
      cy0 += 0.5f * p[9] * X * X; cy1 += p[10] * X; cy2 = 0.5f * p[11];
    }
  }
};

// x-direction cubic of one window row from its four u8 pixels p0..p3 (columns ix-1 .. ix+2), packed in
// `win`. PRMT builds 2^23 + p (exact); the bias cancels in every difference, so one FADD removes it
// where the value itself is needed.
//  PARITY: the reference's monomial coefficients in s = 1 + t (the row's part of a = (B (x) B) v,
//          interpolation_class.cpp:296-336); every intermediate is a small multiple of 1/2: exact.
//  FAST:   Catmull-Rom coefficients in t.
template <int MODE>
__device__ __forceinline__ void row_coeffs_u8(uint32_t win, float c[4]) {
  const float f0 = __uint_as_float(__byte_perm(win, 0x4B000000u, 0x7650u));
  const float f1 = __uint_as_float(__byte_perm(win, 0x4B000000u, 0x7651u));
  const float f2 = __uint_as_float(__byte_perm(win, 0x4B000000u, 0x7652u));
  const float f3 = __uint_as_float(__byte_perm(win, 0x4B000000u, 0x7653u));
  if (MODE == DIC_MODE_PARITY) {
    // 12 exact operations: c0 = p0 - 2 c3 and c2 = c3 + d - c1 (value and slope of the cubic at s = 1)
    const float a = f1 - f2, b = f3 - f0, d = f1 - f0, p0 = f0 - 8388608.0f;
    c[3] = fmaf(1.5f, a, 0.5f * b);
    c[0] = fmaf(-2.f, c[3], p0);
    c[1] = fmaf(1.5f, d, fmaf(8.f, a, 2.5f * b));
    c[2] = (c[3] + d) - c[1];
  } else {
    const float d0 = f0 - f1, d2 = f2 - f1, d3 = f3 - f1;
    c[0] = f1 - 8388608.0f;
    c[1] = 0.5f * (d2 - d0);
    c[2] = fmaf(-0.5f, d3, fmaf(2.f, d2, d0));
    c[3] = fmaf(-1.5f, d2, 0.5f * (d3 - d0));
  }
}

// four consecutive bytes at byte offset `o` of the warp's patch (any alignment)
__device__ __forceinline__ uint32_t patch_window(const uint8_t *patch, int o) {
  const uint32_t *w = reinterpret_cast<const uint32_t *>(patch + (o & ~3));
  return __funnelshift_r(w[0], w[1], (o & 3) * 8);
}

// One staged pixel. The lane walks down a column, so consecutive pixels normally share three of
// their four window rows: the x-direction cubic coefficients of the window rows are kept in
// registers (cw, slot-rotated by the static step S) and only the new bottom row is read from the
// staged patch (two LDS + funnel shift) and converted. When any lane's window moved differently
// (ix changed, iy did not advance by exactly one, first row of a unit) the whole warp rebuilds its
// four rows: the branch is warp-uniform. `member` masks pixels outside the domain (FULL: every pixel
// of the unit is a member).
template <int MODEL, int MODE, bool FULL, int S>
__device__ __forceinline__ void staged_pixel(const float *pw, const LaneWarp<model_nparams(MODEL)> &lw,
                                             float xf, float yf, float ccx, float ccy, const uint8_t *patch,
                                             int px0, int py0, float und_w, bool member, float (&cw)[4][4],
                                             int &wix, int &wiy, float *mom) {
  constexpr int NP = model_nparams(MODEL);
  const float Y = __fsub_rn(yf, ccy);
  float xd, yd;
  if (MODE == DIC_MODE_PARITY) {
    float dxx, dyy;
    warp_point<MODEL, MODE>(pw, xf, yf, ccx, ccy, xd, yd, dxx, dyy);
  } else {
    xd = fmaf(fmaf(lw.cx2, Y, lw.cx1), Y, lw.cx0);
    yd = fmaf(fmaf(lw.cy2, Y, lw.cy1), Y, lw.cy0) + yf;
  }
  int ix, iy;
  const float fx = floor_magic(xd, ix), fy = floor_magic(yd, iy);
  const int o = (iy - 1 - py0) * kPatchW + (ix - 1 - px0); // byte offset of the window's first row
  const bool slide = ix == wix && iy == wiy + 1;
  if (__all_sync(0xffffffffu, slide)) {
    row_coeffs_u8<MODE>(patch_window(patch, o + 3 * kPatchW), cw[(3 + S) & 3]);
  } else {
    row_coeffs_u8<MODE>(patch_window(patch, o), cw[(0 + S) & 3]);
    row_coeffs_u8<MODE>(patch_window(patch, o + kPatchW), cw[(1 + S) & 3]);
    row_coeffs_u8<MODE>(patch_window(patch, o + 2 * kPatchW), cw[(2 + S) & 3]);
    row_coeffs_u8<MODE>(patch_window(patch, o + 3 * kPatchW), cw[(3 + S) & 3]);
  }
  wix = ix; wiy = iy;
  float w, wx, wy;
  if (MODE == DIC_MODE_PARITY)
    bicubic_parity_rows(cw[(0 + S) & 3], cw[(1 + S) & 3], cw[(2 + S) & 3], cw[(3 + S) & 3], xd, yd, ix, iy, w, wx, wy);
  else
    bicubic_fast_rows(cw[(0 + S) & 3], cw[(1 + S) & 3], cw[(2 + S) & 3], cw[(3 + S) & 3], xd - fx, yd - fy, w, wx, wy);
  float V = und_w - w;
  if (!FULL) { V = member ? V : 0.f; wx = member ? wx : 0.f; wy = member ? wy : 0.f; }
  accumulate_moments<NP>(mom, V, wx, wy, Y);
}

// ---- parity mode, second form of the inner loop (DIC_PARITY_LOOP == 2, the default).
// The window state of a lane: cw = x-direction cubic coefficients of the four window rows in ROTATING slots,
// (wix, wiy) = the window's pixel as the bit patterns the floor trick produces (2^23 + ix, and the 2^23 + iy expected
// at the next step), wp = address of the aligned 32-bit word that holds the first byte of the window's LAST row in
// the staged patch, wsh = bit shift of that byte inside the word (the same for every row of a window, because the
// patch pitch is a multiple of four).
//   parity_window_open          (cold: first row of a unit, or a lane's window did not simply move down one row)
//                               positions the window ONE ROW ABOVE the pixel: rows 0..2 of the pixel's window go to
//                               slots 0..2, so that the very next step -- at the same pixel -- loads row 3 and evaluates;
//   parity_slide_step_trip<S>   (hot) one pixel: warp, floor, "did every lane's window move down by exactly one row?";
//                               if not, returns false without side effects (the caller reopens the window at this
//                               pixel). Else one row is read (immediate offset from wp: no address arithmetic from ix,
//                               iy), converted, and the pixel is evaluated on slots (S, S+1, S+2, S+3) mod 4. Four steps
//                               with S = 0, 1, 2, 3 in a row rotate the slots back: the window never moves between
//                               registers (the first form shifted 12 registers per pixel and paid ~7 more copies to
//                               merge its two branches), and the loop-carried values advance once per four pixels.
// The per-pixel results are bit-identical to the first form: same operations on the same values.
// Packed y pass (DIC_YSTAGE_PACKED, the default): the window rows hold their four x-direction coefficients as TWO
// fp32 pairs (c0, c1), (c2, c3), and the y pass of the coefficient stage -- the same 11 exact operations for each of
// the four columns -- runs as 2 x 11 packed instructions (dic_f32x2.cuh) instead of 4 x 11 scalar ones: 22 issue
// slots less per pixel in a loop that is bound by issue slots, the same FMA-pipe passes. Every value of this stage
// is exact in fp32 (multiples of 1/4 below 2^17), so neither the packing nor any contraction ptxas applies to it can
// change a bit; the 40-term polynomial stays scalar and unfused.
__device__ __forceinline__ void row_coeffs_u8_pairs(uint32_t win, f2 (&row)[2]) {
  float c[4];
  row_coeffs_u8<DIC_MODE_PARITY>(win, c);
  row[0] = pk(c[0], c[1]); row[1] = pk(c[2], c[3]);
}
__device__ __forceinline__ void bicubic_parity_rows_pairs(const f2 (&r0)[2], const f2 (&r1)[2], const f2 (&r2)[2],
                                                          const f2 (&r3)[2], float xdef, float ydef, float fix, float fiy,
                                                          float &w, float &wx, float &wy) {
  float a[4][4];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    f2 c[4];
    monomial_from_rows_x2(r0[h], r1[h], r2[h], r3[h], c);
#pragma unroll
    for (int jk = 0; jk < 4; ++jk) { a[jk][2 * h] = lo(c[jk]); a[jk][2 * h + 1] = hi(c[jk]); }
  }
  parity_eval_f(a, xdef, ydef, fix, fiy, w, wx, wy);
}

template <int MODEL>
__device__ __forceinline__ void parity_window_open_pairs(const float *pw, float xf, float yf, float ccx, float ccy,
                                                         const uint8_t *patch, int px0, int py0, f2 (&cw)[4][2],
                                                         int &wix, int &wiy, const uint8_t *&wp, int &wsh) {
  float xd, yd, dxx, dyy;
  warp_point<MODEL, DIC_MODE_PARITY>(pw, xf, yf, ccx, ccy, xd, yd, dxx, dyy);
  int ix, iy;
  floor_magic(xd, ix);
  floor_magic(yd, iy);
  const int o = (iy - 1 - py0) * kPatchW + (ix - 1 - px0); // byte offset of the window's first row
  wsh = (o & 3) * 8;
  const uint8_t *b = patch + (o & ~3);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(b + k * kPatchW);
    row_coeffs_u8_pairs(__funnelshift_r(w[0], w[1], wsh), cw[k]);
  }
  wp = b + 2 * kPatchW;
  wix = __float_as_int(__fadd_rd(xd, 8388608.0f)); wiy = __float_as_int(__fadd_rd(yd, 8388608.0f));
}

template <int MODEL, int S>
__device__ __forceinline__ bool parity_slide_step_trip_pairs(const float *pw, float xf, float yf, float ccx, float ccy,
                                                             f2 (&cw)[4][2], int wix_bits, int &wiy_next_bits,
                                                             const uint8_t *wp, int wsh, float und_w, bool member, float *mom) {
  constexpr int NP = model_nparams(MODEL);
  float xd, yd, dxx, Y;
  warp_point<MODEL, DIC_MODE_PARITY>(pw, xf, S == 0 ? yf : yf + (float)S, ccx, ccy, xd, yd, dxx, Y);
  const float mx = __fadd_rd(xd, 8388608.0f), my = __fadd_rd(yd, 8388608.0f);
  if (!__all_sync(0xffffffffu, __float_as_int(mx) == wix_bits && __float_as_int(my) == wiy_next_bits)) return false;
  wiy_next_bits = __float_as_int(my) + 1;
  const float fx = mx - 8388608.0f, fy = my - 8388608.0f;
  {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(wp + (S + 1) * kPatchW);
    row_coeffs_u8_pairs(__funnelshift_r(w[0], w[1], wsh), cw[(3 + S) & 3]);
  }
  float w, wx, wy;
  bicubic_parity_rows_pairs(cw[(0 + S) & 3], cw[(1 + S) & 3], cw[(2 + S) & 3], cw[(3 + S) & 3], xd, yd, fx, fy, w, wx, wy);
  float V = und_w - w;
  V = member ? V : 0.f; wx = member ? wx : 0.f; wy = member ? wy : 0.f;
  accumulate_moments<NP>(mom, V, wx, wy, Y);
  return true;
}

template <int MODEL>
__device__ __forceinline__ void parity_window_open(const float *pw, float xf, float yf, float ccx, float ccy,
                                                   const uint8_t *patch, int px0, int py0, float (&cw)[4][4],
                                                   int &wix, int &wiy, const uint8_t *&wp, int &wsh) {
  float xd, yd, dxx, dyy;
  warp_point<MODEL, DIC_MODE_PARITY>(pw, xf, yf, ccx, ccy, xd, yd, dxx, dyy);
  int ix, iy;
  floor_magic(xd, ix);
  floor_magic(yd, iy);
  const int o = (iy - 1 - py0) * kPatchW + (ix - 1 - px0); // byte offset of the window's first row
  wsh = (o & 3) * 8;
  const uint8_t *b = patch + (o & ~3);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(b + k * kPatchW);
    row_coeffs_u8<DIC_MODE_PARITY>(__funnelshift_r(w[0], w[1], wsh), cw[k]);
  }
  wp = b + 2 * kPatchW;
  // the window sits one row above the pixel: the next step is at this same pixel (see parity_slide_step_trip)
  wix = __float_as_int(__fadd_rd(xd, 8388608.0f)); wiy = __float_as_int(__fadd_rd(yd, 8388608.0f));
}

// Trip-static form of the step (what the loop below uses): four steps S = 0..3 share one set of loop-carried values
// (yf, wp, the reference-pixel pointer, the membership word), each step addresses its row with an immediate offset,
// and the window's pixel is kept as the BIT PATTERNS of 2^23 + ix and of the 2^23 + iy expected next, so that the
// slide test is two integer compares on the floor trick's own output (no mask, no + 1 on the critical path).
// yf is the row of step 0 of the trip; wp the word holding the first byte of the window's last row before step 0.
template <int MODEL, int S>
__device__ __forceinline__ bool parity_slide_step_trip(const float *pw, float xf, float yf, float ccx, float ccy,
                                                       float (&cw)[4][4], int wix_bits, int &wiy_next_bits,
                                                       const uint8_t *wp, int wsh, float und_w, bool member, float *mom) {
  constexpr int NP = model_nparams(MODEL);
  float xd, yd, dxx, Y;
  warp_point<MODEL, DIC_MODE_PARITY>(pw, xf, S == 0 ? yf : yf + (float)S, ccx, ccy, xd, yd, dxx, Y);
  const float mx = __fadd_rd(xd, 8388608.0f), my = __fadd_rd(yd, 8388608.0f); // floor_magic, integer part in the low bits
  if (!__all_sync(0xffffffffu, __float_as_int(mx) == wix_bits && __float_as_int(my) == wiy_next_bits)) return false;
  wiy_next_bits = __float_as_int(my) + 1;
  const float fx = mx - 8388608.0f, fy = my - 8388608.0f;
  {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(wp + (S + 1) * kPatchW);
    row_coeffs_u8<DIC_MODE_PARITY>(__funnelshift_r(w[0], w[1], wsh), cw[(3 + S) & 3]);
  }
  float w, wx, wy;
  bicubic_parity_rows_f(cw[(0 + S) & 3], cw[(1 + S) & 3], cw[(2 + S) & 3], cw[(3 + S) & 3], xd, yd, fx, fy, w, wx, wy);
  float V = und_w - w;
  V = member ? V : 0.f; wx = member ? wx : 0.f; wy = member ? wy : 0.f;
  accumulate_moments<NP>(mom, V, wx, wy, Y);
  return true;
}

// ---- parity mode, third form of the inner loop (DIC_PARITY_LOOP == 3, an option): TWO pixels per step in the two
// halves of packed fp32 pairs (dic_f32x2.cuh). A unit of nr rows (nr even) is walked as two streams of nr / 2 rows,
// stream A = rows [0, nr/2) in the lo halves, stream B = rows [nr/2, nr) in the hi halves; each stream has its own
// sliding window (rotating slots, as in the second form), and one step
//   warps both pixels, floors, tests "did every lane's two windows move down by exactly one row?", reads one new
//   window row per stream, converts both with ONE set of packed operations, runs the y pass and the 40-term
//   polynomial once on pairs, and adds both pixels to the lane's moments.
// Per pixel the operations and their operands are those of the second form, so every w, dw/dx, dw/dy is bit-identical
// (once the products that feed additions are protected from ptxas's contraction, see mul2_sep); the issue slots per
// pixel fall from ~263 to ~166 and the loop becomes bound by the FMA pipe instead -- which is why it does not win.
template <int MODEL> struct PairWarp {
  // the reference's left-to-right warp expression (model_class.cpp:150-202; warp_point<MODEL, PARITY>) with the
  // prefix that does not depend on y evaluated once per strip
  float kx, p3, q4, p1, p5, cy;
  float qx, p7x, h8, qy, p10x, h11; // quadratic extension
  __device__ __forceinline__ void set(const float *p, float xf, float ccx, float ccy) {
    const float dx = __fsub_rn(xf, ccx);
    kx = __fadd_rn(__fadd_rn(xf, p[0]), __fmul_rn(p[2], dx));
    q4 = __fmul_rn(p[4], dx);
    p1 = p[1]; p3 = p[3]; p5 = p[5]; cy = ccy;
    if (MODEL == DIC_FM_QUADRATIC) {
      qx = __fmul_rn(__fmul_rn(__fmul_rn(0.5f, p[6]), dx), dx); p7x = __fmul_rn(p[7], dx); h8 = __fmul_rn(0.5f, p[8]);
      qy = __fmul_rn(__fmul_rn(__fmul_rn(0.5f, p[9]), dx), dx); p10x = __fmul_rn(p[10], dx); h11 = __fmul_rn(0.5f, p[11]);
    } else {
      qx = p7x = h8 = qy = p10x = h11 = 0.f;
    }
  }
  // zero: see mul2_sep -- every product here is rounded before it is added, as in the reference's expression
  __device__ __forceinline__ void point(f2 Y, uint32_t zero, f2 &XD, f2 &YD, f2 &DY) const {
    static_assert(MODEL == DIC_FM_UVUxUyVxVy || MODEL == DIC_FM_QUADRATIC, "tile kernel models");
    DY = sub2(Y, bc(cy));
    f2 tx = add2(bc(kx), mul2_sep(bc(p3), DY, zero));
    f2 ty = add2(add2(add2(Y, bc(p1)), bc(q4)), mul2_sep(bc(p5), DY, zero));
    if (MODEL == DIC_FM_QUADRATIC) {
      tx = add2(tx, bc(qx)); tx = add2(tx, mul2_sep(bc(p7x), DY, zero)); tx = add2(tx, mul2_sep(mul2(bc(h8), DY), DY, zero));
      ty = add2(ty, bc(qy)); ty = add2(ty, mul2_sep(bc(p10x), DY, zero)); ty = add2(ty, mul2_sep(mul2(bc(h11), DY), DY, zero));
    }
    XD = tx; YD = ty;
  }
};

// window state of the two streams
struct PairWindow {
  uint32_t mx_a, mx_b;     // bit patterns of 2^23 + ix of the windows' pixels (floor_magic)
  uint32_t my_a, my_b;     // bit patterns of 2^23 + iy EXPECTED at the next step
  const uint8_t *wp_a, *wp_b; // aligned word holding the first byte of the windows' last loaded row, at the trip's start
  int sh_a, sh_b;          // bit shift of that byte inside the word
};

template <int MODEL>
__device__ __forceinline__ void parity_pair_open(const PairWarp<MODEL> &pwp, f2 Y, uint32_t zero, const uint8_t *patch,
                                                 int px0, int py0, f2 (&cw)[4][4], PairWindow &win) {
  f2 XD, YD, DY;
  pwp.point(Y, zero, XD, YD, DY);
  const f2 MX = add2_rd(XD, bc(8388608.0f)), MY = add2_rd(YD, bc(8388608.0f));
  const int ixa = lo_bits(MX) & 0x7fffff, ixb = hi_bits(MX) & 0x7fffff;
  const int iya = lo_bits(MY) & 0x7fffff, iyb = hi_bits(MY) & 0x7fffff;
  const int oa = (iya - 1 - py0) * kPatchW + (ixa - 1 - px0), ob = (iyb - 1 - py0) * kPatchW + (ixb - 1 - px0);
  win.sh_a = (oa & 3) * 8; win.sh_b = (ob & 3) * 8;
  const uint8_t *ba = patch + (oa & ~3), *bb = patch + (ob & ~3);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const uint32_t *wa = reinterpret_cast<const uint32_t *>(ba + k * kPatchW);
    const uint32_t *wb = reinterpret_cast<const uint32_t *>(bb + k * kPatchW);
    row_coeffs_u8_x2(__funnelshift_r(wa[0], wa[1], win.sh_a), __funnelshift_r(wb[0], wb[1], win.sh_b), cw[k]);
  }
  win.wp_a = ba + 2 * kPatchW; win.wp_b = bb + 2 * kPatchW;
  win.mx_a = lo_bits(MX); win.mx_b = hi_bits(MX);
  win.my_a = lo_bits(MY); win.my_b = hi_bits(MY); // the window sits one row above: the next step is at this same pixel
}

// U = reference pixels of the two rows, M = 1 / 0 membership of the two pixels
template <int MODEL, int S>
__device__ __forceinline__ bool parity_pair_step(const PairWarp<MODEL> &pwp, f2 Y, uint32_t zero, f2 (&cw)[4][4],
                                                 PairWindow &win, f2 U, f2 M, float *mom) {
  constexpr int NP = model_nparams(MODEL);
  f2 XD, YD, DY;
  pwp.point(Y, zero, XD, YD, DY);
  const f2 MX = add2_rd(XD, bc(8388608.0f)), MY = add2_rd(YD, bc(8388608.0f));
  const bool slide = lo_bits(MX) == win.mx_a && hi_bits(MX) == win.mx_b && lo_bits(MY) == win.my_a && hi_bits(MY) == win.my_b;
  // The vote is taken here but (DIC_PAIR_LATE_BRANCH) acted upon only before the moments are touched: everything in
  // between is register arithmetic on a scratch row, so a failed step has no side effect and the ~60 cycles of
  // compare -> vote -> branch latency overlap the polynomial instead of preceding it.
  const bool ok = __all_sync(0xffffffffu, slide);
#if !DIC_PAIR_LATE_BRANCH
  if (!ok) return false;
#endif
  {
    const uint32_t *wa = reinterpret_cast<const uint32_t *>(win.wp_a + (S + 1) * kPatchW);
    const uint32_t *wb = reinterpret_cast<const uint32_t *>(win.wp_b + (S + 1) * kPatchW);
    row_coeffs_u8_x2(__funnelshift_r(wa[0], wa[1], win.sh_a), __funnelshift_r(wb[0], wb[1], win.sh_b), cw[(3 + S) & 3]);
  }
  f2 a[4][4];
#pragma unroll
  for (int ik = 0; ik < 4; ++ik) {
    f2 c[4];
    monomial_from_rows_x2(cw[(0 + S) & 3][ik], cw[(1 + S) & 3][ik], cw[(2 + S) & 3][ik], cw[(3 + S) & 3][ik], c);
#pragma unroll
    for (int jk = 0; jk < 4; ++jk) a[jk][ik] = c[jk];
  }
  const f2 FX = sub2(MX, bc(8388608.0f)), FY = sub2(MY, bc(8388608.0f));
  const f2 dxf = add2(sub2(XD, FX), bc(1.f)), dyf = add2(sub2(YD, FY), bc(1.f));
  f2 W, WX, WY;
  parity_eval_x2(a, dxf, dyf, zero, W, WX, WY);
  const f2 V = mul2(sub2(U, W), M);
  WX = mul2(WX, M); WY = mul2(WY, M);
#if DIC_PAIR_LATE_BRANCH
  if (!ok) return false;
#endif
  win.my_a = lo_bits(MY) + 1u; win.my_b = hi_bits(MY) + 1u;
  accumulate_moments<NP>(mom, lo(V), lo(WX), lo(WY), lo(DY));
  accumulate_moments<NP>(mom, hi(V), hi(WX), hi(WY), hi(DY));
  return true;
}

// Default: the second form. Measured on B200 (c4, 4096 subsets, 32-row tiles, ticket queue): second form 2.20 ms,
// third form 2.42 ms. The packed form halves the issue slots (166 instead of 263 per pixel) but not the FMA-pipe
// passes (206 instead of 210 per pixel: a packed instruction is two passes), the pipe is what bounds the unfused
// reference arithmetic, and the packed form pays 33 XORs per step to keep ptxas from contracting its products
// (dic_f32x2.cuh) plus spills at the 128-register budget. Kept as a compile-time option (-DDIC_PARITY_LOOP=3).
#ifndef DIC_PARITY_LOOP
#define DIC_PARITY_LOOP 2
#endif
#ifndef DIC_PAIR_UNROLL
#define DIC_PAIR_UNROLL 1
#endif

// What a warp needs to know about a work unit before touching its pixels. A unit is a run of consecutive
// rows of one tile: rows [r0, r0 + nr) after trimming the rows no lane owns.
struct UnitPlan {
  int x0, y0;       // level coordinates of the unit's first pixel
  int nr;           // rows of the unit (0: nothing to do)
  uint32_t colmask; // this lane's column of the membership mask: bit r <=> pixel (x0 + lane, y0 + r)
  int px0, py0;     // origin of the staged deformed-image patch
  int pxr, pxe, pye; // raw (un-aligned) first column, exclusive end column / row of the footprint the patch must hold
  bool full, staged;
};

// Rows [r0, r1) of tile `t`. GRAN = 4 keeps the trimmed range on multiples of four rows (fast mode walks four
// pixels per trip), GRAN = 1 trims to the exact first / last owned row.
// What plan_unit reads of a tile record, fetched ONE UNIT EARLIER than it is needed: the record comes from L2
// (~0.4 us), and a warp that loads it at planning time stalls for that long once per unit (5 % of a 32-row unit).
struct TileHead {
  int x0, y0;
  uint32_t col, full_rows; // col: this lane's column of the membership mask
};
__device__ __forceinline__ TileHead load_tile_head(const TileLevel &tl, int t) {
  const Tile *tp = tl.tiles + t;
  TileHead h;
  h.col = __ldg(&tp->cols[threadIdx.x & 31]);
  h.x0 = __ldg(&tp->x0); h.y0 = __ldg(&tp->y0); h.full_rows = __ldg(&tp->full_rows);
  return h;
}

template <int NP, int GRAN>
__device__ __forceinline__ UnitPlan plan_unit(const TileHead &th, int r0, int r1, const float *p,
                                              float ccx, float ccy, const LevelImage &def) {
  const int lane = threadIdx.x & 31;
  UnitPlan q;
  uint32_t col = (th.col >> r0) & low_bits(r1 - r0);
  const uint32_t any = __reduce_or_sync(0xffffffffu, col); // rows of the range some lane owns
  if (any == 0u) { q.nr = 0; q.x0 = q.y0 = q.px0 = q.py0 = q.pxr = q.pxe = q.pye = 0; q.colmask = 0; q.full = q.staged = false; return q; }
  int lo = __ffs(any) - 1, hi = 32 - __clz(any);
  if (GRAN > 1) { lo &= ~(GRAN - 1); hi = (hi + GRAN - 1) & ~(GRAN - 1); }
  r0 += lo;
  q.nr = hi - lo;
  q.colmask = (col >> lo) & low_bits(q.nr);
  q.x0 = th.x0; q.y0 = th.y0 + r0;
  const uint32_t unit_rows = low_bits(q.nr) << r0;
  q.full = (th.full_rows & unit_rows) == unit_rows;
  // footprint of the unit under the current parameters: one corner per lane (lanes 0-3), min / max by
  // shuffle, widened for the curvature of the quadratic model
  float bx0, bx1, by0, by1;
  {
    const float Xc = ((lane & 1) ? (float)(q.x0 + kTileW - 1) : (float)q.x0) - ccx;
    const float Yc = ((lane & 2) ? (float)(q.y0 + q.nr - 1) : (float)q.y0) - ccy;
    float xd = Xc + ccx + p[0] + p[2] * Xc + p[3] * Yc;
    float yd = Yc + ccy + p[1] + p[4] * Xc + p[5] * Yc;
    if (NP == 12) {
      xd += 0.5f * p[6] * Xc * Xc + p[7] * Xc * Yc + 0.5f * p[8] * Yc * Yc;
      yd += 0.5f * p[9] * Xc * Xc + p[10] * Xc * Yc + 0.5f * p[11] * Yc * Yc;
    }
    bx0 = bx1 = xd; by0 = by1 = yd;
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
      bx0 = fminf(bx0, __shfl_xor_sync(0xffffffffu, bx0, o)); bx1 = fmaxf(bx1, __shfl_xor_sync(0xffffffffu, bx1, o));
      by0 = fminf(by0, __shfl_xor_sync(0xffffffffu, by0, o)); by1 = fmaxf(by1, __shfl_xor_sync(0xffffffffu, by1, o));
    }
    bx0 = __shfl_sync(0xffffffffu, bx0, 0); bx1 = __shfl_sync(0xffffffffu, bx1, 0);
    by0 = __shfl_sync(0xffffffffu, by0, 0); by1 = __shfl_sync(0xffffffffu, by1, 0);
    float slack = 0.01f;
    if (NP == 12)
      slack += 128.f * (fabsf(p[6]) + fabsf(p[9])) + 32.f * (fabsf(p[8]) + fabsf(p[11]));
    bx0 -= slack; by0 -= slack; bx1 += slack; by1 += slack;
  }
  // staged window: columns [px0, px0 + kPatchW), rows [py0, py0 + kPatchH) must hold every 4 x 4 window
  const bool inside = bx0 > 1.f && by0 > 1.f && bx1 < (float)def.cols - 2.f && by1 < (float)def.rows - 2.f;
  q.pxr = (int)floorf(bx0) - 1;
  q.px0 = q.pxr & ~15; q.py0 = (int)floorf(by0) - 1;
  q.pxe = (int)floorf(bx1) + 3; q.pye = (int)floorf(by1) + 3; // exclusive
  q.staged = inside && q.pxe - q.px0 <= kPatchW && q.pye - q.py0 <= kPatchH;
  return q;
}

// lane 0 starts the two bulk tensor copies of a planned unit into the warp's next staging buffer
__device__ __forceinline__ void issue_unit(WarpStage &st, const UnitPlan &q, const CUtensorMap *map_def,
                                           const CUtensorMap *map_und) {
  if (q.nr > 0 && q.staged) {
    if ((threadIdx.x & 31) == 0) {
      uint8_t *dst = st.buf + (st.issued & 1) * kStageBytes;
      uint64_t *bar = st.bar + (st.issued & 1);
      mbar_expect_tx(bar, kStageBytes);
      tma_load_2d(dst, map_def, q.px0, q.py0, bar);
      tma_load_2d(dst + kPatchBytes, map_und, q.x0 & ~15, q.y0, bar);
    }
    ++st.issued;
  }
}

// ---- speculative staging of the NEXT evaluation's first unit (DIC_SPECULATE, OFF: measured, no gain).
// An evaluation starts with a chain nothing can overlap: tile record from L2 (~0.4 us) -> plan -> TMA of the first unit
// (~1 us) -> first pixel. It is paid 10-18 times per solve, and the small domains (c1, c3) and the coarse levels of
// every domain spend a third of an evaluation in it. Most evaluations follow one at the SAME level with parameters that
// moved by a fraction of a pixel, so at the end of a pass every warp stages the first unit of its range again, planned
// with the parameters it has (box moved up by up to two rows / left by one 16-byte step where the footprint leaves
// room). The next pass plans with the real parameters; if its footprint lies inside the box that is already in shared
// memory it starts computing at once, otherwise (level change, large step, next sector) the copy is awaited and
// dropped. The pixels read are the same either way: results do not change by a bit (GPU suite green with it on).
// Measured on B200 with it on / off: c1 0.106 / 0.101 ms, c2 0.527 / 0.515, c3 0.137 / 0.134, c4 2.231 / 2.176, c5 4.98 / 5.06:
// the chain it removes is shorter than it looked (the level data is L2-resident: a TMA lands in ~0.5 us) and the second
// plan_unit plus the slot traffic sit on the same critical path, before the barrier. Kept as a compile-time option.
#ifndef DIC_SPECULATE
#define DIC_SPECULATE 0
#endif
struct SpecSlot {           // one per warp, in shared memory (keeps the hot loop's registers free)
  int pending;              // a speculative copy was issued and has not been consumed or dropped
  int level, quad_begin;    // what it was planned for
  const Tile *tiles;        // identity of the sector's tile list at that level
  int px0, py0;             // origin of the staged box
  int x0, y0;               // TileHead of the first tile (col per lane below)
  uint32_t full_rows;
  uint32_t col[32];
};
__device__ __forceinline__ void spec_drop(WarpStage &st, SpecSlot *sp, int *timeout_flag) {
  // warp-uniform: sp->pending is written by lane 0 and read after a __syncwarp
  if (sp->pending) {
    const bool landed = mbar_wait(st.bar + (st.consumed & 1), (st.consumed >> 1) & 1);
    ++st.consumed;
    if (!landed && (threadIdx.x & 31) == 0) atomicExch(timeout_flag, 1);
    __syncwarp();
    if ((threadIdx.x & 31) == 0) sp->pending = 0;
    __syncwarp();
  }
}

// ---- one evaluation over a range of QUADS (4 consecutive rows of a tile; 4 quads per tile, tiles in column-major
// strip order) of one level. The partition of a level over warps is in quads, so that every warp gets the same
// number of pixel rows to within four -- a whole-tile granule left some warps with 8 tiles and others with 7
// (12 % of a pass spent waiting at the barrier); coarse levels and small subsets spread the same way until every
// warp has at least one quad. A warp walks its range tile by tile (first and last tile possibly partial); the
// staging of the next unit is in flight while the current one is evaluated.
template <int MODEL, int MODE, bool BATCH>
__device__ __forceinline__ void evaluate_tiles(const SolveSettings &cfg, const TileMaps &maps, float cx0, float cy0,
                                               const TileLevel tl, int level, const float *p,
                                               int quad_begin, int quad_end, WarpStage &st, SpecSlot *spec,
                                               float *warp_acc, unsigned int *slow_counter, int *timeout_flag) {
  constexpr int NP = model_nparams(MODEL);
  constexpr int GRAN = MODE == DIC_MODE_PARITY ? (DIC_PARITY_LOOP == 3 ? 2 : 1) : 4;
  using M = Mom<NP>;
  const int lane = threadIdx.x & 31;
  if (__shfl_sync(0xffffffffu, *(volatile int *)timeout_flag, 0)) return; // a copy was lost earlier in this launch: the staging state is void (warp-uniform test)
  const LevelImage und = cfg.und[level];
  const LevelImage def = cfg.def[level];
  const CUtensorMap *map_def = &maps.def[level], *map_und = &maps.und[level];
  const float inv = 1.f / (float)(1 << level);
  const float ccx = cx0 * inv, ccy = cy0 * inv;
  float mom[M::kN];
#pragma unroll
  for (int i = 0; i < M::kN; ++i) mom[i] = 0.f;
  // parity mode replays the reference's warp expression per pixel: keep the parameters in
  // registers there; fast mode only needs them at unit set-up (shared memory is fine)
  float preg[MODE == DIC_MODE_PARITY ? NP : 1];
  if (MODE == DIC_MODE_PARITY) {
#pragma unroll
    for (int i = 0; i < NP; ++i) preg[i] = p[i];
  }
  const float *pw = MODE == DIC_MODE_PARITY ? preg : p;

  int cur_x0 = INT_MIN;
  float X = 0.f, xf = 0.f;
  LaneWarp<NP> lw;
  lw.set(p, 0.f, 0.f);
  PairWarp<MODEL> pwp;
  pwp.set(p, 0.f, ccx, ccy);

  // rows [r0, r1) of tile t that belong to this warp's quad range
  const int t_first = quad_begin / kQuadsPerTile, t_last = (quad_end - 1) / kQuadsPerTile;
  auto rows_of = [&](int t, int &r0, int &r1) {
    r0 = t == t_first ? (quad_begin % kQuadsPerTile) * 4 : 0;
    r1 = t == t_last ? ((quad_end - 1) % kQuadsPerTile) * 4 + 4 : kTileH;
  };
  UnitPlan nxt;
  nxt.nr = 0;
  bool reuse = false;
  if (DIC_SPECULATE && spec->pending) {
    if (quad_begin < quad_end && spec->level == level && spec->quad_begin == quad_begin && spec->tiles == tl.tiles) {
      int r0, r1;
      rows_of(t_first, r0, r1);
      TileHead h;
      h.x0 = spec->x0; h.y0 = spec->y0; h.full_rows = spec->full_rows; h.col = spec->col[lane];
      nxt = plan_unit<NP, GRAN>(h, r0, r1, p, ccx, ccy, def);
      const int sx0 = spec->px0, sy0 = spec->py0;
      // the footprint under the real parameters must lie inside the box that was staged (warp-uniform: plan values are)
      if (nxt.nr > 0 && nxt.staged && nxt.pxr >= sx0 && nxt.pxe <= sx0 + kPatchW && nxt.py0 >= sy0 && nxt.pye <= sy0 + kPatchH) {
        nxt.px0 = sx0; nxt.py0 = sy0;
        reuse = true;
        __syncwarp();
        if (lane == 0) spec->pending = 0; // consumed by the unit loop below like any staged unit
        __syncwarp();
      }
    }
    if (!reuse) spec_drop(st, spec, timeout_flag);
  }
  if (!reuse && quad_begin < quad_end) {
    int r0, r1;
    rows_of(t_first, r0, r1);
    const TileHead head_first = load_tile_head(tl, t_first);
    nxt = plan_unit<NP, GRAN>(head_first, r0, r1, p, ccx, ccy, def);
    issue_unit(st, nxt, map_def, map_und);
    if (DIC_SPECULATE) { // parked in the warp's slot for the end of this pass (not in registers across the pixel loop)
      spec->col[lane] = head_first.col;
      if (lane == 0) { spec->x0 = head_first.x0; spec->y0 = head_first.y0; spec->full_rows = head_first.full_rows; }
    }
  }
  TileHead head_next = {0, 0, 0u, 0u}; // record of tile t + 1, in flight while tile t - 1 ... t is evaluated
  if (quad_begin < quad_end && t_first + 1 <= t_last) head_next = load_tile_head(tl, t_first + 1);
  for (int t = t_first; quad_begin < quad_end && t <= t_last; ++t) {
    const UnitPlan q = nxt;
    if (t + 1 <= t_last) {
      int r0, r1;
      rows_of(t + 1, r0, r1);
      nxt = plan_unit<NP, GRAN>(head_next, r0, r1, p, ccx, ccy, def);
      issue_unit(st, nxt, map_def, map_und);
      if (t + 2 <= t_last) head_next = load_tile_head(tl, t + 2);
    }
    if (q.nr == 0) continue;
    const int x0 = q.x0, y0 = q.y0, nr = q.nr;
    if (x0 != cur_x0) {
      if (cur_x0 != INT_MIN) flush_moments<NP>(mom, X, warp_acc);
      cur_x0 = x0;
      xf = (float)(x0 + lane);
      X = __fsub_rn(xf, ccx);
      lw.set(p, xf, X);
      if (MODE == DIC_MODE_PARITY && DIC_PARITY_LOOP == 3) pwp.set(p, xf, ccx, ccy);
    }
    const uint32_t colmask = q.colmask;
    if (q.staged) {
      const uint8_t *patch = st.buf + (st.consumed & 1) * kStageBytes;
      const uint8_t *ucol = patch + kPatchBytes + (x0 & 15) + lane; // reference pixels of this lane's column
      const bool landed = mbar_wait(st.bar + (st.consumed & 1), (st.consumed >> 1) & 1);
      ++st.consumed;
      if (!landed) { // never expected: end the launch with error_cuda instead of consuming garbage (or hanging)
        if (lane == 0) atomicExch(timeout_flag, 1);
        break;
      }
      const int px0 = q.px0, py0 = q.py0;
      float cw[4][4];
      int wix = INT_MIN, wiy = INT_MIN; // no window yet: the first row rebuilds all four
#define DIC_STEP(FULLV, SV, R)                                                                       \
  staged_pixel<MODEL, MODE, FULLV, SV>(pw, lw, xf, (float)(y0 + (R)), ccx, ccy, patch, px0, py0,      \
                                       (float)ucol[(R) * kUndW], FULLV || ((colmask >> (R)) & 1u) != 0, \
                                       cw, wix, wiy, mom)
      // fast mode: four pixels per trip so that the window rotation is a static renaming of registers.
      // parity mode: one pixel per trip and an explicit 12-register shift -- its per-pixel code is
      // 2.4x longer and the unrolled body overflowed the instruction cache (10 % no-instruction stalls).
#define DIC_SHIFT()                                                                                  \
  _Pragma("unroll") for (int k_ = 0; k_ < 4; ++k_) { cw[0][k_] = cw[1][k_]; cw[1][k_] = cw[2][k_]; cw[2][k_] = cw[3][k_]; }
      if (MODE == DIC_MODE_PARITY && DIC_PARITY_LOOP == 3) {
        // two streams of nr / 2 rows (GRAN = 2 keeps nr even); one code path for full and partial units
        const int half = nr >> 1;
        int r = 0;
        PairWindow win;
        // ONE copy of the step (DIC_PAIR_UNROLL == 1, the default): the window rows move between registers, 12 pair
        // copies per step -- issue slots the packed loop has to spare. The statically rotated four-step form of the
        // scalar loop (DIC_PAIR_UNROLL == 4) is 26 KB of code here and stalled 1.3 cycles per issue on instruction
        // fetch (profiles/r2_pair_loop_unrolled_ncu_full.txt).
#pragma unroll 1
        while (r < half) {
          f2 cw2[4][4];
          f2 Yp = pk((float)(y0 + r), (float)(y0 + half + r));
          const uint8_t *up_a = ucol + r * kUndW, *up_b = ucol + (half + r) * kUndW;
          uint32_t cm_a = colmask >> r, cm_b = colmask >> (half + r);
          parity_pair_open<MODEL>(pwp, Yp, cfg.opaque_zero, patch, px0, py0, cw2, win);
#if DIC_PAIR_UNROLL == 4
#define DIC_PSTEP(SV)                                                                                        \
  {                                                                                                          \
    const f2 U = pk((float)up_a[(SV) * kUndW], (float)up_b[(SV) * kUndW]);                                      \
    const f2 Mk = pk(((cm_a >> (SV)) & 1u) ? 1.f : 0.f, ((cm_b >> (SV)) & 1u) ? 1.f : 0.f);                      \
    if (!parity_pair_step<MODEL, SV>(pwp, Yp, cfg.opaque_zero, cw2, win, U, Mk, mom)) break;                                  \
    Yp = add2(Yp, bc(1.f));                                                                                  \
    if (++r >= half) break;                                                                                  \
  }
#pragma unroll 1
          while (true) {
            DIC_PSTEP(0) DIC_PSTEP(1) DIC_PSTEP(2) DIC_PSTEP(3)
            win.wp_a += 4 * kPatchW; win.wp_b += 4 * kPatchW;
            up_a += 4 * kUndW; up_b += 4 * kUndW;
            cm_a >>= 4; cm_b >>= 4;
          }
#undef DIC_PSTEP
#else
#pragma unroll 1
          do {
            const f2 U = pk((float)*up_a, (float)*up_b);
            const f2 Mk = pk((cm_a & 1u) ? 1.f : 0.f, (cm_b & 1u) ? 1.f : 0.f);
            if (!parity_pair_step<MODEL, 0>(pwp, Yp, cfg.opaque_zero, cw2, win, U, Mk, mom)) break;
#pragma unroll
            for (int k_ = 0; k_ < 4; ++k_) { cw2[0][k_] = cw2[1][k_]; cw2[1][k_] = cw2[2][k_]; cw2[2][k_] = cw2[3][k_]; }
            win.wp_a += kPatchW; win.wp_b += kPatchW;
            up_a += kUndW; up_b += kUndW;
            cm_a >>= 1; cm_b >>= 1;
            Yp = add2(Yp, bc(1.f));
          } while (++r < half);
#endif
        }
      } else if (MODE == DIC_MODE_PARITY && DIC_PARITY_LOOP == 2) {
        // one code path for full and partial units (three selects per pixel buy half the instruction footprint)
        const uint32_t members = q.full ? 0xffffffffu : colmask;
        int r = 0;
#if DIC_YSTAGE_PACKED
        f2 cwp[4][2];
#define DIC_PSTEP(SV)                                                                                              \
  if (!parity_slide_step_trip_pairs<MODEL, SV>(pw, xf, yf, ccx, ccy, cwp, wix, wiy, wp, wsh, (float)up[(SV) * kUndW], \
                                               ((cm >> (SV)) & 1u) != 0, mom)) { r += (SV); break; }               \
  if (left <= (SV) + 1) { r = nr; break; }
#else
#define DIC_PSTEP(SV)                                                                                              \
  if (!parity_slide_step_trip<MODEL, SV>(pw, xf, yf, ccx, ccy, cw, wix, wiy, wp, wsh, (float)up[(SV) * kUndW],      \
                                         ((cm >> (SV)) & 1u) != 0, mom)) { r += (SV); break; }                     \
  if (left <= (SV) + 1) { r = nr; break; }
#endif
#pragma unroll 1
        while (r < nr) {
          // (re)open the window at row r; the loop-carried values of a trip follow from r
          float yf = (float)(y0 + r);
          const uint8_t *wp, *up = ucol + r * kUndW; // wp: window rows in the patch; up: reference pixel of the row
          int wsh, left = nr - r;
          uint32_t cm = members >> r;
#if DIC_YSTAGE_PACKED
          parity_window_open_pairs<MODEL>(pw, xf, yf, ccx, ccy, patch, px0, py0, cwp, wix, wiy, wp, wsh);
#else
          parity_window_open<MODEL>(pw, xf, yf, ccx, ccy, patch, px0, py0, cw, wix, wiy, wp, wsh);
#endif
#pragma unroll 1
          while (true) {
            DIC_PSTEP(0) DIC_PSTEP(1) DIC_PSTEP(2) DIC_PSTEP(3)
            r += 4; left -= 4; yf += 4.f; wp += 4 * kPatchW; up += 4 * kUndW; cm >>= 4;
          }
        }
#undef DIC_PSTEP
      } else if (MODE == DIC_MODE_PARITY) {
        if (q.full) {
#pragma unroll 1
          for (int r = 0; r < nr; ++r) { DIC_STEP(true, 0, r); DIC_SHIFT(); }
        } else {
#pragma unroll 1
          for (int r = 0; r < nr; ++r) { DIC_STEP(false, 0, r); DIC_SHIFT(); }
        }
      } else if (BATCH && DIC_FAST_UNROLL_BATCH == 1) {
        // fast mode, batch form: ONE copy of the pixel step, the window rows move between registers. Four CTAs per SM
        // at four different places of the statically rotated four-step loop stalled 1.65 cycles per issue on
        // instruction fetch (profiles/r2_c4_batch_tiles_fast_ncu_full.txt): 2.00 -> 1.75 ms on c4. The grid form (two
        // CTAs per SM walking in step) keeps the four-step loop: it is 7 % / 15 % faster there (c2 / c5).
#pragma unroll 1
        for (int r = 0; r < nr; ++r) { DIC_STEP(false, 0, r); DIC_SHIFT(); }
      } else if (q.full) {
#pragma unroll 1
        for (int r = 0; r < nr; r += 4) { DIC_STEP(true, 0, r); DIC_STEP(true, 1, r + 1); DIC_STEP(true, 2, r + 2); DIC_STEP(true, 3, r + 3); }
      } else {
#pragma unroll 1
        for (int r = 0; r < nr; r += 4) { DIC_STEP(false, 0, r); DIC_STEP(false, 1, r + 1); DIC_STEP(false, 2, r + 2); DIC_STEP(false, 3, r + 3); }
      }
#undef DIC_SHIFT
#undef DIC_STEP
      __syncwarp(); // every lane is done with this buffer before lane 0 hands it to the next copy
    } else {
      if (slow_counter && lane == 0) atomicAdd(slow_counter, 1u);
      // footprint leaves the image or the staging buffer: per-pixel path with the reference's
      // own bounds test (error 2 + zero contribution, interpolation_class.cpp:129-137)
      const uint8_t *ucol = und.ptr + (size_t)y0 * und.pitch + x0 + lane;
#pragma unroll 1
      for (int r = 0; r < nr; ++r) {
        if (!((colmask >> r) & 1u)) continue;
        const float yf = (float)(y0 + r);
        float xd, yd, dxx, dyy, w, wx, wy;
        warp_point<MODEL, MODE>(p, xf, yf, ccx, ccy, xd, yd, dxx, dyy);
        if (!sample_def<DIC_IM_BICUBIC, MODE>(def, xd, yd, w, wx, wy)) mom[M::kOob] += 1.f;
        accumulate_moments<NP>(mom, (float)__ldg(ucol + (size_t)r * und.pitch) - w, wx, wy, dyy);
      }
    }
  }
  if (DIC_SPECULATE && quad_begin < quad_end && !__shfl_sync(0xffffffffu, *(volatile int *)timeout_flag, 0)) {
    // stage the first unit of this range again for the next evaluation (see SpecSlot), before the flush below
    int r0, r1;
    rows_of(t_first, r0, r1);
    __syncwarp();
    TileHead h;
    h.x0 = spec->x0; h.y0 = spec->y0; h.full_rows = spec->full_rows; h.col = spec->col[lane];
    UnitPlan q = plan_unit<NP, GRAN>(h, r0, r1, p, ccx, ccy, def);
    if (q.nr > 0 && q.staged) {
      const int room_y = kPatchH - (q.pye - q.py0);
      q.py0 -= min(2, room_y >> 1);
      if (q.pxr - q.px0 < 2 && q.pxe - (q.px0 - 16) <= kPatchW) q.px0 -= 16;
      issue_unit(st, q, map_def, map_und);
      if (lane == 0) {
        spec->pending = 1; spec->level = level; spec->quad_begin = quad_begin; spec->tiles = tl.tiles;
        spec->px0 = q.px0; spec->py0 = q.py0;
      }
      __syncwarp();
    }
  }
  if (cur_x0 != INT_MIN) flush_moments<NP>(mom, X, warp_acc);
}

// Duplicate pixels of a blob list (beyond their first occurrence), handled by one warp.
template <int MODEL, int MODE>
__device__ __forceinline__ void evaluate_extras(const SolveSettings &cfg, float cx0, float cy0,
                                                const TileLevel tl, int level, const float *p,
                                                float *warp_acc) {
  constexpr int NP = model_nparams(MODEL);
  using M = Mom<NP>;
  const int lane = threadIdx.x & 31;
  const LevelImage und = cfg.und[level];
  const LevelImage def = cfg.def[level];
  const float inv = 1.f / (float)(1 << level);
  const float ccx = cx0 * inv, ccy = cy0 * inv;
  float mom[M::kN];
#pragma unroll
  for (int i = 0; i < M::kN; ++i) mom[i] = 0.f;
  for (int base = 0; base < tl.n_extra; base += 32) {
    const int i = base + lane;
    float Xe = 0.f;
    if (i < tl.n_extra) {
      float2 q = __ldg(tl.extra + i);
      float xd, yd, dxx, dyy, w, wx, wy;
      warp_point<MODEL, MODE>(p, q.x, q.y, ccx, ccy, xd, yd, dxx, dyy);
      if (!sample_def<DIC_IM_BICUBIC, MODE>(def, xd, yd, w, wx, wy)) mom[M::kOob] += 1.f;
      int uix = (int)(q.x + 0.5f), uiy = (int)(q.y + 0.5f);
      float V = (float)__ldg(und.ptr + (size_t)uiy * und.pitch + uix) - w;
      accumulate_moments<NP>(mom, V, wx, wy, dyy);
      Xe = dxx;
    }
    flush_moments<NP>(mom, Xe, warp_acc);
  }
}

// ---- the solve kernel on tiles (same LM / barrier / solve machinery as gn_solve_kernel)
//   GRID         one sector, every CTA of a cooperative launch works on it (large domains)
//   !GRID, CL=1  each CTA owns whole sectors (BASELINE config 4: thousands of small subsets)
//   !GRID, CL=2  each CTA PAIR (thread-block cluster of 2) owns whole sectors
// Threads per CTA of the BATCH form (grid form: always kThreads). 128 = four warps per subset and four CTAs per SM:
// while one subset's warp 0 runs its LM step + solve only three warps wait instead of seven, and the SM still
// holds 16 warps (measured on c4: see DESIGN.md section 4.1).
#ifndef DIC_BATCH_THREADS
#define DIC_BATCH_THREADS 128
#endif
#ifndef DIC_GRID_THREADS
#define DIC_GRID_THREADS kThreads
#endif
#ifndef DIC_BATCH_REGS
#define DIC_BATCH_REGS 128 // register budget per thread the batch form is compiled for (CTAs per SM = 64 K / regs / threads)
#endif
#ifndef DIC_GRID_CTAS
#define DIC_GRID_CTAS (tile_ctas_per_sm(model) * kThreads / DIC_GRID_THREADS)
#endif
// the 12-parameter model needs one thread per accumulator (92) in a few places: its CTAs keep four warps
// Grid form of the 12-parameter model: ONE CTA of 384 threads per SM (up to 168 registers: no spills) instead of two
// of 256 at 128 registers (DIC_GRID_QUAD_ONE_CTA). Measured on c2 (B200): 0.569 ms with 2 x 256, 0.545 with 3 x 128,
// 0.516 with 1 x 384, 0.560 with 1 x 512. Besides the registers, the warps of ONE CTA are the same age for the warp
// scheduler and finish a pass within 9 % of each other; with several CTAs per SM the oldest CTA is served first and the
// youngest ends 25-40 % later (tools/probe_tl.py, per-CTA end times), which every other CTA then waits for at the
// grid barrier -- and the barrier has 148 participants instead of 296 / 444. The affine grid form keeps two CTAs of
// 256: c1 (few quads per warp) is 13-27 % slower with one big CTA and c5 within 2 % either way.
#ifndef DIC_GRID_QUAD_ONE_CTA
#define DIC_GRID_QUAD_ONE_CTA 1
#endif
__host__ __device__ constexpr bool tile_grid_one_cta(int model, int /*mode: both arithmetic modes gain*/) {
  return DIC_GRID_QUAD_ONE_CTA && model == DIC_FM_QUADRATIC;
}
__host__ __device__ constexpr int tile_cta_threads(int model, int mode, bool grid) {
  return grid ? (tile_grid_one_cta(model, mode) ? 384 : DIC_GRID_THREADS)
              : (model == DIC_FM_QUADRATIC && DIC_BATCH_THREADS < 128 ? 128 : DIC_BATCH_THREADS);
}
__host__ __device__ constexpr int tile_ctas_for(int model, int mode, bool grid) {
  return grid ? (tile_grid_one_cta(model, mode) ? 1 : DIC_GRID_CTAS)
              : 65536 / DIC_BATCH_REGS / tile_cta_threads(model, mode, false);
}

template <int MODEL, int MODE, bool GRID, int CL>
__global__ void __launch_bounds__(tile_cta_threads(MODEL, MODE, GRID), tile_ctas_for(MODEL, MODE, GRID))
gn_solve_tiles_kernel(const SolveSettings cfg, const __grid_constant__ TileMaps maps,
                      const SectorDev *__restrict__ sectors, const SectorTiles *__restrict__ sector_tiles,
                      const float *guesses, const GuessParam guess0, dic_result *__restrict__ results, int first_sector,
                      int n_sectors, GridWork *work) {
  constexpr int NP = model_nparams(MODEL);
  constexpr int NACC = Acc<NP>::kN;
  static_assert(!GRID || CL == 1, "clusters are a batch-mode feature");
  constexpr int NT = tile_cta_threads(MODEL, MODE, GRID), NW = NT / 32; // threads, warps of this CTA
  static_assert(Acc<NP>::kN <= NT && kMaxParams <= NT, "one thread per accumulator / parameter");
  extern __shared__ __align__(128) uint8_t dyn_smem[];
  uint8_t *s_stage = dyn_smem;                                                    // [warps][kWarpStageBytes]
  float *s_wacc = reinterpret_cast<float *>(dyn_smem + NW * kWarpStageBytes);     // [warps][NACC]
  __shared__ SolveShared<NP> sh;
  __shared__ __align__(8) uint64_t s_bar[NW][2];
  __shared__ SpecSlot s_spec[NW];
  __shared__ SectorTiles s_tiles; // this sector's tile lists and centre: read once, not once per evaluation
  __shared__ float s_center[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float *warp_acc = s_wacc + warp * NACC;
  WarpStage st;
  st.buf = s_stage + warp * kWarpStageBytes;
  st.bar = s_bar[warp];
  st.issued = st.consumed = 0;
  if (lane == 0) {
    mbar_init(&st.bar[0], 1);
    mbar_init(&st.bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (lane == 0) s_spec[warp].pending = 0;
  if (tid == 0) { sh.xch_count = 0; sh.timed_out = 0; } // timed_out is sticky for the rest of the launch
  __syncwarp();
  const int crank = CL == 2 ? (int)cluster_ctarank() : 0;
  if (CL == 2) cluster_sync_all(); // the partner's shared memory exists from here on
  const int group = GRID ? 0 : (int)blockIdx.x / CL, n_groups = GRID ? 1 : (int)gridDim.x / CL;

  // Batch form, one CTA per sector: the first sector is the CTA's own index, every further one comes from a launch-wide
  // ticket counter (the sectors of a batch differ in LM iterations, and n_sectors is rarely a multiple of the CTA
  // slots: with a static stride the launch ends when the unluckiest CTA does). A sector is still solved by ONE CTA
  // with the same instruction sequence whichever CTA that is, so the records do not depend on the assignment.
  constexpr bool kQueue = DIC_BATCH_QUEUE && !GRID && CL == 1;
  __shared__ int s_next_sector;
  for (int si = group; si < n_sectors;) {
    const SectorDev *sec = sectors + first_sector + si;
    const SectorTiles *stl = sector_tiles + first_sector + si;
    const float *guess = guesses + (size_t)(first_sector + si) * kMaxParams;
    dic_result *result = results + first_sector + si;
    unsigned int my_gen;
    if (tid < kMaxLevels) s_tiles.lev[tid] = stl->lev[tid];
    if (tid == kMaxLevels) { s_center[0] = sec->cx; s_center[1] = sec->cy; }
    begin_sector<MODEL, GRID>(sh, cfg, sec, guess, guess0, work, my_gen); // ends with a CTA barrier
#if DIC_BATCH_TIMELINE
    int dbg_mark = 0; // diagnostics build: CTA 0 records the phases of its LAST sector's evaluations (tools/probe_batch_tl.py)
#endif
    while (true) {
      const int level = sh.level;
#if DIC_BATCH_TIMELINE
      if (!GRID && blockIdx.x == 0 && tid == 0 && dbg_mark < kMaxMarks) work->marks[dbg_mark][0] = global_ns();
#endif
      for (int k = lane; k < NACC; k += 32) warp_acc[k] = 0.f;
      __syncwarp();
      const TileLevel tl = s_tiles.lev[level];
      // quads of the level over the warps that take part: at least one quad per active warp
      const int n_quads = tl.n_tiles * kQuadsPerTile;
      const int ctas = GRID ? (int)gridDim.x : CL;
      const int n_active = max(1, min((n_quads + NW - 1) / NW, ctas));
      const int cta = GRID ? (int)blockIdx.x : crank;
      const bool active = cta < n_active;
      if (active) {
        const int nw = n_active * NW;
        const int wg = cta * NW + warp;
        // balanced contiguous ranges: the first (n_quads % nw) warps take one quad more
        const int base = n_quads / nw, rem = n_quads - base * nw;
        const int qb = wg * base + min(wg, rem), qe = qb + base + (wg < rem ? 1 : 0);
        evaluate_tiles<MODEL, MODE, !GRID>(cfg, maps, s_center[0], s_center[1], tl, level, sh.p, qb, qe, st, &s_spec[warp],
                                    warp_acc, GRID ? &work->slow_units : nullptr, &sh.timed_out);
        if (tl.n_extra > 0 && wg == 0)
          evaluate_extras<MODEL, MODE>(cfg, s_center[0], s_center[1], tl, level, sh.p, warp_acc);
      }
      __syncthreads();
#if DIC_BATCH_TIMELINE
      // diagnostics: [0] pass started, [1] every warp's pass done, [2] LM step started, [3] LM step done (reduce_and_step)
      if (!GRID && blockIdx.x == 0 && tid == 0 && dbg_mark < kMaxMarks) { work->marks[dbg_mark][1] = global_ns(); sh.mark = dbg_mark; }
#endif
      for (int k = tid; k < NACC; k += NT) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) s += s_wacc[w * NACC + k];
        sh.tot[k] = s;
      }
      __syncthreads();
      reduce_and_step<MODEL, GRID, CL>(sh, active, n_active, cfg, sec, result, work, my_gen);
#if DIC_BATCH_TIMELINE
      if (!GRID && blockIdx.x == 0 && tid == 0 && dbg_mark < kMaxMarks) work->n_marks = ++dbg_mark;
#endif
      if (sh.done) break;
    }
    if (kQueue && tid == 0) s_next_sector = n_groups + (int)atomicAdd(&work->next_sector, 1u);
    __syncthreads();
    si = GRID ? n_sectors : kQueue ? s_next_sector : si + n_groups;
  }
  if (DIC_SPECULATE) spec_drop(st, &s_spec[warp], &sh.timed_out); // no bulk copy may be in flight when the CTA exits
  if (GRID) grid_depart<NACC>(work, sh.rs_seq, sh.rowsplit != 0);
  if (kQueue && tid == 0) { // the last CTA to leave re-arms the ticket counter for the next launch
    __threadfence();
    if (atomicAdd(&work->batch_departed, 1u) == gridDim.x - 1) {
      work->next_sector = 0; work->batch_departed = 0;
      __threadfence();
    }
  }
}

constexpr size_t tiles_dyn_smem(int nacc, int threads) {
  return (size_t)(threads / 32) * kWarpStageBytes + sizeof(float) * (size_t)(threads / 32) * nacc;
}

// ------------------------------------------------------------------ tile construction

// Scatter a level's pixel list into the mask grid (column-major slots). A pixel whose bit is
// already set is a duplicate of the reference list and goes to `extra`.
__global__ void tiles_scatter_kernel(const float2 *__restrict__ xy, long n, int gx0, int gy0, int nty,
                                     uint32_t *masks, float2 *extra, int extra_cap, int *n_extra) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float2 q = xy[i];
  int x = (int)(q.x + 0.5f) - gx0, y = (int)(q.y + 0.5f) - gy0;
  int tx = x / kTileW, ty = y / kTileH;
  uint32_t bit = 1u << (x - tx * kTileW);
  uint32_t old = atomicOr(&masks[((size_t)tx * nty + ty) * kTileH + (y - ty * kTileH)], bit);
  if (old & bit) {
    int k = atomicAdd(n_extra, 1);
    if (k < extra_cap) extra[k] = q;
  }
}

struct TilePred {
  const uint32_t *masks;
  int gx0, gy0, nty;
  __device__ __forceinline__ bool operator()(long idx, Tile &out) const {
    const uint32_t *m = masks + (size_t)idx * kTileH;
    uint32_t any = 0;
#pragma unroll
    for (int r = 0; r < kTileH; ++r) { out.rows[r] = m[r]; any |= m[r]; }
    int tx = (int)(idx / nty), ty = (int)(idx % nty);
    out.x0 = gx0 + tx * kTileW; out.y0 = gy0 + ty * kTileH;
    if (any != 0) tile_finish(out);
    return any != 0;
  }
};

template <class Pred, class Out, bool EMIT>
__global__ void __launch_bounds__(kCompactThreads)
compact_any_kernel(Pred pred, long ncand, unsigned int *block_counts,
                   const unsigned long long *block_offsets, Out *__restrict__ out) {
  __shared__ unsigned int warp_tot[kCompactThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long base = (long)blockIdx.x * kCompactChunk;
  unsigned long long running = EMIT ? block_offsets[blockIdx.x] : 0ull;
  unsigned int total = 0;
  for (int it = 0; it < kCompactItems; ++it) {
    long idx = base + (long)it * kCompactThreads + tid;
    Out q;
    bool keep = idx < ncand && pred(idx, q);
    unsigned int bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    unsigned int before = 0, all = 0;
#pragma unroll
    for (int w = 0; w < kCompactThreads / 32; ++w) {
      unsigned int t = warp_tot[w];
      if (w < warp) before += t;
      all += t;
    }
    if (EMIT && keep) out[running + before + __popc(bal & ((1u << lane) - 1u))] = q;
    running += all;
    total += all;
    __syncthreads();
  }
  if (!EMIT && tid == 0) block_counts[blockIdx.x] = total;
}

// Rectangle in level coordinates [gx0, gx0 + nx) x [gy0, gy0 + ny): tiles in closed form,
// column-major (tile index = tx * nty + ty).
__global__ void rect_tiles_kernel(Tile *__restrict__ out, int ntx, int nty, int gx0, int gy0, int nx, int ny) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntx * nty) return;
  int tx = t / nty, ty = t - tx * nty;
  Tile q;
  q.x0 = gx0 + tx * kTileW; q.y0 = gy0 + ty * kTileH;
  int wcols = min(kTileW, nx - tx * kTileW), hrows = min(kTileH, ny - ty * kTileH);
  uint32_t m = wcols >= 32 ? 0xffffffffu : ((1u << wcols) - 1u);
#pragma unroll
  for (int r = 0; r < kTileH; ++r) q.rows[r] = r < hrows ? m : 0u;
  tile_finish(q);
  out[t] = q;
}

// ---- a whole grid of rectangles at once (dic_reset_polygon_rect_grid): one descriptor per (sector, level)
struct RectDesc {
  int xs, ys, nx, ny, mag; // level-0 coordinates of the first kept pixel, kept pixels per row / column, 2^level
  int ntx, nty;            // tiles across / down
  long long list_off, tile_off; // offsets (elements) into the shared list / tile blocks
};
__global__ void rect_grid_fill_kernel(const RectDesc *__restrict__ desc, int n_desc, float2 *__restrict__ lists) {
  for (int di = blockIdx.y; di < n_desc; di += gridDim.y) {
    const RectDesc d = desc[di];
    const int n = d.nx * d.ny; // one rectangle at one level: far below 2^31 pixels
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) continue;
    const int row = i / d.nx, col = i - row * d.nx;
    const float inv = 1.f / (float)d.mag;
    lists[d.list_off + i] = make_float2((float)(d.xs + col * d.mag) * inv, (float)(d.ys + row * d.mag) * inv);
  }
}
__global__ void rect_grid_tiles_kernel(const RectDesc *__restrict__ desc, int n_desc, Tile *__restrict__ tiles) {
  for (int di = blockIdx.y; di < n_desc; di += gridDim.y) {
    const RectDesc d = desc[di];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= d.ntx * d.nty) continue;
    const int tx = t / d.nty, ty = t - tx * d.nty;
    Tile q;
    q.x0 = d.xs / d.mag + tx * kTileW; q.y0 = d.ys / d.mag + ty * kTileH;
    const int wcols = min(kTileW, d.nx - tx * kTileW), hrows = min(kTileH, d.ny - ty * kTileH);
    const uint32_t m = wcols >= 32 ? 0xffffffffu : ((1u << wcols) - 1u);
#pragma unroll
    for (int r = 0; r < kTileH; ++r) q.rows[r] = r < hrows ? m : 0u;
    tile_finish(q);
    tiles[d.tile_off + t] = q;
  }
}

// integer bounding box of a list (for blob / point-list sectors)
__global__ void bbox_kernel(const float2 *__restrict__ xy, long n, int *box /*minx miny maxx maxy*/) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  int mnx = INT_MAX, mny = INT_MAX, mxx = INT_MIN, mxy = INT_MIN;
  for (; i < n; i += (long)gridDim.x * blockDim.x) {
    float2 q = xy[i];
    int x = (int)(q.x + 0.5f), y = (int)(q.y + 0.5f);
    mnx = min(mnx, x); mny = min(mny, y); mxx = max(mxx, x); mxy = max(mxy, y);
  }
  for (int o = 16; o > 0; o >>= 1) {
    mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, o));
    mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, o)); mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&box[0], mnx); atomicMin(&box[1], mny); atomicMax(&box[2], mxx); atomicMax(&box[3], mxy);
  }
}

} // namespace dic
