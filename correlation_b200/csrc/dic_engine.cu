// dic_engine.cu -- host side of libdic_b200.so: the C-ABI of include/dic_b200.h.
//
// Replaces the reference's cuda_class / cuda_pyramid / cuda_polygon / cuda_solver host code.
// One engine = one CUDA device, one correlation stream, one image stream; device memory is owned
// here: three image pyramids (und / def / nxt, rotated by index like pyramid_class.cpp:211-258),
// per-sector pixel lists for every used pyramid level, and a few hundred bytes of LM state.
#include <cooperative_groups.h>
#include <cuda.h>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h> // header-only; ranges are free unless a profiler injects the NVTX library

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "dic_kernels.cuh"
#include "dic_polygon.hpp"
#include "dic_tiles.cuh"

using namespace dic;

namespace {

// NVTX ranges where the reference has them (cuda_class.cu:133-326 per level / solve / update, :497-554 image
// reads, cuda_polygon.cuh:443, cuda_pyramid.cuh:58): one range per C-ABI call that enqueues device work.
struct NvtxRange {
  explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

enum SectorKind { SK_NONE = 0, SK_RECT, SK_ANNULAR, SK_BLOB, SK_POINTS };

struct Sector {
  SectorKind kind = SK_NONE;
  float2 *buf = nullptr; // one allocation for all levels
  size_t cap = 0;        // in float2
  float2 *xy[kMaxLevels] = {};
  long n[kMaxLevels] = {};
  long n_total[kMaxLevels] = {}; // whole-domain counts when this GPU holds only a band of rows
  bool banded = false;
  float cx = 0.f, cy = 0.f;
  bool integer_grid = true;
  // geometry kept for reference-order regeneration
  int rx0 = 0, ry0 = 0, rx1 = 0, ry1 = 0;
  bool pending = false; // an async correlate is in flight
  // structured form (dic_tiles.cuh): column-major 32x16 tiles per level (+ duplicate pixels)
  Tile *tbuf = nullptr;
  size_t tcap = 0;
  bool buf_owned = false, tbuf_owned = false; // false: lives in the engine arena
  float2 *ebuf = nullptr;
  size_t ecap = 0;
  TileLevel tl[kMaxLevels] = {};
  bool has_tiles = false;
};

struct PyramidSlot {
  uint8_t *base = nullptr;
  size_t cap = 0;
  LevelImage lev[kMaxLevels] = {};
  // TMA descriptors of every level for the two box shapes the tile kernel stages
  CUtensorMap tm_patch[kMaxLevels], tm_tile[kMaxLevels];
  CUtensorMap tm_pyr[kMaxLevels]; // the level as the SOURCE of the pyramid kernel (box kPyrSW x kPyrSH)
  int rows = 0, cols = 0;
  bool valid = false;
};

} // namespace

struct dic_engine {
  int device = 0;
  int num_sms = 0;
  cudaStream_t stream = nullptr, img_stream = nullptr, copy_stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_img = nullptr, ev_gn = nullptr, ev_copy = nullptr, ev_copy_und = nullptr;
  cudaEvent_t ev_rot = nullptr; // correlation-stream position at the last pyramid rotation (see dic_reset_next_pyramid)
  cudaEvent_t ev_step0 = nullptr, ev_step1 = nullptr; // around the whole GPU side of a correlate (copies included)
  float last_step_ms = 0.f;
  PyramidSlot pyr[7];
  // role (0 und, 1 def, 2 nxt, 3 / 4 staged und / def of the next pair, 5 / 6 of the pair after it) -> slot
  int role[7] = {0, 1, 2, 3, 4, 5, 6};
  int staged = 0;                      // pairs staged and not yet advanced to (0..2)
  cudaEvent_t ev_ready[2] = {nullptr, nullptr}; // image-stream position behind the pyramid build of staged pair 0 / 1
  int start = 0, step = 1, stop = 0;
  int max_iters = 50;
  float precision = 1e-3f;
  int model = DIC_FM_UVUxUyVxVy, interp = DIC_IM_BICUBIC, mode = DIC_MODE_PARITY;
  int center_mode = DIC_CENTER_REFERENCE;
  int colors = 1; // channels of the images of the last dic_reset_image_pyramids (1, or 3 interleaved)
  std::vector<Sector> sectors;
  SectorDev *d_sectors = nullptr;
  SectorTiles *d_sector_tiles = nullptr;
  uint32_t *d_masks = nullptr;
  size_t cap_masks = 0;
  SectorDev *h_sectors = nullptr;      // pinned mirrors: async upload of the sector records
  SectorTiles *h_sector_tiles = nullptr;
  // bump arena for small sector buffers (thousands of subsets: no cudaMalloc per sector)
  std::vector<char *> arena_chunks;
  char *arena_cur = nullptr;
  size_t arena_left = 0;
  int kernel_variant = 0; // 0 auto, 1 pixel-list kernel, 2 tile kernel
  Mailbox *d_mailbox = nullptr; // row-split: peers write their sums here
  void *peer_mailbox[kMaxRanks] = {};
  int rs_rank = 0, rs_world = 1;
  float *d_guess = nullptr;
  dic_result *d_results = nullptr;
  dic_result *h_results = nullptr; // pinned and mapped: the kernels write the records straight into it
  dic_result *h_results_dev = nullptr; // the same block as the device sees it
  float *h_guess = nullptr;        // pinned + mapped: batch kernels read their guesses from here (zero-copy)
  float *h_guess_dev = nullptr;    // device-side address of h_guess
  int cluster_mode = 0;            // batch launches: 0 auto, 1 one CTA per sector, 2 one CTA pair per sector
  // diagnostics (dic_pipe_trace): device-time marks of the staged-pair pipeline
  static constexpr int kTrace = 48;
  bool trace_on = false;
  int trace_stage = 0, trace_solve = 0;
  cudaEvent_t trace_base = nullptr, trace_st[kTrace][4] = {}, trace_so[kTrace][2] = {};
  int batch_queue = 0;             // batch launches: 0 / 1 resident CTAs + ticket queue, 2 one CTA per sector (dic_set_batch_queue)
  // device blocks shared by the sectors of one dic_reset_polygon_rect_grid call, keyed by its first sector id:
  // rebuilding the same range reuses (or regrows) its blocks, another range gets its own
  struct GridBlock { int first_id = 0; void *lists = nullptr, *tiles = nullptr, *desc = nullptr; size_t cap_lists = 0, cap_tiles = 0, cap_desc = 0; };
  std::vector<GridBlock> grid_blocks;
  int last_cluster = 1;            // CTAs per sector of the last batch launch (1 or 2)
  int cap_sectors = 0;
  GridWork *d_work = nullptr;
  float *d_partials = nullptr;
  int max_grid = 0;
  float *d_scratch = nullptr; // 256 floats + ints for small kernels
  unsigned int *d_counts = nullptr;
  unsigned long long *d_offsets = nullptr;
  size_t cap_counts = 0;
  uint8_t *d_stage = nullptr;
  size_t cap_stage = 0;
  float last_ms = 0.f;
  bool timing_pending = false;
  std::atomic<long long> launches{0};
  std::string err;
  std::mutex err_mutex;
};

namespace {

#define CU_TRY(e, call)                                                                   \
  do {                                                                                    \
    cudaError_t _r = (call);                                                              \
    if (_r != cudaSuccess) {                                                              \
      set_error(e, std::string(#call) + ": " + cudaGetErrorString(_r));                   \
      return DIC_ERROR_CUDA;                                                              \
    }                                                                                     \
  } while (0)

void set_error(dic_engine *e, const std::string &msg) {
  std::lock_guard<std::mutex> g(e->err_mutex);
  e->err = msg;
}

inline int align_up(int v, int a) { return (v + a - 1) / a * a; }

int np_of(const dic_engine *e) { return model_nparams(e->model); }

PyrWeights pyramid_weights() {
  // pyramid_class.cpp:83-90: fp32 products of the 1-D taps, formed at run time
  const float km[5] = {0.05f, 0.25f, 0.4f, 0.25f, 0.05f};
  PyrWeights w;
  for (int dj = 0; dj < 5; ++dj)
    for (int di = 0; di < 5; ++di) {
      volatile float p = km[di] * km[dj];
      w.w[dj * 5 + di] = p;
    }
  return w;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}

// u8 image, box_w x box_h bytes, no swizzle, out-of-image elements read as 0
int encode_level_map(dic_engine *e, CUtensorMap *map, const LevelImage &li, int box_w, int box_h) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) { set_error(e, "cuTensorMapEncodeTiled is not available from this driver"); return DIC_ERROR_CUDA; }
  const cuuint64_t dims[2] = {(cuuint64_t)li.cols, (cuuint64_t)li.rows};
  const cuuint64_t strides[1] = {(cuuint64_t)li.pitch};
  const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t *>(li.ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error(e, "cuTensorMapEncodeTiled failed: " + std::to_string((int)r)); return DIC_ERROR_CUDA; }
  return DIC_OK;
}

// (Re)shape a pyramid slot for rows x cols, levels 0..stop.
int shape_slot(dic_engine *e, PyramidSlot &s, int rows, int cols, int stop, int colors = 1, cudaStream_t st = nullptr) {
  size_t total = 0;
  size_t off[kMaxLevels];
  int r = rows, c = cols;
  LevelImage lev[kMaxLevels] = {};
  for (int l = 0; l <= stop; ++l) {
    // rows are 128-byte aligned; a row that is already a multiple of 128 stays tight so that the
    // level-0 upload is ONE linear PCIe copy (49 -> 55 GB/s measured against the pitched 2-D copy).
    // Loads past `cols` inside the pitch (or into the next row) are never used by an in-image sample.
    int pitch = align_up(c * colors, 128);
    off[l] = total;
    lev[l].rows = r; lev[l].cols = c; lev[l].pitch = pitch; lev[l].colors = colors;
    total += (size_t)pitch * (r + 8);
    total = (total + 255) / 256 * 256;
    r /= 2; c /= 2;
  }
  if (total > s.cap) {
    if (s.base) CU_TRY(e, cudaFree(s.base));
    s.base = nullptr; s.cap = 0;
    CU_TRY(e, cudaMalloc(&s.base, total));
    s.cap = total;
  }
  // colour: the reference's coefficient builders read up to 5 bytes per column (see sample_def_color); padding
  // and spare rows are zero so that those reads are deterministic. Stream-ordered before the upload that follows
  // on `st` (the legacy default stream is not ordered against the engine's non-blocking streams).
  if (colors != 1) CU_TRY(e, cudaMemsetAsync(s.base, 0, total, st));
  for (int l = 0; l <= stop; ++l) {
    lev[l].ptr = s.base + off[l];
    const bool same = s.lev[l].ptr == lev[l].ptr && s.lev[l].rows == lev[l].rows && s.lev[l].cols == lev[l].cols &&
                      s.lev[l].pitch == lev[l].pitch && s.lev[l].colors == lev[l].colors;
    s.lev[l] = lev[l];
    if (colors == 1 && !same && lev[l].rows > 0 && lev[l].cols > 0) { // TMA descriptors: monochrome (tile kernel) only
      int rc = encode_level_map(e, &s.tm_patch[l], lev[l], kPatchW, kPatchH);
      if (!rc) rc = encode_level_map(e, &s.tm_tile[l], lev[l], kUndW, kTileH);
      if (!rc) rc = encode_level_map(e, &s.tm_pyr[l], lev[l], kPyrSW, kPyrSH);
      if (rc) return rc;
    }
  }
  for (int l = stop + 1; l < kMaxLevels; ++l) s.lev[l] = LevelImage{};
  s.rows = rows; s.cols = cols;
  return DIC_OK;
}

// Level 0 -> levels 1..stop on `st`. With a row band [row_begin, row_end) of level 0 only the target
// rows whose whole 5-row support lies inside the band are rebuilt at each level (the band shrinks by
// two rows per level at either end unless it touches the image border).
int build_levels(dic_engine *e, PyramidSlot &s, int stop, cudaStream_t st, int row_begin = 0, int row_end = -1) {
  static const PyrWeights kw = pyramid_weights();
  int rb = row_begin, re = row_end < 0 ? s.lev[0].rows : row_end;
  for (int l = 1; l <= stop; ++l) {
    const LevelImage &src = s.lev[l - 1];
    const LevelImage &dst = s.lev[l];
    if (dst.rows <= 0 || dst.cols <= 0) break;
    if (src.colors == 3) {
      dim3 grid((dst.cols + 127) / 128, dst.rows);
      pyramid_color_kernel<<<grid, 128, 0, st>>>(src.ptr, src.pitch, const_cast<uint8_t *>(dst.ptr), dst.rows, dst.cols,
                                                 dst.pitch, kw);
      e->launches++;
      continue;
    }
    // target row t reads source rows 2t-2 .. 2t+2
    const int tb = rb <= 0 ? 0 : (rb + 2 + 1) / 2;
    const int te = re >= src.rows ? dst.rows : std::min(dst.rows, (re - 1 - 2) / 2 + 1);
    if (te > tb) {
      static bool attr_set[16] = {false};
      if (!attr_set[e->device & 15]) {
        CU_TRY(e, cudaFuncSetAttribute(pyramid_level_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPyrSmem));
        attr_set[e->device & 15] = true;
      }
      dim3 block(kPyrTX, kPyrTY);
      const int n_tiles = ((dst.cols + kPyrTX - 1) / kPyrTX) * ((te - tb + kPyrTH - 1) / kPyrTH);
      const int grid = std::max(1, std::min(n_tiles, e->num_sms * 5)); // persistent: 5 CTAs of 38 KB fit an SM
      pyramid_level_kernel<<<grid, block, kPyrSmem, st>>>(s.tm_pyr[l - 1], const_cast<uint8_t *>(dst.ptr), dst.rows,
                                                          dst.cols, dst.pitch, kw, tb, te, &e->d_work->img_error);
      e->launches++;
    }
    rb = tb; re = std::max(tb, te);
  }
  CU_TRY(e, cudaGetLastError());
  s.valid = true;
  return DIC_OK;
}

int upload_level0(dic_engine *e, PyramidSlot &s, const void *src, int rows, int cols, int spitch,
                  bool src_on_device, cudaStream_t st) {
  uint8_t *dst = const_cast<uint8_t *>(s.lev[0].ptr);
  const cudaMemcpyKind kind = src_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  const int row_bytes = cols * s.lev[0].colors; // spitch: bytes per source row
  if (spitch == row_bytes && s.lev[0].pitch == row_bytes)
    CU_TRY(e, cudaMemcpyAsync(dst, src, (size_t)rows * row_bytes, kind, st));
  else
    CU_TRY(e, cudaMemcpy2DAsync(dst, s.lev[0].pitch, src, spitch, row_bytes, rows, kind, st));
  return DIC_OK;
}

int set_image(dic_engine *e, int role, const void *src, int rows, int cols, int spitch,
              bool on_device, cudaStream_t st) {
  PyramidSlot &s = e->pyr[e->role[role]];
  int rc = shape_slot(e, s, rows, cols, e->stop, e->colors, st);
  if (rc) return rc;
  rc = upload_level0(e, s, src, rows, cols, spitch, on_device, st);
  if (rc) return rc;
  return build_levels(e, s, e->stop, st);
}


constexpr size_t kArenaChunk = 64u << 20, kArenaMaxAlloc = 4u << 20;

// Device memory for a sector buffer: small requests come from the engine's bump arena (released
// with the engine), large ones get their own allocation.
int sector_alloc(dic_engine *e, size_t bytes, void **out, bool *owned) {
  bytes = (bytes + 255) / 256 * 256;
  if (bytes > kArenaMaxAlloc) {
    CU_TRY(e, cudaMalloc(out, bytes));
    *owned = true;
    return DIC_OK;
  }
  if (bytes > e->arena_left) {
    char *chunk = nullptr;
    CU_TRY(e, cudaMalloc(&chunk, kArenaChunk));
    e->arena_chunks.push_back(chunk);
    e->arena_cur = chunk;
    e->arena_left = kArenaChunk;
  }
  *out = e->arena_cur;
  e->arena_cur += bytes;
  e->arena_left -= bytes;
  *owned = false;
  return DIC_OK;
}

int ensure_sector_capacity(dic_engine *e, int n) {
  if (n <= e->cap_sectors) return DIC_OK;
  int cap = std::max(n, std::max(16, e->cap_sectors * 2));
  SectorDev *ds = nullptr; float *dg = nullptr; dic_result *dr = nullptr; SectorTiles *dt = nullptr;
  dic_result *hr = nullptr; float *hg = nullptr; SectorDev *hs = nullptr; SectorTiles *ht = nullptr;
  CU_TRY(e, cudaMalloc(&ds, sizeof(SectorDev) * cap));
  CU_TRY(e, cudaMalloc(&dg, sizeof(float) * kMaxParams * cap));
  CU_TRY(e, cudaMalloc(&dt, sizeof(SectorTiles) * cap));
  CU_TRY(e, cudaMemset(dt, 0, sizeof(SectorTiles) * cap));
  CU_TRY(e, cudaMalloc(&dr, sizeof(dic_result) * cap));
  CU_TRY(e, cudaHostAlloc(&hr, sizeof(dic_result) * cap, cudaHostAllocMapped | cudaHostAllocPortable));
  CU_TRY(e, cudaHostAlloc(&hg, sizeof(float) * kMaxParams * cap, cudaHostAllocMapped | cudaHostAllocPortable));
  CU_TRY(e, cudaMallocHost(&hs, sizeof(SectorDev) * cap));
  CU_TRY(e, cudaMallocHost(&ht, sizeof(SectorTiles) * cap));
  memset(hs, 0, sizeof(SectorDev) * cap);
  memset(ht, 0, sizeof(SectorTiles) * cap);
  CU_TRY(e, cudaMemset(ds, 0, sizeof(SectorDev) * cap));
  CU_TRY(e, cudaMemset(dg, 0, sizeof(float) * kMaxParams * cap));
  CU_TRY(e, cudaMemset(dr, 0, sizeof(dic_result) * cap));
  memset(hr, 0, sizeof(dic_result) * cap);
  memset(hg, 0, sizeof(float) * kMaxParams * cap);
  if (e->cap_sectors) {
    CU_TRY(e, cudaStreamSynchronize(e->stream));
    CU_TRY(e, cudaMemcpy(ds, e->d_sectors, sizeof(SectorDev) * e->cap_sectors, cudaMemcpyDeviceToDevice));
    CU_TRY(e, cudaMemcpy(dg, e->d_guess, sizeof(float) * kMaxParams * e->cap_sectors, cudaMemcpyDeviceToDevice));
    CU_TRY(e, cudaMemcpy(dt, e->d_sector_tiles, sizeof(SectorTiles) * e->cap_sectors, cudaMemcpyDeviceToDevice));
    CU_TRY(e, cudaMemcpy(dr, e->d_results, sizeof(dic_result) * e->cap_sectors, cudaMemcpyDeviceToDevice));
    memcpy(hr, e->h_results, sizeof(dic_result) * e->cap_sectors);
    memcpy(hg, e->h_guess, sizeof(float) * kMaxParams * e->cap_sectors);
    memcpy(hs, e->h_sectors, sizeof(SectorDev) * e->cap_sectors);
    memcpy(ht, e->h_sector_tiles, sizeof(SectorTiles) * e->cap_sectors);
    cudaFree(e->d_sectors); cudaFree(e->d_guess); cudaFree(e->d_results); cudaFree(e->d_sector_tiles);
    cudaFreeHost(e->h_results); cudaFreeHost(e->h_guess); cudaFreeHost(e->h_sectors); cudaFreeHost(e->h_sector_tiles);
  }
  e->h_sectors = hs; e->h_sector_tiles = ht;
  e->d_sector_tiles = dt;
  e->d_sectors = ds; e->d_guess = dg; e->d_results = dr; e->h_results = hr; e->h_guess = hg;
  void *hg_dev = nullptr;
  CU_TRY(e, cudaHostGetDevicePointer(&hg_dev, hg, 0));
  e->h_guess_dev = static_cast<float *>(hg_dev);
  void *hr_dev = nullptr;
  CU_TRY(e, cudaHostGetDevicePointer(&hr_dev, hr, 0));
  e->h_results_dev = static_cast<dic_result *>(hr_dev);
  // the memsets / copies above ran on the legacy default stream, which the engine's non-blocking streams do
  // not wait for: drain the device once (this path runs only when the sector table grows)
  CU_TRY(e, cudaDeviceSynchronize());
  e->cap_sectors = cap;
  e->sectors.resize(cap);
  return DIC_OK;
}

int ensure_counts(dic_engine *e, size_t nblocks) {
  if (nblocks + 1 <= e->cap_counts) return DIC_OK;
  if (e->d_counts) cudaFree(e->d_counts);
  if (e->d_offsets) cudaFree(e->d_offsets);
  size_t cap = nblocks + 1 + 1024;
  CU_TRY(e, cudaMalloc(&e->d_counts, sizeof(unsigned int) * cap));
  CU_TRY(e, cudaMalloc(&e->d_offsets, sizeof(unsigned long long) * cap));
  e->cap_counts = cap;
  return DIC_OK;
}

// Levels the LM loop visits (correlation_class.cpp:373) == levels that own a list
// (pyramid_class.cpp:299-301) when (stop - start) % step == 0, which dic_reset_image_pyramids enforces.
bool level_used(const dic_engine *e, int l) {
  if (l == 0) return true;
  return l >= e->start && l <= e->stop && (l - e->start) % e->step == 0 && l >= (e->start == 0 ? e->step : e->start);
}

int sector_reserve(dic_engine *e, Sector &s, size_t total) {
  if (total > s.cap) {
    if (s.buf && s.buf_owned) CU_TRY(e, cudaFree(s.buf));
    s.buf = nullptr; s.cap = 0;
    size_t n = std::max<size_t>(total, 1);
    void *p = nullptr;
    int rc = sector_alloc(e, sizeof(float2) * n, &p, &s.buf_owned);
    if (rc) return rc;
    s.buf = static_cast<float2 *>(p);
    s.cap = n;
  }
  return DIC_OK;
}

int push_sector(dic_engine *e, int id) {
  Sector &s = e->sectors[id];
  SectorDev &d = e->h_sectors[id];
  memset(&d, 0, sizeof(d));
  for (int l = 0; l < kMaxLevels; ++l) {
    d.xy[l] = s.xy[l]; d.n[l] = (int)s.n[l];
    d.n_total[l] = (int)(s.banded ? s.n_total[l] : s.n[l]);
  }
  d.cx = s.cx; d.cy = s.cy;
  SectorTiles &t = e->h_sector_tiles[id];
  memset(&t, 0, sizeof(t));
  if (s.has_tiles)
    for (int l = 0; l < kMaxLevels; ++l) t.lev[l] = s.tl[l];
  // the pinned mirrors stay valid until the sector is reset again, which first drains the stream
  CU_TRY(e, cudaMemcpyAsync(e->d_sectors + id, &d, sizeof(SectorDev), cudaMemcpyHostToDevice, e->stream));
  CU_TRY(e, cudaMemcpyAsync(e->d_sector_tiles + id, &t, sizeof(SectorTiles), cudaMemcpyHostToDevice, e->stream));
  return DIC_OK;
}

// Runs an order-preserving compaction of `ncand` candidates; first call (out == nullptr) returns
// the number kept, second call emits.
template <class Pred>
int compact_count(dic_engine *e, const Pred &pred, long ncand, long *kept) {
  size_t nblocks = (size_t)((ncand + kCompactChunk - 1) / kCompactChunk);
  if (nblocks == 0) { *kept = 0; return DIC_OK; }
  int rc = ensure_counts(e, nblocks);
  if (rc) return rc;
  compact_kernel<Pred, false><<<(unsigned)nblocks, kCompactThreads, 0, e->stream>>>(
      pred, ncand, e->d_counts, nullptr, nullptr);
  scan_counts_kernel<<<1, 1024, 0, e->stream>>>(e->d_counts, (int)nblocks, e->d_offsets);
  e->launches += 2;
  unsigned long long total = 0;
  CU_TRY(e, cudaMemcpyAsync(&total, e->d_offsets + nblocks, sizeof(total), cudaMemcpyDeviceToHost, e->stream));
  CU_TRY(e, cudaStreamSynchronize(e->stream));
  *kept = (long)total;
  return DIC_OK;
}
template <class Pred>
int compact_emit(dic_engine *e, const Pred &pred, long ncand, float2 *out) {
  size_t nblocks = (size_t)((ncand + kCompactChunk - 1) / kCompactChunk);
  if (nblocks == 0) return DIC_OK;
  compact_kernel<Pred, true><<<(unsigned)nblocks, kCompactThreads, 0, e->stream>>>(
      pred, ncand, nullptr, e->d_offsets, out);
  e->launches++;
  CU_TRY(e, cudaGetLastError());
  return DIC_OK;
}

// Derive the coarser lists of a sector from its level-0 list by the reference's successive
// decimation (pyramid_class.cpp:298-322). Level 0 must already sit at s.buf[0 .. n0).
int decimate_levels(dic_engine *e, Sector &s) {
  // pass 1: counts (needs temporary storage per level: decimate into scratch after level 0)
  // Lists shrink ~4x per level, so reserve n0/2 extra and emit level by level.
  long n0 = s.n[0];
  int prev = 0;
  size_t used = (size_t)n0;
  int first = (e->start == 0 ? e->step : e->start);
  for (int l = first; l <= e->stop; l += e->step) {
    DecimatePred pred{s.xy[prev], 1 << (l - prev)};
    long kept = 0;
    int rc = compact_count(e, pred, s.n[prev], &kept);
    if (rc) return rc;
    if (used + (size_t)kept > s.cap) {
      // grow, preserving what is already there
      size_t ncap = used + (size_t)kept + (size_t)n0 / 2 + 16;
      float2 *nb = nullptr;
      bool nb_owned = false;
      void *pv = nullptr;
      int rc2 = sector_alloc(e, sizeof(float2) * ncap, &pv, &nb_owned);
      if (rc2) return rc2;
      nb = static_cast<float2 *>(pv);
      CU_TRY(e, cudaMemcpyAsync(nb, s.buf, sizeof(float2) * used, cudaMemcpyDeviceToDevice, e->stream));
      CU_TRY(e, cudaStreamSynchronize(e->stream));
      for (int k = 0; k < kMaxLevels; ++k)
        if (s.xy[k]) s.xy[k] = nb + (s.xy[k] - s.buf);
      if (s.buf_owned) cudaFree(s.buf);
      s.buf = nb; s.cap = ncap; s.buf_owned = nb_owned;
      pred.src = s.xy[prev];
    }
    s.xy[l] = s.buf + used;
    s.n[l] = kept;
    rc = compact_emit(e, pred, s.n[prev], s.xy[l]);
    if (rc) return rc;
    used += (size_t)kept;
    prev = l;
  }
  return DIC_OK;
}

void clear_levels(Sector &s) {
  for (int l = 0; l < kMaxLevels; ++l) { s.xy[l] = nullptr; s.n[l] = 0; s.n_total[l] = 0; }
  s.banded = false;
}

int check_levels_nonempty(dic_engine *e, const Sector &s) {
  for (int l = e->stop; l >= e->start; l -= e->step)
    if ((s.banded ? s.n_total[l] : s.n[l]) <= 0) return DIC_ERROR_BAD_DOMAIN;
  return DIC_OK;
}

// pyramid_class.cpp:325-347 on a host copy of the list, in the order given.
void seq_mean(const std::vector<float2> &pts, float &cx, float &cy) {
  volatile float sx = 0.f, sy = 0.f;
  for (const float2 &q : pts) { sx = sx + q.x; sy = sy + q.y; }
  cx = sx / (float)pts.size();
  cy = sy / (float)pts.size();
}

int exact_center(dic_engine *e, Sector &s) {
  unsigned long long *d = reinterpret_cast<unsigned long long *>(e->d_scratch);
  CU_TRY(e, cudaMemsetAsync(d, 0, 16, e->stream));
  int grid = std::max(1, std::min(e->num_sms * 8, (int)((s.n[0] + 255) / 256)));
  sum_xy_kernel<<<grid, 256, 0, e->stream>>>(s.xy[0], s.n[0], d);
  e->launches++;
  long long h[2];
  CU_TRY(e, cudaMemcpyAsync(h, d, 16, cudaMemcpyDeviceToHost, e->stream));
  CU_TRY(e, cudaStreamSynchronize(e->stream));
  s.cx = (float)((double)h[0] / 1024.0 / (double)s.n[0]);
  s.cy = (float)((double)h[1] / 1024.0 / (double)s.n[0]);
  return DIC_OK;
}


// Builds the structured form of an integer-grid sector from its per-level pixel lists.
int build_tiles(dic_engine *e, Sector &s, bool may_have_duplicates) {
  s.has_tiles = false;
  for (int l = 0; l < kMaxLevels; ++l) s.tl[l] = TileLevel{};
  struct Grid { int l, gx0, gy0, ntx, nty; size_t slots; } g[kMaxLevels];
  int ng = 0;
  size_t total_slots = 0, max_slots = 0;
  int *d_box = reinterpret_cast<int *>(e->d_scratch + 768);
  for (int l = 0; l <= e->stop; ++l) {
    if (!level_used(e, l) || s.n[l] <= 0) continue;
    int box[4] = {INT_MAX, INT_MAX, INT_MIN, INT_MIN};
    CU_TRY(e, cudaMemcpyAsync(d_box, box, sizeof(box), cudaMemcpyHostToDevice, e->stream));
    int grid = (int)std::max<long>(1, std::min<long>((s.n[l] + 255) / 256, e->num_sms * 8));
    bbox_kernel<<<grid, 256, 0, e->stream>>>(s.xy[l], s.n[l], d_box);
    e->launches++;
    CU_TRY(e, cudaMemcpyAsync(box, d_box, sizeof(box), cudaMemcpyDeviceToHost, e->stream));
    CU_TRY(e, cudaStreamSynchronize(e->stream));
    if (box[0] < 0 || box[1] < 0) return DIC_OK; // negative coordinates: keep the list kernel
    Grid q;
    q.l = l;
    q.gx0 = box[0]; q.gy0 = box[1]; // tiles start at the domain's own corner: no partial first row / column
    q.ntx = (box[2] - q.gx0) / kTileW + 1; q.nty = (box[3] - q.gy0) / kTileH + 1;
    q.slots = (size_t)q.ntx * q.nty;
    total_slots += q.slots;
    max_slots = std::max(max_slots, q.slots);
    g[ng++] = q;
  }
  if (ng == 0) return DIC_OK;
  if (total_slots > s.tcap) {
    if (s.tbuf && s.tbuf_owned) cudaFree(s.tbuf);
    s.tbuf = nullptr; s.tcap = 0;
    void *pv = nullptr;
    int rc0 = sector_alloc(e, sizeof(Tile) * total_slots, &pv, &s.tbuf_owned);
    if (rc0) return rc0;
    s.tbuf = static_cast<Tile *>(pv);
    s.tcap = total_slots;
  }
  size_t need_masks = max_slots * kTileH;
  if (need_masks > e->cap_masks) {
    if (e->d_masks) cudaFree(e->d_masks);
    e->d_masks = nullptr; e->cap_masks = 0;
    CU_TRY(e, cudaMalloc(&e->d_masks, sizeof(uint32_t) * need_masks * 2));
    e->cap_masks = need_masks * 2;
  }
  const int extra_cap_per_level = may_have_duplicates ? 4096 : 0;
  if (may_have_duplicates && s.ecap < (size_t)extra_cap_per_level * kMaxLevels) {
    if (s.ebuf) cudaFree(s.ebuf);
    CU_TRY(e, cudaMalloc(&s.ebuf, sizeof(float2) * extra_cap_per_level * kMaxLevels));
    s.ecap = (size_t)extra_cap_per_level * kMaxLevels;
  }
  int *d_nextra = reinterpret_cast<int *>(e->d_scratch + 800);
  size_t used = 0;
  for (int k = 0; k < ng; ++k) {
    const Grid &q = g[k];
    CU_TRY(e, cudaMemsetAsync(e->d_masks, 0, sizeof(uint32_t) * q.slots * kTileH, e->stream));
    CU_TRY(e, cudaMemsetAsync(d_nextra, 0, sizeof(int), e->stream));
    float2 *extra = may_have_duplicates ? s.ebuf + (size_t)q.l * extra_cap_per_level : nullptr;
    tiles_scatter_kernel<<<(unsigned)((s.n[q.l] + 255) / 256), 256, 0, e->stream>>>(
        s.xy[q.l], s.n[q.l], q.gx0, q.gy0, q.nty, e->d_masks, extra, extra_cap_per_level, d_nextra);
    TilePred pred{e->d_masks, q.gx0, q.gy0, q.nty};
    size_t nblocks = (q.slots + kCompactChunk - 1) / kCompactChunk;
    int rc = ensure_counts(e, nblocks);
    if (rc) return rc;
    compact_any_kernel<TilePred, Tile, false><<<(unsigned)nblocks, kCompactThreads, 0, e->stream>>>(
        pred, (long)q.slots, e->d_counts, nullptr, nullptr);
    scan_counts_kernel<<<1, 1024, 0, e->stream>>>(e->d_counts, (int)nblocks, e->d_offsets);
    Tile *dst = s.tbuf + used;
    compact_any_kernel<TilePred, Tile, true><<<(unsigned)nblocks, kCompactThreads, 0, e->stream>>>(
        pred, (long)q.slots, nullptr, e->d_offsets, dst);
    e->launches += 4;
    unsigned long long ntiles = 0;
    int nextra = 0;
    CU_TRY(e, cudaMemcpyAsync(&ntiles, e->d_offsets + nblocks, sizeof(ntiles), cudaMemcpyDeviceToHost, e->stream));
    CU_TRY(e, cudaMemcpyAsync(&nextra, d_nextra, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    CU_TRY(e, cudaStreamSynchronize(e->stream));
    if (nextra > extra_cap_per_level) return DIC_OK; // too many duplicates: keep the list kernel
    s.tl[q.l].tiles = dst;
    s.tl[q.l].n_tiles = (int)ntiles;
    s.tl[q.l].extra = extra;
    s.tl[q.l].n_extra = nextra;
    used += (size_t)ntiles;
  }
  CU_TRY(e, cudaGetLastError());
  s.has_tiles = true;
  return DIC_OK;
}

SolveSettings solve_settings(const dic_engine *e) {
  SolveSettings cfg;
  memset(&cfg, 0, sizeof(cfg));
  const PyramidSlot &u = e->pyr[e->role[0]], &d = e->pyr[e->role[1]];
  for (int l = 0; l < kMaxLevels; ++l) { cfg.und[l] = u.lev[l]; cfg.def[l] = d.lev[l]; }
  cfg.start = e->start; cfg.step = e->step; cfg.stop = e->stop;
  cfg.max_iters = e->max_iters; cfg.precision = e->precision;
  return cfg;
}

// the guess of a single-sector launch rides in the kernel parameters (no copy-engine work per correlate)
GuessParam guess_param(const dic_engine *e, int first) {
  GuessParam g;
  memcpy(g.v, e->h_guess + (size_t)first * kMaxParams, sizeof(g.v));
  return g;
}

template <int MODEL, int INTERP, int MODE>
int launch_solve(dic_engine *e, bool grid_mode, int first, int count) {
  SolveSettings cfg = solve_settings(e);
  const SectorDev *sectors = e->d_sectors;
  const float *guesses = e->h_guess_dev;
  GuessParam g0 = guess_param(e, first);
  dic_result *results = e->h_results_dev; // zero-copy: 176 B per sector, written once when its pyramid ends
  GridWork *work = e->d_work;
  if (grid_mode) {
    auto kern = gn_solve_kernel<MODEL, INTERP, MODE, true>;
    static int per_sm_cached[16] = {0};
    int &per_sm = per_sm_cached[e->device & 15];
    if (per_sm == 0) CU_TRY(e, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, 0));
    if (per_sm < 1) { set_error(e, "gn_solve_kernel does not fit on an SM"); return DIC_ERROR_CUDA; }
    // enough CTAs for ~4 pixels per thread at the finest level, never more than co-resident
    long n0 = e->sectors[first].n[e->start];
    int want = (int)std::min<long>((n0 + kThreads * 4 - 1) / (kThreads * 4), (long)per_sm * e->num_sms);
    int grid = std::max(1, std::min(want, e->max_grid));
    int one = 1;
    void *args[] = {&cfg, &sectors, &guesses, &g0, &results, &first, &one, &work};
    CU_TRY(e, cudaLaunchCooperativeKernel((void *)kern, dim3(grid), dim3(kThreads), args, 0, e->stream));
  } else {
    auto kern = gn_solve_kernel<MODEL, INTERP, MODE, false>;
    static int per_sm_cached[16] = {0};
    int &per_sm = per_sm_cached[e->device & 15];
    if (per_sm == 0) CU_TRY(e, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, 0));
    int grid = std::max(1, std::min(count, std::max(1, per_sm) * e->num_sms));
    kern<<<grid, kThreads, 0, e->stream>>>(cfg, sectors, guesses, g0, results, first, count, work);
    CU_TRY(e, cudaGetLastError());
  }
  e->launches++;
  return DIC_OK;
}

template <int MODEL, int MODE>
int launch_solve_tiles(dic_engine *e, bool grid_mode, int first, int count) {
  SolveSettings cfg = solve_settings(e);
  const PyramidSlot &u = e->pyr[e->role[0]], &d = e->pyr[e->role[1]];
  const SectorDev *sectors = e->d_sectors;
  const SectorTiles *stiles = e->d_sector_tiles;
  const float *guesses = e->h_guess_dev;
  GuessParam g0 = guess_param(e, first);
  dic_result *results = e->h_results_dev; // zero-copy: 176 B per sector, written once when its pyramid ends
  GridWork *work = e->d_work;
  TileMaps maps;
  memset(&maps, 0, sizeof(maps));
  for (int l = 0; l <= e->stop; ++l) { maps.def[l] = d.tm_patch[l]; maps.und[l] = u.tm_tile[l]; }
  constexpr int NACC = Acc<model_nparams(MODEL)>::kN;
  constexpr int NTB = tile_cta_threads(MODEL, MODE, false), NTG = tile_cta_threads(MODEL, MODE, true); // threads per CTA of the batch / grid form
  const size_t smem = tiles_dyn_smem(NACC, grid_mode ? NTG : NTB);
  if (grid_mode) {
    auto kern = gn_solve_tiles_kernel<MODEL, MODE, true, 1>;
    static int per_sm_cached[16] = {0}; // per device: attribute + occupancy queried once, not per launch
    int &per_sm = per_sm_cached[e->device & 15];
    if (per_sm == 0) {
      CU_TRY(e, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      CU_TRY(e, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
      CU_TRY(e, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NTG, smem));
    }
    if (per_sm < 1) { set_error(e, "gn_solve_tiles_kernel does not fit on an SM"); return DIC_ERROR_CUDA; }
    // at least four quads (16 rows of a strip) per warp at the finest level, never more CTAs than are co-resident
    long nt = (long)e->sectors[first].tl[e->start].n_tiles * kQuadsPerTile / 4;
    int want = (int)std::min<long>((nt + NTG / 32 - 1) / (NTG / 32), (long)per_sm * e->num_sms);
    int grid = std::max(1, std::min(want, e->max_grid));
    int one = 1;
    void *args[] = {&cfg, &maps, &sectors, &stiles, &guesses, &g0, &results, &first, &one, &work};
    CU_TRY(e, cudaLaunchCooperativeKernel((void *)kern, dim3(grid), dim3(NTG), args, smem, e->stream));
  } else {
    auto kern1 = gn_solve_tiles_kernel<MODEL, MODE, false, 1>;
    auto kern2 = gn_solve_tiles_kernel<MODEL, MODE, false, 2>;
    static int per_sm_cached[16] = {0};
    int &per_sm = per_sm_cached[e->device & 15];
    if (per_sm == 0) {
      CU_TRY(e, cudaFuncSetAttribute(kern1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      CU_TRY(e, cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      // all of the SM's unified L1 / shared memory as shared memory: the staging buffers (35 KB per CTA) decide how
      // many CTAs are resident, the kernel's global reads are tile records and parameters only
      CU_TRY(e, cudaFuncSetAttribute(kern1, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
      CU_TRY(e, cudaFuncSetAttribute(kern2, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
      CU_TRY(e, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern1, NTB, smem));
    }
    const long slots = (long)std::max(1, per_sm) * e->num_sms;
    // One CTA pair per sector doubles the warps that work on a sector. Measured on c4 (B200, 128-thread CTAs, 4 per
    // SM): a pair costs ~20 % per sector (cluster barrier + DSMEM exchange per evaluation, co-scheduling of the two
    // CTAs inside one GPC), and a CTA does not run faster when its SM is only partly filled (a warp of this kernel
    // is bound by its own dependent fp32 chains), so splitting sectors never shortens a launch that already fills
    // the CTA slots once: 512 subsets per GPU take 0.43 ms as 1024 half-sector CTAs in 1.73 waves and 0.37 ms as
    // 512 whole-sector CTAs in one wave. Pairs pay only when there are too few sectors to occupy the slots at all.
    bool pair = e->cluster_mode == 2 || (e->cluster_mode == 0 && 2L * count <= slots);
    e->last_cluster = pair ? 2 : 1;
    if (pair) {
      cudaLaunchConfig_t lc;
      memset(&lc, 0, sizeof(lc));
      lc.gridDim = dim3((unsigned)std::max(2L, std::min(2L * count, slots / 2 * 2)));
      lc.blockDim = dim3(NTB);
      lc.dynamicSmemBytes = smem;
      lc.stream = e->stream;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      lc.attrs = at; lc.numAttrs = 1;
      CU_TRY(e, cudaLaunchKernelEx(&lc, kern2, cfg, maps, sectors, stiles, guesses, g0, results, first, count, work));
    } else {
      // Resident CTAs with a ticket queue never leave before the launch ends, and they fill the register file: a
      // kernel of another stream (the next pair's pyramid build) waits for the whole solve. One CTA per sector lets
      // the block scheduler interleave it (the ticket of such a CTA is past the end: it leaves after its sector).
      // Measured on c4 (tools/probe_pipe.py): with two pairs staged ahead the step is bound by the PCIe transfer
      // either way (2.45 ms), and the resident kernel is within 1.5 % in both forms: the ticket queue stays the default.
      const bool per_sector = e->batch_queue == 2;
      int grid = (int)std::max(1L, per_sector ? (long)count : std::min((long)count, slots));
      kern1<<<grid, NTB, smem, e->stream>>>(cfg, maps, sectors, stiles, guesses, g0, results, first, count, work);
    }
    CU_TRY(e, cudaGetLastError());
  }
  e->launches++;
  return DIC_OK;
}

bool tiles_applicable(const dic_engine *e, int first, int count) {
  if (e->kernel_variant == 1) return false;
  if (e->colors != 1) return false; // colour images: generic pixel-list kernel
  if (e->interp != DIC_IM_BICUBIC) return false;
  if (e->model != DIC_FM_UVUxUyVxVy && e->model != DIC_FM_QUADRATIC) return false;
  for (int i = 0; i < count; ++i)
    if (!e->sectors[first + i].has_tiles) return false;
  return true;
}

int launch_solve_tiles_any(dic_engine *e, bool grid_mode, int first, int count) {
  if (e->model == DIC_FM_UVUxUyVxVy)
    return e->mode == DIC_MODE_FAST ? launch_solve_tiles<DIC_FM_UVUxUyVxVy, DIC_MODE_FAST>(e, grid_mode, first, count)
                                    : launch_solve_tiles<DIC_FM_UVUxUyVxVy, DIC_MODE_PARITY>(e, grid_mode, first, count);
  return e->mode == DIC_MODE_FAST ? launch_solve_tiles<DIC_FM_QUADRATIC, DIC_MODE_FAST>(e, grid_mode, first, count)
                                  : launch_solve_tiles<DIC_FM_QUADRATIC, DIC_MODE_PARITY>(e, grid_mode, first, count);
}

template <int MODEL, int INTERP>
int launch_solve_mode(dic_engine *e, bool grid_mode, int first, int count) {
  if (e->mode == DIC_MODE_FAST) return launch_solve<MODEL, INTERP, DIC_MODE_FAST>(e, grid_mode, first, count);
  return launch_solve<MODEL, INTERP, DIC_MODE_PARITY>(e, grid_mode, first, count);
}
template <int MODEL>
int launch_solve_interp(dic_engine *e, bool grid_mode, int first, int count) {
  switch (e->interp) {
  case DIC_IM_NEAREST: return launch_solve<MODEL, DIC_IM_NEAREST, DIC_MODE_PARITY>(e, grid_mode, first, count);
  case DIC_IM_BILINEAR: return launch_solve<MODEL, DIC_IM_BILINEAR, DIC_MODE_PARITY>(e, grid_mode, first, count);
  default: return launch_solve_mode<MODEL, DIC_IM_BICUBIC>(e, grid_mode, first, count);
  }
}
int launch_solve_any(dic_engine *e, bool grid_mode, int first, int count) {
  switch (e->model) {
  case DIC_FM_U: return launch_solve_interp<DIC_FM_U>(e, grid_mode, first, count);
  case DIC_FM_UV: return launch_solve_interp<DIC_FM_UV>(e, grid_mode, first, count);
  case DIC_FM_UVQ: return launch_solve_interp<DIC_FM_UVQ>(e, grid_mode, first, count);
  case DIC_FM_UVUxUyVxVy: return launch_solve_interp<DIC_FM_UVUxUyVxVy>(e, grid_mode, first, count);
  default: return launch_solve_interp<DIC_FM_QUADRATIC>(e, grid_mode, first, count);
  }
}

template <int MODEL, int INTERP, int MODE>
int launch_eval(dic_engine *e, int id, int level, const float *d_params, int grid) {
  SolveSettings cfg = solve_settings(e);
  gn_eval_kernel<MODEL, INTERP, MODE><<<grid, kThreads, 0, e->stream>>>(cfg, e->d_sectors + id, level,
                                                                      d_params, e->d_partials);
  e->launches++;
  CU_TRY(e, cudaGetLastError());
  return DIC_OK;
}
template <int MODEL>
int launch_eval_model(dic_engine *e, int id, int level, const float *d_params, int grid) {
  if (e->interp == DIC_IM_NEAREST) return launch_eval<MODEL, DIC_IM_NEAREST, DIC_MODE_PARITY>(e, id, level, d_params, grid);
  if (e->interp == DIC_IM_BILINEAR) return launch_eval<MODEL, DIC_IM_BILINEAR, DIC_MODE_PARITY>(e, id, level, d_params, grid);
  if (e->mode == DIC_MODE_FAST) return launch_eval<MODEL, DIC_IM_BICUBIC, DIC_MODE_FAST>(e, id, level, d_params, grid);
  return launch_eval<MODEL, DIC_IM_BICUBIC, DIC_MODE_PARITY>(e, id, level, d_params, grid);
}

bool sector_ok(const dic_engine *e, int id) {
  return id >= 0 && id < e->cap_sectors && e->sectors[id].kind != SK_NONE;
}

int images_ready(dic_engine *e) {
  const PyramidSlot &u = e->pyr[e->role[0]], &d = e->pyr[e->role[1]];
  if (!u.valid || !d.valid) { set_error(e, "image pyramids not set"); return DIC_ERROR_BAD_ARGUMENT; }
  return DIC_OK;
}

} // namespace

// =============================================================================== C ABI

extern "C" {

int dic_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

dic_engine *dic_create(int device) {
  int n = dic_device_count();
  if (device < 0 || device >= n) return nullptr;
  if (cudaSetDevice(device) != cudaSuccess) return nullptr;
  dic_engine *e = new dic_engine;
  e->device = device;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete e; return nullptr; }
  e->num_sms = prop.multiProcessorCount;
  // The image stream outranks the correlation stream: when a batch solve runs one CTA per sector, the block scheduler
  // hands freed CTA slots to the next pair's pyramid build first, so that build hides inside the solve instead of
  // queueing behind it (dic_set_batch_queue).
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  if (std::getenv("DIC_NO_STREAM_PRIORITY")) prio_hi = prio_lo; // diagnostics: both streams at the default priority
  bool ok = cudaStreamCreateWithPriority(&e->stream, cudaStreamNonBlocking, prio_lo) == cudaSuccess &&
            cudaStreamCreateWithPriority(&e->img_stream, cudaStreamNonBlocking, prio_hi) == cudaSuccess &&
            cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreateWithFlags(&e->ev_copy, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&e->ev_copy_und, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreate(&e->ev0) == cudaSuccess && cudaEventCreate(&e->ev1) == cudaSuccess &&
            cudaEventCreate(&e->ev_step0) == cudaSuccess && cudaEventCreate(&e->ev_step1) == cudaSuccess &&
            cudaEventCreateWithFlags(&e->ev_img, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&e->ev_gn, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&e->ev_rot, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&e->ev_ready[0], cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&e->ev_ready[1], cudaEventDisableTiming) == cudaSuccess;
  e->max_grid = e->num_sms * 8;
  ok = ok && cudaMalloc(&e->d_work, sizeof(GridWork)) == cudaSuccess &&
       cudaMemset(e->d_work, 0, sizeof(GridWork)) == cudaSuccess &&
       cudaMalloc(&e->d_partials, sizeof(float) * kAccStride * (size_t)e->max_grid) == cudaSuccess &&
       cudaMalloc(&e->d_scratch, 4096) == cudaSuccess;
  if (!ok) { cudaGetLastError(); dic_destroy(e); return nullptr; }
  return e;
}

void dic_destroy(dic_engine *e) {
  if (!e) return;
  cudaSetDevice(e->device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  if (e->img_stream) cudaStreamSynchronize(e->img_stream);
  for (auto &s : e->sectors) {
    if (s.buf && s.buf_owned) cudaFree(s.buf);
    if (s.tbuf && s.tbuf_owned) cudaFree(s.tbuf);
    if (s.ebuf) cudaFree(s.ebuf);
  }
  for (char *c : e->arena_chunks) cudaFree(c);
  for (auto &g : e->grid_blocks) { cudaFree(g.lists); cudaFree(g.tiles); cudaFree(g.desc); }
  if (e->h_sectors) cudaFreeHost(e->h_sectors);
  if (e->h_sector_tiles) cudaFreeHost(e->h_sector_tiles);
  cudaFree(e->d_sector_tiles); cudaFree(e->d_masks); cudaFree(e->d_mailbox);
  for (auto &p : e->pyr) if (p.base) cudaFree(p.base);
  cudaFree(e->d_sectors); cudaFree(e->d_guess); cudaFree(e->d_results);
  if (e->h_results) cudaFreeHost(e->h_results);
  if (e->h_guess) cudaFreeHost(e->h_guess);
  cudaFree(e->d_work); cudaFree(e->d_partials); cudaFree(e->d_scratch);
  cudaFree(e->d_counts); cudaFree(e->d_offsets); cudaFree(e->d_stage);
  if (e->ev_step0) cudaEventDestroy(e->ev_step0);
  if (e->ev_step1) cudaEventDestroy(e->ev_step1);
  if (e->ev0) cudaEventDestroy(e->ev0);
  if (e->ev1) cudaEventDestroy(e->ev1);
  if (e->ev_img) cudaEventDestroy(e->ev_img);
  if (e->trace_base) {
    cudaEventDestroy(e->trace_base);
    for (auto &r : e->trace_st) for (auto &ev : r) cudaEventDestroy(ev);
    for (auto &r : e->trace_so) for (auto &ev : r) cudaEventDestroy(ev);
  }
  if (e->ev_gn) cudaEventDestroy(e->ev_gn);
  if (e->ev_rot) cudaEventDestroy(e->ev_rot);
  for (auto &ev : e->ev_ready) if (ev) cudaEventDestroy(ev);
  if (e->ev_copy) cudaEventDestroy(e->ev_copy);
  if (e->ev_copy_und) cudaEventDestroy(e->ev_copy_und);
  if (e->copy_stream) { cudaStreamSynchronize(e->copy_stream); cudaStreamDestroy(e->copy_stream); }
  if (e->stream) cudaStreamDestroy(e->stream);
  if (e->img_stream) cudaStreamDestroy(e->img_stream);
  delete e;
}

const char *dic_last_error(const dic_engine *e) { return e ? e->err.c_str() : "null engine"; }

int dic_set_max_iters(dic_engine *e, int v) { if (!e) return DIC_ERROR_BAD_ARGUMENT; e->max_iters = v; return DIC_OK; }
int dic_set_precision(dic_engine *e, float v) { if (!e) return DIC_ERROR_BAD_ARGUMENT; e->precision = v; return DIC_OK; }
int dic_set_fitting_model(dic_engine *e, int m) {
  if (!e || m < DIC_FM_U || m > DIC_FM_QUADRATIC) return DIC_ERROR_BAD_ARGUMENT;
  e->model = m; return DIC_OK;
}
int dic_set_interpolation_model(dic_engine *e, int m) {
  if (!e || m < DIC_IM_NEAREST || m > DIC_IM_BICUBIC) return DIC_ERROR_BAD_ARGUMENT;
  e->interp = m; return DIC_OK;
}
int dic_set_arith_mode(dic_engine *e, int m) {
  if (!e || (m != DIC_MODE_PARITY && m != DIC_MODE_FAST)) return DIC_ERROR_BAD_ARGUMENT;
  e->mode = m; return DIC_OK;
}
int dic_set_kernel_variant(dic_engine *e, int v) {
  if (!e || v < 0 || v > 2) return DIC_ERROR_BAD_ARGUMENT;
  e->kernel_variant = v; return DIC_OK;
}
int dic_set_center_mode(dic_engine *e, int m) {
  if (!e || (m != DIC_CENTER_REFERENCE && m != DIC_CENTER_EXACT)) return DIC_ERROR_BAD_ARGUMENT;
  e->center_mode = m; return DIC_OK;
}

static int set_pyramid_range(dic_engine *e, int start, int step, int stop) {
  if (step < 1) step = 1;
  if (start < 0 || stop < start || stop >= kMaxLevels || (stop - start) % step != 0) {
    set_error(e, "pyramid start/step/stop must satisfy 0 <= start <= stop < 8 and (stop-start) % step == 0");
    return DIC_ERROR_BAD_ARGUMENT;
  }
  if (e->start != start || e->step != step || e->stop != stop) {
    // level layout changes: every sector list has to be rebuilt by the caller (the reference GUI
    // re-runs resetPolygon after a pyramid change as well, mainapp.cpp:900-912)
    for (auto &s : e->sectors) { s.kind = SK_NONE; clear_levels(s); }
    for (auto &p : e->pyr) p.valid = false;
    e->staged = 0;
  }
  e->start = start; e->step = step; e->stop = stop;
  return DIC_OK;
}

static int reset_pyramids_common(dic_engine *e, const void *und, const void *def, const void *nxt,
                                 int rows, int cols, int pitch, bool on_device, int start, int step,
                                 int stop, int colors = 1) {
  if (!e || !und || !def || rows < 8 || cols < 8) return DIC_ERROR_BAD_ARGUMENT;
  cudaSetDevice(e->device);
  NvtxRange nvtx("dic_reset_image_pyramids");
  int rc = set_pyramid_range(e, start, step, stop);
  if (rc) return rc;
  if (colors == 3 && (cols % (1 << stop)) != 0) {
    // pyramid_class.cpp:93-94 halves the BYTE step of a colour row (targetStep = sourceStep / 2): for an odd width
    // that is not the width of the next level and the reference's own levels shear; refused instead of mimicked
    set_error(e, "colour images need a width that stays even down to the coarsest pyramid level");
    return DIC_ERROR_BAD_ARGUMENT;
  }
  e->colors = colors;
  if ((rc = set_image(e, 0, und, rows, cols, pitch, on_device, e->stream))) return rc;
  if ((rc = set_image(e, 1, def, rows, cols, pitch, on_device, e->stream))) return rc;
  if (nxt && (rc = set_image(e, 2, nxt, rows, cols, pitch, on_device, e->stream))) return rc;
  if (!on_device) CU_TRY(e, cudaStreamSynchronize(e->stream)); // caller may reuse its buffers
  return DIC_OK;
}

int dic_reset_image_pyramids(dic_engine *e, const uint8_t *und, const uint8_t *def, const uint8_t *nxt,
                             int rows, int cols, int channels, int start, int step, int stop) {
  if (channels != 1 && channels != 3) { if (e) set_error(e, "images have 1 or 3 interleaved 8-bit channels"); return DIC_ERROR_BAD_ARGUMENT; }
  return reset_pyramids_common(e, und, def, nxt, rows, cols, cols * channels, false, start, step, stop, channels);
}
int dic_reset_image_pyramids_device(dic_engine *e, const void *und, const void *def, const void *nxt,
                                    int rows, int cols, int pitch, int start, int step, int stop) {
  // the source images were produced on streams this library knows nothing about (e.g. the framework's current
  // stream) and the engine's own streams are non-blocking: drain the device once so that the copies below cannot
  // overtake their producer. A set-up call, not a per-frame one.
  if (e) { cudaSetDevice(e->device); cudaDeviceSynchronize(); }
  return reset_pyramids_common(e, und, def, nxt, rows, cols, pitch, true, start, step, stop);
}

int dic_reset_next_pyramid(dic_engine *e, const uint8_t *nxt, int rows, int cols) {
  if (!e || !nxt) return DIC_ERROR_BAD_ARGUMENT;
  cudaSetDevice(e->device);
  NvtxRange nvtx("dic_reset_next_pyramid");
  // after a rotation the nxt slot is the previous und / def image, which solves enqueued BEFORE the rotation may
  // still be reading: the upload waits for the correlation stream's position at that rotation (not for the
  // current frame's solve, which is what this call is meant to overlap with)
  CU_TRY(e, cudaStreamWaitEvent(e->img_stream, e->ev_rot, 0));
  int rc = set_image(e, 2, nxt, rows, cols, cols * e->colors, false, e->img_stream);
  if (rc) return rc;
  CU_TRY(e, cudaStreamSynchronize(e->img_stream));
  return DIC_OK;
}
// Enqueue-only form: the H2D copy and the pyramid build of the next image are queued on the image stream and the
// call returns; dic_make_def_pyramid_from_nxt orders the correlation stream behind them. The caller keeps `nxt`
// alive and unchanged until then (pinned memory makes the copy truly asynchronous). Lets a frame loop prefetch
// frame k + 2 without a loader thread (the reference spawns one per frame, manager_class.cpp:1438-1447).
int dic_reset_next_pyramid_async(dic_engine *e, const uint8_t *nxt, int rows, int cols) {
  if (!e || !nxt) return DIC_ERROR_BAD_ARGUMENT;
  cudaSetDevice(e->device);
  NvtxRange nvtx("dic_reset_next_pyramid_async");
  CU_TRY(e, cudaStreamWaitEvent(e->img_stream, e->ev_rot, 0)); // see dic_reset_next_pyramid
  return set_image(e, 2, nxt, rows, cols, cols * e->colors, false, e->img_stream);
}
int dic_reset_next_pyramid_device(dic_engine *e, const void *nxt, int rows, int cols, int pitch) {
  if (!e || !nxt || e->colors != 1) return DIC_ERROR_BAD_ARGUMENT;
  cudaSetDevice(e->device);
  CU_TRY(e, cudaStreamWaitEvent(e->img_stream, e->ev_rot, 0)); // see dic_reset_next_pyramid
  int rc = set_image(e, 2, nxt, rows, cols, pitch, true, e->img_stream);
  if (rc) return rc;
  CU_TRY(e, cudaEventRecord(e->ev_img, e->img_stream));
  CU_TRY(e, cudaStreamWaitEvent(e->stream, e->ev_img, 0));
  return DIC_OK;
}
int dic_reset_def_pyramid(dic_engine *e, const uint8_t *def, int rows, int cols) {
  if (!e || !def) return DIC_ERROR_BAD_ARGUMENT;
  cudaSetDevice(e->device);
  int rc = set_image(e, 1, def, rows, cols, cols * e->colors, false, e->stream);
  if (rc) return rc;
  CU_TRY(e, cudaStreamSynchronize(e->stream));
  return DIC_OK;
}
int dic_reset_def_pyramid_device(dic_engine *e, const void *def, int rows, int cols, int pitch) {
  if (!e || !def || e->colors != 1) return DIC_ERROR_BAD_ARGUMENT;
  cudaSetDevice(e->device);
  return set_image(e, 1, def, rows, cols, pitch, true, e->stream);
}

int dic_make_und_pyramid_from_def(dic_engine *e) {
  if (!e) return DIC_ERROR_BAD_ARGUMENT;
  // pyramid_class.cpp:211-226: und takes def's images, def becomes empty
  cudaSetDevice(e->device);
  std::swap(e->role[0], e->role[1]);
  e->pyr[e->role[1]].valid = false;
  CU_TRY(e, cudaEventRecord(e->ev_rot, e->stream));
  return DIC_OK;
}
int dic_make_def_pyramid_from_nxt(dic_engine *e) {
  if (!e) return DIC_ERROR_BAD_ARGUMENT;
  cudaSetDevice(e->device);
  // pyramid_class.cpp:228-243
  CU_TRY(e, cudaEventRecord(e->ev_img, e->img_stream));
  CU_TRY(e, cudaStreamWaitEvent(e->stream, e->ev_img, 0));
  std::swap(e->role[1], e->role[2]);
  e->pyr[e->role[2]].valid = false;
  CU_TRY(e, cudaEventRecord(e->ev_rot, e->stream)); // everything that still reads the slot that is now `nxt`
  return DIC_OK;
}

// Double-buffered ingest of whole image pairs: the upload and the pyramid build of pair k + 1 run on
// the image stream while the solve of pair k runs on the correlation stream.
static int stage_pair_rows(dic_engine *e, const uint8_t *und, const uint8_t *def, int rows, int cols,
                           int row_begin, int row_end) {
  if (!e || !und || !def || rows < 8 || cols < 8) return DIC_ERROR_BAD_ARGUMENT;
  if (e->colors != 1) { set_error(e, "staged image pairs are monochrome"); return DIC_ERROR_BAD_ARGUMENT; }
  NvtxRange nvtx("dic_stage_next_pair");
  row_begin = std::max(0, row_begin); row_end = std::min(rows, row_end);
  if (row_end <= row_begin) return DIC_ERROR_BAD_ARGUMENT;
  cudaSetDevice(e->device);
  if (e->staged >= 2) {
    set_error(e, "two pairs are staged already: dic_advance_pair first");
    return DIC_ERROR_BAD_ARGUMENT;
  }
  const int pos = e->staged; // 0: roles 3 / 4, 1: roles 5 / 6
  // The staging slots were the und / def slots until the last dic_advance_pair and may still be read by the solves
  // enqueued before it: ev_gn was recorded there, in front of that call's wait for the image stream. (Recording it
  // here instead made this pair's transfer wait for the previous pair's pyramid build as well, through the
  // correlation stream: 0.11 ms of idle bus per c4 step.) An event never recorded is a no-op to wait for.
  CU_TRY(e, cudaStreamWaitEvent(e->copy_stream, e->ev_gn, 0));
  const int ti = e->trace_on && e->trace_stage < dic_engine::kTrace ? e->trace_stage++ : -1;
  if (ti >= 0) cudaEventRecord(e->trace_st[ti][0], e->copy_stream);
  // transfers on the copy stream, pyramid kernels on the image stream: the next pair's transfer does
  // not queue behind this pair's pyramid build
  int rc;
  for (int k = 0; k < 2; ++k) {
    PyramidSlot &s = e->pyr[e->role[3 + 2 * pos + k]];
    if ((rc = shape_slot(e, s, rows, cols, e->stop))) return rc;
    const uint8_t *src = (k == 0 ? und : def) + (size_t)row_begin * cols;
    uint8_t *dst = const_cast<uint8_t *>(s.lev[0].ptr) + (size_t)row_begin * s.lev[0].pitch;
    const int nr = row_end - row_begin;
    if (s.lev[0].pitch == cols)
      CU_TRY(e, cudaMemcpyAsync(dst, src, (size_t)nr * cols, cudaMemcpyHostToDevice, e->copy_stream));
    else
      CU_TRY(e, cudaMemcpy2DAsync(dst, s.lev[0].pitch, src, cols, cols, nr, cudaMemcpyHostToDevice, e->copy_stream));
    // one event per image: the reference image's pyramid is built while the deformed image is still on the bus
    CU_TRY(e, cudaEventRecord(k == 0 ? e->ev_copy_und : e->ev_copy, e->copy_stream));
    if (ti >= 0) cudaEventRecord(e->trace_st[ti][1 + k], e->copy_stream);
  }
  for (int k = 0; k < 2; ++k) {
    CU_TRY(e, cudaStreamWaitEvent(e->img_stream, k == 0 ? e->ev_copy_und : e->ev_copy, 0));
    if ((rc = build_levels(e, e->pyr[e->role[3 + 2 * pos + k]], e->stop, e->img_stream, row_begin, row_end))) return rc;
  }
  CU_TRY(e, cudaEventRecord(e->ev_ready[pos], e->img_stream));
  e->staged = pos + 1;
  if (ti >= 0) cudaEventRecord(e->trace_st[ti][3], e->img_stream);
  return DIC_OK;
}
int dic_stage_next_pair(dic_engine *e, const uint8_t *und, const uint8_t *def, int rows, int cols) {
  return stage_pair_rows(e, und, def, rows, cols, 0, rows);
}
int dic_stage_next_pair_rows(dic_engine *e, const uint8_t *und, const uint8_t *def, int rows, int cols,
                             int row_begin, int row_end) {
  return stage_pair_rows(e, und, def, rows, cols, row_begin, row_end);
}
int dic_advance_pair(dic_engine *e) {
  if (!e) return DIC_ERROR_BAD_ARGUMENT;
  cudaSetDevice(e->device);
  if (e->staged < 1 || !e->pyr[e->role[3]].valid || !e->pyr[e->role[4]].valid) {
    set_error(e, "dic_advance_pair without a staged pair");
    return DIC_ERROR_BAD_ARGUMENT;
  }
  CU_TRY(e, cudaEventRecord(e->ev_gn, e->stream)); // every solve that reads the slots about to become staging slots
  // the solves that follow wait for THIS pair's pyramids only, not for a second staged pair still on the bus
  CU_TRY(e, cudaStreamWaitEvent(e->stream, e->ev_ready[0], 0));
  std::swap(e->role[0], e->role[3]);
  std::swap(e->role[1], e->role[4]);
  e->pyr[e->role[3]].valid = false;
  e->pyr[e->role[4]].valid = false;
  if (e->staged == 2) { // the second staged pair moves up; the released slots become the far staging pair
    std::swap(e->role[3], e->role[5]);
    std::swap(e->role[4], e->role[6]);
    std::swap(e->ev_ready[0], e->ev_ready[1]);
  }
  e->staged--;
  return DIC_OK;
}

// ------------------------------------------------------------------ domains

static int download_list(dic_engine *e, const float2 *d, long n, std::vector<float2> &h);
static void to_reference_order(const Sector &s, std::vector<float2> &h);

static int begin_sector(dic_engine *e, int id, Sector **out) {
  if (!e || id < 0 || id > (1 << 24)) return DIC_ERROR_BAD_ARGUMENT;
  cudaSetDevice(e->device);
  int rc = ensure_sector_capacity(e, id + 1);
  if (rc) return rc;
  Sector &s = e->sectors[id];
  if (s.kind != SK_NONE) CU_TRY(e, cudaStreamSynchronize(e->stream)); // its pinned mirror may be in flight
  s.kind = SK_NONE;
  s.pending = false;
  s.has_tiles = false;
  clear_levels(s);
  *out = &s;
  return DIC_OK;
}

static int reset_rect_impl(dic_engine *e, int id, int x0, int y0, int x1, int y1, int band_y0, int band_y1) {
  Sector *sp;
  int rc = begin_sector(e, id, &sp);
  if (rc) return rc;
  Sector &s = *sp;
  if (x1 < x0 || y1 < y0) return DIC_ERROR_BAD_DOMAIN;
  // per level: multiples of 2^l inside [x0, x1] x [y0, y1] (successive decimation of integer
  // points by pyramid_class.cpp:306-316 is exactly that)
  struct Lv { int l, xs, ys, nx, ny, mag; } lv[kMaxLevels];
  int nl = 0;
  size_t total = 0;
  for (int l = 0; l <= e->stop; ++l) {
    if (!level_used(e, l)) continue;
    int mag = 1 << l;
    auto first_mult = [mag](int a) { int q = a / mag; if (q * mag < a) ++q; return q * mag; };
    auto last_mult = [mag](int a) { int q = a / mag; if (q * mag > a) --q; return q * mag; };
    int xs = first_mult(x0), xe = last_mult(x1), ys = first_mult(y0), ye = last_mult(y1);
    int nx = xe >= xs ? (xe - xs) / mag + 1 : 0, ny = ye >= ys ? (ye - ys) / mag + 1 : 0;
    s.n_total[l] = (long)nx * ny;
    // this GPU's band of rows [band_y0, band_y1] (the whole rectangle unless row-split)
    ys = first_mult(std::max(y0, band_y0)); ye = last_mult(std::min(y1, band_y1));
    ny = ye >= ys ? (ye - ys) / mag + 1 : 0;
    lv[nl++] = Lv{l, xs, ys, nx, ny, mag};
    total += (size_t)nx * ny;
  }
  if ((rc = sector_reserve(e, s, total))) return rc;
  size_t used = 0;
  for (int k = 0; k < nl; ++k) {
    long n = (long)lv[k].nx * lv[k].ny;
    s.xy[lv[k].l] = s.buf + used;
    s.n[lv[k].l] = n;
    if (n > 0) {
      rect_fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>(s.xy[lv[k].l], n, lv[k].xs,
                                                                          lv[k].ys, lv[k].nx, lv[k].mag);
      e->launches++;
    }
    used += (size_t)n;
  }
  CU_TRY(e, cudaGetLastError());
  s.cx = (float)(x0 + x1) * 0.5f; // what the manager passes: the sector's integer centre
  s.cy = (float)(y0 + y1) * 0.5f;
  s.rx0 = x0; s.ry0 = y0; s.rx1 = x1; s.ry1 = y1;
  s.integer_grid = true;
  s.banded = (band_y0 > y0 || band_y1 < y1);
  if ((rc = check_levels_nonempty(e, s))) return rc;
  s.kind = SK_RECT;
  { // closed-form tiles: at level l the rectangle is [xs/mag, xe/mag] x [ys/mag, ye/mag]
    size_t total_tiles = 0;
    for (int k = 0; k < nl; ++k)
      total_tiles += (size_t)((lv[k].nx + kTileW - 1) / kTileW) * ((lv[k].ny + kTileH - 1) / kTileH);
    if (total_tiles > s.tcap) {
      if (s.tbuf && s.tbuf_owned) cudaFree(s.tbuf);
      s.tbuf = nullptr; s.tcap = 0;
      void *pv = nullptr;
      if ((rc = sector_alloc(e, sizeof(Tile) * total_tiles, &pv, &s.tbuf_owned))) return rc;
      s.tbuf = static_cast<Tile *>(pv);
      s.tcap = total_tiles;
    }
    for (int l = 0; l < kMaxLevels; ++l) s.tl[l] = TileLevel{};
    size_t tused = 0;
    for (int k = 0; k < nl; ++k) {
      const int ntx = (lv[k].nx + kTileW - 1) / kTileW, nty = (lv[k].ny + kTileH - 1) / kTileH;
      const int nt = ntx * nty;
      if (nt <= 0) continue;
      Tile *dst = s.tbuf + tused;
      rect_tiles_kernel<<<(nt + 127) / 128, 128, 0, e->stream>>>(dst, ntx, nty, lv[k].xs / lv[k].mag,
                                                                 lv[k].ys / lv[k].mag, lv[k].nx, lv[k].ny);
      e->launches++;
      s.tl[lv[k].l].tiles = dst;
      s.tl[lv[k].l].n_tiles = nt;
      tused += (size_t)nt;
    }
    CU_TRY(e, cudaGetLastError());
    s.has_tiles = (x0 >= 0 && y0 >= 0);
  }
  return push_sector(e, id);
}

// Host restatement of the box / corner set-up of manager_class.cpp:838-895 (fp32, same order).
struct AnnulusGeom {
  int x0, y0, x1, y1;
  float c00x = 0, c01x = 0, c10x = 0, c11x = 0, c00y = 0, c01y = 0, c10y = 0, c11y = 0;
  float ri2, ro2;
};
static AnnulusGeom annulus_geom(float r, float dr, float a, float da, float cx, float cy, int as) {
  AnnulusGeom g;
  if (as == 1) {
    g.x0 = (int)(cx - (r + dr)); g.x1 = (int)(cx + (r + dr));
    g.y0 = (int)(cy - (r + dr)); g.y1 = (int)(cy + (r + dr));
  } else {
    float sin0 = (float)std::sin((double)a), cos0 = (float)std::cos((double)a);
    float sin1 = (float)std::sin((double)(a + da)), cos1 = (float)std::cos((double)(a + da));
    float sin2 = (float)std::sin((double)(a + da / 2.f)), cos2 = (float)std::cos((double)(a + da / 2.f));
    volatile float t;
    t = r * cos0; g.c00x = cx + t;
    t = r * cos1; g.c01x = cx + t;
    t = (r + dr) * cos0; t = t * 1.2f; g.c10x = cx + t;
    t = (r + dr) * cos1; t = t * 1.2f; g.c11x = cx + t;
    t = r * sin0; g.c00y = cy + t;
    t = r * sin1; g.c01y = cy + t;
    t = (r + dr) * sin0; t = t * 1.2f; g.c10y = cy + t;
    t = (r + dr) * sin1; t = t * 1.2f; g.c11y = cy + t;
    t = (r + dr) * cos2; float arc_x = cx + t;
    t = (r + dr) * sin2; float arc_y = cy + t;
    g.x0 = (int)std::min(arc_x, std::min(std::min(g.c00x, g.c01x), std::min(g.c10x, g.c11x)));
    g.x1 = (int)std::max(arc_x, std::max(std::max(g.c00x, g.c01x), std::max(g.c10x, g.c11x)));
    g.y0 = (int)std::min(arc_y, std::min(std::min(g.c00y, g.c01y), std::min(g.c10y, g.c11y)));
    g.y1 = (int)std::max(arc_y, std::max(std::max(g.c00y, g.c01y), std::max(g.c10y, g.c11y)));
  }
  volatile float ro = r + dr;
  g.ro2 = ro * ro;
  g.ri2 = r * r;
  return g;
}

// The sequential fp32 centre of the CPU engine for an annulus: its list order is x outer, y inner
// (manager_class.cpp:902-919) and the centre is the running fp32 sum of that list divided by its length
// (pyramid_class.cpp:325-347) -- inherently sequential, so it is replayed on the host. The membership test is not:
// for a full ring (as == 1) the fp32 value r2 = fl(fl(ax * ax) + fl(ay * ay)) is monotone in |ay|, so the members
// of a column are at most two runs of rows whose ends four binary searches find with the reference's own
// predicate; only the 2 x N dependent additions remain (a 9 M pixel ring: 28 ms -> ~10 ms). Annular SECTORS
// (as > 1: the wedge test is not monotone in a column) keep the plain walk over their much smaller box.
static inline float ring_r2(float ax2, float j, float cy) {
  const float ay = j - cy;
  const float ay2 = ay * ay;
  return ax2 + ay2;
}
static void annulus_reference_center(const AnnulusGeom &g, float cx, float cy, int as, float &ocx,
                                     float &ocy, long &count) {
  float sx = 0.f, sy = 0.f;
  long n = 0;
  if (as == 1) {
    // rows j0 .. j1 - 1; r2 falls while j < jc and rises from jc on (jc = first row with j - cy >= 0)
    const int j0 = g.y0, j1 = g.y1;
    int jc = j0;
    while (jc < j1 && (float)jc - cy < 0.f) ++jc;
    for (float i = (float)g.x0; i < (float)g.x1; ++i) {
      const float ax = i - cx;
      const float ax2 = ax * ax;
      // falling side [j0, jc): members are rows with ri2 < r2 < ro2 -> [first r2 < ro2, last r2 > ri2]
      // rising side  [jc, j1): members are                          -> [first r2 > ri2, last r2 < ro2]
      auto first_true = [&](int lo, int hi, auto pred) { // smallest j in [lo, hi) with pred(j), pred monotone false -> true
        while (lo < hi) { int mid = lo + (hi - lo) / 2; if (pred(mid)) hi = mid; else lo = mid + 1; }
        return lo;
      };
      const int a_lo = first_true(j0, jc, [&](int j) { return ring_r2(ax2, (float)j, cy) < g.ro2; });
      const int a_hi = first_true(j0, jc, [&](int j) { return !(ring_r2(ax2, (float)j, cy) > g.ri2); }); // exclusive
      const int b_lo = first_true(jc, j1, [&](int j) { return ring_r2(ax2, (float)j, cy) > g.ri2; });
      const int b_hi = first_true(jc, j1, [&](int j) { return !(ring_r2(ax2, (float)j, cy) < g.ro2); }); // exclusive
      for (int j = a_lo; j < a_hi; ++j) { sx = sx + i; sy = sy + (float)j; }
      for (int j = b_lo; j < b_hi; ++j) { sx = sx + i; sy = sy + (float)j; }
      n += std::max(0, a_hi - a_lo) + std::max(0, b_hi - b_lo);
    }
  } else {
    for (float i = (float)g.x0; i < (float)g.x1; ++i)
      for (int j = g.y0; j < g.y1; ++j) {
        const float ax = i - cx, ay = (float)j - cy;
        const float ax2 = ax * ax, ay2 = ay * ay;
        const float r2 = ax2 + ay2;
        if (r2 > g.ri2 && r2 < g.ro2) {
          const float a1 = g.c11x - i, b1 = g.c01y - g.c11y, a2 = g.c11y - (float)j, b2 = g.c01x - g.c11x;
          const float m1 = a1 * b1, m2 = a2 * b2;
          const float cross1 = m1 - m2;
          const float a3 = g.c00x - i, b3 = g.c10y - g.c00y, a4 = g.c00y - (float)j, b4 = g.c10x - g.c00x;
          const float m3 = a3 * b3, m4 = a4 * b4;
          const float cross2 = m3 - m4;
          const float pr = cross1 * cross2;
          if (pr > 0.f) { sx = sx + i; sy = sy + (float)j; ++n; }
        }
      }
  }
  count = n;
  ocx = n ? sx / (float)n : cx;
  ocy = n ? sy / (float)n : cy;
}

int dic_reset_polygon_rect(dic_engine *e, int id, int x0, int y0, int x1, int y1) {
  return reset_rect_impl(e, id, x0, y0, x1, y1, y0, y1);
}
int dic_reset_polygon_rect_band(dic_engine *e, int id, int x0, int y0, int x1, int y1, int band_y0, int band_y1) {
  if (band_y1 < band_y0) return DIC_ERROR_BAD_ARGUMENT;
  return reset_rect_impl(e, id, x0, y0, x1, y1, std::max(y0, band_y0), std::min(y1, band_y1));
}

// A whole grid of rectangles in one go: what n calls of dic_reset_polygon_rect(first_id + k, boxes[4k..4k+3])
// build (the subdivision loop of manager_class.cpp:274-336 issues exactly those calls on frame 0), with ONE list
// kernel and ONE tile kernel over all sectors and levels, one descriptor upload and one upload of the sector
// records -- 4096 subsets went from 3 launches + a sync each (138 ms) to a few launches in total.
static int ensure_block(dic_engine *e, void **ptr, size_t *cap, size_t bytes) {
  if (bytes <= *cap) return DIC_OK;
  if (*ptr) CU_TRY(e, cudaFree(*ptr));
  *ptr = nullptr; *cap = 0;
  CU_TRY(e, cudaMalloc(ptr, bytes));
  *cap = bytes;
  return DIC_OK;
}

int dic_reset_polygon_rect_grid(dic_engine *e, int first_id, int n, const int *boxes) {
  if (!e || !boxes || n <= 0 || first_id < 0 || first_id + n > (1 << 24)) return DIC_ERROR_BAD_ARGUMENT;
  cudaSetDevice(e->device);
  NvtxRange nvtx("dic_reset_polygon_rect_grid");
  int rc = ensure_sector_capacity(e, first_id + n);
  if (rc) return rc;
  CU_TRY(e, cudaStreamSynchronize(e->stream)); // pinned mirrors / shared blocks of these sectors may be in flight
  int lv_ids[kMaxLevels], nl = 0;
  for (int l = 0; l <= e->stop; ++l)
    if (level_used(e, l)) lv_ids[nl++] = l;
  std::vector<RectDesc> desc((size_t)n * nl);
  size_t total_px = 0, total_tiles = 0;
  int worst = DIC_OK;
  long max_n = 1, max_t = 1;
  for (int k = 0; k < n; ++k) {
    const int x0 = boxes[4 * k], y0 = boxes[4 * k + 1], x1 = boxes[4 * k + 2], y1 = boxes[4 * k + 3];
    Sector &s = e->sectors[first_id + k];
    if (s.buf && s.buf_owned) cudaFree(s.buf);
    if (s.tbuf && s.tbuf_owned) cudaFree(s.tbuf);
    s.buf = nullptr; s.cap = 0; s.buf_owned = false; s.tbuf = nullptr; s.tcap = 0; s.tbuf_owned = false;
    s.kind = SK_NONE; s.pending = false; s.has_tiles = false;
    clear_levels(s);
    for (int l = 0; l < kMaxLevels; ++l) s.tl[l] = TileLevel{};
    const bool bad = x1 < x0 || y1 < y0;
    const size_t slice_begin = total_px;
    for (int j = 0; j < nl; ++j) {
      const int l = lv_ids[j], mag = 1 << l;
      auto first_mult = [mag](int a) { int q = a / mag; if (q * mag < a) ++q; return q * mag; };
      auto last_mult = [mag](int a) { int q = a / mag; if (q * mag > a) --q; return q * mag; };
      RectDesc &d = desc[(size_t)k * nl + j];
      memset(&d, 0, sizeof(d));
      if (bad) continue;
      const int xs = first_mult(x0), xe = last_mult(x1), ys = first_mult(y0), ye = last_mult(y1);
      d.xs = xs; d.ys = ys; d.mag = mag;
      d.nx = xe >= xs ? (xe - xs) / mag + 1 : 0;
      d.ny = ye >= ys ? (ye - ys) / mag + 1 : 0;
      if (d.nx == 0 || d.ny == 0) d.nx = d.ny = 0;
      d.ntx = (d.nx + kTileW - 1) / kTileW; d.nty = (d.ny + kTileH - 1) / kTileH;
      d.list_off = (long long)total_px; d.tile_off = (long long)total_tiles;
      total_px += (size_t)d.nx * d.ny;
      total_tiles += (size_t)d.ntx * d.nty;
      max_n = std::max(max_n, (long)d.nx * d.ny);
      max_t = std::max(max_t, (long)d.ntx * d.nty);
    }
    // a sector's levels are contiguous, with the slack the other builders reserve, so that a Lagrangian update
    // (dic_update_polygon) can re-decimate in place instead of allocating
    const size_t n0 = nl ? (size_t)desc[(size_t)k * nl].nx * desc[(size_t)k * nl].ny : 0;
    total_px = std::max(total_px, slice_begin + n0 + n0 / 2 + 16);
    s.cap = total_px - slice_begin;
  }
  dic_engine::GridBlock *gb = nullptr;
  for (auto &g : e->grid_blocks)
    if (g.first_id == first_id) gb = &g;
  if (!gb) { e->grid_blocks.emplace_back(); gb = &e->grid_blocks.back(); gb->first_id = first_id; }
  if ((rc = ensure_block(e, &gb->lists, &gb->cap_lists, sizeof(float2) * std::max<size_t>(total_px, 1)))) return rc;
  if ((rc = ensure_block(e, &gb->tiles, &gb->cap_tiles, sizeof(Tile) * std::max<size_t>(total_tiles, 1)))) return rc;
  if ((rc = ensure_block(e, &gb->desc, &gb->cap_desc, sizeof(RectDesc) * desc.size()))) return rc;
  float2 *lists = static_cast<float2 *>(gb->lists);
  Tile *tiles = static_cast<Tile *>(gb->tiles);
  CU_TRY(e, cudaMemcpyAsync(gb->desc, desc.data(), sizeof(RectDesc) * desc.size(), cudaMemcpyHostToDevice, e->stream));
  const RectDesc *d_desc = static_cast<const RectDesc *>(gb->desc);
  {
    dim3 g1((unsigned)((max_n + 255) / 256), (unsigned)std::min<size_t>(desc.size(), 65535));
    rect_grid_fill_kernel<<<g1, 256, 0, e->stream>>>(d_desc, (int)desc.size(), lists);
    dim3 g2((unsigned)((max_t + 63) / 64), (unsigned)std::min<size_t>(desc.size(), 65535));
    rect_grid_tiles_kernel<<<g2, 64, 0, e->stream>>>(d_desc, (int)desc.size(), tiles);
    e->launches += 2;
    CU_TRY(e, cudaGetLastError());
  }
  for (int k = 0; k < n; ++k) {
    const int x0 = boxes[4 * k], y0 = boxes[4 * k + 1], x1 = boxes[4 * k + 2], y1 = boxes[4 * k + 3];
    const int id = first_id + k;
    Sector &s = e->sectors[id];
    bool ok = !(x1 < x0 || y1 < y0);
    for (int j = 0; j < nl && ok; ++j) {
      const RectDesc &d = desc[(size_t)k * nl + j];
      const int l = lv_ids[j];
      s.xy[l] = lists + d.list_off;
      s.n[l] = (long)d.nx * d.ny;
      s.tl[l].tiles = tiles + d.tile_off;
      s.tl[l].n_tiles = d.ntx * d.nty;
    }
    if (ok) {
      s.buf = s.xy[0]; // level 0 first: the slice (s.cap elements) a Lagrangian update rewrites in place
      s.cx = (float)(x0 + x1) * 0.5f; s.cy = (float)(y0 + y1) * 0.5f;
      s.rx0 = x0; s.ry0 = y0; s.rx1 = x1; s.ry1 = y1;
      s.integer_grid = true;
      ok = check_levels_nonempty(e, s) == DIC_OK;
    }
    if (!ok) { clear_levels(s); if (worst == DIC_OK) worst = DIC_ERROR_BAD_DOMAIN; }
    else { s.kind = SK_RECT; s.has_tiles = x0 >= 0 && y0 >= 0; }
    // sector records: filled on the host, uploaded in one piece below
    SectorDev &d = e->h_sectors[id];
    memset(&d, 0, sizeof(d));
    for (int l = 0; l < kMaxLevels; ++l) { d.xy[l] = s.xy[l]; d.n[l] = (int)s.n[l]; d.n_total[l] = (int)s.n[l]; }
    d.cx = s.cx; d.cy = s.cy;
    SectorTiles &t = e->h_sector_tiles[id];
    memset(&t, 0, sizeof(t));
    if (s.has_tiles)
      for (int l = 0; l < kMaxLevels; ++l) t.lev[l] = s.tl[l];
  }
  CU_TRY(e, cudaMemcpyAsync(e->d_sectors + first_id, e->h_sectors + first_id, sizeof(SectorDev) * n, cudaMemcpyHostToDevice, e->stream));
  CU_TRY(e, cudaMemcpyAsync(e->d_sector_tiles + first_id, e->h_sector_tiles + first_id, sizeof(SectorTiles) * n,
                            cudaMemcpyHostToDevice, e->stream));
  CU_TRY(e, cudaStreamSynchronize(e->stream)); // `desc` is pageable host memory
  return worst;
}

int dic_last_cluster_size(const dic_engine *e) { return e ? e->last_cluster : 0; }
int dic_set_batch_queue(dic_engine *e, int mode) {
  if (!e || mode < 0 || mode > 2) return DIC_ERROR_BAD_ARGUMENT;
  e->batch_queue = mode;
  return DIC_OK;
}
int dic_set_cluster_mode(dic_engine *e, int mode) {
  if (!e || mode < 0 || mode > 2) return DIC_ERROR_BAD_ARGUMENT;
  e->cluster_mode = mode;
  return DIC_OK;
}

int dic_reset_polygon_annular(dic_engine *e, int id, float r, float dr, float a, float da, float cx,
                              float cy, int as) {
  Sector *sp;
  int rc = begin_sector(e, id, &sp);
  if (rc) return rc;
  Sector &s = *sp;
  if (as < 1 || dr <= 0.f) return DIC_ERROR_BAD_DOMAIN;
  AnnulusGeom g = annulus_geom(r, dr, a, da, cx, cy, as);
  if (g.x1 <= g.x0 || g.y1 <= g.y0) return DIC_ERROR_BAD_DOMAIN;
  AnnulusPred pred[kMaxLevels];
  long ncand[kMaxLevels] = {}, kept[kMaxLevels] = {};
  int lvl[kMaxLevels], nl = 0;
  size_t total = 0;
  for (int l = 0; l <= e->stop; ++l) {
    if (!level_used(e, l)) continue;
    int mag = 1 << l;
    auto first_mult = [mag](int v) { int q = v / mag; if (q * mag < v) ++q; return q * mag; };
    int xs = first_mult(g.x0), ys = first_mult(g.y0);
    int nxc = xs < g.x1 ? (g.x1 - 1 - xs) / mag + 1 : 0;
    int nyc = ys < g.y1 ? (g.y1 - 1 - ys) / mag + 1 : 0;
    AnnulusPred p;
    p.xs = xs; p.ys = ys; p.nxc = std::max(nxc, 1); p.mag = mag; p.as = as;
    p.cx = cx; p.cy = cy; p.ri2 = g.ri2; p.ro2 = g.ro2;
    p.c00x = g.c00x; p.c01x = g.c01x; p.c10x = g.c10x; p.c11x = g.c11x;
    p.c00y = g.c00y; p.c01y = g.c01y; p.c10y = g.c10y; p.c11y = g.c11y;
    pred[nl] = p; ncand[nl] = (long)nxc * nyc; lvl[nl] = l;
    if ((rc = compact_count(e, p, ncand[nl], &kept[nl]))) return rc;
    total += (size_t)kept[nl];
    ++nl;
  }
  if ((rc = sector_reserve(e, s, total))) return rc;
  size_t used = 0;
  for (int k = 0; k < nl; ++k) {
    s.xy[lvl[k]] = s.buf + used;
    s.n[lvl[k]] = kept[k];
    if (kept[k] > 0) {
      // offsets of this level's count pass were overwritten by later levels: recount, then emit
      long again = 0;
      if ((rc = compact_count(e, pred[k], ncand[k], &again))) return rc;
      if ((rc = compact_emit(e, pred[k], ncand[k], s.xy[lvl[k]]))) return rc;
    }
    used += (size_t)kept[k];
  }
  s.integer_grid = true;
  if ((rc = check_levels_nonempty(e, s))) return rc;
  if (e->center_mode == DIC_CENTER_REFERENCE) {
    long cnt = 0;
    annulus_reference_center(g, cx, cy, as, s.cx, s.cy, cnt);
    if (cnt != s.n[0]) { set_error(e, "annulus: host/device membership mismatch"); return DIC_ERROR_BAD_DOMAIN; }
  } else if ((rc = exact_center(e, s))) {
    return rc;
  }
  s.kind = SK_ANNULAR;
  if ((rc = build_tiles(e, s, false))) return rc;
  return push_sector(e, id);
}

static int finish_list_sector(dic_engine *e, int id, Sector &s, SectorKind kind, int use_center, float cx,
                              float cy, const std::vector<float2> *host_list) {
  int rc = decimate_levels(e, s);
  if (rc) return rc;
  if ((rc = check_levels_nonempty(e, s))) return rc;
  if (use_center) {
    s.cx = cx; s.cy = cy;
  } else if (e->center_mode == DIC_CENTER_REFERENCE) {
    std::vector<float2> tmp;
    if (!host_list) {
      tmp.resize(s.n[0]);
      CU_TRY(e, cudaMemcpyAsync(tmp.data(), s.xy[0], sizeof(float2) * s.n[0], cudaMemcpyDeviceToHost, e->stream));
      CU_TRY(e, cudaStreamSynchronize(e->stream));
      host_list = &tmp;
    }
    seq_mean(*host_list, s.cx, s.cy);
  } else if ((rc = exact_center(e, s))) {
    return rc;
  }
  s.kind = kind;
  if (s.integer_grid && (rc = build_tiles(e, s, true))) return rc;
  return push_sector(e, id);
}

int dic_reset_polygon_blob(dic_engine *e, int id, const float *contour_xy, int n_vertices) {
  Sector *sp;
  int rc = begin_sector(e, id, &sp);
  if (rc) return rc;
  Sector &s = *sp;
  if (!contour_xy || n_vertices < 3) return DIC_ERROR_BAD_DOMAIN;
  BlobPolygon poly(contour_xy, n_vertices);
  if (poly.bad()) return DIC_ERROR_BAD_DOMAIN; // manager_class.cpp:1026-1030
  std::vector<HostSpan> hs = poly.spans();
  std::vector<Span> spans(hs.size());
  long total = 0;
  for (size_t i = 0; i < hs.size(); ++i) {
    spans[i] = Span{hs[i].y, hs[i].xb, hs[i].xe, total};
    total += hs[i].xe - hs[i].xb;
  }
  if (total <= 0) return DIC_ERROR_BAD_DOMAIN;
  if ((rc = sector_reserve(e, s, (size_t)total + (size_t)total / 2 + 16))) return rc;
  size_t bytes = sizeof(Span) * spans.size();
  if (bytes > e->cap_stage) {
    if (e->d_stage) cudaFree(e->d_stage);
    e->d_stage = nullptr; e->cap_stage = 0;
    CU_TRY(e, cudaMalloc(&e->d_stage, bytes * 2));
    e->cap_stage = bytes * 2;
  }
  CU_TRY(e, cudaMemcpyAsync(e->d_stage, spans.data(), bytes, cudaMemcpyHostToDevice, e->stream));
  s.xy[0] = s.buf; s.n[0] = total;
  expand_spans_kernel<<<(unsigned)((total + 255) / 256), 256, 0, e->stream>>>(
      reinterpret_cast<const Span *>(e->d_stage), (int)spans.size(), total, s.buf);
  e->launches++;
  CU_TRY(e, cudaGetLastError());
  CU_TRY(e, cudaStreamSynchronize(e->stream));
  s.integer_grid = true;
  if (e->center_mode == DIC_CENTER_REFERENCE) {
    // sequential fp32 mean in emission order, straight from the spans (no list download)
    volatile float sx = 0.f, sy = 0.f;
    for (const HostSpan &h : hs)
      for (int i = h.xb; i < h.xe; ++i) { sx = sx + (float)i; sy = sy + (float)h.y; }
    return finish_list_sector(e, id, s, SK_BLOB, 1, sx / (float)total, sy / (float)total, nullptr);
  }
  return finish_list_sector(e, id, s, SK_BLOB, 0, 0.f, 0.f, nullptr);
}

int dic_reset_polygon_points(dic_engine *e, int id, const float *xy, int64_t n, int use_center, float cx,
                             float cy) {
  Sector *sp;
  int rc = begin_sector(e, id, &sp);
  if (rc) return rc;
  Sector &s = *sp;
  if (!xy || n <= 0) return DIC_ERROR_BAD_DOMAIN;
  if ((rc = sector_reserve(e, s, (size_t)n + (size_t)n / 2 + 16))) return rc;
  CU_TRY(e, cudaMemcpyAsync(s.buf, xy, sizeof(float2) * (size_t)n, cudaMemcpyHostToDevice, e->stream));
  CU_TRY(e, cudaStreamSynchronize(e->stream));
  s.xy[0] = s.buf; s.n[0] = (long)n;
  s.integer_grid = false;
  std::vector<float2> host;
  if (!use_center && e->center_mode == DIC_CENTER_REFERENCE)
    host.assign(reinterpret_cast<const float2 *>(xy), reinterpret_cast<const float2 *>(xy) + n);
  return finish_list_sector(e, id, s, SK_POINTS, use_center, cx, cy, host.empty() ? nullptr : &host);
}

int dic_set_polygon_center(dic_engine *e, int id, float cx, float cy) {
  if (!e || !sector_ok(e, id)) return DIC_ERROR_BAD_ARGUMENT;
  cudaSetDevice(e->device);
  e->sectors[id].cx = cx; e->sectors[id].cy = cy;
  return push_sector(e, id);
}

// CudaClass::updatePolygon with the CPU path's semantics (manager_class.cpp:353-397, 616-676,
// 1050-1110): Lagrangian = every point translated by the centre shift of the last result and
// rounded to the grid (add_pair, :37-47); strict Lagrangian = und points := last deformed points.
int dic_update_polygon(dic_engine *e, int id, int deformation_description) {
  if (!e || !sector_ok(e, id)) return DIC_ERROR_BAD_ARGUMENT;
  if (deformation_description == DIC_DEF_EULERIAN) return DIC_OK; // cuda_polygon.cu:268-275
  if (deformation_description != DIC_DEF_LAGRANGIAN && deformation_description != DIC_DEF_STRICT_LAGRANGIAN)
    return DIC_ERROR_BAD_ARGUMENT;
  cudaSetDevice(e->device);
  CU_TRY(e, cudaStreamSynchronize(e->stream));
  Sector &s = e->sectors[id];
  if (s.banded) { set_error(e, "domain updates of a row-split sector are not supported"); return DIC_ERROR_BAD_ARGUMENT; }
  const dic_result &r = e->h_results[id];
  const int np = np_of(e);
  const float u = r.resultingParameters[0], v = np > 1 ? r.resultingParameters[1] : 0.f;
  const long n0 = s.n[0];
  const unsigned grid = (unsigned)((n0 + 255) / 256);
  if (s.xy[0] != s.buf) return DIC_ERROR_BAD_ARGUMENT;
  if (deformation_description == DIC_DEF_LAGRANGIAN) {
    translate_round_kernel<<<grid, 256, 0, e->stream>>>(s.buf, n0, u, v);
  } else {
    float *d_params = e->d_scratch + 512;
    CU_TRY(e, cudaMemcpyAsync(d_params, r.resultingParameters, sizeof(float) * np, cudaMemcpyHostToDevice, e->stream));
    switch (e->model) {
    case DIC_FM_U: warp_list_kernel<DIC_FM_U><<<grid, 256, 0, e->stream>>>(s.buf, n0, d_params, s.cx, s.cy, s.buf); break;
    case DIC_FM_UV: warp_list_kernel<DIC_FM_UV><<<grid, 256, 0, e->stream>>>(s.buf, n0, d_params, s.cx, s.cy, s.buf); break;
    case DIC_FM_UVQ: warp_list_kernel<DIC_FM_UVQ><<<grid, 256, 0, e->stream>>>(s.buf, n0, d_params, s.cx, s.cy, s.buf); break;
    case DIC_FM_UVUxUyVxVy: warp_list_kernel<DIC_FM_UVUxUyVxVy><<<grid, 256, 0, e->stream>>>(s.buf, n0, d_params, s.cx, s.cy, s.buf); break;
    default: warp_list_kernel<DIC_FM_QUADRATIC><<<grid, 256, 0, e->stream>>>(s.buf, n0, d_params, s.cx, s.cy, s.buf); break;
    }
    s.integer_grid = false;
  }
  e->launches++;
  CU_TRY(e, cudaGetLastError());
  for (int l = 1; l < kMaxLevels; ++l) { s.xy[l] = nullptr; s.n[l] = 0; }
  s.has_tiles = false;
  int rc = decimate_levels(e, s);
  if (rc) return rc;
  if ((rc = check_levels_nonempty(e, s))) return rc;
  if (s.kind == SK_RECT) {
    // the manager passes the rounded deformed centre of the previous frame (:2085-2086)
    s.cx = (float)(int)(s.cx + u + 0.5f);
    s.cy = (float)(int)(s.cy + v + 0.5f);
  } else if (e->center_mode == DIC_CENTER_REFERENCE) {
    std::vector<float2> h;
    if ((rc = download_list(e, s.xy[0], n0, h))) return rc;
    if (deformation_description == DIC_DEF_LAGRANGIAN) to_reference_order(s, h);
    seq_mean(h, s.cx, s.cy);
  } else if ((rc = exact_center(e, s))) {
    return rc;
  }
  if (s.integer_grid && (rc = build_tiles(e, s, s.kind == SK_BLOB || s.kind == SK_POINTS))) return rc;
  return push_sector(e, id);
}

// ------------------------------------------------------------------ row-split (one domain, several GPUs)

int dic_rowsplit_mailbox_handle(dic_engine *e, void *handle_out, int handle_bytes) {
  if (!e || !handle_out || handle_bytes < (int)sizeof(cudaIpcMemHandle_t)) return DIC_ERROR_BAD_ARGUMENT;
  cudaSetDevice(e->device);
  if (!e->d_mailbox) {
    CU_TRY(e, cudaMalloc(&e->d_mailbox, sizeof(Mailbox)));
    CU_TRY(e, cudaMemset(e->d_mailbox, 0, sizeof(Mailbox)));
  }
  cudaIpcMemHandle_t h;
  CU_TRY(e, cudaIpcGetMemHandle(&h, e->d_mailbox));
  memcpy(handle_out, &h, sizeof(h));
  return DIC_OK;
}

int dic_rowsplit_connect(dic_engine *e, int rank, int world, const void *handles, int handle_bytes) {
  if (!e || rank < 0 || world < 1 || world > kMaxRanks || rank >= world) return DIC_ERROR_BAD_ARGUMENT;
  cudaSetDevice(e->device);
  CU_TRY(e, cudaStreamSynchronize(e->stream));
  if (!e->d_mailbox) {
    CU_TRY(e, cudaMalloc(&e->d_mailbox, sizeof(Mailbox)));
  }
  CU_TRY(e, cudaMemset(e->d_mailbox, 0, sizeof(Mailbox)));
  for (int r = 0; r < world; ++r) {
    if (r == rank) { e->peer_mailbox[r] = e->d_mailbox; continue; }
    if (!handles || handle_bytes < (int)sizeof(cudaIpcMemHandle_t)) return DIC_ERROR_BAD_ARGUMENT;
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const char *>(handles) + (size_t)r * handle_bytes, sizeof(h));
    void *p = nullptr;
    CU_TRY(e, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    e->peer_mailbox[r] = p;
  }
  e->rs_rank = rank; e->rs_world = world;
  GridWork w;
  CU_TRY(e, cudaMemcpy(&w, e->d_work, sizeof(GridWork), cudaMemcpyDeviceToHost));
  w.rs_rank = rank; w.rs_world = world; w.rs_seq = 0;
  w.rs_local = e->d_mailbox;
  for (int r = 0; r < kMaxRanks; ++r) w.rs_peer[r] = static_cast<Mailbox *>(r < world ? e->peer_mailbox[r] : nullptr);
  CU_TRY(e, cudaMemcpy(e->d_work, &w, sizeof(GridWork), cudaMemcpyHostToDevice));
  return DIC_OK;
}

int dic_rowsplit_disconnect(dic_engine *e) {
  if (!e) return DIC_ERROR_BAD_ARGUMENT;
  cudaSetDevice(e->device);
  CU_TRY(e, cudaStreamSynchronize(e->stream));
  for (int r = 0; r < e->rs_world; ++r)
    if (r != e->rs_rank && e->peer_mailbox[r]) cudaIpcCloseMemHandle(e->peer_mailbox[r]);
  for (auto &p : e->peer_mailbox) p = nullptr;
  e->rs_rank = 0; e->rs_world = 1;
  GridWork w;
  CU_TRY(e, cudaMemcpy(&w, e->d_work, sizeof(GridWork), cudaMemcpyDeviceToHost));
  w.rs_rank = 0; w.rs_world = 1; w.rs_local = nullptr; w.rs_seq = 0;
  for (auto &p : w.rs_peer) p = nullptr;
  CU_TRY(e, cudaMemcpy(e->d_work, &w, sizeof(GridWork), cudaMemcpyHostToDevice));
  return DIC_OK;
}

// ------------------------------------------------------------------ correlate

static int enqueue_correlate(dic_engine *e, int first, int count, const float *guesses, bool grid_mode) {
  int rc = images_ready(e);
  if (rc) return rc;
  const int np = np_of(e);
  NvtxRange nvtx(grid_mode ? "dic_correlate" : "dic_correlate_batch");
  bool in_flight = false;
  for (int i = 0; i < count; ++i) {
    if (!sector_ok(e, first + i)) { set_error(e, "unknown sector"); return DIC_ERROR_BAD_ARGUMENT; }
    in_flight = in_flight || e->sectors[first + i].pending;
  }
  // The guesses do not travel by memcpy: a single sector's guess rides in the kernel parameters, a batch reads the
  // pinned block below directly (zero-copy). Nothing of a correlate() queues on the H2D copy engine, so it cannot
  // end up behind the next image pair's upload (which used to serialise upload and solve on the batch path).
  if (in_flight) CU_TRY(e, cudaStreamSynchronize(e->stream)); // an un-waited async solve may still read the block
  for (int i = 0; i < count; ++i) {
    float *g = e->h_guess + (size_t)(first + i) * kMaxParams;
    for (int k = 0; k < kMaxParams; ++k) g[k] = k < np ? guesses[(size_t)i * np + k] : 0.f;
    e->sectors[first + i].pending = true;
  }
  CU_TRY(e, cudaEventRecord(e->ev_step0, e->stream));
  CU_TRY(e, cudaEventRecord(e->ev0, e->stream));
  const int tj = e->trace_on && e->trace_solve < dic_engine::kTrace ? e->trace_solve++ : -1;
  if (tj >= 0) cudaEventRecord(e->trace_so[tj][0], e->stream);
  rc = tiles_applicable(e, first, count) ? launch_solve_tiles_any(e, grid_mode, first, count)
                                         : launch_solve_any(e, grid_mode, first, count);
  if (rc) return rc;
  CU_TRY(e, cudaEventRecord(e->ev1, e->stream));
  if (tj >= 0) cudaEventRecord(e->trace_so[tj][1], e->stream);
  e->timing_pending = true;
  // no download: the result records are written by the kernel into the pinned, mapped block the host reads after
  // the stream has drained (a D2H copy of 176 B cost ~10 us of copy-engine latency per correlate)
  CU_TRY(e, cudaEventRecord(e->ev_step1, e->stream));
  return DIC_OK;
}

static int collect(dic_engine *e, int first, int count, float *guesses_out, dic_result *results) {
  CU_TRY(e, cudaStreamSynchronize(e->stream));
  if (e->timing_pending) {
    cudaEventElapsedTime(&e->last_ms, e->ev0, e->ev1);
    cudaEventElapsedTime(&e->last_step_ms, e->ev_step0, e->ev_step1);
    e->timing_pending = false;
  }
  const int np = np_of(e);
  int worst = DIC_OK;
  bool aborted = false;
  for (int i = 0; i < count; ++i) aborted = aborted || e->h_results[first + i].errorCode == DIC_ERROR_CUDA;
  if (aborted) {
    set_error(e, "a bounded wait inside a kernel expired (grid barrier 10 s / TMA transfer 2 s)");
    cudaMemsetAsync(&e->d_work->img_error, 0, sizeof(int), e->stream); // sticky flag of the pyramid kernel: reported once
  }
  for (int i = 0; i < count; ++i) {
    const dic_result &r = e->h_results[first + i];
    e->sectors[first + i].pending = false;
    if (results) results[i] = r;
    if (guesses_out)
      for (int k = 0; k < np; ++k) guesses_out[(size_t)i * np + k] = r.resultingParameters[k];
    if (r.errorCode != DIC_OK && worst == DIC_OK) worst = r.errorCode;
  }
  return worst;
}

int dic_correlate_async(dic_engine *e, int id, const float *guess) {
  if (!e || !guess) return DIC_ERROR_BAD_ARGUMENT;
  cudaSetDevice(e->device);
  return enqueue_correlate(e, id, 1, guess, true);
}
int dic_correlate_wait(dic_engine *e, int id, float *guess_out, dic_result *out) {
  if (!e || !sector_ok(e, id)) return DIC_ERROR_BAD_ARGUMENT;
  cudaSetDevice(e->device);
  return collect(e, id, 1, guess_out, out);
}
int dic_correlate(dic_engine *e, int id, float *guess_inout, dic_result *out) {
  int rc = dic_correlate_async(e, id, guess_inout);
  if (rc) return rc;
  return collect(e, id, 1, guess_inout, out);
}
int dic_correlate_batch_async(dic_engine *e, int first, int n, const float *guesses) {
  if (!e || !guesses || n <= 0) return DIC_ERROR_BAD_ARGUMENT;
  cudaSetDevice(e->device);
  return enqueue_correlate(e, first, n, guesses, false);
}
int dic_correlate_batch_wait(dic_engine *e, int first, int n, float *guesses_out, dic_result *results) {
  if (!e || n <= 0) return DIC_ERROR_BAD_ARGUMENT;
  for (int i = 0; i < n; ++i)
    if (!sector_ok(e, first + i)) return DIC_ERROR_BAD_ARGUMENT;
  cudaSetDevice(e->device);
  return collect(e, first, n, guesses_out, results);
}
int dic_correlate_batch(dic_engine *e, int first, int n, float *guesses_inout, dic_result *results) {
  if (!e || !guesses_inout || n <= 0) return DIC_ERROR_BAD_ARGUMENT;
  cudaSetDevice(e->device);
  int rc = enqueue_correlate(e, first, n, guesses_inout, false);
  if (rc) return rc;
  return collect(e, first, n, guesses_inout, results);
}

// ------------------------------------------------------------------ read-back / introspection

static int download_list(dic_engine *e, const float2 *d, long n, std::vector<float2> &h) {
  h.resize((size_t)n);
  if (n == 0) return DIC_OK;
  CU_TRY(e, cudaMemcpyAsync(h.data(), d, sizeof(float2) * (size_t)n, cudaMemcpyDeviceToHost, e->stream));
  CU_TRY(e, cudaStreamSynchronize(e->stream));
  return DIC_OK;
}

static void to_reference_order(const Sector &s, std::vector<float2> &h) {
  // rect / annulus lists are stored row-major for coalescing; the CPU engine builds them x-outer,
  // y-inner (manager_class.cpp:1607-1611, :902-919)
  if (s.kind == SK_RECT || s.kind == SK_ANNULAR)
    std::stable_sort(h.begin(), h.end(), [](const float2 &a, const float2 &b) {
      return a.x < b.x || (a.x == b.x && a.y < b.y);
    });
}

int dic_get_level_points(dic_engine *e, int id, int level, float *xy, int64_t cap, int64_t *n_needed) {
  if (!e || !sector_ok(e, id) || level < 0 || level >= kMaxLevels) return DIC_ERROR_BAD_ARGUMENT;
  cudaSetDevice(e->device);
  const Sector &s = e->sectors[id];
  if (n_needed) *n_needed = s.n[level];
  if (!xy || cap <= 0) return DIC_OK;
  std::vector<float2> h;
  int rc = download_list(e, s.xy[level], s.n[level], h);
  if (rc) return rc;
  to_reference_order(s, h);
  memcpy(xy, h.data(), sizeof(float2) * (size_t)std::min<int64_t>(cap, s.n[level]));
  return DIC_OK;
}
int dic_get_und_xy0(dic_engine *e, int id, float *xy, int64_t cap, int64_t *n_needed) {
  return dic_get_level_points(e, id, 0, xy, cap, n_needed);
}

int dic_get_def_xy0(dic_engine *e, int id, float *xy, int64_t cap, int64_t *n_needed) {
  if (!e || !sector_ok(e, id)) return DIC_ERROR_BAD_ARGUMENT;
  cudaSetDevice(e->device);
  const Sector &s = e->sectors[id];
  if (n_needed) *n_needed = s.n[0];
  if (!xy || cap <= 0) return DIC_OK;
  // correlation_class.cpp:884-896: the model applied to the level-0 list with the last result
  const int np = np_of(e);
  float *d_params = e->d_scratch + 512;
  CU_TRY(e, cudaMemcpyAsync(d_params, e->h_results[id].resultingParameters, sizeof(float) * np,
                            cudaMemcpyHostToDevice, e->stream));
  float2 *tmp = nullptr;
  CU_TRY(e, cudaMalloc(&tmp, sizeof(float2) * (size_t)s.n[0]));
  unsigned grid = (unsigned)((s.n[0] + 255) / 256);
  switch (e->model) {
  case DIC_FM_U: warp_list_kernel<DIC_FM_U><<<grid, 256, 0, e->stream>>>(s.xy[0], s.n[0], d_params, s.cx, s.cy, tmp); break;
  case DIC_FM_UV: warp_list_kernel<DIC_FM_UV><<<grid, 256, 0, e->stream>>>(s.xy[0], s.n[0], d_params, s.cx, s.cy, tmp); break;
  case DIC_FM_UVQ: warp_list_kernel<DIC_FM_UVQ><<<grid, 256, 0, e->stream>>>(s.xy[0], s.n[0], d_params, s.cx, s.cy, tmp); break;
  case DIC_FM_UVUxUyVxVy: warp_list_kernel<DIC_FM_UVUxUyVxVy><<<grid, 256, 0, e->stream>>>(s.xy[0], s.n[0], d_params, s.cx, s.cy, tmp); break;
  default: warp_list_kernel<DIC_FM_QUADRATIC><<<grid, 256, 0, e->stream>>>(s.xy[0], s.n[0], d_params, s.cx, s.cy, tmp); break;
  }
  e->launches++;
  // keep the pairing with the reference-ordered und list: sort by the und key
  std::vector<float2> hu, hd;
  int rc = download_list(e, s.xy[0], s.n[0], hu);
  if (!rc) rc = download_list(e, tmp, s.n[0], hd);
  cudaFree(tmp);
  if (rc) return rc;
  if (s.kind == SK_RECT || s.kind == SK_ANNULAR) {
    std::vector<size_t> idx(hu.size());
    for (size_t i = 0; i < idx.size(); ++i) idx[i] = i;
    std::stable_sort(idx.begin(), idx.end(), [&](size_t a, size_t b) {
      return hu[a].x < hu[b].x || (hu[a].x == hu[b].x && hu[a].y < hu[b].y);
    });
    std::vector<float2> t(hd.size());
    for (size_t i = 0; i < idx.size(); ++i) t[i] = hd[idx[i]];
    hd.swap(t);
  }
  memcpy(xy, hd.data(), sizeof(float2) * (size_t)std::min<int64_t>(cap, s.n[0]));
  return DIC_OK;
}

int dic_get_level_center(dic_engine *e, int id, int level, float *cx, float *cy) {
  if (!e || !sector_ok(e, id) || level < 0 || level >= kMaxLevels) return DIC_ERROR_BAD_ARGUMENT;
  float inv = 1.f / (float)(1 << level);
  if (cx) *cx = e->sectors[id].cx * inv;
  if (cy) *cy = e->sectors[id].cy * inv;
  return DIC_OK;
}

int dic_get_pyramid_level(dic_engine *e, int which, int level, uint8_t *out, int *rows, int *cols) {
  if (!e || which < 0 || which > 2 || level < 0 || level > e->stop) return DIC_ERROR_BAD_ARGUMENT;
  cudaSetDevice(e->device);
  const PyramidSlot &s = e->pyr[e->role[which]];
  if (!s.valid) return DIC_ERROR_BAD_ARGUMENT;
  const LevelImage &li = s.lev[level];
  if (rows) *rows = li.rows;
  if (cols) *cols = li.cols;
  if (!out) return DIC_OK;
  cudaStream_t st = which == 2 ? e->img_stream : e->stream;
  CU_TRY(e, cudaMemcpy2DAsync(out, (size_t)li.cols * li.colors, li.ptr, li.pitch, (size_t)li.cols * li.colors, li.rows,
                              cudaMemcpyDeviceToHost, st));
  CU_TRY(e, cudaStreamSynchronize(st));
  return DIC_OK;
}

int dic_evaluate(dic_engine *e, int id, int level, const float *params, float *A, float *b, float *chi,
                 int *n_oob) {
  if (!e || !sector_ok(e, id) || !params || level < 0 || level > e->stop) return DIC_ERROR_BAD_ARGUMENT;
  cudaSetDevice(e->device);
  int rc = images_ready(e);
  if (rc) return rc;
  const Sector &s = e->sectors[id];
  if (s.n[level] <= 0) return DIC_ERROR_BAD_DOMAIN;
  const int np = np_of(e);
  const int nacc = np * (np + 1) / 2 + np + 2;
  float *d_params = e->d_scratch + 512;
  float *d_out = e->d_scratch;
  CU_TRY(e, cudaMemcpyAsync(d_params, params, sizeof(float) * np, cudaMemcpyHostToDevice, e->stream));
  int grid = (int)std::max<long>(1, std::min<long>((s.n[level] + kThreads * 4 - 1) / (kThreads * 4), e->max_grid));
  switch (e->model) {
  case DIC_FM_U: rc = launch_eval_model<DIC_FM_U>(e, id, level, d_params, grid); break;
  case DIC_FM_UV: rc = launch_eval_model<DIC_FM_UV>(e, id, level, d_params, grid); break;
  case DIC_FM_UVQ: rc = launch_eval_model<DIC_FM_UVQ>(e, id, level, d_params, grid); break;
  case DIC_FM_UVUxUyVxVy: rc = launch_eval_model<DIC_FM_UVUxUyVxVy>(e, id, level, d_params, grid); break;
  default: rc = launch_eval_model<DIC_FM_QUADRATIC>(e, id, level, d_params, grid); break;
  }
  if (rc) return rc;
  sum_partials_kernel<<<1, 128, 0, e->stream>>>(e->d_partials, grid, nacc, d_out);
  e->launches++;
  float h[kAccStride];
  CU_TRY(e, cudaMemcpyAsync(h, d_out, sizeof(float) * nacc, cudaMemcpyDeviceToHost, e->stream));
  CU_TRY(e, cudaStreamSynchronize(e->stream));
  int k = 0;
  for (int p1 = 0; p1 < np; ++p1)
    for (int p2 = 0; p2 < np; ++p2) A[p1 * np + p2] = 0.f;
  for (int p1 = 0; p1 < np; ++p1)
    for (int p2 = p1; p2 < np; ++p2) A[p1 * np + p2] = h[k++];
  for (int p = 0; p < np; ++p) b[p] = h[k++];
  if (chi) *chi = h[k];
  if (n_oob) *n_oob = (int)(h[k + 1] + 0.5f);
  return DIC_OK;
}

int dic_solve_step(dic_engine *e, const float *A_upper, const float *b, float lambda, float scaling,
                   float *dp) {
  if (!e || !A_upper || !b || !dp) return DIC_ERROR_BAD_ARGUMENT;
  cudaSetDevice(e->device);
  const int np = np_of(e);
  float h[kAccStride];
  int k = 0;
  for (int p1 = 0; p1 < np; ++p1)
    for (int p2 = p1; p2 < np; ++p2) h[k++] = A_upper[p1 * np + p2];
  for (int p = 0; p < np; ++p) h[k++] = b[p];
  float *d_tot = e->d_scratch, *d_dp = e->d_scratch + 256;
  int *d_ok = reinterpret_cast<int *>(e->d_scratch + 300);
  CU_TRY(e, cudaMemcpyAsync(d_tot, h, sizeof(float) * k, cudaMemcpyHostToDevice, e->stream));
  switch (np) {
  case 1: solve_step_kernel<1><<<1, 32, 0, e->stream>>>(d_tot, scaling, lambda, d_dp, d_ok); break;
  case 2: solve_step_kernel<2><<<1, 32, 0, e->stream>>>(d_tot, scaling, lambda, d_dp, d_ok); break;
  case 3: solve_step_kernel<3><<<1, 32, 0, e->stream>>>(d_tot, scaling, lambda, d_dp, d_ok); break;
  case 6: solve_step_kernel<6><<<1, 32, 0, e->stream>>>(d_tot, scaling, lambda, d_dp, d_ok); break;
  default: solve_step_kernel<12><<<1, 32, 0, e->stream>>>(d_tot, scaling, lambda, d_dp, d_ok); break;
  }
  e->launches++;
  int ok = 0;
  CU_TRY(e, cudaMemcpyAsync(dp, d_dp, sizeof(float) * np, cudaMemcpyDeviceToHost, e->stream));
  CU_TRY(e, cudaMemcpyAsync(&ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  CU_TRY(e, cudaStreamSynchronize(e->stream));
  return ok ? DIC_OK : DIC_ERROR_SOLVER;
}

int dic_pipe_trace(dic_engine *e, int on, float *stage_ms, float *solve_ms, int cap, int *n_solves) {
  if (!e) return 0;
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  int n_st = 0;
  if (e->trace_on && stage_ms && solve_ms) {
    n_st = std::min(cap, e->trace_stage);
    const int n_so = std::min(cap, e->trace_solve);
    for (int i = 0; i < n_st; ++i)
      for (int k = 0; k < 4; ++k) cudaEventElapsedTime(&stage_ms[4 * i + k], e->trace_base, e->trace_st[i][k]);
    for (int i = 0; i < n_so; ++i)
      for (int k = 0; k < 2; ++k) cudaEventElapsedTime(&solve_ms[2 * i + k], e->trace_base, e->trace_so[i][k]);
    if (n_solves) *n_solves = n_so;
  }
  if (on && !e->trace_base) {
    cudaEventCreate(&e->trace_base);
    for (auto &r : e->trace_st) for (auto &ev : r) cudaEventCreate(&ev);
    for (auto &r : e->trace_so) for (auto &ev : r) cudaEventCreate(&ev);
  }
  e->trace_on = on != 0;
  e->trace_stage = e->trace_solve = 0;
  if (on) { cudaEventRecord(e->trace_base, e->stream); cudaStreamSynchronize(e->stream); }
  cudaGetLastError(); // a mark that was never recorded (a call that failed half-way) must not surface in a later launch check
  return n_st;
}
int dic_get_timeline(dic_engine *e, unsigned long long *marks, int cap) {
  if (!e || !marks) return 0;
  cudaSetDevice(e->device);
  GridWork h;
  if (cudaMemcpy(&h, e->d_work, sizeof(GridWork), cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
  int n = std::min(std::min(h.n_marks, cap), kMaxMarks);
  memcpy(marks, h.marks, sizeof(unsigned long long) * 4 * n);
  if (n < cap) marks[4 * n] = h.slow_units; // one extra word: units that took the per-pixel path
  return n;
}
int dic_get_cta_times(dic_engine *e, unsigned long long *out, int cap) {
  if (!e || !out) return 0;
  cudaSetDevice(e->device);
  int n = std::min(cap, kMaxCtaMarks);
  if (cudaMemcpy(out, (const char *)e->d_work + offsetof(GridWork, cta_done), sizeof(unsigned long long) * n,
                 cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
  return n;
}
int dic_get_cta_smids(dic_engine *e, unsigned int *out, int cap) {
  if (!e || !out) return 0;
  cudaSetDevice(e->device);
  int n = std::min(cap, kMaxCtaMarks);
  if (cudaMemcpy(out, (const char *)e->d_work + offsetof(GridWork, cta_smid), sizeof(unsigned int) * n,
                 cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
  return n;
}
float dic_last_correlate_ms(dic_engine *e) { return e ? e->last_ms : 0.f; }
float dic_last_step_ms(dic_engine *e) { return e ? e->last_step_ms : 0.f; }
int64_t dic_kernel_launches(const dic_engine *e) { return e ? (int64_t)e->launches.load() : 0; }
void *dic_correlation_stream(dic_engine *e) { return e ? (void *)e->stream : nullptr; }
int dic_synchronize(dic_engine *e) {
  if (!e) return DIC_ERROR_BAD_ARGUMENT;
  cudaSetDevice(e->device);
  CU_TRY(e, cudaStreamSynchronize(e->stream));
  CU_TRY(e, cudaStreamSynchronize(e->img_stream));
  return DIC_OK;
}

} // extern "C"
