/* include/dic_b200.h -- C-ABI of libdic_b200.so, the B200-native DIC engine.
 *
 * Drop-in boundary: one `extern "C"` entry point per public method of
 * the reference's GPU facade `CudaClass` (cuda_class.cuh:46-79), which is the only thing the
 * reference's orchestrator (`managerClass`, manager_class.h:84-86) talks to on its GPU path.
 * Same verbs, same argument meaning, same error enum -- but raw buffers instead of file
 * paths / cv::Mat / v_points, and RESULT SEMANTICS OF THE REFERENCE CPU ENGINE
 * (chi = sum V^2 / N, look-ahead step damped with max(0.4 lambda, 1e-9), `iterations` = index
 * inside the last level, error 2 on out-of-image).
 *
 * Plain pointers and sizes only; no CUDA, torch or C++ types in any signature. All functions
 * return a dic_error (0 = ok) unless stated. The engine owns all device memory. Calls on one
 * engine are not re-entrant, except dic_reset_next_pyramid which may run on a second host
 * thread concurrently with dic_correlate (manager_class.cpp:1438-1447).
 */
#ifndef DIC_B200_H
#define DIC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DIC_MAX_PARAMS 12
#define DIC_MAX_LEVELS 8

/* enums.hpp:25-35 errorEnum, same numeric values */
typedef enum {
  DIC_OK = 0,
  DIC_ERROR_MODEL_OUT_OF_IMAGE = 1,
  DIC_ERROR_INTERPOLATION_OUT_OF_IMAGE = 2,
  DIC_ERROR_MAX_ITERS_REACHED = 3,
  DIC_ERROR_BAD_DOMAIN = 4,
  DIC_ERROR_SOLVER = 5, /* enums.hpp: error_cuSolver -- here: non-SPD normal equations */
  DIC_ERROR_CUDA = 6,
  DIC_ERROR_MULTITHREAD = 7,
  DIC_ERROR_BAD_ARGUMENT = 8 /* extension: API misuse (null pointer, unknown sector, ...) */
} dic_error;

/* enums.hpp:17-23 fittingModelEnum (+ the 12-parameter extension of BASELINE.json) */
typedef enum {
  DIC_FM_U = 0,
  DIC_FM_UV = 1,
  DIC_FM_UVQ = 2,
  DIC_FM_UVUxUyVxVy = 3,
  DIC_FM_QUADRATIC = 4 /* p = u v ux uy vx vy uxx uxy uyy vxx vxy vyy (extension) */
} dic_fitting_model;

/* enums.hpp:10-15 interpolationModelEnum */
typedef enum { DIC_IM_NEAREST = 0, DIC_IM_BILINEAR = 1, DIC_IM_BICUBIC = 2 } dic_interpolation_model;

/* enums.hpp:83-88 deformationDescriptionEnum */
typedef enum { DIC_DEF_STRICT_LAGRANGIAN = 0, DIC_DEF_LAGRANGIAN = 1, DIC_DEF_EULERIAN = 2 } dic_deformation_description;

/* Arithmetic form of the per-pixel evaluation (extension):
 *  PARITY: the reference's fp32 operation order (monomial bicubic in 1+t, unfused mul/add),
 *          bit-identical w, dw/dx, dw/dy per pixel to interpolation_class.cpp:79-138.
 *  FAST:   the same interpolant (Catmull-Rom) in weight form with FMAs; more accurate than
 *          the reference's own rounding noise, ~3x fewer instructions. */
typedef enum { DIC_MODE_PARITY = 0, DIC_MODE_FAST = 1 } dic_arith_mode;

/* How the domain centre is derived when the caller does not give one (annulus, blob):
 *  REFERENCE: sequential fp32 mean over the pixel list in the reference's list order
 *             (pyramid_class.cpp:325-347) -- reproduces the CPU engine's centre bit for bit.
 *  EXACT:     the true mean (integer sums on the device). */
typedef enum { DIC_CENTER_REFERENCE = 0, DIC_CENTER_EXACT = 1 } dic_center_mode;

/* domains.hpp:110-118 CorrelationResult, widened to 12 parameters plus work accounting. */
typedef struct {
  float resultingParameters[DIC_MAX_PARAMS];
  float chi;          /* last_good_chi of the finest level, = sum V^2 / N (CPU semantics) */
  int numberOfPoints; /* level-0 pixel count */
  int iterations;     /* iteration index reached inside the LAST level (correlation_class.cpp:870) */
  int errorCode;      /* dic_error */
  float undCenterX;
  float undCenterY;
  /* extension: per-level accounting (index = pyramid level) */
  int iterationsPerLevel[DIC_MAX_LEVELS];
  int evaluationsPerLevel[DIC_MAX_LEVELS]; /* passes over the level's pixels */
  int pointsPerLevel[DIC_MAX_LEVELS];
} dic_result;

typedef struct dic_engine dic_engine;

/* ---- CudaClass::initialize (cuda_class.cu:40-93): number of usable devices (0 => caller
 *      disables GPU mode, mainapp.cpp:99). */
int dic_device_count(void);

/* ---- CudaClass ctor/dtor. `device` is the CUDA ordinal (reference hard-codes 0,
 *      cuda_class.cu:333-338). Returns NULL on failure. */
dic_engine *dic_create(int device);
void dic_destroy(dic_engine *e);
/* last CUDA / engine error text for this engine (never NULL) */
const char *dic_last_error(const dic_engine *e);

/* ---- setters: CudaClass::set_max_iters / set_precision / set_fitting_model /
 *      set_interpolation_model (cuda_class.cu:95-101, 475-496) */
int dic_set_max_iters(dic_engine *e, int maximum_iterations);
int dic_set_precision(dic_engine *e, float required_precision);
int dic_set_fitting_model(dic_engine *e, int fitting_model);
int dic_set_interpolation_model(dic_engine *e, int interpolation_model);
/* extensions */
int dic_set_arith_mode(dic_engine *e, int arith_mode);
int dic_set_center_mode(dic_engine *e, int center_mode);
/* 0 = automatic (tile kernel for integer-grid domains with the affine / quadratic model and
 * bicubic interpolation, pixel-list kernel otherwise), 1 = force the pixel-list kernel */
int dic_set_kernel_variant(dic_engine *e, int variant);

/* ---- CudaClass::resetImagePyramids(undPath, defPath, nxtPath, color, start, step, stop)
 *      (cuda_class.cu:512-572): host u8 images (row-major, `channels` interleaved: 1 = monochrome,
 *      3 = the reference's color_color mode), builds all three pyramids. nxt may be NULL.
 *      Colour follows the CPU engine as executed: per-channel pyramid (pyramid_class.cpp:52-134), per-colour loop
 *      of the evaluation (interpolation_class.cpp:712-750) INCLUDING the column indexing of its bicubic / bilinear
 *      coefficient builders (ix * (3 + channel), :268-273, :356-359). Colour images take the generic pixel-list
 *      kernel, need a width that stays even down to the coarsest level, and later next / def images passed to
 *      dic_reset_next_pyramid / dic_reset_def_pyramid have the same channel count. */
int dic_reset_image_pyramids(dic_engine *e, const uint8_t *und, const uint8_t *def,
                             const uint8_t *nxt, int rows, int cols, int channels,
                             int pyramid_start, int pyramid_step, int pyramid_stop);
/* same, but the level-0 images already live in device memory (pitch in bytes). The call drains the device first
 * (the images may have been produced on any stream); the per-frame _device variants below do NOT: the caller
 * orders the producer of nxt_dev / def_dev before the call (stream or event synchronisation). */
int dic_reset_image_pyramids_device(dic_engine *e, const void *und_dev, const void *def_dev,
                                    const void *nxt_dev, int rows, int cols, int pitch,
                                    int pyramid_start, int pyramid_step, int pyramid_stop);
/* ---- CudaClass::resetNextPyramid(nxtPath) (cuda_class.cu:498-510) */
int dic_reset_next_pyramid(dic_engine *e, const uint8_t *nxt, int rows, int cols);
/* extension: enqueue only -- returns once the copy and the pyramid build are queued on the image stream; `nxt` must
 * stay alive and unchanged until dic_make_def_pyramid_from_nxt (which orders the solve behind them) */
int dic_reset_next_pyramid_async(dic_engine *e, const uint8_t *nxt, int rows, int cols);
int dic_reset_next_pyramid_device(dic_engine *e, const void *nxt_dev, int rows, int cols, int pitch);
/* replace only the deformed image (host convenience for frame loops without a nxt slot) */
int dic_reset_def_pyramid(dic_engine *e, const uint8_t *def, int rows, int cols);
int dic_reset_def_pyramid_device(dic_engine *e, const void *def_dev, int rows, int cols, int pitch);
/* extension: double-buffered ingest of whole image pairs. dic_stage_next_pair enqueues the upload
 * and the pyramid build of the NEXT (und, def) pair on the image stream and returns at once: both
 * host buffers must stay untouched until the dic_advance_pair that makes the staged pair current
 * (a stream-side wait, no host sync). A loop `advance; stage(k + 1); correlate(k)` overlaps the
 * PCIe transfer of pair k + 1 with the solve of pair k -- the same overlap the reference gets from
 * resetNextPyramid on its loader thread (manager_class.cpp:1438-1447), for both images.
 * Up to TWO pairs may be staged ahead (a third dic_stage_next_pair without a dic_advance_pair is refused):
 * `advance; correlate_async(k); stage(k + 2); correlate_wait(k)` keeps the bus busy while the host waits for
 * the solve -- with one pair ahead the transfer of pair k + 2 cannot be enqueued before the host has seen the
 * end of solve k. dic_advance_pair makes the OLDEST staged pair current and waits (stream-side) for that
 * pair's pyramids only. */
int dic_stage_next_pair(dic_engine *e, const uint8_t *und, const uint8_t *def, int rows, int cols);
int dic_advance_pair(dic_engine *e);
/* same for a GPU that works on a band of the image only (sharded subsets, row-split domain): `und` and
 * `def` still point at the FULL host images, but only rows [row_begin, row_end) are transferred and only
 * the pyramid rows they fully determine are rebuilt (two level-rows fewer per level at each cut). Rows
 * outside keep whatever the slot held before: the caller's band must cover its domains plus their
 * displacement, the 2-pixel bicubic halo and 2^(level + 2) rows of pyramid support. */
int dic_stage_next_pair_rows(dic_engine *e, const uint8_t *und, const uint8_t *def, int rows, int cols,
                             int row_begin, int row_end);
/* ---- CudaClass::makeUndPyramidFromDef / makeDefPyramidFromNxt (cuda_class.cu:607-613):
 *      pointer rotation, no copy */
int dic_make_und_pyramid_from_def(dic_engine *e);
int dic_make_def_pyramid_from_nxt(dic_engine *e);

/* ---- CudaClass::resetPolygon x3 (cuda_class.cu:574-605).
 *      Membership follows the CPU engine's builders, not the reference GPU functors:
 *      rect     manager_class.cpp:1596-1614  all integer (x,y), x0<=x<=x1, y0<=y<=y1;
 *               centre = ((x0+x1)/2, (y0+y1)/2), the value the manager passes (:438-441)
 *      annular  manager_class.cpp:816-940    strict ri^2 < r^2 < ro^2, half-open box, wedge test
 *      blob     polygon_class.cpp            ear-clipped triangles, half-open scanline fill
 *      Returns DIC_ERROR_BAD_DOMAIN for a self-intersecting contour or an empty level. */
int dic_reset_polygon_rect(dic_engine *e, int iSector, int x0, int y0, int x1, int y1);
int dic_reset_polygon_annular(dic_engine *e, int iSector, float r, float dr, float a, float da,
                              float cx, float cy, int as);
int dic_reset_polygon_blob(dic_engine *e, int iSector, const float *contour_xy, int n_vertices);
/* extension (BASELINE config 5): this GPU takes only image rows [band_y0, band_y1] of the rectangle;
 * centre, point counts and chi scaling stay those of the whole rectangle. Use with dic_rowsplit_*. */
int dic_reset_polygon_rect_band(dic_engine *e, int iSector, int x0, int y0, int x1, int y1, int band_y0,
                                int band_y1);
/* extension (BASELINE config 4): n rectangles at once -- the result of n calls
 * dic_reset_polygon_rect(first_sector + k, boxes[4k], boxes[4k+1], boxes[4k+2], boxes[4k+3]), which is what the
 * subdivision loop of manager_class.cpp:274-336 issues on frame 0, built by one list kernel and one tile kernel
 * over all sectors and levels. Returns DIC_ERROR_BAD_DOMAIN if any rectangle is empty at a used level (those
 * sectors stay undefined, the others are valid). */
int dic_reset_polygon_rect_grid(dic_engine *e, int first_sector, int n, const int *boxes);
/* extension: an arbitrary point list (what CorrelationClass::Newton_Raphson(guess, N, xy) takes,
 * correlation_class.cpp:326-343). use_center: 0 = derive per the centre mode. */
int dic_reset_polygon_points(dic_engine *e, int iSector, const float *xy, int64_t n,
                             int use_center, float cx, float cy);
/* extension: override the centre of an existing sector */
int dic_set_polygon_center(dic_engine *e, int iSector, float cx, float cy);

/* ---- CudaClass::updatePolygon(iSector, deformationDescription) (cuda_class.cu:596-605,
 *      cuda_polygon.cu:268-415): Lagrangian = translate the list by the rounded centre shift of
 *      the last result; strict Lagrangian = und points := last deformed points; Eulerian = no-op */
int dic_update_polygon(dic_engine *e, int iSector, int deformation_description);

/* ---- CudaClass::correlate(iSector, guess, results) (cuda_class.cu:104-293).
 *      guess: level-0 units, n parameters; read, then overwritten with the result (as the
 *      reference does, cuda_class.cu:289-290). out may be NULL. Returns out->errorCode. */
int dic_correlate(dic_engine *e, int iSector, float *guess_inout, dic_result *out);
/* extension (BASELINE config 4): n_sectors consecutive sector ids, one CTA per sector, one
 * launch. guesses: n_sectors x n_params (row-major), overwritten; results: n_sectors entries. */
int dic_correlate_batch(dic_engine *e, int first_sector, int n_sectors, float *guesses_inout,
                        dic_result *results);
/* extension: how dic_correlate_batch maps sectors to thread blocks. 0 (default) = automatic: one CTA per
 * sector, or one CTA PAIR per sector (thread-block cluster of 2, partial sums exchanged through distributed
 * shared memory) when there are too few sectors to occupy the GPU's CTA slots at all; 1 = always one CTA; 2 = always a pair.
 * Results are identical to ~1 ulp of the sums (the two halves are added in a fixed order). */
int dic_set_cluster_mode(dic_engine *e, int mode);
/* extension: how a one-CTA-per-sector batch occupies the GPU. 1 = resident CTAs (as many as fit at once) that draw
 * further sectors from a launch-wide ticket counter; 2 = one CTA per sector, left to the hardware block scheduler,
 * so that kernels of the higher-priority image stream (the NEXT pair's pyramid build, dic_stage_next_pair) are
 * interleaved with the solve instead of waiting for all of it; 0 (default) = 1 (measured on BASELINE config 4: the
 * staged-pair loop is bound by the PCIe transfer in both forms, the resident kernel is within 1.5 %). The records do
 * not depend on the choice (one CTA, one instruction sequence per sector either way). */
int dic_set_batch_queue(dic_engine *e, int mode);
/* CTAs per sector the last dic_correlate_batch launch used (1 or 2) */
int dic_last_cluster_size(const dic_engine *e);
/* extension: enqueue only (no host sync); dic_correlate_wait collects. Lets a caller overlap
 * the next upload with the solve, and lets bench.py time the device alone. */
int dic_correlate_async(dic_engine *e, int iSector, const float *guess);
int dic_correlate_wait(dic_engine *e, int iSector, float *guess_out, dic_result *out);
/* the same split for a batch: dic_correlate_batch == dic_correlate_batch_async + dic_correlate_batch_wait. A frame loop
 * that calls `dic_advance_pair; dic_correlate_batch_async(k); dic_stage_next_pair(k + 1); dic_correlate_batch_wait(k)`
 * keeps the host work of the staging call off the solve's critical path. guesses_out / results may be NULL. */
int dic_correlate_batch_async(dic_engine *e, int first_sector, int n_sectors, const float *guesses);
int dic_correlate_batch_wait(dic_engine *e, int first_sector, int n_sectors, float *guesses_out, dic_result *results);

/* ---- extension: one domain row-split over `world` GPUs of one node (one process per GPU).
 *      Every evaluation ends with a sum of the normal equations over the ranks, done inside the
 *      persistent kernel through peer-mapped mailboxes (CUDA IPC, NVLink): no NCCL call, no host
 *      round trip, bitwise-identical totals on every rank. Protocol: each rank calls
 *      dic_rowsplit_mailbox_handle (64-byte opaque handle), the host all-gathers the handles
 *      (torch.distributed / MPI / anything), each rank calls dic_rowsplit_connect with all of them,
 *      then dic_correlate is called collectively by all ranks on sectors built with
 *      dic_reset_polygon_rect_band. A peer that does not answer within 20 s yields
 *      DIC_ERROR_MULTITHREAD instead of a hang. */
int dic_rowsplit_mailbox_handle(dic_engine *e, void *handle_out, int handle_bytes);
int dic_rowsplit_connect(dic_engine *e, int rank, int world, const void *handles, int handle_bytes);
int dic_rowsplit_disconnect(dic_engine *e);

/* ---- CudaClass::getUndXY0ToCPU / getDefXY0ToCPU (cuda_class.cu, cuda_polygon.cu:417-428):
 *      level-0 list in the reference CPU order. Writes min(cap, n) points (interleaved x,y);
 *      *n_needed receives n. */
int dic_get_und_xy0(dic_engine *e, int iSector, float *xy, int64_t cap, int64_t *n_needed);
int dic_get_def_xy0(dic_engine *e, int iSector, float *xy, int64_t cap, int64_t *n_needed);

/* ---- introspection used by the parity tests and bench (extensions) */
/* which: 0 und, 1 def, 2 nxt. out may be NULL to query the size. */
int dic_get_pyramid_level(dic_engine *e, int which, int level, uint8_t *out, int *rows, int *cols);
int dic_get_level_points(dic_engine *e, int iSector, int level, float *xy, int64_t cap,
                         int64_t *n_needed);
int dic_get_level_center(dic_engine *e, int iSector, int level, float *cx, float *cy);
/* one evaluation at `params` (LEVEL units): raw upper-triangular A (n x n row-major), b, chi
 * (sum V^2, unscaled), number of out-of-image pixels. */
int dic_evaluate(dic_engine *e, int iSector, int level, const float *params, float *A, float *b,
                 float *chi, int *n_out_of_image);
/* the damped solve of correlation_class.cpp:642-688 on caller data (device Cholesky) */
int dic_solve_step(dic_engine *e, const float *A_upper, const float *b, float lambda,
                   float scaling, float *dp);
/* device-side time of the last correlate / correlate_batch launch in milliseconds (CUDA events
 * on the correlation stream) and how many kernels this engine has launched so far */
float dic_last_correlate_ms(dic_engine *e);
/* the same over the whole GPU side of the call: solve kernel(s) and result download (the guess travels in the
 * kernel parameters / is read from pinned memory by the kernel: no upload) */
float dic_last_step_ms(dic_engine *e);
/* CTA 0's timeline of the last single-sector correlate: per evaluation 4 device timestamps (ns):
 * pass start, own pass done, all CTAs arrived, LM step published. Returns the evaluation count. */
int dic_get_timeline(dic_engine *e, unsigned long long *marks, int cap);
/* diagnostics of the staged-pair pipeline. on != 0 arms device-time marks for the next (at most 48) dic_stage_next_pair*
 * and correlate calls; every call first drains the device and, if marks were armed and the arrays are given, writes
 * per staging call {transfer may start, reference image landed, deformed image landed, both pyramids built}
 * (stage_ms, 4 floats each) and per correlate {kernel start, kernel end} (solve_ms, 2 floats each), in ms since the
 * arming call. Returns the number of staging calls written (<= cap), *n_solves the correlates. */
int dic_pipe_trace(dic_engine *e, int on, float *stage_ms, float *solve_ms, int cap, int *n_solves);
/* load-balance probe: per CTA of the last single-sector correlate, the device time (ns) at which its
 * pass of the last evaluation ended. Returns the number of entries written (<= cap, <= 1024). */
int dic_get_cta_times(dic_engine *e, unsigned long long *out, int cap);
/* same probe: the SM each CTA of the last single-sector correlate ran on */
int dic_get_cta_smids(dic_engine *e, unsigned int *out, int cap);
int64_t dic_kernel_launches(const dic_engine *e);
/* the CUDA stream handle (cudaStream_t as void*) the GN kernels run on, for event timing */
void *dic_correlation_stream(dic_engine *e);
int dic_synchronize(dic_engine *e);

#ifdef __cplusplus
}
#endif
#endif /* DIC_B200_H */
