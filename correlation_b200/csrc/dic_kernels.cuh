// dic_kernels.cuh -- the CUDA kernels of the B200 DIC engine (sm_100a).
//
//   gn_solve_kernel     one launch per correlate(): every pyramid level, every LM iteration, the
//                       reduction and the 6x6 / 12x12 solve stay on the device
//                       (replaces kCorrelation + k_global_reduction + k_build_LS_problem_in_GPU0 +
//                        cuSOLVER + kUpdateParameters + kScale and the per-iteration host sync of
//                        cuda_class.cu:104-473; semantics of correlation_class.cpp:349-640)
//   gn_eval_kernel      one evaluation only (parity tests: A, b, chi per evaluation)
//   pyramid_level_kernel  pyramid_class.cpp:52-134, bit-exact
//   list builders       manager_class.cpp:1596-1614 / :816-940, pyramid_class.cpp:289-323
#pragma once
#include <float.h>

#include "dic_device.cuh"

namespace dic {

constexpr int kAccStride = 96; // floats per CTA partial record (>= Acc<12>::kN = 92)
constexpr unsigned long long kSpinTimeoutNs = 10000000000ull; // 10 s: grid-barrier waits are bounded

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void red_release_add_u32(unsigned int *p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release_u32(unsigned int *p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

#ifndef DIC_BATCH_TIMELINE
#define DIC_BATCH_TIMELINE 0 // diagnostics only (tools/probe_batch_tl.py)
#endif

// ------------------------------------------------------------------ LM state machine

// scratch of warp_solve (NP * NP) followed by the step dp (NP): lm_step puts dp at kDpOffset
template <int NP> struct SolveScratch {
  static constexpr int kDpOffset = NP * (NP + 1);
  static constexpr int kFloats = kDpOffset + NP + 4 < 16 ? 16 : kDpOffset + NP + 4; // >= kMaxParams: also stages the guess
};


// Executed by ONE full warp after every evaluation. tot = the evaluation's reduced sums.
// Restates correlation_class.cpp:373-591 as "what to evaluate next"; parameter vectors are held
// one element per lane.
template <int MODEL>
__device__ void lm_step(LMState *s, const float *tot, const SolveSettings &cfg,
                        const SectorDev *sec, dic_result *result, float *smem, bool writer = true) {
  constexpr int NP = model_nparams(MODEL);
  using L = Acc<NP>;
  const int lane = threadIdx.x & 31;
  const bool act = lane < NP;
  const int li = act ? lane : 0;
  const float min_lambda = 1e-9f, max_lambda = 1e9f;

  int level = s->level, level_old = s->level_old, phase = s->phase;
  int iteration = s->iteration, use_saved = s->use_saved;
  int error_code = s->error_code, reached = s->reached_iterations;
  float lambda = s->lambda, last_good_chi = s->last_good_chi, scaling = s->scaling;
  float e_p = s->p[li], e_mp = s->mp[li], e_lg = s->last_good[li];
  float e_tent = s->tentative[li], e_saved = s->saved[li];
  int evals = s->evals[level] + 1, iters = s->iters[level];
  const int this_level = level;

  const float chi = tot[L::kChi] * scaling;
  const bool oob = tot[L::kOob] > 0.5f;
  float *dp = smem + SolveScratch<NP>::kDpOffset;
  bool end_level = false, finish = false, begin_iter = false;
  __syncwarp();

  if (phase == PH_INIT) {
    // correlation_class.cpp:410-439
    bool ok = !oob;
    if (ok) {
      last_good_chi = chi;
      ok = warp_solve<NP>(tot, scaling, lambda, smem, dp);
      if (!ok) error_code = DIC_ERROR_SOLVER;
    } else {
      error_code = DIC_ERROR_INTERPOLATION_OUT_OF_IMAGE;
    }
    if (!ok) { // :413-419: give up on the whole pyramid
      e_mp = e_p;
      level_old = level;
      finish = true;
    } else {
      e_saved = e_p + dp[li];
      e_mp = e_saved;
      use_saved = 1;
      iteration = 1;
      begin_iter = true;
    }
  } else if (phase == PH_REDO) {
    // :475-499, parameters evaluated were last_good
    bool ok = !oob;
    if (ok) {
      ok = warp_solve<NP>(tot, scaling, lambda, smem, dp);
      if (!ok) error_code = DIC_ERROR_SOLVER;
    } else {
      error_code = DIC_ERROR_INTERPOLATION_OUT_OF_IMAGE;
    }
    if (!ok) {
      e_mp = e_lg;
      end_level = true;
    } else {
      e_tent = e_lg + dp[li];
      e_mp = e_tent;
      e_p = e_tent;
      phase = PH_TENT;
    }
  } else {
    // :503-585, parameters evaluated were tentative
    bool ok = !oob;
    if (ok) {
      ok = warp_solve<NP>(tot, scaling, fmaxf(lambda * 0.4f, min_lambda), smem, dp);
      if (!ok) error_code = DIC_ERROR_SOLVER;
    } else {
      error_code = DIC_ERROR_INTERPOLATION_OUT_OF_IMAGE;
    }
    if (!ok) {
      e_mp = e_tent;
      end_level = true;
    } else {
      e_saved = e_tent + dp[li];
      e_mp = e_saved;
      float delta_chi = fabsf((last_good_chi - chi) / (fmaxf(last_good_chi, chi) + cfg.precision));
      if (chi <= last_good_chi) {
        last_good_chi = chi;
        lambda = fmaxf(lambda * 0.4f, min_lambda);
        e_lg = e_tent;
        use_saved = 1;
      } else {
        lambda = fminf(lambda * 10.0f, max_lambda);
        use_saved = 0;
      }
      if (delta_chi < cfg.precision) end_level = true;
      else { ++iteration; begin_iter = true; }
    }
  }

  if (begin_iter) { // :441-467
    if (iteration > cfg.max_iters || lambda >= max_lambda) {
      error_code = DIC_ERROR_MAX_ITERS_REACHED;
      end_level = true;
    } else {
      reached = iteration;
      iters = iteration;
      if (use_saved) { e_tent = e_saved; e_p = e_tent; phase = PH_TENT; }
      else { e_p = e_lg; phase = PH_REDO; }
    }
  }

  int next_level = level;
  if (end_level) { // :589 and the top of the level loop :373-407
    level_old = level;
    next_level = level - cfg.step;
    if (next_level < cfg.start) {
      finish = true;
    } else {
      e_mp = translate_param<MODEL>(e_mp, li, level_old, next_level);
      error_code = DIC_OK;
      lambda = 0.0001f;
      last_good_chi = FLT_MAX;
      e_lg = e_mp;
      e_p = e_mp;
      scaling = 1.f / (float)sec->n_total[next_level];
      phase = PH_INIT;
    }
  }
  if (finish) e_mp = translate_param<MODEL>(e_mp, li, level_old, 0); // :638 / :417

  __syncwarp();
  if (act) {
    s->p[li] = e_p; s->mp[li] = e_mp; s->last_good[li] = e_lg;
    s->tentative[li] = e_tent; s->saved[li] = e_saved;
    if (finish && writer) result->resultingParameters[li] = e_mp;
  }
  if (lane == 0) {
    s->evals[this_level] = evals;
    s->iters[this_level] = iters;
    s->level = finish ? level : next_level;
    s->level_old = level_old;
    s->phase = phase; s->iteration = iteration; s->use_saved = use_saved;
    s->error_code = error_code; s->reached_iterations = reached;
    s->lambda = lambda; s->last_good_chi = last_good_chi; s->scaling = scaling;
    s->done = finish ? 1 : 0;
    if (finish && writer) {
      for (int i = NP; i < kMaxParams; ++i) result->resultingParameters[i] = 0.f;
      result->chi = last_good_chi;
      result->numberOfPoints = sec->n_total[0];
      result->iterations = reached;
      result->errorCode = error_code;
      result->undCenterX = sec->cx;
      result->undCenterY = sec->cy;
      for (int l = 0; l < kMaxLevels; ++l) {
        result->iterationsPerLevel[l] = l == this_level ? iters : s->iters[l];
        result->evaluationsPerLevel[l] = l == this_level ? evals : s->evals[l];
        result->pointsPerLevel[l] = (l >= cfg.start && l <= cfg.stop && (l - cfg.start) % cfg.step == 0)
                                        ? sec->n_total[l] : 0;
      }
    }
  }
  __syncwarp();
}

// Initial LM state (top of correlation_class.cpp:349-407 for the coarsest level).
template <int MODEL>
__device__ void lm_init(LMState *s, const SolveSettings &cfg, const SectorDev *sec,
                        const float *guess) {
  constexpr int NP = model_nparams(MODEL);
  const int lane = threadIdx.x & 31;
  if (lane < NP) {
    float v = translate_param<MODEL>(guess[lane], lane, 0, cfg.stop);
    s->p[lane] = v; s->mp[lane] = v; s->last_good[lane] = v; s->tentative[lane] = v;
    s->saved[lane] = v;
  }
  if (lane == 0) {
    s->lambda = 0.0001f; s->last_good_chi = FLT_MAX;
    s->scaling = 1.f / (float)sec->n_total[cfg.stop];
    s->level = cfg.stop; s->level_old = 0; s->iteration = 0; s->use_saved = 1;
    s->phase = PH_INIT; s->done = 0; s->error_code = DIC_OK; s->reached_iterations = 0;
    for (int l = 0; l < kMaxLevels; ++l) { s->evals[l] = 0; s->iters[l] = 0; }
  }
  __syncwarp();
}

// ------------------------------------------------------------------ one evaluation pass

template <int MODEL, int INTERP, int MODE>
__device__ __forceinline__ void evaluate_list(const SolveSettings &cfg, const SectorDev *sec,
                                              int level, const float *p, long first, long stride,
                                              float *acc) {
  const LevelImage und = cfg.und[level];
  const LevelImage def = cfg.def[level];
  const float2 *__restrict__ xy = sec->xy[level];
  const long n = sec->n[level];
  const float inv = 1.f / (float)(1 << level); // pyramid_class.cpp:357-361
  const float cx = sec->cx * inv, cy = sec->cy * inv;
  if (MODE == DIC_MODE_PARITY && und.colors == 3) { // colour images: generic path only, reference arithmetic only
    for (long i = first; i < n; i += stride) {
      float2 q = __ldg(xy + i);
      accumulate_pixel_color<MODEL, INTERP, MODE>(und, def, p, cx, cy, q.x, q.y, acc);
    }
    return;
  }
  for (long i = first; i < n; i += stride) {
    float2 q = __ldg(xy + i);
    accumulate_pixel<MODEL, INTERP, MODE>(und, def, p, cx, cy, q.x, q.y, acc);
  }
}

// ------------------------------------------------------------------ row-split all-reduce
//
// One domain split by pixel rows over several GPUs: each evaluation ends with a sum of the
// (n^2 + n)/2 + n + 2 normal-equation values over the ranks. Done here, inside the persistent kernel:
//   send     CTA 0, after the rank's own grid all-reduce: warp w stores the rank's sums into slot
//            [parity][rank] of peer w's mailbox (plain stores to peer-mapped memory: NVLink), system
//            fence, release-store of the sequence number -- one warp per peer, all peers in parallel;
//   receive  EVERY CTA polls the rank's own mailbox (local HBM, written by the peers) for all ranks'
//            sequence numbers and adds the rows IN RANK ORDER itself: every CTA of every rank gets bitwise
//            the same totals and takes the same LM decisions -- no broadcast, no publish / re-read hop
//            through a master CTA.
// Two parities: a rank cannot be more than one evaluation ahead of a peer, because it needs that
// peer's sums of the current evaluation to get there, and a peer sends those only after all of its
// CTAs have passed its grid barrier, i.e. have finished reading the previous evaluation's slots.
__device__ __forceinline__ unsigned int ld_acquire_sys_u32(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u32(unsigned int *p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float ld_relaxed_sys_f32(const float *p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
constexpr unsigned long long kPeerTimeoutNs = 20000000000ull; // 20 s: a peer that never answers is an error, not a hang

// CTA 0 only, all of its warps: warp w serves peers w, w + 8, ...
template <int NACC>
__device__ __forceinline__ void rowsplit_send(GridWork *work, const float *tot, unsigned int seq) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rank = work->rs_rank, world = work->rs_world;
  const int par = seq & 1;
  for (int r = warp; r < world; r += (int)(blockDim.x >> 5)) {
    Mailbox *mb = work->rs_peer[r];
    for (int k = lane; k < NACC; k += 32) mb->sums[par][rank][k] = tot[k];
    __threadfence_system();
    __syncwarp();
    if (lane == 0) st_release_sys_u32(&mb->seq[par][rank], seq + 1);
  }
}
// every CTA, all threads. Returns false (CTA-uniform) when a peer did not answer in time.
template <int NACC>
__device__ __forceinline__ bool rowsplit_receive(GridWork *work, float *tot, unsigned int seq, int *s_flag) {
  const int tid = threadIdx.x;
  const int world = work->rs_world;
  const int par = seq & 1;
  const Mailbox *mb = work->rs_local;
  if (tid == 0) *s_flag = 1;
  __syncthreads();
  if (tid < world) {
    const unsigned int *flag = &mb->seq[par][tid];
    unsigned long long t0 = 0;
    unsigned int spins = 0;
    while (ld_acquire_sys_u32(flag) != seq + 1) {
      if ((++spins & 0x3ffu) == 0) {
        const unsigned long long now = global_timer_ns();
        if (t0 == 0) t0 = now;
        else if (now - t0 > kPeerTimeoutNs) { *s_flag = 0; break; }
      }
    }
  }
  __syncthreads();
  const bool ok = *s_flag != 0;
  if (ok && tid < NACC) {
    double s = 0.0;
    for (int r = 0; r < world; ++r) s += (double)ld_relaxed_sys_f32(&mb->sums[par][r][tid]);
    tot[tid] = (float)s;
  }
  __syncthreads();
  return ok;
}

// ------------------------------------------------------------------ thread-block cluster helpers
// A subset solved by a PAIR of CTAs (cluster of 2, one CTA per SM): each CTA evaluates half of the
// subset's units, stores its sums into BOTH CTAs' shared memory (its own and, through distributed
// shared memory, the partner's), one cluster barrier, and both add the two rows in rank order --
// bitwise identical totals, both run the same LM step. Halves the scheduling granule of a batch of
// subsets (fills the tail wave when a GPU holds only a few hundred subsets).
__device__ __forceinline__ unsigned int cluster_ctarank() {
  unsigned int r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_dsmem_f32(const float *local_addr, unsigned int cta, float v) {
  unsigned int remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local_addr)), "r"(cta));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(v) : "memory");
}

// Shared-memory block common to both solve kernels.
template <int NP> struct SolveShared {
  static constexpr int NACC = Acc<NP>::kN;
  float tot[NACC];
  float solve[SolveScratch<NP>::kFloats];
  float p[kMaxParams];
  float xch[2][2][NACC]; // cluster pair: [evaluation parity][source CTA] partial sums
  int level, done;
  int rowsplit;  // this launch exchanges its sums with other GPUs (read once per sector, not per evaluation)
  int mark;      // next slot of CTA 0's timeline
  int timed_out; // a bounded wait expired
  int rs_ok;     // row-split: scratch of rowsplit_receive
  unsigned int arrive_target; // value of GridWork::arrive that completes the current evaluation
  unsigned int rs_seq;        // row-split: evaluations exchanged so far (same on all CTAs of all ranks)
  unsigned int xch_count;     // cluster pair: exchanges done by this CTA (parity of the xch buffer)
  LMState state;
};
static_assert(SolveScratch<12>::kDpOffset + 12 <= SolveScratch<12>::kFloats, "dp must lie inside SolveShared::solve");

// After one evaluation pass: sh.tot holds this CTA's sums. Produces the next command in
// sh.p / sh.level / sh.done for every CTA.
//   GRID: all-reduce through GridWork (see there), then every CTA's warp 0 runs the LM step + solve
//   on its own state; CTA 0 alone writes the result record and the timeline. With a row-split over
//   GPUs, CTA 0 additionally sends the rank's sums to the peers and every CTA adds all ranks' rows.
//   Batch: the CTA (CL = 1) or the CTA pair (CL = 2) owns the sector.
template <int MODEL, bool GRID, int CL = 1>
__device__ __forceinline__ void reduce_and_step(SolveShared<model_nparams(MODEL)> &sh, bool active,
                                                int n_active, const SolveSettings &cfg,
                                                const SectorDev *sec, dic_result *result,
                                                GridWork *work, unsigned int &my_gen) {
  constexpr int NP = model_nparams(MODEL);
  constexpr int NACC = Acc<NP>::kN;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (GRID) {
    const bool lead = blockIdx.x == 0;
    double *acc = work->acc[my_gen % 3u];
    if (active && tid < NACC) atomicAdd(&acc[tid], (double)sh.tot[tid]);
    __syncthreads();
    if (tid == 0) {
      const unsigned long long t0 = global_ns();
      const int m = sh.mark;
      if (lead && m < kMaxMarks) work->marks[m][1] = t0;
      if (blockIdx.x < kMaxCtaMarks) work->cta_done[blockIdx.x] = t0;
      // release RMW by one thread after the CTA barrier: cumulative over the other threads' atomics
      // (and over their reads of the previous evaluation's totals)
      red_release_add_u32(&work->arrive, 1u);
      const unsigned int target = sh.arrive_target + gridDim.x;
      sh.arrive_target = target;
      // bounded wait: a lost CTA must never hang the GPU (the launch then ends with error_cuda)
      unsigned int spins = 0;
      while ((int)(ld_acquire_u32(&work->arrive) - target) < 0) {
        if ((++spins & 0xffffu) == 0 && (global_ns() - t0 > kSpinTimeoutNs || *(volatile int *)&work->abort)) {
          atomicExch(&work->abort, 1); sh.timed_out = 1; break;
        }
      }
      if (lead && m < kMaxMarks) work->marks[m][2] = global_ns();
    }
    __syncthreads();
    // order-independent to well below fp32 ulp; every CTA reads the same bits
    if (tid < NACC) sh.tot[tid] = (float)__ldcg(&acc[tid]);
    if (lead && tid < NACC) __stcg(&work->acc[(my_gen + 2u) % 3u][tid], 0.0);
    __syncthreads();
    if (sh.rowsplit) {
      const unsigned int seq = sh.rs_seq;
      if (lead) rowsplit_send<NACC>(work, sh.tot, seq);
      if (!rowsplit_receive<NACC>(work, sh.tot, seq, &sh.rs_ok) && tid == 0) sh.timed_out = 2;
      if (tid == 0) sh.rs_seq = seq + 1;
      __syncthreads();
    }
    if (warp == 0) {
      lm_step<MODEL>(&sh.state, sh.tot, cfg, sec, result, sh.solve, lead);
      if (lane == 0) {
        if (sh.timed_out == 2) { sh.state.done = 1; if (lead) result->errorCode = DIC_ERROR_MULTITHREAD; } // a peer never answered
        else if (sh.timed_out) { sh.state.done = 1; if (lead) result->errorCode = DIC_ERROR_CUDA; } // a CTA never arrived / a TMA copy never landed
      }
      __syncwarp();
      if (lane < NP) sh.p[lane] = sh.state.p[lane];
      if (lane == 0) {
        sh.level = sh.state.level; sh.done = sh.state.done;
        if (lead) {
          const int m = sh.mark;
          const unsigned long long t3 = global_ns();
          if (m < kMaxMarks) { work->marks[m][3] = t3; work->n_marks = m + 1; }
          if (m + 1 < kMaxMarks) work->marks[m + 1][0] = t3;
          sh.mark = m + 1;
        }
      }
    }
    ++my_gen;
    __syncthreads();
  } else {
    bool writer = true;
    if (CL == 2) {
      // exchange with the partner CTA through distributed shared memory
      const unsigned int me = cluster_ctarank();
      const int par = sh.xch_count & 1;
      if (tid < NACC) {
        const float v = sh.tot[tid];
        sh.xch[par][me][tid] = v;
        st_dsmem_f32(&sh.xch[par][me][tid], me ^ 1u, v);
      }
      cluster_sync_all(); // release / acquire at cluster scope: both rows are visible to both CTAs
      if (tid < NACC) sh.tot[tid] = sh.xch[par][0][tid] + sh.xch[par][1][tid];
      if (tid == 0) sh.xch_count += 1;
      __syncthreads();
      writer = me == 0;
    }
    if (warp == 0) {
#if DIC_BATCH_TIMELINE
      if (blockIdx.x == 0 && lane == 0 && sh.mark < kMaxMarks) work->marks[sh.mark][2] = global_ns();
#endif
      lm_step<MODEL>(&sh.state, sh.tot, cfg, sec, result, sh.solve, writer);
#if DIC_BATCH_TIMELINE
      if (blockIdx.x == 0 && lane == 0 && sh.mark < kMaxMarks) work->marks[sh.mark][3] = global_ns();
#endif
      if (lane == 0 && sh.timed_out) { sh.state.done = 1; if (writer) result->errorCode = DIC_ERROR_CUDA; }
      __syncwarp();
      if (lane < NP) sh.p[lane] = sh.state.p[lane];
      if (lane == 0) { sh.level = sh.state.level; sh.done = sh.state.done; }
    }
    __syncthreads();
  }
}

// Initial guess of a single-sector (grid) launch: travels as a kernel parameter, so that a correlate() enqueues
// nothing on a copy engine. Batch launches read their guesses straight from the caller-visible pinned block.
struct GuessParam { float v[kMaxParams]; };

__device__ __forceinline__ float ld_guess(const float *p) { // pinned host memory: never from a stale cache line
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

// Start of a sector: every CTA derives the first command locally and initialises its own copy of the
// LM state (grid mode: all CTAs keep identical copies; batch mode: the CTA or CTA pair owns the sector).
template <int MODEL, bool GRID>
__device__ __forceinline__ void begin_sector(SolveShared<model_nparams(MODEL)> &sh, const SolveSettings &cfg,
                                             const SectorDev *sec, const float *guess, const GuessParam &g0,
                                             GridWork *work, unsigned int &my_gen) {
  constexpr int NP = model_nparams(MODEL);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid < kMaxParams) sh.solve[tid] = GRID ? g0.v[tid] : (tid < NP ? ld_guess(guess + tid) : 0.f);
  __syncthreads();
  if (warp == 0) lm_init<MODEL>(&sh.state, cfg, sec, sh.solve); // grid mode: every CTA keeps its own copy
  if (tid == 0) {
    sh.rowsplit = GRID && work->rs_local != nullptr;
    sh.rs_seq = sh.rowsplit ? work->rs_seq : 0u;
    sh.mark = 0; sh.arrive_target = 0;
    if (work->img_error) sh.timed_out = 1; // a pyramid transfer failed earlier: finish at once with error_cuda
  }
  if (GRID && blockIdx.x == 0 && tid == 0) { work->n_marks = 0; work->slow_units = 0; work->marks[0][0] = global_ns(); }
  if (GRID && tid == 0 && blockIdx.x < kMaxCtaMarks) {
    unsigned int smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    work->cta_smid[blockIdx.x] = smid;
  }
  if (tid < NP) sh.p[tid] = translate_param<MODEL>(sh.solve[tid], tid, 0, cfg.stop);
  if (tid == 0) { sh.level = cfg.stop; sh.done = 0; }
  __syncthreads();
  my_gen = 0u; // evaluations done by this launch
}

// End of a grid launch: the last CTA to leave puts GridWork back into its between-launch state.
template <int NACC>
__device__ __forceinline__ void grid_depart(GridWork *work, unsigned int rs_seq, bool rowsplit) {
  __syncthreads();
  __shared__ int s_last;
  if (threadIdx.x == 0) {
    if (rowsplit && blockIdx.x == 0) work->rs_seq = rs_seq;
    __threadfence();
    s_last = atomicAdd(&work->departed, 1u) == gridDim.x - 1 ? 1 : 0;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    for (int k = threadIdx.x; k < 3 * 96; k += blockDim.x) (&work->acc[0][0])[k] = 0.0;
    if (threadIdx.x == 0) { work->arrive = 0; work->abort = 0; work->departed = 0; }
  }
}

// GRID = true : every CTA of a cooperative launch works on ONE sector (large domains).
// GRID = false: each CTA owns whole sectors (BASELINE config 4: thousands of small subsets).
template <int MODEL, int INTERP, int MODE, bool GRID>
__global__ void __launch_bounds__(kThreads)
gn_solve_kernel(const SolveSettings cfg, const SectorDev *__restrict__ sectors,
                const float *guesses, const GuessParam guess0, dic_result *__restrict__ results, int first_sector,
                int n_sectors, GridWork *work) {
  constexpr int NP = model_nparams(MODEL);
  constexpr int NACC = Acc<NP>::kN;
  __shared__ float s_red[(kThreads / 32) * NACC];
  __shared__ SolveShared<NP> sh;
  const int tid = threadIdx.x;
  if (tid == 0) { sh.timed_out = 0; sh.xch_count = 0; }

  for (int si = GRID ? 0 : blockIdx.x; si < n_sectors; si += GRID ? n_sectors : gridDim.x) {
    const SectorDev *sec = sectors + first_sector + si;
    const float *guess = guesses + (size_t)(first_sector + si) * kMaxParams;
    dic_result *result = results + first_sector + si;
    unsigned int my_gen;
    begin_sector<MODEL, GRID>(sh, cfg, sec, guess, guess0, work, my_gen);
    while (true) {
      const int level = sh.level;
      float p[NP];
#pragma unroll
      for (int i = 0; i < NP; ++i) p[i] = sh.p[i];
      float acc[NACC];
#pragma unroll
      for (int k = 0; k < NACC; ++k) acc[k] = 0.f;
      int n_active = 1;
      bool active = true;
      if (GRID) {
        long need = ((long)sec->n[level] + kThreads * 4 - 1) / (kThreads * 4);
        n_active = (int)max(1l, min(need, (long)gridDim.x));
        active = (int)blockIdx.x < n_active;
        if (active)
          evaluate_list<MODEL, INTERP, MODE>(cfg, sec, level, p, (long)blockIdx.x * kThreads + tid,
                                             (long)n_active * kThreads, acc);
      } else {
        evaluate_list<MODEL, INTERP, MODE>(cfg, sec, level, p, tid, kThreads, acc);
      }
      __syncthreads();
      block_reduce<NACC>(acc, s_red, sh.tot);
      reduce_and_step<MODEL, GRID>(sh, active, n_active, cfg, sec, result, work, my_gen);
      if (sh.done) break;
    }
    __syncthreads();
  }
  if (GRID) grid_depart<NACC>(work, sh.rs_seq, sh.rowsplit != 0);
}

// ------------------------------------------------------------------ single evaluation (tests)

template <int MODEL, int INTERP, int MODE>
__global__ void __launch_bounds__(kThreads)
gn_eval_kernel(const SolveSettings cfg, const SectorDev *__restrict__ sec, int level,
               const float *__restrict__ params, float *partials) {
  constexpr int NP = model_nparams(MODEL);
  constexpr int NACC = Acc<NP>::kN;
  __shared__ float s_red[(kThreads / 32) * NACC];
  __shared__ float s_tot[NACC];
  float p[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) p[i] = params[i];
  float acc[NACC];
#pragma unroll
  for (int k = 0; k < NACC; ++k) acc[k] = 0.f;
  evaluate_list<MODEL, INTERP, MODE>(cfg, sec, level, p, (long)blockIdx.x * kThreads + threadIdx.x,
                                     (long)gridDim.x * kThreads, acc);
  block_reduce<NACC>(acc, s_red, s_tot);
  if (threadIdx.x < NACC) partials[(size_t)blockIdx.x * kAccStride + threadIdx.x] = s_tot[threadIdx.x];
}

// out[k] = sum over CTAs of partials[c][k], fixed order, double accumulation
__global__ void sum_partials_kernel(const float *partials, int n_cta, int nacc, float *out) {
  int k = threadIdx.x;
  if (k >= nacc) return;
  double s = 0.0;
  for (int c = 0; c < n_cta; ++c) s += (double)partials[(size_t)c * kAccStride + k];
  out[k] = (float)s;
}

template <int NP>
__global__ void solve_step_kernel(const float *tot, float scaling, float lambda, float *dp_out,
                                  int *ok_out) {
  __shared__ float smem[SolveScratch<NP>::kFloats];
  __shared__ float s_tot[NP * (NP + 1) / 2 + NP + 2];
  for (int i = threadIdx.x; i < NP * (NP + 1) / 2 + NP; i += 32) s_tot[i] = tot[i];
  __syncwarp();
  float *dp = smem + SolveScratch<NP>::kDpOffset;
  bool ok = warp_solve<NP>(s_tot, scaling, lambda, smem, dp);
  if (threadIdx.x < NP) dp_out[threadIdx.x] = ok ? dp[threadIdx.x] : 0.f;
  if (threadIdx.x == 0) *ok_out = ok ? 1 : 0;
}

// ------------------------------------------------------------------ pyramid

// pyramid_class.cpp:52-134. Persistent CTAs walk over 32 x 32 target tiles; one thread = 4 vertically
// adjacent targets. The (96 x 67) u8 source footprint of the NEXT tile is fetched by TMA
// (cp.async.bulk.tensor.2d, out-of-image = 0) into a double buffer while the current tile is computed.
// The landed u8 tile is converted once to fp32 in shared memory (PRMT + FADD -- no I2F), split into
// even / odd source columns so that the stride-2 taps of neighbouring lanes hit distinct banks.
// A thread walks down its 11 source rows, loads the five taps of a row once and feeds every target whose
// 5 x 5 window contains that row: each target still runs the reference's 25 sequential fp32 mul + add in
// the reference's order (dj outer, di inner, no FMA contraction), the four chains are independent (ILP 4)
// and a tap is read from shared memory 55 / 4 times per target instead of 25. Truncation to u8; target
// border rows / columns are written as 0 (the reference leaves a zero-initialised border, :98-102).
// HBM-bound by design (1.25 B per source pixel), FP32-issue bound in practice: 50 unfused ops per target.
constexpr int kPyrTX = 32, kPyrTY = 8, kPyrK = 4; // threads x, threads y, targets per thread
constexpr int kPyrTH = kPyrTY * kPyrK;            // 32 target rows per tile
constexpr int kPyrSW = 96, kPyrSH = 2 * kPyrTH + 3; // staged source window: 96 columns x 67 rows
constexpr int kPyrSrcBytes = (kPyrSW * kPyrSH + 127) / 128 * 128; // one u8 buffer, 128-byte multiple
constexpr int kPyrHalf = kPyrSW / 2;              // floats per row of each fp32 half
constexpr size_t kPyrSmem = 2 * (size_t)kPyrSrcBytes + 2 * sizeof(float) * kPyrSH * kPyrHalf;
struct PyrWeights { float w[25]; };

__global__ void __launch_bounds__(kPyrTX *kPyrTY)
pyramid_level_kernel(const __grid_constant__ CUtensorMap src_map, uint8_t *__restrict__ dst, int drows, int dcols,
                     int dpitch, PyrWeights kw, int trow_begin, int trow_end, int *err_flag) {
  extern __shared__ __align__(128) uint8_t pyr_smem[];
  uint8_t *raw = pyr_smem;                                                        // [2][kPyrSrcBytes]
  float(*tE)[kPyrHalf] = reinterpret_cast<float(*)[kPyrHalf]>(pyr_smem + 2 * kPyrSrcBytes); // even source columns
  float(*tO)[kPyrHalf] = tE + kPyrSH;                                             // odd source columns
  __shared__ __align__(8) uint64_t bar[2];
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * kPyrTX + tx;
  // only target rows [trow_begin, trow_end) are produced (a GPU that holds a band of the image)
  const int ntx = (dcols + kPyrTX - 1) / kPyrTX, nty = (trow_end - trow_begin + kPyrTH - 1) / kPyrTH;
  const int n_tiles = ntx * nty;
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  // staged window of a tile: columns from 2 ox - 16 (16-byte aligned for TMA), rows from 2 oy - 2
  auto issue = [&](int tile, int b) {
    const int bx = tile % ntx, by = tile / ntx;
    mbar_expect_tx(&bar[b], kPyrSW * kPyrSH);
    tma_load_2d(raw + b * kPyrSrcBytes, &src_map, 2 * bx * kPyrTX - 16, 2 * (trow_begin + by * kPyrTH) - 2, &bar[b]);
  };
  int tile = blockIdx.x;
  if (tile < n_tiles && tid == 0) issue(tile, 0);
  for (int it = 0; tile < n_tiles; ++it, tile += gridDim.x) {
    const int b = it & 1;
    // buffer b ^ 1 was last read by the conversion of iteration it - 1, which every thread left through the
    // barrier that follows it
    if (tile + (int)gridDim.x < n_tiles && tid == 0) issue(tile + gridDim.x, b ^ 1);
    if (!mbar_wait(&bar[b], (it >> 1) & 1)) { // the copy never landed: flag it (GridWork::img_error) and leave, all threads alike
      if (tid == 0) atomicExch(err_flag, 1);
      break;
    }
    const uint8_t *rb = raw + b * kPyrSrcBytes;
    for (int idx = tid; idx < kPyrSH * (kPyrSW / 16); idx += kPyrTX * kPyrTY) {
      const int r = idx / (kPyrSW / 16), c16 = idx - r * (kPyrSW / 16);
      const uint4 v = *reinterpret_cast<const uint4 *>(rb + r * kPyrSW + 16 * c16);
      float4 *e = reinterpret_cast<float4 *>(&tE[r][8 * c16]), *o = reinterpret_cast<float4 *>(&tO[r][8 * c16]);
      e[0] = make_float4(u8_to_float(v.x, 0), u8_to_float(v.x, 2), u8_to_float(v.y, 0), u8_to_float(v.y, 2));
      o[0] = make_float4(u8_to_float(v.x, 1), u8_to_float(v.x, 3), u8_to_float(v.y, 1), u8_to_float(v.y, 3));
      e[1] = make_float4(u8_to_float(v.z, 0), u8_to_float(v.z, 2), u8_to_float(v.w, 0), u8_to_float(v.w, 2));
      o[1] = make_float4(u8_to_float(v.z, 1), u8_to_float(v.z, 3), u8_to_float(v.w, 1), u8_to_float(v.w, 3));
    }
    __syncthreads();
    const int bx = tile % ntx, by = tile / ntx;
    const int ti = bx * kPyrTX + tx, tj0 = trow_begin + by * kPyrTH + ty * kPyrK;
    if (ti < dcols) {
      float acc[kPyrK];
#pragma unroll
      for (int t = 0; t < kPyrK; ++t) acc[t] = 0.f;
      // source column of tap di: 2 ti - 2 + di = window column 2 tx + 14 + di: even di -> tE[tx + 7 + di/2], odd -> tO[tx + 7 + (di-1)/2]
#pragma unroll
      for (int r = 0; r < 2 * kPyrK + 3; ++r) {
        const float *rE = tE[2 * ty * kPyrK + r] + tx + 7, *rO = tO[2 * ty * kPyrK + r] + tx + 7;
        const float s0 = rE[0], s1 = rO[0], s2 = rE[1], s3 = rO[1], s4 = rE[2];
#pragma unroll
        for (int t = 0; t < kPyrK; ++t) {
          const int dj = r - 2 * t; // row of target t's window
          if (dj >= 0 && dj < 5) {
            acc[t] = __fadd_rn(acc[t], __fmul_rn(s0, kw.w[dj * 5 + 0]));
            acc[t] = __fadd_rn(acc[t], __fmul_rn(s1, kw.w[dj * 5 + 1]));
            acc[t] = __fadd_rn(acc[t], __fmul_rn(s2, kw.w[dj * 5 + 2]));
            acc[t] = __fadd_rn(acc[t], __fmul_rn(s3, kw.w[dj * 5 + 3]));
            acc[t] = __fadd_rn(acc[t], __fmul_rn(s4, kw.w[dj * 5 + 4]));
          }
        }
      }
#pragma unroll
      for (int t = 0; t < kPyrK; ++t) {
        const int tj = tj0 + t;
        if (tj >= drows || tj >= trow_end) break;
        const bool interior = ti >= 1 && tj >= 1 && ti < dcols - 1 && tj < drows - 1;
        dst[(size_t)tj * dpitch + ti] = interior ? (uint8_t)__float2uint_rz(acc[t]) : (uint8_t)0;
      }
    }
    __syncthreads(); // the fp32 tile is free for the next conversion
  }
}

// pyramid_class.cpp:52-134 for three interleaved channels (the reference GPU's k_pyramid_color, kernels.cu:841-918,
// is what this replaces; the arithmetic is the CPU engine's): per target pixel and channel the same 25 sequential
// fp32 multiply-adds in dj-outer / di-inner order, truncation, zero border. Not a hot kernel (colour images take
// the generic pixel-list path): one thread per target pixel, plain loads.
__global__ void pyramid_color_kernel(const uint8_t *__restrict__ src, int spitch, uint8_t *__restrict__ dst, int drows,
                                     int dcols, int dpitch, PyrWeights kw) {
  const int ti = blockIdx.x * blockDim.x + threadIdx.x, tj = blockIdx.y;
  if (ti >= dcols || tj >= drows) return;
  const bool interior = ti >= 1 && tj >= 1 && ti < dcols - 1 && tj < drows - 1;
  for (int c = 0; c < 3; ++c) {
    float addition = 0.f;
    if (interior) {
      for (int dj = -2; dj <= 2; ++dj)
        for (int di = -2; di <= 2; ++di) {
          const float sv = (float)__ldg(src + (size_t)(2 * tj + dj) * spitch + (2 * ti + di) * 3 + c);
          addition = __fadd_rn(addition, __fmul_rn(sv, kw.w[(2 + dj) * 5 + (2 + di)]));
        }
    }
    dst[(size_t)tj * dpitch + ti * 3 + c] = interior ? (uint8_t)__float2uint_rz(addition) : (uint8_t)0;
  }
}

// level-0 upload helper: tightly packed (or pitched) u8 rows -> engine pitch
__global__ void copy_rows_kernel(const uint8_t *__restrict__ src, int spitch, uint8_t *__restrict__ dst,
                                 int dpitch, int rows, int cols) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y;
  if (x < cols && y < rows) dst[(size_t)y * dpitch + x] = src[(size_t)y * spitch + x];
}

// ------------------------------------------------------------------ pixel lists

// Rectangle, level list in row-major order. Keeps level-0 pixels whose coordinates are multiples
// of `mag` (pyramid_class.cpp:306-316) and scales them by 1/mag.
__global__ void rect_fill_kernel(float2 *__restrict__ out, long n, int xs, int ys, int nx, int mag) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int row = (int)(i / nx), col = (int)(i % nx);
  float inv = 1.f / (float)mag;
  out[i] = make_float2((float)(xs + col * mag) * inv, (float)(ys + row * mag) * inv);
}

// Annulus / annular sector membership, manager_class.cpp:897-919, evaluated in fp32 with the
// reference's operation order (no contraction). Candidates are the box pixels whose coordinates
// are multiples of `mag`, enumerated row-major.
struct AnnulusPred {
  int xs, ys, nxc, mag, as;
  float cx, cy, ri2, ro2;
  float c00x, c01x, c10x, c11x, c00y, c01y, c10y, c11y;
  __device__ __forceinline__ bool operator()(long idx, float2 &out) const {
    int row = (int)(idx / nxc), col = (int)(idx % nxc);
    float i = (float)(xs + col * mag);
    float j = (float)(ys + row * mag);
    float ax = __fsub_rn(i, cx), ay = __fsub_rn(j, cy);
    float r2 = __fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay));
    if (!(r2 > ri2 && r2 < ro2)) return false;
    if (as != 1) {
      float cross1 = __fsub_rn(__fmul_rn(__fsub_rn(c11x, i), __fsub_rn(c01y, c11y)),
                               __fmul_rn(__fsub_rn(c11y, j), __fsub_rn(c01x, c11x)));
      float cross2 = __fsub_rn(__fmul_rn(__fsub_rn(c00x, i), __fsub_rn(c10y, c00y)),
                               __fmul_rn(__fsub_rn(c00y, j), __fsub_rn(c10x, c00x)));
      if (!(__fmul_rn(cross1, cross2) > 0.f)) return false;
    }
    float inv = 1.f / (float)mag;
    out = make_float2(i * inv, j * inv);
    return true;
  }
};

// Generic list decimation, pyramid_class.cpp:306-316.
struct DecimatePred {
  const float2 *src;
  int mag;
  __device__ __forceinline__ bool operator()(long idx, float2 &out) const {
    float2 q = src[idx];
    int ix = (int)(q.x + 0.5f), iy = (int)(q.y + 0.5f);
    if (ix % mag != 0 || iy % mag != 0) return false;
    float inv = 1.f / (float)mag;
    out = make_float2(q.x * inv, q.y * inv);
    return true;
  }
};

// Order-preserving stream compaction: count pass, scan of CTA counts, emit pass.
constexpr int kCompactThreads = 256, kCompactItems = 8;
constexpr int kCompactChunk = kCompactThreads * kCompactItems;

template <class Pred, bool EMIT>
__global__ void __launch_bounds__(kCompactThreads)
compact_kernel(Pred pred, long ncand, unsigned int *block_counts,
               const unsigned long long *block_offsets, float2 *__restrict__ out) {
  __shared__ unsigned int warp_tot[kCompactThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long base = (long)blockIdx.x * kCompactChunk;
  unsigned long long running = EMIT ? block_offsets[blockIdx.x] : 0ull;
  unsigned int total = 0;
  for (int it = 0; it < kCompactItems; ++it) {
    long idx = base + (long)it * kCompactThreads + tid;
    float2 q;
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

// exclusive scan of CTA counts (single CTA), total written to offsets[n]
__global__ void scan_counts_kernel(const unsigned int *counts, int n, unsigned long long *offsets) {
  __shared__ unsigned long long buf[1024];
  __shared__ unsigned long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    int i = base + threadIdx.x;
    unsigned long long v = i < n ? counts[i] : 0;
    buf[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      unsigned long long t = threadIdx.x >= o ? buf[threadIdx.x - o] : 0;
      __syncthreads();
      buf[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < n) offsets[i] = carry + buf[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += buf[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) offsets[n] = carry;
}

// Blob: row spans (y, x_begin, x_end) produced on the host from the ear-clipped triangles
// (polygon_class.cpp:339-403) are expanded to pixels in the reference's emission order.
struct Span { int y, xb, xe; long offset; };
__global__ void expand_spans_kernel(const Span *__restrict__ spans, int n_spans, long n_pixels,
                                    float2 *__restrict__ out) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pixels) return;
  int lo = 0, hi = n_spans - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (spans[mid].offset <= i) lo = mid; else hi = mid - 1;
  }
  Span s = spans[lo];
  out[i] = make_float2((float)(s.xb + (int)(i - s.offset)), (float)s.y);
}

// kModel_inPlace (correlationKernel.cu:56-110): def positions of a list under `params`.
template <int MODEL>
__global__ void warp_list_kernel(const float2 *__restrict__ src, long n, const float *params,
                                 float cx, float cy, float2 *__restrict__ dst) {
  constexpr int NP = model_nparams(MODEL);
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float p[NP];
#pragma unroll
  for (int k = 0; k < NP; ++k) p[k] = params[k];
  float2 q = src[i];
  float xd, yd, dx, dy;
  warp_point<MODEL, DIC_MODE_PARITY>(p, q.x, q.y, cx, cy, xd, yd, dx, dy);
  dst[i] = make_float2(xd, yd);
}

// Lagrangian domain update of the CPU path: add_pair (manager_class.cpp:37-47) -- translate by the
// centre shift and round to the pixel grid.
__global__ void translate_round_kernel(float2 *xy, long n, float ox, float oy) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float2 q = xy[i];
  xy[i] = make_float2((float)(int)(__fadd_rn(__fadd_rn(ox, q.x), 0.5f)), (float)(int)(__fadd_rn(__fadd_rn(oy, q.y), 0.5f)));
}

// exact integer sums of a list of integer-valued points (DIC_CENTER_EXACT)
__global__ void sum_xy_kernel(const float2 *__restrict__ xy, long n, unsigned long long *sums) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long long sx = 0, sy = 0;
  for (; i < n; i += (long)gridDim.x * blockDim.x) {
    float2 q = xy[i];
    sx += (long long)llrintf(q.x * 1024.f);
    sy += (long long)llrintf(q.y * 1024.f);
  }
  for (int o = 16; o > 0; o >>= 1) {
    sx += __shfl_xor_sync(0xffffffffu, sx, o);
    sy += __shfl_xor_sync(0xffffffffu, sy, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&sums[0], (unsigned long long)sx);
    atomicAdd(&sums[1], (unsigned long long)sy);
  }
}

} // namespace dic
