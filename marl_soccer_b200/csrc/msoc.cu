/*
 * msoc.cu -- B200 (sm_100a) kernels and the C-ABI (include/msoc.h) of the batched 2v2 soccer
 * simulator.  Replaces the reference's per-env Python/pymunk update loop
 * (soccer_simulation/marl_vecenv.py:30-68 -> soccer_env.py:100-154 -> game/game.py:378-437 ->
 * pymunk Space.step) with one fused device step per vectorised step: two launches, see "The fused step".
 *
 * Data layout: one 128-byte record per env in each of three rotating buffers (msoc::Arrays in step_core.cuh);
 * one thread owns one env while it is stepped, one warp a batch of 32 envs (consecutive in the streaming kernel,
 * listed in the contact kernel).  The stacked observations (N,4,66) are never read back: once a warp has stored
 * the new states, the three buffers hold the poses behind the three frames of the stack, and the warp rebuilds
 * the observations of its envs from them -- four lanes per env (one per agent), eight envs at a time, staged in
 * shared memory in the layout of the output and written as whole 1 056-byte blocks by bulk asynchronous copies, so
 * that every DRAM sector is written exactly once and in full.
 *
 * There is no CPU fallback in this library: every entry point needs a CUDA device.
 */
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <string>
#include <vector>

#include "../../include/msoc.h"
#include "step_core.cuh"

using namespace msoc;

#ifdef MSOC_CHECKS
__device__ unsigned int g_msoc_check_bits;
#endif

/* ------------------------------------------------------------------------------------ errors */
static thread_local std::string g_last_error;
static std::atomic<uint64_t> g_launches{0};

static int fail(int code, const char *what, cudaError_t ce = cudaSuccess)
{
    char buf[512];
    if (ce != cudaSuccess) snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(ce));
    else snprintf(buf, sizeof buf, "%s", what);
    g_last_error = buf;
    return code;
}
#define CUDA_TRY(expr)                                                        \
    do {                                                                      \
        cudaError_t _e = (expr);                                              \
        if (_e != cudaSuccess) return fail(MSOC_ERR_CUDA, #expr, _e);         \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

/* ------------------------------------------------------------------------------------ handle */
constexpr int MAX_CHUNKS = 8; /* pipeline stages of the host-buffer step */
/* control block in device memory (ints): two sets of list counters that alternate between steps, the step counters
   that select the ping-pong halves, one set of list counters per pipeline chunk of the host-buffer step */
enum { CTL_LIGHT = 0, CTL_HEAVY = 1, CTL_NEXT_BATCH = 2, CTL_PAIR = 4, CTL_MULTI = 5, CTL_WORDS = 8 };
enum { CTL_STEP_FAST = 16, CTL_STEP_CONTACT = 17, CTL_LAST_COUNTS = 20 /* light, heavy, pair, multi of the last step */, CTL_CHUNK0 = 24, CTL_TOTAL = CTL_CHUNK0 + CTL_WORDS * MAX_CHUNKS };

struct msoc_handle {
    int device;
    int64_t n;
    uint64_t global_offset;
    SimCfg cfg;
    Arrays A;
    int sm_count, blocks_per_sm; /* persistent grid of the contact kernel */
    int heavy_lanes_override; /* MSOC_HEAVY_LANES (experiments): 0 = choose by batch size */
    void *slab;
    /* internal I/O buffers of the host-buffer API */
    float *d_obs, *d_act, *d_rew;
    uint8_t *d_done, *d_mask;
    int8_t *d_goal;
    int32_t *d_score;
    double *d_stats; /* 8 doubles, msoc_stats layout */
    int *d_ctl;      /* CTL_TOTAL ints */
    int *d_list;     /* n ints: contact lists of the step in flight (light from the front of a range, heavy from its back) */
    int *d_list2;    /* n ints: pair class from the front, multi class from the back */
    float *d_frames; /* (N,4,22) newest frames, msoc_step_host_frames; allocated on first use */
    void *d_inject;  /* Arrays::inject, allocated by the first msoc_set_state */
    cudaStream_t pipe_stream;          /* second lane of the chunked host-buffer step */
    cudaEvent_t ev_pipe_fork, ev_pipe_join;
    void *d_stage; size_t stage_bytes; /* get/set_state staging */
};

constexpr int WARPS_PER_BLOCK = 4;
constexpr int BLOCK = WARPS_PER_BLOCK * 32; /* reset kernel */

/* ------------------------------------------------------------------ observation blocks */
constexpr int ENV_STRIDE = SCRATCH_WORDS | 1; /* floats of per-lane solver scratch in the contact kernel; odd: no bank conflicts */
constexpr int OBS_ENVS = 8;                     /* envs whose observations a warp builds at a time */
constexpr int BLOCK_WORDS = OBS_ENVS * 4 * OBS;  /* floats of staged observation blocks: 8 x (4 x 66) */
constexpr int REC_F4 = 7;                        /* float4s of a record that a frame is made of */
constexpr int RECS_WORDS = OBS_ENVS * 3 * REC_F4 * 4; /* floats of staged records: 8 envs x 3 buffers x 7 float4 */
constexpr int STAGE_WORDS = BLOCK_WORDS + RECS_WORDS; /* per warp: 8 448 + 2 688 bytes, both 16-byte multiples */

/* ---------------------------------------------------------------------------- the fused step */
struct StepParams {
    Arrays A;
    SimCfg cfg;
    const float *actions; /* (N,4,3) */
    float *obs_out;       /* (N,4,66) */
    float *frames_out;    /* (N,4,22) or null */
    float *reward;        /* (N,2) */
    uint8_t *done;        /* (N) */
    int8_t *goal;         /* (N) */
    int32_t *score;       /* (N,2) or null */
    double *stats;        /* 8 */
    int *ctl;             /* the handle's control block */
    int *list;            /* N slots: envs that need the contact path -- light from the front of [e0, e1), heavy from its back */
    int *list2;           /* N slots: pair class from the front of [e0, e1), multi class from its back */
    int64_t e0, e1;       /* the envs this launch steps */
    uint64_t global_offset;
    uint32_t flags;
    int heavy_lanes;      /* envs per heavy warp-batch (MSOC_HEAVY_LANES, experiments) or 0: chosen by the contact kernel */
    int chunk;            /* -1: a whole step (list counters alternate with the step counter, the contact kernel advances it);
                             >= 0: one pipeline chunk of a host-buffer step (its own pre-zeroed counters, nobody advances) */
};

#ifdef MSOC_TIMELINE /* debug build only (tools/timeline.py): when did every batch of every step kernel run? */
__device__ unsigned long long g_tl[1 << 18]; /* records of 4 words: kernel id, start, end (globaltimer ns), sm id */
__device__ unsigned int g_tl_n;
extern "C" int msoc_debug_timeline(unsigned long long *out, unsigned int *n) {
    cudaMemcpyFromSymbol(n, g_tl_n, sizeof(unsigned int)); unsigned int z = 0; cudaMemcpyToSymbol(g_tl_n, &z, sizeof z);
    return (int)cudaMemcpyFromSymbol(out, g_tl, sizeof g_tl);
}
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void tl_record(int kid, unsigned long long t0) {
    const unsigned int i = atomicAdd(&g_tl_n, 1u);
    unsigned int smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    if (i < (1u << 16)) { g_tl[4 * i] = (unsigned long long)kid; g_tl[4 * i + 1] = t0; g_tl[4 * i + 2] = gtime(); g_tl[4 * i + 3] = smid; }
}
#define MSOC_TL_BEGIN() const unsigned long long tl0 = gtime()
#define MSOC_TL_END(kid) do { if (lane == 0) tl_record(kid, tl0); } while (0)
#else
#define MSOC_TL_BEGIN() do { } while (0)
#define MSOC_TL_END(kid) do { } while (0)
#endif
/* Per-thread tallies for the per-rollout statistics. */
struct Tally { int done, goals_b, goals_r, contacts, overflow, nonfinite; float ret; };
__device__ __forceinline__ void tally_clear(Tally &T) { T.done = T.goals_b = T.goals_r = T.contacts = T.overflow = T.nonfinite = 0; T.ret = 0.0f; }

__device__ __forceinline__ int *step_ctl(const StepParams &P, int step) { return P.chunk < 0 ? P.ctl + CTL_WORDS * (step & 1) : P.ctl + CTL_CHUNK0 + CTL_WORDS * P.chunk; }

/* Loads env e (the record of buffer step % 3, or the injected record of a flagged env) and its actions, steps it.  On
   success (always, except for the contact-free mode, which declines envs that need the contact path) the per-env
   outputs and the new state are written; the observation follows in obs_tile once the whole warp has stored. */
__device__ __forceinline__ bool step_one_env(const int MODE, const int ALLOWED, const StepParams &P, int step, int64_t e, Work &W, int &load, Tally &T)
{
    float act[12];
    const float4 *a4 = reinterpret_cast<const float4 *>(P.actions + e * 12);
    const float4 x0 = __ldg(a4), x1 = __ldg(a4 + 1), x2 = __ldg(a4 + 2);
    act[0] = x0.x; act[1] = x0.y; act[2] = x0.z; act[3] = x0.w;
    act[4] = x1.x; act[5] = x1.y; act[6] = x1.z; act[7] = x1.w;
    act[8] = x2.x; act[9] = x2.y; act[10] = x2.z; act[11] = x2.w;
    /* soccer_env.py:116-117 raises on non-finite actions; a device-resident caller gets a counter instead (the clip
       turns NaN into -1) */
    bool finite = true;
#pragma unroll
    for (int k = 0; k < 12; k++) finite = finite && (fabsf(act[k]) <= 3.0e38f);
    Env E;
    const float4 *rec = P.A.pose[buf_cur(step)] + e * POSE_F4;
    if ((ALLOWED & (1 << MODE_FULL)) && MODE == MODE_FULL) { /* the only kernel that sees injected states */
        if (__float_as_uint(rec[7].z) & FLAG_INJECT) rec = P.A.inject + e * POSE_F4;
    }
    load_env(P.A, rec, e, E);
    StepOut out;
    if (!env_step(MODE, ALLOWED, E, act, P.cfg, P.A, cache_half(step), e, P.global_offset + (uint64_t)e, P.flags, W, out, load)) return false;
    reinterpret_cast<float2 *>(P.reward)[e] = make_float2(out.reward, out.reward);
    P.done[e] = out.done;
    P.goal[e] = out.goal;
    if (P.score != nullptr) reinterpret_cast<int2 *>(P.score)[e] = make_int2(out.score_b, out.score_r);
    store_env(P.A, buf_next(step), e, E, out.score_dirty);
    if (out.fresh_episode) { /* a fresh episode has no history but its first frame (soccer_env.py:92-96) */
        store_env_record(P.A.pose[buf_cur(step)] + e * POSE_F4, E);
        store_env_record(P.A.pose[buf_prev(step)] + e * POSE_F4, E);
    }
    T.nonfinite += finite ? 0 : 1;
    T.done += out.done; T.goals_b += out.goal > 0; T.goals_r += out.goal < 0;
    T.contacts += out.n_contacts; T.overflow += out.overflow; T.ret += out.finished_return;
    return true;
}

/* ---- bulk asynchronous copies (the TMA engine, PTX cp.async.bulk): shared memory -> global memory */
__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
/* until the engine has READ the shared-memory sources of all committed groups (they may then be overwritten) */
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
/* orders this thread's shared-memory writes before later bulk copies of any thread it then synchronises with */
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

/* Stacked observations [frame(t-2) | frame(t-1) | frame(t)] (soccer_env.py:130-140) of the warp's envs whose bit is set
   in `mask` (lane l owns env `my_env`; its new state is stored), eight envs at a time:
     1. the warp fetches the 8 x 3 records (buffers (step+2)%3, step%3 -- L1/L2 hits, this warp read or prefetched
        them -- and (step+1)%3, just written by sibling lanes: read from L2) with 16-byte loads, 168 in all, the loads of
        the next eight envs in flight while the current ones are worked on, and parks them in shared memory;
     2. lane = (env, agent) rebuilds its agent's three frames and puts them where they belong in the env's 4 x 66 block;
     3. the blocks leave as bulk asynchronous copies, one per env (1 056 contiguous, 32-byte aligned bytes = 33 whole
        sectors; nothing is ever read back or written twice).
   frames_out (or null): the newest frames once more, compact (N,4,22), for the host path. */
struct RecLoad { float4 v[6]; };
/* which record word lane `lane` fetches in round r: slot r*32 + lane in [buffer][env][float4] order, 168 used */
__device__ __forceinline__ void fetch_slot(int r, int lane, int &k, int &e8, int &f)
{
    const int sl = r * 32 + lane;
    k = sl / (OBS_ENVS * REC_F4);
    const int rem = sl - k * (OBS_ENVS * REC_F4);
    e8 = rem / REC_F4; f = rem - e8 * REC_F4;
}
__device__ __forceinline__ void obs_fetch(const Arrays &A, int step, uint32_t mask, int g, int my_env, int lane, RecLoad &R)
{
    const uint32_t m8 = (mask >> (8 * g)) & 0xffu;
#pragma unroll
    for (int r = 0; r < 6; r++) {
        int k, e8, f;
        fetch_slot(r, lane, k, e8, f);
        const int env = __shfl_sync(0xffffffffu, my_env, (8 * g + e8) & 31);
        R.v[r] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if ((r < 5 || lane < 3 * OBS_ENVS * REC_F4 - 160) && ((m8 >> e8) & 1u)) {
            const int b = k == 0 ? buf_prev(step) : k == 1 ? buf_cur(step) : buf_next(step);
            const float4 *p = A.pose[b] + (int64_t)env * POSE_F4 + f;
            R.v[r] = k == 2 ? __ldcg(p) : *p;
        }
    }
}
__device__ __forceinline__ void obs_tile(const Arrays &A, const SimCfg &cfg, float *obs_out, float *frames_out, int step, float *s_stage,
                                         uint32_t lane_mask, int64_t my_env64, int lane)
{
    const int a = lane & 3, el = lane >> 2;
    float4 *recs4 = reinterpret_cast<float4 *>(s_stage + BLOCK_WORDS);
    float2 *mine = reinterpret_cast<float2 *>(s_stage) + lane * 33; /* row (el, a) of the staged blocks */
    __syncwarp(); /* the siblings' stores of the new records are ordered before the loads below */
    /* the envs of the lanes in lane_mask, compacted: the streaming kernel's warps have ~23 of 32 (the others were declined),
       which is three groups of eight instead of four */
    uint32_t mask = lane_mask;
    int my_env = (int)my_env64; /* a handle holds fewer than 2^31 envs */
    if (lane_mask != 0xffffffffu) {
        int *s_idx = reinterpret_cast<int *>(s_stage);
        if ((lane_mask >> lane) & 1u) s_idx[__popc(lane_mask & ((1u << lane) - 1u))] = my_env;
        __syncwarp();
        const int cnt = __popc(lane_mask);
        my_env = lane < cnt ? s_idx[lane] : 0;
        mask = (1u << cnt) - 1u; /* cnt < 32 here */
        __syncwarp();
    }
    int g = __ffs((int)((mask & 0xffu ? 1u : 0u) | (mask & 0xff00u ? 2u : 0u) | (mask & 0xff0000u ? 4u : 0u) | (mask & 0xff000000u ? 8u : 0u))) - 1;
    RecLoad R;
    obs_fetch(A, step, mask, g, my_env, lane, R);
#pragma unroll 1
    while (true) {
        const uint32_t m8 = (mask >> (8 * g)) & 0xffu;
        /* the staging areas are free: every lane has waited for its own bulk copies, then the warp synchronised */
#pragma unroll
        for (int r = 0; r < 6; r++)
            if (r < 5 || lane < 3 * OBS_ENVS * REC_F4 - 160) recs4[r * 32 + lane] = R.v[r];
        __syncwarp();
        /* next group with work, its loads in flight from here on */
        int gn = g + 1;
        while (gn < 4 && ((mask >> (8 * gn)) & 0xffu) == 0u) gn++;
        if (gn < 4) obs_fetch(A, step, mask, gn, my_env, lane, R);
        const int env = __shfl_sync(0xffffffffu, my_env, 8 * g + el);
        if ((m8 >> el) & 1u) {
            MSOC_CHECK(env >= 0 && env < A.n, CHK_OBS_ENV);
#pragma unroll
            for (int k = 0; k < 3; k++) {
                float o[22];
                frame_of_record(recs4 + (k * OBS_ENVS + el) * REC_F4, a, cfg, o);
#pragma unroll
                for (int i = 0; i < 11; i++) mine[k * 11 + i] = make_float2(o[2 * i], o[2 * i + 1]);
            }
        }
        fence_async_smem();
        __syncwarp();
        if (a == 0 && ((m8 >> el) & 1u)) /* one lane per env */
            bulk_store(obs_out + (int64_t)env * (4 * OBS), s_stage + el * (4 * OBS), 4 * OBS * sizeof(float));
        bulk_commit();
        if (frames_out != nullptr) {
            float2 *fr2 = reinterpret_cast<float2 *>(frames_out);
            const float2 *stage2 = reinterpret_cast<const float2 *>(s_stage);
#pragma unroll
            for (int it = 0; it < 11; it++) { /* 8 x 4 x 11 = 352 float2: the third frame of every row */
                const int idx = it * 32 + lane;
                const int row = idx / 11, j = idx - row * 11;
                const int env2 = __shfl_sync(0xffffffffu, my_env, 8 * g + (row >> 2));
                if ((m8 >> (row >> 2)) & 1u) fr2[(int64_t)env2 * 44 + (row & 3) * 11 + j] = stage2[row * 33 + 22 + j];
            }
        }
        bulk_wait_read();
        __syncwarp();
        if (gn >= 4) break;
        g = gn;
    }
}

/* The fused step is two launches on the caller's stream.
   msoc_step_fast_kernel     streams over ALL envs, thread t of block b steps env e0 + b*128 + t in contact-free
                             mode and its warp writes the observation rows.  No contact code is compiled into
                             it: few registers, small shared memory, small instruction footprint -> many
                             resident warps to hide the HBM latency.  Envs whose broad phase finds a candidate
                             pair (~27 % in the benchmark mix) write nothing and are appended, one atomic per
                             warp and class, to the list of their work class (step_core.cuh LOAD_*): light
                             (exactly one agent x wall pair, the bulk), pair (exactly one agent x agent or
                             ball x agent pair, at most one agent x wall pair beside it), multi (only wall candidates, several), heavy (anything else).
   msoc_step_contact_kernel  every warp of a persistent grid pulls batches of listed envs of ONE class -- heavy
                             batches first, then multi, pair, light: longest first -- and every thread steps one
                             env in the mode of its class: the register-only single-body solver (light), islands
                             of one or two bodies with their contacts in the lane's shared-memory slots (pair,
                             multi), or the general Chipmunk path (heavy): narrow phase over all candidate pairs,
                             arbiter cache, 10-iteration impulse solver with bodies and contacts in shared
                             memory.  The warp then writes the rows.
   The divergent, latency-bound contact work therefore always runs on warps of similar work, and the few long
   heavy batches start at once on their own warps while all other warps stream through the rest.  (Separate kernels
   per class on separate streams were measured to run one after the other, not side by side: profiles/.)

   Which half of the ping-pong state is current is a step counter in DEVICE memory (so a captured CUDA graph of any
   number of steps replays correctly): the fast kernel reads ctl[CTL_STEP_FAST] and copies it to ctl[CTL_STEP_CONTACT]
   for the contact kernel of the same step; the contact kernel, which only starts when the fast kernel is complete
   and is complete before the next fast kernel starts, writes the incremented counter back. */
#ifndef MSOC_FAST_BLOCK
#define MSOC_FAST_BLOCK 128
#endif
constexpr int FAST_BLOCK = MSOC_FAST_BLOCK; /* envs (= threads) per block of the fast kernel */
constexpr size_t FAST_SMEM_BYTES = (size_t)(FAST_BLOCK / 32) * STAGE_WORDS * sizeof(float);
#ifndef MSOC_FAST_MIN_BLOCKS
#define MSOC_FAST_MIN_BLOCKS 4
#endif

/* warp reduce of the per-rollout statistics (marl-soccer.ipynb:411-429), one atomic per warp and counter */
__device__ __forceinline__ void flush_tally(const Tally &T, double *stats, int lane)
{
    int nd = T.done, gb = T.goals_b, gr = T.goals_r, nc = T.contacts, ov = T.overflow, nf = T.nonfinite;
    float ret = T.ret;
    if (!__any_sync(0xffffffffu, (nd | gb | gr | nc | ov | nf) != 0)) return; /* the common case of the streaming kernel */
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        nd += __shfl_xor_sync(0xffffffffu, nd, o); gb += __shfl_xor_sync(0xffffffffu, gb, o);
        gr += __shfl_xor_sync(0xffffffffu, gr, o); nc += __shfl_xor_sync(0xffffffffu, nc, o);
        ov += __shfl_xor_sync(0xffffffffu, ov, o);
        nf += __shfl_xor_sync(0xffffffffu, nf, o);
        ret += __shfl_xor_sync(0xffffffffu, ret, o);
    }
    if (lane == 0) {
        if (nd) { atomicAdd(stats + 0, (double)nd); atomicAdd(stats + 1, (double)ret); }
        if (gb) atomicAdd(stats + 2, (double)gb);
        if (gr) atomicAdd(stats + 3, (double)gr);
        if (nc) atomicAdd(stats + 5, (double)nc);
        if (ov) atomicAdd(stats + 6, (double)ov);
        if (nf) atomicAdd(stats + 7, (double)nf);
    }
}

__global__ void __launch_bounds__(FAST_BLOCK, MSOC_FAST_MIN_BLOCKS) msoc_step_fast_kernel(const __grid_constant__ StepParams P)
{
    extern __shared__ __align__(16) float s_dyn[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float *s_warp = s_dyn + warp * STAGE_WORDS;
    cudaGridDependencySynchronize(); /* launched programmatically behind the previous kernel of the stream */
    const int step = P.ctl[CTL_STEP_FAST];
    int *ctl = step_ctl(P, step);
    if (blockIdx.x == 0) {
        if (P.chunk < 0 && tid < CTL_WORDS) P.ctl[CTL_WORDS * ((step & 1) ^ 1) + tid] = 0; /* the next step's list counters */
        if (tid == 0) {
            P.ctl[CTL_STEP_CONTACT] = step;
            atomicAdd(P.stats + 4, (double)(P.e1 - P.e0)); /* every env of the range is stepped by one of the two kernels */
        }
    }
    MSOC_TL_BEGIN();
    const int64_t my_env = P.e0 + (int64_t)blockIdx.x * FAST_BLOCK + tid;
    const bool have = my_env < P.e1;
    Tally T; tally_clear(T);
    Work W; /* never touched in contact-free mode */
    W.ovf = nullptr; W.body = W.pool = W.geom = W.old = W.isl = nullptr; W.pool_count = nullptr;
    bool ok = false;
    int load = 0;
    if (have) {
        /* the oldest of the three records the observation is rebuilt from is not needed by the step: start pulling it */
        asm volatile("prefetch.global.L2 [%0];" ::"l"(P.A.pose[buf_prev(step)] + my_env * POSE_F4));
        ok = step_one_env(MODE_FAST, 1 << MODE_FAST, P, step, my_env, W, load, T);
    }
    const uint32_t mask = __ballot_sync(0xffffffffu, ok);
    /* list slots of the declined envs, by work class: lanes 0..3 each fetch the base of one class (one atomic per warp and
       class that occurs); the atomics are in flight while the observations are built */
    const bool declined = have && !ok;
    const uint32_t mdec = __ballot_sync(0xffffffffu, declined);
    uint32_t mclass = 0u; /* the warp's declined lanes of my class */
    int base = 0;
    if (mdec != 0u) {
        const uint32_t b0 = __ballot_sync(0xffffffffu, declined && (load & 1)), b1 = __ballot_sync(0xffffffffu, declined && (load & 2));
        auto class_mask = [&](int cl) { return mdec & ((cl & 1) ? b0 : ~b0) & ((cl & 2) ? b1 : ~b1); };
        mclass = class_mask(load);
        const uint32_t of_lane = class_mask(lane & 3); /* lane c < 4 speaks for class c */
        if (lane < N_LOADS && of_lane != 0u) {
            const int word = lane == LOAD_LIGHT ? CTL_LIGHT : lane == LOAD_HEAVY ? CTL_HEAVY : lane == LOAD_PAIR ? CTL_PAIR : CTL_MULTI;
            base = atomicAdd(ctl + word, __popc(of_lane));
        }
    }
    if (mask != 0u) obs_tile(P.A, P.cfg, P.obs_out, P.frames_out, step, s_warp, mask, my_env, lane);
    if (mdec != 0u) {
        base = __shfl_sync(0xffffffffu, base, declined ? load : 0);
        if (declined) {
            const int rank = base + __popc(mclass & ((1u << lane) - 1u));
            int *lst = (load == LOAD_LIGHT || load == LOAD_HEAVY) ? P.list : P.list2;
            const bool front = load == LOAD_LIGHT || load == LOAD_PAIR;
            lst[front ? P.e0 + rank : P.e1 - 1 - rank] = (int)my_env;
        }
    }
    flush_tally(T, P.stats, lane);
    if ((blockIdx.x & 15) == 0 && warp == 0) MSOC_TL_END(0);
}

/* Per-warp scratch of 32 x ENV_STRIDE floats: during the contact solve it holds the lanes' solver bodies (30 fields,
   field-major with stride 32: conflict-free), the warp's pool of 32 x CON_FAST contact records (15 fields, field-major;
   an env takes as many records as it has contacts; in the pair / multi modes the same words are the lanes' own island
   slots), the parked poses and the preloaded arbiter cache entries; afterwards the staged observation blocks. */
static_assert(STAGE_WORDS <= 32 * ENV_STRIDE, "the observation staging must fit the per-warp solver scratch");
#ifndef MSOC_HEAVY_BLOCK
#define MSOC_HEAVY_BLOCK 64 /* threads per block of the contact kernel */
#endif
constexpr int HEAVY_BLOCK = MSOC_HEAVY_BLOCK;
#ifndef MSOC_HEAVY_MIN_BLOCKS
#define MSOC_HEAVY_MIN_BLOCKS 4 /* register cap 255: no spills with all four modes compiled in; at 168 registers (5 or 6 blocks) the spills cost more than the extra warps bring */
#endif
constexpr int HEAVY_MIN_BLOCKS = MSOC_HEAVY_MIN_BLOCKS;

constexpr int HEAVY_WARP_WORDS = (32 * ENV_STRIDE + 3) & ~3; /* 16-byte aligned per-warp scratch */
constexpr size_t STEP_SMEM_BYTES = (size_t)(HEAVY_BLOCK / 32) * HEAVY_WARP_WORDS * sizeof(float);

/* All four contact classes in ONE persistent kernel (concurrent kernels with different register / shared-memory shapes
   were measured to run one after the other on this part, not side by side): every warp pulls batches from one counter
   over the combined batch space [heavy | multi | pair | light], i.e. longest batches first, so that the few long
   latency-bound heavy batches start at once on their own warps while all the other warps stream through the rest. */
template <int MODES>
__device__ __forceinline__ void contact_body(const StepParams &P, int *s_pool_count)
{
    extern __shared__ __align__(16) float s_dyn[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float *s_warp = s_dyn + warp * HEAVY_WARP_WORDS;
    cudaGridDependencySynchronize(); /* launched programmatically behind the streaming kernel: wait until it is complete */
    const int step = P.ctl[CTL_STEP_CONTACT];
    int *ctl = step_ctl(P, step);
    /* this step's fast kernel is complete and the next one starts after this kernel: advance the step counter */
    if (P.chunk < 0 && blockIdx.x == 0 && tid == 0) {
        P.ctl[CTL_STEP_FAST] = (step + 1) % 6;
        P.ctl[CTL_LAST_COUNTS + 0] = ctl[CTL_LIGHT]; P.ctl[CTL_LAST_COUNTS + 1] = ctl[CTL_HEAVY];
        P.ctl[CTL_LAST_COUNTS + 2] = ctl[CTL_PAIR]; P.ctl[CTL_LAST_COUNTS + 3] = ctl[CTL_MULTI];
    }
    /* final: the fast kernel has finished */
    const int n_heavy = (MODES & (1 << MODE_FULL)) ? ctl[CTL_HEAVY] : 0, n_multi = (MODES & (1 << MODE_MULTI)) ? ctl[CTL_MULTI] : 0;
    const int n_pair = (MODES & (1 << MODE_PAIR)) ? ctl[CTL_PAIR] : 0, n_light = (MODES & (1 << MODE_LIGHT)) ? ctl[CTL_LIGHT] : 0;
    MSOC_CHECK(step >= 0 && step < 6, CHK_STEP_COUNTER);
    MSOC_CHECK(n_heavy >= 0 && n_light >= 0 && (int64_t)n_heavy + n_light <= P.e1 - P.e0, CHK_LIST_COUNT);
    MSOC_CHECK(n_multi >= 0 && n_pair >= 0 && (int64_t)n_multi + n_pair <= P.e1 - P.e0, CHK_LIST_COUNT);
    /* Envs per heavy batch.  A heavy batch is bound by its latency (the warp walks the union of its lanes' divergent
       contact work: 24 us median / 65 us worst for one env, ~110 us median for 32): as few envs per warp as still gives every heavy batch a warp of
       its own at once; wider batches once there are several times more heavy envs than warps (throughput). */
    int heavy_lanes = P.heavy_lanes;
    if (heavy_lanes == 0) {
        const int n_warps_ = gridDim.x * (HEAVY_BLOCK / 32);
        heavy_lanes = 1;
        while (heavy_lanes < 32 && heavy_lanes * n_warps_ < n_heavy) heavy_lanes *= 2;
        if (heavy_lanes >= 8) heavy_lanes = heavy_lanes >= 16 ? 32 : 16; /* several rounds of batches per warp anyway: throughput
                                                                             (measured: 512 Ki envs 0.282 ms with 16, 0.295 with 8) */
    }
    const int b_heavy = (n_heavy + heavy_lanes - 1) / heavy_lanes;
    const int b_multi = b_heavy + (n_multi + 31) / 32, b_pair = b_multi + (n_pair + 31) / 32;
    const int batches = b_pair + (n_light + 31) / 32;

    Tally T; tally_clear(T);
    float ovf_store[MAXC - CON_FAST][CON_FIELDS]; /* local memory, touched only by envs with more than CON_FAST contacts */
    Work W;
    W.ovf = ovf_store;
    W.body = s_warp + lane;
    W.pool = s_warp + BODY_FIELDS * 5 * 32; /* shared by the warp's lanes */
    W.pool_count = s_pool_count + warp;
    W.geom = s_warp + (BODY_FIELDS * 5 + CON_FIELDS * CON_FAST) * 32 + lane;
    W.old = s_warp + (BODY_FIELDS * 5 + CON_FIELDS * CON_FAST + GEOM_WORDS) * 32 + lane;
    W.isl = s_warp + BODY_FIELDS * 5 * 32 + lane; /* island modes: their contact slots, in the (then idle) contact pool */
    /* A warp's first batch is its own number (so the heavy batches, numbered first, each start at once on a warp of their
       own); the following ones are handed out by the batch counter.  (Claiming batches ahead of time and pulling their
       state into the L2 while the current one is stepped was measured slower: batches waiting behind a long one cost
       more than the latency that is hidden.) */
    const int n_warps = gridDim.x * (HEAVY_BLOCK / 32);
    int b = blockIdx.x * (HEAVY_BLOCK / 32) + warp;
#pragma unroll 1
    while (b < batches) {
        int mode, idx, count, lanes = 32;
        const int *slot;
        if (b < b_heavy)      { mode = MODE_FULL;  lanes = heavy_lanes; idx = b * lanes + lane;  count = n_heavy; slot = P.list + (P.e1 - 1 - idx); }
        else if (b < b_multi) { mode = MODE_MULTI; idx = (b - b_heavy) * 32 + lane; count = n_multi; slot = P.list2 + (P.e1 - 1 - idx); }
        else if (b < b_pair)  { mode = MODE_PAIR;  idx = (b - b_multi) * 32 + lane; count = n_pair;  slot = P.list2 + (P.e0 + idx); }
        else                  { mode = MODE_LIGHT; idx = (b - b_pair) * 32 + lane;  count = n_light; slot = P.list + (P.e0 + idx); }
        const bool have = lane < lanes && idx < count;
        const int64_t my_env = have ? (int64_t)*slot : 0;
        MSOC_CHECK(!have || (my_env >= P.e0 && my_env < P.e1), CHK_LIST_ENV);
        MSOC_TL_BEGIN();
        bool ok = false;
        int load = 0;
        if (lane == 0) *W.pool_count = 0; /* the warp's contact pool is empty */
        __syncwarp();
        if (have) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(P.A.pose[buf_prev(step)] + my_env * POSE_F4));
            ok = step_one_env(mode, MODES, P, step, my_env, W, load, T);
        }
        const uint32_t mask = __ballot_sync(0xffffffffu, ok);
        /* (obs_tile starts with a __syncwarp: the solver scratch of every lane is dead before it is reused) */
        if (mask != 0u) obs_tile(P.A, P.cfg, P.obs_out, P.frames_out, step, s_warp, mask, my_env, lane);
        MSOC_TL_END(mode == MODE_FULL ? 2 : mode == MODE_LIGHT ? 1 : 3);
        if (lane == 0) b = n_warps + atomicAdd(ctl + CTL_NEXT_BATCH, 1);
        b = __shfl_sync(0xffffffffu, b, 0);
    }
    flush_tally(T, P.stats, lane);
}

__global__ void __launch_bounds__(HEAVY_BLOCK, HEAVY_MIN_BLOCKS) msoc_step_contact_kernel(const __grid_constant__ StepParams P)
{
    __shared__ int s_pool_count[HEAVY_BLOCK / 32];
    contact_body<(1 << MODE_FULL) | (1 << MODE_LIGHT) | (1 << MODE_PAIR) | (1 << MODE_MULTI)>(P, s_pool_count);
}

/* after a chunked host-buffer step (whose contact kernels do not): advances the step counter, sums the chunks' class counts */
__global__ void msoc_advance_kernel(int *ctl)
{
    if (threadIdx.x == 0) ctl[CTL_STEP_FAST] = (ctl[CTL_STEP_FAST] + 1) % 6;
    if (threadIdx.x < 4) {
        const int word = threadIdx.x == 0 ? CTL_LIGHT : threadIdx.x == 1 ? CTL_HEAVY : threadIdx.x == 2 ? CTL_PAIR : CTL_MULTI;
        int sum = 0;
        for (int c = 0; c < MAX_CHUNKS; c++) sum += ctl[CTL_CHUNK0 + CTL_WORDS * c + word];
        ctl[CTL_LAST_COUNTS + threadIdx.x] = sum;
    }
}

/* ------------------------------------------------------------------------------------- reset */
struct ResetParams {
    Arrays A;
    SimCfg cfg;
    const uint8_t *mask; /* N or null */
    float *obs_out;      /* (N,4,66) or null */
    const int *ctl;
    uint64_t global_offset, seed;
    int mode, has_seed;
};

__global__ void __launch_bounds__(BLOCK) msoc_reset_kernel(const __grid_constant__ ResetParams P)
{
    __shared__ __align__(16) float s_stage[WARPS_PER_BLOCK][STAGE_WORDS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t block_base = (int64_t)blockIdx.x * BLOCK;
    const int64_t e = block_base + threadIdx.x;
    if (block_base + warp * 32 >= P.A.n) return;
    const int step = P.ctl[CTL_STEP_FAST];
    const bool doit = (e < P.A.n) && (P.mask == nullptr || P.mask[e] != 0);
    if (doit) {
        const uint64_t gidx = P.global_offset + (uint64_t)e;
        uint64_t seed; uint32_t sc;
        if (P.has_seed) { seed = P.seed + gidx; sc = 0; P.A.seed[e] = seed; } /* marl_vecenv.py:23: seed + i */
        else { seed = P.A.seed[e]; sc = P.A.spawn_count[e]; }
        Env E;
        env_full_reset(E, P.mode, seed, gidx, sc);
        P.A.spawn_count[e] = sc;
        /* all three buffers: the observation history of a fresh episode is its first frame (soccer_env.py:92-96).
           The state the next step starts from is buffer step % 3. */
        store_env(P.A, buf_cur(step), e, E, true);
        store_env_record(P.A.pose[buf_prev(step)] + e * POSE_F4, E);
        store_env_record(P.A.pose[buf_next(step)] + e * POSE_F4, E);
    }
    const uint32_t m = __ballot_sync(0xffffffffu, doit);
    if (P.obs_out != nullptr && m != 0u) obs_tile(P.A, P.cfg, P.obs_out, nullptr, step, s_stage[warp], m, e, lane);
}

/* -------------------------------------------------------------------- state inject / extract */
__device__ __forceinline__ void hist_to_record(const msoc_env_state &S, int k, float4 *r)
{
    Pose Q;
    for (int i = 0; i < 5; i++) { Q.px[i] = S.hist_pos[k][i][0]; Q.py[i] = S.hist_pos[k][i][1]; }
    for (int i = 0; i < 4; i++) { Q.vx[i] = S.hist_vel[k][i][0]; Q.vy[i] = S.hist_vel[k][i][1]; Q.ang[i] = S.hist_ang[k][i]; Q.w[i] = S.hist_angvel[k][i]; }
    pose_pack(Q, 0.0f, 0.0f, r);
}
__device__ __forceinline__ void record_to_hist(const float4 *r, msoc_env_state &S, int k)
{
    Pose Q; pose_unpack(r, Q);
    for (int i = 0; i < 5; i++) { S.hist_pos[k][i][0] = Q.px[i]; S.hist_pos[k][i][1] = Q.py[i]; }
    for (int i = 0; i < 4; i++) { S.hist_vel[k][i][0] = Q.vx[i]; S.hist_vel[k][i][1] = Q.vy[i]; S.hist_ang[k][i] = Q.ang[i]; S.hist_angvel[k][i] = Q.w[i]; }
}
/* same pose?  (bit patterns of the 26 floats a frame is made of; the ball velocity in r[4].zw is not part of it) */
__device__ __forceinline__ bool same_pose(const float4 *a, const float4 *b)
{
    bool same = true;
    for (int i = 0; i < 7; i++) {
        same = same && __float_as_uint(a[i].x) == __float_as_uint(b[i].x) && __float_as_uint(a[i].y) == __float_as_uint(b[i].y);
        if (i != 4) same = same && __float_as_uint(a[i].z) == __float_as_uint(b[i].z) && __float_as_uint(a[i].w) == __float_as_uint(b[i].w);
    }
    return same;
}

__global__ void msoc_get_state_kernel(Arrays A, const int *ctl, const int64_t *idx, int64_t n, msoc_env_state *out)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int step = ctl[CTL_STEP_FAST];
    const int64_t e = idx[t];
    const float4 *rec = A.pose[buf_cur(step)] + e * POSE_F4;
    const bool injected = (__float_as_uint(rec[7].z) & FLAG_INJECT) != 0u && A.inject != nullptr;
    Env E;
    load_env(A, injected ? A.inject + e * POSE_F4 : rec, e, E);
    msoc_env_state S;
    memset(&S, 0, sizeof S);
    for (int i = 0; i < 5; i++) {
        S.pos[i][0] = E.px[i]; S.pos[i][1] = E.py[i]; S.vel[i][0] = E.vx[i]; S.vel[i][1] = E.vy[i];
        S.angvel[i] = E.w[i]; S.vbias[i][0] = E.vbx[i]; S.vbias[i][1] = E.vby[i];
    }
    for (int i = 0; i < 4; i++) { S.ang[i] = E.ang[i]; S.wbias[i] = E.wb[i]; }
    S.ep_return = E.ep_return; S.steps = E.steps; S.score[0] = E.score_b; S.score[1] = E.score_r;
    S.mode = (int)((E.flags & FLAG_MODE_MASK) >> FLAG_MODE_SHIFT);
    S.spawn_count = A.spawn_count[e]; S.seed = A.seed[e];
    const uint32_t cnt = E.flags & FLAG_CACHE_MASK;
    S.cache_count = cnt;
    for (uint32_t j = 0; j < cnt; j++) {
        const uint32_t *c = A.cache[cache_half(step)] + cache_slot(e, (int)j);
        S.cache_info[j] = c[0]; S.cache_jn[j] = __uint_as_float(c[1]); S.cache_jt[j] = __uint_as_float(c[2]);
    }
    S.hist_valid = 1u;
    record_to_hist(A.pose[buf_prev(step)] + e * POSE_F4, S, 0);
    record_to_hist(rec, S, 1);
    out[t] = S;
}

__global__ void msoc_set_state_kernel(Arrays A, const int *ctl, const int64_t *idx, int64_t n, const msoc_env_state *in)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int step = ctl[CTL_STEP_FAST];
    const int64_t e = idx[t];
    const msoc_env_state &S = in[t];
    Env E;
    for (int i = 0; i < 5; i++) {
        E.px[i] = S.pos[i][0]; E.py[i] = S.pos[i][1]; E.vx[i] = S.vel[i][0]; E.vy[i] = S.vel[i][1];
        E.w[i] = S.angvel[i]; E.vbx[i] = S.vbias[i][0]; E.vby[i] = S.vbias[i][1];
    }
    for (int i = 0; i < 4; i++) { E.ang[i] = S.ang[i]; E.wb[i] = S.wbias[i]; }
    E.ep_return = S.ep_return; E.steps = S.steps; E.score_b = S.score[0]; E.score_r = S.score[1];
    uint32_t cnt = S.cache_count > (uint32_t)MAX_CACHE ? (uint32_t)MAX_CACHE : S.cache_count;
    E.flags = cnt | (((uint32_t)S.mode & 3u) << FLAG_MODE_SHIFT);
    A.spawn_count[e] = S.spawn_count; A.seed[e] = S.seed;
    for (uint32_t j = 0; j < cnt; j++) {
        uint32_t *c = A.cache[cache_half(step)] + cache_slot(e, (int)j);
        c[0] = S.cache_info[j]; c[1] = __float_as_uint(S.cache_jn[j]); c[2] = __float_as_uint(S.cache_jt[j]);
    }
    /* The observation history.  The frames already emitted stay what they were (the reference's deque is not touched by
       a poke of the bodies): the current record keeps the pose behind the newest of them and the injected state goes to
       Arrays::inject -- unless the two poses are the same bits.  With hist_valid the caller supplies both history poses
       (checkpoint restore, parity injection). */
    float4 *rec = A.pose[buf_cur(step)] + e * POSE_F4;
    float4 newp[7], prev[7];
    { Pose Q; pose_of(E, Q); pose_pack(Q, 0.0f, 0.0f, newp); }
    if (S.hist_valid) {
        float4 h0[7];
        hist_to_record(S, 0, h0);
        for (int i = 0; i < 7; i++) A.pose[buf_prev(step)][e * POSE_F4 + i] = h0[i];
        hist_to_record(S, 1, prev);
    } else {
        for (int i = 0; i < 7; i++) prev[i] = rec[i];
    }
    store_env_bias(A, e, E);
    A.score[e] = make_int2(E.score_b, E.score_r);
    if (same_pose(prev, newp)) {
        store_env_record(rec, E);
    } else {
        store_env_record(A.inject + e * POSE_F4, E);
        for (int i = 0; i < 7; i++) rec[i] = prev[i];
        rec[7] = make_float4(0.0f, __uint_as_float((uint32_t)E.steps), __uint_as_float(FLAG_INJECT), 0.0f);
    }
}

/* rows of the internal observation buffer of a list of envs, packed */
__global__ void msoc_gather_obs_kernel(const float *obs, const int64_t *idx, int64_t n, float *out)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * 4 * OBS) return;
    const int64_t k = t / (4 * OBS);
    out[t] = obs[idx[k] * 4 * OBS + (t - k * 4 * OBS)];
}

__global__ void msoc_init_kernel(Arrays A, uint64_t seed)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= A.n) return;
    A.seed[e] = seed;
    A.spawn_count[e] = 0;
}

/* ------------------------------------------------------------------ rollout: policy inputs */
/* One pass over the blue agents' observation rows (528 contiguous bytes per env): normalised + clipped bf16 policy input
   (padded rows), raw bf16 copy for the rollout buffer, per-feature sum / sum of squares for the running normaliser.
   Thread (x, y): float2 x of the 132 floats of env y (+ 4 per iteration), so a thread always sees the same two features. */
constexpr int PI_X = 66, PI_Y = 4, PI_PAD = 72;
__global__ void __launch_bounds__(PI_X * PI_Y) msoc_policy_inputs_kernel(const float *__restrict__ obs, int64_t n, const float *__restrict__ shift,
                                                                         const float *__restrict__ inv_std, __nv_bfloat162 *__restrict__ x_out,
                                                                         __nv_bfloat162 *__restrict__ raw_out, double *moments)
{
    __shared__ float s_red[PI_Y][PI_X][4];
    const int x = threadIdx.x, y = threadIdx.y;
    const int row = x >= 33 ? 1 : 0, f = 2 * x - 66 * row; /* features f, f + 1 of agent `row` */
    const float sh0 = shift[f], sh1 = shift[f + 1], is0 = inv_std[f], is1 = inv_std[f + 1];
    float s0 = 0.0f, s1 = 0.0f, q0 = 0.0f, q1 = 0.0f;
    for (int64_t e = (int64_t)blockIdx.x * PI_Y + y; e < n; e += (int64_t)gridDim.x * PI_Y) {
        const float2 v = reinterpret_cast<const float2 *>(obs + e * (4 * OBS))[x];
        const float a = fminf(fmaxf(fmaf(v.x, is0, sh0), -10.0f), 10.0f), b = fminf(fmaxf(fmaf(v.y, is1, sh1), -10.0f), 10.0f);
        x_out[((2 * e + row) * PI_PAD + f) >> 1] = __floats2bfloat162_rn(a, b);
        if (raw_out != nullptr) raw_out[(e * (2 * OBS) + 2 * x) >> 1] = __floats2bfloat162_rn(v.x, v.y);
        s0 += v.x; s1 += v.y; q0 = fmaf(v.x, v.x, q0); q1 = fmaf(v.y, v.y, q1);
    }
    if (moments == nullptr) return;
    s_red[y][x][0] = s0; s_red[y][x][1] = s1; s_red[y][x][2] = q0; s_red[y][x][3] = q1;
    __syncthreads();
    if (y == 0) {
        double t0 = 0.0, t1 = 0.0, u0 = 0.0, u1 = 0.0;
        for (int k = 0; k < PI_Y; k++) { t0 += s_red[k][x][0]; t1 += s_red[k][x][1]; u0 += s_red[k][x][2]; u1 += s_red[k][x][3]; }
        atomicAdd(moments + f, t0); atomicAdd(moments + f + 1, t1);
        atomicAdd(moments + OBS + f, u0); atomicAdd(moments + OBS + f + 1, u1);
    }
}

/* ---------------------------------------------------------------------------------- C-ABI */
extern "C" {

const char *msoc_last_error(void) { return g_last_error.c_str(); }
int msoc_version(void) { return MSOC_VERSION; }
uint64_t msoc_launch_count(void) { return g_launches.load(); }
/* Bits of the failed kernel checks since the last call (see MSOC_CHECK in step_core.cuh); -1 in a build without
   -DMSOC_CHECKS.  Synchronises the device. */
int msoc_debug_errors(void)
{
#ifdef MSOC_CHECKS
    unsigned int bits = 0, zero = 0;
    if (cudaDeviceSynchronize() != cudaSuccess) return -2;
    if (cudaMemcpyFromSymbol(&bits, g_msoc_check_bits, sizeof bits) != cudaSuccess) return -2;
    cudaMemcpyToSymbol(g_msoc_check_bits, &zero, sizeof zero);
    return (int)bits;
#else
    return -1;
#endif
}
int64_t msoc_num_envs(const msoc_handle *h) { return h ? h->n : 0; }

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

static int grid_for(int64_t n) { return (int)((n + BLOCK - 1) / BLOCK); }

static void fill_cfg(const msoc_config *c, SimCfg &s)
{
    s.max_velocity = c->max_velocity;
    s.agent_minv = 1.0f / c->agent_mass; s.ball_minv = 1.0f / c->ball_mass;
    s.agent_iinv = 1.0f / c->agent_moment; s.ball_iinv = 1.0f / c->ball_moment;
    s.agent_friction = c->agent_friction; s.ball_friction = c->ball_friction;
    s.force_max = c->action_force_max; s.torque_max = c->action_torque_max;
    s.max_ang_vel = c->max_angular_velocity;
    s.prox_mult = c->ball_proximity_multiplier; s.move_mult = c->move_ball_to_goal_multiplier;
    s.goal_reward = c->goal_scored_reward; s.conceded_penalty = c->goal_conceded_penalty;
    s.alive_penalty = c->alive_penalty; s.score_diff_mult = c->score_difference_multiplier;
    s.max_steps = c->max_steps; s.pad = 0;
    cfg_derive(s);
}

int msoc_reset(msoc_handle *h, const uint8_t *d_mask, int mode, int has_seed, uint64_t seed, float *d_obs_out, void *stream);
int msoc_destroy(msoc_handle *h);

int msoc_create(const msoc_config *cfg, int64_t n_envs, int device, uint64_t seed, uint64_t global_env_offset,
                msoc_handle **out)
{
    if (!cfg || !out || n_envs <= 0) return fail(MSOC_ERR_INVALID, "msoc_create: bad argument");
    if (n_envs > 0x7fffffff) return fail(MSOC_ERR_INVALID, "msoc_create: at most 2^31 - 1 envs per handle");
    if (!(cfg->agent_mass > 0.0f) || !(cfg->ball_mass > 0.0f) || !(cfg->agent_moment > 0.0f) || !(cfg->ball_moment > 0.0f))
        return fail(MSOC_ERR_INVALID, "msoc_create: masses and moments must be positive");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev <= 0)
        return fail(MSOC_ERR_CUDA, "msoc_create: no CUDA device (this library has no CPU fallback)", ce);
    if (device < 0 || device >= ndev) return fail(MSOC_ERR_INVALID, "msoc_create: bad device index");
    DeviceGuard guard(device);

    msoc_handle *h = new msoc_handle();
    memset(h, 0, sizeof *h);
    h->device = device; h->n = n_envs; h->global_offset = global_env_offset;
    fill_cfg(cfg, h->cfg);
    if (const char *hl = getenv("MSOC_HEAVY_LANES")) {
        const int v = atoi(hl);
        if (v == 1 || v == 2 || v == 4 || v == 8 || v == 16 || v == 32) h->heavy_lanes_override = v;
    }
    /* every failure below goes through msoc_destroy, which releases whatever exists so far */
    auto bail = [&](int code, const char *what, cudaError_t e) { msoc_destroy(h); return fail(code, what, e); };

    const size_t n = (size_t)n_envs;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes); return o; };
    size_t o_pose[3], o_cache[2];
    for (int k = 0; k < 3; k++) o_pose[k] = take(n * POSE_F4 * sizeof(float4));
    const size_t o_iscore = take(n * sizeof(int2));
    const size_t o_bias = take(n * 4 * sizeof(float4));
    const size_t o_seed = take(n * sizeof(uint64_t)), o_sc = take(n * sizeof(uint32_t));
    for (int k = 0; k < 2; k++) o_cache[k] = take(n * MAX_CACHE * 3 * sizeof(uint32_t));
    const size_t o_obs = take(n * 4 * OBS * sizeof(float)), o_act = take(n * 12 * sizeof(float));
    const size_t o_rew = take(n * 2 * sizeof(float)), o_done = take(n), o_goal = take(n), o_mask = take(n);
    const size_t o_score = take(n * 2 * sizeof(int32_t));
    const size_t o_stats = take(8 * sizeof(double));
    const size_t o_ctl = take(CTL_TOTAL * sizeof(int));
    const size_t o_list = take(n * sizeof(int)), o_list2 = take(n * sizeof(int));
    const size_t total = off;

    ce = cudaMalloc(&h->slab, total);
    if (ce != cudaSuccess) return bail(MSOC_ERR_ALLOC, "msoc_create: cudaMalloc", ce);
    ce = cudaMemset(h->slab, 0, total);
    if (ce != cudaSuccess) return bail(MSOC_ERR_CUDA, "msoc_create: cudaMemset", ce);
    char *base = (char *)h->slab;
    Arrays &A = h->A;
    A.n = n_envs;
    for (int k = 0; k < 3; k++) A.pose[k] = (float4 *)(base + o_pose[k]);
    for (int k = 0; k < 2; k++) A.cache[k] = (uint32_t *)(base + o_cache[k]);
    A.score = (int2 *)(base + o_iscore);
    A.bias = (float4 *)(base + o_bias);
    A.inject = nullptr;
    A.seed = (uint64_t *)(base + o_seed); A.spawn_count = (uint32_t *)(base + o_sc);
    h->d_obs = (float *)(base + o_obs); h->d_act = (float *)(base + o_act); h->d_rew = (float *)(base + o_rew);
    h->d_done = (uint8_t *)(base + o_done); h->d_goal = (int8_t *)(base + o_goal); h->d_mask = (uint8_t *)(base + o_mask);
    h->d_score = (int32_t *)(base + o_score);
    h->d_stats = (double *)(base + o_stats);
    h->d_ctl = (int *)(base + o_ctl);
    h->d_list = (int *)(base + o_list); h->d_list2 = (int *)(base + o_list2);

    ce = cudaFuncSetAttribute(msoc_step_contact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)STEP_SMEM_BYTES);
    if (ce == cudaSuccess)
        ce = cudaFuncSetAttribute(msoc_step_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FAST_SMEM_BYTES);
    if (ce != cudaSuccess) return bail(MSOC_ERR_CUDA, "msoc_create: smem attribute", ce);
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
    ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->blocks_per_sm, msoc_step_contact_kernel, HEAVY_BLOCK, STEP_SMEM_BYTES);
    if (ce != cudaSuccess || h->blocks_per_sm < 1 || h->sm_count < 1) return bail(MSOC_ERR_CUDA, "msoc_create: occupancy query", ce);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&h->pipe_stream, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&h->ev_pipe_fork, cudaEventDisableTiming);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&h->ev_pipe_join, cudaEventDisableTiming);
    if (ce != cudaSuccess) return bail(MSOC_ERR_CUDA, "msoc_create: stream/event", ce);
    msoc_init_kernel<<<(unsigned)((n_envs + 255) / 256), 256>>>(A, seed);
    g_launches++;
    /* Game.__init__ -> setup_field -> reset(): first spawn in the default random mode */
    int rc = msoc_reset(h, nullptr, MSOC_MODE_RANDOM, 0, 0, h->d_obs, nullptr);
    if (rc != MSOC_OK) { msoc_destroy(h); return rc; }
    ce = cudaDeviceSynchronize();
    if (ce != cudaSuccess) return bail(MSOC_ERR_CUDA, "msoc_create: init", ce);
    *out = h;
    return MSOC_OK;
}

int msoc_destroy(msoc_handle *h)
{
    if (!h) return MSOC_OK;
    DeviceGuard guard(h->device);
    cudaDeviceSynchronize();
    if (h->d_stage) cudaFree(h->d_stage);
    if (h->d_frames) cudaFree(h->d_frames);
    if (h->d_inject) cudaFree(h->d_inject);
    if (h->ev_pipe_fork) cudaEventDestroy(h->ev_pipe_fork);
    if (h->ev_pipe_join) cudaEventDestroy(h->ev_pipe_join);
    if (h->pipe_stream) cudaStreamDestroy(h->pipe_stream);
    if (h->slab) cudaFree(h->slab);
    delete h;
    return MSOC_OK;
}

int msoc_reset(msoc_handle *h, const uint8_t *d_mask, int mode, int has_seed, uint64_t seed, float *d_obs_out, void *stream)
{
    if (!h) return fail(MSOC_ERR_INVALID, "msoc_reset: null handle");
    if (mode < 0 || mode > 2) return fail(MSOC_ERR_INVALID, "msoc_reset: bad mode");
    DeviceGuard guard(h->device);
    ResetParams P;
    P.A = h->A; P.cfg = h->cfg; P.mask = d_mask; P.obs_out = d_obs_out; P.ctl = h->d_ctl;
    P.global_offset = h->global_offset; P.seed = seed; P.mode = mode; P.has_seed = has_seed;
    msoc_reset_kernel<<<grid_for(h->n), BLOCK, 0, (cudaStream_t)stream>>>(P);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return MSOC_OK;
}

/* the two launches of one step over the envs [e0, e1), both on `st` */
static int launch_step(msoc_handle *h, StepParams &P, int64_t e0, int64_t e1, int chunk, cudaStream_t st)
{
    P.e0 = e0; P.e1 = e1; P.chunk = chunk;
    const int64_t m = e1 - e0;
    P.heavy_lanes = h->heavy_lanes_override; /* 0: the contact kernel chooses by the number of heavy envs */
    /* Both kernels are launched programmatically dependent on whatever kernel precedes them in the stream: their grids
       are set up while that kernel drains and their blocks wait in cudaGridDependencySynchronize() until it is complete
       (a few microseconds of launch latency per step, which is what small batches are made of). */
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t lf = {};
    lf.gridDim = dim3((unsigned)((m + FAST_BLOCK - 1) / FAST_BLOCK));
    lf.blockDim = dim3(FAST_BLOCK);
    lf.dynamicSmemBytes = FAST_SMEM_BYTES;
    lf.stream = st;
    lf.attrs = at; lf.numAttrs = 1;
    CUDA_TRY(cudaLaunchKernelEx(&lf, msoc_step_fast_kernel, P));
    g_launches++;
    /* one persistent grid for all contact classes, right behind the streaming kernel */
    const int64_t resident = (int64_t)h->sm_count * h->blocks_per_sm;
    const int64_t blocks_max = (m + HEAVY_BLOCK - 1) / HEAVY_BLOCK;
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3((unsigned)(blocks_max < resident ? blocks_max : resident));
    lc.blockDim = dim3(HEAVY_BLOCK);
    lc.dynamicSmemBytes = STEP_SMEM_BYTES;
    lc.stream = st;
    lc.attrs = at; lc.numAttrs = 1;
    CUDA_TRY(cudaLaunchKernelEx(&lc, msoc_step_contact_kernel, P));
    g_launches++;
    return MSOC_OK;
}

static void fill_step_params(msoc_handle *h, StepParams &P, const float *d_actions, float *d_obs_out, float *d_frames_out,
                             float *d_reward, uint8_t *d_done, int8_t *d_goal, int32_t *d_score, uint32_t flags)
{
    P.A = h->A; P.cfg = h->cfg; P.actions = d_actions; P.obs_out = d_obs_out; P.frames_out = d_frames_out;
    P.reward = d_reward; P.done = d_done; P.goal = d_goal; P.score = d_score; P.stats = h->d_stats;
    P.global_offset = h->global_offset; P.flags = flags;
    P.ctl = h->d_ctl; P.list = h->d_list; P.list2 = h->d_list2;
}

int msoc_step(msoc_handle *h, const float *d_actions, float *d_obs_out, float *d_reward,
              uint8_t *d_done, int8_t *d_goal, int32_t *d_score, uint32_t flags, void *stream)
{
    if (!h || !d_actions || !d_obs_out || !d_reward || !d_done || !d_goal)
        return fail(MSOC_ERR_INVALID, "msoc_step: null argument");
    DeviceGuard guard(h->device);
    StepParams P;
    fill_step_params(h, P, d_actions, d_obs_out, nullptr, d_reward, d_done, d_goal, d_score, flags);
    return launch_step(h, P, 0, h->n, -1, (cudaStream_t)stream);
}

/* Host-buffer step: the envs are cut into chunks that alternate between two stream lanes, so that the H2D copy and the
   kernels of one chunk run while the D2H copies of the previous one are still on the wire. */
static int step_host_impl(msoc_handle *h, const float *h_actions, float *h_obs, float *h_frames, float *h_reward, uint8_t *h_done,
                          int8_t *h_goal, int32_t *h_score, uint32_t flags, void *stream)
{
    if (!h || !h_actions) return fail(MSOC_ERR_INVALID, "msoc_step_host: null argument");
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (h_frames && !h->d_frames) {
        cudaError_t ce = cudaMalloc(&h->d_frames, (size_t)h->n * 4 * FRAME * sizeof(float));
        if (ce != cudaSuccess) return fail(MSOC_ERR_ALLOC, "msoc_step_host_frames: cudaMalloc", ce);
    }
    StepParams P;
    fill_step_params(h, P, h->d_act, h->d_obs, h_frames ? h->d_frames : nullptr, h->d_rew, h->d_done, h->d_goal, h->d_score, flags);
    /* chunks of at least 64 Ki envs, at most MAX_CHUNKS */
    int chunks = (int)(h->n / 65536);
    if (chunks < 1) chunks = 1;
    if (chunks > MAX_CHUNKS) chunks = MAX_CHUNKS;
    const int64_t per = ((h->n + chunks - 1) / chunks + 127) & ~(int64_t)127;
    auto copies_back = [&](int64_t e0, int64_t m, cudaStream_t s) -> int {
        if (h_obs) CUDA_TRY(cudaMemcpyAsync(h_obs + e0 * 4 * OBS, h->d_obs + e0 * 4 * OBS, (size_t)m * 4 * OBS * sizeof(float), cudaMemcpyDeviceToHost, s));
        if (h_frames) CUDA_TRY(cudaMemcpyAsync(h_frames + e0 * 4 * FRAME, h->d_frames + e0 * 4 * FRAME, (size_t)m * 4 * FRAME * sizeof(float), cudaMemcpyDeviceToHost, s));
        if (h_reward) CUDA_TRY(cudaMemcpyAsync(h_reward + e0 * 2, h->d_rew + e0 * 2, (size_t)m * 2 * sizeof(float), cudaMemcpyDeviceToHost, s));
        if (h_done) CUDA_TRY(cudaMemcpyAsync(h_done + e0, h->d_done + e0, (size_t)m, cudaMemcpyDeviceToHost, s));
        if (h_goal) CUDA_TRY(cudaMemcpyAsync(h_goal + e0, h->d_goal + e0, (size_t)m, cudaMemcpyDeviceToHost, s));
        if (h_score) CUDA_TRY(cudaMemcpyAsync(h_score + e0 * 2, h->d_score + e0 * 2, (size_t)m * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        return MSOC_OK;
    };
    if (chunks == 1) {
        CUDA_TRY(cudaMemcpyAsync(h->d_act, h_actions, (size_t)h->n * 12 * sizeof(float), cudaMemcpyHostToDevice, st));
        int rc = launch_step(h, P, 0, h->n, -1, st);
        if (rc != MSOC_OK) return rc;
        rc = copies_back(0, h->n, st);
        if (rc != MSOC_OK) return rc;
        CUDA_TRY(cudaStreamSynchronize(st));
        return MSOC_OK;
    }
    CUDA_TRY(cudaMemsetAsync(h->d_ctl + CTL_CHUNK0, 0, CTL_WORDS * MAX_CHUNKS * sizeof(int), st));
    CUDA_TRY(cudaEventRecord(h->ev_pipe_fork, st));
    CUDA_TRY(cudaStreamWaitEvent(h->pipe_stream, h->ev_pipe_fork, 0));
    for (int c = 0; c < chunks; c++) {
        const int64_t e0 = c * per, e1 = (e0 + per < h->n) ? e0 + per : h->n;
        if (e0 >= e1) break;
        cudaStream_t s = (c & 1) ? h->pipe_stream : st;
        CUDA_TRY(cudaMemcpyAsync(h->d_act + e0 * 12, h_actions + e0 * 12, (size_t)(e1 - e0) * 12 * sizeof(float), cudaMemcpyHostToDevice, s));
        int rc = launch_step(h, P, e0, e1, c, s);
        if (rc != MSOC_OK) return rc;
        rc = copies_back(e0, e1 - e0, s);
        if (rc != MSOC_OK) return rc;
    }
    CUDA_TRY(cudaEventRecord(h->ev_pipe_join, h->pipe_stream));
    CUDA_TRY(cudaStreamWaitEvent(st, h->ev_pipe_join, 0));
    msoc_advance_kernel<<<1, 32, 0, st>>>(h->d_ctl);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(st));
    return MSOC_OK;
}

int msoc_step_host(msoc_handle *h, const float *h_actions, float *h_obs, float *h_reward, uint8_t *h_done,
                   int8_t *h_goal, int32_t *h_score, uint32_t flags, void *stream)
{
    return step_host_impl(h, h_actions, h_obs, nullptr, h_reward, h_done, h_goal, h_score, flags, stream);
}

int msoc_step_host_frames(msoc_handle *h, const float *h_actions, float *h_frames, float *h_reward, uint8_t *h_done,
                          int8_t *h_goal, int32_t *h_score, uint32_t flags, void *stream)
{
    if (!h_frames) return fail(MSOC_ERR_INVALID, "msoc_step_host_frames: null frame buffer");
    return step_host_impl(h, h_actions, nullptr, h_frames, h_reward, h_done, h_goal, h_score, flags, stream);
}

int msoc_reset_host(msoc_handle *h, const uint8_t *h_mask, int mode, int has_seed, uint64_t seed, float *h_obs, void *stream)
{
    if (!h) return fail(MSOC_ERR_INVALID, "msoc_reset_host: null handle");
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)h->n;
    if (h_mask) CUDA_TRY(cudaMemcpyAsync(h->d_mask, h_mask, n, cudaMemcpyHostToDevice, st));
    int rc = msoc_reset(h, h_mask ? h->d_mask : nullptr, mode, has_seed, seed, h->d_obs, stream);
    if (rc != MSOC_OK) return rc;
    if (h_obs) CUDA_TRY(cudaMemcpyAsync(h_obs, h->d_obs, n * 4 * OBS * sizeof(float), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return MSOC_OK;
}

int msoc_read_counters(msoc_handle *h, int32_t *h_score, int32_t *h_steps, void *stream)
{
    if (!h) return fail(MSOC_ERR_INVALID, "msoc_read_counters: null handle");
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    int step = 0;
    CUDA_TRY(cudaMemcpyAsync(&step, h->d_ctl + CTL_STEP_FAST, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (h_steps) {
        std::vector<float4> tmp((size_t)h->n);
        CUDA_TRY(cudaMemcpy2DAsync(tmp.data(), sizeof(float4), h->A.pose[step % 3] + 7, POSE_F4 * sizeof(float4), sizeof(float4), (size_t)h->n, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        for (int64_t i = 0; i < h->n; i++) { int32_t v; memcpy(&v, &tmp[(size_t)i].y, 4); h_steps[i] = v; }
    }
    if (h_score) {
        CUDA_TRY(cudaMemcpyAsync(h_score, h->A.score, (size_t)h->n * sizeof(int2), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
    }
    return MSOC_OK;
}

static int ensure_stage(msoc_handle *h, size_t bytes)
{
    if (h->stage_bytes >= bytes) return MSOC_OK;
    if (h->d_stage) cudaFree(h->d_stage);
    h->d_stage = nullptr; h->stage_bytes = 0;
    cudaError_t ce = cudaMalloc(&h->d_stage, bytes);
    if (ce != cudaSuccess) return fail(MSOC_ERR_ALLOC, "state staging cudaMalloc", ce);
    h->stage_bytes = bytes;
    return MSOC_OK;
}

static int check_idx(const msoc_handle *h, const int64_t *idx, int64_t n, const char *who)
{
    if (!h || !idx || n <= 0) return fail(MSOC_ERR_INVALID, who);
    for (int64_t i = 0; i < n; i++)
        if (idx[i] < 0 || idx[i] >= h->n) return fail(MSOC_ERR_INVALID, "env index out of range");
    return MSOC_OK;
}

int msoc_get_state(msoc_handle *h, const int64_t *h_idx, int64_t n, msoc_env_state *h_out)
{
    int rc = check_idx(h, h_idx, n, "msoc_get_state: bad argument");
    if (rc != MSOC_OK) return rc;
    if (!h_out) return fail(MSOC_ERR_INVALID, "msoc_get_state: null output");
    DeviceGuard guard(h->device);
    const size_t ib = align_up((size_t)n * sizeof(int64_t)), sb = (size_t)n * sizeof(msoc_env_state);
    rc = ensure_stage(h, ib + sb);
    if (rc != MSOC_OK) return rc;
    int64_t *d_idx = (int64_t *)h->d_stage;
    msoc_env_state *d_s = (msoc_env_state *)((char *)h->d_stage + ib);
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(d_idx, h_idx, (size_t)n * sizeof(int64_t), cudaMemcpyHostToDevice));
    msoc_get_state_kernel<<<(unsigned)((n + 127) / 128), 128>>>(h->A, h->d_ctl, d_idx, n, d_s);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(h_out, d_s, sb, cudaMemcpyDeviceToHost));
    return MSOC_OK;
}

static float wrap_host(float x)
{
    double a = (double)x;
    if (a > 3.14159274101257324 || a < -3.14159274101257324) a = atan2(sin(a), cos(a));
    return (float)a;
}

int msoc_set_state(msoc_handle *h, const int64_t *h_idx, int64_t n, const msoc_env_state *h_in)
{
    int rc = check_idx(h, h_idx, n, "msoc_set_state: bad argument");
    if (rc != MSOC_OK) return rc;
    if (!h_in) return fail(MSOC_ERR_INVALID, "msoc_set_state: null input");
    DeviceGuard guard(h->device);
    if (!h->d_inject) { /* records of injected states whose pose differs from the newest emitted frame's */
        cudaError_t ce = cudaMalloc(&h->d_inject, (size_t)h->n * POSE_F4 * sizeof(float4));
        if (ce != cudaSuccess) return fail(MSOC_ERR_ALLOC, "msoc_set_state: cudaMalloc", ce);
        h->A.inject = (float4 *)h->d_inject;
    }
    const size_t ib = align_up((size_t)n * sizeof(int64_t)), sb = (size_t)n * sizeof(msoc_env_state);
    rc = ensure_stage(h, ib + sb);
    if (rc != MSOC_OK) return rc;
    int64_t *d_idx = (int64_t *)h->d_stage;
    msoc_env_state *d_s = (msoc_env_state *)((char *)h->d_stage + ib);
    /* wrap angles on the host in double (the reference keeps them unwrapped) */
    std::vector<msoc_env_state> tmp(h_in, h_in + n);
    for (auto &S : tmp)
        for (int i = 0; i < 4; i++) {
            S.ang[i] = wrap_host(S.ang[i]);
            for (int k = 0; k < 2; k++) S.hist_ang[k][i] = wrap_host(S.hist_ang[k][i]);
        }
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(d_idx, h_idx, (size_t)n * sizeof(int64_t), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d_s, tmp.data(), sb, cudaMemcpyHostToDevice));
    msoc_set_state_kernel<<<(unsigned)((n + 127) / 128), 128>>>(h->A, h->d_ctl, d_idx, n, d_s);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaDeviceSynchronize());
    return MSOC_OK;
}

int msoc_get_obs_host(msoc_handle *h, const int64_t *h_idx, int64_t n, float *h_obs)
{
    int rc = check_idx(h, h_idx, n, "msoc_get_obs_host: bad argument");
    if (rc != MSOC_OK) return rc;
    if (!h_obs) return fail(MSOC_ERR_INVALID, "msoc_get_obs_host: null output");
    DeviceGuard guard(h->device);
    const size_t ib = align_up((size_t)n * sizeof(int64_t)), ob = (size_t)n * 4 * OBS * sizeof(float);
    rc = ensure_stage(h, ib + ob);
    if (rc != MSOC_OK) return rc;
    int64_t *d_idx = (int64_t *)h->d_stage;
    float *d_o = (float *)((char *)h->d_stage + ib);
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(d_idx, h_idx, (size_t)n * sizeof(int64_t), cudaMemcpyHostToDevice));
    msoc_gather_obs_kernel<<<(unsigned)((n * 4 * OBS + 255) / 256), 256>>>(h->d_obs, d_idx, n, d_o);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(h_obs, d_o, ob, cudaMemcpyDeviceToHost));
    return MSOC_OK;
}

int msoc_device_buffers(msoc_handle *h, msoc_buffers *out)
{
    if (!h || !out) return fail(MSOC_ERR_INVALID, "msoc_device_buffers: null argument");
    out->obs = h->d_obs; out->actions = h->d_act; out->reward = h->d_rew; out->done = h->d_done;
    out->goal = h->d_goal; out->score = h->d_score; out->mask = h->d_mask; out->stats = h->d_stats;
    return MSOC_OK;
}

int msoc_stats_device(msoc_handle *h, double *d_out, int reset, void *stream)
{
    if (!h || !d_out) return fail(MSOC_ERR_INVALID, "msoc_stats_device: null argument");
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaMemcpyAsync(d_out, h->d_stats, 8 * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (reset) CUDA_TRY(cudaMemsetAsync(h->d_stats, 0, 8 * sizeof(double), st));
    return MSOC_OK;
}

int msoc_policy_inputs(const float *d_obs, int64_t n_envs, const float *d_shift, const float *d_inv_std, void *d_x_out,
                       void *d_raw_out, double *d_moments, void *stream)
{
    if (!d_obs || !d_shift || !d_inv_std || !d_x_out || n_envs <= 0) return fail(MSOC_ERR_INVALID, "msoc_policy_inputs: bad argument");
    int dev = 0, sms = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int64_t want = (n_envs + PI_Y - 1) / PI_Y, cap = (int64_t)sms * 7; /* 7 blocks of 264 threads per SM; a thread sums n / (4 * grid) rows in fp32 */
    msoc_policy_inputs_kernel<<<(unsigned)(want < cap ? want : cap), dim3(PI_X, PI_Y), 0, (cudaStream_t)stream>>>(
        d_obs, n_envs, d_shift, d_inv_std, (__nv_bfloat162 *)d_x_out, (__nv_bfloat162 *)d_raw_out, d_moments);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return MSOC_OK;
}

int msoc_last_class_counts(msoc_handle *h, int32_t h_out[4], void *stream)
{
    if (!h || !h_out) return fail(MSOC_ERR_INVALID, "msoc_last_class_counts: null argument");
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaMemcpyAsync(h_out, h->d_ctl + CTL_LAST_COUNTS, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return MSOC_OK;
}

int msoc_stats_read(msoc_handle *h, msoc_stats *h_out, int reset, void *stream)
{
    if (!h || !h_out) return fail(MSOC_ERR_INVALID, "msoc_stats_read: null argument");
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaMemcpyAsync(h_out, h->d_stats, 8 * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (reset) CUDA_TRY(cudaMemsetAsync(h->d_stats, 0, 8 * sizeof(double), st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return MSOC_OK;
}

} /* extern "C" */
