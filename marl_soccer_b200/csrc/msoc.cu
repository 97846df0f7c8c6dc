/*
 * msoc.cu -- B200 (sm_100a) kernels and the C-ABI (include/msoc.h) of the batched 2v2 soccer
 * simulator.  Replaces the reference's per-env Python/pymunk update loop
 * (soccer_simulation/marl_vecenv.py:30-68 -> soccer_env.py:100-154 -> game/game.py:378-437 ->
 * pymunk Space.step) with one fused device step per vectorised step: three launches, see "The fused step".
 *
 * Data layout: per-env records of float4s over N envs (msoc::Arrays in step_core.cuh); one thread owns
 * one env, one warp a batch of 32 envs (consecutive in the streaming kernel, listed in the contact
 * kernels).  Observations (N,4,66) are written per warp: the new 22-float frames are staged through
 * shared memory and the 3-frame stack is shifted with coalesced 8-byte accesses (rows of 264 B are
 * 8-byte aligned), so every DRAM sector of an env's 1 056-byte block is written once.
 *
 * There is no CPU fallback in this library: every entry point needs a CUDA device.
 */
#include <cuda_runtime.h>
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
struct msoc_handle {
    int device;
    int64_t n;
    uint64_t global_offset;
    SimCfg cfg;
    Arrays A;
    int cur; /* which half of the ping-pong arbiter cache is current */
    int sm_count, blocks_per_sm, light_blocks_per_sm; /* persistent grids of the contact and light kernels */
    void *slab;
    /* internal I/O buffers of the host-buffer API */
    float *d_obs, *d_act, *d_rew;
    uint8_t *d_done, *d_mask;
    int8_t *d_goal;
    int32_t *d_score;
    double *d_stats; /* 8 doubles, msoc_stats layout */
    int *d_ctl;  /* 2 x CTL_WORDS counters of the step kernels, alternating between steps */
    int *d_list; /* n ints: contact list of the step in flight */
    int step_parity;
    cudaStream_t aux_stream;        /* the light kernel runs here, beside the heavy contact kernel */
    cudaEvent_t ev_listed, ev_light; /* fork after the fast kernel / join after the light kernel */
    void *d_stage; size_t stage_bytes; /* get/set_state staging */
};

constexpr int WARPS_PER_BLOCK = 4;
constexpr int BLOCK = WARPS_PER_BLOCK * 32; /* reset kernel */
constexpr int STEP_BLOCK = 128;             /* step kernel: envs (= threads) per block */
#ifndef MSOC_STEP_MIN_BLOCKS
#define MSOC_STEP_MIN_BLOCKS 3
#endif
constexpr int STEP_MIN_BLOCKS = MSOC_STEP_MIN_BLOCKS; /* resident blocks per SM the register budget is set for */

/* ------------------------------------------------------------------ coalesced observation rows */
constexpr int ENV_STRIDE = (SCRATCH_WORDS > 88 ? SCRATCH_WORDS : 88) | 1; /* floats of per-lane scratch in the step kernel: >= 88 (4 new frames) and >= SCRATCH_WORDS; odd: no bank conflicts */
constexpr int RESET_STRIDE = 89; /* reset kernel: frame staging only */

/* One warp writes the stacked observations of the (up to) 32 envs its lanes own.  Lane l owns env
   `my_env` (envs need not be consecutive).  An env's 4 rows are 1 056 contiguous, 32-byte aligned bytes =
   132 float2; row = 33 float2: [frame t-2 | frame t-1 | frame t].  Frames t-2, t-1 come from obs_in
   shifted by one frame (soccer_env.py:134-137), frame t from shared memory (lane l's 4 frames at
   s_new + l*ENV_STRIDE); envs in `fresh` (reset / auto-reset, soccer_env.py:92-96) get three copies of the
   new frame.  Lane j moves float2 j of every row: fully contiguous 8-byte accesses (the +22-float shift is
   only 8-byte aligned, so float2 is the widest access that keeps loads and stores both contiguous).
   Safe when obs_out == obs_in: within a batch of rows all loads precede all stores, and one env's rows
   are only ever touched by the warp that owns it. */
template <int ENV_STRIDE>
__device__ __forceinline__ void write_obs_tile(const float2 *in2, float2 *out2, const float *s_new, uint32_t mask,
                                               uint32_t fresh, int64_t my_env, int lane)
{
    const bool lane_hist = lane < 22;                   /* float2 0-21 of a row: history; 22-31: frame t */
    const int j_lane = lane < 11 ? lane : lane < 22 ? lane - 11 : lane - 22;
    const float *s_lane = s_new + 2 * j_lane;           /* this lane's float2 of the staged frames */
#ifndef MSOC_WRITER_ENVS
#define MSOC_WRITER_ENVS 8
#endif
    constexpr int EB = MSOC_WRITER_ENVS; /* envs per batch: 4*EB row loads in flight per lane */
#pragma unroll 1
    for (int l0 = 0; l0 < 32; l0 += EB) {
        const uint32_t mb = (mask >> l0) & ((1u << EB) - 1u);
        if (mb == 0u) continue;
        float2 v[EB][4];
        float2 *po[EB];
#pragma unroll
        for (int w = 0; w < EB; w++) {
            const int64_t ew = __shfl_sync(0xffffffffu, my_env, l0 + w);
            const float2 *pi = in2 + ew * 132 + lane + 11;
            po[w] = out2 + ew * 132 + lane;
            const bool ld = ((mb >> w) & 1u) && lane_hist && !((fresh >> (l0 + w)) & 1u);
#pragma unroll
            for (int a = 0; a < 4; a++) {
                const float *sp = s_lane + (l0 + w) * ENV_STRIDE + a * 22;
                v[w][a] = make_float2(sp[0], sp[1]);
                if (ld) v[w][a] = pi[a * 33];
            }
        }
        __syncwarp();
#pragma unroll
        for (int w = 0; w < EB; w++) {
            if (!((mb >> w) & 1u)) continue;
#pragma unroll
            for (int a = 0; a < 4; a++) po[w][a * 33] = v[w][a];
        }
    }
    /* last float2 of every row (frame t, floats 20-21): lane l writes its own env's four */
    if ((mask >> lane) & 1u) {
#pragma unroll
        for (int a = 0; a < 4; a++) {
            const float *sp = s_new + lane * ENV_STRIDE + a * 22 + 20;
            out2[my_env * 132 + a * 33 + 32] = make_float2(sp[0], sp[1]);
        }
    }
}

/* ---------------------------------------------------------------------------- the fused step */
struct StepParams {
    Arrays A;
    SimCfg cfg;
    const float *actions; /* (N,4,3) */
    const float *obs_in;  /* (N,4,66) */
    float *obs_out;       /* (N,4,66) */
    float *reward;        /* (N,2) */
    uint8_t *done;        /* (N) */
    int8_t *goal;         /* (N) */
    int32_t *score;       /* (N,2) or null */
    double *stats;        /* 8 */
    int *ctl;             /* counters of this step (all start at 0): CTL_* below */
    int *ctl_other;       /* counters of the next step: zeroed by this one */
    int *list;            /* N slots: envs that need the contact path -- light from the front, heavy from the back */
    uint64_t global_offset;
    uint32_t flags;
    int cur;
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
struct Tally { int done, goals_b, goals_r, contacts, overflow, envs; float ret; };

__device__ __forceinline__ bool step_one_env(const int MODE, const StepParams &P, int64_t e, Env &E, Work &W, int &load, bool &fresh,
                                             Tally &T)
{
    float act[12];
    const float4 *a4 = reinterpret_cast<const float4 *>(P.actions + e * 12);
    const float4 x0 = __ldg(a4), x1 = __ldg(a4 + 1), x2 = __ldg(a4 + 2);
    act[0] = x0.x; act[1] = x0.y; act[2] = x0.z; act[3] = x0.w;
    act[4] = x1.x; act[5] = x1.y; act[6] = x1.z; act[7] = x1.w;
    act[8] = x2.x; act[9] = x2.y; act[10] = x2.z; act[11] = x2.w;
    load_env(P.A, e, E);
    StepOut out;
    if (!env_step(MODE, E, act, P.cfg, P.A, P.cur, e, P.global_offset + (uint64_t)e, P.flags, W, out, load)) return false;
    store_env(P.A, e, E);
    reinterpret_cast<float2 *>(P.reward)[e] = make_float2(out.reward, out.reward);
    P.done[e] = out.done;
    P.goal[e] = out.goal;
    if (P.score != nullptr) reinterpret_cast<int2 *>(P.score)[e] = make_int2(out.score_b, out.score_r);
    fresh = out.fresh_episode;
    T.done += out.done; T.goals_b += out.goal > 0; T.goals_r += out.goal < 0; T.envs += 1;
    T.contacts += out.n_contacts; T.overflow += out.overflow; T.ret += out.finished_return;
    return true;
}

/* The fused step is three launches, each tuned for its share of the work.
   msoc_step_fast_kernel     streams over ALL envs, thread t of block b steps env b*128 + t in contact-free
                             mode and its warp writes the observation rows.  No contact code is compiled into
                             it: few registers, small shared memory, small instruction footprint -> many
                             resident warps to hide the HBM latency.  Envs whose broad phase finds a candidate
                             pair (~27 % in the benchmark mix) write nothing and are appended, one atomic per
                             warp and class, to the step's contact list: light (exactly one agent x wall pair,
                             the bulk) from the front, heavy (anything else) from the back.
   msoc_step_light_kernel    every warp of a persistent grid takes batches of 32 light envs, thread per env: one
                             narrow-phase call and a register-only single-body impulse solver (at most two
                             contacts); again no solver scratch.
   msoc_step_contact_kernel  every warp of a persistent grid takes batches of 32 heavy envs and every thread steps
                             one of them in full mode: narrow phase over all candidate pairs, arbiter cache,
                             10-iteration impulse solver with bodies and contacts in shared memory; its warp
                             writes the rows.
   The divergent, latency-bound contact work therefore always runs on full warps of similar work, and each
   kind of work gets the register / shared-memory budget (hence the occupancy) that suits it.  The two contact
   kernels only depend on the fast kernel's lists and are launched on two streams (msoc_step). */
enum { CTL_LIGHT = 0, CTL_HEAVY = 1, CTL_NEXT_HEAVY = 2, CTL_NEXT_LIGHT = 3, CTL_WORDS = 4 };

constexpr int FAST_STRIDE = 89; /* floats of per-lane frame staging in the fast kernel (88, odd: no bank conflicts) */
#ifndef MSOC_FAST_BLOCK
#define MSOC_FAST_BLOCK 128
#endif
constexpr int FAST_BLOCK = MSOC_FAST_BLOCK; /* envs (= threads) per block of the fast kernel */
constexpr size_t FAST_SMEM_BYTES = (size_t)FAST_BLOCK * FAST_STRIDE * sizeof(float);
#ifndef MSOC_FAST_MIN_BLOCKS
#define MSOC_FAST_MIN_BLOCKS 4
#endif

/* Appends the envs of the lanes with `want` set at `*tail` (one atomic per warp); slot(i) maps the i-th
   entry to its address. */
template <typename SlotFn>
__device__ __forceinline__ void push_warp(int *tail, bool want, int env, int lane, SlotFn slot)
{
    const uint32_t m = __ballot_sync(0xffffffffu, want);
    if (m == 0u) return;
    int base = 0;
    if (lane == 0) base = atomicAdd(tail, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (want) *slot(base + __popc(m & ((1u << lane) - 1u))) = env;
}

/* warp reduce of the per-rollout statistics (marl-soccer.ipynb:411-429), one atomic per warp and counter */
__device__ __forceinline__ void flush_tally(const Tally &T, double *stats, int lane)
{
    int nd = T.done, gb = T.goals_b, gr = T.goals_r, nc = T.contacts, ov = T.overflow, na = T.envs;
    float ret = T.ret;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        nd += __shfl_xor_sync(0xffffffffu, nd, o); gb += __shfl_xor_sync(0xffffffffu, gb, o);
        gr += __shfl_xor_sync(0xffffffffu, gr, o); nc += __shfl_xor_sync(0xffffffffu, nc, o);
        ov += __shfl_xor_sync(0xffffffffu, ov, o); na += __shfl_xor_sync(0xffffffffu, na, o);
        ret += __shfl_xor_sync(0xffffffffu, ret, o);
    }
    if (lane == 0) {
        if (nd) { atomicAdd(stats + 0, (double)nd); atomicAdd(stats + 1, (double)ret); }
        if (gb) atomicAdd(stats + 2, (double)gb);
        if (gr) atomicAdd(stats + 3, (double)gr);
        if (na) atomicAdd(stats + 4, (double)na);
        if (nc) atomicAdd(stats + 5, (double)nc);
        if (ov) atomicAdd(stats + 6, (double)ov);
    }
}

__global__ void __launch_bounds__(FAST_BLOCK, MSOC_FAST_MIN_BLOCKS) msoc_step_fast_kernel(const __grid_constant__ StepParams P)
{
    extern __shared__ float s_dyn[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float *s_warp = s_dyn + warp * 32 * FAST_STRIDE;
    if (blockIdx.x == 0 && tid < CTL_WORDS) P.ctl_other[tid] = 0;
    MSOC_TL_BEGIN();
    const int64_t my_env = (int64_t)blockIdx.x * FAST_BLOCK + tid;
    const bool have = my_env < P.A.n;
    Tally T; T.done = T.goals_b = T.goals_r = T.contacts = T.overflow = T.envs = 0; T.ret = 0.0f;
    Work W; /* never touched in contact-free mode */
    W.ovf = nullptr; W.body = W.pool = W.geom = W.old = nullptr; W.pool_count = nullptr;
    bool fresh = false, ok = false;
    int load = 0;
    {
        Env E;
        if (have) ok = step_one_env(MODE_FAST, P, my_env, E, W, load, fresh, T);
        if (ok) make_frames<22>(E, P.cfg, s_warp + lane * FAST_STRIDE);
    }
    const uint32_t mask = __ballot_sync(0xffffffffu, ok);
    const uint32_t fmask = __ballot_sync(0xffffffffu, ok && fresh);
    __syncwarp();
    if (mask) write_obs_tile<FAST_STRIDE>(reinterpret_cast<const float2 *>(P.obs_in), reinterpret_cast<float2 *>(P.obs_out), s_warp, mask,
                                          fmask, my_env, lane);
    const bool declined = have && !ok;
    push_warp(P.ctl + CTL_LIGHT, declined && load == 0, (int)my_env, lane, [&](int i) { return P.list + i; });
    push_warp(P.ctl + CTL_HEAVY, declined && load != 0, (int)my_env, lane, [&](int i) { return P.list + (P.A.n - 1 - i); });
    flush_tally(T, P.stats, lane);
    if ((blockIdx.x & 15) == 0 && warp == 0) MSOC_TL_END(0);
}

/* Light envs (exactly one agent x wall candidate pair, ~82 % of the contact envs): thread t of a batch steps
   one listed env with the register-only single-body solver of step_core.cuh (MODE_LIGHT).  Like the fast
   kernel it needs no solver scratch: shared memory only stages the frames. */
#ifndef MSOC_LIGHT_BLOCK
#define MSOC_LIGHT_BLOCK 64 /* small blocks: they slip into an SM as soon as one heavy block has left it */
#endif
constexpr int LIGHT_BLOCK = MSOC_LIGHT_BLOCK;
constexpr int LIGHT_MIN_BLOCKS = 512 / LIGHT_BLOCK; /* 128 registers per thread */
constexpr size_t LIGHT_SMEM_BYTES = (size_t)LIGHT_BLOCK * FAST_STRIDE * sizeof(float);
__global__ void __launch_bounds__(LIGHT_BLOCK, LIGHT_MIN_BLOCKS) msoc_step_light_kernel(const __grid_constant__ StepParams P)
{
    extern __shared__ float s_dyn[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float *s_warp = s_dyn + warp * 32 * FAST_STRIDE;
    const int n_light = P.ctl[CTL_LIGHT]; /* final: the fast kernel has finished */
    const int batches = (n_light + 31) / 32;
    Tally T; T.done = T.goals_b = T.goals_r = T.contacts = T.overflow = T.envs = 0; T.ret = 0.0f;
    Work W; /* never touched in light mode */
    W.ovf = nullptr; W.body = W.pool = W.geom = W.old = nullptr; W.pool_count = nullptr;
#pragma unroll 1
    while (true) {
        /* every warp takes its own batches of 32 envs, handed out dynamically: this kernel runs beside the heavy
           contact kernel and its blocks start whenever an SM has room for them */
        int b = 0;
        if (lane == 0) b = atomicAdd(P.ctl + CTL_NEXT_LIGHT, 1);
        b = __shfl_sync(0xffffffffu, b, 0);
        if (b >= batches) break;
        const int idx = b * 32 + lane;
        const bool have = idx < n_light;
        const int64_t my_env = have ? (int64_t)P.list[idx] : 0;
        MSOC_TL_BEGIN();
        bool fresh = false, ok = false;
        int load = 0;
        {
            Env E;
            if (have) ok = step_one_env(MODE_LIGHT, P, my_env, E, W, load, fresh, T);
            if (ok) make_frames<22>(E, P.cfg, s_warp + lane * FAST_STRIDE);
        }
        const uint32_t mask = __ballot_sync(0xffffffffu, ok);
        const uint32_t fmask = __ballot_sync(0xffffffffu, ok && fresh);
        __syncwarp();
        if (mask) write_obs_tile<FAST_STRIDE>(reinterpret_cast<const float2 *>(P.obs_in), reinterpret_cast<float2 *>(P.obs_out), s_warp,
                                              mask, fmask, my_env, lane);
        __syncwarp(); /* the staging area is rewritten by the next batch */
        MSOC_TL_END(1);
    }
    flush_tally(T, P.stats, lane);
}

/* Per-warp scratch of 32 x ENV_STRIDE floats, time-multiplexed: during the contact solve it holds the
   lanes' solver bodies (30 fields, field-major with stride 32: conflict-free), the warp's pool of
   32 x CON_FAST contact records (15 fields, field-major; an env takes as many records as it has contacts),
   the parked poses and the preloaded arbiter cache entries; afterwards the lanes' four new observation
   frames (lane-major, stride ENV_STRIDE). */
static_assert(SCRATCH_WORDS <= ENV_STRIDE && 88 <= ENV_STRIDE, "per-lane scratch too small");
#ifndef MSOC_HEAVY_BLOCK
#define MSOC_HEAVY_BLOCK 64 /* threads per block of the heavy contact kernel: one warp, so that an SM's registers and shared
                               memory are handed to the light kernel warp by warp as the heavy batches finish */
#endif
constexpr int HEAVY_BLOCK = MSOC_HEAVY_BLOCK;
#ifndef MSOC_HEAVY_MIN_BLOCKS
#define MSOC_HEAVY_MIN_BLOCKS 5 /* register cap 204: ptxas settles on 168 without spills (with 6 it spills 48 B); 6 blocks are resident all the same */
#endif
constexpr int HEAVY_MIN_BLOCKS = MSOC_HEAVY_MIN_BLOCKS;
constexpr size_t STEP_SMEM_BYTES = (size_t)HEAVY_BLOCK * ENV_STRIDE * sizeof(float);

__global__ void __launch_bounds__(HEAVY_BLOCK, HEAVY_MIN_BLOCKS) msoc_step_contact_kernel(const __grid_constant__ StepParams P)
{
    extern __shared__ float s_dyn[];
    __shared__ int s_pool_count[HEAVY_BLOCK / 32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float *s_warp = s_dyn + warp * 32 * ENV_STRIDE;
    const int n_heavy = P.ctl[CTL_HEAVY]; /* final: the fast kernel has finished */
    const int heavy_batches = (n_heavy + 31) / 32;
    const float2 *in2 = reinterpret_cast<const float2 *>(P.obs_in);
    float2 *out2 = reinterpret_cast<float2 *>(P.obs_out);

    Tally T; T.done = T.goals_b = T.goals_r = T.contacts = T.overflow = T.envs = 0; T.ret = 0.0f;
    float ovf_store[MAXC - CON_FAST][CON_FIELDS]; /* local memory, touched only by envs with more than CON_FAST contacts */
    Work W;
    W.ovf = ovf_store;
    W.body = s_warp + lane;
    W.pool = s_warp + BODY_FIELDS * 5 * 32; /* shared by the warp's lanes */
    W.pool_count = &s_pool_count[warp];
    W.geom = s_warp + (BODY_FIELDS * 5 + CON_FIELDS * CON_FAST) * 32 + lane;
    W.old = s_warp + (BODY_FIELDS * 5 + CON_FIELDS * CON_FAST + GEOM_WORDS) * 32 + lane;
#pragma unroll 1
    while (true) {
        /* every WARP takes its own batches of 32 envs, handed out dynamically (the scratch is per warp, so the warps
           of a block never wait for each other; the batches differ a lot in length) */
        int b = 0;
        if (lane == 0) b = atomicAdd(P.ctl + CTL_NEXT_HEAVY, 1);
        b = __shfl_sync(0xffffffffu, b, 0);
        if (b >= heavy_batches) break;
        const int idx = b * 32 + lane;
        const bool have = idx < n_heavy;
        int64_t my_env = 0;
        if (have) my_env = (int64_t)P.list[P.A.n - 1 - idx];
        MSOC_TL_BEGIN();
        bool fresh = false, ok = false;
        int load = 0;
        if (lane == 0) *W.pool_count = 0; /* the warp's contact pool is empty */
        __syncwarp();
        {
            Env E;
            if (have) ok = step_one_env(MODE_FULL, P, my_env, E, W, load, fresh, T);
            __syncwarp(); /* the solver scratch of every lane is dead: reuse it for the frames */
            if (ok) make_frames<22>(E, P.cfg, s_warp + lane * ENV_STRIDE);
        }
        const uint32_t mask = __ballot_sync(0xffffffffu, ok);
        const uint32_t fmask = __ballot_sync(0xffffffffu, ok && fresh);
        __syncwarp();
        if (mask) write_obs_tile<ENV_STRIDE>(in2, out2, s_warp, mask, fmask, my_env, lane);
        __syncwarp();
        MSOC_TL_END(2);
    }
    flush_tally(T, P.stats, lane);
}

/* ------------------------------------------------------------------------------------- reset */
struct ResetParams {
    Arrays A;
    SimCfg cfg;
    const uint8_t *mask; /* N or null */
    float *obs_out;      /* (N,4,66) or null */
    uint64_t global_offset, seed;
    int mode, has_seed, cur;
};

__global__ void __launch_bounds__(BLOCK) msoc_reset_kernel(const __grid_constant__ ResetParams P)
{
    __shared__ float s_frames[WARPS_PER_BLOCK][32 * RESET_STRIDE];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t block_base = (int64_t)blockIdx.x * BLOCK;
    const int64_t e = block_base + threadIdx.x;
    if (block_base + warp * 32 >= P.A.n) return;
    const bool doit = (e < P.A.n) && (P.mask == nullptr || P.mask[e] != 0);
    if (doit) {
        const uint64_t gidx = P.global_offset + (uint64_t)e;
        uint64_t seed; uint32_t sc;
        if (P.has_seed) { seed = P.seed + gidx; sc = 0; P.A.seed[e] = seed; } /* marl_vecenv.py:23: seed + i */
        else { seed = P.A.seed[e]; sc = P.A.spawn_count[e]; }
        Env E;
        env_full_reset(E, P.mode, seed, gidx, sc);
        P.A.spawn_count[e] = sc;
        store_env(P.A, e, E);
        make_frames<22>(E, P.cfg, &s_frames[warp][lane * RESET_STRIDE]);
    }
    const uint32_t m = __ballot_sync(0xffffffffu, doit);
    __syncwarp();
    if (P.obs_out != nullptr && m != 0u) {
        float2 *o2 = reinterpret_cast<float2 *>(P.obs_out);
        write_obs_tile<RESET_STRIDE>(o2, o2, s_frames[warp], m, m, e, lane);
    }
}

/* -------------------------------------------------------------------- state inject / extract */
__global__ void msoc_get_state_kernel(Arrays A, int cur, const int64_t *idx, int64_t n, msoc_env_state *out)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int64_t e = idx[t];
    Env E;
    load_env(A, e, E);
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
        const uint32_t *c = A.cache[cur] + cache_slot(e, (int)j);
        S.cache_info[j] = c[0]; S.cache_jn[j] = __uint_as_float(c[1]); S.cache_jt[j] = __uint_as_float(c[2]);
    }
    out[t] = S;
}

__global__ void msoc_set_state_kernel(Arrays A, int cur, const int64_t *idx, int64_t n, const msoc_env_state *in)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
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
        uint32_t *c = A.cache[cur] + cache_slot(e, (int)j);
        c[0] = S.cache_info[j]; c[1] = __float_as_uint(S.cache_jn[j]); c[2] = __float_as_uint(S.cache_jt[j]);
    }
    store_env(A, e, E);
}

__global__ void msoc_init_kernel(Arrays A, uint64_t seed)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= A.n) return;
    A.seed[e] = seed;
    A.spawn_count[e] = 0;
}

/* ---------------------------------------------------------------------------------- C-ABI */
extern "C" {

const char *msoc_last_error(void) { return g_last_error.c_str(); }
int msoc_version(void) { return MSOC_VERSION; }
uint64_t msoc_launch_count(void) { return g_launches.load(); }
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
}

int msoc_reset(msoc_handle *h, const uint8_t *d_mask, int mode, int has_seed, uint64_t seed, float *d_obs_out, void *stream);

int msoc_create(const msoc_config *cfg, int64_t n_envs, int device, uint64_t seed, uint64_t global_env_offset,
                msoc_handle **out)
{
    if (!cfg || !out || n_envs <= 0) return fail(MSOC_ERR_INVALID, "msoc_create: bad argument");
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
    h->device = device; h->n = n_envs; h->global_offset = global_env_offset; h->cur = 0;
    fill_cfg(cfg, h->cfg);

    const size_t n = (size_t)n_envs;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes); return o; };
    const size_t o_bodies = take(n * 5 * sizeof(float4)), o_misc = take(n * 4 * sizeof(float4));
    const size_t o_bias = take(n * 4 * sizeof(float4));
    const size_t o_seed = take(n * sizeof(uint64_t)), o_sc = take(n * sizeof(uint32_t));
    size_t o_cache[2];
    for (int k = 0; k < 2; k++) o_cache[k] = take(n * MAX_CACHE * 3 * sizeof(uint32_t));
    const size_t o_obs = take(n * 4 * OBS * sizeof(float)), o_act = take(n * 12 * sizeof(float));
    const size_t o_rew = take(n * 2 * sizeof(float)), o_done = take(n), o_goal = take(n), o_mask = take(n);
    const size_t o_score = take(n * 2 * sizeof(int32_t));
    const size_t o_stats = take(8 * sizeof(double));
    const size_t o_tile = take(2 * 4 * sizeof(int));
    const size_t o_list = take(n * sizeof(int));
    const size_t total = off;

    ce = cudaMalloc(&h->slab, total);
    if (ce != cudaSuccess) { delete h; return fail(MSOC_ERR_ALLOC, "msoc_create: cudaMalloc", ce); }
    ce = cudaMemset(h->slab, 0, total);
    if (ce != cudaSuccess) { cudaFree(h->slab); delete h; return fail(MSOC_ERR_CUDA, "msoc_create: cudaMemset", ce); }
    char *base = (char *)h->slab;
    Arrays &A = h->A;
    A.n = n_envs;
    A.bodies = (float4 *)(base + o_bodies); A.misc = (float4 *)(base + o_misc);
    A.bias = (float4 *)(base + o_bias);
    A.seed = (uint64_t *)(base + o_seed); A.spawn_count = (uint32_t *)(base + o_sc);
    for (int k = 0; k < 2; k++) {
        A.cache[k] = (uint32_t *)(base + o_cache[k]);
    }
    h->d_obs = (float *)(base + o_obs); h->d_act = (float *)(base + o_act); h->d_rew = (float *)(base + o_rew);
    h->d_done = (uint8_t *)(base + o_done); h->d_goal = (int8_t *)(base + o_goal); h->d_mask = (uint8_t *)(base + o_mask);
    h->d_score = (int32_t *)(base + o_score);
    h->d_stats = (double *)(base + o_stats);
    h->d_ctl = (int *)(base + o_tile);
    h->d_list = (int *)(base + o_list);

    ce = cudaFuncSetAttribute(msoc_step_contact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)STEP_SMEM_BYTES);
    if (ce == cudaSuccess)
        ce = cudaFuncSetAttribute(msoc_step_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FAST_SMEM_BYTES);
    if (ce == cudaSuccess)
        ce = cudaFuncSetAttribute(msoc_step_light_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LIGHT_SMEM_BYTES);
    if (ce == cudaSuccess)
        ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->light_blocks_per_sm, msoc_step_light_kernel, LIGHT_BLOCK, LIGHT_SMEM_BYTES);
    if (ce != cudaSuccess) { cudaFree(h->slab); delete h; return fail(MSOC_ERR_CUDA, "msoc_create: smem attribute", ce); }
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
    ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->blocks_per_sm, msoc_step_contact_kernel, HEAVY_BLOCK, STEP_SMEM_BYTES);
    if (ce != cudaSuccess || h->blocks_per_sm < 1 || h->sm_count < 1) {
        cudaFree(h->slab); delete h; return fail(MSOC_ERR_CUDA, "msoc_create: occupancy query", ce);
    }
    ce = cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&h->ev_listed, cudaEventDisableTiming);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&h->ev_light, cudaEventDisableTiming);
    if (ce != cudaSuccess) { cudaFree(h->slab); delete h; return fail(MSOC_ERR_CUDA, "msoc_create: stream/event", ce); }
    msoc_init_kernel<<<(unsigned)((n_envs + 255) / 256), 256>>>(A, seed);
    g_launches++;
    /* Game.__init__ -> setup_field -> reset(): first spawn in the default random mode */
    int rc = msoc_reset(h, nullptr, MSOC_MODE_RANDOM, 0, 0, h->d_obs, nullptr);
    if (rc != MSOC_OK) { cudaFree(h->slab); delete h; return rc; }
    ce = cudaDeviceSynchronize();
    if (ce != cudaSuccess) { cudaFree(h->slab); delete h; return fail(MSOC_ERR_CUDA, "msoc_create: init", ce); }
    *out = h;
    return MSOC_OK;
}

int msoc_destroy(msoc_handle *h)
{
    if (!h) return MSOC_OK;
    DeviceGuard guard(h->device);
    cudaDeviceSynchronize();
    if (h->d_stage) cudaFree(h->d_stage);
    if (h->ev_listed) cudaEventDestroy(h->ev_listed);
    if (h->ev_light) cudaEventDestroy(h->ev_light);
    if (h->aux_stream) cudaStreamDestroy(h->aux_stream);
    cudaFree(h->slab);
    delete h;
    return MSOC_OK;
}

int msoc_reset(msoc_handle *h, const uint8_t *d_mask, int mode, int has_seed, uint64_t seed, float *d_obs_out, void *stream)
{
    if (!h) return fail(MSOC_ERR_INVALID, "msoc_reset: null handle");
    if (mode < 0 || mode > 2) return fail(MSOC_ERR_INVALID, "msoc_reset: bad mode");
    DeviceGuard guard(h->device);
    ResetParams P;
    P.A = h->A; P.cfg = h->cfg; P.mask = d_mask; P.obs_out = d_obs_out;
    P.global_offset = h->global_offset; P.seed = seed; P.mode = mode; P.has_seed = has_seed; P.cur = h->cur;
    msoc_reset_kernel<<<grid_for(h->n), BLOCK, 0, (cudaStream_t)stream>>>(P);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return MSOC_OK;
}

int msoc_step(msoc_handle *h, const float *d_actions, const float *d_obs_in, float *d_obs_out, float *d_reward,
              uint8_t *d_done, int8_t *d_goal, int32_t *d_score, uint32_t flags, void *stream)
{
    if (!h || !d_actions || !d_obs_in || !d_obs_out || !d_reward || !d_done || !d_goal)
        return fail(MSOC_ERR_INVALID, "msoc_step: null argument");
    DeviceGuard guard(h->device);
    StepParams P;
    P.A = h->A; P.cfg = h->cfg; P.actions = d_actions; P.obs_in = d_obs_in; P.obs_out = d_obs_out;
    P.reward = d_reward; P.done = d_done; P.goal = d_goal; P.score = d_score; P.stats = h->d_stats;
    P.global_offset = h->global_offset; P.flags = flags; P.cur = h->cur;
    P.ctl = h->d_ctl + 4 * h->step_parity; P.ctl_other = h->d_ctl + 4 * (h->step_parity ^ 1); P.list = h->d_list;
    h->step_parity ^= 1;
    const int64_t n_tiles = (h->n + FAST_BLOCK - 1) / FAST_BLOCK;
    msoc_step_fast_kernel<<<(unsigned)n_tiles, FAST_BLOCK, FAST_SMEM_BYTES, (cudaStream_t)stream>>>(P);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    /* The two contact kernels only depend on the fast kernel's list.  The heavy one (few, long, latency-bound
       batches: one per warp) goes first and stays on the caller's stream; the light one runs beside it on the
       handle's own stream and fills the SMs as the heavy blocks drain.  The caller's stream then waits for it. */
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t resident = (int64_t)h->sm_count * h->blocks_per_sm;
    const int64_t heavy_blocks_max = (h->n + HEAVY_BLOCK - 1) / HEAVY_BLOCK;
    const unsigned grid = (unsigned)(heavy_blocks_max < resident ? heavy_blocks_max : resident);
    CUDA_TRY(cudaEventRecord(h->ev_listed, st));
    msoc_step_contact_kernel<<<grid, HEAVY_BLOCK, STEP_SMEM_BYTES, st>>>(P);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamWaitEvent(h->aux_stream, h->ev_listed, 0));
    const int64_t light_resident = (int64_t)h->sm_count * (h->light_blocks_per_sm > 0 ? h->light_blocks_per_sm : 1);
    const int64_t light_blocks_max = (h->n + LIGHT_BLOCK - 1) / LIGHT_BLOCK;
    msoc_step_light_kernel<<<(unsigned)(light_blocks_max < light_resident ? light_blocks_max : light_resident), LIGHT_BLOCK, LIGHT_SMEM_BYTES, h->aux_stream>>>(P);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(h->ev_light, h->aux_stream));
    CUDA_TRY(cudaStreamWaitEvent(st, h->ev_light, 0));
    h->cur ^= 1;
    CUDA_TRY(cudaGetLastError());
    return MSOC_OK;
}

int msoc_step_host(msoc_handle *h, const float *h_actions, float *h_obs, float *h_reward, uint8_t *h_done,
                   int8_t *h_goal, int32_t *h_score, uint32_t flags, void *stream)
{
    if (!h || !h_actions) return fail(MSOC_ERR_INVALID, "msoc_step_host: null argument");
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)h->n;
    CUDA_TRY(cudaMemcpyAsync(h->d_act, h_actions, n * 12 * sizeof(float), cudaMemcpyHostToDevice, st));
    int rc = msoc_step(h, h->d_act, h->d_obs, h->d_obs, h->d_rew, h->d_done, h->d_goal, h->d_score, flags, stream);
    if (rc != MSOC_OK) return rc;
    if (h_obs) CUDA_TRY(cudaMemcpyAsync(h_obs, h->d_obs, n * 4 * OBS * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (h_reward) CUDA_TRY(cudaMemcpyAsync(h_reward, h->d_rew, n * 2 * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (h_done) CUDA_TRY(cudaMemcpyAsync(h_done, h->d_done, n, cudaMemcpyDeviceToHost, st));
    if (h_goal) CUDA_TRY(cudaMemcpyAsync(h_goal, h->d_goal, n, cudaMemcpyDeviceToHost, st));
    if (h_score) CUDA_TRY(cudaMemcpyAsync(h_score, h->d_score, n * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return MSOC_OK;
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
    std::vector<int4> tmp((size_t)h->n);
    CUDA_TRY(cudaMemcpy2DAsync(tmp.data(), sizeof(int4), h->A.misc + 2, 4 * sizeof(float4), sizeof(int4), (size_t)h->n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    for (int64_t i = 0; i < h->n; i++) {
        if (h_steps) h_steps[i] = tmp[(size_t)i].x;
        if (h_score) { h_score[2 * i] = tmp[(size_t)i].y; h_score[2 * i + 1] = tmp[(size_t)i].z; }
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
    msoc_get_state_kernel<<<(unsigned)((n + 127) / 128), 128>>>(h->A, h->cur, d_idx, n, d_s);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(h_out, d_s, sb, cudaMemcpyDeviceToHost));
    return MSOC_OK;
}

int msoc_set_state(msoc_handle *h, const int64_t *h_idx, int64_t n, const msoc_env_state *h_in)
{
    int rc = check_idx(h, h_idx, n, "msoc_set_state: bad argument");
    if (rc != MSOC_OK) return rc;
    if (!h_in) return fail(MSOC_ERR_INVALID, "msoc_set_state: null input");
    DeviceGuard guard(h->device);
    const size_t ib = align_up((size_t)n * sizeof(int64_t)), sb = (size_t)n * sizeof(msoc_env_state);
    rc = ensure_stage(h, ib + sb);
    if (rc != MSOC_OK) return rc;
    int64_t *d_idx = (int64_t *)h->d_stage;
    msoc_env_state *d_s = (msoc_env_state *)((char *)h->d_stage + ib);
    /* wrap angles on the host in double (the reference keeps them unwrapped) */
    std::vector<msoc_env_state> tmp(h_in, h_in + n);
    for (auto &S : tmp)
        for (int i = 0; i < 4; i++) {
            double a = (double)S.ang[i];
            if (a > 3.14159274101257324 || a < -3.14159274101257324) a = atan2(sin(a), cos(a));
            S.ang[i] = (float)a;
        }
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(d_idx, h_idx, (size_t)n * sizeof(int64_t), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d_s, tmp.data(), sb, cudaMemcpyHostToDevice));
    msoc_set_state_kernel<<<(unsigned)((n + 127) / 128), 128>>>(h->A, h->cur, d_idx, n, d_s);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaDeviceSynchronize());
    return MSOC_OK;
}

int msoc_get_obs_host(msoc_handle *h, const int64_t *h_idx, int64_t n, float *h_obs)
{
    int rc = check_idx(h, h_idx, n, "msoc_get_obs_host: bad argument");
    if (rc != MSOC_OK) return rc;
    DeviceGuard guard(h->device);
    CUDA_TRY(cudaDeviceSynchronize());
    for (int64_t i = 0; i < n; i++)
        CUDA_TRY(cudaMemcpy(h_obs + i * 4 * OBS, h->d_obs + h_idx[i] * 4 * OBS, 4 * OBS * sizeof(float), cudaMemcpyDeviceToHost));
    return MSOC_OK;
}

int msoc_set_obs_host(msoc_handle *h, const int64_t *h_idx, int64_t n, const float *h_obs)
{
    int rc = check_idx(h, h_idx, n, "msoc_set_obs_host: bad argument");
    if (rc != MSOC_OK) return rc;
    DeviceGuard guard(h->device);
    CUDA_TRY(cudaDeviceSynchronize());
    for (int64_t i = 0; i < n; i++)
        CUDA_TRY(cudaMemcpy(h->d_obs + h_idx[i] * 4 * OBS, h_obs + i * 4 * OBS, 4 * OBS * sizeof(float), cudaMemcpyHostToDevice));
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
