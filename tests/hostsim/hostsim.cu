/*
 * hostsim.cu -- TEST INFRASTRUCTURE ONLY.  Compiles the per-env fp32 arithmetic of the product
 * kernels (marl_soccer_b200/csrc/step_core.cuh, all __host__ __device__) for the HOST, so that the
 * fp32 step logic can be checked against the fp64 oracle on a machine without a GPU
 * (`pytest -m "not gpu"`).  The product package never loads this library and has no CPU path; the
 * GPU parity tests (`pytest -m gpu`) check the real kernels through the C-ABI of include/msoc.h.
 *
 * Build: nvcc -O2 -std=c++17 --shared -Xcompiler -fPIC -o libhostsim.so hostsim.cu   (host code only)
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../include/msoc.h"
#include "../../marl_soccer_b200/csrc/step_core.cuh"

using namespace msoc;

struct HostSim {
    int64_t n;
    uint64_t global_offset;
    SimCfg cfg;
    Arrays A;
    int step; /* 0..5, the kernels' device-side step counter */
    std::vector<std::vector<char>> mem;
    double stats[8];
    std::vector<int32_t> last_contacts, last_load;
};

template <typename T>
static T *alloc(HostSim *h, size_t count)
{
    h->mem.emplace_back(count * sizeof(T), 0);
    return reinterpret_cast<T *>(h->mem.back().data());
}

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

extern "C" {

void hsim_reset(HostSim *h, const uint8_t *mask, int mode, int has_seed, uint64_t seed, float *obs_out);

HostSim *hsim_create(const msoc_config *cfg, int64_t n, uint64_t seed, uint64_t global_offset, float *obs_out)
{
    HostSim *h = new HostSim();
    h->n = n; h->global_offset = global_offset; h->step = 0;
    memset(h->stats, 0, sizeof h->stats);
    fill_cfg(cfg, h->cfg);
    Arrays &A = h->A;
    A.n = n;
    const size_t N = (size_t)n;
    for (int k = 0; k < 3; k++) A.pose[k] = alloc<float4>(h, N * POSE_F4);
    A.score = alloc<int2>(h, N);
    A.inject = alloc<float4>(h, N * POSE_F4);
    A.bias = alloc<float4>(h, N * 4);
    A.seed = alloc<uint64_t>(h, N); A.spawn_count = alloc<uint32_t>(h, N);
    for (int k = 0; k < 2; k++) {
        A.cache[k] = alloc<uint32_t>(h, N * MAX_CACHE * 3);
    }
    for (int64_t e = 0; e < n; e++) { A.seed[e] = seed; A.spawn_count[e] = 0; }
    hsim_reset(h, nullptr, MSOC_MODE_RANDOM, 0, 0, obs_out);
    return h;
}

void hsim_destroy(HostSim *h) { delete h; }

/* stacked observation of one env from its three records (the kernels' obs_tile), soccer_env.py:130-140 */
static void write_obs(const HostSim *h, int64_t e, float *obs_out)
{
    const float4 *recs[3] = {h->A.pose[buf_prev(h->step)] + e * POSE_F4, h->A.pose[buf_cur(h->step)] + e * POSE_F4,
                             h->A.pose[buf_next(h->step)] + e * POSE_F4};
    for (int a = 0; a < 4; a++)
        for (int k = 0; k < 3; k++) frame_of_record(recs[k], a, h->cfg, obs_out + a * OBS + k * FRAME);
}

void hsim_reset(HostSim *h, const uint8_t *mask, int mode, int has_seed, uint64_t seed, float *obs_out)
{
    for (int64_t e = 0; e < h->n; e++) {
        if (mask && !mask[e]) continue;
        const uint64_t gidx = h->global_offset + (uint64_t)e;
        uint64_t sd; uint32_t sc;
        if (has_seed) { sd = seed + gidx; sc = 0; h->A.seed[e] = sd; }
        else { sd = h->A.seed[e]; sc = h->A.spawn_count[e]; }
        Env E;
        env_full_reset(E, mode, sd, gidx, sc);
        h->A.spawn_count[e] = sc;
        store_env(h->A, buf_cur(h->step), e, E, true);
        store_env_record(h->A.pose[buf_prev(h->step)] + e * POSE_F4, E);
        store_env_record(h->A.pose[buf_next(h->step)] + e * POSE_F4, E);
        if (obs_out) write_obs(h, e, obs_out + e * 4 * OBS);
    }
}

void hsim_step(HostSim *h, const float *actions, float *obs_out, float *reward, uint8_t *done,
               int8_t *goal, int32_t *score, uint32_t flags)
{
    h->last_contacts.assign((size_t)h->n, 0); h->last_load.assign((size_t)h->n, -1);
    const int step = h->step;
    for (int64_t e = 0; e < h->n; e++) {
        const float4 *rec = h->A.pose[buf_cur(step)] + e * POSE_F4;
        const bool injected = (f2u(rec[7].z) & FLAG_INJECT) != 0u;
        Env E;
        load_env(h->A, rec, e, E);
        StepOut out;
        /* same flow as the kernels: contact-free fast pass first, the light or the general pass if it declines */
        int load = 0;
        float body[BODY_FIELDS * 5], con[CON_FIELDS * CON_FAST], geom[GEOM_WORDS], oldc[3 * OLD_FAST], isl[ISL_FIELDS * ISL_SLOTS];
        float ovf_store[MAXC - CON_FAST][CON_FIELDS];
        Work W;
        W.ovf = ovf_store;
        int pool_count = 0;
        W.body = body; W.pool = con; W.pool_count = &pool_count; W.geom = geom; W.old = oldc; W.isl = isl;
        const uint64_t gidx = h->global_offset + (uint64_t)e;
        if (!env_step(MODE_FAST, 1 << MODE_FAST, E, actions + e * 12, h->cfg, h->A, cache_half(step), e, gidx, flags, W, out, load)) {
            load_env(h->A, (injected && load != 0) ? h->A.inject + e * POSE_F4 : rec, e, E);
            h->last_load[(size_t)e] = load;
            int dummy;
            pool_count = 0;
            const bool force_full = getenv("HSIM_FORCE_FULL") != nullptr; /* debugging: everything through the general path */
            env_step(force_full ? MODE_FULL : mode_of_load(load), 31, E, actions + e * 12, h->cfg, h->A, cache_half(step), e, gidx, flags, W, out, dummy);
        }
        h->last_contacts[(size_t)e] = out.n_contacts;
        store_env(h->A, buf_next(step), e, E, out.score_dirty);
        if (out.fresh_episode) { /* a fresh episode's history is its first frame */
            store_env_record(h->A.pose[buf_cur(step)] + e * POSE_F4, E);
            store_env_record(h->A.pose[buf_prev(step)] + e * POSE_F4, E);
        }
        reward[2 * e] = out.reward; reward[2 * e + 1] = out.reward;
        done[e] = out.done; goal[e] = out.goal;
        if (score) { score[2 * e] = out.score_b; score[2 * e + 1] = out.score_r; }
        write_obs(h, e, obs_out + e * 4 * OBS);
        if (out.done) { h->stats[0] += 1.0; h->stats[1] += out.finished_return; }
        if (out.goal > 0) h->stats[2] += 1.0;
        if (out.goal < 0) h->stats[3] += 1.0;
        h->stats[4] += 1.0; h->stats[5] += out.n_contacts; h->stats[6] += out.overflow;
    }
    h->step = (h->step + 1) % 6;
}

void hsim_last(HostSim *h, int32_t *contacts, int32_t *load)
{
    memcpy(contacts, h->last_contacts.data(), (size_t)h->n * 4);
    memcpy(load, h->last_load.data(), (size_t)h->n * 4);
}

void hsim_stats(HostSim *h, double *out8, int reset)
{
    memcpy(out8, h->stats, sizeof h->stats);
    if (reset) memset(h->stats, 0, sizeof h->stats);
}

static void hist_to_record(const msoc_env_state *S, int k, float4 *r)
{
    Pose Q;
    for (int i = 0; i < 5; i++) { Q.px[i] = S->hist_pos[k][i][0]; Q.py[i] = S->hist_pos[k][i][1]; }
    for (int i = 0; i < 4; i++) {
        double a = (double)S->hist_ang[k][i];
        if (a > 3.14159274101257324 || a < -3.14159274101257324) a = atan2(sin(a), cos(a));
        Q.vx[i] = S->hist_vel[k][i][0]; Q.vy[i] = S->hist_vel[k][i][1]; Q.ang[i] = (float)a; Q.w[i] = S->hist_angvel[k][i];
    }
    pose_pack(Q, 0.0f, 0.0f, r);
}
static void record_to_hist(const float4 *r, msoc_env_state *S, int k)
{
    Pose Q; pose_unpack(r, Q);
    for (int i = 0; i < 5; i++) { S->hist_pos[k][i][0] = Q.px[i]; S->hist_pos[k][i][1] = Q.py[i]; }
    for (int i = 0; i < 4; i++) { S->hist_vel[k][i][0] = Q.vx[i]; S->hist_vel[k][i][1] = Q.vy[i]; S->hist_ang[k][i] = Q.ang[i]; S->hist_angvel[k][i] = Q.w[i]; }
}
static bool same_pose(const float4 *a, const float4 *b)
{
    bool same = true;
    for (int i = 0; i < 7; i++) {
        same = same && f2u(a[i].x) == f2u(b[i].x) && f2u(a[i].y) == f2u(b[i].y);
        if (i != 4) same = same && f2u(a[i].z) == f2u(b[i].z) && f2u(a[i].w) == f2u(b[i].w);
    }
    return same;
}

void hsim_get_state(HostSim *h, int64_t e, msoc_env_state *S)
{
    const Arrays &A = h->A;
    const float4 *rec = A.pose[buf_cur(h->step)] + e * POSE_F4;
    const bool injected = (f2u(rec[7].z) & FLAG_INJECT) != 0u;
    Env E;
    load_env(A, injected ? A.inject + e * POSE_F4 : rec, e, E);
    memset(S, 0, sizeof *S);
    for (int i = 0; i < 5; i++) {
        S->pos[i][0] = E.px[i]; S->pos[i][1] = E.py[i]; S->vel[i][0] = E.vx[i]; S->vel[i][1] = E.vy[i];
        S->angvel[i] = E.w[i]; S->vbias[i][0] = E.vbx[i]; S->vbias[i][1] = E.vby[i];
    }
    for (int i = 0; i < 4; i++) { S->ang[i] = E.ang[i]; S->wbias[i] = E.wb[i]; }
    S->ep_return = E.ep_return; S->steps = E.steps; S->score[0] = E.score_b; S->score[1] = E.score_r;
    S->mode = (int)((E.flags & FLAG_MODE_MASK) >> FLAG_MODE_SHIFT);
    S->spawn_count = A.spawn_count[e]; S->seed = A.seed[e];
    const uint32_t cnt = E.flags & FLAG_CACHE_MASK;
    S->cache_count = cnt;
    for (uint32_t j = 0; j < cnt; j++) {
        const uint32_t *c = A.cache[cache_half(h->step)] + cache_slot(e, (int)j);
        S->cache_info[j] = c[0]; S->cache_jn[j] = u2f(c[1]); S->cache_jt[j] = u2f(c[2]);
    }
    S->hist_valid = 1u;
    record_to_hist(A.pose[buf_prev(h->step)] + e * POSE_F4, S, 0);
    record_to_hist(rec, S, 1);
}

/* mirrors msoc_set_state_kernel */
void hsim_set_state(HostSim *h, int64_t e, const msoc_env_state *S)
{
    const Arrays &A = h->A;
    const int step = h->step;
    Env E;
    for (int i = 0; i < 5; i++) {
        E.px[i] = S->pos[i][0]; E.py[i] = S->pos[i][1]; E.vx[i] = S->vel[i][0]; E.vy[i] = S->vel[i][1];
        E.w[i] = S->angvel[i]; E.vbx[i] = S->vbias[i][0]; E.vby[i] = S->vbias[i][1];
    }
    for (int i = 0; i < 4; i++) {
        double a = (double)S->ang[i];
        if (a > 3.14159274101257324 || a < -3.14159274101257324) a = atan2(sin(a), cos(a));
        E.ang[i] = (float)a; E.wb[i] = S->wbias[i];
    }
    E.ep_return = S->ep_return; E.steps = S->steps; E.score_b = S->score[0]; E.score_r = S->score[1];
    uint32_t cnt = S->cache_count > (uint32_t)MAX_CACHE ? (uint32_t)MAX_CACHE : S->cache_count;
    E.flags = cnt | (((uint32_t)S->mode & 3u) << FLAG_MODE_SHIFT);
    A.spawn_count[e] = S->spawn_count; A.seed[e] = S->seed;
    for (uint32_t j = 0; j < cnt; j++) {
        uint32_t *c = A.cache[cache_half(step)] + cache_slot(e, (int)j);
        c[0] = S->cache_info[j]; c[1] = f2u(S->cache_jn[j]); c[2] = f2u(S->cache_jt[j]);
    }
    float4 *rec = A.pose[buf_cur(step)] + e * POSE_F4;
    float4 newp[7], prev[7];
    { Pose Q; pose_of(E, Q); pose_pack(Q, 0.0f, 0.0f, newp); }
    if (S->hist_valid) {
        hist_to_record(S, 0, A.pose[buf_prev(step)] + e * POSE_F4);
        hist_to_record(S, 1, prev);
    } else {
        memcpy(prev, rec, sizeof prev);
    }
    store_env_bias(A, e, E);
    A.score[e] = make_int2(E.score_b, E.score_r);
    if (same_pose(prev, newp)) {
        store_env_record(rec, E);
    } else {
        store_env_record(A.inject + e * POSE_F4, E);
        memcpy(rec, prev, sizeof prev);
        rec[7] = make_float4(0.0f, u2f((uint32_t)E.steps), u2f(FLAG_INJECT), 0.0f);
    }
}

} /* extern "C" */
