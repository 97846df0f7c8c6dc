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
    int cur;
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
}

extern "C" {

void hsim_reset(HostSim *h, const uint8_t *mask, int mode, int has_seed, uint64_t seed, float *obs_out);

HostSim *hsim_create(const msoc_config *cfg, int64_t n, uint64_t seed, uint64_t global_offset, float *obs_out)
{
    HostSim *h = new HostSim();
    h->n = n; h->global_offset = global_offset; h->cur = 0;
    memset(h->stats, 0, sizeof h->stats);
    fill_cfg(cfg, h->cfg);
    Arrays &A = h->A;
    A.n = n;
    const size_t N = (size_t)n;
    A.bodies = alloc<float4>(h, N * 5); A.misc = alloc<float4>(h, N * 4);
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

static void write_obs(const float *frames /*4x22*/, bool fresh, const float *obs_in, float *obs_out)
{
    for (int a = 0; a < 4; a++) {
        float row[OBS];
        if (fresh) {
            for (int f = 0; f < 3; f++) memcpy(row + f * FRAME, frames + a * FRAME, FRAME * sizeof(float));
        } else {
            memcpy(row, obs_in + a * OBS + FRAME, 2 * FRAME * sizeof(float));
            memcpy(row + 2 * FRAME, frames + a * FRAME, FRAME * sizeof(float));
        }
        memcpy(obs_out + a * OBS, row, sizeof row);
    }
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
        store_env(h->A, e, E);
        if (obs_out) {
            float frames[4 * FRAME];
            make_frames<FRAME>(E, h->cfg, frames);
            write_obs(frames, true, nullptr, obs_out + e * 4 * OBS);
        }
    }
}

void hsim_step(HostSim *h, const float *actions, const float *obs_in, float *obs_out, float *reward, uint8_t *done,
               int8_t *goal, int32_t *score, uint32_t flags)
{
    h->last_contacts.assign((size_t)h->n, 0); h->last_load.assign((size_t)h->n, -1);
    for (int64_t e = 0; e < h->n; e++) {
        Env E;
        load_env(h->A, e, E);
        StepOut out;
        float frames[4 * FRAME];
        /* same two-instantiation flow as the kernel: contact-free fast pass first, full pass if it declines */
        int load = 0;
        float body[BODY_FIELDS * 5], con[CON_FIELDS * CON_FAST], geom[GEOM_WORDS], oldc[3 * OLD_FAST];
        float ovf_store[MAXC - CON_FAST][CON_FIELDS];
        Work W;
        W.ovf = ovf_store;
        int pool_count = 0;
        W.body = body; W.pool = con; W.pool_count = &pool_count; W.geom = geom; W.old = oldc;
        const uint64_t gidx = h->global_offset + (uint64_t)e;
        if (!env_step(MODE_FAST, E, actions + e * 12, h->cfg, h->A, h->cur, e, gidx, flags, W, out, load)) {
            load_env(h->A, e, E);
            h->last_load[(size_t)e] = load;
            /* light class (one agent x wall pair): the register path; anything else: the general path */
            int dummy;
            pool_count = 0;
            env_step(load == 0 ? MODE_LIGHT : MODE_FULL, E, actions + e * 12, h->cfg, h->A, h->cur, e, gidx, flags, W, out, dummy);
        }
        h->last_contacts[(size_t)e] = out.n_contacts;
        float snap[SNAP_FIELDS];
        snapshot_env(E, snap, 1);
        for (int a = 0; a < 4; a++) make_frame_dyn(snap, 1, a, h->cfg, frames + a * FRAME);
        store_env(h->A, e, E);
        reward[2 * e] = out.reward; reward[2 * e + 1] = out.reward;
        done[e] = out.done; goal[e] = out.goal;
        if (score) { score[2 * e] = out.score_b; score[2 * e + 1] = out.score_r; }
        write_obs(frames, out.fresh_episode, obs_in + e * 4 * OBS, obs_out + e * 4 * OBS);
        if (out.done) { h->stats[0] += 1.0; h->stats[1] += out.finished_return; }
        if (out.goal > 0) h->stats[2] += 1.0;
        if (out.goal < 0) h->stats[3] += 1.0;
        h->stats[4] += 1.0; h->stats[5] += out.n_contacts; h->stats[6] += out.overflow;
    }
    h->cur ^= 1;
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

void hsim_get_state(HostSim *h, int64_t e, msoc_env_state *S)
{
    const Arrays &A = h->A;
    Env E;
    load_env(A, e, E);
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
        const uint32_t *c = A.cache[h->cur] + cache_slot(e, (int)j);
        S->cache_info[j] = c[0]; S->cache_jn[j] = u2f(c[1]); S->cache_jt[j] = u2f(c[2]);
    }
}

void hsim_set_state(HostSim *h, int64_t e, const msoc_env_state *S)
{
    const Arrays &A = h->A;
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
        uint32_t *c = A.cache[h->cur] + cache_slot(e, (int)j);
        c[0] = S->cache_info[j]; c[1] = f2u(S->cache_jn[j]); c[2] = f2u(S->cache_jt[j]);
    }
    store_env(A, e, E);
}

} /* extern "C" */
