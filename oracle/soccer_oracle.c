/*
 * soccer_oracle.c -- CPU ORACLE (test infrastructure, NOT a product path).
 *
 * A plain-C, fp64, scalar restatement of the reference hot path
 *     SoccerEnv.step -> Game.step -> pymunk Space.step(1/60)
 * for the 2v2 soccer scene of sdace9719/marl-soccer.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library; the product (marl_soccer_b200/) never does and has no
 * CPU fallback.
 *
 * PARITY UNPINNED.  The arithmetic of the reference lives in the third-party
 * dependency pymunk (cffi wrapper of Chipmunk2D), which is unpinned
 * (soccer_simulation/requirements.txt:2), not vendored in /root/reference and
 * not installable here (no network).  The reference's own tests hold no golden
 * vectors for this path (test_rewards.py asserts reward SIGNS only).  What is
 * restated below from Chipmunk2D 7.0.3 (cpSpaceStep.c, cpBody.c, cpArbiter.c,
 * cpCollision.c, cpPolyShape.c, cpSpace.c defaults) is its published
 * algorithm, anchored on the reference's call sites:
 *     game/game.py:24-25   Space(), gravity (0,0)
 *     game/game.py:50-72   6 wall segments r=2, 2 goal-line segments r=1
 *     game/entities.py:11-35   agent: Body(mass,100), box 30x30, e .2 u .8
 *     game/entities.py:62-84   ball:  Body(mass,10), circle r=10, e .95 u .2
 *     game/game.py:378-437 Game.step
 *     game/game.py:76-249  Game.reset and the three spawn modes
 *     game/game.py:258-322 22-float observation frame
 *     game/game.py:324-375 rewards
 *     soccer_env.py:100-154 action clip/scale (float32), 3-frame stacking
 *     marl_vecenv.py:30-68 auto-reset in full-random mode
 * What IS pinned (tests/test_oracle_*.py): the observation layout constants
 * of test_rewards.py:37-58, its seven behavioural scenarios, Philox4x32-10
 * known-answer vectors, and analytic contact-free / single-contact cases.
 *
 * Two deliberate, documented differences from the reference:
 *  (1) Spawn randomness is a counter-based Philox4x32-10 stream keyed by the
 *      per-env seed and global env index (NumPy's PCG64 stream cannot be
 *      matched by a device RNG).  The draw ORDER and
 *      distributions follow game/game.py:154-249.  Draws are float32 so that
 *      the device kernel reproduces spawn states bit-exactly.
 *  (2) Arbiter order inside one step (implementation-defined in Chipmunk's
 *      BBTree) is canonical: ascending pair id = agent x segment (agent-major,
 *      setup_field order), agent x agent, ball x agent, ball x wall.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "soccer_oracle.h"

/* ------------------------------------------------------------------ vec2 */
typedef struct { double x, y; } v2;
static inline v2 V(double x, double y) { v2 r = {x, y}; return r; }
static inline v2 vadd(v2 a, v2 b) { return V(a.x + b.x, a.y + b.y); }
static inline v2 vsub(v2 a, v2 b) { return V(a.x - b.x, a.y - b.y); }
static inline v2 vmul(v2 a, double s) { return V(a.x * s, a.y * s); }
static inline v2 vneg(v2 a) { return V(-a.x, -a.y); }
static inline double vdot(v2 a, v2 b) { return a.x * b.x + a.y * b.y; }
static inline double vcross(v2 a, v2 b) { return a.x * b.y - a.y * b.x; }
static inline v2 vperp(v2 a) { return V(-a.y, a.x); }   /* cpvperp  */
static inline v2 vrperp(v2 a) { return V(a.y, -a.x); }  /* cpvrperp */
static inline double vlen(v2 a) { return sqrt(vdot(a, a)); }
static inline v2 vlerp(v2 a, v2 b, double t) { return vadd(vmul(a, 1.0 - t), vmul(b, t)); }
static inline v2 vrotate(v2 a, v2 b) { return V(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
static inline v2 vnormalize(v2 a) { return vmul(a, 1.0 / (vlen(a) + DBL_MIN)); }
static inline double clamp01(double t) { return t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t); }
static inline double dclamp(double f, double lo, double hi) { return fmin(fmax(f, lo), hi); }

/* ------------------------------------------------------- scene constants */
/* game/constants.py:2-19 */
#define SCREEN_W 800.0
#define SCREEN_H 600.0
#define FIELD_MARGIN 10.0
#define GOAL_HEIGHT 150.0
#define AGENT_SIZE 30.0
#define BALL_RADIUS 10.0
#define CAT_BALL 1u
#define CAT_AGENT 2u
#define CAT_WALL 4u
#define CAT_GOALWALL 8u

#define N_AGENTS 4
#define BALL 4
#define STATIC_BODY 5
#define N_BODIES 6
#define N_SEGS 8

#define ARB_FIRST 0
#define ARB_NORMAL 1
#define ARB_CACHED 3

typedef struct {
    v2 p, v;
    double a, w;
    v2 v_bias;
    double w_bias;
    v2 f;
    double t;
    double m_inv, i_inv;
} Body;

typedef struct { v2 a, b, n; double r; double e, u; unsigned cat, mask; } Segment;

typedef struct {
    v2 r1, r2;
    double nMass, tMass, bounce, jnAcc, jtAcc, jBias, bias;
    int key; /* canonical feature key standing in for Chipmunk's contact hash */
} Contact;

typedef struct {
    int exists, state;
    long stamp;
    int body_a, body_b;
    int count;
    Contact con[2];
    v2 n;
    double e, u;
} Arbiter;

struct OracleEnv {
    OracleConfig cfg;
    Body body[N_BODIES];
    Segment seg[N_SEGS];
    v2 box_local[4];
    Arbiter arb[ORACLE_N_PAIRS];
    int active[ORACLE_N_PAIRS];
    int n_active;
    long stamp;
    double prev_dt;
    /* game */
    int steps, score_blue, score_red;
    int mode;
    uint64_t seed, global_index;
    uint32_t spawn_count;
    double prev_d[4], prev_D_blue, prev_D_red;
    /* env (soccer_env.py:37-39) */
    float frames[4][3][ORACLE_FRAME];
};

/* --------------------------------------------------------------- Philox */
static inline void mulhilo(uint32_t a, uint32_t b, uint32_t *hi, uint32_t *lo)
{
    uint64_t p = (uint64_t)a * (uint64_t)b;
    *hi = (uint32_t)(p >> 32);
    *lo = (uint32_t)p;
}

void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint32_t hi0, lo0, hi1, lo1;
        mulhilo(0xD2511F53u, c0, &hi0, &lo0);
        mulhilo(0xCD9E8D57u, c2, &hi1, &lo1);
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* float32 uniform in [lo, hi): 24 random bits, one fused multiply-add. */
static inline float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
static inline float uni(uint32_t x, float lo, float hi) { return fmaf(u01(x), hi - lo, lo); }

/* ----------------------------------------------------------- scene setup */
static void setup_field(OracleEnv *E)
{
    /* game/game.py:45-72 */
    const double gy_top = SCREEN_H / 2 + GOAL_HEIGHT / 2, gy_bot = SCREEN_H / 2 - GOAL_HEIGHT / 2;
    const double L = FIELD_MARGIN, R = SCREEN_W - FIELD_MARGIN, B = FIELD_MARGIN, T = SCREEN_H - FIELD_MARGIN;
    const v2 pts[N_SEGS][2] = {
        {{L, B}, {R, B}}, {{L, T}, {R, T}},
        {{L, B}, {L, gy_bot}}, {{L, gy_top}, {L, T}},
        {{R, B}, {R, gy_bot}}, {{R, gy_top}, {R, T}},
        {{L, gy_bot}, {L, gy_top}}, {{R, gy_bot}, {R, gy_top}},
    };
    for (int s = 0; s < N_SEGS; s++) {
        Segment *g = &E->seg[s];
        g->a = pts[s][0]; g->b = pts[s][1];
        g->n = vrperp(vnormalize(vsub(g->b, g->a))); /* cpSegmentShapeInit */
        if (s < 6) { g->r = 2.0; g->e = 0.95; g->u = 0.2; g->cat = CAT_WALL; g->mask = CAT_AGENT | CAT_BALL; }
        else       { g->r = 1.0; g->e = 0.95; g->u = 0.0; g->cat = CAT_GOALWALL; g->mask = CAT_AGENT; }
    }
    /* cpBoxShapeInit2 vertex order */
    const double h = AGENT_SIZE / 2;
    E->box_local[0] = V(h, -h); E->box_local[1] = V(h, h);
    E->box_local[2] = V(-h, h); E->box_local[3] = V(-h, -h);
    Body *S = &E->body[STATIC_BODY];
    memset(S, 0, sizeof *S);
}

static void new_bodies(OracleEnv *E)
{
    /* game/game.py:88-106: bodies and shapes are re-created; Chipmunk drops the
       cached arbiters of removed shapes (cpSpaceRemoveShape -> cpSpaceFilterArbiters). */
    for (int i = 0; i < 5; i++) {
        Body *b = &E->body[i];
        memset(b, 0, sizeof *b);
        if (i < N_AGENTS) { b->m_inv = 1.0 / E->cfg.agent_mass; b->i_inv = 1.0 / E->cfg.agent_moment; }
        else              { b->m_inv = 1.0 / E->cfg.ball_mass;  b->i_inv = 1.0 / E->cfg.ball_moment; }
    }
    E->body[2].a = M_PI; E->body[3].a = M_PI;
    memset(E->arb, 0, sizeof E->arb);
    E->n_active = 0;
}

/* game/game.py:129-249.  Positions only; velocities/angles per the reference. */
static void apply_spawn(OracleEnv *E)
{
    double px[5], py[5];
    if (E->mode == ORACLE_MODE_FIXED) {
        px[0] = SCREEN_W * 0.25; py[0] = SCREEN_H * 0.33;
        px[1] = SCREEN_W * 0.25; py[1] = SCREEN_H * 0.66;
        px[2] = SCREEN_W * 0.75; py[2] = SCREEN_H * 0.33;
        px[3] = SCREEN_W * 0.75; py[3] = SCREEN_H * 0.66;
        px[4] = SCREEN_W / 2;    py[4] = SCREEN_H / 2;
    } else {
        uint32_t key[2] = {(uint32_t)E->seed, (uint32_t)(E->seed >> 32)};
        uint32_t ctr[4] = {(uint32_t)E->global_index, (uint32_t)(E->global_index >> 32), E->spawn_count, 0};
        uint32_t A[4], Bq[4], C[4], D[4];
        ctr[3] = 0; oracle_philox4x32_10(ctr, key, A);
        ctr[3] = 1; oracle_philox4x32_10(ctr, key, Bq);
        ctr[3] = 2; oracle_philox4x32_10(ctr, key, C);
        ctr[3] = 3; oracle_philox4x32_10(ctr, key, D);
        E->spawn_count++;
        const float xmin = 30.0f, xmax = 770.0f, ymin = 30.0f, ymax = 570.0f;
        if (E->mode == ORACLE_MODE_FULL_RANDOM) {
            int corners = u01(A[0]) < 0.75f;
            if (corners) {
                int c[2] = {(int)(A[1] >> 30), (int)(A[2] >> 30)};
                for (int k = 0; k < 2; k++) {
                    int left = (c[k] == 0 || c[k] == 1), top = (c[k] == 0 || c[k] == 2);
                    float cx = left ? 18.0f : 782.0f, cy = top ? 582.0f : 18.0f;
                    px[k] = cx + uni(Bq[2 * k], -5.0f, 5.0f);
                    py[k] = cy + uni(Bq[2 * k + 1], -5.0f, 5.0f);
                }
            } else {
                for (int k = 0; k < 2; k++) { px[k] = uni(Bq[2 * k], xmin, xmax); py[k] = uni(Bq[2 * k + 1], ymin, ymax); }
            }
            for (int k = 0; k < 2; k++) { px[2 + k] = uni(C[2 * k], xmin, xmax); py[2 + k] = uni(C[2 * k + 1], ymin, ymax); }
            px[4] = uni(D[0], xmin, xmax); py[4] = uni(D[1], ymin, ymax);
        } else {
            for (int k = 0; k < 2; k++) { px[k] = uni(Bq[2 * k], 30.0f, 380.0f); py[k] = uni(Bq[2 * k + 1], ymin, ymax); }
            for (int k = 0; k < 2; k++) { px[2 + k] = uni(C[2 * k], 420.0f, 770.0f); py[2 + k] = uni(C[2 * k + 1], ymin, ymax); }
            px[4] = 400.0f + uni(D[0], -40.0f, 40.0f); py[4] = 300.0f + uni(D[1], -40.0f, 40.0f);
        }
        /* float32 sums so that the device reproduces them exactly */
        for (int k = 0; k < 5; k++) { px[k] = (double)(float)px[k]; py[k] = (double)(float)py[k]; }
    }
    for (int i = 0; i < 5; i++) {
        Body *b = &E->body[i];
        b->p = V(px[i], py[i]);
        b->v = V(0, 0);
        if (i < N_AGENTS) { b->w = 0.0; b->a = (i < 2) ? 0.0 : M_PI; }
        /* ball: angular velocity is NOT reset (game/game.py:151-152, 189-190) */
    }
}

static void update_reward_state(OracleEnv *E)
{
    /* game/game.py:251-256 */
    v2 bp = E->body[BALL].p;
    for (int i = 0; i < 4; i++) E->prev_d[i] = vlen(vsub(E->body[i].p, bp));
    E->prev_D_blue = vlen(vsub(bp, V(FIELD_MARGIN, SCREEN_H / 2)));
    E->prev_D_red = vlen(vsub(bp, V(SCREEN_W - FIELD_MARGIN, SCREEN_H / 2)));
}

/* game/game.py:258-322 with the dtype flow of NumPy:
   vel is cast to f32 and divided in f32; everything else is f64 then cast. */
static void frame_for_agent(const OracleEnv *E, int i, float *o)
{
    const OracleConfig *c = &E->cfg;
    const Body *me = &E->body[i];
    double vmax = fmax(c->max_velocity, 1e-6);
    o[0] = (float)me->v.x / (float)vmax;
    o[1] = (float)me->v.y / (float)vmax;
    double aw = atan2(sin(me->a), cos(me->a));
    o[2] = (float)(aw / M_PI);
    o[3] = (float)(me->w / fmax(c->max_angular_velocity, 1e-6));
    int mate = (i == 0) ? 1 : (i == 1) ? 0 : (i == 2) ? 3 : 2;
    int opp0 = (i < 2) ? 2 : 0, opp1 = (i < 2) ? 3 : 1;
    v2 blue_goal = V(FIELD_MARGIN, SCREEN_H / 2), red_goal = V(SCREEN_W - FIELD_MARGIN, SCREEN_H / 2);
    v2 targets[6] = {E->body[mate].p, E->body[opp0].p, E->body[opp1].p, E->body[BALL].p,
                     (i < 2) ? blue_goal : red_goal, (i < 2) ? red_goal : blue_goal};
    double diag = hypot(SCREEN_W, SCREEN_H);
    for (int k = 0; k < 6; k++) {
        v2 d = vsub(targets[k], me->p);
        double mag = vlen(d);
        float ux = 0.0f, uy = 0.0f;
        if (mag > 1e-8) { ux = (float)(d.x / mag); uy = (float)(d.y / mag); } else mag = 0.0;
        o[4 + 3 * k] = ux; o[5 + 3 * k] = uy;
        o[6 + 3 * k] = (float)(mag / fmax(diag, 1e-6));
    }
}

/* ------------------------------------------------------- narrow phase */
static void box_world(const OracleEnv *E, int i, v2 *verts, v2 *normals)
{
    /* cpPolyShapeCacheData: v0 = T*v, n = R*n; plane k = (vert k, normal of edge k-1 -> k) */
    const Body *b = &E->body[i];
    v2 rot = V(cos(b->a), sin(b->a));
    for (int k = 0; k < 4; k++) verts[k] = vadd(b->p, vrotate(E->box_local[k], rot));
    const v2 ln[4] = {{0, -1}, {1, 0}, {0, 1}, {-1, 0}};
    for (int k = 0; k < 4; k++) normals[k] = vrotate(ln[k], rot);
}

/* Closest features of two convex polygons (a segment is a 2-gon).  Returns the
   (n, d) that Chipmunk's GJK (separated) / EPA (overlapping) converge to:
   n points from A to B, d is the signed distance. */
static void closest_convex(const v2 *A, int nA, const v2 *B, int nB, v2 *n_out, double *d_out)
{
    double best = -INFINITY; v2 bn = V(0, 0);
    for (int k = 0; k < nA; k++) {
        v2 e = vsub(A[(k + 1) % nA], A[k]);
        v2 nk = vnormalize(vrperp(e));
        double s = INFINITY;
        for (int j = 0; j < nB; j++) s = fmin(s, vdot(vsub(B[j], A[k]), nk));
        if (s > best) { best = s; bn = nk; }
    }
    for (int k = 0; k < nB; k++) {
        v2 e = vsub(B[(k + 1) % nB], B[k]);
        v2 nk = vnormalize(vrperp(e));
        double s = INFINITY;
        for (int i = 0; i < nA; i++) s = fmin(s, vdot(vsub(A[i], B[k]), nk));
        if (s > best) { best = s; bn = vneg(nk); }
    }
    if (best <= 0.0) { *n_out = bn; *d_out = best; return; }
    double bd = INFINITY; v2 bdn = bn;
    for (int k = 0; k < nA; k++) {
        v2 e = vsub(A[(k + 1) % nA], A[k]);
        for (int j = 0; j < nB; j++) {
            double t = clamp01(vdot(vsub(B[j], A[k]), e) / vdot(e, e));
            v2 delta = vsub(B[j], vadd(A[k], vmul(e, t)));
            double dist = vlen(delta);
            if (dist < bd) { bd = dist; bdn = vmul(delta, 1.0 / dist); }
        }
    }
    for (int k = 0; k < nB; k++) {
        v2 e = vsub(B[(k + 1) % nB], B[k]);
        for (int i = 0; i < nA; i++) {
            double t = clamp01(vdot(vsub(A[i], B[k]), e) / vdot(e, e));
            v2 delta = vsub(vadd(B[k], vmul(e, t)), A[i]);
            double dist = vlen(delta);
            if (dist < bd) { bd = dist; bdn = vmul(delta, 1.0 / dist); }
        }
    }
    *n_out = bdn; *d_out = bd;
}

typedef struct { v2 pa, pb; int ia, ib; double r; } Edge; /* endpoints + vertex indices */

static Edge support_edge_poly(const v2 *verts, const v2 *normals, v2 n)
{
    /* cpCollision.c SupportEdgeForPoly / PolySupportPointIndex */
    double mx = -INFINITY; int i1 = 0;
    for (int i = 0; i < 4; i++) { double d = vdot(verts[i], n); if (d > mx) { mx = d; i1 = i; } }
    int i0 = (i1 + 3) % 4, i2 = (i1 + 1) % 4;
    Edge e; e.r = 0.0;
    if (vdot(n, normals[i1]) > vdot(n, normals[i2])) { e.pa = verts[i0]; e.ia = i0; e.pb = verts[i1]; e.ib = i1; }
    else                                              { e.pa = verts[i1]; e.ia = i1; e.pb = verts[i2]; e.ib = i2; }
    return e;
}

static Edge support_edge_segment(const Segment *s, v2 n)
{
    Edge e; e.r = s->r;
    if (vdot(s->n, n) > 0.0) { e.pa = s->a; e.ia = 0; e.pb = s->b; e.ib = 1; }
    else                     { e.pa = s->b; e.ia = 1; e.pb = s->a; e.ib = 0; }
    return e;
}

typedef struct { int count; v2 n; v2 p1[2], p2[2]; int key[2]; } Manifold;

static void contact_points(Edge e1, Edge e2, v2 n, double d, Manifold *m)
{
    /* cpCollision.c ContactPoints.  key = (vertex on shape a)*4 + (vertex on shape b)
       identifies the same feature pair that CP_HASH_PAIR(e1.x.hash, e2.y.hash) does. */
    m->count = 0;
    double mindist = e1.r + e2.r;
    if (!(d <= mindist)) return;
    m->n = n;
    double d_e1_a = vcross(e1.pa, n), d_e1_b = vcross(e1.pb, n);
    double d_e2_a = vcross(e2.pa, n), d_e2_b = vcross(e2.pb, n);
    double e1_denom = 1.0 / (d_e1_b - d_e1_a + DBL_MIN);
    double e2_denom = 1.0 / (d_e2_b - d_e2_a + DBL_MIN);
    {
        v2 p1 = vadd(vmul(n, e1.r), vlerp(e1.pa, e1.pb, clamp01((d_e2_b - d_e1_a) * e1_denom)));
        v2 p2 = vadd(vmul(n, -e2.r), vlerp(e2.pa, e2.pb, clamp01((d_e1_a - d_e2_a) * e2_denom)));
        if (vdot(vsub(p2, p1), n) <= 0.0) {
            m->p1[m->count] = p1; m->p2[m->count] = p2; m->key[m->count] = e1.ia * 4 + e2.ib; m->count++;
        }
    }
    {
        v2 p1 = vadd(vmul(n, e1.r), vlerp(e1.pa, e1.pb, clamp01((d_e2_a - d_e1_a) * e1_denom)));
        v2 p2 = vadd(vmul(n, -e2.r), vlerp(e2.pa, e2.pb, clamp01((d_e1_b - d_e2_a) * e2_denom)));
        if (vdot(vsub(p2, p1), n) <= 0.0) {
            m->p1[m->count] = p1; m->p2[m->count] = p2; m->key[m->count] = e1.ib * 4 + e2.ia; m->count++;
        }
    }
}

static void collide_segment_box(const OracleEnv *E, int s, int agent, Manifold *m)
{
    /* cpCollision.c SegmentToPoly: a = segment, b = poly */
    const Segment *g = &E->seg[s];
    v2 verts[4], normals[4];
    box_world(E, agent, verts, normals);
    v2 sv[2] = {g->a, g->b};
    v2 n; double d;
    closest_convex(sv, 2, verts, 4, &n, &d);
    m->count = 0;
    if (d - g->r <= 0.0)
        contact_points(support_edge_segment(g, n), support_edge_poly(verts, normals, vneg(n)), n, d, m);
}

static void collide_box_box(const OracleEnv *E, int i, int j, Manifold *m)
{
    /* cpCollision.c PolyToPoly */
    v2 va[4], na[4], vb[4], nb[4];
    box_world(E, i, va, na);
    box_world(E, j, vb, nb);
    v2 n; double d;
    closest_convex(va, 4, vb, 4, &n, &d);
    m->count = 0;
    if (d <= 0.0)
        contact_points(support_edge_poly(va, na, n), support_edge_poly(vb, nb, vneg(n)), n, d, m);
}

static void collide_ball_box(const OracleEnv *E, int agent, Manifold *m)
{
    /* cpCollision.c CircleToPoly: a = circle, b = poly; GJK with the circle as its centre */
    v2 verts[4], normals[4];
    box_world(E, agent, verts, normals);
    v2 c = E->body[BALL].p;
    m->count = 0;
    double best = -INFINITY; int bk = 0;
    for (int k = 0; k < 4; k++) {
        double s = vdot(vsub(c, verts[k]), normals[k]);
        if (s > best) { best = s; bk = k; }
    }
    v2 n, pb; double d;
    if (best <= 0.0) {
        /* centre inside the box: EPA -> closest face */
        n = vneg(normals[bk]); d = best;
        pb = vsub(c, vmul(normals[bk], best));
    } else {
        double bd = INFINITY; pb = c; n = V(1, 0);
        for (int k = 0; k < 4; k++) {
            v2 a0 = verts[(k + 3) % 4], e = vsub(verts[k], a0);
            double t = clamp01(vdot(vsub(c, a0), e) / vdot(e, e));
            v2 q = vadd(a0, vmul(e, t));
            double dist = vlen(vsub(q, c));
            if (dist < bd) { bd = dist; pb = q; }
        }
        d = bd; n = vmul(vsub(pb, c), 1.0 / d);
    }
    if (d <= BALL_RADIUS) {
        m->count = 1; m->n = n; m->key[0] = 0;
        m->p1[0] = vadd(c, vmul(n, BALL_RADIUS));
        m->p2[0] = pb;
    }
}

static void collide_ball_segment(const OracleEnv *E, int s, Manifold *m)
{
    /* cpCollision.c CircleToSegment (tangents are zero: no end-cap rejection) */
    const Segment *g = &E->seg[s];
    v2 center = E->body[BALL].p;
    v2 seg_delta = vsub(g->b, g->a);
    double t = clamp01(vdot(seg_delta, vsub(center, g->a)) / vdot(seg_delta, seg_delta));
    v2 closest = vadd(g->a, vmul(seg_delta, t));
    double mindist = BALL_RADIUS + g->r;
    v2 delta = vsub(closest, center);
    double distsq = vdot(delta, delta);
    m->count = 0;
    if (distsq < mindist * mindist) {
        double dist = sqrt(distsq);
        v2 n = dist ? vmul(delta, 1.0 / dist) : g->n;
        m->count = 1; m->n = n; m->key[0] = 0;
        m->p1[0] = vadd(center, vmul(n, BALL_RADIUS));
        m->p2[0] = vadd(closest, vmul(n, -g->r));
    }
}

/* ------------------------------------------------------------- arbiters */
static int pair_agent_seg(int i, int s) { return i * 8 + s; }
static int pair_agent_agent(int i, int j)
{
    static const int idx[4][4] = {{-1, 0, 1, 2}, {-1, -1, 3, 4}, {-1, -1, -1, 5}, {-1, -1, -1, -1}};
    return 32 + idx[i][j];
}
static int pair_ball_agent(int i) { return 38 + i; }
static int pair_ball_wall(int s) { return 42 + s; }

static void arbiter_update(OracleEnv *E, int pair, int body_a, int body_b, double e, double u, const Manifold *m)
{
    /* cpSpaceCollideShapes + cpArbiterUpdate */
    Arbiter *arb = &E->arb[pair];
    if (!arb->exists) { memset(arb, 0, sizeof *arb); arb->exists = 1; arb->state = ARB_FIRST; }
    Contact fresh[2];
    for (int i = 0; i < m->count; i++) {
        Contact *c = &fresh[i];
        memset(c, 0, sizeof *c);
        c->r1 = vsub(m->p1[i], E->body[body_a].p);
        c->r2 = vsub(m->p2[i], E->body[body_b].p);
        c->key = m->key[i];
        for (int j = 0; j < arb->count; j++)
            if (arb->con[j].key == c->key) { c->jnAcc = arb->con[j].jnAcc; c->jtAcc = arb->con[j].jtAcc; }
    }
    for (int i = 0; i < m->count; i++) arb->con[i] = fresh[i];
    arb->count = m->count;
    arb->n = m->n;
    arb->e = e; arb->u = u;
    arb->body_a = body_a; arb->body_b = body_b;
    if (arb->state == ARB_CACHED) arb->state = ARB_FIRST;
    E->active[E->n_active++] = pair;
    arb->stamp = E->stamp;
}

static inline double k_scalar_body(const Body *b, v2 r, v2 n)
{
    double rcn = vcross(r, n);
    return b->m_inv + b->i_inv * rcn * rcn;
}
static inline v2 relative_velocity(const Body *a, const Body *b, v2 r1, v2 r2)
{
    v2 v1 = vadd(a->v, vmul(vperp(r1), a->w));
    v2 v2_ = vadd(b->v, vmul(vperp(r2), b->w));
    return vsub(v2_, v1);
}
static inline void apply_impulse(Body *b, v2 j, v2 r)
{
    b->v = vadd(b->v, vmul(j, b->m_inv));
    b->w += b->i_inv * vcross(r, j);
}
static inline void apply_bias_impulse(Body *b, v2 j, v2 r)
{
    b->v_bias = vadd(b->v_bias, vmul(j, b->m_inv));
    b->w_bias += b->i_inv * vcross(r, j);
}

static void velocity_func(OracleEnv *E, int i, double dt)
{
    /* cpBodyUpdateVelocity (gravity 0, damping 1^dt = 1) then the reference's
       custom_velocity_func: game/entities.py:19-28 (agent), :69-77 (ball). */
    Body *b = &E->body[i];
    const OracleConfig *c = &E->cfg;
    b->v = vadd(vmul(b->v, 1.0), vmul(vadd(V(0, 0), vmul(b->f, b->m_inv)), dt));
    b->w = b->w * 1.0 + b->t * b->i_inv * dt;
    b->f = V(0, 0); b->t = 0.0;
    if (i < N_AGENTS) { b->v = vmul(b->v, c->agent_friction); b->w *= c->agent_friction; }
    else              { b->v = vmul(b->v, c->ball_friction); }
    double len = vlen(b->v);
    if (len > c->max_velocity) b->v = vmul(vmul(b->v, 1.0 / len), c->max_velocity);
}

static void space_step(OracleEnv *E, double dt)
{
    /* Chipmunk2D cpSpaceStep */
    E->stamp++;
    double prev_dt = E->prev_dt;
    E->prev_dt = dt;
    for (int k = 0; k < E->n_active; k++) E->arb[E->active[k]].state = ARB_NORMAL;
    E->n_active = 0;

    /* cpBodyUpdatePosition */
    for (int i = 0; i < 5; i++) {
        Body *b = &E->body[i];
        b->p = vadd(b->p, vmul(vadd(b->v, b->v_bias), dt));
        b->a = b->a + (b->w + b->w_bias) * dt;
        b->v_bias = V(0, 0); b->w_bias = 0.0;
    }

    /* collision detection, canonical order = ascending pair id */
    Manifold m;
    for (int i = 0; i < N_AGENTS; i++)
        for (int s = 0; s < N_SEGS; s++) {
            collide_segment_box(E, s, i, &m);
            if (m.count) arbiter_update(E, pair_agent_seg(i, s), STATIC_BODY, i, E->seg[s].e * 0.2, E->seg[s].u * 0.8, &m);
        }
    for (int i = 0; i < N_AGENTS; i++)
        for (int j = i + 1; j < N_AGENTS; j++) {
            collide_box_box(E, i, j, &m);
            if (m.count) arbiter_update(E, pair_agent_agent(i, j), i, j, 0.2 * 0.2, 0.8 * 0.8, &m);
        }
    for (int i = 0; i < N_AGENTS; i++) {
        collide_ball_box(E, i, &m);
        if (m.count) arbiter_update(E, pair_ball_agent(i), BALL, i, 0.95 * 0.2, 0.2 * 0.8, &m);
    }
    for (int s = 0; s < 6; s++) { /* goal lines are filtered out for the ball (mask) */
        collide_ball_segment(E, s, &m);
        if (m.count) arbiter_update(E, pair_ball_wall(s), BALL, STATIC_BODY, 0.95 * E->seg[s].e, 0.2 * E->seg[s].u, &m);
    }

    /* cpSpaceArbiterSetFilter: collision_persistence = 3 */
    for (int p = 0; p < ORACLE_N_PAIRS; p++) {
        Arbiter *arb = &E->arb[p];
        if (!arb->exists) continue;
        long ticks = E->stamp - arb->stamp;
        if (ticks >= 1 && arb->state != ARB_CACHED) arb->state = ARB_CACHED;
        if (ticks >= 3) { arb->exists = 0; arb->count = 0; }
    }

    /* cpArbiterPreStep: slop 0.1, bias = 1 - collision_bias^dt */
    const double slop = 0.1;
    const double biasCoef = 1.0 - pow(pow(1.0 - 0.1, 60.0), dt);
    for (int k = 0; k < E->n_active; k++) {
        Arbiter *arb = &E->arb[E->active[k]];
        Body *a = &E->body[arb->body_a], *b = &E->body[arb->body_b];
        v2 n = arb->n;
        v2 body_delta = vsub(b->p, a->p);
        for (int i = 0; i < arb->count; i++) {
            Contact *c = &arb->con[i];
            c->nMass = 1.0 / (k_scalar_body(a, c->r1, n) + k_scalar_body(b, c->r2, n));
            c->tMass = 1.0 / (k_scalar_body(a, c->r1, vperp(n)) + k_scalar_body(b, c->r2, vperp(n)));
            double dist = vdot(vadd(vsub(c->r2, c->r1), body_delta), n);
            c->bias = -biasCoef * fmin(0.0, dist + slop) / dt;
            c->jBias = 0.0;
            c->bounce = vdot(relative_velocity(a, b, c->r1, c->r2), n) * arb->e;
        }
    }

    for (int i = 0; i < 5; i++) velocity_func(E, i, dt);

    /* cpArbiterApplyCachedImpulse */
    double dt_coef = (prev_dt == 0.0 ? 0.0 : dt / prev_dt);
    for (int k = 0; k < E->n_active; k++) {
        Arbiter *arb = &E->arb[E->active[k]];
        if (arb->state == ARB_FIRST) continue;
        Body *a = &E->body[arb->body_a], *b = &E->body[arb->body_b];
        for (int i = 0; i < arb->count; i++) {
            Contact *c = &arb->con[i];
            v2 j = vmul(vrotate(arb->n, V(c->jnAcc, c->jtAcc)), dt_coef);
            apply_impulse(a, vneg(j), c->r1);
            apply_impulse(b, j, c->r2);
        }
    }

    /* cpArbiterApplyImpulse x iterations (10) */
    for (int it = 0; it < 10; it++) {
        for (int k = 0; k < E->n_active; k++) {
            Arbiter *arb = &E->arb[E->active[k]];
            Body *a = &E->body[arb->body_a], *b = &E->body[arb->body_b];
            v2 n = arb->n;
            for (int i = 0; i < arb->count; i++) {
                Contact *c = &arb->con[i];
                v2 r1 = c->r1, r2 = c->r2;
                v2 vb1 = vadd(a->v_bias, vmul(vperp(r1), a->w_bias));
                v2 vb2 = vadd(b->v_bias, vmul(vperp(r2), b->w_bias));
                v2 vr = relative_velocity(a, b, r1, r2);
                double vbn = vdot(vsub(vb2, vb1), n);
                double vrn = vdot(vr, n);
                double vrt = vdot(vr, vperp(n));
                double jbn = (c->bias - vbn) * c->nMass;
                double jbnOld = c->jBias;
                c->jBias = fmax(jbnOld + jbn, 0.0);
                double jn = -(c->bounce + vrn) * c->nMass;
                double jnOld = c->jnAcc;
                c->jnAcc = fmax(jnOld + jn, 0.0);
                double jtMax = arb->u * c->jnAcc;
                double jt = -vrt * c->tMass;
                double jtOld = c->jtAcc;
                c->jtAcc = dclamp(jtOld + jt, -jtMax, jtMax);
                v2 jb = vmul(n, c->jBias - jbnOld);
                apply_bias_impulse(a, vneg(jb), r1);
                apply_bias_impulse(b, jb, r2);
                v2 j = vrotate(n, V(c->jnAcc - jnOld, c->jtAcc - jtOld));
                apply_impulse(a, vneg(j), r1);
                apply_impulse(b, j, r2);
            }
        }
    }
    /* the static body never moves (m_inv = i_inv = 0) */
}

/* ----------------------------------------------------------- game level */
static void game_reset(OracleEnv *E, int mode)
{
    /* game/game.py:76-118 + soccer_env.py:92-96 */
    E->steps = 0;
    E->mode = mode;
    E->score_blue = 0; E->score_red = 0;
    new_bodies(E);
    apply_spawn(E);
    update_reward_state(E);
    for (int i = 0; i < 4; i++) {
        float f[ORACLE_FRAME];
        frame_for_agent(E, i, f);
        for (int k = 0; k < 3; k++) memcpy(E->frames[i][k], f, sizeof f);
    }
}

OracleEnv *oracle_create(const OracleConfig *cfg, uint64_t seed, uint64_t global_index)
{
    OracleEnv *E = (OracleEnv *)calloc(1, sizeof *E);
    if (!E) return NULL;
    E->cfg = *cfg;
    E->seed = seed; E->global_index = global_index; E->spawn_count = 0;
    E->stamp = 0; E->prev_dt = 0.0;
    setup_field(E);
    game_reset(E, ORACLE_MODE_RANDOM); /* Game.__init__ -> setup_field -> reset() */
    return E;
}

void oracle_destroy(OracleEnv *E) { free(E); }

void oracle_reset(OracleEnv *E, int mode, int has_seed, uint64_t seed)
{
    if (has_seed) { E->seed = seed; E->spawn_count = 0; }
    game_reset(E, mode);
}

void oracle_get_obs(const OracleEnv *E, float *obs)
{
    for (int i = 0; i < 4; i++)
        for (int k = 0; k < 3; k++)
            memcpy(obs + i * ORACLE_OBS + k * ORACLE_FRAME, E->frames[i][k], sizeof(float) * ORACLE_FRAME);
}

/* the four 22-float frames of the CURRENT bodies (game/game.py:258-322), without touching the 3-frame history:
   lets the tests build the stacked observation that belongs to a given pair of history poses */
void oracle_frames(const OracleEnv *E, float *frames)
{
    for (int i = 0; i < 4; i++) frame_for_agent(E, i, frames + i * ORACLE_FRAME);
}

void oracle_step(OracleEnv *E, const float *actions, float *obs, double *reward, uint8_t *done, int8_t *goal)
{
    const OracleConfig *c = &E->cfg;
    /* soccer_env.py:110-125: clip to [-1,1] and scale in float32 */
    double fx[4], fy[4], tq[4];
    for (int i = 0; i < 4; i++) {
        float a0 = fminf(fmaxf(actions[3 * i + 0], -1.0f), 1.0f);
        float a1 = fminf(fmaxf(actions[3 * i + 1], -1.0f), 1.0f);
        float a2 = fminf(fmaxf(actions[3 * i + 2], -1.0f), 1.0f);
        fx[i] = (double)(a0 * (float)c->action_force_max);
        fy[i] = (double)(a1 * (float)c->action_force_max);
        tq[i] = (double)(a2 * (float)c->action_torque_max);
    }
    /* game/game.py:379-397 */
    update_reward_state(E);
    E->steps += 1;
    for (int i = 0; i < 5; i++) { E->body[i].f = V(0, 0); E->body[i].t = 0.0; }
    for (int i = 0; i < 4; i++) {
        Body *b = &E->body[i];
        v2 rot = V(cos(b->a), sin(b->a));
        b->f = vadd(b->f, vrotate(V(fx[i], fy[i]), rot)); /* apply_force_at_local_point(.., (0,0)) */
        b->t = tq[i];
    }
    space_step(E, 1.0 / 60.0);

    /* game/game.py:401-412 */
    v2 bp = E->body[BALL].p;
    const double gy_top = SCREEN_H / 2 + GOAL_HEIGHT / 2, gy_bot = SCREEN_H / 2 - GOAL_HEIGHT / 2;
    int g = 0;
    if (bp.x < FIELD_MARGIN && gy_bot < bp.y && bp.y < gy_top) { g = -1; E->score_red++; }
    else if (bp.x > SCREEN_W - FIELD_MARGIN && gy_bot < bp.y && bp.y < gy_top) { g = +1; E->score_blue++; }

    /* game/game.py:324-375 */
    double r = 0.0;
    if (c->ball_proximity_multiplier != 0.0) {
        double d0 = vlen(vsub(E->body[0].p, bp)), d1 = vlen(vsub(E->body[1].p, bp));
        r += c->ball_proximity_multiplier * ((E->prev_d[0] - d0) + (E->prev_d[1] - d1));
    }
    double D = vlen(vsub(bp, V(SCREEN_W - FIELD_MARGIN, SCREEN_H / 2)));
    r += (E->prev_D_red - D) * c->move_ball_to_goal_multiplier;
    if (g > 0) r += c->goal_scored_reward;
    if (g < 0) r -= c->goal_conceded_penalty;
    r -= c->alive_penalty;

    if (g != 0) apply_spawn(E); /* soft reset, game/game.py:421-422 */

    int dn = 0;
    if (c->max_steps > 0 && E->steps >= c->max_steps) { /* game/game.py:425-433 */
        dn = 1;
        r = c->score_difference_multiplier * (double)(E->score_blue - E->score_red);
    }

    /* game/game.py:435 + soccer_env.py:130-140 */
    for (int i = 0; i < 4; i++) {
        memcpy(E->frames[i][0], E->frames[i][1], sizeof(float) * ORACLE_FRAME);
        memcpy(E->frames[i][1], E->frames[i][2], sizeof(float) * ORACLE_FRAME);
        frame_for_agent(E, i, E->frames[i][2]);
    }
    if (obs) oracle_get_obs(E, obs);
    reward[0] = r; reward[1] = r;
    *done = (uint8_t)dn;
    *goal = (int8_t)g;
}

/* ----------------------------------------------------- state inject/extract */
void oracle_get_state(const OracleEnv *E, OracleState *S)
{
    memset(S, 0, sizeof *S);
    for (int i = 0; i < 5; i++) {
        const Body *b = &E->body[i];
        S->pos[i][0] = b->p.x; S->pos[i][1] = b->p.y;
        S->vel[i][0] = b->v.x; S->vel[i][1] = b->v.y;
        S->ang[i] = b->a; S->angvel[i] = b->w;
        S->vbias[i][0] = b->v_bias.x; S->vbias[i][1] = b->v_bias.y; S->wbias[i] = b->w_bias;
    }
    S->steps = E->steps; S->score[0] = E->score_blue; S->score[1] = E->score_red;
    S->mode = E->mode; S->spawn_count = E->spawn_count; S->seed = E->seed;
    for (int i = 0; i < 4; i++)
        for (int k = 0; k < 3; k++)
            memcpy(&S->obs[i][k * ORACLE_FRAME], E->frames[i][k], sizeof(float) * ORACLE_FRAME);
    int n = 0;
    for (int p = 0; p < ORACLE_N_PAIRS; p++) {
        const Arbiter *arb = &E->arb[p];
        if (!arb->exists) continue;
        for (int i = 0; i < arb->count && n < ORACLE_MAX_CACHE; i++) {
            S->cache_pair[n] = p; S->cache_key[n] = arb->con[i].key;
            S->cache_age[n] = (int)(E->stamp - arb->stamp);
            S->cache_jn[n] = arb->con[i].jnAcc; S->cache_jt[n] = arb->con[i].jtAcc;
            n++;
        }
    }
    S->cache_count = n;
}

void oracle_set_state(OracleEnv *E, const OracleState *S)
{
    for (int i = 0; i < 5; i++) {
        Body *b = &E->body[i];
        b->p = V(S->pos[i][0], S->pos[i][1]);
        b->v = V(S->vel[i][0], S->vel[i][1]);
        b->a = S->ang[i]; b->w = S->angvel[i];
        b->v_bias = V(S->vbias[i][0], S->vbias[i][1]); b->w_bias = S->wbias[i];
        b->f = V(0, 0); b->t = 0.0;
    }
    E->steps = S->steps; E->score_blue = S->score[0]; E->score_red = S->score[1];
    E->mode = S->mode; E->spawn_count = S->spawn_count; E->seed = S->seed;
    for (int i = 0; i < 4; i++)
        for (int k = 0; k < 3; k++)
            memcpy(E->frames[i][k], &S->obs[i][k * ORACLE_FRAME], sizeof(float) * ORACLE_FRAME);
    memset(E->arb, 0, sizeof E->arb);
    E->n_active = 0;
    if (E->stamp < 8) E->stamp = 8;
    if (E->prev_dt == 0.0) E->prev_dt = 1.0 / 60.0;
    for (uint32_t k = 0; k < S->cache_count; k++) {
        int p = S->cache_pair[k];
        Arbiter *arb = &E->arb[p];
        if (!arb->exists) {
            arb->exists = 1; arb->count = 0;
            arb->stamp = E->stamp - S->cache_age[k];
            if (S->cache_age[k] == 0) { arb->state = ARB_NORMAL; E->active[E->n_active++] = p; }
            else arb->state = ARB_CACHED;
        }
        if (arb->count < 2) {
            Contact *c = &arb->con[arb->count++];
            memset(c, 0, sizeof *c);
            c->key = S->cache_key[k]; c->jnAcc = S->cache_jn[k]; c->jtAcc = S->cache_jt[k];
        }
    }
    update_reward_state(E);
}

int oracle_contact_count(const OracleEnv *E)
{
    int n = 0;
    for (int k = 0; k < E->n_active; k++) n += E->arb[E->active[k]].count;
    return n;
}

/* ------------------------------------------------------------- vec level */
struct OracleVec {
    int64_t n;
    OracleEnv **env;
};

OracleVec *oracle_vec_create(const OracleConfig *cfg, int64_t n, uint64_t seed, uint64_t global_offset)
{
    OracleVec *Vv = (OracleVec *)calloc(1, sizeof *Vv);
    Vv->n = n;
    Vv->env = (OracleEnv **)calloc((size_t)n, sizeof(OracleEnv *));
    for (int64_t i = 0; i < n; i++) Vv->env[i] = oracle_create(cfg, seed, global_offset + (uint64_t)i);
    return Vv;
}

void oracle_vec_destroy(OracleVec *Vv)
{
    if (!Vv) return;
    for (int64_t i = 0; i < Vv->n; i++) oracle_destroy(Vv->env[i]);
    free(Vv->env); free(Vv);
}

OracleEnv *oracle_vec_env(OracleVec *Vv, int64_t i) { return Vv->env[i]; }

void oracle_vec_reset(OracleVec *Vv, const uint8_t *mask, int mode, int has_seed, uint64_t seed, float *obs)
{
    /* marl_vecenv.py:18-28: env i is seeded with seed + i (i = global env index) */
    for (int64_t i = 0; i < Vv->n; i++) {
        if (mask && !mask[i]) continue;
        oracle_reset(Vv->env[i], mode, has_seed, seed + Vv->env[i]->global_index);
    }
    if (obs) for (int64_t i = 0; i < Vv->n; i++) oracle_get_obs(Vv->env[i], obs + i * 4 * ORACLE_OBS);
}

void oracle_vec_step(OracleVec *Vv, const float *actions, float *obs, double *reward, uint8_t *done,
                     int8_t *goal, int auto_reset, int nthreads)
{
    /* marl_vecenv.py:30-68: sequential loop in the reference; threads here only
       for the "all host cores" CPU baseline. */
    int64_t n = Vv->n;
#ifdef _OPENMP
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for num_threads(nthreads) schedule(static)
#endif
    for (int64_t i = 0; i < n; i++) {
        OracleEnv *E = Vv->env[i];
        oracle_step(E, actions + i * 12, obs ? obs + i * 4 * ORACLE_OBS : NULL, reward + i * 2, done + i, goal + i);
        if (auto_reset && done[i]) {
            oracle_reset(E, ORACLE_MODE_FULL_RANDOM, 0, 0);
            if (obs) oracle_get_obs(E, obs + i * 4 * ORACLE_OBS);
        }
    }
    (void)nthreads;
}
