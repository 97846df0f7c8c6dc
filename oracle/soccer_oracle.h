/*
 * soccer_oracle.h -- C interface of the CPU oracle (TEST INFRASTRUCTURE ONLY).
 * See soccer_oracle.c for provenance; "parity unpinned" at the pymunk boundary.
 */
#ifndef SOCCER_ORACLE_H
#define SOCCER_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_FRAME 22
#define ORACLE_OBS 66
#define ORACLE_N_PAIRS 48
#define ORACLE_MAX_CACHE 96

#define ORACLE_MODE_RANDOM 0      /* game/game.py:154-190 (default) */
#define ORACLE_MODE_FIXED 1       /* game/game.py:129-152 */
#define ORACLE_MODE_FULL_RANDOM 2 /* game/game.py:192-249 */

/* config.json keys (soccer_env.py:42-64, game/game.py:27,264,330-372,430) */
typedef struct {
    double max_velocity, agent_mass, ball_mass, agent_friction, ball_friction;
    double agent_moment, ball_moment; /* entities.py:11,62: literals 100 and 10 */
    double action_force_max, action_torque_max, max_angular_velocity;
    double ball_proximity_multiplier, move_ball_to_goal_multiplier, goal_scored_reward;
    double goal_conceded_penalty, alive_penalty, score_difference_multiplier;
    int32_t max_steps;
    int32_t _pad;
} OracleConfig;

/* Pair ids shared with the device state format (include/msoc.h):
   agent i x segment s: i*8+s; agents (i<j): 32+{01,02,03,12,13,23}; ball x agent i: 38+i;
   ball x wall s: 42+s. */
typedef struct {
    double pos[5][2], vel[5][2], ang[5], angvel[5], vbias[5][2], wbias[5];
    int32_t steps, score[2], mode;
    uint32_t spawn_count, cache_count;
    uint64_t seed;
    float obs[4][ORACLE_OBS];
    int32_t cache_pair[ORACLE_MAX_CACHE], cache_key[ORACLE_MAX_CACHE], cache_age[ORACLE_MAX_CACHE];
    double cache_jn[ORACLE_MAX_CACHE], cache_jt[ORACLE_MAX_CACHE];
} OracleState;

typedef struct OracleEnv OracleEnv;
typedef struct OracleVec OracleVec;

void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

OracleEnv *oracle_create(const OracleConfig *cfg, uint64_t seed, uint64_t global_index);
void oracle_destroy(OracleEnv *E);
void oracle_reset(OracleEnv *E, int mode, int has_seed, uint64_t seed);
void oracle_get_obs(const OracleEnv *E, float *obs /* 4*66 */);
void oracle_frames(const OracleEnv *E, float *frames /* 4*22, of the current bodies */);
void oracle_step(OracleEnv *E, const float *actions /* 4*3 */, float *obs /* 4*66 or NULL */,
                 double *reward /* 2 */, uint8_t *done, int8_t *goal /* +1 blue, -1 red */);
void oracle_get_state(const OracleEnv *E, OracleState *S);
void oracle_set_state(OracleEnv *E, const OracleState *S);
int oracle_contact_count(const OracleEnv *E);

OracleVec *oracle_vec_create(const OracleConfig *cfg, int64_t n, uint64_t seed, uint64_t global_offset);
void oracle_vec_destroy(OracleVec *V);
OracleEnv *oracle_vec_env(OracleVec *V, int64_t i);
void oracle_vec_reset(OracleVec *V, const uint8_t *mask, int mode, int has_seed, uint64_t seed, float *obs);
void oracle_vec_step(OracleVec *V, const float *actions, float *obs, double *reward, uint8_t *done,
                     int8_t *goal, int auto_reset, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
