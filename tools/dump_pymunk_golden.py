#!/usr/bin/env python
"""Anchor for the oracle at the pymunk boundary ("parity unpinned").

Wherever the reference's dependencies exist (pymunk, pygame, numpy; NOT in the build image of this repo: no network),
this script drives the UNMODIFIED reference `SoccerEnv` / `Game` (soccer_simulation/soccer_env.py:100-154,
game/game.py:378-437) from the 256 start states of tests/golden/step_v2.npz with the same two actions per env, and
writes what the real Chipmunk2D step produces to tests/golden/pymunk_v2.npz.  tests/test_pymunk_golden.py then checks
the oracle, the host build of the kernel arithmetic and the CUDA kernels against that file (and skips, saying why,
while the file is absent).

    python tools/dump_pymunk_golden.py /path/to/marl-soccer/soccer_simulation

The start states carry no cached arbiters and zero bias velocities, so setting the bodies' position / velocity /
angle / angular velocity reproduces them exactly in a fresh pymunk space; what the FIRST step returns is then a pure
function of the state and the action.  The second step continues from pymunk's own state (its arbiter cache and bias
velocities are not visible through the Python API), so it is compared with the oracle's free-running second step."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden", "step_v2.npz")
OUT = os.path.join(ROOT, "tests", "golden", "pymunk_v2.npz")


def main():
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    ref = os.path.abspath(sys.argv[1])
    sys.path.insert(0, ref)
    try:
        import pymunk  # noqa: F401
        import pygame  # noqa: F401  (game/game.py:1 imports it unconditionally)
    except ImportError as ex:
        raise SystemExit(f"the reference cannot run here: {ex}")
    import soccer_env as ref_env  # the reference's own module

    z = np.load(GOLD)
    n = len(z["s0_steps"])
    agents = [f"agent_{i}" for i in range(4)]
    out = {k: [] for k in ("pos1", "vel1", "ang1", "angvel1", "rew1", "done1", "goal1", "frame1",
                           "pos2", "vel2", "ang2", "angvel2", "rew2", "done2", "goal2", "frame2")}
    for e in range(n):
        env = ref_env.soccerenv()
        env.reset(seed=0, options={"use_fixed_positions": True})
        g = env._game
        bodies = [a.body for a in g.agents] + [g.ball.body]
        for i, b in enumerate(bodies):
            b.position = tuple(float(x) for x in z["s0_pos"][e, i])
            b.velocity = tuple(float(x) for x in z["s0_vel"][e, i])
            b.angle = float(z["s0_ang"][e, i])
            b.angular_velocity = float(z["s0_angvel"][e, i])
        g.steps = int(z["s0_steps"][e])
        g.score = {"blue": int(z["s0_score"][e, 0]), "red": int(z["s0_score"][e, 1])}
        mode = int(z["s0_mode"][e])
        g._use_fixed_positions, g._use_full_random_positions = mode == 1, mode == 2
        g._rng = np.random.default_rng(int(z["s0_seed"][e]))  # PCG64: re-spawn positions after a goal are NOT comparable
        for k, act in ((1, z["act1"][e]), (2, z["act2"][e])):
            obs, rew, term, trunc, info = env.step({a: act[i] for i, a in enumerate(agents)})
            out[f"pos{k}"].append([tuple(b.position) for b in bodies])
            out[f"vel{k}"].append([tuple(b.velocity) for b in bodies])
            out[f"ang{k}"].append([b.angle for b in bodies])
            out[f"angvel{k}"].append([b.angular_velocity for b in bodies])
            out[f"rew{k}"].append([rew["agent_0"], rew["agent_1"]])
            out[f"done{k}"].append(bool(trunc["agent_0"]))
            gs = info["agent_0"].get("goal_scored_by")
            out[f"goal{k}"].append(0 if gs is None else (1 if gs == "blue" else -1))
            out[f"frame{k}"].append(np.stack([obs[a][44:] for a in agents]))  # the newest frame of every agent
        env.close()
    import pymunk
    np.savez_compressed(OUT, pymunk_version=np.array(pymunk.version), chipmunk_version=np.array(pymunk.chipmunk_version),
                        **{k: np.asarray(v) for k, v in out.items()})
    print(f"wrote {OUT}: {n} envs x 2 steps from pymunk {pymunk.version} (Chipmunk {pymunk.chipmunk_version})")


if __name__ == "__main__":
    main()
