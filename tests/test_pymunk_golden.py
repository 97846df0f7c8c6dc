"""The oracle, the host build of the kernel arithmetic and the CUDA kernels against vectors produced by the REAL
reference on real pymunk (tests/golden/pymunk_v2.npz, written by tools/dump_pymunk_golden.py wherever pymunk and
pygame exist).  The build image of this repo has neither and no network, so the file cannot be generated here: the
tests skip and say so.  This is the hook that would move the oracle from "parity unpinned" to pinned."""
import os

import numpy as np
import pytest

import golden_util as G
import oracle_lib as O
import parity_util as P

HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(HERE, "golden", "pymunk_v2.npz")
GOLD = os.path.join(HERE, "golden", "step_v2.npz")

needs_file = pytest.mark.skipif(
    not os.path.exists(PATH),
    reason="tests/golden/pymunk_v2.npz is absent: pymunk / pygame are not installable in this image (no network); "
           "run tools/dump_pymunk_golden.py <reference>/soccer_simulation where they exist")


def _first_step_against_pymunk(step_fn, n, z, ref):
    """step_fn(act) -> (obs, rew, done, goal, states after the step as oracle-style dicts)."""
    obs, rew, done, goal, states = step_fn(z["act1"])
    assert np.array_equal(done.astype(bool), ref["done1"])
    assert np.array_equal(goal, ref["goal1"])
    keep = goal == 0  # after a goal the re-spawn draws come from PCG64 in the reference, from Philox here
    worst = {"pos": 0.0, "vel": 0.0, "angvel": 0.0, "ang": 0.0, "reward": 0.0, "frame": 0.0}
    for i in np.nonzero(keep)[0]:
        s = states[i]
        for key, rk in (("pos", "pos1"), ("vel", "vel1"), ("angvel", "angvel1")):
            a, b = np.asarray(s[key], np.float64), np.asarray(ref[rk][i], np.float64)
            worst[key] = max(worst[key], float(np.max(np.abs(a - b) / (P.ATOL[key] + P.RTOL * np.abs(b)))))
        worst["ang"] = max(worst["ang"], float(np.max(P.ang_diff(np.asarray(s["ang"])[:4], ref["ang1"][i][:4]) / P.ATOL["ang"])))
        worst["reward"] = max(worst["reward"], abs(float(rew[i, 0]) - float(ref["rew1"][i, 0])) /
                              (P.ATOL["reward"] + P.RTOL * abs(float(ref["rew1"][i, 0]))))
        f = np.asarray(obs[i]).reshape(4, 3, 22)[:, 2]
        stacked_d = np.concatenate([f, f, f], axis=1)
        stacked_r = np.concatenate([ref["frame1"][i]] * 3, axis=1)
        worst["frame"] = max(worst["frame"], P.compare_obs(stacked_d, stacked_r))
    P.record("pymunk/first_step", {"envs": int(keep.sum()), "worst_violation_ratio": worst})
    assert max(worst.values()) <= P.MAX_RATIO, worst


@needs_file
def test_oracle_first_step_equals_real_pymunk():
    z, ref = np.load(GOLD), np.load(PATH)
    s0 = G.unpack(z, "s0")
    n = len(s0)
    ora = O.OracleVec(n, P.CONFIG, seed=0)
    ora.set_states(s0)

    def step(act):
        o, r, d, g = ora.step(act, auto_reset=False)
        return o, r, d, g, ora.get_states()
    _first_step_against_pymunk(step, n, z, ref)


@needs_file
def test_host_build_first_step_equals_real_pymunk():
    import hostsim_lib as H
    z, ref = np.load(GOLD), np.load(PATH)
    s0 = G.unpack(z, "s0")
    n = len(s0)
    sim = H.HostSim(n, P.CONFIG, seed=0)
    P.add_batch_api(H.HostSim)
    sim.set_states(np.arange(n), [P.oracle_to_dev_state(s) for s in s0])

    def step(act):
        o, r, d, g = sim.step(act, auto_reset=False)
        st = sim.get_states(np.arange(n))
        return o, r, d, g, [P.dev_to_oracle_state(st[i], o[i]) for i in range(n)]
    _first_step_against_pymunk(step, n, z, ref)


@needs_file
@pytest.mark.gpu
def test_kernels_first_step_equals_real_pymunk():
    from marl_soccer_b200.host_api import HostBufferSim
    z, ref = np.load(GOLD), np.load(PATH)
    s0 = G.unpack(z, "s0")
    n = len(s0)
    sim = HostBufferSim(n, P.CONFIG, seed=0)
    sim.set_states(np.arange(n), [P.oracle_to_dev_state(s) for s in s0])

    def step(act):
        o, r, d, g = sim.step(act, auto_reset=False)
        st = sim.get_states(np.arange(n))
        return o, r, d, g, [P.dev_to_oracle_state(st[i], o[i]) for i in range(n)]
    _first_step_against_pymunk(step, n, z, ref)


def test_the_hook_is_wired():
    """Without the file: the dump tool exists, names the reference entry points it drives, and the golden start states
    it reads carry no arbiter cache and no bias velocities (so a fresh pymunk space reproduces them exactly)."""
    tool = open(os.path.join(os.path.dirname(HERE), "tools", "dump_pymunk_golden.py")).read()
    assert "soccerenv()" in tool and "env.step(" in tool and "pymunk_v2.npz" in tool
    z = np.load(GOLD)
    assert int(z["s0_cache_n"].sum()) == 0 and float(np.abs(z["s0_vbias"]).max()) == 0.0 and float(np.abs(z["s0_wbias"]).max()) == 0.0
