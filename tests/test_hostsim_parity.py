"""fp32 step arithmetic of the kernels (host build of step_core.cuh, tests/hostsim) against the fp64
oracle.  Runs without a GPU; the same checks run on the real kernels in test_gpu_parity.py."""
import numpy as np
import pytest

import hostsim_lib as H
import oracle_lib as O
import parity_util as P

P.add_batch_api(H.HostSim)


def test_spawn_bit_exact_all_modes():
    n = 512
    for mode in (O.MODE_RANDOM, O.MODE_FIXED, O.MODE_FULL_RANDOM):
        sim = H.HostSim(n, P.CONFIG, seed=5, global_offset=1000)
        ora = O.OracleVec(n, P.CONFIG, seed=5, global_offset=1000)
        o_d = sim.reset(mode, seed=77)
        o_o = ora.reset(mode, seed=77)
        for i in range(0, n, 7):
            sd, so = sim.get_state(i), ora.env(i).get_state()
            assert np.array_equal(np.array(sd.pos, np.float64), so["pos"]), (mode, i)
            assert sd.spawn_count == so["spawn_count"] and sd.seed == so["seed"]
        assert np.allclose(o_d, o_o, atol=P.ATOL["obs"])


def test_single_step_injected_states():
    worst, goals = P.check_single_step(H.HostSim, 1024, seed=11, name="hostsim/single_step_1024")
    assert goals > 50  # the 'goal' scenario kind really crosses the line


def test_single_step_non_default_config():
    """Every config key moved (P.ALT_CONFIG): explicit force / angular-velocity scales, the prox == 0 branch, a conceded
    penalty, a terminal bonus, max_steps 37 (readers: soccer_env.py:63-64, game/game.py:264,330,368,430)."""
    worst, goals = P.check_single_step(H.HostSim, 512, seed=12, config=P.ALT_CONFIG, name="hostsim/single_step_alt_config")
    assert goals > 20


def test_tracked_rollout_non_default_config():
    bad, total, ev, worst = P.check_tracked_rollout(H.HostSim, 32, 80, seed=6, mode=O.MODE_FULL_RANDOM, config=P.ALT_CONFIG,
                                                    name="hostsim/tracked_alt_config")
    assert ev["dones"] >= 32 and ev["contacts"] > 200   # 37-step episodes: everybody truncates at least once
    assert bad <= total * P.MAX_TRACKED_FRACTION, (bad, total, worst)


def test_high_torque_fallback_config_keeps_angles_wrapped():
    """config without action_torque_max: the reference's fall-back of 100000 (soccer_env.py:64) spins the agents by
    several turns per step; the kernels' angle stays wrapped and the observation's angle feature stays in [-1, 1]."""
    bad, total, ev, worst = P.check_tracked_rollout(H.HostSim, 16, 40, seed=8, mode=O.MODE_RANDOM, config=P.SPIN_CONFIG,
                                                    name="hostsim/tracked_high_torque")
    assert bad <= total * P.MAX_TRACKED_FRACTION, (bad, total, worst)
    sim = H.HostSim(8, P.SPIN_CONFIG, seed=1)
    sim.reset(O.MODE_FIXED)
    for t in range(30):
        o, *_ = sim.step(np.ones((8, 4, 3), np.float32), auto_reset=False)
        ang = o.reshape(8, 4, 3, 22)[:, :, 2, 2]
        assert np.all(np.abs(ang) <= 1.0 + 1e-6), float(np.abs(ang).max())
    w = np.array([sim.get_state(0).angvel[i] for i in range(4)])
    assert np.all(w > 400.0)  # > 2 pi per step: more than one full turn between two position updates


def test_tracked_rollout_full_random():
    bad, total, ev, worst = P.check_tracked_rollout(H.HostSim, 48, 100, seed=3, mode=O.MODE_FULL_RANDOM, name="hostsim/tracked_full_random")
    assert ev["dones"] > 0 and ev["contacts"] > 500
    assert bad <= total * P.MAX_TRACKED_FRACTION, (bad, total, worst)


def test_tracked_rollout_default_mode():
    bad, total, ev, worst = P.check_tracked_rollout(H.HostSim, 32, 60, seed=4, mode=O.MODE_RANDOM)
    assert bad <= total * P.MAX_TRACKED_FRACTION, (bad, total, worst)


def test_free_running_contact_free_100_steps():
    """100-step free-running rollout (no re-synchronisation) from the fixed kickoff with small forces:
    nobody touches anything, so fp32 and fp64 stay within the single-step tolerance bands."""
    n = 8
    sim = H.HostSim(n, P.CONFIG, seed=1)
    ora = O.OracleVec(n, P.CONFIG, seed=1)
    sim.reset(O.MODE_FIXED)
    ora.reset(O.MODE_FIXED)
    rng = np.random.default_rng(2)
    for t in range(100):
        act = (rng.uniform(-1, 1, (n, 4, 3)) * [0.02, 0.02, 1.0]).astype(np.float32)
        o_d, r_d, d_d, g_d = sim.step(act, auto_reset=False)
        o_o, r_o, d_o, g_o = ora.step(act, auto_reset=False)
    worst, failing, cache_bad = P.compare_all(sim, ora, o_d, o_o, r_d, r_o, n)
    assert not cache_bad
    # 100 steps of accumulated rounding stay inside the single-step band (observed worst ratio 0.17)
    assert max(worst.values()) < 1.0, worst


def test_global_offset_shards_agree():
    """Env i of a 2-shard run equals env i of the 1-shard run (Philox keyed by the global env index)."""
    full = H.HostSim(64, P.CONFIG, seed=9)
    lo = H.HostSim(32, P.CONFIG, seed=9, global_offset=0)
    hi = H.HostSim(32, P.CONFIG, seed=9, global_offset=32)
    a = full.reset(O.MODE_FULL_RANDOM, seed=123)
    b = np.concatenate([lo.reset(O.MODE_FULL_RANDOM, seed=123), hi.reset(O.MODE_FULL_RANDOM, seed=123)])
    assert np.array_equal(a, b)
    rng = np.random.default_rng(0)
    for _ in range(30):
        act = rng.uniform(-1, 1, (64, 4, 3)).astype(np.float32)
        fa = full.step(act)
        la, ha = lo.step(act[:32]), hi.step(act[32:])
        for x, y, z in zip(fa, la, ha):
            assert np.array_equal(x, np.concatenate([y, z]))


def test_class_modes_equal_the_general_path(monkeypatch):
    """Every work class of the contact kernel (light: one agent x wall pair; pair: one agent x agent / ball x agent pair;
    multi: several wall pairs) solves its islands on their own.  From identical states the result must be the general
    path's (HSIM_FORCE_FULL routes every declined env through MODE_FULL) -- the same numbers in every state field on the
    host build, where nothing is contracted -- arbiter cache included."""
    n = 512
    a, b = H.HostSim(n, P.CONFIG, seed=1), H.HostSim(n, P.CONFIG, seed=1)
    a.reset(O.MODE_FULL_RANDOM, seed=1)
    for i in range(n):
        s = a.get_state(i)
        s.steps = (i * 2654435761) % 1000
        a.set_state(i, s)
    rng = np.random.default_rng(0)
    for _ in range(150):
        a.step(rng.uniform(-1, 1, (n, 4, 3)).astype(np.float32))
    seen = {0: 0, 1: 0, 2: 0, 3: 0}
    for _ in range(25):
        for i in range(n):
            b.set_state(i, a.get_state(i))
        act = rng.uniform(-1, 1, (n, 4, 3)).astype(np.float32)
        monkeypatch.delenv("HSIM_FORCE_FULL", raising=False)
        out_a = a.step(act)
        cont_a, load_a = a.last()
        monkeypatch.setenv("HSIM_FORCE_FULL", "1")
        out_b = b.step(act)
        cont_b, _ = b.last()
        monkeypatch.delenv("HSIM_FORCE_FULL", raising=False)
        for x, y in zip(out_a, out_b):
            assert np.array_equal(x, y)
        assert np.array_equal(cont_a, cont_b)
        for k in seen:
            seen[k] += int(np.sum((load_a == k) & (cont_a > 0)))
        for i in np.nonzero(load_a >= 0)[0]:
            sa, sb = a.get_state(int(i)), b.get_state(int(i))
            for name, _ in type(sa)._fields_:
                # equal VALUES in every field (a zero tangent impulse of a frictionless goal-line contact may come out as
                # -0.0 on one path and +0.0 on the other: min / max of two zeros, which the host compiler may order either way)
                assert np.array_equal(np.array(getattr(sa, name)), np.array(getattr(sb, name))), (int(i), int(load_a[i]), name)
    # every class with a path of its own really solved contacts (class 1, the general path, is what the other side runs
    # for everything; on this mix it is down to ~0.1 % of the envs)
    assert min(seen[k] for k in (0, 2, 3)) > 20, seen
