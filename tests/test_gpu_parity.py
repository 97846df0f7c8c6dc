"""Parity of the CUDA kernels (through the C-ABI, include/msoc.h) with the CPU oracle.  -m gpu."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
import parity_util as P

pytestmark = pytest.mark.gpu


def Dev(n, config, seed=0, global_offset=0):
    from marl_soccer_b200.host_api import HostBufferSim
    return HostBufferSim(n, config, seed=seed, global_offset=global_offset)


def test_spawn_bit_exact_all_modes():
    n = 4096
    for mode in (O.MODE_RANDOM, O.MODE_FIXED, O.MODE_FULL_RANDOM):
        sim = Dev(n, P.CONFIG, seed=5, global_offset=1000)
        ora = O.OracleVec(n, P.CONFIG, seed=5, global_offset=1000)
        o_d = sim.reset(mode, seed=77)
        o_o = ora.reset(mode, seed=77)
        idx = np.arange(0, n, 5)
        for i, sd in zip(idx, sim.get_states(idx)):
            so = ora.env(int(i)).get_state()
            assert np.array_equal(np.array(sd.pos, np.float64), so["pos"]), (mode, i)
            assert sd.spawn_count == so["spawn_count"] and sd.seed == so["seed"]
        assert np.allclose(o_d, o_o, atol=P.ATOL["obs"])


def test_single_step_injected_states_4096():
    """BASELINE config 2: 4 096 batched envs, parity vs the restated Game on injected states."""
    worst, goals = P.check_single_step(Dev, 4096, seed=11, name="cuda/single_step_4096")
    assert goals > 200
    print("worst violation ratios", worst)


def test_single_step_non_default_config_4096():
    """Every config key moved (P.ALT_CONFIG): explicit force / angular-velocity scales, the prox == 0 branch, a conceded
    penalty, a terminal bonus, max_steps 37 (readers: soccer_env.py:63-64, game/game.py:264,330,368,430)."""
    worst, goals = P.check_single_step(Dev, 4096, seed=12, config=P.ALT_CONFIG, name="cuda/single_step_alt_config_4096")
    assert goals > 100


def test_tracked_rollout_non_default_config():
    bad, total, ev, worst = P.check_tracked_rollout(Dev, 96, 80, seed=6, mode=O.MODE_FULL_RANDOM, config=P.ALT_CONFIG,
                                                    name="cuda/tracked_alt_config")
    assert ev["dones"] >= 96 and ev["contacts"] > 500
    assert bad <= total * P.MAX_TRACKED_FRACTION, (bad, total, worst)


def test_high_torque_fallback_config_keeps_angles_wrapped():
    """config without action_torque_max: the reference's fall-back of 100000 (soccer_env.py:64) spins the agents by
    several turns per step; the kernels' angle stays wrapped and the observation's angle feature stays in [-1, 1]."""
    bad, total, ev, worst = P.check_tracked_rollout(Dev, 64, 40, seed=8, mode=O.MODE_RANDOM, config=P.SPIN_CONFIG,
                                                    name="cuda/tracked_high_torque")
    assert bad <= total * P.MAX_TRACKED_FRACTION, (bad, total, worst)
    sim = Dev(256, P.SPIN_CONFIG, seed=1)
    sim.reset(O.MODE_FIXED)
    for t in range(30):
        o, *_ = sim.step(np.ones((256, 4, 3), np.float32), auto_reset=False)
        ang = o.reshape(256, 4, 3, 22)[:, :, 2, 2]
        assert np.all(np.abs(ang) <= 1.0 + 1e-6), float(np.abs(ang).max())
    assert all(sim.get_state(0).angvel[i] > 400.0 for i in range(4))  # > 2 pi per step


def test_tracked_rollout_full_random_100_steps():
    bad, total, ev, worst = P.check_tracked_rollout(Dev, 128, 100, seed=3, mode=O.MODE_FULL_RANDOM, name="cuda/tracked_full_random")
    assert ev["dones"] > 0 and ev["contacts"] > 1000
    assert bad <= total * P.MAX_TRACKED_FRACTION, (bad, total, worst)
    print("tracked rollout", bad, total, ev, worst)


def test_tracked_rollout_default_mode():
    bad, total, ev, worst = P.check_tracked_rollout(Dev, 64, 60, seed=4, mode=O.MODE_RANDOM, name="cuda/tracked_default_mode")
    assert bad <= total * P.MAX_TRACKED_FRACTION, (bad, total, worst)


def test_free_running_contact_free_100_steps():
    n = 64
    sim = Dev(n, P.CONFIG, seed=1)
    ora = O.OracleVec(n, P.CONFIG, seed=1)
    sim.reset(O.MODE_FIXED)
    ora.reset(O.MODE_FIXED)
    rng = np.random.default_rng(2)
    for t in range(100):
        act = (rng.uniform(-1, 1, (n, 4, 3)) * [0.02, 0.02, 1.0]).astype(np.float32)
        o_d, r_d, d_d, g_d = sim.step(act, auto_reset=False)
        o_o, r_o, d_o, g_o = ora.step(act, auto_reset=False)
        assert np.array_equal(d_d, d_o) and np.array_equal(g_d, g_o)
    worst, failing, cache_bad = P.compare_all(sim, ora, o_d, o_o, r_d, r_o, n)
    assert not cache_bad
    P.record("cuda/free_running_contact_free_100_steps", P.summarize(failing, worst, n))
    assert max(worst.values()) < 1.0, worst  # observed on the B200: 0.17


def test_free_running_through_contacts_divergence_curve():
    """North star: "100-step rollouts from shared states".  Through contacts fp32 and fp64 trajectories separate
    chaotically (a contact one step earlier or later), so beyond the first steps this cannot be an assertion about
    every env: the test MEASURES it -- fraction of envs with every quantity inside the single-step band, and goal /
    done flag mismatches, after k = 1, 10, 30, 100 free-running steps of 4 096 envs from one shared full-random reset
    (corner spawns: a quarter of the envs start in contact) -- records the curve in the parity report and asserts
    what must hold: done flags always (a pure step count), the first step inside the band for (almost) all envs, and
    a majority still inside after 100 steps."""
    n = 4096
    sim = Dev(n, P.CONFIG, seed=31)
    ora = O.OracleVec(n, P.CONFIG, seed=31)
    sim.reset(O.MODE_FULL_RANDOM, seed=6)
    ora.reset(O.MODE_FULL_RANDOM, seed=6)
    rng = np.random.default_rng(9)
    curve = {}
    goal_mismatch = 0
    for t in range(1, 101):
        act = rng.uniform(-1, 1, (n, 4, 3)).astype(np.float32)
        o_d, r_d, d_d, g_d = sim.step(act, auto_reset=False)
        o_o, r_o, d_o, g_o = ora.step(act, auto_reset=False, nthreads=8)
        assert np.array_equal(d_d, d_o)
        goal_mismatch += int((g_d != g_o).sum())
        if t in (1, 10, 30, 100):
            worst, failing, cache_bad = P.compare_all(sim, ora, o_d, o_o, r_d, r_o, n)
            bad = {i for i, _ in failing} | set(cache_bad)
            curve[t] = {"fraction_in_band": round(1.0 - len(bad) / n, 4), "cache_structure_mismatches": len(set(cache_bad)),
                        "goal_flag_mismatches_so_far": goal_mismatch,
                        "median_worst_ratio_of_out_of_band_envs": round(float(np.median([max(d.values()) for _, d in failing])), 1) if failing else 0.0}
    P.record("cuda/free_running_through_contacts_4096", {"envs": n, "curve": curve})
    print("free-running divergence curve", curve)
    assert curve[1]["fraction_in_band"] >= 0.995
    assert curve[100]["fraction_in_band"] >= 0.5


def test_global_offset_shards_agree():
    full = Dev(4096, P.CONFIG, seed=9)
    lo = Dev(2048, P.CONFIG, seed=9, global_offset=0)
    hi = Dev(2048, P.CONFIG, seed=9, global_offset=2048)
    a = full.reset(O.MODE_FULL_RANDOM, seed=123)
    b = np.concatenate([lo.reset(O.MODE_FULL_RANDOM, seed=123), hi.reset(O.MODE_FULL_RANDOM, seed=123)])
    assert np.array_equal(a, b)
    rng = np.random.default_rng(0)
    for _ in range(40):
        act = rng.uniform(-1, 1, (4096, 4, 3)).astype(np.float32)
        fa = full.step(act)
        la, ha = lo.step(act[:2048]), hi.step(act[2048:])
        for x, y, z in zip(fa, la, ha):
            assert np.array_equal(x, np.concatenate([y, z]))


def test_chunked_host_step_equals_the_whole_step():
    """msoc_step_host cuts large batches into pipeline chunks on two streams (H2D and kernels of one chunk overlap the
    D2H copies of the previous one); msoc_step_host_frames returns only the newest frame.  Both must be the very same
    simulation as the one-launch device step: bit-identical outputs over steps that include goals and auto-resets."""
    import torch
    from marl_soccer_b200 import _capi
    from marl_soccer_b200.host_api import HostBufferSim
    from marl_soccer_b200.sim import BatchedSoccerSim
    n = 200_000  # three chunks
    cfg = {**P.CONFIG, "simulation": {"max_steps": 9}}
    host = HostBufferSim(n, cfg, seed=5)
    fr = HostBufferSim(n, cfg, seed=5, pinned=True)  # persistent page-locked buffers: views, overwritten by the next step
    dev = BatchedSoccerSim(n, config=cfg, seed=5)
    a = host.reset(O.MODE_FULL_RANDOM, seed=2)
    fr.reset(O.MODE_FULL_RANDOM, seed=2)
    b = dev.reset(O.MODE_FULL_RANDOM, seed=2)
    assert np.array_equal(a, b.cpu().numpy())
    rng = np.random.default_rng(1)
    stack = a.copy()
    seen = {}
    for t in range(12):
        act = rng.uniform(-1, 1, (n, 4, 3)).astype(np.float32)
        o_h, r_h, d_h, g_h = host.step(act)
        o_d, r_d, d_d, g_d = dev.step(torch.from_numpy(act).cuda())
        assert np.array_equal(o_h, o_d.cpu().numpy()) and np.array_equal(r_h, r_d.cpu().numpy()), t
        assert np.array_equal(d_h, d_d.cpu().numpy()) and np.array_equal(g_h, g_d.cpu().numpy()), t
        assert np.array_equal(host.score, dev.score.cpu().numpy()), t
        # the contact kernel's four work classes all occur, and the chunks of the host step add up to the device step's
        cc = dev.class_counts()
        assert sum(cc.values()) < n, cc
        for k, v in cc.items():
            seen[k] = seen.get(k, 0) + v
        hc = (C.c_int32 * 4)()
        _capi.check(_capi.lib().msoc_last_class_counts(host._h, C.byref(hc), None))
        assert list(hc) == [cc["light"], cc["heavy"], cc["pair"], cc["multi"]], (list(hc), cc)
        # frames-only variant: the caller keeps the 3-frame stack (soccer_env.py:130-140, :92-96)
        frames, r_f, d_f, g_f = fr.step_frames(act)
        s4 = stack.reshape(n, 4, 3, 22)
        s4[:, :, 0], s4[:, :, 1], s4[:, :, 2] = s4[:, :, 1].copy(), s4[:, :, 2].copy(), frames
        fresh = d_f.astype(bool)
        s4[fresh] = frames[fresh][:, :, None, :]
        assert np.array_equal(stack, o_h), t
        assert np.array_equal(r_f, r_h) and np.array_equal(d_f, d_h) and np.array_equal(g_f, g_h)
    assert d_h.sum() == 0 and host.stats()["episodes"] == n  # everybody truncated once, at step 9
    # (the corner spawns of the first steps are all multi / pair; light comes later; heavy is down to pile-ups of three)
    assert min(seen[k] for k in ("light", "pair", "multi")) > 1000 and seen["heavy"] > 0, seen


def test_ragged_sizes_and_masked_reset():
    """Env counts that are not multiples of the warp/block tile, and masked resets."""
    for n in (1, 31, 33, 127, 129, 1000):
        sim = Dev(n, P.CONFIG, seed=2)
        ora = O.OracleVec(n, P.CONFIG, seed=2)
        o_d, o_o = sim.reset(O.MODE_FULL_RANDOM, seed=5), ora.reset(O.MODE_FULL_RANDOM, seed=5)
        assert np.allclose(o_d, o_o, atol=P.ATOL["obs"])
        act = np.random.default_rng(n).uniform(-1, 1, (n, 4, 3)).astype(np.float32)
        o_d, r_d, d_d, g_d = sim.step(act)
        o_o, r_o, d_o, g_o = ora.step(act)
        worst, failing, cache_bad = P.compare_all(sim, ora, o_d, o_o, r_d, r_o, n)
        assert not cache_bad and not failing, (n, failing[:3])
        mask = (np.arange(n) % 3 == 0).astype(np.uint8)
        m_d = sim.reset(O.MODE_RANDOM, mask=mask)
        m_o = ora.reset(O.MODE_RANDOM, mask=mask)
        keep = mask == 0
        assert np.allclose(m_d[mask == 1], m_o[mask == 1], atol=P.ATOL["obs"])
        assert np.array_equal(m_d[keep], o_d[keep])  # untouched envs keep their observation rows


def test_full_size_invariants_65536():
    """BASELINE config 3 size: properties that do not need the oracle at full size."""
    import torch
    from marl_soccer_b200.sim import BatchedSoccerSim
    n = 65536
    sim = BatchedSoccerSim(n, config=P.CONFIG, seed=0)
    sim.reset(2, seed=0)
    st = sim.get_states(np.arange(0, n, 997))
    g = torch.Generator(device="cuda").manual_seed(1)
    prev = sim.obs.clone()
    total_done = 0
    # put every env 5 steps before truncation so that the auto-reset path runs at full size
    some = sim.get_states(np.arange(n))
    for s in some:
        s.steps = 995
    sim.set_states(np.arange(n), some)
    for t in range(8):
        act = torch.rand((n, 4, 3), generator=g, device="cuda") * 2 - 1
        obs, rew, done, goal = sim.step(act)
        obs_c = obs.clone()
        assert torch.equal(rew[:, 0], rew[:, 1])
        fresh = done.bool()
        total_done += int(fresh.sum())
        assert bool(((t == 4) == fresh).all()), "done <=> steps reached max_steps"
        # frame shift: old frames 1,2 become frames 0,1 unless the env was auto-reset
        keep = ~fresh
        o4 = obs_c.view(n, 4, 3, 22)
        p4 = prev.view(n, 4, 3, 22)
        assert torch.equal(o4[keep][:, :, :2], p4[keep][:, :, 1:])
        if fresh.any():
            f = o4[fresh]
            assert torch.equal(f[:, :, 0], f[:, :, 1]) and torch.equal(f[:, :, 1], f[:, :, 2])
        # unit vectors have norm 1 (or 0)
        u = o4[:, :, 2, 4:].reshape(n, 4, 6, 3)[..., :2]
        nrm = (u * u).sum(-1).sqrt()
        assert bool(((nrm - 1).abs() < 1e-5).logical_or(nrm == 0).all())
        assert bool(torch.isfinite(obs_c).all())
        prev = obs_c
    assert total_done == n
    stats = sim.stats()
    assert stats["episodes"] == n and stats["env_steps"] == 8 * n and stats["contact_overflow"] == 0


def test_free_running_episode_statistics_match_the_oracle():
    """Free-running rollouts through contacts diverge chaotically between fp32 and fp64 (a contact one step earlier or
    later), so whole episodes are compared as populations: 4 096 envs x 1 000 steps (one full episode incl. the corner
    spawns, truncation and auto-reset) with the same i.i.d. actions on both sides; contacts per env-step, goals, episode
    returns and the distribution of the final observations must agree within sampling error."""
    n, steps = 4096, 1001
    sim = Dev(n, P.CONFIG, seed=21)
    ora = O.OracleVec(n, P.CONFIG, seed=21)
    sim.reset(O.MODE_FULL_RANDOM, seed=4)
    ora.reset(O.MODE_FULL_RANDOM, seed=4)
    sim.stats(reset=True)
    rng = np.random.default_rng(8)
    ret_d, ret_o = np.zeros(n), np.zeros(n)
    goals_d = goals_o = 0
    done_d = done_o = 0
    contacts_o = 0
    for t in range(steps):
        act = rng.uniform(-1, 1, (n, 4, 3)).astype(np.float32)
        o_d, r_d, d_d, g_d = sim.step(act)
        o_o, r_o, d_o, g_o = ora.step(act, nthreads=8)
        assert np.array_equal(d_d, d_o), "truncation is a pure step count: it must agree exactly"
        ret_d += r_d[:, 0]; ret_o += r_o[:, 0]
        goals_d += int(np.abs(g_d).sum()); goals_o += int(np.abs(g_o).sum())
        done_d += int(d_d.sum()); done_o += int(d_o.sum())
    st = sim.stats()
    assert done_d == done_o == n and st["episodes"] == n and st["contact_overflow"] == 0
    # goals: Poisson counts of a few dozen
    assert abs(goals_d - goals_o) <= 4 * np.sqrt(max(goals_o, 1)) + 4, (goals_d, goals_o)
    # episode returns: same mean within 4 standard errors, same spread within 10 %
    se = np.sqrt(ret_o.var() / n + ret_d.var() / n)
    assert abs(ret_d.mean() - ret_o.mean()) <= 4 * se + 1e-3, (ret_d.mean(), ret_o.mean(), se)
    assert abs(ret_d.std() / ret_o.std() - 1) < 0.1, (ret_d.std(), ret_o.std())
    # where the agents and the ball end up: mean of every observation feature of the last step
    diff = np.abs(o_d.reshape(n, -1).mean(0) - o_o.reshape(n, -1).mean(0))
    spread = o_o.reshape(n, -1).std(0) / np.sqrt(n)
    assert np.all(diff <= 6 * spread + 2e-3), float((diff / (6 * spread + 2e-3)).max())
    print("episode statistics: goals", goals_d, goals_o, "mean return", ret_d.mean(), ret_o.mean(),
          "contacts/env-step (device)", st["contacts"] / st["env_steps"])


def test_class_modes_match_the_general_path_on_the_device():
    """The light / pair / multi classes solve their contact islands on their own; MSOC_STEP_GENERAL_PATH sends the same
    envs through the general path instead.  From identical states (re-synchronised every step) the two must be the same
    simulation: flags, counters and the arbiter-cache structure bit-exact, everything else within the oracle bands of each
    other (the two code paths only differ in how the compiler contracts their arithmetic, which ten Gauss-Seidel
    iterations over stiff contacts amplify a little; on the host build, where nothing is contracted, they are bit-identical:
    test_hostsim_parity.py)."""
    import torch
    from marl_soccer_b200.sim import BatchedSoccerSim
    n = 4096
    a = BatchedSoccerSim(n, config=P.CONFIG, seed=3)
    b = BatchedSoccerSim(n, config=P.CONFIG, seed=3)
    a.reset(O.MODE_FULL_RANDOM, seed=4)
    idx = np.arange(n)
    st = a.get_states(idx)
    for i, s in enumerate(st):
        s.steps = (i * 2654435761) % 1000
    a.set_states(idx, st)
    g = torch.Generator(device="cuda").manual_seed(5)
    for _ in range(150):
        a.step(torch.rand((n, 4, 3), generator=g, device="cuda") * 2 - 1)
    seen = {}
    worst = {}
    for t in range(40):
        b.set_states(idx, a.get_states(idx))
        act = torch.rand((n, 4, 3), generator=g, device="cuda") * 2 - 1
        oa, ra, da, ga = (x.clone() for x in a.step(act))
        for k, v in a.class_counts().items():
            seen[k] = seen.get(k, 0) + v
        ob, rb, db, gb = b.step(act, general_path=True)
        cb = b.class_counts()
        assert cb["light"] == cb["pair"] == cb["multi"] == 0 and cb["heavy"] > 0
        assert torch.equal(da, db) and torch.equal(ga, gb) and torch.equal(a.score, b.score)
        worst["reward"] = max(worst.get("reward", 0.0), float((ra - rb).abs().max()) / P.ATOL["reward"])
        worst["obs"] = max(worst.get("obs", 0.0), float((oa - ob).abs().max()) / P.ATOL["obs"])
        for sa, sb in zip(a.get_states(idx), b.get_states(idx)):
            assert sa.cache_count == sb.cache_count and list(sa.cache_info[:sa.cache_count]) == list(sb.cache_info[:sb.cache_count])
            assert sa.steps == sb.steps and sa.spawn_count == sb.spawn_count
            for name, band in (("pos", "pos"), ("vel", "vel"), ("angvel", "angvel"), ("vbias", "vbias"), ("wbias", "wbias"), ("cache_jn", "impulse"), ("cache_jt", "impulse")):
                xa, xb = np.array(getattr(sa, name), np.float64), np.array(getattr(sb, name), np.float64)
                err = float(np.max(np.abs(xa - xb) / (P.ATOL[band] + P.RTOL * np.abs(xb))))
                worst[name] = max(worst.get(name, 0.0), err)
    P.record("cuda/class_modes_vs_general_path", {"env_steps_compared": 40 * n, "class_env_steps": seen,
                                                  "worst_violation_ratio": {k: round(v, 3) for k, v in worst.items()}})
    assert min(seen[k] for k in ("light", "pair", "multi")) > 100 and seen["heavy"] > 0, seen
    assert max(worst.values()) < 0.7, worst  # observed on the B200: 0.35 (an observation), 0.13 (an angular velocity)
