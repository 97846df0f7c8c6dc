"""Shared helpers of the parity tests: state conversion between the oracle (fp64, tests/oracle_lib.py)
and the device format (include/msoc.h msoc_env_state), seeded scenario generators, tolerant
comparison, and DeviceSim = the product's C-ABI driven with NumPy host buffers.

Tolerances (north-star: goals/dones/steps/reset indices bit-exact; states and rewards within 1e-5
relative, fp32 vs the oracle's fp64): every comparison is |a-b| <= atol + 1e-5*|b| with an explicit
per-quantity atol that reflects one fp32 ulp at the quantity's natural scale (positions <= 800 px ->
ulp 6.1e-5; velocities <= ~450 px/s -> ulp 3.1e-5)."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

import oracle_lib as O
from marl_soccer_b200 import _capi

RTOL = 1e-5
ATOL = {
    "pos": 2.5e-4,     # px: 4 ulp at 800
    "vel": 4e-3,       # px/s: 1e-5 * 400 (post-impact relative speeds reach 2 * max_velocity)
    "ang": 2e-5,       # rad, compared modulo 2 pi
    "angvel": 5e-4,    # rad/s: 1e-5 * 50 (the light boxes, moment 100, spin up to ~1e2 rad/s on impact)
    "vbias": 4e-3,     # px/s, same scale as vel
    "wbias": 5e-4,
    "reward": 2e-6,    # alive penalty is 1e-5; shaping terms are O(1e-2..3e-1)
    "obs": 2e-5,       # unit vectors / normalised magnitudes, O(1)
    "impulse": 2e-2,   # cached jnAcc/jtAcc, O(1e2..2e3): 1e-5 * 2000
}

CONFIG = O.DEFAULT_CONFIG


def wrap(a):
    return np.arctan2(np.sin(a), np.cos(a))


def dev_to_oracle_state(S: _capi.MsocEnvState, obs: np.ndarray) -> dict:
    cache = []
    for j in range(int(S.cache_count)):
        info = int(S.cache_info[j])
        cache.append((info & 63, (info >> 6) & 15, (info >> 10) & 3, float(S.cache_jn[j]), float(S.cache_jt[j])))
    return {
        "pos": np.array(S.pos, np.float64), "vel": np.array(S.vel, np.float64),
        "ang": np.array(list(S.ang) + [0.0], np.float64), "angvel": np.array(S.angvel, np.float64),
        "vbias": np.array(S.vbias, np.float64), "wbias": np.array(list(S.wbias) + [0.0], np.float64),
        "steps": int(S.steps), "score": (int(S.score[0]), int(S.score[1])), "mode": int(S.mode),
        "spawn_count": int(S.spawn_count), "seed": int(S.seed), "obs": np.asarray(obs, np.float32).reshape(4, 66),
        "cache": cache,
    }


def oracle_to_dev_state(d: dict) -> _capi.MsocEnvState:
    S = _capi.MsocEnvState()
    for i in range(5):
        for k in range(2):
            S.pos[i][k] = float(d["pos"][i][k])
            S.vel[i][k] = float(d["vel"][i][k])
            S.vbias[i][k] = float(d["vbias"][i][k]) if "vbias" in d else 0.0
        S.angvel[i] = float(d["angvel"][i])
    for i in range(4):
        S.ang[i] = float(wrap(float(d["ang"][i])))
        S.wbias[i] = float(d["wbias"][i]) if "wbias" in d else 0.0
    S.ep_return = float(d.get("ep_return", 0.0))
    S.steps = int(d.get("steps", 0))
    S.score[0], S.score[1] = (int(x) for x in d.get("score", (0, 0)))
    S.mode = int(d.get("mode", 0))
    S.spawn_count = int(d.get("spawn_count", 0))
    S.seed = int(d.get("seed", 0))
    cache = d.get("cache", [])[: _capi.MAX_CACHE]
    S.cache_count = len(cache)
    for j, (p, key, age, jn, jt) in enumerate(cache):
        S.cache_info[j] = int(p) | (int(key) << 6) | (int(age) << 10)
        S.cache_jn[j] = float(jn)
        S.cache_jt[j] = float(jt)
    return S


def f32(x):
    return np.asarray(x, np.float32).astype(np.float64)


def random_state(rng: np.random.Generator, kind: str = "open") -> dict:
    """One fp32-representable env state.  kind: 'open' (SURVEY section 8d config-2 recipe: uniform over the
    field box, velocities in the disc <= 200, angles U(-pi,pi), ang-vel U(-10,10), steps U{0..998}),
    'walls' (agents and ball hugging walls / corners / goal mouths), 'scrum' (bodies packed around the
    ball), 'goal' (ball about to cross a goal line)."""
    pos = np.zeros((5, 2))
    if kind == "open":
        pos[:, 0] = rng.uniform(30, 770, 5)
        pos[:, 1] = rng.uniform(30, 570, 5)
    elif kind == "walls":
        for i in range(5):
            side = rng.integers(0, 6)
            r = 10.0 if i == 4 else 15.0
            d = rng.uniform(-1.5, 6.0) + r + 2.0  # centre distance to the wall core line
            if side == 0:
                pos[i] = (10 + d, rng.uniform(30, 570))
            elif side == 1:
                pos[i] = (790 - d, rng.uniform(30, 570))
            elif side == 2:
                pos[i] = (rng.uniform(30, 770), 10 + d)
            elif side == 3:
                pos[i] = (rng.uniform(30, 770), 590 - d)
            elif side == 4:  # corner
                cx = 10 + d if rng.random() < 0.5 else 790 - d
                d2 = rng.uniform(-1.5, 6.0) + r + 2.0
                cy = 10 + d2 if rng.random() < 0.5 else 590 - d2
                pos[i] = (cx, cy)
            else:  # goal mouth / posts
                x = 10 + rng.uniform(-2, 25) if rng.random() < 0.5 else 790 - rng.uniform(-2, 25)
                pos[i] = (x, rng.uniform(200, 400))
    elif kind == "scrum":
        c = np.array([rng.uniform(100, 700), rng.uniform(100, 500)])
        pos[4] = c
        for i in range(4):
            ang = rng.uniform(-math.pi, math.pi)
            pos[i] = c + rng.uniform(18, 40) * np.array([math.cos(ang), math.sin(ang)])
    elif kind == "goal":
        pos[:4, 0] = rng.uniform(100, 700, 4)
        pos[:4, 1] = rng.uniform(30, 570, 4)
        left = rng.random() < 0.5
        pos[4] = ((10 + rng.uniform(0.5, 3.0)) if left else (790 - rng.uniform(0.5, 3.0)), rng.uniform(240, 360))
    else:
        raise ValueError(kind)
    vel = np.zeros((5, 2))
    for i in range(5):
        sp = 200.0 * math.sqrt(rng.random())
        th = rng.uniform(-math.pi, math.pi)
        vel[i] = (sp * math.cos(th), sp * math.sin(th))
    if kind == "goal":
        vel[4] = ((-1 if pos[4, 0] < 400 else 1) * rng.uniform(150, 200), rng.uniform(-20, 20))
    ang = np.concatenate([rng.uniform(-math.pi, math.pi, 4), [0.0]])
    if rng.random() < 0.25:  # axis-aligned agents (the spawn orientation), the degenerate clipping case
        ang[:4] = np.array([0.0, 0.0, math.pi, math.pi])[:4]
    angvel = rng.uniform(-10, 10, 5)
    d = {
        "pos": f32(pos), "vel": f32(vel), "ang": f32(ang), "angvel": f32(angvel),
        "vbias": np.zeros((5, 2)), "wbias": np.zeros(5),
        "steps": int(rng.integers(0, 999)), "score": (int(rng.integers(0, 3)), int(rng.integers(0, 3))),
        "mode": int(rng.integers(0, 3)), "spawn_count": int(rng.integers(0, 5)), "seed": int(rng.integers(0, 2**31)),
        "obs": f32(rng.uniform(-1, 1, (4, 66))).astype(np.float32), "cache": [],
    }
    return d


def close(a, b, atol, rtol=RTOL):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b) <= atol + rtol * np.abs(b)


def ang_diff(a, b):
    return np.abs(wrap(np.asarray(a, np.float64) - np.asarray(b, np.float64)))


def compare_state(dev: dict, ora: dict) -> dict:
    """Max violation ratio err/(atol+rtol*|ref|) per quantity (<= 1 passes) and raw max abs error."""
    out = {}
    for key in ("pos", "vel", "angvel", "vbias"):
        a, b = np.asarray(dev[key], np.float64), np.asarray(ora[key], np.float64)
        out[key] = float(np.max(np.abs(a - b) / (ATOL[key] + RTOL * np.abs(b))))
    out["ang"] = float(np.max(ang_diff(dev["ang"][:4], ora["ang"][:4]) / ATOL["ang"]))
    a, b = np.asarray(dev["wbias"][:4], np.float64), np.asarray(ora["wbias"][:4], np.float64)
    out["wbias"] = float(np.max(np.abs(a - b) / (ATOL["wbias"] + RTOL * np.abs(b))))
    return out


def compare_obs(dev_obs, ora_obs) -> float:
    """Observation rows (4,66); the angle feature (index 2 of every frame) is compared modulo 2
    (+1 and -1 are the same angle: float32(pi) wraps to either side of the atan2 branch cut)."""
    a = np.asarray(dev_obs, np.float64).reshape(4, 3, 22).copy()
    b = np.asarray(ora_obs, np.float64).reshape(4, 3, 22).copy()
    da = np.abs(wrap((a[:, :, 2] - b[:, :, 2]) * math.pi)) / math.pi
    a[:, :, 2] = 0.0
    b[:, :, 2] = 0.0
    atol = np.full(22, ATOL["obs"])
    atol[0:2] = max(ATOL["obs"], ATOL["vel"] / 200.0)   # velocity / max_velocity
    atol[3] = max(ATOL["obs"], ATOL["angvel"] / 10.0)   # angular velocity / max_angular_velocity
    v = np.max(np.abs(a - b) / (atol + RTOL * np.abs(b)))
    return float(max(v, np.max(da) / ATOL["obs"]))


def cache_dict(cache):
    return {(p, k): (age, jn, jt) for (p, k, age, jn, jt) in cache}


def compare_cache(dev_cache, ora_cache) -> tuple[bool, float]:
    """Same (pair, key, age) sets bit-exactly; impulses within tolerance."""
    d, o = cache_dict(dev_cache), cache_dict(ora_cache)
    same = set(d) == set(o) and all(d[k][0] == o[k][0] for k in d)
    worst = 0.0
    if same:
        for k in d:
            for x, y in ((d[k][1], o[k][1]), (d[k][2], o[k][2])):
                worst = max(worst, abs(x - y) / (ATOL["impulse"] + RTOL * abs(y)))
    return same, worst


class DeviceSim:
    """The product's C-ABI (include/msoc.h) with NumPy host buffers: msoc_create / msoc_reset_host /
    msoc_step_host / msoc_get_state / msoc_set_state.  Needs a CUDA device."""

    name = "device"

    def __init__(self, n: int, config: dict, seed: int = 0, global_offset: int = 0, device: int = 0):
        self._L = _capi.lib()
        self.n = int(n)
        self._cfg = _capi.make_config(config)
        h = C.c_void_p()
        _capi.check(self._L.msoc_create(C.byref(self._cfg), self.n, device, seed, global_offset, C.byref(h)))
        self._h = h

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.msoc_destroy(self._h)
            self._h = None

    def reset(self, mode: int = 0, seed: int | None = None, mask=None) -> np.ndarray:
        obs = np.zeros((self.n, 4, 66), np.float32)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        _capi.check(self._L.msoc_reset_host(self._h, None if m is None else m.ctypes.data, mode,
                                            0 if seed is None else 1, 0 if seed is None else int(seed),
                                            obs.ctypes.data, None))
        return obs

    def step(self, actions, auto_reset: bool = True):
        a = np.ascontiguousarray(actions, dtype=np.float32).reshape(self.n, 12)
        obs = np.zeros((self.n, 4, 66), np.float32)
        rew = np.zeros((self.n, 2), np.float32)
        done = np.zeros(self.n, np.uint8)
        goal = np.zeros(self.n, np.int8)
        self.score = np.zeros((self.n, 2), np.int32)
        _capi.check(self._L.msoc_step_host(self._h, a.ctypes.data, obs.ctypes.data, rew.ctypes.data,
                                           done.ctypes.data, goal.ctypes.data, self.score.ctypes.data,
                                           1 if auto_reset else 0, None))
        return obs, rew, done, goal

    def get_states(self, idx) -> list:
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        arr = (_capi.MsocEnvState * len(idx))()
        _capi.check(self._L.msoc_get_state(self._h, idx.ctypes.data, len(idx), C.byref(arr)))
        return list(arr)

    def set_states(self, idx, states, obs=None) -> None:
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        arr = (_capi.MsocEnvState * len(idx))(*states)
        _capi.check(self._L.msoc_set_state(self._h, idx.ctypes.data, len(idx), C.byref(arr)))
        if obs is not None:
            o = np.ascontiguousarray(obs, dtype=np.float32).reshape(len(idx), 4, 66)
            _capi.check(self._L.msoc_set_obs_host(self._h, idx.ctypes.data, len(idx), o.ctypes.data))

    def get_state(self, i: int):
        return self.get_states([i])[0]

    def set_state(self, i: int, S, obs=None) -> None:
        self.set_states([i], [S], None if obs is None else np.asarray(obs, np.float32)[None])

    def get_obs(self, i: int) -> np.ndarray:
        idx = np.array([i], np.int64)
        o = np.zeros((1, 4, 66), np.float32)
        _capi.check(self._L.msoc_get_obs_host(self._h, idx.ctypes.data, 1, o.ctypes.data))
        return o[0]

    def stats(self, reset: bool = False) -> dict:
        s = _capi.MsocStats()
        _capi.check(self._L.msoc_stats_read(self._h, C.byref(s), 1 if reset else 0, None))
        return {k: getattr(s, k) for k, _ in _capi.MsocStats._fields_}


def add_batch_api(cls):
    """HostSim gets the list-based get_states/set_states of DeviceSim."""
    if not hasattr(cls, "get_states"):
        cls.get_states = lambda self, idx: [self.get_state(int(i)) for i in idx]
        cls.set_states = lambda self, idx, states, obs=None: [
            self.set_state(int(i), s, None if obs is None else obs[k]) for k, (i, s) in enumerate(zip(idx, states))]
    return cls


# ----------------------------------------------------------------------------------------- checks
KINDS = ("open", "walls", "scrum", "goal")


def inject(sim, ora, states):
    n = len(states)
    sim.set_states(np.arange(n), [oracle_to_dev_state(s) for s in states],
                   np.stack([s["obs"] for s in states]))
    ora.set_states(states)


def compare_all(sim, ora, obs_d, obs_o, rew_d, rew_o, n, label=""):
    """Per-env worst violation ratio of every quantity after a step; returns (worst dict, list of
    failing env indices, cache mismatches)."""
    worst, failing, cache_bad = {}, [], []
    dev_states = sim.get_states(np.arange(n))
    for i in range(n):
        sd = dev_to_oracle_state(dev_states[i], obs_d[i])
        so = ora.env(i).get_state()
        c = compare_state(sd, so)
        c["obs"] = compare_obs(obs_d[i], obs_o[i])
        c["reward"] = abs(float(rew_d[i, 0]) - float(rew_o[i, 0])) / (ATOL["reward"] + RTOL * abs(float(rew_o[i, 0])))
        same, w = compare_cache(sd["cache"], so["cache"])
        c["impulse"] = w
        if not same:
            cache_bad.append(i)
        exact = (sd["steps"] == so["steps"] and sd["score"] == so["score"] and sd["mode"] == so["mode"]
                 and sd["spawn_count"] == so["spawn_count"] and sd["seed"] == so["seed"])
        if not exact:
            cache_bad.append(i)
        if max(c.values()) > 1.0:
            failing.append((i, {k: round(v, 2) for k, v in c.items() if v > 1.0}))
        for k, v in c.items():
            worst[k] = max(worst.get(k, 0.0), v)
    return worst, failing, cache_bad


def check_single_step(sim_cls, n: int, seed: int, **kw):
    """Inject n seeded states (all four scenario kinds) into both sides, one step with out-of-range
    actions (clipping is exercised), compare everything."""
    rng = np.random.default_rng(seed)
    sim = sim_cls(n, CONFIG, seed=0, **kw)
    ora = O.OracleVec(n, CONFIG, seed=0)
    states = [random_state(rng, KINDS[i % 4]) for i in range(n)]
    inject(sim, ora, states)
    act = rng.uniform(-1.2, 1.2, (n, 4, 3)).astype(np.float32)
    o_d, r_d, d_d, g_d = sim.step(act, auto_reset=False)
    o_o, r_o, d_o, g_o = ora.step(act, auto_reset=False)
    assert np.array_equal(d_d, d_o), "done flags differ"
    assert np.array_equal(g_d, g_o), "goal flags differ"
    assert np.array_equal(r_d[:, 0], r_d[:, 1]), "blue rewards must be identical (game/game.py:324-375)"
    worst, failing, cache_bad = compare_all(sim, ora, o_d, o_o, r_d, r_o, n)
    assert not cache_bad, f"arbiter cache / counters differ for envs {cache_bad[:10]}"
    # Multi-contact Gauss-Seidel solves on the injected deep-overlap states amplify fp32 rounding a little
    # beyond the per-quantity band: allow <= 0.5 % of the envs to exceed it, and none by more than 5x.
    assert len(failing) <= max(1, n // 200), f"{len(failing)} envs out of tolerance, e.g. {failing[:5]}; worst {worst}"
    assert max(worst.values()) < 5.0, worst
    return worst, int(np.abs(g_o).sum())


def check_tracked_rollout(sim_cls, n: int, steps: int, seed: int, mode: int = 2, **kw):
    """`steps`-step rollout from a shared reset; after every step the oracle is re-synchronised to the
    device state, so each of the steps is an independent single-step parity check on states the sim
    itself reaches (spawn overlap with walls, resting contacts, warm-started arbiters, goals, truncation,
    auto-reset).  Returns the number of (env, step) pairs that were out of tolerance and the total."""
    rng = np.random.default_rng(seed)
    sim = sim_cls(n, CONFIG, seed=seed, **kw)
    ora = O.OracleVec(n, CONFIG, seed=seed)
    o_d = sim.reset(mode, seed=seed)
    o_o = ora.reset(mode, seed=seed)
    assert np.allclose(o_d, o_o, atol=ATOL["obs"]), "reset observations differ"
    # start late in the episode for half of the envs so that truncation + auto-reset are exercised
    st = sim.get_states(np.arange(n))
    for i in range(0, n, 2):
        st[i].steps = CONFIG["simulation"]["max_steps"] - 1 - (i % max(steps, 1))
    sim.set_states(np.arange(n), st)
    out_of_tol, total, events = 0, 0, {"goals": 0, "dones": 0, "contacts": 0}
    worst_all = {}
    for t in range(steps):
        dev_states = sim.get_states(np.arange(n))
        ora.set_states([dev_to_oracle_state(dev_states[i], o_d[i]) for i in range(n)])
        act = rng.uniform(-1.0, 1.0, (n, 4, 3)).astype(np.float32)
        o_d, r_d, d_d, g_d = sim.step(act, auto_reset=True)
        o_o, r_o, d_o, g_o = ora.step(act, auto_reset=True)
        assert np.array_equal(d_d, d_o), f"done flags differ at step {t}"
        assert np.array_equal(g_d, g_o), f"goal flags differ at step {t}"
        worst, failing, cache_bad = compare_all(sim, ora, o_d, o_o, r_d, r_o, n)
        assert not cache_bad, f"step {t}: arbiter cache / counters differ for envs {cache_bad[:10]}"
        out_of_tol += len(failing)
        total += n
        events["goals"] += int(np.abs(g_o).sum())
        events["dones"] += int(d_o.sum())
        events["contacts"] += sum(1 for i in range(n) if ora.env(i).contact_count() > 0)
        for k, v in worst.items():
            worst_all[k] = max(worst_all.get(k, 0.0), v)
    return out_of_tol, total, events, worst_all
