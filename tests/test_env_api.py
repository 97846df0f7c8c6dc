"""Host logic of the Python drop-ins (marl_soccer_b200.soccer_env / marl_vecenv) and the behavioural
scenarios of the reference's test_rewards.py, run without a GPU by injecting a checker-backed sim through
the `_sim_factory` test seam (tests/backends.py).  The same scenarios run on the real kernels in
test_gpu_env.py."""
import math

import numpy as np
import pytest

import backends as B
import parity_util as P
from marl_soccer_b200 import marl_vecenv, soccer_env

FACTORIES = {"oracle": B.OracleBackedSim, "hostsim": B.HostSimBacked}
AGENTS = [f"agent_{i}" for i in range(4)]
# test_rewards.py:37-51
FRAME_SIZE, STACK_SIZE = 22, 3
ANG_IDX, BALL_START, OPP_GOAL_START, OWN_GOAL_START = 2, 13, 19, 16


def make_env(kind="hostsim", **kw):
    return soccer_env.soccerenv(_sim_factory=FACTORIES[kind], _seed=kw.pop("seed", 5), **kw)


def latest(obs_vec):
    return np.asarray(obs_vec, np.float32)[(STACK_SIZE - 1) * FRAME_SIZE:]


def vec_from(frame, start):
    return frame[start:start + 2] * (frame[start + 2] * 1000.0)


def world_to_local(v, angle):
    c, s = math.cos(angle), math.sin(angle)
    return np.array([v[0] * c + v[1] * s, -v[0] * s + v[1] * c])


def zero():
    return {a: [0.0, 0.0, 0.0] for a in AGENTS}


def toward(v, mag=150000.0):
    d = v / (np.linalg.norm(v) + 1e-8)
    return [float(d[0] * mag), float(d[1] * mag), 0.0]


# ------------------------------------------------------------------------------ API surface
@pytest.mark.parametrize("kind", ["oracle", "hostsim"])
def test_api_surface(kind):
    env = make_env(kind)
    assert env.metadata == {"render_modes": ["human"], "name": "soccer_sim_v1"}
    assert env.possible_agents == AGENTS and env.agents == AGENTS
    assert env.observation_space("agent_0").shape == (66,) and env.observation_space("agent_0").dtype == np.float32
    sp = env.action_space("agent_3")
    assert sp.shape == (3,) and float(sp.low.min()) == -1.0 and float(sp.high.max()) == 1.0
    obs, infos = env.reset(seed=3)
    assert set(obs) == set(AGENTS) and all(o.shape == (66,) and o.dtype == np.float32 for o in obs.values())
    assert infos == {a: {} for a in AGENTS}
    obs2, rew, term, trunc, infos = env.step({a: env.action_space(a).sample() for a in AGENTS})
    assert set(rew) == set(AGENTS) and rew["agent_2"] == 0.0 and rew["agent_3"] == 0.0 and rew["agent_0"] == rew["agent_1"]
    assert all(isinstance(r, float) for r in rew.values())
    assert term == {a: False for a in AGENTS} and trunc == {a: False for a in AGENTS}
    assert infos["agent_0"]["score"] == {"blue": 0, "red": 0} and infos["agent_1"] == infos["agent_3"]
    # newest frame last: frames 0,1 of the new obs are frames 1,2 of the old one
    for a in AGENTS:
        assert np.array_equal(obs2[a][:44], obs[a][22:])
    sc = soccer_env.get_observation_scalers(env)
    assert sc == {"max_velocity": 200.0, "max_angular_velocity": 10.0, "field_diagonal": 1000.0, "stack_size": 3, "frame_size": 22}
    assert soccer_env.make_env(_sim_factory=FACTORIES[kind]).possible_agents == AGENTS


def test_constructor_and_step_errors(kind="hostsim"):
    with pytest.raises(ValueError):
        soccer_env.SoccerEnv(env=2, _sim_factory=FACTORIES[kind])
    with pytest.raises(ValueError):
        soccer_env.SoccerEnv(num_envs=8, _sim_factory=FACTORIES[kind])
    soccer_env.SoccerEnv(num_envs=1, env=1, _sim_factory=FACTORIES[kind])  # allowed values are ignored
    env = make_env(kind)
    env.reset()
    with pytest.raises(ValueError, match="Missing actions"):
        env.step({a: [0, 0, 0] for a in AGENTS[:3]})
    with pytest.raises(ValueError, match="unknown agents"):
        env.step({**zero(), "agent_9": [0, 0, 0]})
    with pytest.raises(ValueError, match="shape"):
        env.step({**zero(), "agent_1": [0, 0]})
    with pytest.raises(ValueError, match="non-finite"):
        env.step({**zero(), "agent_2": [0.0, float("nan"), 0.0]})


def test_reset_options_and_seed_determinism(kind="hostsim"):
    env = make_env(kind)
    o1, _ = env.reset(seed=11, options={"use_full_random_positions": True})
    o2, _ = env.reset(seed=11, options={"use_full_random_positions": True})
    o3, _ = env.reset(seed=12, options={"use_full_random_positions": True})
    assert all(np.array_equal(o1[a], o2[a]) for a in AGENTS) and not all(np.array_equal(o1[a], o3[a]) for a in AGENTS)
    of, _ = env.reset(options={"use_fixed_positions": True})
    assert np.allclose(latest(of["agent_0"])[4:7], [0.0, 1.0, 0.198], atol=1e-6)


def test_truncation_clears_agents_and_reports_terminal_reward(kind="hostsim"):
    cfg = {**P.CONFIG, "rewards": {**P.CONFIG["rewards"], "score_difference_multiplier": 5.0}, "simulation": {"max_steps": 4}}
    env = make_env(kind, config=cfg)
    env.reset(seed=1)
    for t in range(4):
        obs, rew, term, trunc, infos = env.step(zero())
        assert trunc["agent_0"] == (t == 3)
    assert env.agents == [] and rew["agent_0"] == 0.0 and not any(term.values())
    env.reset()
    assert env.agents == AGENTS


# ------------------------------------------------------------- behavioural scenarios (test_rewards.py)
@pytest.mark.parametrize("kind", ["oracle", "hostsim"])
@pytest.mark.parametrize("agent_idx", [0, 1])
def test_proximity_reward_sign(kind, agent_idx):
    """test_rewards.py:139-199: 6 steps toward the ball beat the idle baseline, 6 steps away lose to it."""
    for sign in (+1, -1):
        env = make_env(kind)
        env.reset(seed=21)
        obs, rew, *_ = env.step(zero())
        base = rew[f"agent_{agent_idx}"]
        total = 0.0
        for _ in range(6):
            fr = latest(obs[f"agent_{agent_idx}"])
            ang = float(fr[ANG_IDX] * math.pi)
            act = zero()
            act[f"agent_{agent_idx}"] = toward(world_to_local(sign * vec_from(fr, BALL_START), ang))
            obs, rew, *_ = env.step(act)
            total += rew[f"agent_{agent_idx}"]
        assert sign * (total - 6 * base) > 0.0


def approach_ball(env, obs, agent, limit=35.0, max_steps=120):
    total = 0.0
    for _ in range(max_steps):
        fr = latest(obs[agent])
        bv = vec_from(fr, BALL_START)
        if np.linalg.norm(bv) < limit:
            break
        act = zero()
        act[agent] = toward(world_to_local(bv, float(fr[ANG_IDX] * math.pi)))
        obs, rew, term, trunc, infos = env.step(act)
        assert not trunc[agent] and "goal_scored_by" not in infos[agent]
        total += rew["agent_0"] + rew["agent_1"]
    return obs, total


@pytest.mark.parametrize("kind", ["oracle", "hostsim"])
def test_pushing_ball_towards_red_goal_is_rewarded(kind):
    """test_rewards.py:202-251"""
    env = make_env(kind)
    obs, _ = env.reset(seed=33)
    obs, total = approach_ball(env, obs, "agent_0")
    for _ in range(5):
        fr = latest(obs["agent_0"])
        act = zero()
        act["agent_0"] = toward(world_to_local(vec_from(fr, OPP_GOAL_START), float(fr[ANG_IDX] * math.pi)))
        obs, rew, term, trunc, infos = env.step(act)
        total += rew["agent_0"] + rew["agent_1"]
    assert total > 0.0


def drive_ball_into_goal(env, agent, push_dir):
    """Scripted blue agent: get behind the ball (on the side opposite to push_dir), line up with the line
    goal-centre -> ball, then push.  Velocity-tracking control from the observation only (own velocity =
    frame[0:2] * max_velocity), forces expressed in the body frame like test_rewards.py:111-119."""
    obs, _ = env.reset(options={"use_fixed_positions": True})
    total, scored, infos = 0.0, None, {}
    goal_slot = OPP_GOAL_START if push_dir > 0 else OWN_GOAL_START
    for step in range(990):
        fr = latest(obs[agent])
        ang = float(fr[ANG_IDX] * math.pi)
        vel = fr[0:2] * 200.0
        bv = vec_from(fr, BALL_START)                 # agent -> ball
        gv = vec_from(fr, goal_slot)                  # agent -> target goal centre
        u = (gv - bv) / (np.linalg.norm(gv - bv) + 1e-8)  # ball -> goal
        station = bv - u * 34.0                       # where to stand: behind the ball on the goal line
        along = float(np.dot(-bv, u))                 # agent's position relative to the ball along u
        lateral = float(u[0] * -bv[1] - u[1] * -bv[0])
        if along > -20.0:                             # beside / in front of the ball: go around it
            side = 1.0 if lateral >= 0 else -1.0
            tgt = bv + np.array([-u[1], u[0]]) * side * 45.0 - u * 50.0
            vdes = tgt / (np.linalg.norm(tgt) + 1e-8) * 150.0
        elif np.linalg.norm(station) > 8.0 and abs(lateral) > 5.0:
            vdes = np.clip(station * 4.0, -120.0, 120.0)
        else:
            vdes = u * 190.0
        f_world = (vdes - vel) * 10.0 * 60.0 * 0.5    # half of the force that would reach vdes in one tick
        n = np.linalg.norm(f_world)
        if n > 150000.0:
            f_world *= 150000.0 / n
        fl = world_to_local(f_world, ang)
        act = zero()
        act[agent] = [float(fl[0] / 150000.0), float(fl[1] / 150000.0), float(np.clip(-fr[3] * 2.0, -1, 1))]
        obs, rew, term, trunc, infos = env.step(act)
        total += rew["agent_0"] + rew["agent_1"]
        if "goal_scored_by" in infos[agent]:
            scored = infos[agent]["goal_scored_by"]
            break
        assert not trunc[agent]
    return obs, total, scored, infos


@pytest.mark.parametrize("kind", ["oracle", "hostsim"])
def test_goal_scored_and_terminal_reward(kind):
    """test_rewards.py:415-513: blue scores; after idling to the end of the episode the terminal step's
    reward is score_difference_multiplier * (blue - red) for both blue agents."""
    cfg = {**P.CONFIG, "rewards": {**P.CONFIG["rewards"], "score_difference_multiplier": 5.0}}
    env = make_env(kind, config=cfg)
    obs, total, scored, infos = drive_ball_into_goal(env, "agent_0", +1)
    assert scored == "blue" and infos["agent_0"]["score"] == {"blue": 1, "red": 0}
    last = None
    for _ in range(1000):
        obs, rew, term, trunc, infos = env.step(zero())
        last = rew
        if trunc["agent_0"]:
            break
    assert trunc["agent_0"] and env.agents == []
    sc = infos["agent_0"]["score"]
    assert abs((last["agent_0"] + last["agent_1"]) - 2 * 5.0 * (sc["blue"] - sc["red"])) <= 0.5


@pytest.mark.parametrize("kind", ["oracle", "hostsim"])
def test_own_goal_is_penalised_by_shaping(kind):
    """test_rewards.py:254-363 and :516-612: pushing the ball into the own (blue) goal concedes a goal and
    yields a negative cumulative reward during the push."""
    env = make_env(kind)
    obs, total, scored, infos = drive_ball_into_goal(env, "agent_1", -1)
    assert scored == "red" and infos["agent_1"]["score"]["red"] == 1
    assert total < 0.0


# ------------------------------------------------------------------------------ vec env
def test_vec_env_contract(kind="hostsim"):
    n = 6
    cfg = {**P.CONFIG, "simulation": {"max_steps": 5}}
    vec = marl_vecenv.SyncMultiAgentVecEnv([lambda: make_env(kind, config=cfg)] * n, seed=7, _sim_factory=FACTORIES[kind])
    assert vec.num_envs == n and len(vec.envs) == n and vec.possible_agents == AGENTS
    assert vec.single_observation_space.shape == (66,) and vec.single_action_space.shape == (3,)
    obs = vec.reset(seed=1)
    assert isinstance(obs, np.ndarray) and obs.shape == (n, 4, 66) and obs.dtype == np.float32
    rng = np.random.default_rng(0)
    for t in range(5):
        o, r, term, trunc, infos = vec.step(rng.uniform(-1, 1, (n, 4, 3)))
        assert o.shape == (n, 4, 66) and r.shape == (n, 4) and r.dtype == np.float64
        assert term.shape == (n, 4) and term.dtype == bool and not term.any()
        assert trunc.shape == (n, 4) and trunc.dtype == bool and trunc.all() == (t == 4)
        assert (r[:, 2:] == 0).all() and np.array_equal(r[:, 0], r[:, 1])
        assert len(infos) == n and set(infos[0]) == set(AGENTS) and "score" in infos[0]["agent_0"]
        assert [i["agent_2"]["score"] for i in infos][0] == infos[0]["agent_0"]["score"]
    # auto-reset: the observation returned on the truncation step is the next episode's first one
    f = o.reshape(n, 4, 3, 22)
    assert np.array_equal(f[:, :, 0], f[:, :, 1]) and np.array_equal(f[:, :, 1], f[:, :, 2])
    o2, r2, term2, trunc2, _ = vec.step(np.zeros((n, 4, 3), np.float32))
    assert not trunc2.any()
    with pytest.raises(ValueError):
        vec.step(np.zeros((n, 4, 2)))
    bad = np.zeros((n, 4, 3)); bad[2, 1, 0] = np.inf
    with pytest.raises(ValueError):
        vec.step(bad)
    vec.close()


def test_vec_env_matches_single_envs_on_shared_actions(kind="hostsim"):
    """SyncMultiAgentVecEnv == the per-env loop of the reference (marl_vecenv.py:39-53) when the per-env
    states are the same: run the batched sim and n single-env sims from identical injected states."""
    n = 5
    vec = marl_vecenv.SyncMultiAgentVecEnv(None, num_envs=n, config=P.CONFIG, seed=3, _sim_factory=FACTORIES[kind])
    vec.reset(seed=40, options={"use_full_random_positions": True})
    singles = [make_env(kind, seed=3) for _ in range(n)]
    for i, e in enumerate(singles):
        e.reset()
        e._sim.set_state(0, vec._sim.get_state(i))  # the history poses travel in the state
    rng = np.random.default_rng(1)
    for t in range(30):
        a = rng.uniform(-1, 1, (n, 4, 3)).astype(np.float32)
        o, r, term, trunc, infos = vec.step(a)
        for i, e in enumerate(singles):
            oo, rr, tt, tr, inf = e.step(vec._array_to_dict(a[i]))
            assert np.array_equal(vec._dict_to_array(oo), o[i])
            assert rr["agent_0"] == r[i, 0] and inf["agent_0"]["score"] == infos[i]["agent_0"]["score"]


# ------------------------------------------------------------------------------ render pull-back
def test_render_pulls_back_one_env_and_builds_the_scene():
    """render() (soccer_env.py:156-162): None unless render_mode == "human"; otherwise the five poses of the env, and
    the scene primitives flip the y axis like the reference's drawing code (renderer.py:30-42, entities.py:37-57)."""
    from marl_soccer_b200 import renderer
    assert make_env().render() is None
    env = make_env(render_mode="human")
    env.reset(options={"use_fixed_positions": True})
    poses = env.render()
    assert [p for p, _ in poses["agents"]] == [(200.0, 198.0), (200.0, 396.0), (600.0, 198.0), (600.0, 396.0)]
    assert poses["ball"] == (400.0, 300.0) and abs(poses["agents"][2][1] - math.pi) < 1e-6
    prims = renderer.scene(poses)
    polys = [p for p in prims if p[0] == "poly"]
    assert len(polys) == 8 and polys[0][1] == renderer.BLUE_RGB and polys[4][1] == renderer.RED_RGB
    # agent_0 at (200, 198), angle 0: the box is 30 x 30 around (200, 600 - 198) on the screen, its nose points to +x
    xs, ys = [q[0] for q in polys[0][2]], [q[1] for q in polys[0][2]]
    assert (min(xs), max(xs), min(ys), max(ys)) == (185.0, 215.0, 387.0, 417.0)
    assert max(q[0] for q in polys[1][2]) == 215.0
    # agent_2 faces -x (angle pi): its nose tip is on the left edge
    assert abs(min(q[0] for q in polys[5][2]) - 585.0) < 1e-4
    ball = prims[-1]
    assert ball[0] == "circle" and ball[2] == (400.0, 300.0) and ball[3] == 10
    env.close()
