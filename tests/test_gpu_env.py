"""The Python drop-in classes over the REAL CUDA handle (-m gpu): the reference's behavioural scenarios
(soccer_simulation/test_rewards.py:127-612) and the vec-env contract (marl_vecenv.py:18-68) run through
soccerenv() / SyncMultiAgentVecEnv / TorchSoccerVecEnv / make_sharded_sim on the GPU, and the same classes are
compared step by step with their oracle-backed twins (tests/backends.py) on shared states and actions."""
import numpy as np
import pytest

import backends as B
import oracle_lib as O
import parity_util as P
import test_env_api as T
from marl_soccer_b200 import marl_vecenv, soccer_env

pytestmark = pytest.mark.gpu


def _cuda_factory(n, config, seed):
    from marl_soccer_b200.host_api import HostBufferSim
    return HostBufferSim(n, config, seed=seed)


T.FACTORIES["cuda"] = _cuda_factory


def test_product_classes_build_the_cuda_handle_by_default():
    """Without the test seam the classes construct the CUDA simulator themselves (no CPU path)."""
    from marl_soccer_b200.host_api import HostBufferSim
    env = soccer_env.soccerenv()
    assert isinstance(env._sim, HostBufferSim)
    obs, infos = env.reset(seed=1)
    assert obs["agent_0"].shape == (66,)
    env.close()
    vec = marl_vecenv.SyncMultiAgentVecEnv([soccer_env.make_env] * 3, seed=2)
    assert isinstance(vec._sim, HostBufferSim) and vec.reset(seed=0).shape == (3, 4, 66)
    vec.close()


def test_api_surface_on_cuda():
    T.test_api_surface("cuda")
    T.test_constructor_and_step_errors("cuda")
    T.test_reset_options_and_seed_determinism("cuda")
    T.test_truncation_clears_agents_and_reports_terminal_reward("cuda")


@pytest.mark.parametrize("agent_idx", [0, 1])
def test_proximity_reward_sign_on_cuda(agent_idx):
    T.test_proximity_reward_sign("cuda", agent_idx)   # test_rewards.py:139-199


def test_pushing_ball_towards_red_goal_on_cuda():
    T.test_pushing_ball_towards_red_goal_is_rewarded("cuda")   # test_rewards.py:202-251


def test_goal_scored_and_terminal_reward_on_cuda():
    T.test_goal_scored_and_terminal_reward("cuda")   # test_rewards.py:415-513


def test_own_goal_is_penalised_on_cuda():
    T.test_own_goal_is_penalised_by_shaping("cuda")   # test_rewards.py:254-363, :516-612


def test_vec_env_contract_on_cuda():
    T.test_vec_env_contract("cuda")
    T.test_vec_env_matches_single_envs_on_shared_actions("cuda")


def test_soccerenv_on_cuda_tracks_the_oracle_backed_class():
    """SoccerEnv over the kernels against SoccerEnv over the oracle: the same reset, the same scripted dict actions
    (agent_0 chases the ball and pushes it for a while: contacts with the ball), every step re-synchronised from the
    CUDA env's state (history poses included), observations / rewards / infos compared."""
    dev = T.make_env("cuda", seed=9)
    ora = T.make_env("oracle", seed=9)
    od, _ = dev.reset(seed=4, options={"use_fixed_positions": True})
    oo, _ = ora.reset(seed=4, options={"use_fixed_positions": True})
    for a in T.AGENTS:
        assert P.compare_obs(np.stack([od[b] for b in T.AGENTS]), np.stack([oo[b] for b in T.AGENTS])) <= 1.0
    worst = 0.0
    contacts = 0
    for t in range(150):
        fr = T.latest(od["agent_0"])
        act = T.zero()
        act["agent_0"] = T.toward(T.world_to_local(T.vec_from(fr, T.BALL_START), float(fr[T.ANG_IDX] * np.pi)))
        act["agent_2"] = [0.3, -0.2, 0.5]
        # re-synchronise the oracle-backed env to the device state of this step (tracked form)
        S = dev._sim.get_state(0)
        ora._sim._v.env(0).set_state(P.dev_to_oracle_state(S, np.stack([od[b] for b in T.AGENTS])))
        od, rd, td, trd, infd = dev.step(act)
        oo, ro, to, tro, info = ora.step(act)
        assert td == to and trd == tro and infd == info
        worst = max(worst, P.compare_obs(np.stack([od[b] for b in T.AGENTS]), np.stack([oo[b] for b in T.AGENTS])))
        assert abs(rd["agent_0"] - ro["agent_0"]) <= P.ATOL["reward"] + P.RTOL * abs(ro["agent_0"])
        assert rd["agent_2"] == 0.0 and rd["agent_3"] == 0.0
        contacts += ora._sim._v.env(0).contact_count()
    assert contacts > 5, "the scripted agent never reached the ball"
    assert worst <= 2.0, worst


def test_torch_vec_env_equals_numpy_vec_env():
    """TorchSoccerVecEnv (device tensors, zero copies) and SyncMultiAgentVecEnv (NumPy, host buffers) are the same
    simulator: bit-identical observations, rewards, truncations and infos, auto-resets included."""
    import torch
    n = 300
    cfg = {**P.CONFIG, "simulation": {"max_steps": 7}}
    tv = marl_vecenv.TorchSoccerVecEnv(n, config=cfg, device="cuda:0", seed=11)
    nv = marl_vecenv.SyncMultiAgentVecEnv(None, num_envs=n, config=cfg, seed=11)
    ot = tv.reset(seed=5, options={"use_full_random_positions": True})
    on = nv.reset(seed=5, options={"use_full_random_positions": True})
    assert np.array_equal(ot.cpu().numpy(), on)
    rng = np.random.default_rng(3)
    for t in range(16):
        a = rng.uniform(-1, 1, (n, 4, 3)).astype(np.float32)
        o, r, trunc, goal = tv.step(torch.from_numpy(a).cuda())
        o2, r2, term2, trunc2, infos2 = nv.step(a)
        assert np.array_equal(o.cpu().numpy(), o2)
        assert np.array_equal(tv.rewards4().cpu().numpy().astype(np.float64), r2)
        assert np.array_equal(trunc.cpu().numpy(), trunc2[:, 0]) and bool(trunc.all()) == (t % 7 == 6)
        infos = tv.infos()
        assert len(infos) == n and infos[0] == infos2[0] and infos[n - 1] == infos2[n - 1]
    st = tv.sim.stats()
    assert st["episodes"] == 2 * n and st["env_steps"] == 16 * n and st["nonfinite_actions"] == 0
    tv.close(); nv.close()


def test_vec_env_with_persistent_pinned_buffers():
    """SyncMultiAgentVecEnv(pinned_buffers=True): one set of page-locked host buffers for the lifetime of the env, step()
    returns views of them -- the same numbers as the default (fresh arrays), and the previous step's arrays are reused."""
    n = 257
    cfg = {**P.CONFIG, "simulation": {"max_steps": 5}}
    a_env = marl_vecenv.SyncMultiAgentVecEnv(None, num_envs=n, config=cfg, seed=4)
    b_env = marl_vecenv.SyncMultiAgentVecEnv(None, num_envs=n, config=cfg, seed=4, pinned_buffers=True)
    oa = a_env.reset(seed=9, options={"use_full_random_positions": True})
    ob = b_env.reset(seed=9, options={"use_full_random_positions": True})
    assert np.array_equal(oa, ob)
    rng = np.random.default_rng(1)
    first = None
    for t in range(11):
        act = rng.uniform(-1, 1, (n, 4, 3)).astype(np.float32)
        ra, rb = a_env.step(act), b_env.step(act)
        for x, y in zip(ra[:4], rb[:4]):
            assert np.array_equal(x, y)
        assert ra[4][0] == rb[4][0] and ra[4][n - 1] == rb[4][n - 1]
        if first is None:
            first = rb[0]
        else:
            assert np.shares_memory(first, rb[0])  # the same page-locked buffer every step
    a_env.close(); b_env.close()


def test_make_sharded_sim_single_process():
    """distributed.make_sharded_sim without a process group is the whole range on this GPU (rank 0 of 1), and two
    explicit shards with global offsets reproduce it (marl_vecenv.py:39-42: envs never interact)."""
    import torch
    from marl_soccer_b200.distributed import make_sharded_sim, shard_range
    from marl_soccer_b200.sim import BatchedSoccerSim
    n = 1000
    full = make_sharded_sim(n, config=P.CONFIG, seed=4)
    assert full.num_envs == n and full.global_env_offset == 0
    parts = []
    for r in range(2):
        lo, hi = shard_range(n, r, 2)
        parts.append(BatchedSoccerSim(hi - lo, config=P.CONFIG, seed=4, global_env_offset=lo))
    a = full.reset(2, seed=8).clone()
    b = torch.cat([p.reset(2, seed=8) for p in parts])
    assert torch.equal(a, b)
    g = torch.Generator(device="cuda").manual_seed(0)
    for _ in range(10):
        act = torch.rand((n, 4, 3), generator=g, device="cuda") * 2 - 1
        o, r, d, gl = full.step(act)
        outs = [p.step(act[lo:hi]) for p, (lo, hi) in zip(parts, [shard_range(n, r, 2) for r in range(2)])]
        assert torch.equal(o, torch.cat([x[0] for x in outs])) and torch.equal(r, torch.cat([x[1] for x in outs]))


def test_nonfinite_actions_are_counted_on_the_device_path():
    """soccer_env.py:116-117 raises on NaN/Inf; the device-resident path cannot raise per step: it clips (NaN acts as
    -1) and counts the offending env-steps in the statistics block."""
    import torch
    from marl_soccer_b200.sim import BatchedSoccerSim
    sim = BatchedSoccerSim(64, config=P.CONFIG, seed=0)
    sim.stats(reset=True)
    act = torch.zeros((64, 4, 3), device="cuda")
    act[3, 1, 0] = float("nan"); act[40, 2, 2] = float("inf")
    obs, *_ = sim.step(act)
    assert bool(torch.isfinite(obs).all())
    assert sim.stats()["nonfinite_actions"] == 2
