"""Device-resident rollout half of the reference's PPO trainer (marl_soccer_b200/rollout.py):
network / normaliser / GAE arithmetic against plain restatements of marl-soccer.ipynb on the CPU, and the
in-loop rollout on the simulator on the GPU."""
import numpy as np
import pytest
import torch

from marl_soccer_b200.rollout import Agent, PackedPolicy, RolloutBuffer, RunningMeanStd, collect_rollout, compute_gae


def test_agent_architecture_matches_reference_checkpoint_layout():
    """marl-soccer.ipynb:125-190: 66 -> 512 -> 256 -> 128 -> 64 -> (3 | 1), tanh, state-independent log-std; the
    parameter names are the reference's, so its `.ppo_model` state dicts load unchanged."""
    a = Agent()
    sd = a.state_dict()
    assert sd["actor_logstd"].shape == (1, 3)
    for net, out in (("critic", 1), ("actor_mean", 3)):
        dims = [(512, 66), (256, 512), (128, 256), (64, 128), (out, 64)]
        for k, d in zip((0, 2, 4, 6, 8), dims):
            assert tuple(sd[f"{net}.{k}.weight"].shape) == d
            assert tuple(sd[f"{net}.{k}.bias"].shape) == (d[0],)
    x = torch.randn(7, 66)
    act, logp, ent, val = a.get_action_and_value(x)
    assert act.shape == (7, 3) and logp.shape == (7,) and ent.shape == (7,) and val.shape == (7, 1)
    # the log-prob of a given action is the diagonal Gaussian's
    mean = a.actor_mean(x)
    ref = torch.distributions.Normal(mean, torch.ones_like(mean)).log_prob(act).sum(1)
    _, logp2, _, _ = a.get_action_and_value(x, act)
    assert torch.allclose(logp2, ref, atol=1e-6)


def test_packed_policy_is_the_agent():
    """PackedPolicy (first layers of both nets as one GEMM over 72-padded inputs, hidden layers batched over the two nets,
    outputs padded to 8) computes the Agent's means, values and log-probabilities; refresh() follows parameter updates."""
    torch.manual_seed(3)
    a = Agent()
    with torch.no_grad():
        for p in a.parameters():
            p.add_(0.05 * torch.randn_like(p))
    pp = PackedPolicy(a, torch.float32)
    x = torch.randn(41, 66)
    x72 = torch.zeros(41, PackedPolicy.IN)
    x72[:, :66] = x
    mean, value = pp(x72)
    assert torch.allclose(mean, a.actor_mean(x), atol=1e-5) and torch.allclose(value, a.critic(x), atol=1e-5)
    action, logprob, value2 = pp.act(x72)
    _, ref_logprob, _, ref_value = a.get_action_and_value(x, action)
    assert torch.allclose(logprob, ref_logprob, atol=1e-4) and torch.allclose(value2, ref_value, atol=1e-5)
    with torch.no_grad():
        a.actor_mean[0].weight.mul_(2.0)
        a.actor_logstd.fill_(-0.5)
    pp.refresh()
    assert torch.allclose(pp(x72)[0], a.actor_mean(x), atol=1e-5)
    action, logprob, _ = pp.act(x72)
    assert torch.allclose(logprob, a.get_action_and_value(x, action)[1], atol=1e-4)


def test_running_mean_std_matches_batch_statistics():
    """marl-soccer.ipynb:264-296: parallel-variance update; after all batches mean/var equal those of the
    concatenated stream (up to the unit prior count the reference does not use: count starts at 0 here)."""
    rng = np.random.default_rng(0)
    chunks = [rng.normal(3.0, 2.0, (n, 66)) for n in (100, 1, 57, 300)]
    rms = RunningMeanStd((66,), "cpu")
    for c in chunks:
        rms.update(torch.as_tensor(c))
    allx = np.concatenate(chunks)
    assert np.allclose(rms.mean.numpy(), allx.mean(0), atol=1e-10)
    assert np.allclose(rms.var.numpy(), allx.var(0), atol=1e-10)
    x = torch.as_tensor(allx[:5], dtype=torch.float32)
    z = rms.normalize(x)
    ref = np.clip((allx[:5] - allx.mean(0)) / (allx.std(0) + 1e-8), -10, 10)
    assert np.allclose(z.numpy(), ref, atol=1e-4)


def test_gae_matches_the_reference_recursion():
    """marl-soccer.ipynb:454-464, restated with plain Python loops."""
    T, n = 9, 5
    g = torch.Generator().manual_seed(1)
    buf = RolloutBuffer(T, n, "cpu")
    buf.rewards = torch.randn((T, n, 2), generator=g)
    buf.values = torch.randn((T, n, 2), generator=g)
    buf.dones = (torch.rand((T, n, 2), generator=g) < 0.2).float()
    next_done = (torch.rand((n, 2), generator=g) < 0.5).float()

    class ConstCritic(Agent):
        def get_value(self, x):
            return torch.full((x.shape[0], 1), 0.25)
    agent = ConstCritic()
    rms = RunningMeanStd((66,), "cpu")
    next_obs = torch.zeros((n, 2, 66))
    adv, ret = compute_gae(agent, rms, buf, next_obs, next_done, gamma=0.9, gae_lambda=0.8)
    exp = np.zeros((T, n, 2))
    last = np.zeros((n, 2))
    for t in reversed(range(T)):
        if t == T - 1:
            nonterm, nv = 1.0 - next_done.numpy(), np.full((n, 2), 0.25)
        else:
            nonterm, nv = 1.0 - buf.dones[t + 1].numpy(), buf.values[t + 1].numpy()
        delta = buf.rewards[t].numpy() + 0.9 * nv * nonterm - buf.values[t].numpy()
        last = delta + 0.9 * 0.8 * nonterm * last
        exp[t] = last
    assert np.allclose(adv.numpy(), exp, atol=1e-5)
    assert np.allclose(ret.numpy(), exp + buf.values.numpy(), atol=1e-5)


@pytest.mark.gpu
def test_rollout_loop_on_the_device():
    """128 steps of 4 096 envs with the policy in the loop: everything stays on the GPU, the buffers hold what
    the simulator returned, red actions are uniform in [-1, 1], env-steps and episode statistics add up."""
    import parity_util as P
    from marl_soccer_b200.sim import BatchedSoccerSim
    dev = torch.device("cuda:0")
    n, T = 4096, 128
    cfg = dict(P.CONFIG)
    cfg["simulation"] = {"max_steps": 50}
    sim = BatchedSoccerSim(n, config=cfg, device=dev, seed=3)
    torch.manual_seed(0)
    agent = Agent().to(dev)
    rms = RunningMeanStd((66,), dev)
    buf = RolloutBuffer(T, n, dev)
    obs = sim.reset(2, seed=5)[:, :2].clone()
    done = torch.zeros((n, 2), device=dev)
    sim.stats(reset=True)
    gen = torch.Generator(device=dev).manual_seed(7)
    next_obs, next_done = collect_rollout(sim, agent, rms, buf, obs, done, generator=gen)
    st = sim.stats()
    assert st["env_steps"] == n * T
    assert st["episodes"] == n * (T // 50)          # truncation at max_steps, auto-reset in the step kernel
    assert torch.equal(buf.obs[0], obs)
    assert float(buf.dones.sum()) == 2 * (st["episodes"] - n * (1 if T % 50 == 0 else 0))
    assert torch.isfinite(buf.rewards).all() and torch.isfinite(buf.values).all() and torch.isfinite(buf.logprobs).all()
    assert next_obs.shape == (n, 2, 66) and next_done.shape == (n, 2)
    assert rms.count == n * T * 2
    # red agents: uniform random actions (marl-soccer.ipynb:397-400)
    red = sim.actions[:, 2:]
    assert float(red.min()) >= -1.0 and float(red.max()) <= 1.0 and abs(float(red.mean())) < 0.02
    adv, ret = compute_gae(agent, rms, buf, next_obs, next_done)
    assert adv.shape == (T, n, 2) and torch.isfinite(adv).all() and torch.isfinite(ret).all()


@pytest.mark.gpu
def test_deterministic_rollout_matches_the_same_loop_over_the_oracle():
    """The rollout half against the oracle: 256 envs from the fixed kick-off (contact-free for the 8 steps), blue plays
    the policy MEAN on normalised observations, red stands still; the same loop written out over the CPU oracle with the
    same weights must produce the same observations, rewards and values step by step (fp32 GPU matmuls vs fp32 CPU
    matmuls feed slightly different forces into two simulators: a few fp32 ulps on the observations)."""
    import oracle_lib as O
    import parity_util as P
    from marl_soccer_b200.sim import BatchedSoccerSim
    dev = torch.device("cuda:0")
    n, T = 256, 8
    torch.manual_seed(4)
    agent = Agent()
    rms = RunningMeanStd((66,), "cpu")
    rms.mean = torch.linspace(-0.3, 0.3, 66, dtype=torch.float64)
    rms.var = torch.linspace(0.5, 2.0, 66, dtype=torch.float64)
    # --- device: collect_rollout over the CUDA simulator
    sim = BatchedSoccerSim(n, config=P.CONFIG, device=dev, seed=3)
    agent_d = Agent().to(dev)
    agent_d.load_state_dict(agent.state_dict())
    rms_d = RunningMeanStd((66,), dev)
    rms_d.mean, rms_d.var = rms.mean.to(dev), rms.var.to(dev)
    buf = RolloutBuffer(T, n, dev)
    obs = sim.reset(O.MODE_FIXED, seed=5)[:, :2]
    done = torch.zeros((n, 2), device=dev)
    nxt, _ = collect_rollout(sim, agent_d, rms_d, buf, obs, done, update_normalizer=False, deterministic=True)
    # --- oracle: the same loop, plain
    ora = O.OracleVec(n, P.CONFIG, seed=3)
    o = ora.reset(O.MODE_FIXED, seed=5)
    for t in range(T):
        blue = torch.from_numpy(o[:, :2].copy())
        assert np.allclose(buf.obs[t].cpu().numpy(), blue.numpy(), atol=2e-4), t
        with torch.no_grad():
            x = rms.normalize(blue.reshape(-1, 66))
            act = agent.get_deterministic_action(x).reshape(n, 2, 3)
            val = agent.get_value(x).reshape(n, 2)
        assert np.allclose(buf.values[t].cpu().numpy(), val.numpy(), atol=2e-3), t
        assert np.allclose(buf.actions[t].cpu().numpy(), act.numpy(), atol=2e-4), t
        full = np.zeros((n, 4, 3), np.float32)
        full[:, :2] = act.numpy()
        o, r, d, g = ora.step(full, auto_reset=True)
        assert not d.any() and not g.any()
        assert np.allclose(buf.rewards[t].cpu().numpy(), r, atol=1e-5), t
    assert np.allclose(nxt.cpu().numpy(), o[:, :2], atol=2e-4)


@pytest.mark.gpu
def test_graphed_rollout_runs_the_whole_rollout_as_one_cuda_graph():
    """GraphedRollout: the first run is eager and captures; replays step the simulator exactly T times each (the
    device-side step counter keeps the captured sequence replayable for odd T too) and fill the buffers."""
    import parity_util as P
    from marl_soccer_b200.rollout import GraphedRollout
    from marl_soccer_b200.sim import BatchedSoccerSim
    dev = torch.device("cuda:0")
    n, T = 2048, 7
    cfg = dict(P.CONFIG)
    cfg["simulation"] = {"max_steps": 10}
    sim = BatchedSoccerSim(n, config=cfg, device=dev, seed=3)
    torch.manual_seed(0)
    agent = Agent().to(dev)
    rms = RunningMeanStd((66,), dev)
    buf = RolloutBuffer(T, n, dev)
    sim.reset(2, seed=5)
    sim.stats(reset=True)
    ro = GraphedRollout(sim, agent, rms, buf)
    for k in range(4):
        prev = sim.obs.clone()
        nxt, nd = ro.run()
        torch.cuda.synchronize()
        assert torch.equal(buf.obs[0], prev[:, :2])   # the rollout started from the observation the last one ended on
        assert torch.isfinite(buf.rewards).all() and torch.isfinite(buf.values).all()
    st = sim.stats()
    assert st["env_steps"] == 4 * T * n and st["episodes"] == n * (4 * T // 10)
    # frames shift consistently across the replays: the history really is the previous steps' frames
    o4 = sim.obs.view(n, 4, 3, 22)
    keep = ~sim.done.bool()
    last_blue = buf.obs[T - 1].view(n, 2, 3, 22)
    assert torch.equal(o4[keep][:, :2, 1], last_blue[keep][:, :, 2])


@pytest.mark.gpu
def test_fused_policy_inputs_kernel_and_bf16_graphed_rollout():
    """msoc_policy_inputs: one pass over the blue agents' observation rows = the normalised, clipped, padded bf16 policy
    input + the raw bf16 copy + per-feature sums for the running normaliser (marl-soccer.ipynb:385, :403, :264-296);
    against the same three results computed with torch.  Then the bf16 GraphedRollout that uses it (with the packed
    policy): its normaliser ends where a float64 pass over the raw observations ends."""
    import parity_util as P
    from marl_soccer_b200 import _capi
    from marl_soccer_b200.rollout import GraphedRollout
    from marl_soccer_b200.sim import BatchedSoccerSim
    dev = torch.device("cuda:0")
    L = _capi.lib()
    g = torch.Generator(device=dev).manual_seed(2)
    for n in (1, 5, 1000, 70001):
        obs = torch.randn((n, 4, 66), generator=g, device=dev) * 3.0
        obs[0, 0, :4] = torch.tensor([1e4, -1e4, 0.0, 1.0], device=dev)  # clipped at +-10
        mean, std = torch.randn(66, generator=g, device=dev), torch.rand(66, generator=g, device=dev) + 0.5
        inv_std = 1.0 / (std + 1e-8)
        shift = -mean * inv_std
        x72 = torch.zeros((2 * n, 72), dtype=torch.bfloat16, device=dev)
        raw = torch.zeros((n, 2, 66), dtype=torch.bfloat16, device=dev)
        mom = torch.zeros((2, 66), dtype=torch.float64, device=dev)
        for _ in range(2):  # moments accumulate
            _capi.check(L.msoc_policy_inputs(obs.data_ptr(), n, shift.data_ptr(), inv_std.data_ptr(), x72.data_ptr(), raw.data_ptr(),
                                             mom.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
        rows = obs[:, :2].reshape(-1, 66)
        ref = torch.clamp(torch.addcmul(shift, rows, inv_std), -10.0, 10.0).to(torch.bfloat16)
        assert torch.equal(x72[:, :66], ref) and bool((x72[:, 66:] == 0).all())
        assert torch.equal(raw, obs[:, :2].to(torch.bfloat16))
        r64 = rows.to(torch.float64)
        assert torch.allclose(mom[0], 2 * r64.sum(0), rtol=1e-6, atol=1e-3) and torch.allclose(mom[1], 2 * r64.square().sum(0), rtol=1e-6)
    # the bf16 rollout on top of it
    n, T = 4096, 6
    sim = BatchedSoccerSim(n, config=P.CONFIG, device=dev, seed=3)
    torch.manual_seed(0)
    agent = Agent().to(dev)
    rms, ref_rms = RunningMeanStd((66,), dev), RunningMeanStd((66,), dev)
    buf = RolloutBuffer(T, n, dev, obs_dtype=torch.bfloat16)
    sim.reset(2, seed=5)
    ro = GraphedRollout(sim, agent, rms, buf, policy_dtype=torch.bfloat16)
    assert ro.fused_inputs
    for k in range(3):
        first = sim.obs[:, :2].clone()
        ro.run()
        torch.cuda.synchronize()
        assert torch.equal(buf.obs[0], first.to(torch.bfloat16))
        ref_rms.update(buf.obs.float())  # the buffer holds the bf16-rounded observations the moments were taken from in fp32
        assert torch.isfinite(buf.values).all() and torch.isfinite(buf.logprobs).all() and bool((buf.actions.abs() < 20).all())
    assert torch.allclose(rms.mean, ref_rms.mean, atol=2e-3) and torch.allclose(rms.var, ref_rms.var, rtol=2e-2, atol=1e-3)
    assert sim.stats()["env_steps"] == 3 * T * n
