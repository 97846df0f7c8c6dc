"""Drop-in for the reference's `soccer_env.py`: the PettingZoo-style `SoccerEnv` (one 2v2 env), the
factories `soccer_raw_env` / `soccerenv` / `make_env` and `get_observation_scalers`, with the same
spaces, dict packaging, option keys and exceptions (soccer_simulation/soccer_env.py:16-221).

The embedded `Game` of the reference (soccer_env.py:59) is replaced by a one-env handle of the CUDA
simulator (include/msoc.h); `SoccerEnv.step` is one fused device step (three kernel launches).  There is no
CPU physics path: constructing a SoccerEnv needs the built extension and a CUDA device.
"""
from __future__ import annotations

import json
import os
from typing import Any, Dict, Optional

import numpy as np

from . import _capi
from .spaces import Box, ParallelEnv

# Kept for parity with the reference module (soccer_env.py:12); unused there as well.
FRAME_SKIPS = 6

_AGENTS = [f"agent_{i}" for i in range(4)]


class _GameView:
    """What callers reach through `env._game` in the reference: `.config` (get_observation_scalers,
    soccer_env.py:211), `.score`, `.steps`, `.max_steps`; body poses for the renderer."""

    def __init__(self, env: "SoccerEnv"):
        self._env = env
        self.config = env._config
        self.max_steps = env._config["simulation"]["max_steps"]

    @property
    def score(self) -> dict:
        s, _ = self._env._sim.counters()
        return {"blue": int(s[0, 0]), "red": int(s[0, 1])}

    @property
    def steps(self) -> int:
        _, st = self._env._sim.counters()
        return int(st[0])

    def poses(self) -> dict:
        """Render pull-back of this env (renderer.py:30-42): positions and angles of the five bodies."""
        S = self._env._sim.get_state(0)
        return {"agents": [((S.pos[i][0], S.pos[i][1]), S.ang[i]) for i in range(4)],
                "ball": (S.pos[4][0], S.pos[4][1])}


class SoccerEnv(ParallelEnv):
    metadata = {"render_modes": ["human"], "name": "soccer_sim_v1"}

    def __init__(self, render_mode: Optional[str] = None, config: Optional[Dict[str, Any]] = None, **kwargs):
        # soccer_env.py:21-24
        if "env" in kwargs and kwargs["env"] != 1:
            raise ValueError("SoccerEnv supports only a single environment (env must be 1).")
        if "num_envs" in kwargs and kwargs["num_envs"] != 1:
            raise ValueError("SoccerEnv supports only a single environment (num_envs must be 1).")
        self.render_mode = render_mode
        self.possible_agents = list(_AGENTS)
        self.agents = self.possible_agents[:]
        self.agent_name_mapping = {a: i for i, a in enumerate(self.possible_agents)}
        self._action_space = Box(low=-1.0, high=1.0, shape=(3,), dtype=np.float32)
        self._stack_size = 3
        self._frame_size = 22
        if config is None:  # soccer_env.py:42-55
            path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "config.json")
            if not os.path.exists(path):
                raise FileNotFoundError("Could not find config.json next to soccer_env.py; pass config explicitly.")
            with open(path, "r") as f:
                config = json.load(f)
        self._config = config
        physics_cfg = config.get("physics", {}) if isinstance(config, dict) else {}
        self._force_max = float(physics_cfg.get("action_force_max", 150000.0))
        self._torque_max = float(physics_cfg.get("action_torque_max", 100000.0))
        self._observation_space = Box(low=-np.inf, high=np.inf, shape=(self._frame_size * self._stack_size,),
                                      dtype=np.float32)
        # `_sim_factory` is a test seam (tests inject a checker-backed sim to exercise this wrapper's
        # host logic without a GPU); the product always builds the CUDA handle.
        factory = kwargs.get("_sim_factory")
        seed = kwargs.get("_seed")
        if seed is None:  # the reference's Game starts from np.random.default_rng() (game/game.py:17)
            seed = int(np.random.SeedSequence().generate_state(1, dtype=np.uint64)[0] >> 1)
        if factory is None:
            from .host_api import HostBufferSim
            self._sim = HostBufferSim(1, config, seed=seed, device=int(kwargs.get("device", 0)))
        else:
            self._sim = factory(1, config, seed)
        self._game = _GameView(self)
        self._renderer = None

    def observation_space(self, agent):
        return self._observation_space

    def action_space(self, agent):
        return self._action_space

    def reset(self, seed=None, options=None):
        self.agents = self.possible_agents[:]
        use_fixed = use_full_random = False
        if isinstance(options, dict):  # soccer_env.py:86-88
            use_fixed = bool(options.get("use_fixed_positions", False))
            use_full_random = bool(options.get("use_full_random_positions", False))
        mode = _capi.MODE_FIXED if use_fixed else _capi.MODE_FULL_RANDOM if use_full_random else _capi.MODE_RANDOM
        obs = self._sim.reset(mode, seed=None if seed is None else int(seed))
        observations = {a: obs[0, i].astype(np.float32) for i, a in enumerate(self.possible_agents)}
        infos = {a: {} for a in self.possible_agents}
        return observations, infos

    def step(self, actions):
        expected = list(self.possible_agents)
        missing = [a for a in expected if a not in actions]
        if missing:
            raise ValueError(f"Missing actions for agents: {missing}. Expected actions for {expected}.")
        extra = [a for a in actions.keys() if a not in expected]
        if extra:
            raise ValueError(f"Received actions for unknown agents: {extra}. Expected only {expected}.")
        batch = np.zeros((1, 4, 3), np.float32)
        for i, a in enumerate(expected):
            arr = np.asarray(actions.get(a), dtype=np.float32)
            if arr.shape != (3,):
                raise ValueError(f"Action for agent '{a}' must have shape (3,), got {arr.shape}.")
            if not np.all(np.isfinite(arr)):
                raise ValueError(f"Action contains non-finite values for agent '{a}': {arr.tolist()}")
            batch[0, i] = arr  # clip + scale happen in the kernel (soccer_env.py:119-124)
        obs, rew, done, goal = self._sim.step(batch, auto_reset=False)
        score = self._sim.score
        info = {"score": {"blue": int(score[0, 0]), "red": int(score[0, 1])}}
        if goal[0] != 0:
            info["goal_scored_by"] = "blue" if goal[0] > 0 else "red"
        observations = {a: obs[0, i].astype(np.float32) for i, a in enumerate(self.possible_agents)}
        rewards = {"agent_0": float(rew[0, 0]), "agent_1": float(rew[0, 1]), "agent_2": 0.0, "agent_3": 0.0}
        terminations = {a: False for a in self.possible_agents}
        truncations = {a: bool(done[0]) for a in self.possible_agents}
        infos = {a: dict(info) for a in self.possible_agents}
        if any(terminations.values()) or any(truncations.values()):
            self.agents = []
        return observations, rewards, terminations, truncations, infos

    def render(self):
        """Pulls this one env's poses back to the host (msoc_get_state) and draws them with pygame when it is installed
        (marl_soccer_b200/renderer.py; the reference: soccer_env.py:156-162).  Returns the poses; without pygame nothing
        is drawn."""
        if self.render_mode != "human":
            return None
        poses = self._game.poses()
        if self._renderer is None:
            from .renderer import PygameRenderer
            try:
                self._renderer = PygameRenderer()
            except ImportError:  # pygame is not installed: pose pull-back only
                self._renderer = False
        if self._renderer:
            self._renderer.draw(poses)
        return poses

    def close(self):
        if self._renderer:
            self._renderer.close()
        self._renderer = None
        sim = getattr(self, "_sim", None)
        if sim is not None and hasattr(sim, "close"):
            sim.close()


def soccer_raw_env(**kwargs):
    """Return the raw, unwrapped environment (soccer_env.py:174-178)."""
    return SoccerEnv(**kwargs)


def soccerenv(**kwargs):
    """Return the environment the reference calls "wrapped" (its wrapper is commented out, soccer_env.py:181-187)."""
    return soccer_raw_env(**kwargs)


def make_env(**kwargs):
    """Instantiate and return the soccer environment (soccer_env.py:191-197)."""
    return soccerenv(**kwargs)


def get_observation_scalers(env: SoccerEnv):
    """Maximum ranges used to scale observation components (soccer_env.py:200-221)."""
    physics_cfg = env._game.config.get("physics", {})
    max_velocity = float(physics_cfg.get("max_velocity", 400.0))
    max_ang_vel = float(physics_cfg.get("max_angular_velocity", physics_cfg.get("action_torque_max", 100000.0) / 100.0))
    field_diag = float((800 ** 2 + 600 ** 2) ** 0.5)
    return {
        "max_velocity": max_velocity,
        "max_angular_velocity": max_ang_vel,
        "field_diagonal": field_diag,
        "stack_size": env._stack_size,
        "frame_size": env._frame_size,
    }
