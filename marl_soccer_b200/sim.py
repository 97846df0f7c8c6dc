"""BatchedSoccerSim -- the device-resident form of the reference's env stack.

N independent 2v2 soccer envs live in HBM buffers (contiguous per-env records) owned by one C-ABI handle
(include/msoc.h); one fused CUDA kernel per `step` replaces
SyncMultiAgentVecEnv.step -> SoccerEnv.step -> Game.step -> pymunk Space.step
(soccer_simulation/marl_vecenv.py:30-68, soccer_env.py:100-154, game/game.py:378-437).
Actions, observations, rewards, dones and goal flags are torch CUDA tensors; nothing crosses PCIe.

There is no CPU path: constructing a sim without the built extension or without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import json
import os

import numpy as np
import torch

from . import _capi

_DEFAULT_CONFIG_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "config.json")


def load_default_config() -> dict:
    """The shipped config.json (same keys and values as soccer_simulation/config.json)."""
    with open(_DEFAULT_CONFIG_PATH) as f:
        return json.load(f)


class _DevMem:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def _view(ptr, shape, typestr, device) -> torch.Tensor:
    with torch.cuda.device(device):
        return torch.as_tensor(_DevMem(ptr, shape, typestr), device=device)


class BatchedSoccerSim:
    """N envs on one GPU.

    step() returns views of internal buffers that stay valid until the next step():
      obs    (N, 4, 66) float32   3 stacked 22-float frames per agent, newest last (soccer_env.py:37-39)
      reward (N, 2)     float32   blue agents only; red is always 0.0 (soccer_env.py:141-146)
      done   (N,)       uint8     truncation at max_steps (soccer_env.py:148); never "terminated"
      goal   (N,)       int8      +1 blue scored, -1 red scored (info["goal_scored_by"])
      score  (N, 2)     int32     info["score"] of the step, before any auto-reset
    """

    def __init__(self, num_envs: int, config: dict | None = None, device: str | int | torch.device = "cuda:0",
                 seed: int = 0, global_env_offset: int = 0):
        if not torch.cuda.is_available():
            raise _capi.MsocError("BatchedSoccerSim needs a CUDA device (there is no CPU fallback)")
        self._L = _capi.lib()
        self.device = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        if self.device.type != "cuda":
            raise _capi.MsocError("BatchedSoccerSim runs on CUDA devices only")
        self.num_envs = int(num_envs)
        self.config = config if config is not None else load_default_config()
        self.global_env_offset = int(global_env_offset)
        self._cfg = _capi.make_config(self.config)
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        h = C.c_void_p()
        _capi.check(self._L.msoc_create(C.byref(self._cfg), self.num_envs, dev_index, int(seed) & (2**64 - 1),
                                        self.global_env_offset, C.byref(h)))
        self._h = h
        n, d = self.num_envs, self.device
        # zero-copy torch views of the handle's device buffers (valid while the handle lives); they
        # already hold the first spawn's observation (Game.__init__ -> reset, game/game.py:43,74)
        bufs = _capi.MsocBuffers()
        _capi.check(self._L.msoc_device_buffers(self._h, C.byref(bufs)))
        self.obs = _view(bufs.obs, (n, 4, 66), "<f4", d)
        self.actions = _view(bufs.actions, (n, 4, 3), "<f4", d)
        self.reward = _view(bufs.reward, (n, 2), "<f4", d)
        self.done = _view(bufs.done, (n,), "|u1", d)
        self.goal = _view(bufs.goal, (n,), "|i1", d)
        self.score = _view(bufs.score, (n, 2), "<i4", d)
        self._stats = torch.zeros((8,), dtype=torch.float64, device=d)

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._L.msoc_destroy(self._h)
            self._h = None

    __del__ = close

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def reset(self, mode: int = _capi.MODE_RANDOM, seed: int | None = None, mask: torch.Tensor | None = None) -> torch.Tensor:
        """Game.reset for all (or the masked) envs; env i is seeded with seed + global index
        (marl_vecenv.py:23).  Returns the stacked observation (3 copies of frame 0, soccer_env.py:92-96)."""
        mptr = None
        if mask is not None:
            mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
            mptr = mask.data_ptr()
        _capi.check(self._L.msoc_reset(self._h, mptr, int(mode), 0 if seed is None else 1,
                                       0 if seed is None else int(seed) & (2**64 - 1), self.obs.data_ptr(),
                                       self._stream()))
        return self.obs

    def step(self, actions: torch.Tensor, auto_reset: bool = True, general_path: bool = False):
        """One env-step for every env.  actions: (N, 4, 3) float32 CUDA tensor in [-1, 1] (clipped on
        the device like soccer_env.py:119).  general_path (debugging / tests): every env that touches something is
        stepped by the general contact path instead of its work class's own."""
        if actions.device != self.device or actions.dtype != torch.float32 or not actions.is_contiguous():
            actions = actions.to(device=self.device, dtype=torch.float32).contiguous()
        if actions.numel() != self.num_envs * 12:
            raise ValueError(f"actions must have shape ({self.num_envs}, 4, 3), got {tuple(actions.shape)}")
        _capi.check(self._L.msoc_step(self._h, actions.data_ptr(), self.obs.data_ptr(),
                                      self.reward.data_ptr(), self.done.data_ptr(), self.goal.data_ptr(),
                                      self.score.data_ptr(),
                                      (_capi.STEP_AUTO_RESET if auto_reset else 0) | (_capi.STEP_GENERAL_PATH if general_path else 0),
                                      self._stream()))
        return self.obs, self.reward, self.done, self.goal

    def stats_tensor(self, reset: bool = False) -> torch.Tensor:
        """8 float64 on the device (msoc_stats layout): episodes, return sum, goals blue/red, env-steps,
        contacts, contact overflow -- ready for one all-reduce per rollout."""
        _capi.check(self._L.msoc_stats_device(self._h, self._stats.data_ptr(), 1 if reset else 0, self._stream()))
        return self._stats

    def stats(self, reset: bool = False) -> dict:
        t = self.stats_tensor(reset).cpu().tolist()
        return dict(zip([k for k, _ in _capi.MsocStats._fields_], t))

    def class_counts(self) -> dict:
        """Work classes of the last step (instrumentation): envs handed to the contact kernel per class."""
        out = (C.c_int32 * 4)()
        _capi.check(self._L.msoc_last_class_counts(self._h, C.byref(out), self._stream()))
        return {"light": out[0], "heavy": out[1], "pair": out[2], "multi": out[3]}

    def get_states(self, idx) -> list:
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        arr = (_capi.MsocEnvState * len(idx))()
        torch.cuda.synchronize(self.device)
        _capi.check(self._L.msoc_get_state(self._h, idx.ctypes.data, len(idx), C.byref(arr)))
        return list(arr)

    def set_states(self, idx, states) -> None:
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        arr = (_capi.MsocEnvState * len(idx))(*states)
        torch.cuda.synchronize(self.device)
        _capi.check(self._L.msoc_set_state(self._h, idx.ctypes.data, len(idx), C.byref(arr)))

    def counters(self):
        """(score (N,2) int32, steps (N,) int32) of the live episodes, on the host."""
        score = np.zeros((self.num_envs, 2), np.int32)
        steps = np.zeros((self.num_envs,), np.int32)
        _capi.check(self._L.msoc_read_counters(self._h, score.ctypes.data, steps.ctypes.data, self._stream()))
        return score, steps
