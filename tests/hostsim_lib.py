"""TEST INFRASTRUCTURE: host build of the product's per-env fp32 step arithmetic
(tests/hostsim/hostsim.cu includes marl_soccer_b200/csrc/step_core.cuh).  It exists so that the fp32
logic of the kernels can be compared with the fp64 oracle on machines without a GPU.  The product
package never imports this module; the real kernels are checked by the `-m gpu` tests."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from marl_soccer_b200 import _capi

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostsim", "hostsim.cu")
LIB = os.path.join(HERE, "hostsim", "libhostsim.so")
CORE = os.path.join(os.path.dirname(HERE), "marl_soccer_b200", "csrc", "step_core.cuh")

_lib = None


def build(force: bool = False) -> str:
    stale = (not os.path.exists(LIB)) or any(
        os.path.getmtime(p) > os.path.getmtime(LIB) for p in (SRC, CORE))
    if force or stale:
        subprocess.run(["nvcc", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "--shared", "-Xcompiler",
                        "-fPIC", "-o", LIB, SRC], check=True, capture_output=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        vp, u64, i64 = C.c_void_p, C.c_uint64, C.c_int64
        L.hsim_create.restype = vp
        L.hsim_create.argtypes = [C.POINTER(_capi.MsocConfig), i64, u64, u64, vp]
        L.hsim_destroy.argtypes = [vp]
        L.hsim_reset.argtypes = [vp, vp, C.c_int, C.c_int, u64, vp]
        L.hsim_step.argtypes = [vp, vp, vp, vp, vp, vp, vp, C.c_uint32]
        L.hsim_stats.argtypes = [vp, vp, C.c_int]
        L.hsim_last.argtypes = [vp, vp, vp]
        L.hsim_get_state.argtypes = [vp, i64, C.POINTER(_capi.MsocEnvState)]
        L.hsim_set_state.argtypes = [vp, i64, C.POINTER(_capi.MsocEnvState)]
        _lib = L
    return _lib


class HostSim:
    """Same surface as tests/parity_util.DeviceSim, on the host."""

    name = "hostsim"

    def __init__(self, n: int, config: dict, seed: int = 0, global_offset: int = 0):
        self._L = lib()
        self.n = int(n)
        self._cfg = _capi.make_config(config)
        self.obs = np.zeros((self.n, 4, 66), np.float32)
        self._h = self._L.hsim_create(C.byref(self._cfg), self.n, seed, global_offset, self.obs.ctypes.data)

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.hsim_destroy(self._h)
            self._h = None

    def reset(self, mode: int = 0, seed: int | None = None, mask=None) -> np.ndarray:
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        self._L.hsim_reset(self._h, None if m is None else m.ctypes.data, mode, 0 if seed is None else 1,
                           0 if seed is None else int(seed), self.obs.ctypes.data)
        return self.obs.copy()

    def step(self, actions, auto_reset: bool = True):
        a = np.ascontiguousarray(actions, dtype=np.float32).reshape(self.n, 12)
        out = np.zeros_like(self.obs)
        rew = np.zeros((self.n, 2), np.float32)
        done = np.zeros(self.n, np.uint8)
        goal = np.zeros(self.n, np.int8)
        self.score = np.zeros((self.n, 2), np.int32)
        self._L.hsim_step(self._h, a.ctypes.data, out.ctypes.data, rew.ctypes.data,
                          done.ctypes.data, goal.ctypes.data, self.score.ctypes.data, 1 if auto_reset else 0)
        self.obs = out
        return out.copy(), rew, done, goal

    def get_state(self, i: int) -> _capi.MsocEnvState:
        S = _capi.MsocEnvState()
        self._L.hsim_get_state(self._h, i, C.byref(S))
        return S

    def set_state(self, i: int, S: _capi.MsocEnvState) -> None:
        self._L.hsim_set_state(self._h, i, C.byref(S))

    def last(self):
        """(contacts solved, work class) of every env in the last step; class -1 = contact-free, else step_core.cuh LOAD_*."""
        contacts = np.zeros(self.n, np.int32)
        load = np.zeros(self.n, np.int32)
        self._L.hsim_last(self._h, contacts.ctypes.data, load.ctypes.data)
        return contacts, load

    def get_obs(self, i: int) -> np.ndarray:
        return self.obs[i].copy()

    def stats(self, reset: bool = False) -> dict:
        out = np.zeros(8, np.float64)
        self._L.hsim_stats(self._h, out.ctypes.data, 1 if reset else 0)
        keys = [k for k, _ in _capi.MsocStats._fields_]
        return dict(zip(keys, out.tolist()))
