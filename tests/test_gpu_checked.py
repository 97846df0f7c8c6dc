"""The kernels' own bounds / invariant checks (libmsoc_checked.so = the same sources with -DMSOC_CHECKS; see MSOC_CHECK in
marl_soccer_b200/csrc/step_core.cuh): list entries inside the stepped range, contact-pool and overflow slots, arbiter
cache counts, env indices of the observation builder, the device-side step counter, non-NaN state.  compute-sanitizer
is closed on the GPU pool this repo is developed on (profiles/r02_sanitizer.log), so these checks stand in for its
memcheck on the index arithmetic.  Runs in a subprocess because the library is chosen at import time (MSOC_LIB)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import sys
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
import numpy as np, torch
import parity_util as P
from marl_soccer_b200 import _capi
from marl_soccer_b200.sim import BatchedSoccerSim
from marl_soccer_b200.host_api import HostBufferSim
L = _capi.lib()
assert L.msoc_debug_errors() == 0, "checks are compiled in and clean at start"
cfg = {**P.CONFIG, "simulation": {"max_steps": 23}}
g = torch.Generator(device="cuda").manual_seed(0)
for n in (1, 31, 33, 129, 1000, 8192, 70001):
    sim = BatchedSoccerSim(n, config=cfg, seed=n)
    sim.reset(2, seed=1)
    for t in range(60):
        sim.step(torch.rand((n, 4, 3), generator=g, device="cuda") * 2.4 - 1.2)
        if t %% 17 == 5:
            sim.reset(0, mask=(torch.arange(n, device="cuda") %% 5 == 0))
    st = sim.get_states(np.arange(min(n, 64)))
    sim.set_states(np.arange(min(n, 64)), st)
    sim.step(torch.zeros((n, 4, 3), device="cuda"))
    bits = L.msoc_debug_errors()
    assert bits == 0, (n, bin(bits))
    stats = sim.stats()
    assert stats["contact_overflow"] == 0 and stats["env_steps"] == 61 * n, stats
    sim.close()
# the chunked host-buffer step (several pipeline chunks on two streams)
n = 200000
hs = HostBufferSim(n, cfg, seed=3)
hs.reset(2, seed=2)
rng = np.random.default_rng(0)
for t in range(5):
    hs.step(rng.uniform(-1, 1, (n, 4, 3)).astype(np.float32))
bits = L.msoc_debug_errors()
assert bits == 0, ('chunked host step', bin(bits))
print("checked build clean")
"""


@pytest.mark.gpu
def test_checked_build_reports_no_violation():
    from marl_soccer_b200 import build
    lib = build.build_checked()
    env = dict(os.environ, MSOC_LIB=lib)
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "checked build clean" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_product_build_has_the_checks_compiled_out():
    import ctypes as C
    from marl_soccer_b200 import build
    L = C.CDLL(build.build())
    L.msoc_debug_errors.restype = C.c_int
    assert L.msoc_debug_errors() == -1
