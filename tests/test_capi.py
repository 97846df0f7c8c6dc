"""The C-ABI shared library loads on a machine without a GPU and exports every symbol that
include/msoc.h declares; the ctypes struct mirrors have the C sizes.  No compute calls here."""
import ctypes as C
import os
import re
import subprocess

import pytest

from marl_soccer_b200 import _capi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "msoc.h")


@pytest.fixture(scope="module")
def libpath():
    return build.build()


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(msoc_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported(libpath):
    L = C.CDLL(libpath)
    names = declared_functions()
    assert len(names) >= 15
    for name in names:
        assert hasattr(L, name), f"{name} declared in include/msoc.h but not exported by libmsoc.so"
    assert set(names) == set(_capi.EXPORTS), "marl_soccer_b200._capi.EXPORTS is out of sync with the header"


def test_version_and_error_string(libpath):
    L = C.CDLL(libpath)
    _capi.declare(L)
    assert L.msoc_version() == 2
    assert isinstance(L.msoc_last_error(), bytes)
    assert L.msoc_launch_count() == 0


def test_struct_sizes_match_c(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "msoc.h"\nint main(void){printf("%zu %zu %zu\\n", '
                   'sizeof(msoc_config), sizeof(msoc_env_state), sizeof(msoc_stats));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert [int(x) for x in out] == [C.sizeof(_capi.MsocConfig), C.sizeof(_capi.MsocEnvState), C.sizeof(_capi.MsocStats)]


def test_create_without_gpu_fails_loudly(libpath):
    """No silent CPU fallback: without a CUDA device msoc_create returns an error."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = C.CDLL(libpath)
    _capi.declare(L)
    cfg = _capi.make_config({"physics": {"max_velocity": 200, "agent_mass": 10, "ball_mass": 1, "agent_friction": 0.99,
                                         "ball_friction": 0.97},
                             "rewards": {"move_ball_to_goal_multiplier": 0.1, "goal_scored_reward": 4.0,
                                         "goal_conceded_penalty": 0.0, "alive_penalty": 1e-5},
                             "simulation": {"max_steps": 1000}})
    h = C.c_void_p()
    rc = L.msoc_create(C.byref(cfg), 8, 0, 0, 0, C.byref(h))
    assert rc != 0 and not h.value
    assert b"no CUDA device" in L.msoc_last_error() or b"CUDA" in L.msoc_last_error()
    with pytest.raises(_capi.MsocError):
        from marl_soccer_b200.sim import BatchedSoccerSim
        BatchedSoccerSim(8)
