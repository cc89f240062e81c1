"""The specialiser is optional: without NVRTC (simulated with NSGYM_B200_NVRTC=none) or with NSGYM_B200_NO_JIT=1
the precompiled kernels run, results unchanged, and the library says why it did not specialise."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHECK = r"""
import ctypes as C, json, sys
sys.path.insert(0, %r)
import ns_gym_b200.schedulers as S, ns_gym_b200.update_functions as U
from ns_gym_b200 import native as nv
from ns_gym_b200.compile import compile_program
lib = nv.load()
p = compile_program("CartPole-v1", {"masspole": U.IncrementUpdate(S.ContinuousScheduler(), k=0.1)}, 16)
log = C.create_string_buffer(4096)
rc = lib.nsgym_jit_check(C.byref(p.spec), 0, 0, 0, None, 0, log, len(log))
print(json.dumps({"rc": rc, "err": (lib.nsgym_last_error() or b"").decode()}))
""" % ROOT

RUN = r"""
import ctypes as C, json, sys
sys.path.insert(0, %r)
import torch
from tests import parity_util as pu
from tests.cases import CASES
from ns_gym_b200 import native as nv
env = pu.gpu_env(CASES["c1_cartpole_readme"], 1 << 16, precision="fp64")
env.reset(seed=2)
a = torch.zeros(env.num_envs, dtype=torch.int32, device=env.device)
for _ in range(5):
    env.step_raw(a)
torch.cuda.synchronize()
failed, why = C.c_int64(0), C.create_string_buffer(512)
on = nv.load().nsgym_jit_stats(None, None, C.byref(failed), why, len(why))
print(json.dumps({"specialized": env.last_kernel_specialized, "enabled": on, "failed": failed.value,
                  "why": why.value.decode(), "checksum": float(env.buffers["state"].double().sum())}))
""" % ROOT


def _py(code, **env):
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600,
                       env={**os.environ, **env})
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_jit_check_reports_a_missing_nvrtc():
    out = _py(CHECK, NSGYM_B200_NVRTC="none")
    assert out["rc"] == -3 and "not found" in out["err"], out
    assert _py(CHECK)["rc"] > 0


@pytest.mark.gpu
def test_large_batches_fall_back_to_the_precompiled_kernels():
    base = _py(RUN)
    assert base["specialized"] and base["enabled"] == 1 and base["failed"] == 0, base
    missing = _py(RUN, NSGYM_B200_NVRTC="none")
    assert not missing["specialized"] and missing["failed"] >= 1 and "not found" in missing["why"], missing
    off = _py(RUN, NSGYM_B200_NO_JIT="1")
    assert not off["specialized"] and off["enabled"] == 0, off
    # fp64: the three runs leave bit-identical states
    assert base["checksum"] == missing["checksum"] == off["checksum"]
