"""CPU: the C-ABI library loads, exports every symbol include/nsgym_b200.h declares, its struct
layouts agree with the ctypes mirror, and it refuses to compute without a GPU."""
import ctypes as C
import os
import re

import pytest

from ns_gym_b200 import native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    return native.load(build_if_missing=True)


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "nsgym_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nsgym_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    declared = _declared_symbols()
    assert len(declared) >= 15
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in the header but not exported"
    assert set(declared) == set(native.EXPORTS)


def test_struct_layouts_match(lib):
    for which, st in enumerate((native.NsgymSlot, native.NsgymSpec, native.NsgymLayout,
                                native.NsgymBuffers, native.NsgymHostOut)):
        assert lib.nsgym_sizeof(which) == C.sizeof(st)
    assert lib.nsgym_abi_version() == native.ABI_VERSION


def test_header_enums_match_python():
    text = open(os.path.join(ROOT, "include", "nsgym_b200.h")).read()
    for name, value in re.findall(r"NSGYM_([A-Z0-9_]+)\s*=\s*(\d+)", text):
        if hasattr(native, name):
            assert getattr(native, name) == int(value), name


def test_spec_validation_errors(lib):
    spec = native.NsgymSpec()
    h = C.c_void_p()
    assert lib.nsgym_create(C.byref(spec), C.byref(h)) < 0          # abi_version 0
    assert b"ABI version" in lib.nsgym_last_error()
    spec.abi_version = native.ABI_VERSION
    spec.env_kind = 99
    assert lib.nsgym_create(C.byref(spec), C.byref(h)) < 0
    spec.env_kind = native.ENV_CARTPOLE
    spec.n_envs = 0
    assert lib.nsgym_create(C.byref(spec), C.byref(h)) < 0
    spec.n_envs = 8
    spec.n_slots = 7                                                 # CartPole has 6 parameters
    assert lib.nsgym_create(C.byref(spec), C.byref(h)) < 0
    assert b"tunable parameters" in lib.nsgym_last_error()
    spec.n_slots = 1
    spec.slots[0].upd_op = native.UPD_D_INC                          # distribution opcode on a scalar env
    assert lib.nsgym_create(C.byref(spec), C.byref(h)) < 0
    assert h.value is None


def test_no_cpu_execution_path(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    spec = native.NsgymSpec()
    spec.abi_version = native.ABI_VERSION
    spec.env_kind = native.ENV_CARTPOLE
    spec.n_envs = 4
    h = C.c_void_p()
    rc = lib.nsgym_create(C.byref(spec), C.byref(h))
    assert rc == -2 and b"no CPU execution path" in lib.nsgym_last_error()
    from ns_gym_b200.vector_env import NSVectorEnv

    with pytest.raises(native.NsgymError):
        NSVectorEnv("CartPole-v1", {}, 4)


def test_descriptions_do_not_execute_on_host():
    import ns_gym_b200.schedulers as S
    import ns_gym_b200.update_functions as U

    fn = U.IncrementUpdate(S.ContinuousScheduler(), k=1.0)
    with pytest.raises(RuntimeError):
        fn(1.0, 0)
    with pytest.raises(RuntimeError):
        fn.scheduler(0)


def test_stale_library_is_never_loaded(monkeypatch):
    """A library older than its sources is rebuilt (or refused with NSGYM_B200_NO_AUTOBUILD): the
    kernels that run are the ones in the tree -- and there is no CPU path to fall back to."""
    from ns_gym_b200 import build, native

    assert build.is_current()
    monkeypatch.setattr(native, "_lib", None)
    monkeypatch.setattr(build, "is_current", lambda: False)
    monkeypatch.setenv("NSGYM_B200_NO_AUTOBUILD", "1")
    with pytest.raises(native.NsgymError, match="older than the sources"):
        native.load()
    rebuilt = []
    monkeypatch.delenv("NSGYM_B200_NO_AUTOBUILD")
    monkeypatch.setattr(build, "build_library", lambda *a, **k: rebuilt.append(1) or build.LIB_PATH)
    assert native.load() is not None and rebuilt == [1]
