"""The C ABI is a C ABI: ``include/nsgym_b200.h`` compiles as plain C99, and (GPU) a C program
drives the library without Python or torch (tests/c/abi_driver.c)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def test_header_is_plain_c(tmp_path):
    src = tmp_path / "probe.c"
    src.write_text('#include "nsgym_b200.h"\n'
                   "int probe(void) { NsgymSpec s; NsgymSlot t; NsgymLayout l; NsgymBuffers b; NsgymHostOut o;\n"
                   "  (void)s; (void)t; (void)l; (void)b; (void)o; return NSGYM_ABI_VERSION + NSGYM_UPD_D_RANDOM; }\n")
    r = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                        "-fsyntax-only", str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


@pytest.mark.gpu
def test_c_driver_runs_the_quickstart_through_the_abi(tmp_path):
    from ns_gym_b200 import native

    native.load()
    lib_dir = os.path.dirname(native.lib_path())
    exe = str(tmp_path / "abi_driver")
    cc = shutil.which("gcc")
    r = subprocess.run([cc, "-std=c99", "-O1", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CUDA, "include"),
                        os.path.join(ROOT, "tests", "c", "abi_driver.c"), "-o", exe, "-L", lib_dir, "-lnsgym_b200",
                        "-L", os.path.join(CUDA, "lib64"), "-lcudart", "-lm",
                        f"-Wl,-rpath,{lib_dir}", f"-Wl,-rpath,{os.path.join(CUDA, 'lib64')}"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "C ABI driver OK" in r.stdout
