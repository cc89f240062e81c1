"""Build ``libnsgym_b200.so`` in-tree with nvcc for sm_100a (cross-compiles without a GPU).

Fifteen translation units, compiled in parallel:

* ``nsgym_classic_kind.cu``  x 10: the classic-control kernels of one env kind in one precision
                          (fp32 fast mode with FMA contraction; fp64 parity mode with ``-fmad=false``:
                          NumPy rounds every op)
* ``nsgym_f32.cu`` / ``nsgym_f64.cu``  dispatch over the kinds + the known-answer evaluation kernels
* ``nsgym_gridworld.cu``  gridworld kernels, ``-fmad=false`` (fp64 cumulative sums / W1)
* ``nsgym_rows.cu``       lowering of per-env rows (heterogeneous batches, host code)
* ``nsgym_abi.cu``        the C ABI (host code)
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIB_DIR = os.path.join(HERE, "_lib")
LIB_PATH = os.path.join(LIB_DIR, "libnsgym_b200.so")
OBJ_DIR = os.path.join(LIB_DIR, "obj")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", INCLUDE, "-I", CSRC,
          "--expt-relaxed-constexpr"]
# object name -> (source, extra flags).  The classic-control kernels build as one unit per
# (precision, env kind): ten units in parallel instead of two long ones.
UNITS = {
    "nsgym_f32.o": ("nsgym_f32.cu", []),
    "nsgym_f64.o": ("nsgym_f64.cu", ["-fmad=false"]),
    "nsgym_gridworld.o": ("nsgym_gridworld.cu", ["-fmad=false"]),
    "nsgym_rows.o": ("nsgym_rows.cu", []),
    "nsgym_abi.o": ("nsgym_abi.cu", []),
}
for _kind in range(5):
    UNITS[f"nsgym_kind{_kind}_f32.o"] = ("nsgym_classic_kind.cu", [f"-DNSGYM_TU_KIND={_kind}", "-DNSGYM_TU_REAL=float"])
    UNITS[f"nsgym_kind{_kind}_f64.o"] = ("nsgym_classic_kind.cu",
                                         [f"-DNSGYM_TU_KIND={_kind}", "-DNSGYM_TU_REAL=double", "-fmad=false"])


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: ns_gym_b200 needs the CUDA toolkit to build its kernels")


def _source_digest() -> str:
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + ["../../include/nsgym_b200.h"]
    for name in files:
        path = os.path.normpath(os.path.join(CSRC, name))
        if os.path.isfile(path):
            h.update(name.encode())
            with open(path, "rb") as f:
                h.update(f.read())
    # flags without the absolute include paths: the tree may be mounted elsewhere (GPU box)
    h.update(repr((ARCH, [c for c in COMMON if c not in (INCLUDE, CSRC)], UNITS)).encode())
    return h.hexdigest()


def _includes(path: str, seen: set) -> set:
    """Files `path` includes with quotes, transitively (searched in csrc/ and include/)."""
    import re

    if path in seen or not os.path.isfile(path):
        return seen
    seen.add(path)
    with open(path) as f:
        for inc in re.findall(r'^\s*#\s*include\s*"([^"]+)"', f.read(), flags=re.M):
            for d in (CSRC, INCLUDE):
                _includes(os.path.join(d, inc), seen)
    return seen


def _unit_digest(name: str, extra, defines) -> str:
    h = hashlib.sha256()
    for path in sorted(_includes(os.path.join(CSRC, name), set())):
        h.update(os.path.basename(path).encode())
        with open(path, "rb") as f:
            h.update(f.read())
    h.update(repr((ARCH, [c for c in COMMON if c not in (INCLUDE, CSRC)], list(extra), list(defines))).encode())
    return h.hexdigest()


def is_current() -> bool:
    stamp = LIB_PATH + ".sha256"
    if not (os.path.exists(LIB_PATH) and os.path.exists(stamp)):
        return False
    with open(stamp) as f:
        return f.read().strip() == _source_digest()


def build_library(force: bool = False, verbose: bool = False, ptxas_info: bool = False,
                  defines: tuple = (), out: str | None = None) -> str:
    """Compile and link the shared library; returns its path.  No-op when up to date.
    ``defines`` / ``out`` build an experimental variant next to the default library."""
    variant = bool(defines or out)
    if not force and not variant and is_current():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    # one builder at a time (torchrun: every rank may find the library stale at once): the others
    # wait on the lock and find it current when they get it
    import fcntl

    with open(os.path.join(LIB_DIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not variant and is_current():
            return LIB_PATH
        return _build_locked(verbose, ptxas_info, defines, out, variant)


def _build_locked(verbose, ptxas_info, defines, out, variant) -> str:
    nvcc = _nvcc()
    lib_path = out or LIB_PATH
    obj_dir = OBJ_DIR if not variant else os.path.join(LIB_DIR, "obj_" + os.path.basename(lib_path))
    os.makedirs(obj_dir, exist_ok=True)

    def compile_unit(item):
        obj_name, (name, extra) = item
        obj = os.path.join(obj_dir, obj_name)
        # an object is reused while the sources it includes (transitively) and its flags are unchanged
        stamp, digest = obj + ".sha256", _unit_digest(name, extra, defines)
        if not ptxas_info and os.path.exists(obj) and os.path.exists(stamp):
            with open(stamp) as f:
                if f.read().strip() == digest:
                    return obj
        cmd = [nvcc, *ARCH, *COMMON, *extra, *[f"-D{d}" for d in defines], "-c", os.path.join(CSRC, name), "-o", obj]
        if ptxas_info:
            cmd[1:1] = ["-Xptxas", "-v"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {name}:\n{r.stdout}\n{r.stderr}")
        if verbose or ptxas_info:
            sys.stderr.write(r.stderr)
        with open(stamp, "w") as f:
            f.write(digest)
        return obj

    # longest units first; one nvcc per core
    order = sorted(UNITS.items(), key=lambda kv: (0 if "kind" in kv[0] else 1, kv[0]))
    with ThreadPoolExecutor(max_workers=max(2, os.cpu_count() or 4)) as pool:
        objs = list(pool.map(compile_unit, order))
    # link next to the target and rename into place: a process that is dlopen'ing the old library
    # keeps its (unlinked) file, nobody ever maps a half-written one
    tmp_path = f"{lib_path}.tmp{os.getpid()}"
    link = [nvcc, *ARCH, "-shared", "-Xcompiler", "-fPIC", "-o", tmp_path, *objs]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp_path, lib_path)
    if not variant:
        with open(LIB_PATH + ".sha256.tmp", "w") as f:
            f.write(_source_digest())
        os.replace(LIB_PATH + ".sha256.tmp", LIB_PATH + ".sha256")
    return lib_path


if __name__ == "__main__":
    _defs = tuple(a[2:] for a in sys.argv if a.startswith("-D"))
    _out = next((a[6:] for a in sys.argv if a.startswith("--out=")), None)
    print(build_library(force="--force" in sys.argv, verbose=True, ptxas_info="--ptxas" in sys.argv,
                        defines=_defs, out=_out))
