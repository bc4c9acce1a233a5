"""In-tree builds of libnqb.so and of the per-signature kernel libraries (sm_100a only).

``nvcc -gencode arch=compute_100a,code=sm_100a`` cross-compiles without a GPU, so
``__graft_entry__.build()`` runs this on the CPU box and the resulting ``.so`` files
travel to the B200 box with the repo snapshot.  At run time a signature that has no
prebuilt library is generated and compiled on the spot (the same thing the
reference's OpenEquivariance backend does with its JIT, nequip/nn/_tp_scatter_oeq.py:29-47);
if ``nvcc`` is missing that is a hard error -- there is no CPU fallback.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import threading
from concurrent.futures import ThreadPoolExecutor
from typing import Iterable, List, Optional, Tuple

from .codegen import CODEGEN_VERSION, GenOptions, TPSignature, generate

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIBDIR = os.path.join(_HERE, "lib")
GENDIR = os.path.join(_HERE, "_gen")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")

ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON_FLAGS = ["-O3", "-lineinfo", "-std=c++17", "-shared", "-Xcompiler", "-fPIC"]

_lock = threading.Lock()


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError(
            "nequip_b200: nvcc not found -- the B200 kernels cannot be built and there is no CPU fallback"
        )
    return p


def _run(cmd: List[str]):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nequip_b200 build failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return r.stdout + r.stderr


def _newer(src_files: Iterable[str], target: str) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in src_files)


def runtime_lib_path() -> str:
    return os.path.join(LIBDIR, "libnqb.so")


def ensure_runtime(force: bool = False) -> str:
    """Build (if stale) and return the path of libnqb.so."""
    out = runtime_lib_path()
    cus = [os.path.join(CSRC, n) for n in ("nqb_runtime.cu", "nqb_mlp.cu", "nqb_gemm.cu", "nqb_gemm_t.cu", "nqb_nl.cu")]
    srcs = cus + [os.path.join(INCLUDE, "nqb.h"), os.path.join(CSRC, "nqb_tc.cuh")]
    with _lock:
        if force or _newer(srcs, out):
            os.makedirs(LIBDIR, exist_ok=True)
            tmp = out + f".tmp{os.getpid()}"
            extra = os.environ.get("NQB_EXTRA_NVCC_FLAGS", "").split()  # e.g. -DNQB_GEMM_PROF (tools/bench_gemm.py --prof)
            _run([nvcc_path(), *ARCH_FLAGS, *COMMON_FLAGS, *extra, "-I", INCLUDE, "-Xptxas", "-v", "-o", tmp, *cus, "-ldl"])
            os.replace(tmp, out)
    return out


def _device_header_hash() -> str:
    with open(os.path.join(CSRC, "nqb_tp_device.cuh"), "rb") as f:
        return hashlib.sha1(f.read()).hexdigest()[:8]


def spec_lib_path(sig: TPSignature, opts: Optional[GenOptions] = None) -> str:
    opts = opts or GenOptions()
    return os.path.join(LIBDIR, f"nqbspec_{sig.key(opts)}_{_device_header_hash()}.so")


def ensure_spec(sig: TPSignature, opts: Optional[GenOptions] = None, force: bool = False, verbose: bool = False) -> str:
    """Generate + compile the kernel library of one signature (cached in-tree)."""
    opts = opts or GenOptions()
    out = spec_lib_path(sig, opts)
    if os.path.exists(out) and not force:
        return out
    with _lock:
        if os.path.exists(out) and not force:
            return out
        os.makedirs(LIBDIR, exist_ok=True)
        os.makedirs(GENDIR, exist_ok=True)
        # several ranks may JIT the same signature at once: every process writes its own source and
        # temporary library, the final rename is atomic
        cu = os.path.join(GENDIR, os.path.basename(out)[:-3] + f".{os.getpid()}.cu")
        with open(cu, "w") as f:
            f.write(generate(sig, opts))
        final_cu = os.path.join(GENDIR, os.path.basename(out)[:-3] + ".cu")
        tmp = out + f".tmp{os.getpid()}"
        cmd = [nvcc_path(), *ARCH_FLAGS, *COMMON_FLAGS, "-I", CSRC, "-o", tmp, cu]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        log = _run(cmd)
        os.replace(tmp, out)
        os.replace(cu, final_cu)
        if verbose:
            print(log)
    return out


def ensure_specs(sigs: Iterable[Tuple[TPSignature, Optional[GenOptions]]], jobs: int = 0) -> List[str]:
    """Parallel build of many signatures (used by __graft_entry__.build)."""
    todo = list(sigs)
    jobs = jobs or min(8, os.cpu_count() or 1)

    def one(item):
        sig, opts = item
        opts = opts or GenOptions()
        out = spec_lib_path(sig, opts)
        if os.path.exists(out):
            return out
        os.makedirs(LIBDIR, exist_ok=True)
        os.makedirs(GENDIR, exist_ok=True)
        cu = os.path.join(GENDIR, os.path.basename(out)[:-3] + ".cu")
        with open(cu, "w") as f:
            f.write(generate(sig, opts))
        tmp = out + f".tmp{os.getpid()}_{threading.get_ident()}"
        _run([nvcc_path(), *ARCH_FLAGS, *COMMON_FLAGS, "-I", CSRC, "-o", tmp, cu])
        os.replace(tmp, out)
        return out

    with ThreadPoolExecutor(max_workers=jobs) as ex:
        return list(ex.map(one, todo))


__all__ = [
    "ensure_runtime",
    "ensure_spec",
    "ensure_specs",
    "spec_lib_path",
    "runtime_lib_path",
    "CODEGEN_VERSION",
]
