"""Ahead-of-time, in-tree build of the native library ``cpd_b200/_C*.so``.

The reference JIT-compiles at import with no arch flags (CPDtorch/quant/quant_function.py:9-19).
Here every CUDA source is compiled explicitly for ``sm_100a``
(``-gencode arch=compute_100a,code=sm_100a -lineinfo``), objects are cached under
``build/`` by content hash, and the shared object lands next to the python package so
it travels with the source tree (no JIT cache, no site-packages install).

    python -m cpd_b200.build            # incremental
    python -m cpd_b200.build --force    # rebuild everything
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
BUILD = ROOT / "build" / "cpd_b200"

CUDA_SOURCES = [
    "cuda/quantize_sm100.cu",
    "cuda/aps_fold_sm100.cu",
    "cuda/lars_sm100.cu",
    "cuda/gemm_exact_sm100.cu",
    "cuda/gemm_tcgen05_sm100a.cu",
    "cuda/gemm_tcgen05_2cta_sm100a.cu",
    "cuda/gemm_tcgen05_mx_sm100a.cu",
    "cuda/mx_quantize_sm100.cu",
    "cuda/operand_stage_sm100.cu",
    "cuda/bn_act_sm100.cu",
    "cuda/allreduce_fused_sm100a.cu",
]
CPP_SOURCES = [
    "cpu/cpu_ops.cpp",
    "comm/symm_mem.cpp",
    "comm/topo.cpp",
]
TORCH_SOURCES = [
    "bindings.cpp",      # pybind11 only: raw pointers in, no libtorch dependency
]

ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _cuda_home() -> str:
    for c in (os.environ.get("CUDA_HOME"), "/usr/local/cuda"):
        if c and Path(c, "bin", "nvcc").exists():
            return c
    nvcc = shutil.which("nvcc")
    if nvcc:
        return str(Path(nvcc).resolve().parent.parent)
    raise RuntimeError("cpd_b200.build: nvcc not found (set CUDA_HOME)")


def so_path() -> Path:
    suffix = sysconfig.get_config_var("EXT_SUFFIX") or ".so"
    return PKG / f"_C{suffix}"


def _hash_inputs(src: Path, cmd: list[str]) -> str:
    h = hashlib.sha256()
    h.update(" ".join(cmd).encode())
    h.update(src.read_bytes())
    # headers: hash every header in csrc (cheap, and correct for our small tree)
    for hdr in sorted(CSRC.rglob("*.h")) + sorted(CSRC.rglob("*.cuh")) + sorted(CSRC.rglob("*.inc")):
        h.update(hdr.read_bytes())
    return h.hexdigest()[:20]


def _compile(src_rel: str, cmd_prefix: list[str], force: bool, verbose: bool) -> Path:
    src = CSRC / src_rel
    obj = BUILD / (src_rel.replace("/", "_") + ".o")
    stamp = obj.with_suffix(".hash")
    cmd = cmd_prefix + ["-c", str(src), "-o", str(obj)]
    digest = _hash_inputs(src, cmd_prefix)
    if not force and obj.exists() and stamp.exists() and stamp.read_text() == digest:
        return obj
    if verbose:
        print(f"[cpd_b200.build] compiling {src_rel}", flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"cpd_b200.build: compiling {src_rel} failed")
    if verbose and "-v" in " ".join(cmd_prefix) and r.stderr:
        (BUILD / (src_rel.replace("/", "_") + ".ptxas.txt")).write_text(r.stderr)
    stamp.write_text(digest)
    return obj


def build(force: bool = False, verbose: bool = True) -> Path:
    """Build (if needed) and return the path of the shared object."""
    import pybind11

    BUILD.mkdir(parents=True, exist_ok=True)
    cuda_home = _cuda_home()
    nvcc = str(Path(cuda_home, "bin", "nvcc"))
    cxx = os.environ.get("CXX", "g++")
    py_inc = sysconfig.get_paths()["include"]
    common_defs: list[str] = []
    nvcc_cmd = [nvcc, *ARCH_FLAGS, "-O3", "-std=c++17", "-lineinfo", "-Xptxas", "-v", "--expt-relaxed-constexpr",
                "-Xcompiler", "-fPIC", f"-I{CSRC}", *common_defs]
    cpp_cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-Wall", f"-I{CSRC}",
               f"-I{cuda_home}/include", *common_defs]
    torch_cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-fvisibility=hidden", f"-I{CSRC}", f"-I{cuda_home}/include",
                 f"-I{py_inc}", f"-I{pybind11.get_include()}", *common_defs]

    jobs = [(s, nvcc_cmd) for s in CUDA_SOURCES] + [(s, cpp_cmd) for s in CPP_SOURCES] + \
           [(s, torch_cmd) for s in TORCH_SOURCES]
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
        objs = list(ex.map(lambda j: _compile(j[0], j[1], force, verbose), jobs))

    out = so_path()
    link_stamp = BUILD / "link.hash"
    digest = hashlib.sha256(b"".join(o.with_suffix(".hash").read_bytes() for o in objs) + b"link-v2").hexdigest()
    if force or not out.exists() or not link_stamp.exists() or link_stamp.read_text() != digest:
        if verbose:
            print(f"[cpd_b200.build] linking {out.name}", flush=True)
        # static CUDA runtime (nvcc's own default): the kernels were built by this toolkit, and the
        # library must not depend on which libcudart the host framework happens to have loaded
        # -Bsymbolic + --exclude-libs: the static runtime's cuda* symbols stay private to this library, so
        # they can neither interpose nor be interposed by the libcudart that PyTorch / NCCL loaded
        link = [cxx, "-shared", "-o", str(out), *map(str, objs), f"-L{cuda_home}/lib64", "-lcudart_static", "-lrt",
                "-ldl", "-lpthread", "-Wl,-Bsymbolic", "-Wl,--exclude-libs,ALL"]
        r = subprocess.run(link, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("cpd_b200.build: link failed")
        link_stamp.write_text(digest)
    return out


if __name__ == "__main__":
    p = build(force="--force" in sys.argv)
    print(p)
