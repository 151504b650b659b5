"""Loader for the in-tree native library (``quintnet_b200/lib/libquintnet_b200.so``).

The kernels are plain CUDA C++ with a C ABI (no torch headers → seconds to
compile, one ``nvcc -gencode arch=compute_100a,code=sm_100a`` invocation per
file); Python calls them through ctypes with raw device pointers and the
current CUDA stream.  ``build()`` compiles them; ``lib()`` loads them.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path
from typing import List, Optional

import torch

PKG = Path(__file__).resolve().parent.parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"
LIBNAME = "libquintnet_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "--use_fast_math",
]

_lib: Optional[ctypes.CDLL] = None
_load_error: Optional[str] = None

# ops whose native kernel is used on CUDA tensors.  Env override:
#   QUINTNET_NATIVE=all | none | comma-separated list
_DEFAULT_ENABLED = {
    "layernorm", "gelu", "adamw", "dropout", "rope", "cross_entropy", "gemm", "attention",
    "allreduce", "zero1", "tp_fused",
}


def _enabled_set():
    v = os.environ.get("QUINTNET_NATIVE", "").strip()
    if not v or v == "all":
        dis = {s for s in os.environ.get("QUINTNET_NATIVE_DISABLE", "").split(",") if s}
        extra = {s for s in os.environ.get("QUINTNET_NATIVE_ENABLE", "").split(",") if s}
        return (_DEFAULT_ENABLED | extra) - dis
    if v == "none":
        return set()
    return {s.strip() for s in v.split(",") if s.strip()}


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def sources() -> List[Path]:
    return sorted(CSRC.glob("*.cu"))


def build(force: bool = False, verbose: bool = True, jobs: int = 8) -> Path:
    """Compile every ``csrc/*.cu`` for sm_100a and link the shared library."""
    LIBDIR.mkdir(exist_ok=True)
    objdir = LIBDIR / "obj"
    objdir.mkdir(exist_ok=True)
    out = LIBDIR / LIBNAME
    srcs = sources()
    hdrs = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))
    newest_hdr = max([h.stat().st_mtime for h in hdrs], default=0.0)
    procs, objs = [], []
    for s in srcs:
        o = objdir / (s.stem + ".o")
        objs.append(o)
        if not force and o.exists() and o.stat().st_mtime > max(s.stat().st_mtime, newest_hdr):
            continue
        # every artefact is written under a temporary name and renamed into place: a snapshot of the tree (gpurun
        # copies it to the GPU box at an arbitrary moment) never sees a half-written object or library
        tmp = o.with_suffix(".o.tmp")
        cmd = [nvcc_path(), *NVCC_FLAGS, "-I", str(CSRC), "-c", str(s), "-o", str(tmp)]
        if verbose:
            print("[quintnet_b200.build]", " ".join(cmd), flush=True)
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT), tmp, o))
        if len(procs) >= jobs:
            _wait(procs)
    _wait(procs)
    need_link = force or not out.exists() or any(o.stat().st_mtime > out.stat().st_mtime for o in objs) \
        or not _loadable(out)
    if need_link:
        tmp = out.with_suffix(".so.tmp")
        cmd = [nvcc_path(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(tmp),
               *[str(o) for o in objs], "-lcudart"]
        if verbose:
            print("[quintnet_b200.build]", " ".join(cmd), flush=True)
        subprocess.check_call(cmd)
        os.replace(tmp, out)
    return out


def _loadable(path: Path) -> bool:
    """ELF sanity check without dlopen (a CPU-only box may lack libcudart): magic + section table inside the file."""
    try:
        with open(path, "rb") as f:
            head = f.read(64)
            if len(head) < 64 or head[:4] != b"\x7fELF":
                return False
            import struct
            shoff, = struct.unpack_from("<Q", head, 0x28)
            shentsize, shnum = struct.unpack_from("<HH", head, 0x3A)
            return path.stat().st_size >= shoff + shentsize * shnum
    except OSError:
        return False


def _wait(procs):
    while procs:
        s, p, tmp, o = procs.pop(0)
        outp = p.communicate()[0].decode()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed for {s}:\n{outp}")
        os.replace(tmp, o)
        if outp.strip():
            print(outp)


def lib_path() -> Path:
    return LIBDIR / LIBNAME


def available() -> bool:
    return lib() is not None


def lib() -> Optional[ctypes.CDLL]:
    global _lib, _load_error
    if _lib is not None or _load_error is not None:
        return _lib
    p = lib_path()
    if not p.exists():
        _load_error = f"{p} does not exist (run `python -c 'import __graft_entry__ as g; g.build()'`)"
        return None
    try:
        _lib = ctypes.CDLL(str(p))
    except OSError as e:  # e.g. no libcudart on a CPU-only box
        _load_error = str(e)
        return None
    _declare(_lib)
    return _lib


_I, _L, _F, _P = ctypes.c_int, ctypes.c_longlong, ctypes.c_float, ctypes.c_void_p
_U = ctypes.c_ulonglong
_SIGNATURES = {
    "qn_comm_set_flag_watchdog": [ctypes.c_longlong],
    "qn_push_bench": [_P, _P, _L, _I, _I, _I, _I, _L, _P],
    "qn_push_bench_bulk": [_P, _P, _L, _I, _I, _P],
    "qn_gemm_set_flag_watchdog": [ctypes.c_longlong],
    "qn_gemm_bf16_wgrad_accum": [_P, _P, _P, _P, _I, _I, _I, _L, _L, _L, _I, _I, _P],
    "qn_gemm_bf16_wgrad_accum_f32": [_P, _P, _P, _I, _I, _I, _L, _L, _L, _I, _I, _P],
    "qn_small_attn_fits": [_I, _I],
    "qn_small_attn_fwd": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _F, _I, _P],
    "qn_small_attn_bwd": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _F, _I, _P],
    "qn_write_flag": [_P, _L, ctypes.c_uint, _P],
    "qn_ag_done_base": [],
    "qn_gemm_fp8": [_P, _P, _P, _I, _I, _I, _L, _L, _L, _P, _P, _I, _P, _P, _L, _I, _P],
    "qn_quantize_rowwise_e4m3": [_P, _P, _P, _I, _I, _L, _L, _P],
    "qn_quantize_mxfp8": [_P, _P, _P, _I, _I, _L, _L, _P],
    "qn_gemm_mxfp8": [_P, _P, _P, _P, _P, _I, _I, _I, _L, _L, _L, _P, _P],
    "qn_gemm_bf16": [_P, _P, _P, _I, _I, _I, _L, _L, _L, _I, _I, _I, _P, _P, _L, _I, _I, _P],
    "qn_gemm_bf16_splitk": [_P, _P, _P, _P, _I, _I, _I, _L, _L, _L, _I, _I, _I, _I, _P],
    "qn_layernorm_fwd": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _F, _I, _P],
    "qn_layernorm_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "qn_dropout_add_layernorm_fwd": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _F, _F, _U, _U, _P],
    "qn_dropout_add_layernorm_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _U, _U, _P, _P, _P, _P],
    "qn_colsum_accum_f32": [_P, _P, _I, _I, _L, _P],
    "qn_dropout_add_fwd": [_P, _P, _P, _L, _F, _U, _U, _P],
    "qn_dropout_bwd": [_P, _P, _L, _F, _U, _U, _P],
    "qn_gelu_fwd": [_P, _P, _L, _P],
    "qn_gelu_bwd": [_P, _P, _P, _L, _P],
    "qn_colsum": [_P, _P, _P, _I, _I, _L, _P],
    "qn_rope": [_P, _P, _P, _P, _I, _I, _I, _I, _L, _L, _L, _I, _P],
    "qn_adamw": [_P, _I, _P, _I, _P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _F, _F, _P],
    "qn_cross_entropy": [_P, _P, _P, _P, _I, _I, _L, _L, _I, _P],
    "qn_scale_bf16": [_P, _P, _P, _L, _P],
    "qn_attn_fwd": [_P, _P, _P, _I, _I, _I, _F, _I, _F, _U, _P],
    "qn_attn_bwd": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _I, _F, _U, _P],
}


def _declare(lib) -> None:
    for name, args in _SIGNATURES.items():
        if hasattr(lib, name):
            fn = getattr(lib, name)
            fn.argtypes = args
            fn.restype = ctypes.c_int


def declare(name: str, args) -> None:
    """Register the signature of an extra exported function (used by comm ops)."""
    _SIGNATURES[name] = args
    if _lib is not None and hasattr(_lib, name):
        fn = getattr(_lib, name)
        fn.argtypes = args
        fn.restype = ctypes.c_int


def require() -> ctypes.CDLL:
    l = lib()
    if l is None:
        raise RuntimeError(f"quintnet_b200 native library unavailable: {_load_error}")
    return l


def op_enabled(name: str) -> bool:
    """True if CUDA tensors should use the native kernel for ``name``."""
    if name not in _enabled_set():
        return False
    if lib() is None:
        if os.environ.get("QUINTNET_ALLOW_FALLBACK", "0") == "1":
            return False
        raise RuntimeError(
            f"native op '{name}' requested on a CUDA tensor but the native library is not "
            f"loadable: {_load_error}.  Build it (graft build()) or set QUINTNET_ALLOW_FALLBACK=1."
        )
    return True


# ------------------------------------------------------------- ctypes helpers
def ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def check(code: int, what: str) -> None:
    if code != 0:
        raise RuntimeError(f"{what} failed with CUDA error {code}")


# launch counter (bench.py reports gpu_launches)
launch_count = 0


def count(n: int = 1) -> None:
    global launch_count
    launch_count += n
