"""eqtlbma_b200 -- B200-native hot path of eqtlbma_bf (timflutre/eqtlbma v1.3.3).

The product is the C-ABI shared library ``libeqtlbma_b200.so`` (include/eqtlbma_b200.h) built from
``csrc/`` for sm_100a; this package is the thin Python host mirror used by the tests and the
benchmark: it builds / loads the library and exposes :class:`Engine`, a ctypes wrapper whose
methods map one-to-one onto the C entry points.  There is no CPU fallback: loading fails loudly
when the CUDA library is missing.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

from ._capi import Engine as _Engine

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "libeqtlbma_b200.so")
HOST_BIN = os.path.join(_HERE, "eqtlbma_bf")
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC"]

_lib = None


def _digest(paths, extra=""):
    import hashlib
    h = hashlib.sha256(extra.encode())
    for p in paths:
        h.update(os.path.basename(p).encode())
        with open(p, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def build_library(force: bool = False) -> str:
    """Compile every csrc/*.cu for sm_100a (one object per translation unit, in parallel) and link them into
    libeqtlbma_b200.so (in-tree).  Freshness is decided by CONTENT hashes kept next to the objects (file times do not
    survive the snapshot that carries the tree to the GPU box), so an unchanged tree is never recompiled there and a
    changed source always is."""
    from concurrent.futures import ThreadPoolExecutor
    src_dir = os.path.join(_HERE, "csrc")
    obj_dir = os.path.join(src_dir, "build")
    os.makedirs(obj_dir, exist_ok=True)
    names = sorted(os.listdir(src_dir))
    units = [os.path.join(src_dir, f) for f in names if f.endswith(".cu")]

    def deps(path, seen):
        """Project headers a translation unit includes (quoted includes, followed recursively)."""
        import re
        for inc in re.findall(r'^\s*#\s*include\s+"([^"]+)"', open(path).read(), flags=re.M):
            h = os.path.normpath(os.path.join(os.path.dirname(path), inc))
            if os.path.exists(h) and h not in seen:
                seen.append(h)
                deps(h, seen)
        return seen

    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    flags = NVCC_FLAGS + (["-DEQB_TUNING"] if os.environ.get("EQB_BUILD_TUNING") else [])
    todo, objs = [], []
    for u in units:
        obj = os.path.join(obj_dir, os.path.basename(u)[:-3] + ".o")
        objs.append(obj)
        want = _digest([u] + sorted(deps(u, [])), " ".join(flags))
        stamp = obj + ".sha256"
        have = open(stamp).read().strip() if os.path.exists(stamp) and os.path.exists(obj) else ""
        if force or have != want:
            todo.append((u, obj, stamp, want))

    def compile_one(job):
        u, obj, stamp, want = job
        subprocess.check_call([nvcc] + flags + ["-ccbin", "g++", "-c", "-o", obj, u])
        with open(stamp, "w") as fh:
            fh.write(want)

    if todo:
        with ThreadPoolExecutor(max_workers=max(1, min(len(todo), os.cpu_count() or 1))) as ex:
            list(ex.map(compile_one, todo))
    link_stamp = LIB_PATH + ".sha256"
    want = _digest(objs)
    have = open(link_stamp).read().strip() if os.path.exists(link_stamp) and os.path.exists(LIB_PATH) else ""
    if todo or have != want:
        subprocess.check_call([nvcc, "-shared", "-ccbin", "g++", "-o", LIB_PATH] + objs)
        with open(link_stamp, "w") as fh:
            fh.write(want)
    return LIB_PATH


def load_library() -> ctypes.CDLL:
    """Load the CUDA library; raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(the eqtlbma_b200 hot path has no CPU fallback)")
        _lib = ctypes.CDLL(LIB_PATH)
    return _lib


class Engine(_Engine):
    """Context of the CUDA library (prefix ``eqb_``)."""

    def __init__(self, ds, **kw):
        super().__init__(load_library(), "eqb_", ds, **kw)

    def raw_abfs_device(self):
        """(device address, n_pairs, gene ids, gene offsets) of the raw ABFs of the last chunk of the true pass
        (eqb_raw_abfs_device): what HmEngine.append_device takes."""
        import numpy as np
        ptr, n_pairs, n_genes = ctypes.c_void_p(), ctypes.c_int64(0), ctypes.c_int64(0)
        self._call("raw_abfs_device", ctypes.byref(ptr), ctypes.byref(n_pairs), ctypes.byref(n_genes))
        ids = np.zeros(n_genes.value, dtype=np.int64)
        off = np.zeros(n_genes.value + 1, dtype=np.int64)
        self._call("raw_abfs_layout", ids.ctypes.data_as(ctypes.c_void_p), off.ctypes.data_as(ctypes.c_void_p))
        return ptr.value, n_pairs.value, ids, off

    def launch_count(self) -> int:
        f = self.lib.eqb_launch_count
        f.restype = ctypes.c_int64
        return int(f(self.ctx))

    def last_pair_kernel_ms(self) -> float:
        f = self.lib.eqb_last_pair_kernel_ms
        f.restype = ctypes.c_float
        return float(f(self.ctx))

    def fast_gene_count(self) -> int:
        f = self.lib.eqb_fast_gene_count
        f.restype = ctypes.c_int64
        return int(f(self.ctx))

    def set_perm_timing(self, on=True):
        self._call("set_perm_timing", ctypes.c_int32(int(on)))

    def last_perm_timing(self) -> dict:
        """Per-kernel device time of the last permutation run (enable with set_perm_timing before the run)."""
        out = (ctypes.c_double * 8)()
        self._call("last_perm_timing", out)
        keys = ["prep_ms", "gemm_ms", "bf_ms", "merge_ms", "gemm_flops", "gemm_useful_flops", "items", "path"]
        return dict(zip(keys, list(out)))

    def run_device_only(self, lo=0, hi=None, raw=False) -> float:
        hi = self.ds.n_genes if hi is None else hi
        ms = ctypes.c_float(0)
        self._call("run_device_only", ctypes.c_int64(lo), ctypes.c_int64(hi), ctypes.c_int32(int(raw)), ctypes.byref(ms))
        return float(ms.value)

    def run_permutations_device_only(self, nperm, seed, lo=0, hi=None, **kw) -> float:
        hi = self.ds.n_genes if hi is None else hi
        pc = self.perm_config(nperm, seed, **kw)
        ms = ctypes.c_float(0)
        self._call("run_permutations_device_only", ctypes.c_int64(lo), ctypes.c_int64(hi), ctypes.byref(pc),
                   ctypes.byref(ms))
        return float(ms.value)


def measure_fp64_peaks(device: int = 0) -> dict:
    """DFMA / DMMA peaks of the device (register-resident loops): the denominators of the FP64 rooflines."""
    lib = load_library()
    out = (ctypes.c_double * 4)()
    f = lib.eqb_measure_fp64_peaks
    f.restype = ctypes.c_int
    rc = f(ctypes.c_int32(device), out)
    if rc != 0:
        raise RuntimeError(f"eqb_measure_fp64_peaks failed ({rc})")
    return {"dfma_tflops": out[0], "dmma_tflops": out[1], "implied_sm_mhz": out[2], "n_sm": int(out[3])}


def selftest_perm_gemm(n_rows: int, n_cols: int, ldn: int, device: int = 0) -> dict:
    """Self-test and throughput of the TMA / DMMA product kernel of the permutation path."""
    lib = load_library()
    out = (ctypes.c_double * 3)()
    f = lib.eqb_selftest_perm_gemm
    f.restype = ctypes.c_int
    rc = f(ctypes.c_int32(device), ctypes.c_int64(n_rows), ctypes.c_int64(n_cols), ctypes.c_int32(ldn), out)
    if rc != 0:
        raise RuntimeError(f"eqb_selftest_perm_gemm failed ({rc})")
    return {"worst_rel_err": out[0], "tflops": out[1], "tiles": int(out[2])}
