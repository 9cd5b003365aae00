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
              "-Xcompiler", "-fPIC", "-shared"]

_lib = None


def build_library(force: bool = False) -> str:
    """Compile csrc/eqtlbma_b200.cu for sm_100a into libeqtlbma_b200.so (in-tree)."""
    src_dir = os.path.join(_HERE, "csrc")
    srcs = [os.path.join(src_dir, f) for f in sorted(os.listdir(src_dir))] + \
           [os.path.join(ROOT, "include", "eqtlbma_b200.h")]
    if not force and os.path.exists(LIB_PATH) and \
            all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-ccbin", "g++", "-o", LIB_PATH, os.path.join(src_dir, "eqtlbma_b200.cu")]
    subprocess.check_call(cmd)
    return LIB_PATH


def load_library() -> ctypes.CDLL:
    """Load the CUDA library; raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is None:
        path = os.environ.get("EQB_LIB", LIB_PATH)  # tuning builds (variants/), never a different implementation
        if path != LIB_PATH:
            _lib = ctypes.CDLL(path)
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(the eqtlbma_b200 hot path has no CPU fallback)")
        _lib = ctypes.CDLL(LIB_PATH)
    return _lib


class Engine(_Engine):
    """Context of the CUDA library (prefix ``eqb_``)."""

    def __init__(self, ds, **kw):
        super().__init__(load_library(), "eqb_", ds, **kw)

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
