"""Gene sharding over the GPUs of one box: contiguous ranges of whole write-groups,
balanced on cis-window cost; no collective on the data path, a final host gather of the per-gene /
per-pair results in gene order.  Thin wrapper over the C entry point eqb_partition_by_cost."""
from __future__ import annotations

import ctypes

import numpy as np


def gene_costs(cis_begin, cis_end, nperm=0):
    """cost model: cis SNPs x (1 + permutations) pair evaluations per gene"""
    return ((np.asarray(cis_end) - np.asarray(cis_begin)).astype(np.int64) * (1 + int(nperm))).astype(np.int64)


def partition(lib: ctypes.CDLL, costs, wrtsize: int, n_shards: int) -> np.ndarray:
    costs = np.ascontiguousarray(costs, dtype=np.int64)
    out = np.zeros(n_shards + 1, dtype=np.int64)
    f = lib.eqb_partition_by_cost
    f.restype = ctypes.c_int
    rc = f(costs.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(len(costs)), ctypes.c_int64(wrtsize),
           ctypes.c_int32(n_shards), out.ctypes.data_as(ctypes.c_void_p))
    if rc != 0:
        raise RuntimeError("eqb_partition_by_cost failed")
    return out
