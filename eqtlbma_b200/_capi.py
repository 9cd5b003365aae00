"""ctypes binding of the C ABI declared in include/eqtlbma_b200.h.

The binding is symbol-prefix agnostic so that the test-suite can drive the CPU oracle
(oracle/liboracle.so, prefix ``eqo_``) through the very same Python code as the product library
(libeqtlbma_b200.so, prefix ``eqb_``).  The package itself only ever loads the product library.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

ABI_VERSION = 1
ANALYSIS = {"sep": 0, "join": 1}
BFS = {"gen": 0, "sin": 1, "all": 2}
PBF = {"none": 0, "gen": 1, "gen-sin": 2, "all": 3}
ERROR = {"uvlr": 0, "mvlr": 1, "hybrid": 2}
ANCHOR = {"TSS": 0, "TSS+TES": 1}


class Config(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("n_subgroups", C.c_int32), ("n_samples_all", C.c_int32),
                ("analysis", C.c_int32), ("n_snps", C.c_int64), ("n_genes", C.c_int64),
                ("bfs", C.c_int32), ("error_model", C.c_int32), ("qnorm", C.c_int32),
                ("device", C.c_int32), ("fiterr", C.c_double)]


class SubgroupC(C.Structure):
    _fields_ = [("geno_id", C.c_int32), ("n_exp_cols", C.c_int32), ("n_covariates", C.c_int32),
                ("n_cov_cols", C.c_int32), ("all2geno", C.c_void_p), ("all2exp", C.c_void_p),
                ("all2cov", C.c_void_p), ("snp_has_geno", C.c_void_p), ("gene_has_exp", C.c_void_p),
                ("Y", C.c_void_p), ("C", C.c_void_p)]


class ResultsC(C.Structure):
    _fields_ = [("n", C.c_void_p), ("sstats", C.c_void_p), ("abf_gen", C.c_void_p),
                ("abf_cfg", C.c_void_p), ("abf_w", C.c_void_p), ("gene_analyzed", C.c_void_p)]


class PermConfigC(C.Structure):
    _fields_ = [("nperm", C.c_int64), ("seed", C.c_uint64), ("trick", C.c_int32), ("tricut", C.c_int32),
                ("permsep", C.c_int32), ("pbf", C.c_int32), ("maxbf", C.c_int32), ("wrtsize", C.c_int32)]


class PermResultsC(C.Structure):
    _fields_ = [("pval", C.c_void_p), ("nperm_done", C.c_void_p), ("count", C.c_void_p),
                ("true_stat", C.c_void_p), ("median_perm", C.c_void_p), ("perm_stats", C.c_void_p)]


@dataclass
class TrueResults:
    offsets: np.ndarray  # [genes+1]
    gene_analyzed: np.ndarray
    n: np.ndarray  # [pairs,S]
    sstats: np.ndarray  # [pairs,S,5]
    abf_gen: np.ndarray | None  # [pairs,3,L]
    abf_cfg: np.ndarray | None  # [pairs,C,K]
    abf_w: np.ndarray | None  # [pairs,5+C]


@dataclass
class PermResults:
    pval: np.ndarray
    nperm_done: np.ndarray
    count: np.ndarray
    true_stat: np.ndarray
    median_perm: np.ndarray
    perm_stats: np.ndarray | None


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Engine:
    """One context of the C ABI (``prefix`` selects eqb_ = CUDA library, eqo_ = CPU oracle)."""

    timing = None  # set to a dict to accumulate host wall time per C entry point (diagnostics)

    def __init__(self, lib: C.CDLL, prefix: str, ds, analysis="join", bfs="all", error="uvlr",
                 fiterr=0.5, qnorm=False, device=0, pinned=False):
        self.lib, self.prefix = lib, prefix
        self.ds = ds
        self.S = len(ds.subgroups)
        self.analysis, self.bfs, self.error = analysis, bfs, error
        self._keep = []
        cfg = Config(ABI_VERSION, self.S, ds.n_all, ANALYSIS[analysis], ds.n_snps, ds.n_genes,
                     BFS[bfs], ERROR[error], int(qnorm), device, float(fiterr))
        self.ctx = C.c_void_p()
        self._call("create", C.byref(self.ctx), C.byref(cfg), ctx_first=False)
        for s, sg in enumerate(ds.subgroups):
            arrs = dict(all2geno=np.ascontiguousarray(sg.all2geno, dtype=np.int32),
                        all2exp=np.ascontiguousarray(sg.all2exp, dtype=np.int32),
                        all2cov=np.ascontiguousarray(sg.all2cov, dtype=np.int32),
                        snp=np.ascontiguousarray(sg.snp_has_geno, dtype=np.uint8),
                        gene=np.ascontiguousarray(sg.gene_has_exp, dtype=np.uint8),
                        Y=np.ascontiguousarray(sg.Y, dtype=np.float64),
                        Cc=np.ascontiguousarray(sg.C, dtype=np.float64))
            self._keep.append(arrs)
            Q = sg.C.shape[0]
            sc = SubgroupC(sg.geno_id, sg.Y.shape[1], Q, sg.C.shape[1] if Q else 0, _ptr(arrs["all2geno"]),
                           _ptr(arrs["all2exp"]), _ptr(arrs["all2cov"]), _ptr(arrs["snp"]),
                           _ptr(arrs["gene"]), _ptr(arrs["Y"]), _ptr(arrs["Cc"]) if Q else None)
            self._call("set_subgroup", C.c_int32(s), C.byref(sc))
        gl, gs = np.ascontiguousarray(ds.gridL), np.ascontiguousarray(ds.gridS)
        pl, ol = np.ascontiguousarray(gl[:, 0]), np.ascontiguousarray(gl[:, 1])
        ps, os_ = np.ascontiguousarray(gs[:, 0]), np.ascontiguousarray(gs[:, 1])
        self.L, self.K = len(pl), len(ps)
        self._call("set_grids", _ptr(pl), _ptr(ol), C.c_int32(self.L), _ptr(ps), _ptr(os_), C.c_int32(self.K))
        self.cis_begin = np.zeros(ds.n_genes, dtype=np.int64)
        self.cis_end = np.zeros(ds.n_genes, dtype=np.int64)
        gc = np.ascontiguousarray(ds.gene_chr, dtype=np.int32)
        gst = np.ascontiguousarray(ds.gene_start, dtype=np.int64)
        gen = np.ascontiguousarray(ds.gene_end, dtype=np.int64)
        sc_ = np.ascontiguousarray(ds.snp_chr, dtype=np.int32)
        sp = np.ascontiguousarray(ds.snp_pos, dtype=np.int64)
        self._call("build_cis_windows", _ptr(gc), _ptr(gst), _ptr(gen), _ptr(sc_), _ptr(sp),
                   C.c_int32(ANCHOR[ds.anchor]), C.c_int64(ds.radius), _ptr(self.cis_begin), _ptr(self.cis_end))
        # genotypes last: their upload is asynchronous and everything queued after it on the copy engine would wait
        denoms = getattr(ds, "geno_denoms", None) or [1.0] * len(ds.genos)
        for gi, G in enumerate(ds.genos):
            if G.dtype in (np.uint8, np.uint16):
                # compact lossless transport (eqb_set_genotypes_fixed): value = k / denom
                if prefix == "eqb_":
                    Gc = G if G.flags.c_contiguous else np.ascontiguousarray(G)
                    self._keep.append(Gc)
                    self._call("set_genotypes_fixed", C.c_int32(gi), Gc.ctypes.data_as(C.c_void_p),
                               C.c_int32(Gc.dtype.itemsize), C.c_double(float(denoms[gi])), C.c_int64(Gc.shape[0]),
                               C.c_int32(Gc.shape[1]))
                    continue
                G = G.astype(np.float64) / float(denoms[gi])  # the oracle takes doubles
            if getattr(ds, "_clean", False) and G.flags.c_contiguous and G.dtype == np.float64:
                Gc = G  # already NaN-free and contiguous (e.g. pinned by the caller): no host copy
            else:
                Gc = np.ascontiguousarray(np.nan_to_num(G, nan=0.0), dtype=np.float64)
            self._keep.append(Gc)  # the upload is asynchronous: valid until the first run returns
            self._call("set_genotypes", C.c_int32(gi), _ptr(Gc), C.c_int64(Gc.shape[0]), C.c_int32(Gc.shape[1]))
        self._call("finalize")
        f = getattr(lib, prefix + "n_configs")
        f.restype = C.c_int64
        self.n_configs = int(f(self.ctx))

    # ------------------------------------------------------------------
    def _call(self, name, *args, ctx_first=True):
        f = getattr(self.lib, self.prefix + name)
        f.restype = C.c_int
        if Engine.timing is not None:
            import time
            t0 = time.perf_counter()
            rc = f(self.ctx, *args) if ctx_first else f(*args)
            Engine.timing[name] = Engine.timing.get(name, 0.0) + time.perf_counter() - t0
        else:
            rc = f(self.ctx, *args) if ctx_first else f(*args)
        if rc != 0:
            e = getattr(self.lib, self.prefix + "last_error")
            e.restype = C.c_char_p
            msg = e(self.ctx).decode() if self.ctx else "create failed"
            raise RuntimeError(f"{self.prefix}{name} failed ({rc}): {msg}")

    def close(self):
        if self.ctx:
            f = getattr(self.lib, self.prefix + "destroy")
            f.restype = None
            f(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def pair_offsets(self, lo=0, hi=None):
        hi = self.ds.n_genes if hi is None else hi
        off = np.zeros(hi - lo + 1, dtype=np.int64)
        self._call("pair_offsets", C.c_int64(lo), C.c_int64(hi), _ptr(off))
        return off

    def alloc_results(self, lo=0, hi=None, raw=True, pinned=False) -> TrueResults:
        """Host result buffers for run(); pinned (page-locked, via torch) makes the D2H copies DMA."""
        hi = self.ds.n_genes if hi is None else hi
        off = self.pair_offsets(lo, hi)
        P, S, Cn = int(off[-1]), self.S, self.n_configs
        join = self.analysis == "join"

        def mk(shape, dtype, fill):
            if pinned:
                import torch
                t = torch.empty(shape, dtype={np.float64: torch.float64, np.int32: torch.int32,
                                              np.uint8: torch.uint8}[dtype]).pin_memory()
                self._keep.append(t)
                a = t.numpy()
                a[...] = fill
                return a
            return np.full(shape, fill, dtype=dtype)

        n = mk((P, S), np.int32, 0)
        ss = mk((P, S, 5), np.float64, np.nan)
        ga = mk((hi - lo,), np.uint8, 0)
        ag = mk((P, 3, self.L), np.float64, np.nan) if join and raw else None
        ac = mk((P, Cn, self.K), np.float64, np.nan) if join and raw else None
        aw = mk((P, 5 + Cn), np.float64, np.nan) if join else None
        return TrueResults(off, ga, n, ss, ag, ac, aw)

    def run(self, lo=0, hi=None, raw=True, out: TrueResults | None = None) -> TrueResults:
        hi = self.ds.n_genes if hi is None else hi
        r = out if out is not None else self.alloc_results(lo, hi, raw)
        rc = ResultsC(_ptr(r.n), _ptr(r.sstats), _ptr(r.abf_gen), _ptr(r.abf_cfg), _ptr(r.abf_w),
                      _ptr(r.gene_analyzed))
        self._call("run", C.c_int64(lo), C.c_int64(hi), C.byref(rc))
        return r

    def perm_config(self, nperm, seed, trick=0, tricut=10, permsep=0, pbf="none", maxbf=False, wrtsize=10):
        return PermConfigC(int(nperm), int(seed) & 0xFFFFFFFFFFFFFFFF, trick, tricut, permsep, PBF[pbf],
                           int(maxbf), wrtsize)

    def run_permutations(self, nperm, seed, lo=0, hi=None, trick=0, tricut=10, permsep=0, pbf="none",
                         maxbf=False, wrtsize=10, keep_stats=True) -> PermResults:
        hi = self.ds.n_genes if hi is None else hi
        n = hi - lo
        per = self.S if (self.analysis == "sep" and permsep == 2) else 1
        pv = np.full(n * per, np.nan)
        nd = np.zeros(n * per, dtype=np.int64)
        cnt = np.zeros(n * per, dtype=np.int64)
        ts = np.full(n * per, np.nan)
        med = np.full(n * per, np.nan)
        st = np.full((n * per, nperm), np.nan) if keep_stats else None
        pc = self.perm_config(nperm, seed, trick, tricut, permsep, pbf, maxbf, wrtsize)
        pr = PermResultsC(_ptr(pv), _ptr(nd), _ptr(cnt), _ptr(ts), _ptr(med), _ptr(st))
        self._call("run_permutations", C.c_int64(lo), C.c_int64(hi), C.byref(pc), C.byref(pr))
        shp = (n, per) if per > 1 else (n,)
        return PermResults(pv.reshape(shp), nd.reshape(shp), cnt.reshape(shp), ts.reshape(shp),
                           med.reshape(shp), None if st is None else st.reshape(shp + (nperm,)))
