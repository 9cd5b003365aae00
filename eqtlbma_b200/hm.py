"""Host mirror of the reference's eqtlbma_hm Controller (`--model configs`, /root/reference/src/eqtlbma_hm.cpp:50-191)
over the C ABI of include/eqtlbma_hm_b200.h: same sequence (load_data -> init_params -> run_EM ->
compute_posterior -> estimate_profile_ci), every number computed by libeqtlbma_b200.so on the GPU.
No CPU fallback: construction fails when the CUDA library is missing."""
from __future__ import annotations

import ctypes as C

import numpy as np


class _Options(C.Structure):
    _fields_ = [("thresh", C.c_double), ("maxit", C.c_int64), ("stepmax", C.c_double), ("fixed_pi0", C.c_int32),
                ("fixed_grid", C.c_int32), ("fixed_configs", C.c_int32), ("verbose", C.c_int32),
                ("log", C.c_void_p), ("user", C.c_void_p)]


class _Fit(C.Structure):
    _fields_ = [("pi0", C.c_double), ("grid_wts", C.c_void_p), ("config_prior", C.c_void_p), ("loglik", C.c_double),
                ("iters", C.c_int64), ("pi0_ci", C.c_double * 2), ("grid_ci", C.c_void_p), ("config_ci", C.c_void_p)]


_LOGFN = C.CFUNCTYPE(None, C.c_void_p, C.c_char_p)
_GATHERFN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int32)


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class HmFit:
    """Parameters, likelihood and intervals of one fit (arrays owned here, pointed to by the C struct)."""

    def __init__(self, pi0, grid_wts, config_prior):
        self.grid_wts = np.ascontiguousarray(grid_wts, dtype=np.float64).copy()
        self.config_prior = np.ascontiguousarray(config_prior, dtype=np.float64).copy()
        self.grid_ci = np.full((len(self.grid_wts), 2), np.nan)
        self.config_ci = np.full((len(self.config_prior), 2), np.nan)
        self.c = _Fit(float(pi0), _dp(self.grid_wts), _dp(self.config_prior), float("nan"), 0,
                      (C.c_double * 2)(float("nan"), float("nan")), _dp(self.grid_ci), _dp(self.config_ci))

    pi0 = property(lambda self: self.c.pi0)
    loglik = property(lambda self: self.c.loglik)
    iters = property(lambda self: int(self.c.iters))
    pi0_ci = property(lambda self: (self.c.pi0_ci[0], self.c.pi0_ci[1]))


class HmEngine:
    """One eqb_hm context.  `B` [pairs][dim][grid] float64 host array, or `device_ptr` (an integer device address, e.g.
    Engine.raw_abfs_device()) for data already resident on the GPU."""

    def __init__(self, dim: int, grid: int, device: int = 0):
        from . import load_library
        self.lib = load_library()
        self.dim, self.grid = int(dim), int(grid)
        self.ctx = C.c_void_p()
        self.lib.eqb_hm_last_error.restype = C.c_char_p
        self.lib.eqb_hm_n_genes.restype = C.c_int64
        self.lib.eqb_hm_n_pairs.restype = C.c_int64
        self.lib.eqb_hm_launch_count.restype = C.c_int64
        rc = self.lib.eqb_hm_create(C.byref(self.ctx), C.c_int32(device), C.c_int32(dim), C.c_int32(grid))
        if rc != 0:
            msg = self.lib.eqb_hm_last_error(self.ctx) if self.ctx else b"eqb_hm_create failed"
            raise RuntimeError(msg.decode() if isinstance(msg, bytes) else str(msg))

    def close(self):
        if self.ctx:
            self.lib.eqb_hm_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"eqb_hm_{what}: " + self.lib.eqb_hm_last_error(self.ctx).decode())

    def append(self, B: np.ndarray, gene_off: np.ndarray):
        B = np.ascontiguousarray(B, dtype=np.float64)
        assert B.ndim == 3 and B.shape[1] == self.dim and B.shape[2] == self.grid
        off = np.ascontiguousarray(gene_off, dtype=np.int64)
        self._check(self.lib.eqb_hm_append(self.ctx, _dp(B), C.c_int64(B.shape[0]), _dp(off), C.c_int64(len(off) - 1)), "append")

    def append_device(self, device_ptr: int, n_pairs: int, gene_off: np.ndarray):
        off = np.ascontiguousarray(gene_off, dtype=np.int64)
        self._check(self.lib.eqb_hm_append_device(self.ctx, C.c_void_p(device_ptr), C.c_int64(n_pairs), _dp(off),
                                                  C.c_int64(len(off) - 1)), "append_device")

    def finalize(self):
        self._check(self.lib.eqb_hm_finalize(self.ctx), "finalize")

    def set_collective(self, group=None, native=None):
        """Genes sharded over the ranks of a torch.distributed group.  native (default on an NCCL group): the partial sums
        of every evaluation are exchanged by hm_xchg_kernel over peer memory (eqb_hm_ipc_export / eqb_hm_ipc_connect; the
        group only carries the 64-byte IPC handles once).  Otherwise they are all-gathered through the group (NCCL on the
        context's device, or gloo on the host) and combined on the host in rank order."""
        import torch
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if native is None:
            native = dist.get_backend(group) == "nccl"
        if native:
            mine = C.create_string_buffer(64)
            self._check(self.lib.eqb_hm_ipc_export(self.ctx, mine), "ipc_export")
            handles = [None] * world
            dist.all_gather_object(handles, bytes(mine.raw), group=group)
            blob = C.create_string_buffer(b"".join(handles), 64 * world)
            self._check(self.lib.eqb_hm_ipc_connect(self.ctx, C.c_int32(world), C.c_int32(rank), blob), "ipc_connect")
            return
        on_gpu = dist.get_backend(group) == "nccl"
        dev = torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu")

        def gather(user, send, recv, n):
            try:
                mine = torch.from_numpy(np.ctypeslib.as_array(send, shape=(n,)).copy()).to(dev)
                parts = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(world)]
                dist.all_gather(parts, mine, group=group)
                np.ctypeslib.as_array(recv, shape=(world * n,))[:] = torch.stack(parts).cpu().numpy().ravel()
                return 0
            except Exception as exc:  # the C side reports the failure
                print("eqb_hm all-gather failed:", exc)
                return 1

        self._gather_cb = _GATHERFN(gather)
        self._check(self.lib.eqb_hm_set_collective(self.ctx, C.c_int32(world), C.c_int32(rank), self._gather_cb, None), "set_collective")

    n_genes = property(lambda self: int(self.lib.eqb_hm_n_genes(self.ctx)))
    n_pairs = property(lambda self: int(self.lib.eqb_hm_n_pairs(self.ctx)))
    launch_count = property(lambda self: int(self.lib.eqb_hm_launch_count(self.ctx)))

    def loglik(self, pi0, grid_wts, config_prior, keep=False) -> float:
        gw = np.ascontiguousarray(grid_wts, dtype=np.float64)
        cp = np.ascontiguousarray(config_prior, dtype=np.float64)
        out = C.c_double(0)
        self._check(self.lib.eqb_hm_loglik(self.ctx, C.c_double(pi0), _dp(gw), _dp(cp), C.c_int32(int(keep)), C.byref(out)), "loglik")
        return out.value

    def esums(self, pi0, grid_wts, config_prior) -> np.ndarray:
        gw = np.ascontiguousarray(grid_wts, dtype=np.float64)
        cp = np.ascontiguousarray(config_prior, dtype=np.float64)
        out = np.zeros(1 + self.dim + self.grid)
        self._check(self.lib.eqb_hm_esums(self.ctx, C.c_double(pi0), _dp(gw), _dp(cp), _dp(out)), "esums")
        return out

    def em(self, fit: HmFit, thresh=0.05, maxit=None, stepmax=1.0, fixed=None, log=None) -> HmFit:
        fixed = fixed or {}
        lines = []
        cb = _LOGFN(lambda user, text: (log or lines.append)(text.decode()))
        opt = _Options(float(thresh), -1 if maxit is None else int(maxit), float(stepmax), int(bool(fixed.get("pi0"))),
                       int(bool(fixed.get("grid"))), int(bool(fixed.get("configs"))), 1, C.cast(cb, C.c_void_p), None)
        self._check(self.lib.eqb_hm_em(self.ctx, C.byref(opt), C.byref(fit.c)), "em")
        fit.log_lines = lines
        return fit

    def profile_ci(self, fit: HmFit) -> HmFit:
        self._check(self.lib.eqb_hm_profile_ci(self.ctx, C.byref(fit.c)), "profile_ci")
        return fit

    def posteriors(self, fit: HmFit) -> dict:
        G, P, D = self.n_genes, self.n_pairs, self.dim
        out = dict(gene_post=np.zeros(G), gene_bf=np.zeros(G), snp_bf=np.zeros(P), snp_post=np.zeros(P),
                   cfg_bf=np.zeros((P, D)), gene_cfg_post=np.zeros((G, D)))
        self._check(self.lib.eqb_hm_posteriors(self.ctx, C.byref(fit.c), _dp(out["gene_post"]), _dp(out["gene_bf"]),
                                               _dp(out["snp_bf"]), _dp(out["snp_post"]), _dp(out["cfg_bf"]),
                                               _dp(out["gene_cfg_post"])), "posteriors")
        return out

    def estep_device_only(self, grid_wts, config_prior, reps=10) -> float:
        gw = np.ascontiguousarray(grid_wts, dtype=np.float64)
        cp = np.ascontiguousarray(config_prior, dtype=np.float64)
        ms = C.c_float(0)
        self._check(self.lib.eqb_hm_estep_device_only(self.ctx, _dp(gw), _dp(cp), C.c_int32(reps), C.byref(ms)), "estep_device_only")
        return float(ms.value)
