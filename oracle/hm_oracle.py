"""TEST INFRASTRUCTURE -- never imported by the product path (eqtlbma_b200/).

numpy restatement of the reference's hierarchical-model EM for `--model configs`
(/root/reference/src/eqtlbma_hm.cpp:617-1114 + src/hm_methods.cpp), vectorised over genes but with the
reference's order of averaging: grid points, then configurations, then SNPs, each through
log10_weighted_sum (src/eqtlbma/utils/utils_math.cpp:135-159).  Classical EM and the posteriors only;
SQUAREM and the profile-likelihood intervals are pinned directly against the dumps of the compiled
reference (tests/golden/hm/*.dump.gz, oracle/make_golden_hm.py).

Parity pinned: tests/test_hm_oracle.py checks this file against those dumps."""
from __future__ import annotations

import numpy as np

DBL_EPSILON = np.finfo(np.float64).eps


def l10ws(vec: np.ndarray, wts: np.ndarray, axis: int = -1) -> np.ndarray:
    """log10_weighted_sum (utils_math.cpp:135-159) along `axis`; wts broadcastable to vec."""
    mx = np.max(vec, axis=axis, keepdims=True)
    s = np.sum(wts * np.power(10.0, vec - mx), axis=axis, keepdims=True)
    res = np.squeeze(mx + np.log10(s), axis=axis)
    return np.where(np.abs(res) <= DBL_EPSILON, 0.0, res)


class HmOracle:
    def __init__(self, B: np.ndarray, gene_off: np.ndarray):
        self.B = np.asarray(B, dtype=np.float64)  # [pairs][dim][grid]
        self.gene_off = np.asarray(gene_off, dtype=np.int64)
        self.G = len(gene_off) - 1
        self.P, self.dim, self.grid = self.B.shape
        self.m = np.diff(self.gene_off)
        self.gene_of = np.repeat(np.arange(self.G), self.m)

    def _over_snps(self, v: np.ndarray) -> np.ndarray:
        """log10_weighted_sum over the SNPs of each gene with weights 1/m_g (gene_eQTL::compute_log10_BF,
        hm_methods.cpp:395-432); v: [pairs, ...] -> [genes, ...]."""
        mx = np.full((self.G,) + v.shape[1:], -np.inf)
        np.maximum.at(mx, self.gene_of, v)
        w = (1.0 / self.m)[self.gene_of].reshape((-1,) + (1,) * (v.ndim - 1))
        s = np.zeros_like(mx)
        np.add.at(s, self.gene_of, w * np.power(10.0, v - mx[self.gene_of]))
        res = mx + np.log10(s)
        return np.where(np.abs(res) <= DBL_EPSILON, 0.0, res)

    # snp_eQTL::compute_log10_config_BF (hm_methods.cpp:343-349)
    def cfg_bf(self, gw):
        return l10ws(self.B, gw[None, None, :], axis=2)  # [pairs][dim]

    # snp_eQTL::compute_log10_BF (hm_methods.cpp:74-126)
    def snp_bf(self, gw, cp):
        return l10ws(self.cfg_bf(gw), cp[None, :], axis=1)

    def gene_bf(self, gw, cp):
        return self._over_snps(self.snp_bf(gw, cp))

    # gene_eQTL::compute_log10_obs_lik (hm_methods.cpp:479-500), Controller::compute_log10_obs_lik (eqtlbma_hm.cpp:617-650)
    def gene_lik(self, pi0, gw, cp):
        gb = self.gene_bf(gw, cp)
        vec = np.stack([np.zeros_like(gb), gb], axis=1)
        return l10ws(vec, np.array([pi0, 1.0 - pi0])[None, :], axis=1)

    def loglik(self, pi0, gw, cp):
        return float(np.sum(self.gene_lik(pi0, gw, cp)))

    # Controller::run_EM_fixedpoint (eqtlbma_hm.cpp:924-1009) with the per-gene likelihoods kept at (pi0, gw, cp)
    def fixedpoint(self, pi0, gw, cp, fixed):
        lik = self.gene_lik(pi0, gw, cp)
        ones = np.ones(self.G)
        new_pi0 = pi0 if fixed["pi0"] else float(np.sum(np.power(10.0, np.log10(pi0) - lik)) / self.G)
        new_cp = cp.copy()
        if self.dim > 1 and not fixed["configs"]:
            cg = self._over_snps(self.cfg_bf(gw)) - lik[:, None]  # gene_eQTL::em_update_config
            t = l10ws(cg, ones[:, None], axis=0) + np.log10(cp)
            new_cp = np.power(10.0, t - l10ws(t, np.ones(self.dim)))
        new_gw = gw.copy()
        if not fixed["grid"]:
            per = l10ws(self.B, cp[None, :, None], axis=1)  # snp_eQTL::em_update_grid: over configs at each grid point
            gg = self._over_snps(per) - lik[:, None]
            t = l10ws(gg, ones[:, None], axis=0) + np.log10(gw)
            new_gw = np.power(10.0, t - l10ws(t, np.ones(self.grid)))
        return new_pi0, new_gw, new_cp

    # Controller::run_EM_classic (eqtlbma_hm.cpp:1076-1108)
    def run_em_classic(self, pi0, gw, cp, fixed, thresh=0.05, maxit=None):
        gw, cp = gw.copy(), cp.copy()
        lik = self.loglik(pi0, gw, cp)
        it = 0
        while True:
            it += 1
            n_pi0, n_gw, n_cp = self.fixedpoint(pi0, gw, cp, fixed)
            n_lik = self.loglik(n_pi0, n_gw, n_cp)
            if n_lik < lik:
                raise RuntimeError("observed log-likelihood is decreasing")
            if abs(n_lik - lik) < thresh or (maxit is not None and it == maxit - 1):
                break
            pi0, gw, cp, lik = n_pi0, n_gw, n_cp, n_lik
        pi0, gw, cp, lik = n_pi0, n_gw, n_cp, n_lik
        it += 1
        n_pi0, n_gw, n_cp = self.fixedpoint(pi0, gw, cp, fixed)
        lik = self.loglik(n_pi0, n_gw, n_cp)
        return n_pi0, n_gw, n_cp, lik, it

    # gene_eQTL::compute_posterior (hm_methods.cpp:752-781) + the BF columns of save_result (eqtlbma_hm.cpp:1713-1739)
    def posteriors(self, pi0, gw, cp):
        cb = self.cfg_bf(gw)
        sb = l10ws(cb, cp[None, :], axis=1)
        gb = self._over_snps(sb)
        lik = l10ws(np.stack([np.zeros_like(gb), gb], axis=1), np.array([pi0, 1.0 - pi0])[None, :], axis=1)
        gene_post = np.minimum(1.0, np.power(10.0, np.log10(1.0 - pi0) + gb - lik))
        prior = (1.0 / self.m)[self.gene_of]
        snp_post = np.minimum(1.0, np.power(10.0, np.log10(1.0 - pi0) + np.log10(prior) + sb - lik[self.gene_of]))
        contrib = (prior * (1.0 - pi0))[:, None] * cp[None, :] * np.power(10.0, cb - lik[self.gene_of][:, None])
        gcp = np.zeros((self.G, self.dim))
        np.add.at(gcp, self.gene_of, contrib)
        return dict(gene_post=gene_post, gene_bf=gb, snp_bf=sb, snp_post=snp_post, cfg_bf=cb,
                    gene_cfg_post=np.minimum(1.0, gcp))
