/*
 * eqtlbma_oracle.h -- CPU ORACLE (test infrastructure only; never linked into the product path).
 *
 * Same flat interface as include/eqtlbma_b200.h (the structs are shared), entry points prefixed
 * eqo_ instead of eqb_.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load liboracle.so.
 */
#ifndef EQTLBMA_ORACLE_H
#define EQTLBMA_ORACLE_H

#include "../include/eqtlbma_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct eqo_ctx eqo_ctx;

int eqo_create(eqo_ctx **ctx, const eqb_config *cfg);
void eqo_destroy(eqo_ctx *ctx);
const char *eqo_last_error(const eqo_ctx *ctx);
int eqo_set_genotypes(eqo_ctx *ctx, int32_t geno_id, const double *G, int64_t n_snps, int32_t n_cols);
int eqo_set_subgroup(eqo_ctx *ctx, int32_t s, const eqb_subgroup *sg);
int eqo_set_grids(eqo_ctx *ctx, const double *phi2L, const double *oma2L, int32_t L,
                  const double *phi2S, const double *oma2S, int32_t K);
int eqo_build_cis_windows(eqo_ctx *ctx, const int32_t *gene_chr, const int64_t *gene_start,
                          const int64_t *gene_end, const int32_t *snp_chr, const int64_t *snp_pos,
                          int32_t anchor, int64_t radius, int64_t *begin_out, int64_t *end_out);
int eqo_set_cis_windows(eqo_ctx *ctx, const int64_t *begin, const int64_t *end);
int eqo_finalize(eqo_ctx *ctx);
int64_t eqo_n_configs(const eqo_ctx *ctx);
int eqo_pair_offsets(eqo_ctx *ctx, int64_t gene_lo, int64_t gene_hi, int64_t *offsets);
int eqo_run(eqo_ctx *ctx, int64_t gene_lo, int64_t gene_hi, eqb_results *res);
int eqo_run_permutations(eqo_ctx *ctx, int64_t gene_lo, int64_t gene_hi, const eqb_perm_config *pc,
                         eqb_perm_results *res);
/* number of OpenMP threads for the permutation loops (the reference's --thread) */
void eqo_set_threads(eqo_ctx *ctx, int32_t n);
/* MT19937 + gsl_ran_shuffle replay used by tests: fills perms[n_perm][n] with the cumulative
 * shuffles of the identity, starting from a freshly seeded generator (gene.cpp:617,639). */
void eqo_shuffle_table(uint64_t seed, int64_t n_skip_shuffles, int64_t n_perm, int32_t n, int32_t *perms);

#ifdef __cplusplus
}
#endif
#endif
