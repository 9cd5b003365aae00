/*
 * eqtlbma_b200.h -- C ABI of the B200-native eqtlbma_bf hot path (libeqtlbma_b200.so).
 *
 * The reference (timflutre/eqtlbma v1.3.3) has no plugin/FFI layer; the seam this ABI
 * replaces is the body of the write-group loop of run() in src/eqtlbma_bf.cpp:1546-1574:
 *     testForAssociations(...)   eqtlbma_bf.cpp:707-771   -> eqb_run()
 *     makePermutations(...)      eqtlbma_bf.cpp:866-917   -> eqb_run_permutations()
 * whose results the writers (writeRes*, eqtlbma_bf.cpp:919-1447) read only through the
 * GeneSnpPair / Gene getters (gene_snp_pair.hpp:144,251-262; gene.hpp:153-190).  Every entry
 * point below cites the reference interface it stands for.  Plain pointers and sizes only.
 *
 * Index spaces (all produced by the host loader):
 *   subgroups  0..S-1       sorted subgroup ids                (data_loader.cpp:141-143)
 *   samples    0..N_all-1   sorted union of sample names       (samples.cpp:60-68)
 *   snps       0..M-1       chromosomes in name order, position order inside (data_loader.cpp:827-837)
 *   genes      0..G-1       byte-wise gene-name order = std::map<string,Gene> order (eqtlbma_bf.cpp:1546)
 *   configs    0..C-1       gsl_combination lexicographic order by size k=1..S (gene_snp_pair.cpp:504-550);
 *                           C = S for --bfs sin, 2^S-1 for --bfs all, 0 for --bfs gen
 *   pairs                   for gene g in order, for snp in [cis_begin[g], cis_end[g]) in order;
 *                           genes that are not analyzed contribute no pair
 *
 * All functions return 0 on success, non-zero on error (message via eqb_last_error()).
 * A context is re-entrant per ctx, not thread-safe within a ctx.  There is NO CPU fallback:
 * every compute entry point fails if no CUDA device is usable.
 */
#ifndef EQTLBMA_B200_H
#define EQTLBMA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EQB_ABI_VERSION 1

/* --analys (eqtlbma_bf.cpp:1595) */
enum { EQB_ANALYSIS_SEP = 0, EQB_ANALYSIS_JOIN = 1 };
/* --bfs (eqtlbma_bf.cpp:1596, gene_snp_pair.cpp:604-622) */
enum { EQB_BFS_GEN = 0, EQB_BFS_SIN = 1, EQB_BFS_ALL = 2 };
/* --pbf (eqtlbma_bf.cpp:652-657) */
enum { EQB_PBF_NONE = 0, EQB_PBF_GEN = 1, EQB_PBF_GEN_SIN = 2, EQB_PBF_ALL = 3 };
/* --error (eqtlbma_bf.cpp:1596).  MVLR: gene_snp_pair.cpp:624-758 + MVLR.cpp (same individuals in every subgroup,
 * at most 16 subgroups).  HYBRID: gene_snp_pair.cpp:760-1423 (individuals common to / unique to each pair of subgroups;
 * at most 16 subgroups and 6 covariates; with covariates their files must be in the order of the sorted sample names,
 * because the reference indexes them by the all-sample index, :931-933 -- checked by eqb_finalize). */
enum { EQB_ERROR_UVLR = 0, EQB_ERROR_MVLR = 1, EQB_ERROR_HYBRID = 2 };
/* --anchor (snp.cpp:274-297) */
enum { EQB_ANCHOR_TSS = 0, EQB_ANCHOR_TSS_TES = 1 };

typedef struct eqb_ctx eqb_ctx;

/* Options that shape the device layouts (subset of eqtlbma_bf.cpp:1586-1597). */
typedef struct {
  int32_t abi_version; /* EQB_ABI_VERSION */
  int32_t n_subgroups; /* S */
  int32_t n_samples_all; /* N_all = size of the sorted sample union (samples.cpp:60-68) */
  int32_t analysis; /* EQB_ANALYSIS_* */
  int64_t n_snps; /* M */
  int64_t n_genes; /* G */
  int32_t bfs; /* EQB_BFS_* */
  int32_t error_model; /* EQB_ERROR_* */
  int32_t qnorm; /* --qnorm (utils_math.cpp:80-96) */
  int32_t device; /* CUDA device ordinal */
  double fiterr; /* --fiterr, MVLR only (MVLR.cpp:148-173) */
} eqb_config;

/* One subgroup's data, the flat equivalent of Samples / Gene / Snp / Covariates for that
 * subgroup (samples.hpp:34-70, gene.hpp:56, snp.hpp:39-87, covariates.hpp:34-50). */
typedef struct {
  int32_t geno_id; /* which matrix of eqb_set_genotypes() holds this subgroup's dosages */
  int32_t n_exp_cols; /* columns of Y */
  int32_t n_covariates; /* Q, rows of C (name-sorted: covariates.cpp:66-75) */
  int32_t n_cov_cols; /* columns of C */
  const int32_t *all2geno; /* [N_all] column of G for each sample, -1 absent (samples.cpp:139) */
  const int32_t *all2exp; /* [N_all] column of Y, -1 absent (samples.cpp:133) */
  const int32_t *all2cov; /* [N_all] column of C, -1 absent (samples.cpp:145); may be NULL if Q=0 */
  const uint8_t *snp_has_geno; /* [M] Snp::HasGenotypes(subgroup) (snp.cpp:299); NULL = all */
  const uint8_t *gene_has_exp; /* [G] Gene::HasExplevels(subgroup) (gene.cpp:185); NULL = all */
  const double *Y; /* [G][n_exp_cols] expression levels, NaN = missing (gene.cpp:106-124) */
  const double *C; /* [Q][n_cov_cols] covariates */
} eqb_subgroup;

/* Results of the non-permuted pass; caller-allocated HOST buffers (pinned memory makes the
 * copies asynchronous), any pointer may be NULL to skip that output.  n_pairs comes from
 * eqb_pair_offsets().  Layouts are pair-major. */
typedef struct {
  int32_t *n; /* [pairs][S]     GetSampleSize, 0 = no result (gene_snp_pair.cpp:61-68) */
  double *sstats; /* [pairs][S][5]  pve, sigmahat, betahat.geno, sebetahat.geno, betapval.geno
                     (GetPve..GetBetapvalGeno, gene_snp_pair.cpp:1430-1453); NaN when n = 0 */
  double *abf_gen; /* [pairs][3][L]  unweighted "gen","gen-fix","gen-maxh" (gene_snp_pair.cpp:364-416) */
  double *abf_cfg; /* [pairs][C][K]  unweighted per-configuration ABFs (gene_snp_pair.cpp:422-550) */
  double *abf_w; /* [pairs][5+C]   weighted: gen, gen-fix, gen-maxh, gen-sin, all, configs...
                    (GetWeightedAbf, gene_snp_pair.cpp:1467); NaN where --bfs does not produce it */
  uint8_t *gene_analyzed; /* [gene_hi-gene_lo] 1 if the gene was analyzed (eqtlbma_bf.cpp:747-768) */
} eqb_results;

/* Permutation options (eqtlbma_bf.cpp:1586-1597, 773-917). */
typedef struct {
  int64_t nperm; /* --nperm */
  uint64_t seed; /* --seed, passed to gsl_rng_set (MT19937) once per write-group */
  int32_t trick; /* --trick 0|1|2 */
  int32_t tricut; /* --tricut */
  int32_t permsep; /* --permsep 0|1|2 (sep analysis) */
  int32_t pbf; /* EQB_PBF_* (join analysis) */
  int32_t maxbf; /* --maxbf */
  int32_t wrtsize; /* --wrtsize: the RNG is re-seeded every wrtsize genes (eqtlbma_bf.cpp:847) */
} eqb_perm_config;

/* Per-gene permutation results; caller-allocated host buffers of gene_hi-gene_lo entries
 * (x S for --permsep 2, subgroup-minor).  Genes that are not analyzed get NaN / 0. */
typedef struct {
  double *pval; /* GetPermutationPvalueJoin / Sep (gene.hpp:176,183) */
  int64_t *nperm_done; /* GetNbPermutationsJoin / Sep */
  int64_t *count; /* 1 + #{perm stat >= true stat} (resp. <=) BEFORE division: the bit-exact quantity */
  double *true_stat; /* GetTrueL10Abf / GetTrueMinPval */
  double *median_perm; /* GetMedianPermL10Abf: median of the permuted statistics (join only; the
                          reference reads one element past its vector, gene.cpp:713 -- documented tie) */
  double *perm_stats; /* optional [genes][nperm] (x S for permsep 2): every permuted statistic, NaN if not evaluated */
} eqb_perm_results;

/* Number of usable CUDA devices (0 if none); lets a launcher place one shard process per GPU. */
int eqb_device_count(void);
/* Optional: create the device's primary CUDA context ahead of eqb_create (driver start-up, 0.5 - 1.5 s per process), e.g. on
 * a thread while the host parses its input files (loadRawInputData, eqtlbma_bf.cpp:1496-1501, has no device work). */
int eqb_warmup(int32_t device);
int eqb_create(eqb_ctx **ctx, const eqb_config *cfg);
void eqb_destroy(eqb_ctx *ctx);
const char *eqb_last_error(const eqb_ctx *ctx);

/* Snp::subgroup2genotypes_ (snp.hpp:45): dosages of one genotype file, SNP-major
 * G[snp * n_cols + col] (= column-major N x M); rows of SNPs the file lacks may hold anything.
 * The upload is asynchronous (row chunks on a copy stream, overlapped with the projection of the
 * chunks that already landed and with the results of eqb_run travelling back): G must stay valid
 * and unchanged until the first eqb_run* call on this context has returned, or eqb_destroy();
 * pinned host memory (cudaHostAlloc / cudaHostRegister) is what makes the overlap effective. */
int eqb_set_genotypes(eqb_ctx *ctx, int32_t geno_id, const double *G, int64_t n_snps, int32_t n_cols);
/* The same matrix in the compact, LOSSLESS transport format a text parser can produce for free: unsigned
 * integers k of elem_bytes (1 or 2) bytes, element value = (double)k / denom.  Hard calls (VCF GT -> 0/1/2,
 * snp.cpp:130-185) use denom 1; a dosage / IMPUTE file written with d decimals (snp.cpp:88-128,
 * data_loader.cpp:570-1010) uses denom 10^d: the IEEE quotient of two exactly representable integers is the
 * correctly rounded value of k/10^d, i.e. the very double strtod() yields for the decimal text, so the device
 * sees bit-identical dosages while the host->device transfer (the end-to-end bound of a no-permutation run)
 * shrinks 8x / 4x.  Same lifetime and asynchrony rules as eqb_set_genotypes(). */
int eqb_set_genotypes_fixed(eqb_ctx *ctx, int32_t geno_id, const void *G, int32_t elem_bytes, double denom,
                            int64_t n_snps, int32_t n_cols);
/* The expression matrix sg->Y is uploaded asynchronously too: it must stay valid and unchanged until
 * eqb_finalize() has returned (the small arrays of the struct are copied before eqb_set_subgroup returns). */
int eqb_set_subgroup(eqb_ctx *ctx, int32_t s, const eqb_subgroup *sg);
/* Grid (grid.cpp:28-65): phi2/oma2 columns of --gridL (L points) and --gridS (K points). */
int eqb_set_grids(eqb_ctx *ctx, const double *phi2L, const double *oma2L, int32_t L,
                  const double *phi2S, const double *oma2S, int32_t K);
/* Gene::SetCisSnps + Snp::IsInCis (gene.cpp:140-157, snp.cpp:274-297) evaluated on the device:
 * snp_chr/gene_chr are chromosome indexes, snp_pos 1-based positions sorted inside each
 * chromosome, gene_start 1-based (BED start + 1, gene.cpp:44).  Writes [begin,end) per gene into
 * the optional host arrays and keeps them in the context. */
int eqb_build_cis_windows(eqb_ctx *ctx, const int32_t *gene_chr, const int64_t *gene_start,
                          const int64_t *gene_end, const int32_t *snp_chr, const int64_t *snp_pos,
                          int32_t anchor, int64_t radius, int64_t *begin_out, int64_t *end_out);
/* Alternative: windows computed by the caller. */
int eqb_set_cis_windows(eqb_ctx *ctx, const int64_t *begin, const int64_t *end);
/* Builds the device-resident all-sample-space layouts; call once after the setters. */
int eqb_finalize(eqb_ctx *ctx);

int64_t eqb_n_configs(const eqb_ctx *ctx);
/* offsets[i] = number of pairs of analyzed genes in [gene_lo, gene_lo+i); length gene_hi-gene_lo+1 */
int eqb_pair_offsets(eqb_ctx *ctx, int64_t gene_lo, int64_t gene_hi, int64_t *offsets);

/* testForAssociations (eqtlbma_bf.cpp:707-771) for genes [gene_lo, gene_hi). */
int eqb_run(eqb_ctx *ctx, int64_t gene_lo, int64_t gene_hi, eqb_results *res);
/* makePermutations (eqtlbma_bf.cpp:866-917) for genes [gene_lo, gene_hi); gene_lo must be a
 * multiple of wrtsize (write-groups are formed from gene 0).  eqb_run() need not be called first. */
int eqb_run_permutations(eqb_ctx *ctx, int64_t gene_lo, int64_t gene_hi, const eqb_perm_config *pc,
                         eqb_perm_results *res);

/* Device-resident variant used for throughput measurement: runs the same kernels as eqb_run /
 * eqb_run_permutations but leaves results in device memory (no D2H), returning the device time
 * in milliseconds measured with CUDA events on the library's stream. */
int eqb_run_device_only(eqb_ctx *ctx, int64_t gene_lo, int64_t gene_hi, int32_t want_raw, float *ms);
int eqb_run_permutations_device_only(eqb_ctx *ctx, int64_t gene_lo, int64_t gene_hi,
                                     const eqb_perm_config *pc, float *ms);
/* --inss: Bayes factors from summary statistics instead of raw data -- what loadSummaryStats +
 * fillGeneSnpPairsWithSstats (data_loader.cpp:1251-1343, GeneSnpPair::SetSstats gene_snp_pair.cpp:241-254) feed into
 * TestForAssociations(hasDataNotSstats = false) (gene.cpp:293-311): standardisation with nu = n - 2
 * (gene_snp_pair.cpp:256-290; the covariate count is unknown on this path) and CalcAbfsUvlr.  Inputs are pair-major
 * [pairs][S]; n <= 0 marks a subgroup without an entry for the pair.  Needs eqb_create (join analysis, uvlr) and
 * eqb_set_grids only; fills res->abf_gen / abf_cfg / abf_w (NULL pointers are skipped). */
int eqb_bf_from_sstats(eqb_ctx *ctx, int64_t n_pairs, const int32_t *n, const double *sigmahat, const double *betahat,
                       const double *sebetahat, eqb_results *res);
/* Multi-GPU sharding (replaces scripts/eqtlbma_bf_parallel.bash:248-262, one OS process per gene batch):
 * genes are independent, so the G genes are cut into n_shards CONTIGUOUS ranges of whole write-groups
 * (the generator is re-seeded per write-group, eqtlbma_bf.cpp:847, so a group never straddles two
 * GPUs and results do not depend on the GPU count), balanced on the caller's per-gene cost (e.g.
 * cis SNPs x (1 + nperm)).  Pure host function, no context, no device: shard k owns genes
 * [shard_begin[k], shard_begin[k+1]); concatenating the shards' outputs in shard order restores the
 * reference's gene order (the final host gather). */
int eqb_partition_by_cost(const int64_t *cost_per_gene, int64_t n_genes, int64_t wrtsize, int32_t n_shards,
                          int64_t *shard_begin /* n_shards + 1 */);
/* Number of genes whose (gene, subgroup) row sets are gene-independent (K1 outputs reusable: the
 * split projection / contraction path); the others take the general fused kernel. Diagnostic. */
int64_t eqb_fast_gene_count(const eqb_ctx *ctx);
/* Duration (ms, CUDA events on the library's stream) of the last K2+K3 launch (fast_pair_kernel) issued by
 * eqb_run_device_only: the dominant kernel of the non-permuted pass, for roofline accounting. */
float eqb_last_pair_kernel_ms(const eqb_ctx *ctx);
/* Number of kernel launches issued by this context so far. */
int64_t eqb_launch_count(const eqb_ctx *ctx);
/* Raw per-configuration log10 ABFs of the LAST chunk computed by eqb_run (with results->abf_cfg) or
 * eqb_run_device_only(want_raw = 1), still resident in device memory as [pair][config][small-grid point] doubles -- the
 * layout eqb_hm_append_device (include/eqtlbma_hm_b200.h) takes, so that the hierarchical model can be fitted without the
 * `_l10abfs_raw.txt.gz` round trip (eqtlbma_bf.cpp:1083-1229 writes it, eqtlbma_hm.cpp:287-371 parses it back).  The
 * pointer stays valid until the next eqb_run* call on this context.  A run that fits the device budget is one chunk.
 * eqb_raw_abfs_layout: ids of the genes with at least one pair (or NULL) and their pair offsets [n_genes + 1]. */
int eqb_raw_abfs_device(eqb_ctx *ctx, const double **d_B, int64_t *n_pairs, int64_t *n_genes);
int eqb_raw_abfs_layout(eqb_ctx *ctx, int64_t *gene_ids, int64_t *gene_off);
/* diagnostics: worst deviation of the table-driven rcp / log / rsqrt / exp of the permutation BF kernel (perm_gemm.cuh)
 * from the CUDA library versions over n pseudo-random arguments:
 * out5 = { rcp rel, log abs, rsqrt rel, exp rel, special-value mismatches }. */
int eqb_math_selftest(int32_t device, int64_t n, double *out5);

/* FP64 pipe peaks of the device measured with register-resident loops (the denominators of the FP64 rooflines,
 * BASELINE.md section 1): out4 = { DFMA TFLOP/s, DMMA (mma.sync.m8n8k4.f64) TFLOP/s, implied SM clock in MHz, SMs }. */
int eqb_measure_fp64_peaks(int32_t device, double *out4);
/* Self-test + throughput of the TMA / DMMA product kernel of the permutation path on pseudo-random operands
 * (n_rows genotype rows x n_cols operand rows of ldn doubles, ldn a multiple of 16):
 * out3 = { worst |D - reference| / sum |terms|, TFLOP/s (CUDA events, best of 3), tiles }. */
int eqb_selftest_perm_gemm(int32_t device, int64_t n_rows, int64_t n_cols, int32_t ldn, double *out3);
/* Per-kernel device time of the permutation path (Gene::MakePermutations*, gene.cpp:380-717), for roofline accounting:
 * enable before eqb_run_permutations*(); out8 = { prep ms, GEMM ms, BF ms, merge ms, GEMM flop issued, GEMM flop
 * useful (rows x columns that belong to a (SNP, subgroup, permutation)), (SNP, permutation) items, path }. */
int eqb_set_perm_timing(eqb_ctx *ctx, int32_t on);
int eqb_last_perm_timing(const eqb_ctx *ctx, double *out8);

#ifdef __cplusplus
}
#endif
#endif
