/*
 * eqtlbma_hm_b200.h -- C ABI of the B200-native EM of the hierarchical model (eqtlbma_hm, `--model configs`),
 * part of libeqtlbma_b200.so.
 *
 * the EM of src/eqtlbma_hm.cpp:652-1345 + src/hm_methods.cpp consuming the raw
 * per-configuration log10 Bayes factors of eqtlbma_bf -- from host arrays (what a loader of `_l10abfs_raw.txt.gz`
 * produces) or straight from the device buffer of eqb_run() (eqb_raw_abfs_device(), no text round trip).
 * The reference has no plugin layer; the seam is its Controller (eqtlbma_hm.cpp:50-191): load_data -> eqb_hm_append*,
 * compute_log10_obs_lik -> eqb_hm_loglik, em_update_{pi0,config,grid} -> eqb_hm_esums, run_EM -> eqb_hm_em,
 * estimate_profile_ci -> eqb_hm_profile_ci, compute_posterior + the BF columns of save_result -> eqb_hm_posteriors.
 *
 * Data layout: B[pair][config][grid point] doubles (the layout of eqb_results.abf_cfg), pairs of a gene contiguous,
 * gene g owning pairs [gene_off[g], gene_off[g+1]).  SNP priors are uniform within a gene (gene_eQTL::set_snp_prior,
 * hm_methods.cpp:385-389; the reference never applies its SNP-prior update, eqtlbma_hm.cpp:665,1026).
 * All functions return 0 on success, non-zero on error (message via eqb_hm_last_error()).  No CPU fallback.
 */
#ifndef EQTLBMA_HM_B200_H
#define EQTLBMA_HM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct eqb_hm_ctx eqb_hm_ctx;

/* Controller::Controller (eqtlbma_hm.cpp:197-285): dim = number of configurations kept (--dim), grid = --ngrid.
 * Limits: 1 <= dim <= 4096, 1 <= grid <= 32. */
int eqb_hm_create(eqb_hm_ctx **hm, int32_t device, int32_t dim, int32_t grid);
void eqb_hm_destroy(eqb_hm_ctx *hm);
const char *eqb_hm_last_error(const eqb_hm_ctx *hm);

/* Controller::load_data_one_file (eqtlbma_hm.cpp:287-371) after parsing: append n_genes genes with their pairs.
 * B: HOST array [n_pairs][dim][grid]; gene_off: [n_genes + 1], gene_off[0] = 0, gene_off[n_genes] = n_pairs, every
 * gene with at least one pair.  May be called once per input file. */
int eqb_hm_append(eqb_hm_ctx *hm, const double *B, int64_t n_pairs, const int64_t *gene_off, int64_t n_genes);
/* Same with B in DEVICE memory of the context's device (e.g. the pointer of eqb_raw_abfs_device()): device-to-device copy. */
int eqb_hm_append_device(eqb_hm_ctx *hm, const double *d_B, int64_t n_pairs, const int64_t *gene_off, int64_t n_genes);
/* End of loading: work units are formed; fails if a value is NaN or infinite (the reference's likelihood is NaN
 * there and it aborts, eqtlbma_hm.cpp:640-647). */
int eqb_hm_finalize(eqb_hm_ctx *hm);
int64_t eqb_hm_n_genes(const eqb_hm_ctx *hm);
int64_t eqb_hm_n_pairs(const eqb_hm_ctx *hm);

/* Multi-GPU: the genes are sharded over `world` processes (one per GPU), every rank holding whole genes.  The one exchange
 * step of the EM is the sum over genes (eqtlbma_hm.cpp:625-633, 659-868: the OpenMP reductions of the reference): each
 * likelihood / E-step evaluation all-gathers 2 + dim + grid doubles per rank through `fn` (recv = [world][n] in rank order;
 * e.g. torch.distributed.all_gather over NCCL) and combines them in rank order, so every rank takes identical decisions.
 * Call after eqb_hm_finalize on every rank; eqb_hm_posteriors stays local (the genes of the rank).
 * eqb_hm_combine_partials is the host-side combination itself (entries 0-1 sums, the others log10 of sums). */
typedef int (*eqb_hm_allgather_fn)(void *user, const double *send, double *recv, int32_t n);
int eqb_hm_set_collective(eqb_hm_ctx *hm, int32_t world, int32_t rank, eqb_hm_allgather_fn fn, void *user);
int eqb_hm_combine_partials(const double *gathered, int32_t world, int32_t n, double *out);
/* The same exchange WITHOUT the host: every rank exports an exchange buffer (CUDA IPC handle, 64 bytes), the caller gathers
 * the handles of all ranks (any transport: torch.distributed.all_gather_object, MPI, a file) and connects.  From then on
 * each evaluation ends with one kernel (hm_xchg_kernel) that stores the rank's partial sums into every peer's buffer over
 * NVLink / NVSwitch, publishes an epoch flag, waits for the peers' flags (bounded) and combines in rank order -- one
 * launch instead of device -> host -> NCCL -> host.  One process per GPU on one node, world <= 16.  Call both after
 * eqb_hm_finalize; every rank must have exported before any rank connects. */
int eqb_hm_ipc_export(eqb_hm_ctx *hm, void *handle64);
int eqb_hm_ipc_connect(eqb_hm_ctx *hm, int32_t world, int32_t rank, const void *handles /* [world][64] */);

/* Controller::compute_log10_obs_lik (eqtlbma_hm.cpp:617-650): sum over genes of log10(pi0 + (1 - pi0) BF_g) with
 * BF_g the average over SNPs, configurations (config_prior[dim]) and grid points (grid_wts[grid]).  keep != 0 stores the
 * per-gene values the E-step and the posteriors read (the reference's `keep`). */
int eqb_hm_loglik(eqb_hm_ctx *hm, double pi0, const double *grid_wts, const double *config_prior, int32_t keep,
                  double *loglik);
/* E-step sums of Controller::em_update_pi0 / em_update_config / em_update_grid (eqtlbma_hm.cpp:659-868) with the per-gene
 * likelihoods KEPT by the last eqb_hm_loglik(keep = 1):
 *   out[0]            = sum_g 10^(log10 pi0 - lik_g)                      (new pi0 = out[0] / genes)
 *   out[1 + k]        = log10 sum_g 10^(config_genes[k][g])               k < dim   (before "+ log10(config_prior[k])")
 *   out[1 + dim + l]  = log10 sum_g 10^(grid_genes[l][g])                 l < grid */
int eqb_hm_esums(eqb_hm_ctx *hm, double pi0, const double *grid_wts, const double *config_prior, double *out);

/* Options of the fit (eqtlbma_hm.cpp:2106-2124 defaults in brackets). */
typedef struct {
  double thresh;         /* --thresh [0.05] */
  int64_t maxit;         /* --maxit, < 0 = none */
  double stepmax;        /* --msl [1 = classical EM], > 1: SQUAREM (run_EM_square, eqtlbma_hm.cpp:1215-1324) */
  int32_t fixed_pi0;     /* param2fixed_ (--pi0, third column of --init) */
  int32_t fixed_grid;
  int32_t fixed_configs;
  int32_t verbose;       /* > 0: the reference's progress lines (show_state_EM, eqtlbma_hm.cpp:870-923) go to `log` */
  void (*log)(void *user, const char *text); /* may be NULL */
  void *user;
} eqb_hm_options;

/* Parameters and their profile-likelihood intervals; caller-allocated arrays.  IN: initial values
 * (Controller::init_params); OUT: estimates. */
typedef struct {
  double pi0;
  double *grid_wts;     /* [grid] */
  double *config_prior; /* [dim] */
  double loglik;        /* log10 observed likelihood at the estimates */
  int64_t iters;        /* fixed-point iterations run */
  double pi0_ci[2];     /* left, right (NaN until eqb_hm_profile_ci) */
  double *grid_ci;      /* [grid][2] or NULL */
  double *config_ci;    /* [dim][2] or NULL */
} eqb_hm_fit;

/* Controller::run_EM (eqtlbma_hm.cpp:1076-1346). */
int eqb_hm_em(eqb_hm_ctx *hm, const eqb_hm_options *opt, eqb_hm_fit *fit);
/* Controller::estimate_profile_ci (eqtlbma_hm.cpp:1348-1573), tick 0.001, 2 log-likelihood units. */
int eqb_hm_profile_ci(eqb_hm_ctx *hm, eqb_hm_fit *fit);
/* Controller::compute_posterior + the Bayes-factor columns of save_result (eqtlbma_hm.cpp:1598-1748,
 * hm_methods.cpp:743-781) at the parameters of `fit`; any pointer may be NULL.
 *   gene_post[g], gene_bf[g]       gene.posterior.prob, gene.log10.bf
 *   snp_bf[p], snp_post[p]         snp.log10.bf, P(SNP p is the eQTL, gene is an eQTL gene | Y)
 *   cfg_bf[p][dim]                 log10.bf.<config>
 *   gene_cfg_post[g][dim]          post_prob_config_ */
int eqb_hm_posteriors(eqb_hm_ctx *hm, const eqb_hm_fit *fit, double *gene_post, double *gene_bf, double *snp_bf,
                      double *snp_post, double *cfg_bf, double *gene_cfg_post);

/* Measurement: `reps` launches of the streaming E-step kernel (hm_estep_kernel) with the given parameters, CUDA events on
 * the context's stream; ms = average per launch.  Algorithmic bytes per launch = 8 * pairs * dim * grid. */
int eqb_hm_estep_device_only(eqb_hm_ctx *hm, const double *grid_wts, const double *config_prior, int32_t reps, float *ms);
/* Kernel launches issued by this context so far. */
int64_t eqb_hm_launch_count(const eqb_hm_ctx *hm);

#ifdef __cplusplus
}
#endif
#endif
