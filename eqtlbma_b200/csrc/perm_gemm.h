// perm_gemm.h -- host interface of the batched-GEMM permutation path (perm_gemm.cu), used by eqtlbma_b200.cu.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

namespace eqb {

struct DevParams;
struct FastParams;

struct Perm2Env {
  int device = 0, n_sm = 148;
  cudaStream_t stream = nullptr;
  const DevParams *d_prm = nullptr, *hp = nullptr;   // device / host copies of the parameter block
  const FastParams *d_fp = nullptr, *hfp = nullptr;  // K1 outputs (fixed bases, x~'x~, t -> z tables)
  const long long *cb = nullptr, *ce = nullptr;      // host cis windows [G]
  const std::vector<double> *phi2L = nullptr, *oma2L = nullptr, *phi2S = nullptr, *oma2S = nullptr;
  const int *sub_xvar = nullptr;                     // [S] genotype variant of each subgroup
  double *const *d_X = nullptr;                      // [n_xvar] all-sample-space genotype matrices [M][ldn]
  int n_xvar = 0;
  const uint8_t *const *cell_generic = nullptr;      // [S] -> [G]: no NaN / absent sample inside the subgroup's individuals
  const uint8_t *sub_complete = nullptr;             // [S]: every sample of the union has genotype, expression, covariates
  int *d_err = nullptr;
  size_t free_bytes = 0;
};

struct Perm2State;
Perm2State *perm2_create();
void perm2_destroy(Perm2State *st, cudaStream_t stream);

// can this context / request run on the GEMM path?  (otherwise the caller keeps the general fused kernel)
bool perm2_supported(const Perm2Env &env, int which, int stat_kind);

// statistic of the true data (column -1) and of the P permuted data sets for n_items (gene, table) items:
// out_true[(i * per + s)], out_stat[(i * per + s) * P + p]; returns 0 or fills err
int perm2_eval(Perm2State *st, const Perm2Env &env, const int *genes, const int *tabs, size_t n_items,
               const unsigned short *d_perm, long long P, int which, int stat_kind, double *out_true, double *out_stat,
               long long *launches, std::string *err);

struct Perm2Timing { // accumulated device time of the last perm2_eval (CUDA events on env.stream), for rooflines
  float prep_ms = 0.f, gemm_ms = 0.f, bf_ms = 0.f, merge_ms = 0.f;
  double gemm_flops = 0.0; // 2 * 128 * 128 * ldn per tile issued (includes the padding of partial tiles)
  double gemm_useful_flops = 0.0; // rows x columns that belong to a (SNP, subgroup, permutation)
  long long bf_items = 0;  // (SNP, column) items
};
const Perm2Timing &perm2_last_timing(const Perm2State *st);
void perm2_set_timing(Perm2State *st, bool on);

} // namespace eqb
