// hm.cu -- the EM of the hierarchical model (eqtlbma_hm --model configs) behind the C ABI of include/eqtlbma_hm_b200.h:
// device data set, the three device operations (heavy pass, likelihood, E-step sums) and the host-side control flow of
// Controller::run_EM_classic / run_EM_square / estimate_profile_ci / compute_posterior (src/eqtlbma_hm.cpp:659-1610),
// restated around them.  The reference makes one full pass over the data for every log10_weighted_sum nest; here a pass
// (hm_estep_kernel) is made only when the grid weights or the configuration prior change, and everything that depends on
// pi0 alone (likelihood, pi0 sum, the whole pi0 profile of the confidence intervals) reuses its per-gene results.
#include "../../include/eqtlbma_hm_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "hm_kernels.cuh"

using namespace eqb;

struct eqb_hm_ctx {
  int device = 0, dim = 0, grid = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::string err;
  // data
  char *d_alloc = nullptr;   // allocation: 16 bytes | rows x grid doubles | 16 bytes
  size_t cap_pairs = 0;
  long long n_pairs = 0;
  std::vector<long long> gene_off{0};
  bool finalized = false;
  // work units and per-gene scratch
  long long n_units = 0;
  long long *d_unit_row0 = nullptr, *d_gene_unit0 = nullptr, *d_gene_off = nullptr;
  int *d_unit_rows = nullptr;
  double *d_U = nullptr, *d_PA = nullptr, *d_BF = nullptr, *d_kept_lik = nullptr, *d_kept_bf = nullptr;
  double *d_cfg = nullptr, *d_gw = nullptr, *d_out = nullptr, *d_rowA = nullptr, *d_snp = nullptr;
  double *h_out = nullptr; // pinned
  HmArgs args;
  size_t smem = 0;
  // cache of the heavy pass and of the E-step sums
  std::vector<double> heavy_gw, heavy_cfg;
  bool heavy_valid = false;
  std::vector<double> sums_gw, sums_cfg, sums_val;
  double sums_pi0 = 0;
  bool sums_valid = false;
  long long launches = 0, heavy_passes = 0;
  int estep_ctas = 0; // resident CTAs of hm_estep_kernel on this device (persistent grid)
  int *d_counter = nullptr;
  double *d_lik_partial = nullptr; // per-CTA partial sums of hm_lik_kernel
  unsigned int *d_ticket = nullptr;
  // multi-GPU: genes sharded over ranks, partial sums exchanged through the caller's all-gather (eqb_hm_set_collective)
  int world = 1, rank = 0;
  eqb_hm_allgather_fn gather = nullptr;
  void *gather_user = nullptr;
  double total_genes = 0; // over all ranks
  // native exchange over peer memory (eqb_hm_ipc_export / eqb_hm_ipc_connect): hm_xchg_kernel
  bool native = false;
  void *d_xchg = nullptr;
  int xchg_cap = 0;
  HmPeers peers{};
  std::vector<void *> opened;
  unsigned long long epoch = 0;
  int *d_status = nullptr;
  std::vector<double> gather_buf;
  bool ranged = false; // every value within +-1e6 (hm_check_kernel): unclamped exponentials of differences
};

static int fail(eqb_hm_ctx *hm, int code, const char *fmt, ...)
{
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (hm) hm->err = buf;
  return code;
}
#define HCK(call)                                                                                             \
  do {                                                                                                        \
    cudaError_t e_ = (call);                                                                                  \
    if (e_ != cudaSuccess) return fail(hm, 100, "CUDA error %s at %s:%d", cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

template <int G, bool EXACT, bool RANGED>
static cudaError_t launch_estep_t2(eqb_hm_ctx *hm)
{
  cudaError_t e = cudaFuncSetAttribute(hm_estep_kernel<G, EXACT, RANGED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hm->smem);
  if (e != cudaSuccess) return e;
  if (hm->estep_ctas <= 0) { // persistent CTAs: as many as are resident at once
    int per_sm = 0, n_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hm_estep_kernel<G, EXACT, RANGED>, HM_THREADS, hm->smem);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, hm->device);
    if (e != cudaSuccess) return e;
    hm->estep_ctas = std::max(1, per_sm) * std::max(1, n_sm);
  }
  const unsigned ctas = (unsigned)std::min<long long>((hm->n_units + HM_WARPS - 1) / HM_WARPS, hm->estep_ctas);
  e = cudaMemsetAsync(hm->d_counter, 0, sizeof(int), hm->stream);
  if (e != cudaSuccess) return e;
  hm_estep_kernel<G, EXACT, RANGED><<<ctas, HM_THREADS, hm->smem, hm->stream>>>(hm->args);
  return cudaGetLastError();
}
template <int G, bool EXACT>
static cudaError_t launch_estep_t(eqb_hm_ctx *hm)
{
  return hm->ranged ? launch_estep_t2<G, EXACT, true>(hm) : launch_estep_t2<G, EXACT, false>(hm);
}
static cudaError_t launch_estep(eqb_hm_ctx *hm)
{
  ++hm->launches;
  switch (hm->grid) {
  case 5: return launch_estep_t<5, true>(hm);
  case 10: return launch_estep_t<10, true>(hm);
  case 25: return launch_estep_t<25, true>(hm);
  default: break;
  }
  if (hm->grid <= 8) return launch_estep_t<8, false>(hm);
  if (hm->grid <= 16) return launch_estep_t<16, false>(hm);
  return launch_estep_t<32, false>(hm);
}

extern "C" {

int eqb_hm_create(eqb_hm_ctx **out, int32_t device, int32_t dim, int32_t grid)
{
  if (!out) return 1;
  *out = nullptr;
  eqb_hm_ctx *hm = new eqb_hm_ctx();
  *out = hm; // (kept on failure so that the message can be read; the caller destroys it)
  if (dim < 1 || dim > HM_MAXDIM) return fail(hm, 2, "eqb_hm_create: dim %d outside 1..%d", dim, HM_MAXDIM);
  if (grid < 1 || grid > HM_MAXGRID) return fail(hm, 2, "eqb_hm_create: grid %d outside 1..%d", grid, HM_MAXGRID);
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0)
    return fail(hm, 3, "eqb_hm_create: no CUDA device is usable (the hierarchical-model path has no CPU fallback)");
  if (device < 0 || device >= n_dev) return fail(hm, 3, "eqb_hm_create: device %d of %d", device, n_dev);
  hm->device = device;
  hm->dim = dim;
  hm->grid = grid;
  HCK(cudaSetDevice(device));
  HCK(cudaStreamCreateWithFlags(&hm->stream, cudaStreamNonBlocking));
  HCK(cudaEventCreate(&hm->ev0));
  HCK(cudaEventCreate(&hm->ev1));
  HCK(cudaMalloc(&hm->d_cfg, (size_t)dim * 8));
  HCK(cudaMalloc(&hm->d_gw, (size_t)HM_MAXGRID * 8));
  HCK(cudaMalloc(&hm->d_out, (size_t)(dim + grid + 2) * 8));
  HCK(cudaMallocHost(&hm->h_out, (size_t)(dim + grid + 2) * 8));
  return 0;
}

void eqb_hm_destroy(eqb_hm_ctx *hm)
{
  if (!hm) return;
  if (hm->stream) {
    cudaSetDevice(hm->device);
    cudaStreamSynchronize(hm->stream);
  }
  for (void *p : hm->opened) cudaIpcCloseMemHandle(p);
  void *ptrs[] = {hm->d_xchg, hm->d_status, hm->d_alloc, hm->d_counter, hm->d_lik_partial, hm->d_ticket, hm->d_unit_row0, hm->d_gene_unit0, hm->d_gene_off, hm->d_unit_rows, hm->d_U, hm->d_PA, hm->d_BF,
                  hm->d_kept_lik, hm->d_kept_bf, hm->d_cfg, hm->d_gw, hm->d_out, hm->d_rowA, hm->d_snp};
  for (void *p : ptrs)
    if (p) cudaFree(p);
  if (hm->h_out) cudaFreeHost(hm->h_out);
  if (hm->ev0) cudaEventDestroy(hm->ev0);
  if (hm->ev1) cudaEventDestroy(hm->ev1);
  if (hm->stream) cudaStreamDestroy(hm->stream);
  delete hm;
}

const char *eqb_hm_last_error(const eqb_hm_ctx *hm) { return hm ? hm->err.c_str() : "null context"; }
int64_t eqb_hm_n_genes(const eqb_hm_ctx *hm) { return hm ? (int64_t)hm->gene_off.size() - 1 : 0; }
int64_t eqb_hm_n_pairs(const eqb_hm_ctx *hm) { return hm ? hm->n_pairs : 0; }
int64_t eqb_hm_launch_count(const eqb_hm_ctx *hm) { return hm ? hm->launches : 0; }

static int append_impl(eqb_hm_ctx *hm, const double *B, int64_t n_pairs, const int64_t *gene_off, int64_t n_genes, bool on_device)
{
  if (!hm) return 1;
  if (hm->finalized) return fail(hm, 2, "eqb_hm_append: the data set is finalized");
  if (!B || !gene_off || n_pairs <= 0 || n_genes <= 0) return fail(hm, 2, "eqb_hm_append: empty input");
  if (gene_off[0] != 0 || gene_off[n_genes] != n_pairs) return fail(hm, 2, "eqb_hm_append: gene offsets do not span the pairs");
  for (int64_t g = 0; g < n_genes; ++g)
    if (gene_off[g + 1] <= gene_off[g]) return fail(hm, 2, "eqb_hm_append: gene %lld has no pair", (long long)g);
  HCK(cudaSetDevice(hm->device));
  const size_t pair_bytes = (size_t)hm->dim * hm->grid * 8;
  const size_t need = (size_t)hm->n_pairs + (size_t)n_pairs;
  if (need > hm->cap_pairs) {
    size_t cap = std::max(need, hm->cap_pairs + hm->cap_pairs / 2);
    char *fresh = nullptr;
    HCK(cudaMalloc(&fresh, cap * pair_bytes + 32));
    HCK(cudaMemsetAsync(fresh, 0, 16, hm->stream));
    if (hm->d_alloc) {
      HCK(cudaMemcpyAsync(fresh + 16, hm->d_alloc + 16, (size_t)hm->n_pairs * pair_bytes, cudaMemcpyDeviceToDevice, hm->stream));
      HCK(cudaStreamSynchronize(hm->stream));
      cudaFree(hm->d_alloc);
    }
    hm->d_alloc = fresh;
    hm->cap_pairs = cap;
  }
  HCK(cudaMemcpyAsync(hm->d_alloc + 16 + (size_t)hm->n_pairs * pair_bytes, B, (size_t)n_pairs * pair_bytes,
                      on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, hm->stream));
  HCK(cudaStreamSynchronize(hm->stream));
  const long long base = hm->n_pairs;
  for (int64_t g = 1; g <= n_genes; ++g) hm->gene_off.push_back(base + gene_off[g]);
  hm->n_pairs += n_pairs;
  return 0;
}

int eqb_hm_append(eqb_hm_ctx *hm, const double *B, int64_t n_pairs, const int64_t *gene_off, int64_t n_genes)
{
  return append_impl(hm, B, n_pairs, gene_off, n_genes, false);
}
int eqb_hm_append_device(eqb_hm_ctx *hm, const double *d_B, int64_t n_pairs, const int64_t *gene_off, int64_t n_genes)
{
  return append_impl(hm, d_B, n_pairs, gene_off, n_genes, true);
}

int eqb_hm_finalize(eqb_hm_ctx *hm)
{
  if (!hm) return 1;
  if (hm->finalized) return 0;
  if (hm->n_pairs <= 0) return fail(hm, 2, "eqb_hm_finalize: no data");
  HCK(cudaSetDevice(hm->device));
  const int dim = hm->dim, grid = hm->grid;
  const long long G = (long long)hm->gene_off.size() - 1;
  const size_t pair_bytes = (size_t)dim * grid * 8;
  HCK(cudaMemsetAsync(hm->d_alloc + 16 + (size_t)hm->n_pairs * pair_bytes, 0, 16, hm->stream));
  // non-finite values: the reference's likelihood is NaN / infinite and it stops (eqtlbma_hm.cpp:640-647)
  {
    unsigned long long *d_bad = nullptr, bad[2] = {0, 0};
    HCK(cudaMalloc(&d_bad, 16));
    HCK(cudaMemsetAsync(d_bad, 0, 16, hm->stream));
    hm_check_kernel<<<592, 256, 0, hm->stream>>>(reinterpret_cast<const double *>(hm->d_alloc + 16), hm->n_pairs * dim * grid, d_bad);
    ++hm->launches;
    HCK(cudaGetLastError());
    HCK(cudaMemcpyAsync(bad, d_bad, 16, cudaMemcpyDeviceToHost, hm->stream));
    HCK(cudaStreamSynchronize(hm->stream));
    cudaFree(d_bad);
    hm->ranged = bad[1] == 0;
    if (bad[0]) return fail(hm, 4, "ERROR: %llu raw log10(BF) values are NaN or +-Inf", bad[0]);
  }
  // work units: runs of whole pairs of one gene, about 1024 rows each (one warp per unit)
  const int pairs_per_unit = std::max(1, 1024 / dim);
  std::vector<long long> unit_row0, gene_unit0(G + 1);
  std::vector<int> unit_rows;
  for (long long g = 0; g < G; ++g) {
    gene_unit0[g] = (long long)unit_row0.size();
    for (long long p = hm->gene_off[g]; p < hm->gene_off[g + 1]; p += pairs_per_unit) {
      const long long pe = std::min(hm->gene_off[g + 1], p + pairs_per_unit);
      unit_row0.push_back(p * dim);
      unit_rows.push_back((int)((pe - p) * dim));
    }
  }
  gene_unit0[G] = (long long)unit_row0.size();
  hm->n_units = (long long)unit_row0.size();
  const int nout = dim + grid;
  HCK(cudaMalloc(&hm->d_unit_row0, hm->n_units * 8));
  HCK(cudaMalloc(&hm->d_counter, sizeof(int)));
  HCK(cudaMalloc(&hm->d_lik_partial, (size_t)((G + HM_LIK_THREADS - 1) / HM_LIK_THREADS) * 8));
  HCK(cudaMalloc(&hm->d_ticket, sizeof(unsigned int)));
  HCK(cudaMemsetAsync(hm->d_ticket, 0, sizeof(unsigned int), hm->stream));
  HCK(cudaMalloc(&hm->d_unit_rows, hm->n_units * 4));
  HCK(cudaMalloc(&hm->d_gene_unit0, (G + 1) * 8));
  HCK(cudaMalloc(&hm->d_gene_off, (G + 1) * 8));
  HCK(cudaMalloc(&hm->d_U, (size_t)hm->n_units * nout * 8));
  HCK(cudaMalloc(&hm->d_PA, (size_t)G * nout * 8));
  HCK(cudaMalloc(&hm->d_BF, G * 8));
  HCK(cudaMalloc(&hm->d_kept_lik, G * 8));
  HCK(cudaMalloc(&hm->d_kept_bf, G * 8));
  HCK(cudaMemcpyAsync(hm->d_unit_row0, unit_row0.data(), hm->n_units * 8, cudaMemcpyHostToDevice, hm->stream));
  HCK(cudaMemcpyAsync(hm->d_unit_rows, unit_rows.data(), hm->n_units * 4, cudaMemcpyHostToDevice, hm->stream));
  HCK(cudaMemcpyAsync(hm->d_gene_unit0, gene_unit0.data(), (G + 1) * 8, cudaMemcpyHostToDevice, hm->stream));
  HCK(cudaMemcpyAsync(hm->d_gene_off, hm->gene_off.data(), (G + 1) * 8, cudaMemcpyHostToDevice, hm->stream));
  HCK(cudaMemsetAsync(hm->d_kept_lik, 0, G * 8, hm->stream));
  HCK(cudaMemsetAsync(hm->d_kept_bf, 0, G * 8, hm->stream));
  HCK(cudaStreamSynchronize(hm->stream));
  HmArgs &a = hm->args;
  memset(&a, 0, sizeof(a));
  a.B = hm->d_alloc + 16;
  a.unit_row0 = hm->d_unit_row0;
  a.unit_rows = hm->d_unit_rows;
  a.U = hm->d_U;
  a.cfg = hm->d_cfg;
  a.rowA = nullptr;
  a.counter = hm->d_counter;
  a.dim = dim;
  a.grid = grid;
  if (hm->n_units > 0x7fffffffLL) return fail(hm, 2, "eqb_hm_finalize: too many work units");
  a.n_units = (int)hm->n_units;
  a.rpr = (dim >= 32) ? 32 : (32 / dim) * dim;
  a.nslot = (dim >= 32) ? dim : a.rpr;
#ifndef HM_STAGES_SMALL
#define HM_STAGES_SMALL 3
#endif
  a.stages = (grid <= 16) ? HM_STAGES_SMALL : 2;
  a.stage_bytes = (int)hm_align16((size_t)a.rpr * grid * 8 + 16);
  hm->smem = hm_smem_bytes(a.nslot, a.stages, a.stage_bytes);
  if (hm->smem > 200 * 1024) return fail(hm, 2, "eqb_hm_finalize: dim %d x grid %d needs %zu bytes of shared memory", dim, grid, hm->smem);
  hm->finalized = true;
  return 0;
}

} // extern "C"

// ---------------------------------------------------------------- device operations
// heavy pass: per-gene A[g][k], Gd[g][l], BF[g] for (grid weights, configuration prior); cached on the pair
static int heavy(eqb_hm_ctx *hm, const double *gw, const double *cfg, bool want_rows)
{
  const int dim = hm->dim, grid = hm->grid;
  if (!hm->finalized) return fail(hm, 2, "eqb_hm: eqb_hm_finalize() has not been called");
  if (!want_rows && hm->heavy_valid && memcmp(hm->heavy_gw.data(), gw, grid * 8) == 0 && memcmp(hm->heavy_cfg.data(), cfg, dim * 8) == 0)
    return 0;
  HCK(cudaSetDevice(hm->device));
  const long long G = (long long)hm->gene_off.size() - 1;
  hm->heavy_valid = false;
  hm->heavy_gw.assign(gw, gw + grid);
  hm->heavy_cfg.assign(cfg, cfg + dim);
  for (int l = 0; l < HM_MAXGRID; ++l) hm->args.gw[l] = (l < grid) ? gw[l] : 0.0;
  HCK(cudaMemcpyAsync(hm->d_cfg, hm->heavy_cfg.data(), (size_t)dim * 8, cudaMemcpyHostToDevice, hm->stream));
  HCK(cudaMemcpyAsync(hm->d_gw, hm->args.gw, (size_t)HM_MAXGRID * 8, cudaMemcpyHostToDevice, hm->stream));
  if (want_rows && !hm->d_rowA) HCK(cudaMalloc(&hm->d_rowA, (size_t)hm->n_pairs * dim * 8));
  hm->args.rowA = want_rows ? hm->d_rowA : nullptr;
  HCK(launch_estep(hm));
  hm->args.rowA = nullptr;
  hm_gene_kernel<<<(unsigned)((G + HM_GENE_WARPS - 1) / HM_GENE_WARPS), HM_GENE_WARPS * 32, 0, hm->stream>>>(hm->d_U, hm->d_gene_unit0, hm->d_gene_off, dim, grid, G, hm->d_gw, hm->d_PA, hm->d_BF);
  ++hm->launches;
  ++hm->heavy_passes;
  HCK(cudaGetLastError());
  hm->heavy_valid = true;
  return 0;
}

static bool same_vec(const std::vector<double> &a, const double *b, size_t n) { return a.size() == n && memcmp(a.data(), b, n * 8) == 0; }

// partial results of the ranks -> totals, in rank order (deterministic): entries 0 and 1 are plain sums (log-likelihood, pi0
// sum), the others log10 of sums (NaN propagates, -inf = empty)
static void combine_partials(const double *gathered, int world, int n, double *out)
{
  for (int j = 0; j < n; ++j) {
    if (j < 2) {
      double acc = 0.0;
      for (int r = 0; r < world; ++r) acc += gathered[(size_t)r * n + j];
      out[j] = acc;
      continue;
    }
    double mx = -INFINITY;
    bool bad = false;
    for (int r = 0; r < world; ++r) {
      const double v = gathered[(size_t)r * n + j];
      bad = bad || (v != v);
      mx = fmax(mx, v);
    }
    if (bad)
      out[j] = NAN;
    else if (!(mx > -INFINITY) || std::isinf(mx))
      out[j] = mx;
    else {
      double acc = 0.0;
      for (int r = 0; r < world; ++r) acc += pow(10.0, gathered[(size_t)r * n + j] - mx);
      out[j] = mx + log10(acc);
    }
  }
}
// native mode: enqueue the peer-memory exchange of d_out[0..n) on the context's stream (before the read-back)
static int exchange_device(eqb_hm_ctx *hm, size_t n)
{
  if (hm->world <= 1 || !hm->native) return 0;
  if ((int)n > hm->xchg_cap) return fail(hm, 7, "eqb_hm: exchange of %zu values exceeds the buffer", n);
  ++hm->epoch;
  hm_xchg_kernel<<<1, 256, 0, hm->stream>>>(hm->peers, hm->world, hm->rank, (int)n, hm->xchg_cap, hm->epoch, hm->d_out, hm->d_status);
  ++hm->launches;
  HCK(cudaGetLastError());
  return 0;
}
static int exchange_status(eqb_hm_ctx *hm)
{
  if (hm->world <= 1 || !hm->native) return 0;
  int st = 0;
  HCK(cudaMemcpyAsync(&st, hm->d_status, sizeof(int), cudaMemcpyDeviceToHost, hm->stream));
  HCK(cudaStreamSynchronize(hm->stream));
  if (st) return fail(hm, 7, "eqb_hm: a peer did not publish its partial sums within the time limit");
  return 0;
}

static int exchange(eqb_hm_ctx *hm, double *v, size_t n)
{
  if (hm->world <= 1 || !hm->gather || hm->native) return 0;
  hm->gather_buf.resize((size_t)hm->world * n);
  if (hm->gather(hm->gather_user, v, hm->gather_buf.data(), (int32_t)n) != 0) return fail(hm, 7, "eqb_hm: the all-gather callback failed");
  combine_partials(hm->gather_buf.data(), hm->world, (int)n, v);
  return 0;
}

// compute_log10_obs_lik; with keep, the E-step sums of the same parameters are produced in the same breath and cached
static int loglik(eqb_hm_ctx *hm, double pi0, const double *gw, const double *cfg, bool keep, double *out)
{
  const int dim = hm->dim, grid = hm->grid;
  const long long G = (long long)hm->gene_off.size() - 1;
  int rc = heavy(hm, gw, cfg, false);
  if (rc) return rc;
  hm_lik_kernel<<<(unsigned)((G + HM_LIK_THREADS - 1) / HM_LIK_THREADS), HM_LIK_THREADS, 0, hm->stream>>>(hm->d_BF, G, pi0, keep ? 1 : 0, hm->d_kept_lik, hm->d_kept_bf,
                                                                                                         hm->d_lik_partial, hm->d_ticket, hm->d_out);
  ++hm->launches;
  size_t n_out = 1;
  if (keep) {
    hm_sums_kernel<<<dim + grid + 1, HM_SUMS_THREADS, 0, hm->stream>>>(hm->d_PA, hm->d_kept_lik, G, dim + grid, pi0, hm->d_out);
    ++hm->launches;
    n_out = (size_t)dim + grid + 2;
    hm->sums_valid = false;
  }
  HCK(cudaGetLastError());
  if ((rc = exchange_device(hm, n_out))) return rc;
  HCK(cudaMemcpyAsync(hm->h_out, hm->d_out, n_out * 8, cudaMemcpyDeviceToHost, hm->stream));
  HCK(cudaStreamSynchronize(hm->stream));
  if ((rc = exchange_status(hm))) return rc;
  if ((rc = exchange(hm, hm->h_out, n_out))) return rc;
  *out = hm->h_out[0];
  if (keep) {
    hm->sums_pi0 = pi0;
    hm->sums_gw.assign(gw, gw + grid);
    hm->sums_cfg.assign(cfg, cfg + dim);
    hm->sums_val.assign(hm->h_out + 1, hm->h_out + n_out);
    hm->sums_valid = true;
  }
  return 0;
}

// E-step sums with the KEPT per-gene likelihoods: out[0] pi0 sum, out[1 + k], out[1 + dim + l]
static int esums(eqb_hm_ctx *hm, double pi0, const double *gw, const double *cfg, double *out)
{
  const int dim = hm->dim, grid = hm->grid;
  const long long G = (long long)hm->gene_off.size() - 1;
  const size_t n = (size_t)dim + grid + 1;
  if (hm->sums_valid && memcmp(&hm->sums_pi0, &pi0, 8) == 0 && same_vec(hm->sums_gw, gw, grid) && same_vec(hm->sums_cfg, cfg, dim)) {
    memcpy(out, hm->sums_val.data(), n * 8);
    return 0;
  }
  // parameters differ from those of the kept likelihoods (the reference's stale-likelihood corner after a rejected
  // SQUAREM extrapolation, eqtlbma_hm.cpp:1296-1307): heavy pass for these parameters, sums with the kept values
  int rc = heavy(hm, gw, cfg, false);
  if (rc) return rc;
  hm_sums_kernel<<<dim + grid + 1, HM_SUMS_THREADS, 0, hm->stream>>>(hm->d_PA, hm->d_kept_lik, G, dim + grid, pi0, hm->d_out);
  ++hm->launches;
  HCK(cudaGetLastError());
  if (hm->native) HCK(cudaMemsetAsync(hm->d_out, 0, 8, hm->stream)); // (slot of the log-likelihood: not produced here)
  if ((rc = exchange_device(hm, n + 1))) return rc;
  HCK(cudaMemcpyAsync(hm->h_out, hm->d_out, (n + 1) * 8, cudaMemcpyDeviceToHost, hm->stream));
  HCK(cudaStreamSynchronize(hm->stream));
  if ((rc = exchange_status(hm))) return rc;
  hm->h_out[0] = 0.0;
  if ((rc = exchange(hm, hm->h_out, n + 1))) return rc;
  memcpy(out, hm->h_out + 1, n * 8);
  return 0;
}

// ---------------------------------------------------------------- host control flow (Controller, eqtlbma_hm.cpp)
namespace {

// utils::log10_weighted_sum (utils_math.cpp:135-159)
double lws(const double *vec, const double *w, size_t n)
{
  double mx = vec[0];
  for (size_t i = 0; i < n; ++i)
    if (vec[i] > mx) mx = vec[i];
  double sum = 0.0;
  for (size_t i = 0; i < n; ++i)
    if (!std::isnan(vec[i])) sum += w[i] * pow(10.0, vec[i] - mx);
  double res = mx + log10(sum);
  if (std::fabs(res) <= DBL_EPSILON) res = 0.0;
  return res;
}

constexpr int EXC = 1; // the reference's `throw 1`

struct Em {
  eqb_hm_ctx *hm;
  const eqb_hm_options *o;
  int dim, grid;
  double G;
  double lik = NAN, lik0 = NAN, lik1 = NAN, lik2 = NAN, new_lik = NAN;
  double pi0 = NAN, pi0_0 = NAN, pi0_1 = NAN, pi0_2 = NAN, new_pi0 = NAN;
  std::vector<double> gw, gw0, gw1, gw2, new_gw, cfg, cfg0, cfg1, cfg2, new_cfg, ones, sums;
  long long fixedpoints = 0;

  Em(eqb_hm_ctx *h, const eqb_hm_options *opt) : hm(h), o(opt), dim(h->dim), grid(h->grid), G(h->world > 1 ? h->total_genes : (double)(h->gene_off.size() - 1))
  {
    gw.assign(grid, NAN);
    gw0 = gw1 = gw2 = new_gw = gw;
    cfg.assign(dim, NAN);
    cfg0 = cfg1 = cfg2 = new_cfg = cfg;
    ones.assign(std::max(dim, grid), 1.0);
    sums.assign((size_t)dim + grid + 1, 0.0);
  }
  void say(const char *fmt, ...)
  {
    if (!o->log || o->verbose <= 0) return;
    char buf[256];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    o->log(o->user, buf);
  }
  // show_state_EM (eqtlbma_hm.cpp:870-923)
  void show(size_t iter)
  {
    if (!o->log || o->verbose <= 0) return;
    std::string s;
    char b[64];
    snprintf(b, sizeof(b), "iter %4zu", iter);
    s += b;
    snprintf(b, sizeof(b), "  loglik %f", iter == 0 ? lik : new_lik);
    s += b;
    snprintf(b, sizeof(b), "  pi0 %7.4e", iter == 0 ? pi0 : new_pi0);
    s += b;
    s += "  configs";
    for (int k = 0; k < dim; ++k) {
      snprintf(b, sizeof(b), " %7.4e", iter == 0 ? cfg[k] : new_cfg[k]);
      s += b;
    }
    s += "  grid-points";
    for (int l = 0; l < grid; ++l) {
      snprintf(b, sizeof(b), " %7.4e", iter == 0 ? gw[l] : new_gw[l]);
      s += b;
    }
    s += "\n";
    o->log(o->user, s.c_str());
  }
  // compute_log10_obs_lik (eqtlbma_hm.cpp:617-650): EXC for a NaN or infinite sum
  int obs_lik(double p, const std::vector<double> &w, const std::vector<double> &c, bool keep, double *out)
  {
    int rc = loglik(hm, p, w.data(), c.data(), keep, out);
    if (rc) return rc;
    if (std::isnan(*out)) {
      hm->err = "ERROR: log10(obslik) is NaN";
      return EXC;
    }
    if (std::isinf(*out)) {
      hm->err = "ERROR: log10(obslik) is +-Inf";
      return EXC;
    }
    return 0;
  }
  // run_EM_fixedpoint (eqtlbma_hm.cpp:925-1011)
  int fixedpoint(size_t iter)
  {
    ++fixedpoints;
    int rc = esums(hm, pi0, gw.data(), cfg.data(), sums.data());
    if (rc) return rc;
    new_pi0 = o->fixed_pi0 ? pi0 : sums[0] / G; // em_update_pi0
    if (dim > 1) {                              // em_update_config
      if (!o->fixed_configs) {
        for (int k = 0; k < dim; ++k) {
          new_cfg[k] = sums[1 + k] + log10(cfg[k]);
          if (std::isnan(new_cfg[k])) {
            char b[96];
            snprintf(b, sizeof(b), "ERROR: new_config_prior_[%d] is NaN", k);
            hm->err = b;
            return EXC;
          }
        }
        const double denom = lws(new_cfg.data(), ones.data(), dim);
        for (int k = 0; k < dim; ++k) new_cfg[k] = pow(10.0, new_cfg[k] - denom);
      } else
        new_cfg = cfg;
    } else
      new_cfg = cfg;
    if (!o->fixed_grid) { // em_update_grid
      for (int l = 0; l < grid; ++l) {
        new_gw[l] = sums[1 + dim + l] + log10(gw[l]);
        if (std::isnan(new_gw[l])) {
          char b[96];
          snprintf(b, sizeof(b), "ERROR: new_grid_wts_[%d] is NaN", l);
          hm->err = b;
          return EXC;
        }
      }
      const double denom = lws(new_gw.data(), ones.data(), grid);
      for (int l = 0; l < grid; ++l) new_gw[l] = pow(10.0, new_gw[l] - denom);
    } else
      new_gw = gw;
    rc = obs_lik(new_pi0, new_gw, new_cfg, true, &new_lik);
    if (rc) return rc;
    show(iter);
    return 0;
  }
  // update_params (eqtlbma_hm.cpp:1013-1074), including the fall-through from step 1 into step 2
  void update(int step = 0)
  {
    pi0 = new_pi0;
    if (dim > 1) cfg = new_cfg;
    gw = new_gw;
    lik = new_lik;
    if (step == 1) {
      pi0_1 = new_pi0;
      if (dim > 1) cfg1 = new_cfg;
      gw1 = new_gw;
      lik1 = new_lik;
    }
    if (step == 1 || step == 2) {
      pi0_2 = new_pi0;
      if (dim > 1) cfg2 = new_cfg;
      gw2 = new_gw;
      lik2 = new_lik;
    }
  }
  bool last_iter(size_t iter) const { return o->maxit >= 0 && iter == (size_t)o->maxit - 1; }
  // run_EM_classic (eqtlbma_hm.cpp:1076-1107)
  int classic()
  {
    size_t iter = 0;
    int rc = obs_lik(pi0, gw, cfg, true, &lik);
    if (rc) return rc;
    show(iter);
    while (true) {
      ++iter;
      if ((rc = fixedpoint(iter))) return rc;
      if (new_lik < lik) {
        char b[160];
        snprintf(b, sizeof(b), "ERROR: observed log-likelihood is decreasing (%f < %f)", new_lik, lik);
        hm->err = b;
        return 5;
      }
      if (std::fabs(new_lik - lik) < o->thresh || last_iter(iter)) break;
      update();
    }
    update();
    ++iter;
    if ((rc = fixedpoint(iter))) return rc;
    update();
    return 0;
  }
  // compute_steplength (eqtlbma_hm.cpp:1109-1167), equation 9 of Varadhan & Roland (2008)
  double steplength(double stepmin, double stepmax) const
  {
    double sr2 = pow(pi0_1 - pi0_0, 2), sv2 = pow(pi0_2 - 2 * pi0_1 + pi0_0, 2);
    if (dim > 1)
      for (int k = 0; k < dim; ++k) {
        sr2 += pow(cfg1[k] - cfg0[k], 2);
        sv2 += pow(cfg2[k] - 2 * cfg1[k] + cfg0[k], 2);
      }
    for (int l = 0; l < grid; ++l) {
      sr2 += pow(gw1[l] - gw0[l], 2);
      sv2 += pow(gw2[l] - 2 * gw1[l] + gw0[l], 2);
    }
    double alpha = sqrt(sr2 / sv2);
    alpha = std::max(stepmin, std::min(stepmax, alpha));
    alpha = std::min(o->stepmax, alpha);
    return alpha;
  }
  // proposals_squarem (eqtlbma_hm.cpp:1169-1213)
  int proposals(double alpha)
  {
    new_pi0 = pi0_0 + 2.0 * alpha * (pi0_1 - pi0_0) + pow(alpha, 2) * (pi0_2 - 2.0 * pi0_1 + pi0_0);
    if (dim > 1)
      for (int k = 0; k < dim; ++k)
        new_cfg[k] = cfg0[k] + 2.0 * alpha * (cfg1[k] - cfg0[k]) + pow(alpha, 2) * (cfg2[k] - 2.0 * cfg1[k] + cfg0[k]);
    for (int l = 0; l < grid; ++l)
      new_gw[l] = gw0[l] + 2.0 * alpha * (gw1[l] - gw0[l]) + pow(alpha, 2) * (gw2[l] - 2.0 * gw1[l] + gw0[l]);
    int rc = obs_lik(new_pi0, new_gw, new_cfg, true, &new_lik);
    if (rc) return rc;
    show(99999);
    return 0;
  }
  // run_EM_square (eqtlbma_hm.cpp:1215-1324)
  int square()
  {
    const double stepmin0 = 1, stepmax0 = 1, mstep = 4, maxdist = 1;
    size_t iter = 0, iter_main = 0;
    double steplen = 1.0, stepmin = stepmin0, stepmax = stepmax0;
    bool extrap = false;
    int rc = obs_lik(pi0, gw, cfg, true, &lik);
    if (rc) return rc;
    show(iter);
    while (true) {
      ++iter_main;
      say("main loop iter %zu\n", iter_main);
      lik0 = lik;
      pi0_0 = pi0;
      gw0 = gw;
      cfg0 = cfg;
      ++iter;
      if ((rc = fixedpoint(iter))) return rc;
      update(1);
      if (std::fabs(lik1 - lik0) < o->thresh || last_iter(iter)) break;
      ++iter;
      if ((rc = fixedpoint(iter))) return rc;
      update(2);
      if (std::fabs(lik2 - lik1) < o->thresh || last_iter(iter)) break;
      steplen = steplength(stepmin, stepmax);
      say("steplen %f\n", steplen);
      if ((rc = proposals(steplen))) return rc;
      update();
      extrap = true;
      if (std::fabs(steplen - 1) > 0.01) {
        say("step length is large, check consistency\n");
        ++iter;
        rc = fixedpoint(iter);
        if (rc == EXC) {
          say("failure, go back to previous iter\n");
          pi0 = pi0_2;
          if (dim > 1) cfg = cfg2;
          gw = gw2;
          if ((rc = obs_lik(pi0, gw, cfg, true, &lik))) return rc;
          extrap = false;
          if (steplen == stepmax) stepmax = std::max(stepmax0, stepmax / mstep);
          steplen = 1;
          if (steplen == stepmax) stepmax = mstep * stepmax;
          if (stepmin < 0 && steplen == stepmin) stepmin = mstep * stepmin;
          continue;
        } else if (rc)
          return rc;
        say("success, keep going\n");
        update();
      }
      if (extrap && lik < lik0 - maxdist) {
        say("log-lik after squarem is too bad, keep classical iter\n");
        lik = lik2;
        pi0 = pi0_2;
        gw = gw2;
        if (dim > 1) cfg = cfg2;
        if (steplen == stepmax) stepmax = std::max(stepmax0, stepmax / mstep);
        steplen = 1;
      }
      if (steplen == stepmax) stepmax = mstep * stepmax;
      if (stepmin < 0 && steplen == stepmin) stepmin = mstep * stepmin;
      say("extrap %s  stepmax %f  stepmin %f\n", extrap ? "true" : "false", stepmax, stepmin);
    }
    ++iter;
    if ((rc = fixedpoint(iter))) return rc;
    update();
    return 0;
  }
};

bool usable_fit(eqb_hm_ctx *hm, const eqb_hm_fit *fit)
{
  if (!fit || !fit->grid_wts || !fit->config_prior) {
    if (hm) hm->err = "eqb_hm: the fit needs grid_wts and config_prior arrays";
    return false;
  }
  return true;
}

} // namespace

extern "C" {

int eqb_hm_set_collective(eqb_hm_ctx *hm, int32_t world, int32_t rank, eqb_hm_allgather_fn fn, void *user)
{
  if (!hm) return 1;
  if (!hm->finalized) return fail(hm, 2, "eqb_hm_set_collective: call eqb_hm_finalize() first");
  if (world < 1 || rank < 0 || rank >= world || (world > 1 && !fn)) return fail(hm, 2, "eqb_hm_set_collective: invalid arguments");
  hm->world = world;
  hm->rank = rank;
  hm->gather = fn;
  hm->gather_user = user;
  hm->native = false;
  hm->sums_valid = false;
  hm->total_genes = (double)(hm->gene_off.size() - 1);
  if (world > 1) {
    double mine[2] = {hm->total_genes, 0.0};
    int rc = exchange(hm, mine, 2); // (two plain sums)
    if (rc) return rc;
    hm->total_genes = mine[0];
  }
  return 0;
}

int eqb_hm_ipc_export(eqb_hm_ctx *hm, void *handle64)
{
  if (!hm || !handle64) return 1;
  if (!hm->finalized) return fail(hm, 2, "eqb_hm_ipc_export: call eqb_hm_finalize() first");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  HCK(cudaSetDevice(hm->device));
  if (!hm->d_xchg) {
    hm->xchg_cap = hm->dim + hm->grid + 2;
    const size_t bytes = (size_t)HM_XCHG_DATA_OFF + (size_t)2 * HM_MAXWORLD * hm->xchg_cap * 8;
    HCK(cudaMalloc(&hm->d_xchg, bytes));
    HCK(cudaMemset(hm->d_xchg, 0, bytes));
    HCK(cudaMalloc(&hm->d_status, sizeof(int)));
    HCK(cudaMemset(hm->d_status, 0, sizeof(int)));
    HCK(cudaDeviceSynchronize());
  }
  cudaIpcMemHandle_t h;
  HCK(cudaIpcGetMemHandle(&h, hm->d_xchg));
  memcpy(handle64, &h, 64);
  return 0;
}

int eqb_hm_ipc_connect(eqb_hm_ctx *hm, int32_t world, int32_t rank, const void *handles)
{
  if (!hm || !handles) return 1;
  if (!hm->d_xchg) return fail(hm, 2, "eqb_hm_ipc_connect: call eqb_hm_ipc_export() first");
  if (world < 1 || world > HM_MAXWORLD || rank < 0 || rank >= world) return fail(hm, 2, "eqb_hm_ipc_connect: world must be 1..%d", HM_MAXWORLD);
  HCK(cudaSetDevice(hm->device));
  const bool again = !hm->opened.empty() && hm->world == world && hm->rank == rank; // peers already mapped: switch back to native
  for (int r = 0; r < world && !again; ++r) {
    if (r == rank) {
      hm->peers.base[r] = hm->d_xchg;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const char *>(handles) + (size_t)r * 64, 64);
    void *p = nullptr;
    HCK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    hm->opened.push_back(p);
    hm->peers.base[r] = p;
  }
  hm->world = world;
  hm->rank = rank;
  hm->native = world > 1;
  hm->sums_valid = false;
  hm->total_genes = (double)(hm->gene_off.size() - 1);
  if (world > 1) { // total number of genes: the first exchange (two plain sums)
    double mine[2] = {hm->total_genes, 0.0};
    HCK(cudaMemcpyAsync(hm->d_out, mine, 16, cudaMemcpyHostToDevice, hm->stream));
    int rc = exchange_device(hm, 2);
    if (rc) return rc;
    HCK(cudaMemcpyAsync(mine, hm->d_out, 16, cudaMemcpyDeviceToHost, hm->stream));
    HCK(cudaStreamSynchronize(hm->stream));
    if ((rc = exchange_status(hm))) return rc;
    hm->total_genes = mine[0];
  }
  return 0;
}

int eqb_hm_combine_partials(const double *gathered, int32_t world, int32_t n, double *out)
{
  if (!gathered || !out || world < 1 || n < 0) return 1;
  combine_partials(gathered, world, n, out);
  return 0;
}

int eqb_hm_loglik(eqb_hm_ctx *hm, double pi0, const double *grid_wts, const double *config_prior, int32_t keep, double *out)
{
  if (!hm || !grid_wts || !config_prior || !out) return 1;
  return loglik(hm, pi0, grid_wts, config_prior, keep != 0, out);
}

int eqb_hm_esums(eqb_hm_ctx *hm, double pi0, const double *grid_wts, const double *config_prior, double *out)
{
  if (!hm || !grid_wts || !config_prior || !out) return 1;
  if (!hm->finalized) return fail(hm, 2, "eqb_hm: eqb_hm_finalize() has not been called");
  return esums(hm, pi0, grid_wts, config_prior, out);
}

int eqb_hm_em(eqb_hm_ctx *hm, const eqb_hm_options *opt, eqb_hm_fit *fit)
{
  if (!hm || !opt) return 1;
  if (!usable_fit(hm, fit)) return 2;
  if (!hm->finalized) return fail(hm, 2, "eqb_hm: eqb_hm_finalize() has not been called");
  if (!(opt->thresh > 0.0)) return fail(hm, 2, "ERROR: --thresh %g is invalid", opt->thresh);
  Em em(hm, opt);
  em.pi0 = fit->pi0;
  em.gw.assign(fit->grid_wts, fit->grid_wts + hm->grid);
  em.cfg.assign(fit->config_prior, fit->config_prior + hm->dim);
  em.new_cfg = em.cfg; // (dim == 1: Controller::init_params sets new_config_prior_[0] = 1 as well)
  const int rc = (opt->stepmax == 1.0) ? em.classic() : em.square(); // run_EM (eqtlbma_hm.cpp:1338-1341)
  if (rc) return rc == EXC ? 6 : rc;
  fit->pi0 = em.pi0;
  std::copy(em.gw.begin(), em.gw.end(), fit->grid_wts);
  std::copy(em.cfg.begin(), em.cfg.end(), fit->config_prior);
  fit->loglik = em.lik;
  fit->iters = em.fixedpoints;
  fit->pi0_ci[0] = fit->pi0_ci[1] = NAN;
  return 0;
}

// estimate_profile_ci (eqtlbma_hm.cpp:1348-1573): walk each parameter away from its estimate in ticks of 0.001 (the others
// rescaled to keep the sum) until the log-likelihood has dropped by 2 natural-log units
int eqb_hm_profile_ci(eqb_hm_ctx *hm, eqb_hm_fit *fit)
{
  if (!hm) return 1;
  if (!usable_fit(hm, fit)) return 2;
  if (!hm->finalized) return fail(hm, 2, "eqb_hm: eqb_hm_finalize() has not been called");
  const int dim = hm->dim, grid = hm->grid;
  const double tick = 0.001, l10e = log10(exp(1.0));
  const std::vector<double> gw_mle(fit->grid_wts, fit->grid_wts + grid), cfg_mle(fit->config_prior, fit->config_prior + dim);
  const double pi0_mle = fit->pi0;
  double max_lik = 0, cur = 0;
  int rc;
  auto lik_at = [&](double p, const std::vector<double> &w, const std::vector<double> &c, double *out) -> int {
    int r = loglik(hm, p, w.data(), c.data(), true, out);
    if (r) return r;
    if (std::isnan(*out) || std::isinf(*out)) return fail(hm, 6, std::isnan(*out) ? "ERROR: log10(obslik) is NaN" : "ERROR: log10(obslik) is +-Inf");
    return 0;
  };
  if ((rc = lik_at(pi0_mle, gw_mle, cfg_mle, &max_lik))) return rc;
  const double floor_ln = max_lik / l10e - 2.0;
  // pi0 (eqtlbma_hm.cpp:1348-1391)
  double left = pi0_mle, right = pi0_mle;
  while (left >= 0) {
    left -= tick;
    if (left < 0) {
      left = 0;
      break;
    }
    if ((rc = lik_at(left, gw_mle, cfg_mle, &cur))) return rc;
    if (cur / l10e < floor_ln) {
      left += tick;
      break;
    }
  }
  while (right <= 1) {
    right += tick;
    if (right > 1) {
      right = 1;
      break;
    }
    if ((rc = lik_at(right, gw_mle, cfg_mle, &cur))) return rc;
    if (cur / l10e < floor_ln) {
      right -= tick;
      break;
    }
  }
  fit->pi0_ci[0] = left;
  fit->pi0_ci[1] = right;
  // one simplex parameter vector at a time (configs: eqtlbma_hm.cpp:1393-1453, grid points: 1485-1545)
  auto profile = [&](const std::vector<double> &mle, bool is_cfg, double *ci) -> int {
    std::vector<double> v(mle);
    const size_t n = mle.size();
    for (size_t i = 0; i < n; ++i) {
      double lo = mle[i], hi = mle[i];
      const double cp = mle[i], st = 1 - cp;
      for (int side = 0; side < 2; ++side) {
        double &edge = side == 0 ? lo : hi;
        while (side == 0 ? edge >= 0 : edge <= 1) {
          edge += side == 0 ? -tick : tick;
          if (side == 0 ? edge < 0 : edge > 1) {
            edge = side == 0 ? 0 : 1;
            break;
          }
          const double diff = cp - edge;
          for (size_t j = 0; j < n; ++j)
            if (j != i) v[j] = mle[j] + diff * mle[j] / st;
          v[i] = edge;
          int r = is_cfg ? lik_at(pi0_mle, gw_mle, v, &cur) : lik_at(pi0_mle, v, cfg_mle, &cur);
          if (r) return r;
          if (cur / l10e < floor_ln) {
            edge += side == 0 ? tick : -tick;
            break;
          }
        }
      }
      if (ci) {
        ci[2 * i] = lo;
        ci[2 * i + 1] = hi;
      }
    }
    return 0;
  };
  if ((rc = profile(cfg_mle, true, fit->config_ci))) return rc;
  if ((rc = profile(gw_mle, false, fit->grid_ci))) return rc;
  return 0;
}

int eqb_hm_posteriors(eqb_hm_ctx *hm, const eqb_hm_fit *fit, double *gene_post, double *gene_bf, double *snp_bf, double *snp_post,
                      double *cfg_bf, double *gene_cfg_post)
{
  if (!hm) return 1;
  if (!usable_fit(hm, fit)) return 2;
  if (!hm->finalized) return fail(hm, 2, "eqb_hm: eqb_hm_finalize() has not been called");
  HCK(cudaSetDevice(hm->device));
  const int dim = hm->dim, grid = hm->grid;
  const long long G = (long long)hm->gene_off.size() - 1, P = hm->n_pairs;
  const double pi0 = fit->pi0;
  const bool rows = snp_bf || snp_post || cfg_bf;
  int rc = heavy(hm, fit->grid_wts, fit->config_prior, rows);
  if (rc) return rc;
  double lik = 0;
  if ((rc = loglik(hm, pi0, fit->grid_wts, fit->config_prior, true, &lik))) return rc;
  std::vector<double> bf(G), gl(G);
  HCK(cudaMemcpyAsync(bf.data(), hm->d_kept_bf, G * 8, cudaMemcpyDeviceToHost, hm->stream));
  HCK(cudaMemcpyAsync(gl.data(), hm->d_kept_lik, G * 8, cudaMemcpyDeviceToHost, hm->stream));
  std::vector<double> sb;
  if (snp_bf || snp_post) {
    if (!hm->d_snp) HCK(cudaMalloc(&hm->d_snp, (size_t)P * 8));
    hm_snp_kernel<<<(unsigned)((P + 3) / 4), 128, 0, hm->stream>>>(hm->d_rowA, hm->d_cfg, dim, P, hm->d_snp);
    ++hm->launches;
    HCK(cudaGetLastError());
    sb.resize(P);
    HCK(cudaMemcpyAsync(sb.data(), hm->d_snp, (size_t)P * 8, cudaMemcpyDeviceToHost, hm->stream));
  }
  if (cfg_bf) HCK(cudaMemcpyAsync(cfg_bf, hm->d_rowA, (size_t)P * dim * 8, cudaMemcpyDeviceToHost, hm->stream));
  std::vector<double> pa;
  if (gene_cfg_post) {
    pa.resize((size_t)G * dim);
    HCK(cudaMemcpyAsync(pa.data(), hm->d_PA, (size_t)G * dim * 8, cudaMemcpyDeviceToHost, hm->stream));
  }
  HCK(cudaStreamSynchronize(hm->stream));
  // gene_eQTL::compute_posterior (hm_methods.cpp:743-781)
  const double l1 = log10(1.0 - pi0);
  for (long long g = 0; g < G; ++g) {
    if (gene_bf) {
      double v = bf[g];
      if (std::fabs(v) <= DBL_EPSILON) v = 0.0;
      gene_bf[g] = v;
    }
    if (gene_post) gene_post[g] = std::min(1.0, pow(10.0, l1 + bf[g] - gl[g]));
    if (snp_bf || snp_post) {
      const long long p0 = hm->gene_off[g], p1 = hm->gene_off[g + 1];
      const double lprior = log10(1.0 / (double)(p1 - p0));
      for (long long p = p0; p < p1; ++p) {
        if (snp_bf) snp_bf[p] = sb[p];
        if (snp_post) snp_post[p] = std::min(1.0, pow(10.0, l1 + lprior + sb[p] - gl[g]));
      }
    }
    if (gene_cfg_post)
      for (int k = 0; k < dim; ++k)
        gene_cfg_post[(size_t)g * dim + k] = std::min(1.0, (1.0 - pi0) * fit->config_prior[k] * pow(10.0, pa[(size_t)k * G + g] - gl[g]));
  }
  (void)grid;
  return 0;
}

int eqb_hm_estep_device_only(eqb_hm_ctx *hm, const double *grid_wts, const double *config_prior, int32_t reps, float *ms)
{
  if (!hm || !grid_wts || !config_prior || !ms || reps < 1) return 1;
  int rc = heavy(hm, grid_wts, config_prior, false); // uploads the parameters (and warms up unless cached)
  if (rc) return rc;
  HCK(cudaEventRecord(hm->ev0, hm->stream));
  for (int r = 0; r < reps; ++r) HCK(launch_estep(hm));
  HCK(cudaEventRecord(hm->ev1, hm->stream));
  HCK(cudaEventSynchronize(hm->ev1));
  float t = 0;
  HCK(cudaEventElapsedTime(&t, hm->ev0, hm->ev1));
  *ms = t / (float)reps;
  return 0;
}

} // extern "C"
