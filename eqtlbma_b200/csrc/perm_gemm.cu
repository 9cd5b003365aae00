// perm_gemm.cu -- host side of the batched-GEMM permutation path (kernels in perm_gemm.cuh) and the FP64
// micro-benchmarks / GEMM self-test of the C ABI (eqb_measure_fp64_peaks, eqb_selftest_perm_gemm).
//
// Batching: permutation COLUMNS (column 0 = the identity = the true data, column 1 + p = permutation p) are
// processed in batches of PB <= 512 columns, genes in groups whose product matrix D fits the memory budget; every
// batch is prep -> GEMM -> BF -> merge on the context's stream, no host synchronisation in between.
#include "perm_gemm.h"

#include <cuda.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "../../include/eqtlbma_b200.h"
#include "perm_gemm.cuh"

namespace eqb {

namespace {

template <class T>
struct Buf {
  T *p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t n, cudaStream_t st)
  {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFreeAsync(p, st);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMallocAsync((void **)&p, std::max<size_t>(n, 1) * sizeof(T), st);
    if (e == cudaSuccess) cap = n;
    return e;
  }
  void release(cudaStream_t st)
  {
    if (p) cudaFreeAsync(p, st);
    p = nullptr;
    cap = 0;
  }
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn()
{
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D tensor [rows][ldn] of doubles, box = 128 rows x 16 doubles (one 128-byte swizzle row per tile row)
bool make_map(CUtensorMap *m, const void *base, unsigned long long rows, int ldn)
{
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)ldn, (cuuint64_t)std::max<unsigned long long>(rows, 1)};
  cuuint64_t strides[1] = {(cuuint64_t)ldn * 8};
  cuuint32_t box[2] = {(cuuint32_t)PG_KC, (cuuint32_t)PG_TM};
  cuuint32_t estr[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int gemm_grid(int n_sm, long long n_tiles) { return (int)std::max<long long>(1, std::min<long long>(n_sm, n_tiles)); }

} // namespace

struct Perm2State {
  Buf<double> Bmat, D, sc_yy, sc_tss, sc_ybar, part_sep;
  Buf<int> sc_n, sc_rankz, ints; // ints: genes | slots | bbase | dbase | task0 | nchunk
  Buf<unsigned int> sc_colvalid;
  Buf<uint8_t> complete;
  Buf<long long> lls; // drow0 | mg
  Buf<unsigned long long> counter; // next BF task (persistent warps)
  Buf<PgTile> tiles;
  Buf<BfTask> tasks;
  Buf<BfPartial> part;
  PgMaps maps;
  const void *map_x_base[PG_MAXX] = {nullptr};
  const void *map_b_base = nullptr;
  size_t map_b_rows = 0;
  bool attrs_set = false;
  bool timing = false;
  Perm2Timing last;
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
};

Perm2State *perm2_create() { return new Perm2State(); }

void perm2_destroy(Perm2State *st, cudaStream_t s)
{
  if (!st) return;
  st->Bmat.release(s);
  st->D.release(s);
  st->sc_yy.release(s);
  st->sc_tss.release(s);
  st->sc_ybar.release(s);
  st->part_sep.release(s);
  st->sc_n.release(s);
  st->sc_rankz.release(s);
  st->ints.release(s);
  st->sc_colvalid.release(s);
  st->complete.release(s);
  st->lls.release(s);
  st->counter.release(s);
  st->tiles.release(s);
  st->tasks.release(s);
  st->part.release(s);
  for (auto &e : st->ev)
    if (e) cudaEventDestroy(e);
  delete st;
}

const Perm2Timing &perm2_last_timing(const Perm2State *st) { return st->last; }
void perm2_set_timing(Perm2State *st, bool on) { st->timing = on; }

static int prep_warps(const Perm2Env &env)
{
  const size_t per = prep_warp_doubles(env.hp->Qmax, env.hp->ldn) * sizeof(double);
  return (int)std::min<size_t>(8, (200 * 1024) / std::max<size_t>(per, 1));
}

static int bf_warps(const Perm2Env &env, int which, int stat_kind)
{
  const size_t per = bf_warp_doubles(env.hp->S, which, stat_kind) * sizeof(double);
  return (int)std::min<size_t>(8, (200 * 1024) / std::max<size_t>(per, 1));
}

// gen row of gridL grouped by unique phi2 + gridS + BMA weights, by value (kernel parameter)
static bool build_grid(const Perm2Env &env, PermGrid &pg)
{
  memset(&pg, 0, sizeof(pg));
  const std::vector<double> &pL = *env.phi2L, &oL = *env.oma2L, &pS = *env.phi2S, &oS = *env.oma2S;
  const int L = (int)pL.size(), K = (int)pS.size(), S = env.hp->S;
  if (L > PGR_L || K > PGR_K || L > 255) return false;
  std::vector<double> uphi;
  for (int k = 0; k < L; ++k)
    if (std::find(uphi.begin(), uphi.end(), pL[k]) == uphi.end()) uphi.push_back(pL[k]);
  if ((int)uphi.size() > PGR_U) return false;
  int n = 0;
  for (size_t u = 0; u < uphi.size(); ++u) {
    pg.uphi[u] = uphi[u];
    pg.ustart[u] = (unsigned char)n;
    for (int k = 0; k < L; ++k)
      if (pL[k] == uphi[u]) {
        pg.omaL[n] = oL[k];
        pg.kL[n] = (unsigned char)k;
        ++n;
      }
  }
  pg.ustart[uphi.size()] = (unsigned char)n;
  // gridS grouped by phi2 + oma2 (the singleton configurations share one logarithm per group)
  std::vector<double> utot;
  for (int k = 0; k < K; ++k)
    if (std::find(utot.begin(), utot.end(), pS[k] + oS[k]) == utot.end()) utot.push_back(pS[k] + oS[k]);
  n = 0;
  for (size_t u = 0; u < utot.size(); ++u) {
    pg.utot[u] = utot[u];
    pg.tstart[u] = (unsigned char)n;
    for (int k = 0; k < K; ++k)
      if (pS[k] + oS[k] == utot[u]) {
        pg.phiS[n] = pS[k];
        pg.omaS[n] = oS[k];
        pg.phiH[n] = 0.5 * pS[k];
        pg.omaH[n] = 0.5 * oS[k];
        pg.kS[n] = (unsigned char)k;
        ++n;
      }
  }
  pg.tstart[utot.size()] = (unsigned char)n;
  pg.UT = (int)utot.size();
  for (int k = 0; k <= S && k < PGR_S; ++k) pg.size_weight[k] = env.hp->size_weight[k];
  pg.size_weight[0] = 0.0;
  pg.UG = (int)uphi.size();
  pg.invL = L > 0 ? 1.0 / (double)L : 0.0;
  pg.invK = K > 0 ? 1.0 / (double)K : 0.0;
  pg.L = L;
  pg.K = K;
  return true;
}

bool perm2_supported(const Perm2Env &env, int which, int stat_kind)
{
  if (!env.d_fp || !env.hfp || !env.hp) return false;
  const DevParams &hp = *env.hp;
  if (hp.qnorm) return false; // per-(gene, permutation) rank transform: general kernel
  if (hp.analysis == EQB_ANALYSIS_JOIN && hp.error_model != EQB_ERROR_UVLR) return false;
  if (env.n_xvar > PG_MAXX || env.n_xvar < 1) return false;
  if (which == 3 && hp.S >= PGR_S) return false;
  if (prep_warps(env) < 1 || bf_warps(env, which, stat_kind) < 1) return false;
  if (hp.analysis == EQB_ANALYSIS_JOIN) {
    PermGrid pg;
    if (!build_grid(env, pg)) return false;
  }
  if (!encode_fn()) return false;
  return true;
}

#define P2CK(call)                                                                      \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess) {                                                            \
      if (err) *err = std::string("perm2: " #call ": ") + cudaGetErrorString(e_);       \
      return 100;                                                                       \
    }                                                                                   \
  } while (0)

int perm2_eval(Perm2State *st, const Perm2Env &env, const int *genes, const int *tabs, size_t n_items_all,
               const unsigned short *d_perm, long long P, int which, int stat_kind, double *out_true, double *out_stat,
               long long *launches, std::string *err)
{
  const DevParams &hp = *env.hp;
  const int S = hp.S, ldn = hp.ldn;
  const bool join = hp.analysis == EQB_ANALYSIS_JOIN;
  cudaStream_t sm = env.stream;
  PermGrid pg;
  memset(&pg, 0, sizeof(pg));
  if (join && !build_grid(env, pg)) {
    if (err) *err = "perm2: grid too large";
    return 1;
  }
  if (!st->attrs_set) {
    P2CK(cudaFuncSetAttribute(perm_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PG_SMEM_BYTES));
    P2CK(cudaFuncSetAttribute(perm_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    P2CK(cudaFuncSetAttribute(perm_bf_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    P2CK(cudaFuncSetAttribute(perm_bf_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (auto &e : st->ev) P2CK(cudaEventCreate(&e));
    st->attrs_set = true;
  }
  st->last = Perm2Timing();
  // genotype tensor maps (fixed per context)
  for (int v = 0; v < env.n_xvar; ++v)
    if (st->map_x_base[v] != env.d_X[v]) {
      if (!make_map(&st->maps.x[v], env.d_X[v], (unsigned long long)hp.M, ldn)) {
        if (err) *err = "perm2: cuTensorMapEncodeTiled failed (genotypes)";
        return 1;
      }
      st->map_x_base[v] = env.d_X[v];
    }
  const long long cols_total = P + 1;
  const int pw = prep_warps(env), bw = bf_warps(env, which, stat_kind);
  const size_t d_budget = std::max<size_t>((size_t)256 << 20, std::min<size_t>(env.free_bytes / 6, (size_t)6 << 30));
  const size_t b_budget = std::max<size_t>((size_t)128 << 20, std::min<size_t>(env.free_bytes / 12, (size_t)2 << 30));

  // per-item block counts
  std::vector<int> nblkB(n_items_all), nblkD(n_items_all);
  std::vector<uint8_t> complete_all(n_items_all * (size_t)S);
  for (size_t i = 0; i < n_items_all; ++i) {
    int nb = 0, nd = 0;
    for (int s = 0; s < S; ++s) {
      const bool comp = env.sub_complete[s] && env.cell_generic[s] && env.cell_generic[s][genes[i]];
      complete_all[i * S + s] = comp ? 1 : 0;
      nb += comp ? 1 : hp.sub[s].Q + 2;
      nd += comp ? 1 : hp.sub[s].Q + 3;
    }
    nblkB[i] = nb;
    nblkD[i] = nd;
  }

  size_t i0 = 0;
  while (i0 < n_items_all) {
    // ---- items of this group and the column batch size: D and B must fit their budgets
    int PB = (int)std::min<long long>(cols_total, 512);
    auto padded = [](int pbv) { return (pbv + PG_TN - 1) / PG_TN * PG_TN; };
    size_t i1 = i0;
    long long rows = 0;
    int maxD = 0;
    size_t browsum = 0;
    {
      // shrink the column batch until the first item fits
      const long long m_first = env.ce[genes[i0]] - env.cb[genes[i0]];
      while (PB > PG_TN && ((size_t)m_first * nblkD[i0] * padded(PB) * 8 > d_budget ||
                            (size_t)nblkB[i0] * padded(PB) * ldn * 8 > b_budget))
        PB = std::max(PG_TN, PB / 2);
    }
    const int PBpad = padded(PB);
    while (i1 < n_items_all) {
      const long long mg = env.ce[genes[i1]] - env.cb[genes[i1]];
      const int md = std::max(maxD, nblkD[i1]);
      if (i1 > i0 && ((size_t)(rows + mg) * md * PBpad * 8 > d_budget || (browsum + nblkB[i1]) * (size_t)PBpad * ldn * 8 > b_budget ||
                      (i1 - i0) >= 65536))
        break;
      rows += mg;
      maxD = md;
      browsum += nblkB[i1];
      ++i1;
    }
    const int n_items = (int)(i1 - i0);
    const long long ldd = (long long)maxD * PBpad;

    // ---- per-group tables: genes | slots | bbase | dbase, drow0 | mg, complete
    std::vector<int> h_int((size_t)n_items * 2 + (size_t)n_items * S * 2 + (size_t)n_items * 2);
    int *h_genes = h_int.data(), *h_slots = h_genes + n_items, *h_bbase = h_slots + n_items, *h_dbase = h_bbase + (size_t)n_items * S,
        *h_task0 = h_dbase + (size_t)n_items * S, *h_nchunk = h_task0 + n_items;
    std::vector<long long> h_ll((size_t)n_items * 2);
    long long *h_drow0 = h_ll.data(), *h_mg = h_drow0 + n_items;
    {
      long long r = 0;
      int bacc = 0;
      for (int il = 0; il < n_items; ++il) {
        const int g = genes[i0 + il];
        h_genes[il] = g;
        h_slots[il] = tabs[i0 + il];
        h_drow0[il] = r;
        h_mg[il] = env.ce[g] - env.cb[g];
        r += h_mg[il];
        int dacc = 0;
        for (int s = 0; s < S; ++s) {
          const bool comp = complete_all[(i0 + il) * S + s] != 0;
          h_bbase[(size_t)il * S + s] = bacc;
          h_dbase[(size_t)il * S + s] = dacc;
          bacc += comp ? 1 : hp.sub[s].Q + 2;
          dacc += comp ? 1 : hp.sub[s].Q + 3;
        }
      }
    }
    // ---- GEMM tiles: m-tile outermost so that concurrent CTAs share the genotype tile
    std::vector<PgTile> h_tiles;
    double useful_rows_cols = 0.0;
    for (int il = 0; il < n_items; ++il) {
      const int g = h_genes[il];
      const long long mb = env.cb[g], mg = h_mg[il];
      for (long long mt = 0; mt < mg; mt += PG_TM) {
        const int mrows = (int)std::min<long long>(PG_TM, mg - mt);
        for (int s = 0; s < S; ++s) {
          const bool comp = complete_all[(i0 + il) * S + s] != 0;
          const int nb = comp ? 1 : hp.sub[s].Q + 3;
          for (int jj = 0; jj < nb; ++jj) {
            const bool sq = !comp && jj == hp.sub[s].Q + 2;
            for (int cb = 0; cb < PBpad; cb += PG_TN) {
              PgTile t;
              t.m0 = (int)(mb + mt);
              t.mrows = mrows;
              t.drow0 = h_drow0[il] + mt;
              t.brow0 = (h_bbase[(size_t)il * S + s] + (sq ? 0 : jj)) * PBpad + cb; // squares contract with the mask row
              t.dcol0 = (h_dbase[(size_t)il * S + s] + jj) * PBpad + cb;
              t.xvar = (short)env.sub_xvar[s];
              t.square = sq ? 1 : 0;
              h_tiles.push_back(t);
              useful_rows_cols += (double)mrows * std::min(PG_TN, PB - std::min(PB, cb));
            }
          }
        }
      }
    }
    // ---- BF tasks: (item, 32-column group, SNP chunk); chunk length sized for ~32 warps per SM
    std::vector<BfTask> h_tasks;
    {
      const int ncg = (PB + 31) / 32;
      const double units = (double)rows * ncg;
      // fixed chunk lengths: the merge order of the chunk partials (hence the last bits of a statistic) must not
      // depend on which other genes share the batch (results independent of the sharding)
      (void)units;
      const int chunk = (join && which == 3) ? 4 : 16;
      for (int il = 0; il < n_items; ++il) {
        const int g = h_genes[il];
        const long long mb = env.cb[g], me = env.ce[g];
        const int nch = (int)std::max<long long>(1, (me - mb + chunk - 1) / chunk);
        h_task0[il] = (int)h_tasks.size();
        h_nchunk[il] = nch;
        for (int cg = 0; cg < ncg; ++cg)
          for (int ch = 0; ch < nch; ++ch) {
            BfTask t;
            t.il = il;
            t.c_lo = cg * 32;
            t.m_begin = mb + (long long)ch * chunk;
            t.m_end = std::min<long long>(me, t.m_begin + chunk);
            t.first = ch == 0;
            t.pad = 0;
            h_tasks.push_back(t);
          }
      }
    }
    const size_t n_tasks = h_tasks.size();
    // ---- device buffers
    const size_t sc_count = (size_t)n_items * S * PBpad;
    P2CK(st->ints.ensure(h_int.size(), sm));
    P2CK(st->lls.ensure(h_ll.size(), sm));
    P2CK(st->complete.ensure((size_t)n_items * S, sm));
    P2CK(st->tiles.ensure(h_tiles.size(), sm));
    P2CK(st->tasks.ensure(n_tasks, sm));
    P2CK(st->part.ensure(n_tasks * 32, sm));
    P2CK(st->counter.ensure(1, sm));
    if (stat_kind == STAT_SEP_PER) P2CK(st->part_sep.ensure(n_tasks * 2 * S * 32, sm));
    P2CK(st->sc_n.ensure(sc_count, sm));
    P2CK(st->sc_rankz.ensure(sc_count, sm));
    P2CK(st->sc_colvalid.ensure(sc_count, sm));
    P2CK(st->sc_yy.ensure(sc_count, sm));
    P2CK(st->sc_tss.ensure(sc_count, sm));
    P2CK(st->sc_ybar.ensure(sc_count, sm));
    const size_t b_rows = browsum * (size_t)PBpad;
    if (b_rows * ldn > st->Bmat.cap) {
      P2CK(st->Bmat.ensure(b_rows * ldn, sm));
      P2CK(cudaMemsetAsync(st->Bmat.p, 0, st->Bmat.cap * sizeof(double), sm)); // padding columns stay finite
    }
    P2CK(st->D.ensure((size_t)std::max<long long>(rows, 1) * ldd, sm));
    // the previous group's kernels still read the tables: stream-ordered copies from pageable memory are staged by
    // the runtime before returning, so the host vectors may go out of scope
    P2CK(cudaMemcpyAsync(st->ints.p, h_int.data(), h_int.size() * sizeof(int), cudaMemcpyHostToDevice, sm));
    P2CK(cudaMemcpyAsync(st->lls.p, h_ll.data(), h_ll.size() * sizeof(long long), cudaMemcpyHostToDevice, sm));
    P2CK(cudaMemcpyAsync(st->complete.p, &complete_all[i0 * S], (size_t)n_items * S, cudaMemcpyHostToDevice, sm));
    P2CK(cudaMemcpyAsync(st->tiles.p, h_tiles.data(), h_tiles.size() * sizeof(PgTile), cudaMemcpyHostToDevice, sm));
    P2CK(cudaMemcpyAsync(st->tasks.p, h_tasks.data(), n_tasks * sizeof(BfTask), cudaMemcpyHostToDevice, sm));
    P2CK(cudaStreamSynchronize(sm)); // (pageable sources)
    if (st->map_b_base != st->Bmat.p || st->map_b_rows != st->Bmat.cap / ldn) {
      if (!make_map(&st->maps.b, st->Bmat.p, (unsigned long long)(st->Bmat.cap / ldn), ldn)) {
        if (err) *err = "perm2: cuTensorMapEncodeTiled failed (operand matrix)";
        return 1;
      }
      st->map_b_base = st->Bmat.p;
      st->map_b_rows = st->Bmat.cap / ldn;
    }
    PermBatch pb;
    memset(&pb, 0, sizeof(pb));
    pb.genes = st->ints.p;
    pb.slots = st->ints.p + n_items;
    pb.bbase = st->ints.p + 2 * (size_t)n_items;
    pb.dbase = pb.bbase + (size_t)n_items * S;
    pb.complete = st->complete.p;
    pb.drow0 = st->lls.p;
    pb.n_items = n_items;
    pb.PBpad = PBpad;
    pb.P_total = P;
    pb.perm_tab = d_perm;
    pb.Bmat = st->Bmat.p;
    pb.sc_n = st->sc_n.p;
    pb.sc_rankz = st->sc_rankz.p;
    pb.sc_colvalid = st->sc_colvalid.p;
    pb.sc_yy = st->sc_yy.p;
    pb.sc_tss = st->sc_tss.p;
    pb.sc_ybar = st->sc_ybar.p;
    pb.D = st->D.p;
    pb.ldd = ldd;
    pb.which = join ? which : 1;
    pb.stat_kind = stat_kind;
    pb.err_flag = env.d_err;
    MergeArgs ma;
    ma.task0 = pb.dbase + (size_t)n_items * S;
    ma.nchunk = ma.task0 + n_items;
    ma.mg = st->lls.p + n_items;
    ma.part = st->part.p;
    ma.part_sep = st->part_sep.p;
    ma.out_true = out_true;
    ma.out_stat = out_stat;
    ma.row0 = (long long)i0;

    for (long long c0 = 0; c0 < cols_total; c0 += PB) {
      pb.c0 = c0;
      pb.PB = (int)std::min<long long>(PB, cols_total - c0);
      const int pbpad_now = padded(pb.PB); // fewer column tiles for the last, shorter batch
      const long long prep_tasks = (long long)n_items * pb.PB * S;
      const bool tm = st->timing;
      if (tm) P2CK(cudaEventRecord(st->ev[0], sm));
      perm_prep_kernel<<<(unsigned)((prep_tasks + pw - 1) / pw), pw * 32, pw * prep_warp_doubles(hp.Qmax, ldn) * sizeof(double), sm>>>(
          env.d_prm, env.d_fp, pb, pw);
      if (tm) P2CK(cudaEventRecord(st->ev[1], sm));
      // tiles of column blocks past the batch's last column are skipped (the list is ordered ... cb fastest)
      {
        const PgTile *tl = st->tiles.p;
        long long nt = (long long)h_tiles.size();
        if (pbpad_now < PBpad) {
          // rebuild the (shorter) list for the last batch
          std::vector<PgTile> h2;
          h2.reserve(h_tiles.size());
          for (const PgTile &t : h_tiles)
            if ((t.dcol0 % PBpad) < pbpad_now) h2.push_back(t);
          P2CK(cudaMemcpyAsync(st->tiles.p, h2.data(), h2.size() * sizeof(PgTile), cudaMemcpyHostToDevice, sm));
          P2CK(cudaStreamSynchronize(sm));
          nt = (long long)h2.size();
          useful_rows_cols = 0.0;
          for (const PgTile &t : h2) useful_rows_cols += (double)t.mrows * std::min(PG_TN, pb.PB - std::min(pb.PB, (t.dcol0 % PBpad)));
        }
        perm_gemm_kernel<<<gemm_grid(env.n_sm, nt), PG_THREADS, PG_SMEM_BYTES, sm>>>(st->maps, tl, (int)nt, ldn / PG_KC, st->D.p, ldd);
        st->last.gemm_flops += 2.0 * PG_TM * PG_TN * (double)ldn * (double)nt;
        st->last.gemm_useful_flops += 2.0 * useful_rows_cols * (double)ldn;
      }
      if (tm) P2CK(cudaEventRecord(st->ev[2], sm));
      {
        // tasks of column groups past the batch's last column do nothing useful but are cheap to skip in-kernel:
        // restrict the grid instead (tasks are ordered item, column group, chunk -> not contiguous): keep them all
        const size_t smem = (size_t)bw * bf_warp_doubles(S, pb.which, stat_kind) * sizeof(double);
        int occ = 1;
        if (pb.which == 3) P2CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, perm_bf_kernel<true>, bw * 32, smem));
        else P2CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, perm_bf_kernel<false>, bw * 32, smem));
        const unsigned grid = (unsigned)std::max<size_t>(1, std::min<size_t>((n_tasks + bw - 1) / bw, (size_t)env.n_sm * std::max(occ, 1)));
        P2CK(cudaMemsetAsync(st->counter.p, 0, sizeof(unsigned long long), sm));
        if (pb.which == 3)
          perm_bf_kernel<true><<<grid, bw * 32, smem, sm>>>(env.d_prm, env.d_fp, pb, pg, st->tasks.p, (long long)n_tasks, bw,
                                                            st->counter.p, st->part.p, st->part_sep.p);
        else
          perm_bf_kernel<false><<<grid, bw * 32, smem, sm>>>(env.d_prm, env.d_fp, pb, pg, st->tasks.p, (long long)n_tasks, bw,
                                                             st->counter.p, st->part.p, st->part_sep.p);
        st->last.bf_items += rows * pb.PB;
      }
      if (tm) P2CK(cudaEventRecord(st->ev[3], sm));
      {
        const long long n = (long long)n_items * pb.PB;
        perm_merge_kernel<<<(unsigned)((n + 127) / 128), 128, 0, sm>>>(pb, ma, S);
      }
      if (launches) *launches += 4;
      P2CK(cudaGetLastError());
      if (tm) {
        P2CK(cudaEventRecord(st->ev[4], sm));
        P2CK(cudaEventSynchronize(st->ev[4]));
        float a = 0, b = 0, c = 0, d = 0;
        cudaEventElapsedTime(&a, st->ev[0], st->ev[1]);
        cudaEventElapsedTime(&b, st->ev[1], st->ev[2]);
        cudaEventElapsedTime(&c, st->ev[2], st->ev[3]);
        cudaEventElapsedTime(&d, st->ev[3], st->ev[4]);
        st->last.prep_ms += a;
        st->last.gemm_ms += b;
        st->last.bf_ms += c;
        st->last.merge_ms += d;
      }
    }
    i0 = i1;
  }
  return 0;
}

} // namespace eqb

// ---------------------------------------------------------------- C ABI diagnostics
extern "C" {

// FP64 pipe peaks of the device, measured with register-resident loops (CUDA events, best of 5):
// out4 = { DFMA TFLOP/s, DMMA (mma.sync.m8n8k4.f64) TFLOP/s, SM clock estimate during the DFMA loop in MHz, SM count }
int eqb_measure_fp64_peaks(int32_t device, double *out4)
{
  using namespace eqb;
  if (!out4) return 1;
  if (cudaSetDevice(device) != cudaSuccess) return 2;
  int n_sm = 0;
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device);
  double *d = nullptr;
  if (cudaMalloc((void **)&d, 64) != cudaSuccess) return 3;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 4000, ctas = n_sm * 8;
  double best_fma = 0.0, best_mma = 0.0;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0);
    fp64_dfma_peak_kernel<<<ctas, 256>>>(d, iters, 1.0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fl = 2.0 * 64.0 * iters * 256.0 * ctas; // 64 FMAs per thread per iteration
    if (rep > 0) best_fma = std::max(best_fma, fl / (ms * 1e-3) / 1e12);
    cudaEventRecord(e0);
    fp64_dmma_peak_kernel<<<ctas, 256>>>(d, iters, 1.0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    const double fl2 = 512.0 * 64.0 * iters * 8.0 * ctas; // 64 DMMAs (512 flop each) per warp per iteration
    if (rep > 0) best_mma = std::max(best_mma, fl2 / (ms * 1e-3) / 1e12);
  }
  const cudaError_t e = cudaGetLastError();
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  out4[0] = best_fma;
  out4[1] = best_mma;
  out4[2] = best_fma * 1e12 / (2.0 * 64.0 * n_sm) / 1e6; // clock at which 64 DFMA / clk / SM gives that rate
  out4[3] = (double)n_sm;
  return e == cudaSuccess ? 0 : 4;
}

// worst deviation of the permutation BF kernel's table-driven elementary functions from the CUDA library versions over
// n pseudo-random arguments: out5 = { rcp rel, log abs, rsqrt rel, exp / exp10 rel, special-value mismatches }
int eqb_math_selftest(int32_t device, int64_t n, double *out5)
{
  if (!out5 || n <= 0) return 1;
  if (cudaSetDevice(device) != cudaSuccess) return 2;
  double *d = nullptr;
  if (cudaMalloc((void **)&d, 5 * sizeof(double)) != cudaSuccess) return 3;
  cudaMemset(d, 0, 5 * sizeof(double));
  eqb::math_selftest_kernel<<<296, 256>>>(n, d);
  cudaError_t e = cudaMemcpy(out5, d, 5 * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(d);
  return e == cudaSuccess ? 0 : 4;
}

// Self-test and throughput of perm_gemm_kernel on pseudo-random operands: D = X . B^T (and the squared variant) for
// n_rows genotype rows x n_cols operand rows of length ldn (a multiple of 16); out3 = { worst |D - reference| / sum
// |terms|, TFLOP/s of the GEMM kernel (CUDA events, best of 3), tiles }.
int eqb_selftest_perm_gemm(int32_t device, int64_t n_rows, int64_t n_cols, int32_t ldn, double *out3)
{
  using namespace eqb;
  if (!out3 || n_rows < 1 || n_cols < 1 || ldn < 16 || ldn % 16 != 0) return 1;
  if (cudaSetDevice(device) != cudaSuccess) return 2;
  int n_sm = 0;
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device);
  const long long colpad = (n_cols + PG_TN - 1) / PG_TN * PG_TN;
  std::vector<double> hx((size_t)n_rows * ldn), hb((size_t)colpad * ldn, 0.0);
  unsigned long long s = 0x9E3779B97F4A7C15ull;
  auto unif = [&]() {
    s ^= s >> 12;
    s ^= s << 25;
    s ^= s >> 27;
    return (double)((s * 0x2545F4914F6CDD1Dull) >> 11) * (1.0 / 9007199254740992.0);
  };
  for (auto &v : hx) v = floor(unif() * 3.0) + ((unif() < 0.1) ? unif() : 0.0);
  for (long long c = 0; c < n_cols; ++c)
    for (int k = 0; k < ldn; ++k) hb[(size_t)c * ldn + k] = unif() - 0.5;
  std::vector<PgTile> tiles;
  for (int sq = 0; sq < 2; ++sq)
    for (long long m0 = 0; m0 < n_rows; m0 += PG_TM)
      for (long long c0 = 0; c0 < colpad; c0 += PG_TN) {
        PgTile t;
        t.m0 = (int)m0;
        t.mrows = (int)std::min<long long>(PG_TM, n_rows - m0);
        t.drow0 = m0;
        t.brow0 = (int)c0;
        t.dcol0 = (int)(sq * colpad + c0);
        t.xvar = 0;
        t.square = (short)sq;
        tiles.push_back(t);
      }
  const long long ldd = 2 * colpad;
  double *dX = nullptr, *dB = nullptr, *dD = nullptr, *dW = nullptr;
  PgTile *dT = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  int rc = 0;
  do {
    PgMaps maps;
    memset(&maps, 0, sizeof(maps));
    if (cudaMalloc((void **)&dX, hx.size() * 8) != cudaSuccess || cudaMalloc((void **)&dB, hb.size() * 8) != cudaSuccess ||
        cudaMalloc((void **)&dD, (size_t)n_rows * ldd * 8) != cudaSuccess || cudaMalloc((void **)&dW, 8) != cudaSuccess ||
        cudaMalloc((void **)&dT, tiles.size() * sizeof(PgTile)) != cudaSuccess) {
      rc = 3;
      break;
    }
    cudaMemcpy(dX, hx.data(), hx.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hb.data(), hb.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dT, tiles.data(), tiles.size() * sizeof(PgTile), cudaMemcpyHostToDevice);
    cudaMemset(dW, 0, 8);
    cudaMemset(dD, 0xff, (size_t)n_rows * ldd * 8);
    if (!make_map(&maps.x[0], dX, (unsigned long long)n_rows, ldn) || !make_map(&maps.b, dB, (unsigned long long)colpad, ldn)) {
      rc = 4;
      break;
    }
    if (cudaFuncSetAttribute(perm_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PG_SMEM_BYTES) != cudaSuccess) {
      rc = 5;
      break;
    }
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4 && rc == 0; ++rep) {
      cudaEventRecord(e0);
      perm_gemm_kernel<<<gemm_grid(n_sm, (long long)tiles.size()), PG_THREADS, PG_SMEM_BYTES>>>(maps, dT, (int)tiles.size(), ldn / PG_KC,
                                                                                                dD, ldd);
      cudaEventRecord(e1);
      if (cudaEventSynchronize(e1) != cudaSuccess) rc = 6;
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0) best = std::min(best, ms);
    }
    if (rc) break;
    perm_gemm_check_kernel<<<(unsigned)tiles.size(), 256>>>(dX, dB, dT, (int)tiles.size(), ldn, dD, ldd, dW);
    double worst = 0.0;
    if (cudaMemcpy(&worst, dW, 8, cudaMemcpyDeviceToHost) != cudaSuccess) {
      rc = 7;
      break;
    }
    out3[0] = worst;
    out3[1] = 2.0 * PG_TM * PG_TN * (double)ldn * (double)tiles.size() / (best * 1e-3) / 1e12;
    out3[2] = (double)tiles.size();
  } while (0);
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  cudaFree(dX);
  cudaFree(dB);
  cudaFree(dD);
  cudaFree(dW);
  cudaFree(dT);
  if (rc == 0 && cudaGetLastError() != cudaSuccess) rc = 8;
  return rc;
}

} // extern "C"
