// perm_gemm.cuh -- K4 (round 2): the permutation pass as a batched GEMM + a lane-per-item Bayes-factor kernel.
//
// Reference: Gene::MakePermutationsJoin / ...SepAllSubgroups / ...SepPerSubgroup (gene.cpp:380-717) call
// GeneSnpPair::CalcSstatsOneSbgrp + CalcAbfsUvlr for every (cis SNP, subgroup, permutation).  Here the work of one
// batch of (gene, permutation column) items is split into four kernels:
//
//   perm_prep_kernel   warp per (gene, column, subgroup): permuted phenotype gather, keep-mask, CGS2 basis of
//                      [1, covariates] on the kept rows, residual phenotype -> rows of the operand matrix B
//                      (one row of ldn doubles per (gene, subgroup, j, column); j = 0 mask, 1..Q basis, Q+1 y~)
//                      + per-(gene, subgroup, column) scalars (n, rank, y~'y~, tss, ybar).  gene_snp_pair.cpp:79-170.
//                      Subgroups whose kept rows do not depend on the permutation ("complete": every sample has
//                      genotype, expression and covariates, no NaN in the gene's row) project on the FIXED basis of
//                      K1 and need only the y~ row: x~'x~ is permutation-invariant (K1c output).
//   perm_gemm_kernel   D = X_gene . B^T and (X o X) . mask^T as 128 x 128 tiles: TMA (cp.async.bulk.tensor, 128-byte
//                      swizzle) stages 128 x 16 tiles of X and B in shared memory through a 6-deep mbarrier ring fed
//                      by a producer warp; 8 consumer warps (32 x 64 each) issue FP64 mma.sync.m8n8k4 (DMMA: tcgen05
//                      has no f64 kind and m16n8k* lower to the same DMMA.8x8x4 on sm_100a) with conflict-free
//                      fragment loads.  Permutations are the N dimension (128 columns per tile), persistent CTAs.
//   perm_bf_kernel     lane per (SNP, column): x~'x~ = R2 - sum_k H_k^2 (Gram form, explicit CGS2 when it cancels),
//                      summary statistics + standardisation (stats_from_dots, gene_snp_pair.cpp:175-290), ABFs of
//                      the requested kind (gene_snp_pair.cpp:297-622) and the running statistic over the SNPs of
//                      the warp's chunk (gene.cpp:643-697); every grid point / configuration is warp-uniform.
//                      --pbf all: 2^S-1 configurations in the linear domain, subsets enumerated as (high part: Gray
//                      code in registers) x (8 low subsets: shared-memory table), exp through a 16-entry 2^(j/16)
//                      table + degree-4 polynomial and rsqrt through MUFU + one Newton step (relative error < 1e-10
//                      per term, inside the 1e-8 absolute budget of a log10 BF by two orders of magnitude).
//   perm_merge_kernel  thread per (gene, column): chunk partials in SNP order -> the gene-level statistic with the
//                      reference's NaN rules (gene.cpp:429-433, 575-596, 678-697).
#pragma once

#include <cuda.h> // CUtensorMap

#include "fast_kernels.cuh"
#include "table_math.cuh"

namespace eqb {

constexpr int PG_TM = 128, PG_TN = 128; // CTA tile of D
constexpr int PG_KC = 16;               // doubles per k-chunk = one 128-byte swizzled shared-memory row
constexpr int PG_STAGES = 6;
constexpr int PG_CONSUMERS = 8;         // consumer warps (4 along M x 2 along N), + 1 producer warp
constexpr int PG_THREADS = (PG_CONSUMERS + 1) * 32;
constexpr int PG_STAGE_BYTES = (PG_TM + PG_TN) * PG_KC * 8; // 32 KB
constexpr int PG_SMEM_BYTES = PG_STAGES * PG_STAGE_BYTES + 1024 /* alignment */ + 256 /* barriers */;
constexpr int PG_MAXX = 8;              // genotype variants (tensor maps) per launch

struct PgTile { // one 128 x 128 tile of D
  int m0;          // first genotype row (global SNP index) of the A tile
  int mrows;       // rows of the tile that belong to the gene (<= 128)
  long long drow0; // first row of D
  int brow0;       // first row of B (128 consecutive rows = 128 columns of one (gene, subgroup, j) block)
  int dcol0;       // first column of D
  short xvar;      // tensor map of the genotype variant
  short square;    // A operand squared element-wise (the masked sums of squares)
};

struct PgMaps {
  CUtensorMap b;
  CUtensorMap x[PG_MAXX];
};

// ---------------------------------------------------------------- PTX helpers (mbarrier, TMA)
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1)
{
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(map), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}

// ---------------------------------------------------------------- the GEMM
// D[drow0 + r][dcol0 + c] = sum_k A[m0 + r][k] * B[brow0 + c][k]   (A = X or X o X; K = ldn = kchunks * 16)
__global__ void __launch_bounds__(PG_THREADS, 1) perm_gemm_kernel(const __grid_constant__ PgMaps maps, const PgTile *__restrict__ tiles,
                                                                  int n_tiles, int kchunks, double *__restrict__ D, long long ldd)
{
  extern __shared__ uint8_t pg_smem_raw[];
  const uint32_t base = (smem_u32(pg_smem_raw) + 1023u) & ~1023u; // 128-byte swizzle atoms are 1024-byte aligned
  const uint32_t bar0 = base + PG_STAGES * PG_STAGE_BYTES;        // full[PG_STAGES], empty[PG_STAGES]
  uint8_t *gen_base = pg_smem_raw + (base - smem_u32(pg_smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < PG_STAGES; ++s) {
      mbar_init(bar0 + 8 * s, 1);                            // full: one arrive.expect_tx by the producer
      mbar_init(bar0 + 8 * (PG_STAGES + s), PG_CONSUMERS);   // empty: one arrive per consumer warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (warp == PG_CONSUMERS) {
    // ---------------- producer: one lane streams the A / B k-chunks of this CTA's tiles through the ring
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const PgTile tl = tiles[t];
        const CUtensorMap *mx = &maps.x[tl.xvar];
        for (int kc = 0; kc < kchunks; ++kc) {
          mbar_wait(bar0 + 8 * (PG_STAGES + stage), phase ^ 1u);
          const uint32_t full = bar0 + 8 * stage;
          const uint32_t sa = base + stage * PG_STAGE_BYTES, sb = sa + PG_TM * PG_KC * 8;
          mbar_expect_tx(full, PG_STAGE_BYTES);
          tma_load_2d(sa, mx, full, kc * PG_KC, tl.m0);
          tma_load_2d(sb, &maps.b, full, kc * PG_KC, tl.brow0);
          if (++stage == PG_STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
    return;
  }
  // ---------------- consumers: warp (wm, wn) owns rows [32 wm, +32) x columns [64 wn, +64) of the tile
  const int g = lane >> 2, kk = lane & 3;
  const int wm = warp & 3, wn = warp >> 2;
  // byte offset of element (row r, k) inside a tile: r * 128 + (((k >> 1) ^ (r & 7)) << 4) + (k & 1) * 8; r & 7 == g here
  int koff[4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) koff[ks] = ((((ks * 2 + (kk >> 1)) ^ g) << 4) | ((kk & 1) << 3));
  const int arow = (wm * 32 + g) * 128, brow = (PG_TM * PG_KC * 8) + (wn * 64 + g) * 128;
  int stage = 0;
  uint32_t phase = 0;
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const PgTile tl = tiles[t];
    const bool sq = tl.square != 0;
    double acc[4][8][2];
#pragma unroll
    for (int mb = 0; mb < 4; ++mb)
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;
    for (int kc = 0; kc < kchunks; ++kc) {
      mbar_wait(bar0 + 8 * stage, phase);
      const uint8_t *st = gen_base + stage * PG_STAGE_BYTES;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        double a[4], b[8];
#pragma unroll
        for (int mb = 0; mb < 4; ++mb) {
          a[mb] = *reinterpret_cast<const double *>(st + arow + mb * 1024 + koff[ks]);
          if (sq) a[mb] *= a[mb];
        }
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) b[nb] = *reinterpret_cast<const double *>(st + brow + nb * 1024 + koff[ks]);
#pragma unroll
        for (int mb = 0; mb < 4; ++mb)
#pragma unroll
          for (int nb = 0; nb < 8; ++nb) dmma_m8n8k4(acc[mb][nb][0], acc[mb][nb][1], a[mb], b[nb]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar0 + 8 * (PG_STAGES + stage));
      if (++stage == PG_STAGES) {
        stage = 0;
        phase ^= 1u;
      }
    }
#pragma unroll
    for (int mb = 0; mb < 4; ++mb) {
      const int r = wm * 32 + mb * 8 + g;
      if (r < tl.mrows) {
        double *dr = D + (size_t)(tl.drow0 + r) * ldd + tl.dcol0 + wn * 64 + 2 * kk;
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) *reinterpret_cast<double2 *>(dr + nb * 8) = make_double2(acc[mb][nb][0], acc[mb][nb][1]);
      }
    }
  }
}

// plain reference of the same product (self-test only)
__global__ void perm_gemm_check_kernel(const double *__restrict__ X, const double *__restrict__ B, const PgTile *__restrict__ tiles,
                                       int n_tiles, int ldn, const double *__restrict__ D, long long ldd, double *__restrict__ worst)
{
  const int t = blockIdx.x;
  if (t >= n_tiles) return;
  const PgTile tl = tiles[t];
  double w = 0.0;
  for (int e = threadIdx.x; e < tl.mrows * PG_TN; e += blockDim.x) {
    const int r = e / PG_TN, c = e % PG_TN;
    const double *x = X + (size_t)(tl.m0 + r) * ldn, *b = B + (size_t)(tl.brow0 + c) * ldn;
    double s = 0.0, sa = 0.0;
    for (int k = 0; k < ldn; ++k) {
      const double a = tl.square ? x[k] * x[k] : x[k];
      s = fma(a, b[k], s);
      sa += fabs(a * b[k]);
    }
    const double d = D[(size_t)(tl.drow0 + r) * ldd + tl.dcol0 + c];
    const double err = fabs(d - s) / fmax(sa, 1e-300); // relative to the sum of magnitudes (summation order differs)
    w = fmax(w, (err == err) ? err : 1.0);
  }
  w = warp_max_nonan(w);
  if ((threadIdx.x & 31) == 0) atomicMax((unsigned long long *)worst, (unsigned long long)__double_as_longlong(w));
}

// ---------------------------------------------------------------- FP64 pipe peaks (diagnostic micro-benchmarks)
// DFMA: 8 independent chains per thread; DMMA: 16 independent accumulator tiles per warp.  Register-resident,
// no memory traffic: the numbers are the denominators of the FP64 rooflines (eqb_measure_fp64_peaks).
__global__ void __launch_bounds__(256) fp64_dfma_peak_kernel(double *out, int iters, double seed)
{
  double a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed + i + threadIdx.x * 1e-3;
  const double m = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fma(a[i], m, c);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 12345.678) out[0] = s; // never true: keeps the chains alive
}

__global__ void __launch_bounds__(256) fp64_dmma_peak_kernel(double *out, int iters, double seed)
{
  double c[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = 0.0;
  const double a = seed + threadIdx.x * 1e-3, b = 1e-3 * seed;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < 16; ++i) dmma_m8n8k4(c[i][0], c[i][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
}

// ---------------------------------------------------------------- batch description shared by prep / bf / merge
constexpr int PGR_U = 32, PGR_L = 64, PGR_K = 32, PGR_S = 24;
struct PermGrid { // kernel parameter (constant bank): the "gen" row of gridL grouped by unique phi2, gridS, BMA weights
  double uphi[PGR_U];
  double omaL[PGR_L];       // oma2 of gen-row entry i (grouped order)
  double phiS[PGR_K], omaS[PGR_K];
  double size_weight[PGR_S]; // (1/S)(1/choose(S, size)); [0] = 0 (the empty configuration is not part of the model)
  double utot[PGR_K];       // unique values of phi2 + oma2 on gridS; phiS / omaS / kS are stored grouped by them
  unsigned char ustart[PGR_U + 1];
  unsigned char kL[PGR_L];  // grid point k of entry i (element 0 carries the NaN rule of log10_weighted_sum)
  unsigned char tstart[PGR_K + 1];
  unsigned char kS[PGR_K];  // grid point k of gridS entry i
  double phiH[PGR_K], omaH[PGR_K]; // phiS / 2, omaS / 2
  double invL, invK;
  int UG, L, K, UT;
};

struct PermBatch {
  // items (genes) of the batch
  const int *genes;         // [n_items] gene id
  const int *slots;         // [n_items] permutation table of the gene (slot in its write-group / table index)
  const uint8_t *complete;  // [n_items][S] kept rows of (gene, subgroup) do not depend on the permutation
  const long long *drow0;   // [n_items] first row of D of the item
  int n_items;
  // permutation columns: column c of the batch is permutation c0 + c - 1 of the run (-1 = identity = the true data)
  long long c0;
  int PB, PBpad;            // columns of the batch, padded to a multiple of 128
  long long P_total;
  const unsigned short *perm_tab; // [table][P_total][N]
  // operand matrix B and per-(item, subgroup, column) scalars.  Rows of ldn doubles, in blocks of PBpad rows (one row
  // per column): block bbase[item * S + s] + jj with jj = 0 mask, 1..Q basis, Q+1 y~ (general case) or jj = 0 y~
  // ("complete" subgroups: only the residual phenotype depends on the permutation)
  double *Bmat;
  const int *bbase;         // [n_items][S]
  int *sc_n, *sc_rankz;     // [(item * S + s) * PBpad + c]
  unsigned int *sc_colvalid;
  double *sc_yy, *sc_tss, *sc_ybar;
  // D = X . B^T: row (drow0[item] + m - cis_begin), column (dbase[item * S + s] + jj) * PBpad + c with the jj of B,
  // plus jj = Q+2: (X o X) . mask (the masked sum of squares); complete subgroups have the single block jj = 0 (x'y~)
  const double *D;
  const int *dbase;         // [n_items][S]
  long long ldd;
  int which;                // 1 gen, 2 gen-sin, 3 all (join); sep: 1
  int stat_kind;
  int *err_flag;
};

// shared memory (doubles) of one warp of perm_prep_kernel
__host__ __device__ inline size_t prep_warp_doubles(int Qmax, int ldn) { return (size_t)(Qmax + 2) * (ldn + 1); }

// warp task = (item, column, subgroup)
__global__ void __launch_bounds__(256) perm_prep_kernel(const DevParams *__restrict__ prm_, const FastParams *__restrict__ fp_,
                                                        const PermBatch pb, int warps_per_cta)
{
  const DevParams &prm = *prm_;
  extern __shared__ double prep_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = prm.S, N = prm.N, ldn = prm.ldn, Qmax = prm.Qmax, lds = ldn + 1;
  const long long task = (long long)blockIdx.x * warps_per_cta + warp;
  if (warp >= warps_per_cta || task >= (long long)pb.n_items * pb.PB * S) return;
  const int s = (int)(task % S);
  const int c = (int)((task / S) % pb.PB);
  const int il = (int)(task / ((long long)S * pb.PB));
  const int g = pb.genes[il];
  const long long p = pb.c0 + c - 1;
  const unsigned short *perm = (p >= 0) ? pb.perm_tab + ((size_t)pb.slots[il] * pb.P_total + p) * N : nullptr;
  const SubDev &sb = prm.sub[s];
  const FastSub &fs = fp_->sub[s];
  const size_t sc = ((size_t)il * S + s) * pb.PBpad + c;
  const int bb = pb.bbase[(size_t)il * S + s];
  auto brow = [&](int jj) { return pb.Bmat + ((size_t)(bb + jj) * pb.PBpad + c) * ldn; };
  double *q = prep_smem + (size_t)warp * prep_warp_doubles(Qmax, ldn); // [Q+1][lds] basis rows, then the y row
  double *yt = q + (size_t)(Qmax + 1) * lds;
  const double *Yg = sb.Yall + (size_t)g * ldn;
  const int Q = sb.Q;
  if (pb.complete[(size_t)il * S + s]) {
    // kept rows = every sample, whatever the permutation: fixed basis of K1 (prep_basis_kernel), only y~ changes
    double ysum = 0.0;
    for (int i = lane; i < ldn; i += 32) {
      const double yv = (i < N) ? Yg[perm ? (int)perm[i] : i] : 0.0;
      yt[i] = yv;
      ysum += yv;
    }
    ysum = warp_sum(ysum);
    const int n = fs.n;
    const double ybar = ysum / n;
    double tss = 0.0;
    for (int i = lane; i < N; i += 32) {
      const double d = yt[i] - ybar;
      tss += d * d;
    }
    tss = warp_sum(tss);
    __syncwarp();
    for (int pass = 0; pass < 2; ++pass)
      for (int j = 0; j <= Q; ++j) {
        if (!((fs.colvalid >> j) & 1u)) continue;
        const double *qj = fs.Bs + (size_t)j * ldn;
        double h = 0.0;
        for (int i = lane; i < ldn; i += 32) h += qj[i] * yt[i];
        h = warp_sum(h);
        for (int i = lane; i < ldn; i += 32) yt[i] -= h * qj[i];
        __syncwarp();
      }
    double yy = 0.0;
    double *out = brow(0);
    for (int i = lane; i < ldn; i += 32) {
      const double v = yt[i];
      yy += v * v;
      out[i] = v;
    }
    yy = warp_sum(yy);
    if (lane == 0) {
      pb.sc_n[sc] = n;
      pb.sc_rankz[sc] = fs.rankz;
      pb.sc_colvalid[sc] = fs.colvalid;
      pb.sc_yy[sc] = yy;
      pb.sc_tss[sc] = tss;
      pb.sc_ybar[sc] = ybar;
    }
    return;
  }
  // general case (ragged individuals, NaN expression values, absent genes): masks and basis per (gene, permutation)
  int n = 0;
  if (sb.gene_has[g]) {
    for (int i = lane; i < lds; i += 32) {
      double yv = 0.0;
      bool keep = false;
      if (i < N) {
        yv = Yg[perm ? (int)perm[i] : i];
        keep = sb.gmask[i] && !isnan(yv);
      }
      yt[i] = keep ? yv : 0.0;
      q[i] = keep ? 1.0 : 0.0;
      n += keep ? 1 : 0;
    }
    n = warp_sum_int(n);
  }
  __syncwarp();
  if (n == 0) {
    for (int j = 0; j <= Q + 1; ++j) {
      double *out = brow(j);
      for (int i = lane; i < ldn; i += 32) out[i] = 0.0;
    }
    if (lane == 0) {
      pb.sc_n[sc] = 0;
      pb.sc_rankz[sc] = 0;
      pb.sc_colvalid[sc] = 0;
      pb.sc_yy[sc] = 0.0;
      pb.sc_tss[sc] = 0.0;
      pb.sc_ybar[sc] = 0.0;
    }
    return;
  }
  // row 0 stays the 0/1 mask (the intercept direction is mask / sqrt(n); the scaling is applied by the consumer)
  const double inv_sqrt_n = 1.0 / sqrt((double)n);
  unsigned int colvalid = 1u;
  int rankz = 1;
  for (int k = 1; k <= Q; ++k) {
    double *qk = q + (size_t)k * lds;
    const double *Ck = sb.Call + (size_t)(k - 1) * ldn;
    double nrm0 = 0.0;
    int missing = 0;
    for (int i = lane; i < ldn; i += 32) {
      const bool keep = q[i] != 0.0;
      const double v = keep ? Ck[i] : 0.0;
      if (keep && !sb.cmask[i]) missing = 1;
      qk[i] = v;
      nrm0 += v * v;
    }
    nrm0 = warp_sum(nrm0);
    if (__any_sync(0xffffffffu, missing) && lane == 0) atomicExch(pb.err_flag, 1);
    __syncwarp();
    for (int pass = 0; pass < 2; ++pass)
      for (int j = 0; j < k; ++j) {
        if (!((colvalid >> j) & 1u)) continue;
        const double *qj = q + (size_t)j * lds;
        const double sc_j = (j == 0) ? inv_sqrt_n : 1.0;
        double h = 0.0;
        for (int i = lane; i < ldn; i += 32) h += qj[i] * qk[i];
        h = warp_sum(h) * sc_j * sc_j;
        for (int i = lane; i < ldn; i += 32) qk[i] -= h * qj[i];
        __syncwarp();
      }
    double nrm1 = 0.0;
    for (int i = lane; i < ldn; i += 32) nrm1 += qk[i] * qk[i];
    nrm1 = warp_sum(nrm1);
    if (nrm1 > 1e-20 * nrm0 && nrm1 > 0.0) {
      const double inv = 1.0 / sqrt(nrm1);
      for (int i = lane; i < ldn; i += 32) qk[i] *= inv;
      colvalid |= (1u << k);
      rankz++;
    } else
      for (int i = lane; i < ldn; i += 32) qk[i] = 0.0;
    __syncwarp();
  }
  double ysum = 0.0;
  for (int i = lane; i < ldn; i += 32) ysum += yt[i];
  ysum = warp_sum(ysum);
  const double ybar = ysum / n;
  double tss = 0.0;
  for (int i = lane; i < ldn; i += 32)
    if (q[i] != 0.0) {
      const double d = yt[i] - ybar;
      tss += d * d;
    }
  tss = warp_sum(tss);
  for (int pass = 0; pass < 2; ++pass)
    for (int j = 0; j <= Q; ++j) {
      if (!((colvalid >> j) & 1u)) continue;
      const double *qj = q + (size_t)j * lds;
      const double sc_j = (j == 0) ? inv_sqrt_n : 1.0;
      double h = 0.0;
      for (int i = lane; i < ldn; i += 32) h += qj[i] * yt[i];
      h = warp_sum(h) * sc_j * sc_j;
      for (int i = lane; i < ldn; i += 32) yt[i] -= h * qj[i];
      __syncwarp();
    }
  double yy = 0.0;
  for (int i = lane; i < ldn; i += 32) yy += yt[i] * yt[i];
  yy = warp_sum(yy);
  for (int j = 0; j <= Q + 1; ++j) {
    const double *src = (j == Q + 1) ? yt : q + (size_t)j * lds;
    double *out = brow(j);
    for (int i = lane; i < ldn; i += 32) out[i] = src[i];
  }
  if (lane == 0) {
    pb.sc_n[sc] = n;
    pb.sc_rankz[sc] = rankz;
    pb.sc_colvalid[sc] = colvalid;
    pb.sc_yy[sc] = yy;
    pb.sc_tss[sc] = tss;
    pb.sc_ybar[sc] = ybar;
  }
}

// ---------------------------------------------------------------- Bayes factors, lane per (SNP, column)
struct BfTask { // one warp: 32 columns x the SNPs [m_begin, m_end) of one gene
  int il;        // item
  int c_lo;      // first column (multiple of 32)
  long long m_begin, m_end;
  int first;     // the chunk holds the gene's first cis SNP
  int pad;
};

// partial statistic of a (task, column)
struct BfPartial {
  double m, acc;   // running max / online log-sum-exp (join); running minimum p-value (sep, all subgroups) in m
  int cnt_nonnan;  // non-NaN SNP values
  int flags;       // bit 0: the first SNP's value is NaN, bit 1: at least one value accumulated
};

constexpr int PBF_SA = 3; // subgroups of the low part of a configuration (their 2^3 subset sums live in registers)

// shared memory (doubles) of one warp of perm_bf_kernel: st[3S][32] + hi[S-3][3][32] (--pbf all) + sep[2S][32]
__host__ __device__ inline size_t bf_warp_doubles(int S, int which, int stat_kind)
{
  size_t d = (size_t)3 * S * 32;
  if (which == 3) d += (size_t)(S > PBF_SA ? S - PBF_SA : 0) * 3 * 32;
  if (stat_kind == STAT_SEP_PER) d += (size_t)2 * S * 32;
  return d;
}

// diagnostics: worst deviation of the table-driven forms from the CUDA library over n pseudo-random arguments
__global__ void math_selftest_kernel(long long n, double *out)
{
  __shared__ BfTabs Tsm;
  __shared__ double worst[5];
  if (threadIdx.x < 16) {
    const double mj = 1.0 + ((double)threadIdx.x + 0.5) / 16.0;
    Tsm.exp16[threadIdx.x] = exp2((double)threadIdx.x / 16.0);
    Tsm.log16[threadIdx.x] = make_double2(1.0 / mj, log(mj));
  }
  if (threadIdx.x < 5) worst[threadIdx.x] = 0.0;
  __syncthreads();
  TabRef T;
  T.base = smem_u32(&Tsm);
  double w[5] = {0, 0, 0, 0, 0};
  unsigned long long st = 0x9E3779B97F4A7C15ull * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x + 1);
  auto unif = [&]() { // xorshift64*, uniform in [0, 1)
    st ^= st >> 12;
    st ^= st << 25;
    st ^= st >> 27;
    return (double)((st * 0x2545F4914F6CDD1Dull) >> 11) * (1.0 / 9007199254740992.0);
  };
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int kind = (int)(i % 3);
    double x;
    if (kind == 0) x = exp10(-280.0 + 560.0 * unif());
    else if (kind == 1) x = exp10(-12.0 + 24.0 * unif());
    else x = 1.0 + (unif() - 0.5) * ((i % 2) ? 2e-3 : 1.2);
    w[0] = fmax(w[0], fabs(rcp_n(x) - 1.0 / x) * x);
    w[1] = fmax(w[1], fabs(log_tab16(x, T) - log(x)));
    w[1] = fmax(w[1], fabs(log_tab16_pos(x, T) - log(x)));
    w[2] = fmax(w[2], fabs(rsqrt_newton1(x) - rsqrt(x)) * sqrt(x));
    const double y = (kind == 0) ? -690.0 + 1380.0 * unif() : ((kind == 1) ? -60.0 * unif() : 2.0 * unif() - 1.0);
    const double e0 = exp(y);
    w[3] = fmax(w[3], fabs(exp_tab16<true>(y, T) - e0) / e0);
    w[3] = fmax(w[3], fabs(exp_tab16<false>(y, T) - e0) / e0);
    const double d0 = exp10(y * 0.4342944819032518);
    w[3] = fmax(w[3], fabs(exp10_tab16<true>(y * 0.4342944819032518, T) - d0) / d0);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    // special values: the clamped exponential must return (nearly) zero (NaN arguments are excluded by the callers);
    // the logarithm forwards to the library
    double bad = 0.0;
    const double se[4] = {-INFINITY, -1e300, -5000.0, -700.5};
    for (int k = 0; k < 4; ++k) {
      const double v = exp_tab16<true>(se[k], T);
      if (!(v >= 0.0 && v < 1e-300)) bad += 1.0;
    }
    if (!(exp_tab16<false>(-5000.0, T) < 1e-290)) bad += 1.0;
    const double sl[5] = {0.0, -1.0, INFINITY, nan(""), 1e-310};
    for (int k = 0; k < 5; ++k) {
      const double a = log(sl[k]), b = log_tab16(sl[k], T);
      if (!((isnan(a) && isnan(b)) || a == b)) bad += 1.0;
    }
    w[4] = bad;
  }
  for (int k = 0; k < 5; ++k) atomicMax((unsigned long long *)&worst[k], (unsigned long long)__double_as_longlong(w[k]));
  __syncthreads();
  if (threadIdx.x < 5) atomicMax((unsigned long long *)&out[threadIdx.x], (unsigned long long)__double_as_longlong(worst[threadIdx.x]));
}

// Standardised statistics of one (SNP, subgroup, column) in the regular case -- full-rank design, residual sum of
// squares > 0, tabulated t -> z map valid -- with b = sign(x~'y~) |z| / sqrt(x~'x~), v = 1 / x~'x~, t = z
// (gene_snp_pair.cpp:256-290 after substituting se = sigmahat / sqrt(x~'x~): sigmahat cancels).  |t|^2 / nu =
// ess / rss, so w^2 = nu log1p(t^2 / nu) = -nu log(rss / yy).  Returns false for everything else (NaN rules, rank
// deficiency, far tail): the caller then takes stats_from_dots.
__device__ __forceinline__ bool stats_lean(double xy, double xx, double xraw2, double yy, int n, int Q, int rankz,
                                           const double *__restrict__ tz, double tz_nu, double tz_wmax, const TabRef T,
                                           double &b, double &v, double &t)
{
  const double nu = (double)(n - 2 - Q);
  if (tz == nullptr || n < Q + 3 || rankz != Q + 1 || nu != tz_nu || !(xraw2 > 0.0) || !(xx > 1e-24 * xraw2) || !(yy > 0.0))
    return false;
  const double ixx = rcp_n(xx);
  const double ess = xy * xy * ixx;
  const double rss = yy - ess;
  const double larg = rss * rcp_n(yy);
  if (!(larg > 1e-290)) return false; // (also rss <= 0: exact fit, NaN rules of the full function)
  double w2 = -nu * log_tab16_pos(larg, T);
  w2 = (w2 > 0.0) ? w2 : 0.0;
  const double w = (w2 > 0.0) ? w2 * rsqrt_newton1(w2) : 0.0;
  if (!(w < tz_wmax)) return false;
  const double z = (w > 0.0) ? -w * tz_eval(tz, w) : 0.0;
  if (fabs(z) > 1e-8) {
    const double sbh = fabs(z) * rsqrt_newton1(xx); // |z| / sqrt(x~'x~)
    b = (xy < 0.0) ? -sbh : sbh;
    v = ixx;
  } else {
    b = 0.0;
    v = INFINITY;
  }
  t = z;
  return true;
}

// log-domain evaluation of the BMA over all configurations for ONE lane (fallback of the linear-domain form: NaN /
// infinite statistics, or a sum outside the representable window); st = b, v, t^2 of the lane, stride 32
static __device__ __noinline__ double bma_all_logdomain(const double *__restrict__ st, int S, unsigned long long has_mask,
                                                        const PermGrid &pg)
{
  Lse bma;
  bma.init();
  const unsigned long long nconf = 1ull << S;
  for (unsigned long long cfg = 1; cfg < nconf; ++cfg) {
    Lse b;
    b.init();
    for (int k = 0; k < pg.K; ++k) {
      double den = 0.0, num = 0.0, sing = 0.0;
      unsigned long long mm = cfg & has_mask;
      while (mm) {
        const int s = __ffsll((long long)mm) - 1;
        mm &= mm - 1;
        double d, bd, sg;
        term_entry(st[s * 32], st[(S + s) * 32], sqrt(st[(2 * S + s) * 32]), pg.phiS[k], d, bd, sg); // (third row = t^2)
        den += d;
        num += bd;
        sing += sg;
      }
      b.add(abf_from_sums(den, num, sing, pg.omaS[k]), 1.0 / (double)pg.K, pg.kS[k] == 0);
    }
    // CalcBMA (gene_snp_pair.cpp:572-602); the order of the configurations does not matter for a weighted sum, except
    // for the NaN rule of element 0 (configuration "1" = subgroup 0 alone = cfg 1 here as well)
    bma.add(b.result(), pg.size_weight[__popcll(cfg)], cfg == 1);
  }
  return bma.result();
}

// explicit residual sum of squares of one genotype row against the basis rows kept in B (Gram form cancelled)
static __device__ __noinline__ void explicit_xx(const double *__restrict__ Xm, const double *__restrict__ brow0, size_t jstride,
                                                int Q, unsigned int colvalid, int n, int ldn, double &xx, double &xraw2,
                                                double &xsum)
{
  // row j of the basis: brow0 + j * jstride (row 0 = 0/1 mask, unit direction mask / sqrt(n))
  double h[MAXQ + 1], h2[MAXQ + 1];
  const double inv_n = 1.0 / (double)n;
  xraw2 = 0.0;
  xsum = 0.0;
  for (int k = 0; k <= Q; ++k) h[k] = 0.0;
  for (int i = 0; i < ldn; ++i) {
    const double mk = brow0[i];
    const double x = (mk != 0.0) ? Xm[i] : 0.0;
    xraw2 += x * x;
    xsum += x;
    for (int k = 1; k <= Q; ++k) h[k] += brow0[(size_t)k * jstride + i] * x;
  }
  h[0] = xsum * inv_n; // coefficient on the mask row
  for (int k = 1; k <= Q; ++k)
    if (!((colvalid >> k) & 1u)) h[k] = 0.0;
  for (int k = 0; k <= Q; ++k) h2[k] = 0.0;
  for (int i = 0; i < ldn; ++i) {
    const double mk = brow0[i];
    double x = (mk != 0.0) ? Xm[i] - h[0] : 0.0;
    for (int k = 1; k <= Q; ++k) x -= h[k] * brow0[(size_t)k * jstride + i];
    h2[0] += (mk != 0.0) ? x : 0.0;
    for (int k = 1; k <= Q; ++k) h2[k] += brow0[(size_t)k * jstride + i] * x;
  }
  h2[0] *= inv_n;
  for (int k = 1; k <= Q; ++k)
    if (!((colvalid >> k) & 1u)) h2[k] = 0.0;
  xx = 0.0;
  for (int i = 0; i < ldn; ++i) {
    const double mk = brow0[i];
    double x = (mk != 0.0) ? Xm[i] - h[0] - h2[0] : 0.0;
    for (int k = 1; k <= Q; ++k) x -= (h[k] + h2[k]) * brow0[(size_t)k * jstride + i];
    xx += x * x;
  }
}

// online-LSE forms of the "gen" row and of one singleton row (any value range, NaN rules of log10_weighted_sum): the
// fallbacks of the linear-domain accumulations of perm_bf_kernel
static __device__ __noinline__ double gen_row_slow(const double *__restrict__ st, int S, unsigned long long has_mask,
                                                   const PermGrid &pg, const TabRef T)
{
  LseTab rg;
  rg.init();
  const double wL = 1.0 / (double)pg.L;
  for (int u = 0; u < pg.UG; ++u) {
    double den, num, sing;
    {
      const double phi2 = pg.uphi[u];
      double tsum = 0.0, prod = 1.0, slog = 0.0;
      den = 0.0;
      num = 0.0;
      for (int s = 0; s < S; ++s) {
        const double t2 = st[(2 * S + s) * 32]; // (0 for the neutral element of a subgroup that contributes nothing)
        if (((has_mask >> s) & 1ull) && t2 != 0.0) {
          const double b = st[s * 32], v = st[(S + s) * 32];
          const double inv = 1.0 / (v + phi2);
          den += inv;
          num += b * inv;
          tsum += t2 * inv;
          prod *= v * inv;
          if (prod < 1e-200) {
            slog += log(prod);
            prod = 1.0;
          }
        }
      }
      sing = (phi2 == 0.0) ? 0.0 : (0.5 * (slog + log(prod)) + 0.5 * phi2 * tsum) * EQB_INV_LN10;
    }
    for (int i = pg.ustart[u]; i < pg.ustart[u + 1]; ++i) rg.add(abf_from_sums(den, num, sing, pg.omaL[i]), wL, pg.kL[i] == 0, T);
  }
  return rg.result(T);
}

static __device__ __noinline__ double single_row_slow(double b, double vv, double tt, const PermGrid &pg, const TabRef T)
{
  LseTab rc;
  rc.init();
  const double wK = 1.0 / (double)pg.K;
  for (int i = 0; i < pg.K; ++i) rc.add(singleton_value(b, vv, tt, pg.phiS[i], pg.omaS[i]), wK, pg.kS[i] == 0, T);
  return rc.result(T);
}

// Persistent warps: every warp draws tasks from a global counter (uniform cost per SNP, but the number of tasks is
// not a multiple of the resident warps: a static grid left a fifth of the SM time idle in its last wave).
template <bool ALLCFG>
__global__ void __launch_bounds__(256, 2) perm_bf_kernel(const DevParams *__restrict__ prm_, const FastParams *__restrict__ fp_,
                                                         const PermBatch pb, const __grid_constant__ PermGrid pg,
                                                         const BfTask *__restrict__ tasks, long long n_tasks, int warps_per_cta,
                                                         unsigned long long *__restrict__ next_task,
                                                         BfPartial *__restrict__ part, double *__restrict__ part_sep)
{
  const DevParams &prm = *prm_;
  extern __shared__ double bf_smem[];
  __shared__ BfTabs Tsm;
  __shared__ double Exp64[64];
  if (threadIdx.x < 16) {
    const double mj = 1.0 + ((double)threadIdx.x + 0.5) / 16.0;
    Tsm.exp16[threadIdx.x] = exp2((double)threadIdx.x / 16.0);
    Tsm.log16[threadIdx.x] = make_double2(1.0 / mj, log(mj));
  }
  if (ALLCFG && threadIdx.x < 64) Exp64[threadIdx.x] = exp2((double)threadIdx.x / 64.0);
  __syncthreads();
  TabRef T;
  T.base = smem_u32(&Tsm);
  const uint32_t base64 = smem_u32(Exp64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= warps_per_cta) return;
  const int S = prm.S, ldn = prm.ldn;
  const bool join = prm.analysis == 1;
  const int which = pb.which, kind = pb.stat_kind;
  double *wsm = bf_smem + (size_t)warp * bf_warp_doubles(S, which, kind);
  double *st = wsm + lane;                         // st[(r * S + s) * 32]: r = 0 b, 1 v, 2 t^2 (neutral element if inactive)
  double *hi = wsm + (size_t)3 * S * 32 + lane;    // hi[(sbit * 3 + r) * 32]  (--pbf all)
  const int SA = (S < PBF_SA) ? S : PBF_SA, SB = S - SA;
  double *sepm = wsm + bf_warp_doubles(S, which, 0) + lane; // sepm[s * 32] minima, sepm[(S + s) * 32] NaN flags

  while (true) {
    unsigned long long t64 = 0;
    if (lane == 0) t64 = atomicAdd(next_task, 1ull);
    t64 = __shfl_sync(0xffffffffu, t64, 0);
    if (t64 >= (unsigned long long)n_tasks) break;
    const long long task = (long long)t64;
    const BfTask tk = tasks[task];
    if (tk.c_lo >= pb.PB) continue; // column group past the (shorter) last batch
    const int il = tk.il, c = tk.c_lo + lane; // this lane's column
    const int g = pb.genes[il];
    const int cc = (c < pb.PB) ? c : pb.PB - 1; // lanes past the last column repeat it (results dropped)
    const long long mbeg = prm.cis_begin[g];
    const double *Drow = pb.D + (size_t)(pb.drow0[il] + (tk.m_begin - mbeg)) * pb.ldd + cc;
    const size_t scb = (size_t)il * S * pb.PBpad + cc;

    Lse acc_stat;
    acc_stat.init();
    double max_stat = -INFINITY, sep_all_min = 1.0;
    bool first_nan = false;
    int cnt_nonnan = 0;
    if (kind == STAT_SEP_PER)
      for (int s = 0; s < S; ++s) {
        sepm[s * 32] = INFINITY;
        sepm[(S + s) * 32] = 0.0;
      }

    for (long long m = tk.m_begin; m < tk.m_end; ++m, Drow += pb.ldd) {
      // ---- summary statistics + standardisation of every subgroup
      unsigned long long has_mask = 0ull;
      double snp_pmin = 1.0;
      for (int s = 0; s < S; ++s) {
        const SubDev &sb = prm.sub[s];
        const FastSub &fs = fp_->sub[s];
        const int n = pb.sc_n[scb + (size_t)s * pb.PBpad];
        const bool have = (n > 0) && sb.snp_has[m];
        double sb_b = nan(""), sb_v = nan(""), sb_t = nan(""), sb_p = nan("");
        if (have) {
          const double *d = Drow + (size_t)pb.dbase[(size_t)il * S + s] * pb.PBpad;
          const int Q = sb.Q;
          const int rankz = pb.sc_rankz[scb + (size_t)s * pb.PBpad];
          double xy, xx, r2, xsum;
          const bool comp = pb.complete[(size_t)il * S + s] != 0;
          if (m + 1 < tk.m_end) {
            // the next SNP's products of this subgroup: requested now, used one row of arithmetic later (D is read once,
            // straight from HBM: without the prefetch every row pays the DRAM latency in its dependency chain)
            const double *dn = d + pb.ldd;
            asm volatile("prefetch.global.L1 [%0];" ::"l"(dn));
            if (!comp) {
              asm volatile("prefetch.global.L1 [%0];" ::"l"(dn + (size_t)(Q + 1) * pb.PBpad));
              asm volatile("prefetch.global.L1 [%0];" ::"l"(dn + (size_t)(Q + 2) * pb.PBpad));
            }
          }
          if (comp) {
            const double *xs = fs.xstat + (size_t)m * 3; // K1c output: permutation-invariant
            xy = d[0];
            xx = xs[0];
            r2 = xs[1];
            xsum = xs[2];
          } else {
            xsum = d[0];
            xy = d[(size_t)(Q + 1) * pb.PBpad];
            r2 = d[(size_t)(Q + 2) * pb.PBpad];
            double hh = xsum * xsum / (double)n;
            for (int k = 1; k <= Q; ++k) {
              const double h = d[(size_t)k * pb.PBpad];
              hh = fma(h, h, hh);
            }
            xx = r2 - hh;
            if (r2 > 0.0 && xx < 1e-5 * r2) {
              // Gram form cancelled (genotype nearly inside span([1, covariates]) on the kept rows): explicit CGS2
              const double *b0 = pb.Bmat + ((size_t)pb.bbase[(size_t)il * S + s] * pb.PBpad + cc) * ldn;
              explicit_xx(sb.X + (size_t)m * ldn, b0, (size_t)pb.PBpad * ldn, Q, pb.sc_colvalid[scb + (size_t)s * pb.PBpad], n,
                          ldn, xx, r2, xsum);
            }
          }
          const double yy = pb.sc_yy[scb + (size_t)s * pb.PBpad];
          if (!join || !stats_lean(xy, xx, r2, yy, n, Q, rankz, fs.tz, fs.tz_nu, fs.tz_wmax, T, sb_b, sb_v, sb_t)) {
            const double nu = (double)n - 2.0 - Q;
            const bool use_tab = (fs.tz != nullptr) && (fs.tz_nu == nu);
            PairStat ps;
            stats_from_dots(xy, xx, r2, xsum, yy, pb.sc_tss[scb + (size_t)s * pb.PBpad], pb.sc_ybar[scb + (size_t)s * pb.PBpad], n,
                            Q, rankz, use_tab ? fs.tz : nullptr, fs.tz_nu, fs.tz_wmax, ps);
            sb_b = ps.b;
            sb_v = ps.v;
            sb_t = ps.t;
            sb_p = ps.pval;
          }
          has_mask |= 1ull << s;
        }
        // A subgroup without a result, or with |t| < 1e-8 (gene_snp_pair.cpp:314: it contributes nothing), is stored as the
        // NEUTRAL element b = 0, v = 1e300, t^2 = 0: 1 / (v + phi2) ~ 1e-300 and v / (v + phi2) = 1, so the sums over the
        // subgroups below need no branch.  NaN statistics stay NaN (they poison the sums exactly as in the reference).
        {
          const bool inactive = !have || fabs(sb_t) < 1e-8;
          st[s * 32] = inactive ? 0.0 : sb_b;
          st[(S + s) * 32] = inactive ? 1e300 : sb_v;
          st[(2 * S + s) * 32] = inactive ? 0.0 : sb_t * sb_t;
        }
        if (!join) {
          const double pval = have ? sb_p : nan("");
          if (pval < snp_pmin) snp_pmin = pval;
          if (kind == STAT_SEP_PER) {
            const double v = (sb.gene_has[g] && sb.snp_has[m]) ? pval : 1.0;
            if (isnan(v)) {
              if (m == mbeg) sepm[(S + s) * 32] = 1.0;
            } else if (v < sepm[s * 32])
              sepm[s * 32] = v;
          }
        }
      }
      if (!join) {
        if (snp_pmin < sep_all_min) sep_all_min = snp_pmin;
        continue;
      }
      double val; // the SNP's weighted ABF of the requested kind
      if (!ALLCFG || which != 3) {
        // ---- "gen": consistent configuration on gridL, one pass per unique phi2 (gene_snp_pair.cpp:364-416).
        // log10_weighted_sum in the LINEAR domain against the bound Mg = sum_s t_s^2 / 2 >= every ln ABF (a Bayes factor
        // cannot exceed the maximised likelihood ratio): no running maximum, one exp per grid point; a sum outside the
        // representable window (or a NaN) takes the online form of gen_row_slow
        double Mg = 0.0;
        for (int s = 0; s < S; ++s) Mg = fma(0.5, st[(2 * S + s) * 32], Mg);
        double accg = 0.0;
        for (int u = 0; u < pg.UG; ++u) {
          const double phi2 = pg.uphi[u];
          double den = 0.0, num = 0.0, tsum = 0.0, prod = 1.0, slog = 0.0;
          for (int s0 = 0; s0 < S; s0 += 8) {
            const int s1 = min(S, s0 + 8);
            for (int s = s0; s < s1; ++s) {
              const double b = st[s * 32], v = st[(S + s) * 32], t2 = st[(2 * S + s) * 32];
              const double inv = rcp_n(v + phi2);
              den += inv;
              num = fma(b, inv, num);
              tsum = fma(t2, inv, tsum);
              prod *= v * inv;
            }
            // (a factor v / (v + phi2) is never below ~1e-6: eight of them cannot underflow a product kept above 1e-200)
            if (prod < 1e-200) {
              slog += log(prod);
              prod = 1.0;
            }
          }
          // ln of the product of the single-subgroup ABFs (one logarithm per phi2), shifted by the bound
          const double sing = ((phi2 == 0.0) ? 0.0 : fma(0.5, slog + log_tab16_pos(prod, T), 0.5 * phi2 * tsum)) - Mg;
          const bool live = num != 0.0 && den != 0.0 && den == den; // (CalcLog10AbfUvlr's guards, see abf_from_sums)
          const double hn2 = 0.5 * num * num;
          for (int i = pg.ustart[u]; i < pg.ustart[u + 1]; ++i) {
            const double oma2 = pg.omaL[i];
            double xn = -Mg;
            if (live) {
              xn = sing;
              if (oma2 != 0.0) {
                const double z = fma(oma2, den, 1.0);
                xn += fma(-0.5, log_tab16_pos(z, T), hn2 * oma2 * rcp_n(z));
              }
            }
            accg += exp_tab16<false>(xn, T);
          }
        }
        if (accg > 1e-280 && accg < INFINITY) {
          val = (Mg + log_tab16_pos(accg * pg.invL, T)) * PGK[11];
          if (fabs(val) <= DBL_EPSILON) val = 0.0;
        } else
          val = gen_row_slow(st, S, has_mask, pg, T);
        if (which == 2) {
          // ---- singletons on gridS + BMAlite (gene_snp_pair.cpp:422-463, 552-570): grid points grouped by
          // phi2 + oma2 (the logarithm 0.5 ln(v / (v + phi2 + oma2)) is shared inside a group), linear domain
          // against the bound t^2 / 2
          LseTab lite;
          lite.init();
          const double wS = 0.5 / (double)S;
          for (int s = 0; s < S; ++s) {
            const double b = st[s * 32], vv = st[(S + s) * 32], t2 = st[(2 * S + s) * 32];
            const bool live = t2 != 0.0 && b != 0.0 && vv == vv; // (neutral element: t2 = 0; NaN statistics: vv != vv)
            double wc = 0.0; // every value of a subgroup without a usable statistic is 0 (gene_snp_pair.cpp:436-457)
            if (pg.K == 0)
              wc = nan("");
            else if (live) {
              const double Ms = 0.5 * t2, b2 = b * b;
              double acc = 0.0;
              for (int u = 0; u < pg.UT; ++u) {
                const double w = rcp_n(vv + pg.utot[u]);
                const double lg = fma(0.5, log_tab16_pos(vv * w, T), -Ms);
                const double c1 = b2 * w;
                for (int i = pg.tstart[u]; i < pg.tstart[u + 1]; ++i) {
                  const double inv = rcp_n(vv + pg.phiS[i]);
                  acc += exp_tab16<false>(fma(inv, fma(t2, pg.phiH[i], c1 * pg.omaH[i]), lg), T);
                }
              }
              if (acc > 1e-280 && acc < INFINITY) {
                wc = (Ms + log_tab16_pos(acc * pg.invK, T)) * PGK[11];
                if (fabs(wc) <= DBL_EPSILON) wc = 0.0;
              } else
                wc = single_row_slow(b, vv, sqrt(t2), pg, T);
            }
            lite.add(wc, wS, s == 0, T);
          }
          lite.add(val, 0.5, false, T);
          val = lite.result(T);
        }
      } else {
        // ---- all 2^S - 1 configurations, linear domain against Mref = sum_s t_s^2 / 2 - 350 (>= every exponent - 350):
        //   10^abf = exp(A) (1 + oma2 den)^-1/2 exp(num^2 oma2 / (2 (1 + oma2 den)))      gene_snp_pair.cpp:504-602
        bool fast_ok = true;
        double Mref = 0.0;
        for (int s = 0; s < S; ++s)
          if ((has_mask >> s) & 1ull) {
            const double t2 = st[(2 * S + s) * 32], b = st[s * 32], v = st[(S + s) * 32];
            if (isnan(t2) || isnan(b) || isnan(v) || isinf(b)) fast_ok = false;
            else Mref += 0.5 * t2;
          }
        Mref -= 350.0;
        double total = 0.0;
        if (fast_ok) {
          for (int k = 0; k < pg.K; ++k) {
            const double phi2 = pg.phiS[k], oma2 = pg.omaS[k], hom2 = 0.5 * oma2;
            // per-subgroup terms { oma2 / (v + phi2), b / (v + phi2), ln of the single-subgroup ABF }
            auto term = [&](int s, double &dz, double &dn, double &dA) { // (neutral elements give ~0, 0, ~0)
              const double b = st[s * 32], v = st[(S + s) * 32], t2 = st[(2 * S + s) * 32];
              const double inv = rcp_n(v + phi2);
              dz = oma2 * inv;
              dn = b * inv;
              dA = (phi2 == 0.0) ? 0.0 : fma(0.5, log_tab16_pos(v * inv, T), 0.5 * t2 * phi2 * inv);
            };
            // low part: the 2^SA subset sums of the subgroups 0..SA-1, in registers
            double lz[1 << PBF_SA], ln_[1 << PBF_SA], lA[1 << PBF_SA];
            {
              double tz_[PBF_SA], tn_[PBF_SA], tA_[PBF_SA];
#pragma unroll
              for (int s = 0; s < PBF_SA; ++s) {
                tz_[s] = tn_[s] = tA_[s] = 0.0;
                if (s < SA) term(s, tz_[s], tn_[s], tA_[s]);
              }
#pragma unroll
              for (int a = 0; a < (1 << PBF_SA); ++a) {
                lz[a] = ln_[a] = lA[a] = 0.0;
#pragma unroll
                for (int s = 0; s < PBF_SA; ++s)
                  if ((a >> s) & 1) {
                    lz[a] += tz_[s];
                    ln_[a] += tn_[s];
                    lA[a] += tA_[s];
                  }
              }
            }
            for (int s = 0; s < SB; ++s) {
              double dz, dn, dA;
              term(SA + s, dz, dn, dA);
              hi[(s * 3 + 0) * 32] = dz;
              hi[(s * 3 + 1) * 32] = dn;
              hi[(s * 3 + 2) * 32] = dA;
            }
            // high part: Gray code over the 2^SB subsets of the subgroups SA..S-1, sums kept in registers
            double hz = 1.0, hn = 0.0, hA = -Mref, ksum = 0.0;
            unsigned int gray = 0;
            const unsigned int nhi = 1u << SB;
            for (unsigned int ih = 0; ih < nhi; ++ih) {
              if (ih > 0) {
                const int bit = __ffs(ih) - 1; // warp-uniform
                gray ^= 1u << bit;
                const double sg = ((gray >> bit) & 1u) ? 1.0 : -1.0;
                hz = fma(sg, hi[(bit * 3 + 0) * 32], hz);
                hn = fma(sg, hi[(bit * 3 + 1) * 32], hn);
                hA = fma(sg, hi[(bit * 3 + 2) * 32], hA);
              }
              const double *wq = pg.size_weight + __popc(gray);
#pragma unroll
              for (int a = 0; a < (1 << PBF_SA); ++a) {
                if (a >= (1 << SA)) break; // (S < 3)
                const double r = rsqrt_newton1(hz + lz[a]);
                const double q = (hn + ln_[a]) * r; // num / sqrt(1 + oma2 den)
                const double e = exp_tab64(fma(q * q, hom2, hA + lA[a]), base64);
                ksum = fma(wq[__popc(a)], r * e, ksum);
              }
            }
            total += ksum;
          }
        }
        if (fast_ok && total > 1e-280 && total < INFINITY) {
          val = (Mref + log(total / (double)pg.K)) * EQB_INV_LN10;
          if (fabs(val) <= DBL_EPSILON) val = 0.0;
        } else
          val = bma_all_logdomain(st, S, has_mask, pg);
      }
      // ---- running statistic over the SNPs of the gene (gene.cpp:643-697)
      if (isnan(val)) {
        if (m == mbeg) first_nan = true;
      } else {
        cnt_nonnan++;
        if (val > max_stat) max_stat = val;
        acc_stat.add(val, 1.0, false);
      }
    }
    if (kind == STAT_SEP_PER) {
      for (int s = 0; s < S; ++s) {
        part_sep[((size_t)task * 2 * S + s) * 32 + lane] = sepm[s * 32];
        part_sep[((size_t)task * 2 * S + S + s) * 32 + lane] = sepm[(S + s) * 32];
      }
      continue;
    }
    BfPartial o;
    if (kind == STAT_JOIN_MAX) {
      o.m = max_stat;
      o.acc = 0.0;
    } else if (kind == STAT_JOIN_AVG) {
      o.m = acc_stat.m;
      o.acc = acc_stat.acc;
    } else {
      o.m = sep_all_min;
      o.acc = 0.0;
    }
    o.cnt_nonnan = cnt_nonnan;
    o.flags = (first_nan ? 1 : 0) | (acc_stat.any ? 2 : 0);
    part[(size_t)task * 32 + lane] = o;
  }
}

// thread per (item, column): merges the chunk partials of the gene in SNP order and applies the reference's rules
struct MergeArgs {
  const int *task0;     // [n_items] first task of the item (tasks ordered item, column group, chunk)
  const int *nchunk;    // [n_items]
  const long long *mg;  // [n_items] cis SNPs of the gene
  const BfPartial *part;
  const double *part_sep;
  double *out_true;     // [(row0 + item) * per + s]
  double *out_stat;     // [((row0 + item) * per + s) * P_total + p]
  long long row0;
};

__global__ void perm_merge_kernel(const PermBatch pb, const MergeArgs ma, int S)
{
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)pb.n_items * pb.PB) return;
  const int il = (int)(idx / pb.PB), c = (int)(idx % pb.PB);
  const long long p = pb.c0 + c - 1;
  const bool true_rules = p < 0; // statistic of the true data: NaN entries dropped first (gene.cpp:575-596)
  const int kind = pb.stat_kind;
  const int per = (kind == STAT_SEP_PER) ? S : 1;
  const int cg = c >> 5, lane = c & 31, nch = ma.nchunk[il];
  const long long t0 = (long long)ma.task0[il] + (long long)cg * nch;
  const long long Mg = ma.mg[il];
  const size_t row = (size_t)(ma.row0 + il) * per;
  if (kind == STAT_SEP_PER) {
    for (int s = 0; s < S; ++s) {
      double v = INFINITY;
      bool nanf = false;
      for (int ch = 0; ch < nch; ++ch) {
        v = fmin(v, ma.part_sep[((size_t)(t0 + ch) * 2 * S + s) * 32 + lane]);
        nanf = nanf || ma.part_sep[((size_t)(t0 + ch) * 2 * S + S + s) * 32 + lane] != 0.0;
      }
      if (true_rules) {
        if (!(v < 1.0)) v = 1.0;
      } else if (nanf)
        v = nan("");
      else if (Mg == 0 || isinf(v))
        v = 1.0;
      if (true_rules) ma.out_true[row + s] = v;
      else ma.out_stat[(row + s) * (size_t)pb.P_total + p] = v;
    }
    return;
  }
  bool fn = false;
  int nn = 0;
  double res;
  if (kind == STAT_JOIN_MAX) {
    double v = -INFINITY;
    for (int ch = 0; ch < nch; ++ch) {
      const BfPartial o = ma.part[(size_t)(t0 + ch) * 32 + lane];
      v = fmax(v, o.m);
      fn = fn || (o.flags & 1);
      nn += o.cnt_nonnan;
    }
    res = (fn && !true_rules) ? nan("") : v;
  } else if (kind == STAT_JOIN_AVG) {
    Lse t;
    t.init();
    for (int ch = 0; ch < nch; ++ch) {
      const BfPartial o = ma.part[(size_t)(t0 + ch) * 32 + lane];
      Lse q;
      q.m = o.m;
      q.acc = o.acc;
      q.any = (o.flags & 2) != 0;
      q.first_nan = false;
      t.merge(q);
      fn = fn || (o.flags & 1);
      nn += o.cnt_nonnan;
    }
    const double size = true_rules ? (double)nn : (double)Mg;
    if ((fn && !true_rules) || nn == 0)
      res = nan("");
    else {
      res = t.m + log10(t.acc * (1.0 / size));
      if (fabs(res) <= DBL_EPSILON) res = 0.0;
    }
  } else {
    double v = 1.0;
    for (int ch = 0; ch < nch; ++ch) v = fmin(v, ma.part[(size_t)(t0 + ch) * 32 + lane].m);
    res = v;
  }
  if (true_rules) ma.out_true[row] = res;
  else ma.out_stat[row * (size_t)pb.P_total + p] = res;
}

} // namespace eqb
