// fast_kernels.cuh -- the split K1 / K2+K3 path for the common case in which the set of
// individuals entering a subgroup's regression does not depend on the gene (no NaN expression
// inside the subgroup, gene expressed there) and the data are not permuted.
//
//   prep_basis_kernel  (K1a) one CTA per subgroup: orthonormal basis of [1, covariates] on the
//                      subgroup's individuals (CGS2), its rank.              gene_snp_pair.cpp:130-150
//   prep_y_kernel      (K1b) one warp per (gene, subgroup): residual phenotype, yy, tss, mean; flags
//                      the (gene, subgroup) cells that need the general path (NaN, absent gene).
//   prep_x_kernel      (K1c) one warp per SNP, all subgroups in one pass over the genotype row:
//                      residual genotype sum of squares xx = |x - QQ'x|^2 (+ raw moments).
//   fast_pair_kernel   (K2+K3) one CTA per gene, tiles of T cis SNPs, three compact phases:
//                      A  warp per SNP: x . ytil_s for every subgroup (the cis-banded contraction)
//                      B  thread per (pair, subgroup): betahat, se, sigmahat, pve, p-value and the
//                         standardisation (one Student tail + one normal quantile)
//                                           utils_math.cpp:166-209, gene_snp_pair.cpp:256-290
//                      C  thread per (pair, row): ABFs over the grid with online log-sum-exp
//                                           gene_snp_pair.cpp:297-622
// Genes with any non-generic (gene, subgroup) cell are routed to pair_kernel (general path).
#pragma once

#include "pair_kernel.cuh"
#include "table_math.cuh"

namespace eqb {

struct FastSub {
  const double *Bs;    // [Q+1][ldn] orthonormal basis (generic mask)
  const double *Ytil;  // [G][ldn] residual phenotype (valid where ystat flag = 1)
  const double *ystat; // [G][4] yy, tss, ybar, generic-flag
  const double *xstat; // [M][3] xx, xraw2, xsum
  int n, rankz;        // kept individuals, rank of [1, covariates]
  unsigned int colvalid;
  int pad;
  const double *tz;    // [TZ_NI][TZ_NC] Chebyshev coefficients of r(w) = -z(w)/w (see build_tz_kernel), or nullptr
  double tz_nu;        // degrees of freedom the table was built for
  double tz_wmax;      // the table is valid for w < tz_wmax
};

constexpr int TZ_NI = 39; // unit intervals of w in [0, 39)
constexpr int TZ_NC = 11; // Chebyshev coefficients per interval (degree 10)

struct FastParams {
  FastSub sub[MAXS];
};

struct FastArgs {
  const int *genes;            // fast genes of this launch
  int n_genes;
  int T;                       // pairs per tile
  int which;                   // 1 gen, 2 sin, 3 all
  long long n_pairs;           // compact pair range of this launch: [q_begin, n_pairs)
  long long q_begin;
  const int *tile_gene;        // [tiles of this launch] index (into genes) of the gene holding the tile's first pair
  const long long *fast_base;  // [n_genes] first compact pair index of each fast gene
  const long long *pair_off;   // [n_genes] first OUTPUT pair index of each fast gene
  int *out_n;
  double *out_ss, *out_gen, *out_cfg, *out_w;
  int use_dmma;                // fast_pair_warp_kernel: phase A on the FP64 tensor cores
  int delay_ns, delay_ctas, delay_sm; // timing experiment only (-DEQB_TUNING, EQB_FASTW_DELAY_US)
  int debug;                   // timing experiments only (-DEQB_TUNING, EQB_FASTW_DEBUG): 1 no raw-value stores, 2 no phase A, 4 no phase C
  long long n_tiles;           // fast_pair_warp_kernel: tiles of this launch
  const long long *tile_q0;    // [n_tiles + 1] first compact pair index of each tile (variable size, <= 32 pairs)
  // --bfs all (which == 3): fast_pair_warp_kernel is the first pass (contraction, summary statistics, the consistent
  // configuration) and leaves b, v, t of every (pair, subgroup) here for fast_pair_all_kernel (fast_all_kernel.cuh)
  double *st_all;              // [compact pair][3 S]
  unsigned long long *has_all; // [compact pair] subgroups with a result
  // --inss (eqb_bf_from_sstats): the standardised statistics are INPUT (st_all / has_all filled by sstats_std_kernel), the
  // contraction and the summary statistics are skipped, output pair = compact pair
  int from_st;
  // fixed-point genotype transport, one matrix for all subgroups: the resident integer numerators [M][ldn] and the table of
  // exact quotients k2v[k] = k / denom; phase A of fast_pair_warp_kernel reads these instead of the doubles (same values)
  const unsigned short *x16;
  const double *k2v;
};

// ---------------------------------------------------------------- K1a
static __global__ void __launch_bounds__(32) prep_basis_kernel(const DevParams *__restrict__ prm_, double *const *Bs_all,
                                                        const uint8_t *const *emask_all, int *__restrict__ n_out,
                                                        int *__restrict__ rankz_out, unsigned int *__restrict__ colvalid_out,
                                                        int *__restrict__ err_flag)
{
  const DevParams &prm = *prm_;
  const int s = blockIdx.x, lane = threadIdx.x, ldn = prm.ldn, N = prm.N;
  const SubDev &sb = prm.sub[s];
  double *q = Bs_all[s];
  const uint8_t *emask = emask_all[s];
  int n = 0;
  for (int i = lane; i < ldn; i += 32) {
    const bool keep = (i < N) && sb.gmask[i] && emask[i];
    q[i] = keep ? 1.0 : 0.0;
    n += keep ? 1 : 0;
  }
  n = warp_sum_int(n);
  __syncwarp();
  unsigned int colvalid = 0;
  int rankz = 0;
  if (n > 0) {
    const double inv_sqrt_n = 1.0 / sqrt((double)n);
    for (int i = lane; i < ldn; i += 32) q[i] = (q[i] != 0.0) ? inv_sqrt_n : 0.0;
    __syncwarp();
    colvalid = 1u;
    rankz = 1;
    for (int k = 1; k <= sb.Q; ++k) {
      double *qk = q + (size_t)k * ldn;
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
      if (__any_sync(0xffffffffu, missing) && lane == 0) atomicExch(err_flag + 2, 1); // only fatal if a gene uses it
      __syncwarp();
      for (int pass = 0; pass < 2; ++pass)
        for (int j = 0; j < k; ++j) {
          if (!((colvalid >> j) & 1u)) continue;
          const double *qj = q + (size_t)j * ldn;
          double h = 0.0;
          for (int i = lane; i < ldn; i += 32) h += qj[i] * qk[i];
          h = warp_sum(h);
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
  }
  if (lane == 0) {
    n_out[s] = n;
    rankz_out[s] = rankz;
    colvalid_out[s] = colvalid;
  }
}

// ---------------------------------------------------------------- K1b
// One warp per R (gene, subgroup) expression rows, the rows held in registers (element lane + 32 j), the
// subgroup's orthonormal basis staged once per CTA in shared memory when it fits (basis_in_smem), CTAs
// persistent over the genes of their subgroup (blockIdx.y).  Projection = blocks of 4 independent dot
// products, two passes (CGS2); every basis element read from shared memory is used for the R rows (the
// kernel is bound by shared-memory bandwidth, not by HBM: Y is small).
template <int NPL, int R>
__global__ void __launch_bounds__(THREADS) prep_y_kernel(const DevParams *__restrict__ prm_, const FastParams *__restrict__ fp_,
                                                         double *const *Ytil_all, double *const *ystat_all, int basis_in_smem)
{
  const DevParams &prm = *prm_;
  extern __shared__ double ysm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = blockIdx.y, ldn = prm.ldn;
  const SubDev &sb = prm.sub[s];
  const FastSub &fs = fp_->sub[s];
  const int Q = sb.Q, n = fs.n;
  const double *q = fs.Bs;
  if (basis_in_smem) {
    for (int idx = threadIdx.x; idx < (Q + 1) * ldn; idx += THREADS) ysm[idx] = q[idx];
    __syncthreads();
    q = ysm;
  }
  unsigned long long keepm = 0ull; // bit j: element lane + 32 j belongs to the subgroup's individuals
#pragma unroll
  for (int j = 0; j < NPL; ++j) {
    const int i = lane + 32 * j;
    if ((i < ldn) && q[i] != 0.0) keepm |= 1ull << j;
  }
  for (long long g0 = ((long long)blockIdx.x * WARPS + warp) * R; g0 < prm.G; g0 += (long long)gridDim.x * WARPS * R) {
    double yr[R][NPL], ybar[R], tss[R];
    bool generic[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long g = g0 + r;
      generic[r] = (g < prm.G) && sb.gene_has[g] && n > 0 && !prm.qnorm; // --qnorm goes through the general path
      const double *Yg = sb.Yall + (size_t)min(g, prm.G - 1) * ldn;
      int bad = 0;
      double ysum = 0.0;
#pragma unroll
      for (int j = 0; j < NPL; ++j) {
        const bool keep = generic[r] && ((keepm >> j) & 1ull);
        const double v = keep ? __ldcs(Yg + lane + 32 * j) : 0.0;
        if (keep && isnan(v)) bad = 1;
        yr[r][j] = v;
        ysum += v;
      }
      if (__any_sync(0xffffffffu, bad)) {
        generic[r] = false;
#pragma unroll
        for (int j = 0; j < NPL; ++j) yr[r][j] = 0.0;
      }
      ysum = warp_sum(ysum);
      ybar[r] = generic[r] ? ysum / n : 0.0;
      double ts = 0.0;
#pragma unroll
      for (int j = 0; j < NPL; ++j)
        if ((keepm >> j) & 1ull) {
          const double d = yr[r][j] - ybar[r];
          ts += d * d;
        }
      tss[r] = warp_sum(ts);
    }
    // sum of squares before the projection: the second (re-orthogonalisation) pass is only needed when the
    // projection removes almost everything (|y~|^2 < 1e-3 |y|^2); otherwise one pass on the orthonormal basis is
    // accurate to ~eps |y| / |y~| < 1e-14
    double yraw2[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      double a = 0.0;
#pragma unroll
      for (int j = 0; j < NPL; ++j) a += yr[r][j] * yr[r][j];
      yraw2[r] = warp_sum(a);
    }
    for (int pass = 0; pass < 2; ++pass) {
      if (pass == 1) {
        bool need = false;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          double a = 0.0;
#pragma unroll
          for (int j = 0; j < NPL; ++j) a += yr[r][j] * yr[r][j];
          a = warp_sum(a);
          need = need || (a < 1e-3 * yraw2[r]);
        }
        if (!need) break; // warp-uniform
      }
      for (int k0 = 0; k0 <= Q; k0 += 4) {
        double h[R][4];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int a = 0; a < 4; ++a) h[r][a] = 0.0;
#pragma unroll
        for (int j = 0; j < NPL; ++j) {
          const int i = lane + 32 * j;
          if (i < ldn) {
#pragma unroll
            for (int a = 0; a < 4; ++a)
              if (k0 + a <= Q) {
                const double qv = q[(size_t)(k0 + a) * ldn + i];
#pragma unroll
                for (int r = 0; r < R; ++r) h[r][a] += qv * yr[r][j];
              }
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
          for (int r = 0; r < R; ++r)
#pragma unroll
            for (int a = 0; a < 4; ++a) h[r][a] += __shfl_xor_sync(0xffffffffu, h[r][a], o);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
          if (!(k0 + a <= Q && ((fs.colvalid >> (k0 + a)) & 1u))) {
#pragma unroll
            for (int r = 0; r < R; ++r) h[r][a] = 0.0;
          }
#pragma unroll
        for (int j = 0; j < NPL; ++j) {
          const int i = lane + 32 * j;
          if (i < ldn) {
#pragma unroll
            for (int a = 0; a < 4; ++a)
              if (k0 + a <= Q) {
                const double qv = q[(size_t)(k0 + a) * ldn + i];
#pragma unroll
                for (int r = 0; r < R; ++r) yr[r][j] -= h[r][a] * qv;
              }
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long g = g0 + r;
      if (g >= prm.G) continue;
      double *yt = Ytil_all[s] + (size_t)g * ldn;
      double *ys = ystat_all[s] + (size_t)g * 4;
      double yy = 0.0;
#pragma unroll
      for (int j = 0; j < NPL; ++j) {
        const int i = lane + 32 * j;
        if (i < ldn) {
          yt[i] = generic[r] ? yr[r][j] : 0.0;
          yy += yr[r][j] * yr[r][j];
        }
      }
      yy = warp_sum(yy);
      if (lane == 0) {
        if (generic[r]) {
          ys[0] = yy;
          ys[1] = tss[r];
          ys[2] = ybar[r];
          ys[3] = 1.0;
        } else {
          ys[0] = 0.0;
          ys[1] = 0.0;
          ys[2] = 0.0;
          // 2 = nothing to compute in this cell (gene not expressed here / no individual), 0 = general path
          ys[3] = (!sb.gene_has[g] || n == 0) ? 2.0 : 0.0;
        }
      }
    }
  }
}

// ---------------------------------------------------------------- K1c
// One warp per SNP.  The projection on the basis is classical Gram-Schmidt applied twice (CGS2):
// inside a pass the Q+1 dot products are independent, so their partial sums and the warp
// reductions overlap instead of forming a dependent chain.  Subgroups that share the genotype
// matrix, the mask and the covariates (dup_of[s] >= 0) reuse the result of the earlier subgroup.
template <int NPL>
__global__ void __launch_bounds__(THREADS) prep_x_kernel(const DevParams *__restrict__ prm_, const FastParams *__restrict__ fp_,
                                                         double *const *xstat_all, const int *__restrict__ dup_of,
                                                         const int fixup_only, const unsigned long long *__restrict__ fix_list,
                                                         int fix_cap, long long m_lo, long long m_hi)
{
  const DevParams &prm = *prm_;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = prm.S, ldn = prm.ldn;
  // fix-up mode: grid-stride walk over the entries queued by the DMMA pass (fix_list[0] = count; if the
  // list overflowed, every SNP is re-checked)
  const bool listed = fixup_only && fix_list != nullptr && fix_list[0] + 1 < (unsigned long long)fix_cap;
  const long long n_items = listed ? (long long)fix_list[0] : m_hi - m_lo; // unlisted: SNP rows [m_lo, m_hi)
  for (long long item = (long long)blockIdx.x * WARPS + warp; item < n_items; item += (long long)gridDim.x * WARPS) {
  const long long m = listed ? (long long)(fix_list[item + 1] >> 8) : m_lo + item;
  const int s_only = listed ? (int)(fix_list[item + 1] & 0xffull) : -1;
  for (int s = 0; s < S; ++s) {
    if (s_only >= 0 && s != s_only) continue;
    const SubDev &sb = prm.sub[s];
    const FastSub &fs = fp_->sub[s];
    double *xs = xstat_all[s] + (size_t)m * 3;
    if (!sb.snp_has[m] || fs.n == 0) {
      if (lane == 0) {
        xs[0] = 0.0;
        xs[1] = 0.0;
        xs[2] = 0.0;
      }
      continue;
    }
    if (fixup_only) {
      if (dup_of[s] >= 0) continue; // aliases the output of an identical earlier subgroup
      // second pass after the DMMA projection: only the entries whose Gram-form residual lost
      // accuracy (x nearly inside span([1, covariates])) are recomputed with explicit CGS2
      const double xx0 = xs[0], xr0 = xs[1];
      if (!(xr0 > 0.0 && xx0 < 1e-2 * xr0)) continue;
    } else if (dup_of[s] >= 0 && prm.sub[dup_of[s]].snp_has[m]) {
      // same genotype row, same individuals, same covariates: identical residual
      if (lane == 0) {
        const double *src = xstat_all[dup_of[s]] + (size_t)m * 3;
        xs[0] = src[0];
        xs[1] = src[1];
        xs[2] = src[2];
      }
      continue;
    }
    const double *q = fs.Bs;
    const double *Xm = sb.X + (size_t)m * ldn;
    double xr[NPL];
    double xraw2 = 0.0, xsum = 0.0;
#pragma unroll
    for (int j = 0; j < NPL; ++j) {
      const int i = lane + 32 * j;
      const double v = (i < ldn && q[i] != 0.0) ? Xm[i] : 0.0;
      xr[j] = v;
      xraw2 += v * v;
      xsum += v;
    }
    xraw2 = warp_sum(xraw2);
    xsum = warp_sum(xsum);
    const int Q = sb.Q;
    for (int pass = 0; pass < 2; ++pass) {
      for (int k0 = 0; k0 <= Q; k0 += 4) {
        double h[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int j = 0; j < NPL; ++j) {
          const int i = lane + 32 * j;
          if (i < ldn) {
#pragma unroll
            for (int a = 0; a < 4; ++a)
              if (k0 + a <= Q) h[a] += q[(size_t)(k0 + a) * ldn + i] * xr[j];
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
          for (int a = 0; a < 4; ++a) h[a] += __shfl_xor_sync(0xffffffffu, h[a], o);
        }
        // NB: the four projections of a block are removed together (CGS inside the block)
#pragma unroll
        for (int j = 0; j < NPL; ++j) {
          const int i = lane + 32 * j;
          if (i < ldn) {
            double v = xr[j];
#pragma unroll
            for (int a = 0; a < 4; ++a)
              if (k0 + a <= Q && ((fs.colvalid >> (k0 + a)) & 1u)) v -= h[a] * q[(size_t)(k0 + a) * ldn + i];
            xr[j] = v;
          }
        }
      }
    }
    double xx = 0.0;
#pragma unroll
    for (int j = 0; j < NPL; ++j) xx += xr[j] * xr[j];
    xx = warp_sum(xx);
    if (lane == 0) {
      xs[0] = xx;
      xs[1] = xraw2;
      xs[2] = xsum;
    }
  }
  }
}

// ---------------------------------------------------------------- K1c on the FP64 tensor cores
// H = X * Bcat (projections on every basis column) and R2 = (X o X) * Mcat (masked sums of squares)
// with mma.sync m8n8k4 f64 (DMMA; tcgen05 has no f64 kind).  One warp = 8 SNP rows; the k index is
// permuted so that lane (row g, k-slot kk) streams the contiguous quarter kk of its genotype row
// straight from global memory (16-byte loads), Bcat / Mcat sit in shared memory with a row stride
// = 1 mod 16 doubles (conflict-free fragment loads).  Gram form: xx = R2 - sum_k H_k^2.
struct PrepCols {
  int n_sub;            // subgroups of this chunk
  int sub[16];          // their indexes
  int col0[16];         // first basis column of each
  int ncol[16];         // number of basis columns (Q+1)
  int mcol[16];         // mask column of each
  double sqrt_n[16];
};

__device__ __forceinline__ void dmma_m8n8k4(double &d0, double &d1, double a, double b)
{
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

template <int NT, int NM, int NW>
__global__ void __launch_bounds__(NW * 32) prep_x_dmma_kernel(const DevParams *__restrict__ prm_, const double *__restrict__ X,
                                                              const double *__restrict__ Bcat, const double *__restrict__ Mcat,
                                                              const PrepCols pc, double *const *xstat_all,
                                                              unsigned long long *__restrict__ fix_list, int fix_cap,
                                                              long long blk_lo, long long blk_hi)
{
  // Blocks of 8 SNP rows [blk_lo, blk_hi) (a row chunk of the upload pipeline, or every row).
  // Persistent CTAs: Bcat / Mcat are staged once per CTA, then every warp walks its blocks of 8 SNP rows.
  // The genotype stream is register double-buffered in chunks of 8 x 16 bytes per lane, ACROSS block
  // boundaries, so that a warp always has its next chunk in flight while the tensor pipe works.
  const DevParams &prm = *prm_;
  extern __shared__ double psm[];
  const int ldn = prm.ldn, ldn4 = ldn >> 2, strideB = ldn + 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, kk = lane & 3;
  double *Bsm = psm;                                  // [NT*8][strideB]
  double *Msm = Bsm + (size_t)NT * 8 * strideB;       // [NM*8][strideB]
  double *Hsm = Msm + (size_t)NM * 8 * strideB;       // [NW][8][(NT+NM)*8]
  for (int idx = threadIdx.x; idx < NT * 8 * ldn; idx += NW * 32) {
    const int c = idx / ldn, i = idx % ldn;
    Bsm[(size_t)c * strideB + i] = Bcat[idx];
  }
  for (int idx = threadIdx.x; idx < NM * 8 * ldn; idx += NW * 32) {
    const int c = idx / ldn, i = idx % ldn;
    Msm[(size_t)c * strideB + i] = Mcat[idx];
  }
  __syncthreads();
  constexpr int U = 8, CH = 2 * U;                    // doubles per lane per chunk
  const long long M = prm.M, nblk = blk_hi;
  const long long bstep = (long long)gridDim.x * NW;
  long long blk = blk_lo + (long long)blockIdx.x * NW + warp;
  if (blk >= nblk) return;
  const int nch = (ldn4 + CH - 1) / CH;
  const double *brow = Bsm + (size_t)g * strideB + (size_t)kk * ldn4;
  const double *mrowp = Msm + (size_t)g * strideB + (size_t)kk * ldn4;
  const int W = (NT + NM) * 8;
  double *Hw = Hsm + (size_t)warp * 8 * W;
  double c[NT][2], c2[NM][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) c[nt][0] = c[nt][1] = 0.0;
#pragma unroll
  for (int nm = 0; nm < NM; ++nm) c2[nm][0] = c2[nm][1] = 0.0;
  double2 cur[U], nxt[U];
  {
    const double *xr = X + (size_t)min(blk * 8 + g, M - 1) * ldn + (size_t)kk * ldn4;
#pragma unroll
    for (int u = 0; u < U; ++u)
      cur[u] = (2 * u < ldn4) ? __ldcs(reinterpret_cast<const double2 *>(xr + 2 * u)) : make_double2(0.0, 0.0);
  }
  int ch = 0;
  while (true) {
    // prefetch the next chunk (of this block, or the first one of the warp's next block)
    const bool last = (ch + 1 == nch);
    const long long nblk_id = last ? blk + bstep : blk;
    const int nch_id = last ? 0 : ch + 1;
    const bool more = nblk_id < nblk;
    if (more) {
      const double *xr = X + (size_t)min(nblk_id * 8 + g, M - 1) * ldn + (size_t)kk * ldn4 + nch_id * CH;
#pragma unroll
      for (int u = 0; u < U; ++u)
        nxt[u] = (nch_id * CH + 2 * u < ldn4) ? __ldcs(reinterpret_cast<const double2 *>(xr + 2 * u)) : make_double2(0.0, 0.0);
    }
    const int t = ch * CH;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (t + 2 * u < ldn4) { // warp-uniform
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          dmma_m8n8k4(c[nt][0], c[nt][1], cur[u].x, brow[(size_t)nt * 8 * strideB + t + 2 * u]);
          dmma_m8n8k4(c[nt][0], c[nt][1], cur[u].y, brow[(size_t)nt * 8 * strideB + t + 2 * u + 1]);
        }
        const double ax2 = cur[u].x * cur[u].x, ay2 = cur[u].y * cur[u].y;
#pragma unroll
        for (int nm = 0; nm < NM; ++nm) {
          dmma_m8n8k4(c2[nm][0], c2[nm][1], ax2, mrowp[(size_t)nm * 8 * strideB + t + 2 * u]);
          dmma_m8n8k4(c2[nm][0], c2[nm][1], ay2, mrowp[(size_t)nm * 8 * strideB + t + 2 * u + 1]);
        }
      }
    }
    if (last) {
      // epilogue through shared memory: per (row, subgroup) xx = R2 - sum H^2, xraw2 = R2, xsum = sqrt(n) H_0
      const long long m0 = blk * 8;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        Hw[g * W + nt * 8 + 2 * kk] = c[nt][0];
        Hw[g * W + nt * 8 + 2 * kk + 1] = c[nt][1];
        c[nt][0] = c[nt][1] = 0.0;
      }
#pragma unroll
      for (int nm = 0; nm < NM; ++nm) {
        Hw[g * W + NT * 8 + nm * 8 + 2 * kk] = c2[nm][0];
        Hw[g * W + NT * 8 + nm * 8 + 2 * kk + 1] = c2[nm][1];
        c2[nm][0] = c2[nm][1] = 0.0;
      }
      __syncwarp();
      for (int it = lane; it < 8 * pc.n_sub; it += 32) {
        const int r = it / pc.n_sub, si = it % pc.n_sub;
        const long long m = m0 + r;
        if (m >= M) continue;
        const int s = pc.sub[si];
        double *xs = xstat_all[s] + (size_t)m * 3;
        if (!prm.sub[s].snp_has[m]) {
          xs[0] = 0.0;
          xs[1] = 0.0;
          xs[2] = 0.0;
          continue;
        }
        const double *h = Hw + r * W + pc.col0[si];
        double hh = 0.0;
        for (int k = 0; k < pc.ncol[si]; ++k) hh += h[k] * h[k];
        const double r2 = Hw[r * W + NT * 8 + pc.mcol[si]];
        xs[0] = r2 - hh;
        xs[1] = r2;
        xs[2] = pc.sqrt_n[si] * h[0];
        if (r2 > 0.0 && (r2 - hh) < 1e-2 * r2) {
          // Gram-form residual lost accuracy (x nearly inside span([1, covariates])): queue for the explicit pass
          const unsigned long long slot = atomicAdd(fix_list, 1ull);
          if (slot + 1 < (unsigned long long)fix_cap) fix_list[slot + 1] = ((unsigned long long)m << 8) | (unsigned long long)s;
        }
      }
      __syncwarp();
    }
    if (!more) break;
#pragma unroll
    for (int u = 0; u < U; ++u) cur[u] = nxt[u];
    blk = nblk_id;
    ch = nch_id;
  }
}

// ---------------------------------------------------------------- K2 + K3
// ---------------------------------------------------------------- Student-t -> normal score table
// The standardisation maps |t| (nu degrees of freedom) to z = Phi^-1(T_nu(-|t|))
// (gene_snp_pair.cpp:273-274).  In the variable w = sqrt(nu log1p(t^2/nu)) the ratio r(w) = -z/w is
// a smooth function close to 1; it is tabulated per subgroup (nu is shared by every pair of the
// subgroup on this path) as piecewise Chebyshev series fitted to the exact evaluation
// (tdist_P + ugaussian_Pinv), |error| ~ 1e-13, so that the per-pair cost is one short Clenshaw
// recurrence instead of a data-dependent continued fraction.  p-value = 2 Phi(z) = erfc(|z|/sqrt 2).
static __global__ void __launch_bounds__(TZ_NI * 16) build_tz_kernel(const double *__restrict__ nus, double *const *tz_all,
                                                              double *__restrict__ wmax_out)
{
  __shared__ double f[TZ_NI][16];
  __shared__ int bad[TZ_NI];
  const int s = blockIdx.x, iv = threadIdx.x / 16, k = threadIdx.x % 16;
  const double nu = nus[s];
  if (k == 0) bad[iv] = 0;
  __syncthreads();
  if (k < TZ_NC && nu > 0.0) {
    const double x = cospi((k + 0.5) / TZ_NC);
    const double w = iv + 0.5 + 0.5 * x;
    const double a = sqrt(nu * expm1(w * w / nu));
    const double z = ugaussian_Pinv(tdist_P(-a, nu));
    const double r = -z / w;
    f[iv][k] = r;
    if (!(fabs(r) < 1e300)) atomicExch(&bad[iv], 1);
  }
  __syncthreads();
  if (k < TZ_NC && nu > 0.0) {
    double c = 0.0;
    for (int j = 0; j < TZ_NC; ++j) c += f[iv][j] * cospi(k * (j + 0.5) / TZ_NC);
    c *= 2.0 / TZ_NC;
    if (k == 0) c *= 0.5;
    tz_all[s][iv * TZ_NC + k] = c;
  }
  if (threadIdx.x == 0) {
    int first_bad = TZ_NI;
    for (int i = TZ_NI - 1; i >= 0; --i)
      if (bad[i]) first_bad = i;
    wmax_out[s] = (nu > 0.0) ? (double)first_bad : 0.0;
  }
}

__device__ __forceinline__ double tz_eval(const double *__restrict__ tz, double w)
{
  const int iv = (int)w;
  const double x = 2.0 * (w - iv) - 1.0;
  const double *c = tz + iv * TZ_NC;
  double b1 = 0.0, b2 = 0.0;
#pragma unroll
  for (int j = TZ_NC - 1; j >= 1; --j) {
    const double t = 2.0 * x * b1 - b2 + c[j];
    b2 = b1;
    b1 = t;
  }
  return x * b1 - b2 + c[0];
}

// per-thread standardisation + summary statistics of one (pair, subgroup)
struct PairStat {
  double pve, sigmahat, betahat, se, pval; // outputs of utils::FitSingleGeneWithSingleSnp
  double b, v, t;                          // standardised (gene_snp_pair.cpp:256-290)
};

static __device__ __noinline__ void stats_from_dots(double xy, double xx, double xraw2, double xsum, double yy, double tss,
                                             double ybar, int n, int Q, int rankz, const double *__restrict__ tz,
                                             double tz_nu, double tz_wmax, PairStat &o)
{
  const double qn = nan("");
  o.pve = o.sigmahat = o.betahat = o.se = o.pval = qn;
  o.b = o.v = o.t = qn;
  if (n < (2 + Q) + 1) return; // utils_math.cpp:175
  double tail = qn; // two-sided Student tail of |betahat/se| with n - rank degrees of freedom
  int rank;
  if (xx > 1e-24 * xraw2 && xraw2 > 0.0) {
    rank = rankz + 1;
    o.betahat = xy / xx;
    double rss = yy - xy * o.betahat;
    if (rss < 0.0) rss = 0.0;
    o.pve = 1.0 - rss / tss;
    o.sigmahat = sqrt(rss / (double)(n - rank));
    o.se = o.sigmahat / sqrt(xx);
  } else {
    rank = rankz;
    const double rss = yy;
    o.pve = 1.0 - rss / tss;
    o.sigmahat = sqrt(rss / (double)(n - rank));
    if (xraw2 == 0.0) {
      o.betahat = 0.0;
      o.se = 0.0;
    } else if (Q == 0) {
      const double cst = xsum / n;
      double f0 = 1.0, f1 = 1.0, s0 = (double)n, s1 = fabs(cst) * n;
      while (s0 > 1.0) { s0 /= 2.0; f0 *= 2.0; }
      while (s0 < 0.5) { s0 *= 2.0; f0 /= 2.0; }
      while (s1 > 1.0) { s1 /= 2.0; f1 *= 2.0; }
      while (s1 < 0.5) { s1 *= 2.0; f1 /= 2.0; }
      const double a = 1.0 / f0, b = cst / f1, ab2 = a * a + b * b;
      o.betahat = b * ybar / (ab2 * f1);
      const double s2 = rss / (double)(n - rank);
      o.se = sqrt(s2 * b * b / (ab2 * ab2 * n) / (f1 * f1));
    } else
      return; // documented unsupported degenerate design (NaN)
  }
  const double tt = o.betahat / o.se;
  double bhat = o.betahat / o.sigmahat, sebhat = o.se / o.sigmahat, t = bhat / sebhat;
  const double nu = (double)n - 2.0 - Q;
  bool done = false;
  if (tz != nullptr && !isnan(tt) && !isnan(t) && nu == (double)(n - rank) && nu == tz_nu) {
    // tabulated map |t| -> z (same degrees of freedom for the p-value and the standardisation)
    const double a = fabs(t);
    const double w = sqrt(nu * log1p(a * a / nu));
    if (w < tz_wmax) {
      const double z = (w > 0.0) ? -w * tz_eval(tz, w) : 0.0;
      o.pval = (fabs(tt) > 0.0) ? erfc(-z * 0.70710678118654752440) : 1.0;
      t = z;
      done = true;
    }
  }
  if (!done) {
    double central;
    if (!isnan(tt)) {
      tdist_tails(fabs(tt), (double)(n - rank), tail, central);
      o.pval = (fabs(tt) > 0.0) ? tail : 1.0; // 2 * gsl_cdf_tdist_Q(|t|, n - rank)
    }
    if (isnan(t)) return;
    double P;
    if (nu == (double)(n - rank) && !isnan(tail))
      P = (t == 0.0) ? 0.5 : 0.5 * tail; // gsl_cdf_tdist_P(-|t|, nu) shares the tail
    else
      P = tdist_P(-fabs(t), nu);
    t = ugaussian_Pinv(P);
  }
  if (fabs(t) > 1e-8) {
    const double sg2 = fabs(o.betahat) / (fabs(t) * sebhat);
    bhat = o.betahat / sg2;
    sebhat = fabs(bhat / t);
  } else {
    bhat = 0.0;
    sebhat = INFINITY;
  }
  o.b = bhat;
  o.v = sebhat * sebhat;
  o.t = t;
}

// --inss: GeneSnpPair::SetSstats + StandardizeSstatsAndCorrectSmallSampleSize (gene_snp_pair.cpp:241-290) for summary
// statistics read from files: thread per (pair, subgroup); n <= 0 = no entry for that subgroup.  The number of covariates
// is unknown to the reference on this path (its map lookup default-constructs 0): nu = n - 2.
static __global__ void sstats_std_kernel(long long n_items, int S, const int *__restrict__ nn, const double *__restrict__ sigmahat,
                                         const double *__restrict__ betahat, const double *__restrict__ sebetahat,
                                         double *__restrict__ st_all, unsigned long long *__restrict__ has_all)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_items) return;
  const long long pair = i / S;
  const int s = (int)(i - pair * S);
  double b = nan(""), v = nan(""), t = nan("");
  const int n = nn[i];
  if (n > 0) {
    atomicOr(&has_all[pair], 1ull << s);
    double bhat = betahat[i] / sigmahat[i], sebhat = sebetahat[i] / sigmahat[i];
    t = bhat / sebhat;
    if (!isnan(t)) {
      const double nu = (double)n - 2.0;
      t = ugaussian_Pinv(tdist_P(-fabs(bhat / sebhat), nu));
      if (fabs(t) > 1e-8) {
        const double sg2 = fabs(betahat[i]) / (fabs(t) * sebhat);
        bhat = betahat[i] / sg2;
        sebhat = fabs(bhat / t);
      } else {
        bhat = 0.0;
        sebhat = INFINITY;
      }
      b = bhat;
      v = sebhat * sebhat;
    }
  }
  double *o = st_all + pair * 3 * S;
  o[s] = b;
  o[S + s] = v;
  o[2 * S + s] = t;
}

// unique phi2 values of the consistent-configuration rows (gen / gen-fix / gen-maxh on gridL):
// uphi[UL]; row r, grid point k uses uphi[idxL[r*L+k]] and omaL[r*L+k]
struct GridTab {
  const double *uphi;
  const int *idxL;
  const double *omaL;
  int UL;
  int pad; // number of leading entries of uphi used by the gen row (the permutation statistic needs only those)
};

// ES-model log10 ABF from the sums over the active subgroups (CalcLog10AbfUvlr, gene_snp_pair.cpp:332-355)
// lbar = 0.5 log10(V) - 0.5 log10(V+oma2) + 0.5 T2 oma2/(V+oma2)/ln10 with V = 1/den, T2 = num^2/den,
// rewritten as -0.5 log10(1 + oma2 den) + 0.5 num^2 oma2 / (1 + oma2 den) / ln10.
// The reference's guards "bbar != 0 && V < +Inf" and "T2 != 0" are the exact-zero cases num == 0 / den == 0
// (every contributing term has |t| >= 1e-8, so neither quotient can underflow), tested without the divisions.
#define EQB_INV_LN10 0.43429448190325182765
__device__ __forceinline__ double abf_from_sums(double den, double num, double sing, double oma2)
{
  if (num != 0.0 && den != 0.0 && den == den) { // (a NaN denominator fails the reference's "V < +Inf")
    if (oma2 == 0.0) return sing; // (a zero numerator would take the slow division path)
    const double z = 1.0 + oma2 * den; // log(1 + x) has an ABSOLUTE error of ~1e-16 here: no log1p needed
    return sing + (-0.5 * log_fast(z) + 0.5 * num * num * oma2 * rcp_fast(z)) * EQB_INV_LN10;
  }
  return 0.0;
}

// one (subgroup, phi2) term: { 1/(v+phi2), b/(v+phi2), single-subgroup log10 ABF }, the latter as
// -0.5 log10(1 + phi2/v) + 0.5 t^2 phi2/(v+phi2)/ln10  (gene_snp_pair.cpp:314-324)
__device__ __forceinline__ void term_entry(double b, double v, double t, double phi2, double &d, double &bd, double &sg)
{
  if (fabs(t) < 1e-8) {
    d = 0.0;
    bd = 0.0;
    sg = 0.0;
  } else {
    const double inv = rcp_fast(v + phi2);
    d = inv;
    bd = b * inv;
    // -0.5 log10(1 + phi2/v) = 0.5 log10(v / (v + phi2)) = 0.5 log10(v * inv)
    sg = (phi2 == 0.0) ? 0.0 : (0.5 * log_fast(v * inv) + 0.5 * t * t * phi2 * inv) * EQB_INV_LN10;
  }
}

// sums over the subgroups in `mask` for one phi2: den = sum 1/(v+phi2), num = sum b/(v+phi2) and the sum of
// the single-subgroup log10 ABFs with ONE logarithm:
//   sum_s 0.5 log10(v_s / (v_s + phi2)) = 0.5 log10(prod_s v_s / (v_s + phi2))
// (the running product is folded into slog long before it could underflow).  st = {b[S], v[S], t[S]}.
__device__ __forceinline__ void consistent_sums(const double *__restrict__ st, int S, unsigned long long mask, double phi2,
                                                double &den, double &num, double &sing)
{
  double tsum = 0.0, prod = 1.0, slog = 0.0;
  den = 0.0;
  num = 0.0;
  while (mask) {
    const int s = __ffsll((long long)mask) - 1;
    mask &= mask - 1;
    const double b = st[s], v = st[S + s], tt = st[2 * S + s];
    if (!(fabs(tt) < 1e-8)) { // (gene_snp_pair.cpp:314: |t| < 1e-8 contributes nothing)
      const double inv = rcp_fast(v + phi2);
      den += inv;
      num += b * inv;
      tsum += tt * tt * inv;
      prod *= v * inv;
      if (prod < 1e-200) {
        slog += log(prod);
        prod = 1.0;
      }
    }
  }
  sing = (phi2 == 0.0) ? 0.0 : (0.5 * (slog + log_fast(prod)) + 0.5 * phi2 * tsum) * EQB_INV_LN10;
}

// the same three helpers on the table-driven elementary functions (table_math.cuh: absolute error of a log10 ABF
// < 1e-12, four orders of magnitude inside the 1e-8 budget), used by the warp-autonomous kernel whose phase C is bound
// by instruction issue: log / exp10 / division of the CUDA library were two thirds of its instructions
__device__ __forceinline__ double abf_from_sums_t(double den, double num, double sing, double oma2, const TabRef T)
{
  if (num != 0.0 && den != 0.0 && den == den) {
    if (oma2 == 0.0) return sing;
    const double z = fma(oma2, den, 1.0);
    return fma(fma(-0.5, log_tab16(z, T), 0.5 * num * num * oma2 * rcp_n(z)), EQB_INV_LN10, sing);
  }
  return 0.0;
}
__device__ __forceinline__ void consistent_sums_t(const double *__restrict__ st, int S, unsigned long long mask, double phi2,
                                                  double &den, double &num, double &sing, const TabRef T)
{
  double tsum = 0.0, prod = 1.0, slog = 0.0;
  den = 0.0;
  num = 0.0;
  while (mask) {
    const int s = __ffsll((long long)mask) - 1;
    mask &= mask - 1;
    const double b = st[s], v = st[S + s], tt = st[2 * S + s];
    if (!(fabs(tt) < 1e-8)) { // (gene_snp_pair.cpp:314: |t| < 1e-8 contributes nothing)
      const double inv = rcp_n(v + phi2);
      den += inv;
      num = fma(b, inv, num);
      tsum = fma(tt * tt, inv, tsum);
      prod *= v * inv;
      if (prod < 1e-200) {
        slog += log(prod);
        prod = 1.0;
      }
    }
  }
  sing = (phi2 == 0.0) ? 0.0 : (0.5 * (slog + log_tab16(prod, T)) + 0.5 * phi2 * tsum) * EQB_INV_LN10;
}
__device__ __forceinline__ double singleton_value_t(double b, double vv, double tt, double phi2, double oma2, const TabRef T)
{
  // (vv is +inf for |t| <= 1e-8: the guard comes first, the reciprocal seed is not defined there)
  if (!(fabs(tt) < 1e-8) && b != 0.0 && vv == vv && vv < INFINITY) {
    const double inv = rcp_n(vv + phi2), w = rcp_n(vv + phi2 + oma2);
    return (0.5 * log_tab16(vv * w, T) + 0.5 * inv * fma(tt * tt, phi2, b * b * oma2 * w)) * EQB_INV_LN10;
  }
  return 0.0;
}

// singleton configuration: term + ABF of one subgroup merged,
// 0.5 log10(v/(v+phi2)) - 0.5 log10(1 + oma2/(v+phi2)) = 0.5 log10(v / (v + phi2 + oma2))  (one logarithm);
// guards of CalcLog10AbfUvlr as in abf_from_sums
__device__ __forceinline__ double singleton_value(double b, double vv, double tt, double phi2, double oma2)
{
  const double inv = rcp_fast(vv + phi2);
  if (!(fabs(tt) < 1e-8) && b != 0.0 && inv != 0.0 && inv == inv) {
    const double w = rcp_fast(vv + phi2 + oma2);
    return (0.5 * log_fast(vv * w) + 0.5 * inv * (tt * tt * phi2 + b * b * oma2 * w)) * EQB_INV_LN10;
  }
  return 0.0;
}

// phase A helper: SN contractions of one genotype row against the SN residualised expression rows of the gene
// (one warp; 16-byte loads -- rows are ldn*8 bytes apart with ldn a multiple of 16)
template <int SN>
__device__ __forceinline__ void contract_shared_x(const double *__restrict__ Xm, const FastSub *__restrict__ fsub, size_t grow,
                                                  int ldn, int lane, double *__restrict__ out)
{
  const double2 *x2 = reinterpret_cast<const double2 *>(Xm);
  const double2 *y2[SN];
#pragma unroll
  for (int a = 0; a < SN; ++a) y2[a] = reinterpret_cast<const double2 *>(fsub[a].Ytil + grow);
  double acc[SN];
#pragma unroll
  for (int a = 0; a < SN; ++a) acc[a] = 0.0;
  const int h = ldn >> 1;
#pragma unroll(SN <= 4 ? 2 : 1)
  for (int i = lane; i < h; i += 32) {
    const double2 x = x2[i];
#pragma unroll
    for (int a = 0; a < SN; ++a) {
      const double2 y = y2[a][i];
      acc[a] += x.x * y.x;
      acc[a] += x.y * y.y;
    }
  }
#pragma unroll
  for (int a = 0; a < SN; ++a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[a] += __shfl_xor_sync(0xffffffffu, acc[a], o);
    if (lane == a) out[a] = acc[a];
  }
}

// Tile of T pairs (T a power of two) per CTA, phase-synchronous.  Work items of the ABF phases are numbered
// (value, pair) with the PAIR index fastest, so that a warp evaluates ONE grid point / configuration for 32 pairs:
// the data-independent branches (oma2 == 0 of the gen-fix row, phi2 == 0 of the gen-maxh row, consistent vs
// singleton value) are warp-uniform.  Values are staged in shared memory ([value][T+1], conflict-free both ways),
// reduced there, and written to the output rows with coalesced stores.
#ifndef EQB_FAST_MINB
#define EQB_FAST_MINB 3 // resident CTAs per SM the register allocation is capped for
#endif
static __global__ void __launch_bounds__(THREADS, EQB_FAST_MINB) fast_pair_kernel(const DevParams *__restrict__ prm_,
                                                            const FastParams *__restrict__ fp_, const FastArgs fa,
                                                            const GridTab gt)
{
  const DevParams &prm = *prm_;
  extern __shared__ double fsm[];
  const int S = prm.S, ldn = prm.ldn, L = prm.L, K = prm.K, T = fa.T, UL = gt.UL;
  const int lgT = 31 - __clz(T), T1 = T + 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long C = (fa.which == 1) ? 0 : ((fa.which == 2) ? S : prm.C);
  const bool join = prm.analysis == 1;
  const int nrow_small = 3 + ((fa.which == 2) ? S : 0); // rows whose values are staged in shared memory
  const int vals_per_pair = 3 * L + ((fa.which == 2) ? S * K : 0);
  const int sst = (3 * S) | 1, sag = (3 * UL) | 1; // odd strides: conflict-free when the pair index varies across lanes
  // shared memory carve-up
  double *xy = fsm;                                          // [T][S]
  double *st = xy + (size_t)T * S;                           // [T][sst]  b, v, t per subgroup
  double *agg = st + (size_t)T * sst;                        // [T][sag]  den, num, sing per unique phi2
  double *wrow = agg + (size_t)T * sag;                      // [T][3+S] weighted small rows
  double *vs = wrow + (size_t)T * (3 + S);                   // [vals_per_pair][T+1] staged values
  double *tab = vs + (size_t)vals_per_pair * T1;             // [T][K][S][3] (which == 3)
  unsigned long long *hasm = (unsigned long long *)(tab + ((fa.which == 3) ? (size_t)T * K * S * 3 : 0)); // [T]
  long long *s_pair = (long long *)(hasm + T);               // [T] output pair index
  long long *s_m = s_pair + T;                               // [T] SNP index
  int *s_gene = (int *)(s_m + T);                            // [T] gene id

  const long long q0 = fa.q_begin + (long long)blockIdx.x * T;
  const int tn = (int)min((long long)T, fa.n_pairs - q0);
  // map the tile's pairs to (gene, SNP, output index)
  for (int j = threadIdx.x; j < tn; j += THREADS) {
    const long long q = q0 + j;
    int lo = fa.tile_gene[blockIdx.x]; // gene of the tile's first pair (host-computed); walk forward from it
    while (lo + 1 < fa.n_genes && fa.fast_base[lo + 1] <= q) ++lo;
    const int g = fa.genes[lo];
    const long long off = q - fa.fast_base[lo];
    s_gene[j] = g;
    s_m[j] = prm.cis_begin[g] + off;
    s_pair[j] = fa.pair_off[lo] + off;
    hasm[j] = 0ull;
  }
  __syncthreads();
  // ---------------- phase A: contraction x . ytil_s (one warp per pair)
  // the genotype rows of the tile are requested from HBM up front (one L2 prefetch per 128-byte line), so that
  // the per-pair loads below pay an L2 hit instead of a DRAM round trip each
  for (int s0 = 0; s0 < S; s0 += 8) {
    if (s0 > 0 && prm.sub[s0].X == prm.sub[0].X) continue;
    for (int j = warp; j < tn; j += WARPS) {
      const double *row = prm.sub[s0].X + (size_t)s_m[j] * ldn;
      for (int l = lane; l < (ldn >> 4); l += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + 16 * l));
    }
  }
  for (int s0 = 0; s0 < S; s0 += 8) {
    const int sn = min(8, S - s0);
    bool same = true;
    for (int a = 1; a < sn; ++a) same = same && (prm.sub[s0 + a].X == prm.sub[s0].X);
    const FastSub *fsub = fp_->sub + s0;
    for (int j = warp; j < tn; j += WARPS) {
      const size_t xoff = (size_t)s_m[j] * ldn, grow = (size_t)s_gene[j] * ldn;
      double *out = xy + (size_t)j * S + s0;
      if (same) {
        const double *Xm = prm.sub[s0].X + xoff;
        switch (sn) {
        case 1: contract_shared_x<1>(Xm, fsub, grow, ldn, lane, out); break;
        case 2: contract_shared_x<2>(Xm, fsub, grow, ldn, lane, out); break;
        case 3: contract_shared_x<3>(Xm, fsub, grow, ldn, lane, out); break;
        case 4: contract_shared_x<4>(Xm, fsub, grow, ldn, lane, out); break;
        case 5: contract_shared_x<5>(Xm, fsub, grow, ldn, lane, out); break;
        case 6: contract_shared_x<6>(Xm, fsub, grow, ldn, lane, out); break;
        case 7: contract_shared_x<7>(Xm, fsub, grow, ldn, lane, out); break;
        default: contract_shared_x<8>(Xm, fsub, grow, ldn, lane, out); break;
        }
      } else {
        for (int a = 0; a < sn; ++a) contract_shared_x<1>(prm.sub[s0 + a].X + xoff, fsub + a, grow, ldn, lane, out + a);
      }
    }
  }
  __syncthreads();
  // ---------------- phase B: thread per (pair, subgroup): summary statistics + standardisation
  for (int it = threadIdx.x; it < tn * S; it += THREADS) {
    const int j = it / S, s = it % S;
    const long long m = s_m[j];
    const int g = s_gene[j];
    const SubDev &sb = prm.sub[s];
    const FastSub &fs = fp_->sub[s];
    const double *ys = fs.ystat + (size_t)g * 4;
    const bool have = sb.gene_has[g] && sb.snp_has[m] && fs.n > 0;
    PairStat ps;
    ps.pve = ps.sigmahat = ps.betahat = ps.se = ps.pval = nan("");
    ps.b = ps.v = ps.t = nan("");
    if (have) {
      const double *xs = fs.xstat + (size_t)m * 3;
      stats_from_dots(xy[(size_t)j * S + s], xs[0], xs[1], xs[2], ys[0], ys[1], ys[2], fs.n, sb.Q, fs.rankz, fs.tz, fs.tz_nu,
                      fs.tz_wmax, ps);
      atomicOr(&hasm[j], 1ull << s);
    }
    st[(size_t)j * sst + s] = ps.b;
    st[(size_t)j * sst + S + s] = ps.v;
    st[(size_t)j * sst + 2 * S + s] = ps.t;
    const long long pair = s_pair[j];
    if (fa.out_n) fa.out_n[pair * S + s] = have ? fs.n : 0;
    if (fa.out_ss) {
      double *o = fa.out_ss + (pair * S + s) * 5;
      o[0] = ps.pve;
      o[1] = ps.sigmahat;
      o[2] = ps.betahat;
      o[3] = ps.se;
      o[4] = ps.pval;
    }
  }
  __syncthreads();
  if (!join) return;
  // ---------------- phase C0: sums over the subgroups with results, per unique phi2 (consistent configuration)
  for (int it = threadIdx.x; it < (UL << lgT); it += THREADS) {
    const int u = it >> lgT, j = it & (T - 1);
    if (j >= tn) continue;
    const double *stj = st + (size_t)j * sst;
    unsigned long long mask = hasm[j];
    const double phi2 = gt.uphi[u];
    double den, num, sing;
    consistent_sums(stj, S, mask, phi2, den, num, sing); // ONE logarithm per (pair, phi2)
    double *a = agg + (size_t)j * sag + 3 * u;
    a[0] = den;
    a[1] = num;
    a[2] = sing;
  }
  if (fa.which == 3) {
    for (int it = threadIdx.x; it < tn * K * S; it += THREADS) {
      const int j = it / (K * S), e = it % (K * S), k = e / S, s = e % S;
      const double *stj = st + (size_t)j * sst;
      double *te = tab + ((size_t)j * K * S + e) * 3;
      if ((hasm[j] >> s) & 1ull)
        term_entry(stj[s], stj[S + s], stj[2 * S + s], prm.phi2S[k], te[0], te[1], te[2]);
      else {
        te[0] = 0.0;
        te[1] = 0.0;
        te[2] = 0.0;
      }
    }
  }
  __syncthreads();
  // ---------------- phase C1: thread per (value, pair): the 3L consistent values (+ S*K singleton values)
  for (int it = threadIdx.x; it < (vals_per_pair << lgT); it += THREADS) {
    const int e = it >> lgT, j = it & (T - 1);
    if (j >= tn) continue;
    double v;
    if (e < 3 * L) { // warp-uniform
      const double *a = agg + (size_t)j * sag + 3 * gt.idxL[e];
      v = abf_from_sums(a[0], a[1], a[2], gt.omaL[e]);
    } else {
      const int e2 = e - 3 * L, c = e2 / K, k = e2 - c * K;
      const double *stj = st + (size_t)j * sst;
      v = 0.0;
      if ((hasm[j] >> c) & 1ull) {
        v = singleton_value(stj[c], stj[S + c], stj[2 * S + c], prm.phi2S[k], prm.oma2S[k]);
      }
    }
    vs[(size_t)e * T1 + j] = v;
  }
  __syncthreads();
  // ---------------- phase C2: log10_weighted_sum of each staged row (utils_math.cpp:100-131)
  for (int it = threadIdx.x; it < (nrow_small << lgT); it += THREADS) {
    const int r = it >> lgT, j = it & (T - 1);
    if (j >= tn) continue;
    const int nk = (r < 3) ? L : K;
    const double *v = vs + (size_t)((r < 3) ? r * L : 3 * L + (r - 3) * K) * T1 + j;
    // two passes exactly as the reference: max (seeded with element 0), then the weighted sum of 10^(x - max)
    double w = nan("");
    if (nk > 0) {
      double mx = v[0];
      for (int k = 1; k < nk; ++k) {
        const double x = v[(size_t)k * T1];
        mx = (x > mx) ? x : mx;
      }
      double sum = 0.0;
      const double wk = 1.0 / (double)nk;
      for (int k = 0; k < nk; ++k) {
        const double x = v[(size_t)k * T1];
        const double e = exp10_fast(x - mx);
        sum += isnan(x) ? 0.0 : wk * e;
      }
      w = mx + log10(sum);
      if (fabs(w) <= DBL_EPSILON) w = 0.0;
    }
    wrow[(size_t)j * (3 + S) + r] = w;
    double *o = fa.out_w + s_pair[j] * (5 + C);
    if (r < 3) o[r] = w;
    else o[5 + (r - 3)] = w;
  }
  // coalesced copy of the staged values to the output rows (warp per pair)
  for (int j = warp; j < tn; j += WARPS) {
    const long long pair = s_pair[j];
    double *og = fa.out_gen + pair * 3 * L;
    for (int e = lane; e < 3 * L; e += 32) og[e] = vs[(size_t)e * T1 + j];
    if (fa.which == 2) {
      double *oc = fa.out_cfg + pair * C * K;
      for (int e = lane; e < S * K; e += 32) oc[e] = vs[(size_t)(3 * L + e) * T1 + j];
    }
  }
  if (fa.which == 3) {
    // all configurations: thread per (pair, configuration), values from the shared table
    for (long long it = threadIdx.x; it < (long long)tn * C; it += THREADS) {
      const int j = (int)(it / C);
      const long long c = it % C;
      const long long pair = s_pair[j];
      const unsigned long long mask = prm.cfg_mask[c] & hasm[j];
      Lse a;
      a.init();
      for (int k = 0; k < K; ++k) {
        const double *tk = tab + ((size_t)j * K + k) * S * 3;
        double den = 0.0, num = 0.0, sing = 0.0;
        unsigned long long mm = mask;
        while (mm) {
          const int s = __ffsll((long long)mm) - 1;
          mm &= mm - 1;
          den += tk[3 * s];
          num += tk[3 * s + 1];
          sing += tk[3 * s + 2];
        }
        const double v = abf_from_sums(den, num, sing, prm.oma2S[k]);
        if (fa.out_cfg) fa.out_cfg[(pair * C + c) * K + k] = v;
        a.add(v, 1.0 / (double)K, k == 0);
      }
      fa.out_w[pair * (5 + C) + 5 + c] = a.result();
    }
  }
  __syncthreads();
  // ---------------- phase C3: BMAlite / BMA (gene_snp_pair.cpp:552-602)
  if (fa.which == 1) {
    for (int j = threadIdx.x; j < tn; j += THREADS) {
      double *o = fa.out_w + s_pair[j] * (5 + C);
      o[3] = nan("");
      o[4] = nan("");
    }
    return;
  }
  if (fa.which == 2) {
    // singletons only: one thread per pair walks its S+1 terms (CalcBMAlite, gene_snp_pair.cpp:552-570)
    for (int j = threadIdx.x; j < tn; j += THREADS) {
      Lse lite;
      lite.init();
      for (int c = 0; c < S; ++c) lite.add(wrow[(size_t)j * (3 + S) + 3 + c], 0.5 / (double)S, c == 0);
      lite.add(wrow[(size_t)j * (3 + S) + 0], 0.5, false);
      double *o = fa.out_w + s_pair[j] * (5 + C);
      o[3] = lite.result();
      o[4] = nan("");
    }
    return;
  }
  for (int j = warp; j < tn; j += WARPS) {
    const long long pair = s_pair[j];
    const double *wcfg = fa.out_w + pair * (5 + C) + 5;
    Lse lite, bma;
    lite.init();
    bma.init();
    for (long long c = lane; c < C; c += 32) {
      const double wc = wcfg[c];
      if (c < S) lite.add(wc, 0.5 / (double)S, c == 0);
      bma.add(wc, prm.cfg_weight[c], c == 0);
    }
    lite = warp_merge(lite);
    lite.add(wrow[(size_t)j * (3 + S) + 0], 0.5, false);
    const double w_gensin = lite.result();
    bma = warp_merge(bma);
    const double w_all = bma.result();
    if (lane == 0) {
      double *o = fa.out_w + pair * (5 + C);
      o[3] = w_gensin;
      o[4] = w_all;
    }
  }
}

// shared-memory bytes of fast_pair_kernel for a tile of T pairs
__host__ __device__ inline size_t fast_smem_bytes(int T, int S, int L, int K, int UL, int which)
{
  const size_t vals = (size_t)3 * L + ((which == 2) ? (size_t)S * K : 0);
  size_t d = (size_t)T * S + (size_t)T * ((3 * S) | 1) + (size_t)T * ((3 * UL) | 1) + (size_t)T * (3 + S) + vals * (T + 1);
  if (which == 3) d += (size_t)T * K * S * 3;
  return d * 8 + (size_t)T * (8 + 8 + 8 + 4) + 16;
}


// =====================================================================================================
// fast_pair_warp_kernel (K2+K3 for --bfs gen|sin and --analys sep): the same three phases as fast_pair_kernel, but
// every WARP owns a tile of up to 32 pairs from the contraction to the last output row.  There is no barrier and no
// per-CTA prologue: the memory-bound contraction of one warp overlaps the transcendental-bound ABF phase of its
// neighbours (fast_pair_kernel's phases are CTA-synchronous: barrier stalls were a third of its issue slots and its
// thread-per-(pair, subgroup) phase B used 96 of 256 threads).
//   A  mma.sync.m8n8k4 (DMMA) tile product  xy[pairs][S] = X_tile . Ytil_gene^T            (contraction)
//   B  lane per (pair, subgroup): summary statistics + standardisation                  (S rounds of 32 items)
//   C  lane per PAIR: loop over the unique phi2 values (the sums over the subgroups are computed once per
//      phi2 and reused by every grid point of gen / gen-fix / gen-maxh that shares it), then over the singleton
//      configurations; values go straight to their output rows, log10_weighted_sum is accumulated online in
//      registers (utils_math.cpp:100-131; same NaN rules), BMAlite at the end.  The grid point / configuration
//      is uniform across the warp (the pair varies across lanes), so grid-dependent branches never diverge.
// The grid tables of phase C are warp-uniform reads: they travel as a kernel parameter (constant bank, GridConst)
// when they fit, else they are read from global memory.  Tiles come from a host-built list (tile_q0, <= 32 pairs).
// Measured on the c2 step (EQB_FASTW_DEBUG): phase B alone 0.10 ms, A + B 0.26 ms, everything 0.58 ms, of which the
// uncoalesced raw-value stores 0.05 ms -- the phases add up, i.e. the warps of a wave move through them together.
// Neither occupancy (64 / 80 / 128 registers: 0.57 / 0.61 / 0.56 ms) nor persistent warps that request their next
// tile's genotype rows from HBM before phase C (0.65 ms) changed that.
// Shared memory per warp: xy[32][S] + (b, v, t)[32][3S|1] + masks and pair indices.
struct GridOrder {
  const int *ustart; // [UL+1] entries of unique phi2 value u: uent[ustart[u] .. ustart[u+1])
  const int *uent;   // [3L]   r * L + k of the entry (row r of gen / gen-fix / gen-maxh, grid point k)
};

// the same tables by value (kernel parameter -> constant bank): limits UL <= 64, 3L <= 192, K <= 32
constexpr int GC_UL = 64, GC_3L = 192, GC_K = 32;
struct GridConst {
  double uphi[GC_UL];
  double oma[GC_3L]; // omega2 of entry i (grouped order)
  double phiS[GC_K], omaS[GC_K];
  short ustart[GC_UL + 1];
  unsigned char ent[GC_3L]; // r * L + k of entry i
};

struct LseOnline { // log10_weighted_sum accumulated one element at a time (max tracked with rescaling)
  double m, acc;
  bool poisoned; // element 0 was NaN: the reference's max is NaN and so is its result
  __device__ __forceinline__ void init()
  {
    m = -INFINITY;
    acc = 0.0;
    poisoned = false;
  }
  __device__ __forceinline__ void add(double v, double w, bool is_first)
  {
    if (v != v) {
      poisoned = poisoned || is_first;
      return;
    }
    const double d = v - m; // +inf on the first element
    const bool up = d > 0.0;
    const double e = exp10_fast(up ? -d : d);
    acc = up ? fma(acc, e, w) : fma(w, e, acc);
    m = up ? v : m;
  }
  __device__ __forceinline__ double result() const
  {
    if (poisoned) return nan("");
    double r = m + log10(acc);
    if (fabs(r) <= DBL_EPSILON) r = 0.0;
    return r;
  }
};

// phase A helper: two genotype rows (two pairs) against their SN residual phenotype rows, one warp
template <int SN>
__device__ __forceinline__ void contract_two(const double *__restrict__ Xa, const double *__restrict__ Xb,
                                             const FastSub *__restrict__ fsub, size_t growa, size_t growb, int ldn, int lane,
                                             double *__restrict__ outa, double *__restrict__ outb)
{
  const double2 *xa2 = reinterpret_cast<const double2 *>(Xa), *xb2 = reinterpret_cast<const double2 *>(Xb);
  double acca[SN], accb[SN];
#pragma unroll
  for (int a = 0; a < SN; ++a) acca[a] = accb[a] = 0.0;
  const int h = ldn >> 1;
  for (int i = lane; i < h; i += 32) {
    const double2 xa = xa2[i], xb = xb2[i];
#pragma unroll
    for (int a = 0; a < SN; ++a) {
      const double2 ya = reinterpret_cast<const double2 *>(fsub[a].Ytil + growa)[i];
      const double2 yb = reinterpret_cast<const double2 *>(fsub[a].Ytil + growb)[i];
      acca[a] = fma(xa.x, ya.x, acca[a]);
      accb[a] = fma(xb.x, yb.x, accb[a]);
      acca[a] = fma(xa.y, ya.y, acca[a]);
      accb[a] = fma(xb.y, yb.y, accb[a]);
    }
  }
#pragma unroll
  for (int a = 0; a < SN; ++a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      acca[a] += __shfl_xor_sync(0xffffffffu, acca[a], o);
      accb[a] += __shfl_xor_sync(0xffffffffu, accb[a], o);
    }
    if (lane == a) {
      outa[a] = acca[a];
      outb[a] = accb[a];
    }
  }
}

template <int SN>
__device__ __forceinline__ void contract_tile(const double *__restrict__ X, const FastSub *__restrict__ fsub, const long long *s_m,
                                              const int *s_gene, int tn, int S, int ldn, int lane, double *__restrict__ xy)
{
  int j = 0;
  for (; j + 1 < tn; j += 2)
    contract_two<SN>(X + (size_t)s_m[j] * ldn, X + (size_t)s_m[j + 1] * ldn, fsub, (size_t)s_gene[j] * ldn,
                     (size_t)s_gene[j + 1] * ldn, ldn, lane, xy + (size_t)j * S, xy + (size_t)(j + 1) * S);
  if (j < tn) contract_shared_x<SN>(X + (size_t)s_m[j] * ldn, fsub, (size_t)s_gene[j] * ldn, ldn, lane, xy + (size_t)j * S);
}

// phase A on the FP64 tensor cores: the tile's contraction IS a small matrix product
//   xy[32 pairs][sn subgroups] = X_tile[32][ldn] . Ytil_gene[sn][ldn]^T
// done as 4 row blocks of mma.sync.m8n8k4 (DMMA): A fragments straight from the genotype rows (one 16-byte load per
// lane covers two k-steps: the k index inside a chunk of 8 individuals is permuted the same way for A and B),
// B fragments from the gene's residual phenotype rows (L1/L2 hits: 7 KB per gene; columns >= sn repeat the last
// subgroup and are dropped).  The rows of a tile that belong to different genes are handled as runs: every run
// multiplies the row blocks it touches by ITS gene's phenotype block and keeps its own rows (1.6 runs per tile at
// 50 SNPs per gene, 1 at the GTEx shape).  ~600 warp instructions per tile instead of ~8000 for shuffle-reduced
// dot products.  (Measured dead ends: a cp.async ring for the fragments, 0.65 vs 0.58 ms; a CTA-wide shared-memory
// copy of the phenotype rows, same time as the L1 path but it needs a barrier and a per-CTA prologue.)
__device__ __forceinline__ void contract_tile_dmma(const double *__restrict__ X, const FastSub *__restrict__ fsub, int sn,
                                                   const long long *s_m, const int *s_gene, int tn, int S, int ldn, int lane,
                                                   double *__restrict__ xy)
{
  const int r = lane >> 2, kq = lane & 3;
  const double *xp[4];
#pragma unroll
  for (int mb = 0; mb < 4; ++mb) {
    const int j = min(mb * 8 + r, tn - 1); // rows past the end of the tile repeat the last one (results dropped)
    xp[mb] = X + (size_t)s_m[j] * ldn + 2 * kq;
  }
  const int col = min(r, sn - 1);
  const int nchunk = ldn >> 3;
  int j0 = 0;
  while (j0 < tn) {
    const int g = s_gene[j0];
    const unsigned same = __ballot_sync(0xffffffffu, lane < tn && s_gene[lane < tn ? lane : 0] == g);
    const int j1 = j0 + __popc(same >> j0 << j0); // rows of a gene are consecutive
    const double *yp = fsub[col].Ytil + (size_t)g * ldn + 2 * kq;
    double acc[4][2];
#pragma unroll
    for (int mb = 0; mb < 4; ++mb) acc[mb][0] = acc[mb][1] = 0.0;
    const int mb0 = j0 >> 3, mb1 = (j1 - 1) >> 3; // only the row blocks that hold rows of this run (warp-uniform)
#pragma unroll 2
    for (int c = 0; c < nchunk; ++c) {
      const double2 b2 = *reinterpret_cast<const double2 *>(yp + 8 * c);
#pragma unroll
      for (int mb = 0; mb < 4; ++mb) {
        if (mb >= mb0 && mb <= mb1) {
          const double2 a2 = *reinterpret_cast<const double2 *>(xp[mb] + 8 * c);
          dmma_m8n8k4(acc[mb][0], acc[mb][1], a2.x, b2.x);
          dmma_m8n8k4(acc[mb][0], acc[mb][1], a2.y, b2.y);
        }
      }
    }
#pragma unroll
    for (int mb = 0; mb < 4; ++mb) {
      const int j = mb * 8 + r;
      if (j >= j0 && j < j1) {
        if (2 * kq < sn) xy[(size_t)j * S + 2 * kq] = acc[mb][0];
        if (2 * kq + 1 < sn) xy[(size_t)j * S + 2 * kq + 1] = acc[mb][1];
      }
    }
    j0 = j1;
  }
}

// The same product with the A fragments taken from the resident integer numerators: a lane loads the two u16 of its
// k-steps (4 bytes instead of 16) and looks the doubles up in k2v (512 KB, the handful of entries a dosage file uses stay in
// L1): the HBM bytes of the contraction drop 4x, the fragment VALUES are the doubles of the f64 matrix, bit for bit.
__device__ __forceinline__ void contract_tile_dmma_u16(const unsigned short *__restrict__ X16, const double *__restrict__ k2v,
                                                       const FastSub *__restrict__ fsub, int sn, const long long *s_m,
                                                       const int *s_gene, int tn, int S, int ldn, int lane,
                                                       double *__restrict__ xy)
{
  const int r = lane >> 2, kq = lane & 3;
  const unsigned short *xp[4];
#pragma unroll
  for (int mb = 0; mb < 4; ++mb) {
    const int j = min(mb * 8 + r, tn - 1); // rows past the end of the tile repeat the last one (results dropped)
    xp[mb] = X16 + (size_t)s_m[j] * ldn + 2 * kq;
  }
  const int col = min(r, sn - 1);
  const int nchunk = ldn >> 3;
  int j0 = 0;
  while (j0 < tn) {
    const int g = s_gene[j0];
    const unsigned same = __ballot_sync(0xffffffffu, lane < tn && s_gene[lane < tn ? lane : 0] == g);
    const int j1 = j0 + __popc(same >> j0 << j0); // rows of a gene are consecutive
    const double *yp = fsub[col].Ytil + (size_t)g * ldn + 2 * kq;
    double acc[4][2];
#pragma unroll
    for (int mb = 0; mb < 4; ++mb) acc[mb][0] = acc[mb][1] = 0.0;
    const int mb0 = j0 >> 3, mb1 = (j1 - 1) >> 3; // only the row blocks that hold rows of this run (warp-uniform)
#pragma unroll 4
    for (int c = 0; c < nchunk; ++c) {
      const double2 b2 = *reinterpret_cast<const double2 *>(yp + 8 * c);
#pragma unroll
      for (int mb = 0; mb < 4; ++mb) {
        if (mb >= mb0 && mb <= mb1) {
          const unsigned int w = *reinterpret_cast<const unsigned int *>(xp[mb] + 8 * c);
          const double ax = __ldg(k2v + (w & 0xffffu)), ay = __ldg(k2v + (w >> 16));
          dmma_m8n8k4(acc[mb][0], acc[mb][1], ax, b2.x);
          dmma_m8n8k4(acc[mb][0], acc[mb][1], ay, b2.y);
        }
      }
    }
#pragma unroll
    for (int mb = 0; mb < 4; ++mb) {
      const int j = mb * 8 + r;
      if (j >= j0 && j < j1) {
        if (2 * kq < sn) xy[(size_t)j * S + 2 * kq] = acc[mb][0];
        if (2 * kq + 1 < sn) xy[(size_t)j * S + 2 * kq + 1] = acc[mb][1];
      }
    }
    j0 = j1;
  }
}

// shared memory of fast_pair_warp_kernel: one region per warp
__host__ __device__ inline size_t fast_warp_smem_bytes(int S)
{
  return ((size_t)32 * S + (size_t)32 * ((3 * S) | 1)) * 8 + (size_t)32 * (8 + 8 + 8 + 4);
}

#ifndef EQB_FASTW_MINB
#define EQB_FASTW_MINB 2
#endif
// TP: grid tables from the GridConst kernel parameter (constant bank) instead of global memory.
// DM: every group of 8 subgroups shares one genotype matrix and phase A runs on the tensor cores (the common case);
//     the instantiation without DM carries the shuffle-reduced fallbacks, whose register needs would otherwise
//     set the allocation (and the occupancy) of the common case too.
// X16: the contraction reads the resident u16 numerators (fa.x16 / fa.k2v); its own instantiation so that the f64 kernel keeps
//      its register allocation and schedule (with both loops in one kernel the f64 path ran 4 % slower).
#ifndef EQB_FASTW_MINB_X16
#define EQB_FASTW_MINB_X16 EQB_FASTW_MINB
#endif
template <bool TP, bool DM, bool X16 = false>
__global__ void __launch_bounds__(THREADS, X16 ? EQB_FASTW_MINB_X16 : (DM ? EQB_FASTW_MINB : 2)) fast_pair_warp_kernel(const DevParams *__restrict__ prm_,
                                                                 const FastParams *__restrict__ fp_, const FastArgs fa,
                                                                 const GridTab gt, const GridOrder go,
                                                                 const __grid_constant__ GridConst gc)
{
  const DevParams &prm = *prm_;
  extern __shared__ double fsm[];
  const int S = prm.S, ldn = prm.ldn, L = prm.L, K = prm.K, UL = gt.UL;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  __shared__ BfTabs Tsm; // exp / log tables of phase C: the only CTA-wide step of the kernel
  bf_tabs_init(Tsm);
  __syncthreads();
  TabRef T;
  T.base = smem_u32(&Tsm);
  const long long tile = (long long)blockIdx.x * nwarp + warp;
  if (tile >= fa.n_tiles) return; // (no barrier anywhere below)
#ifdef EQB_TUNING
  if (fa.delay_ns > 0 && (int)blockIdx.x < fa.delay_ctas) {
    // timing experiment (EQB_FASTW_DELAY_US): CTAs of the first wave that share an SM start out of phase
    const long long t_end = clock64() + (long long)(blockIdx.x / fa.delay_sm) * fa.delay_ns * 2; // ~2 cycles per ns
    while (clock64() < t_end) __nanosleep(2000);
  }
  const int dbg = fa.debug;
#else
  constexpr int dbg = 0; // the phase-skipping switches of the timing experiments do not exist in the product build
#endif
  const long long q0 = fa.tile_q0[tile];
  const int tn = (int)(fa.tile_q0[tile + 1] - q0); // 1 .. 32
  const long long C = (fa.which == 1) ? 0 : (fa.which == 3 ? prm.C : S); // configurations in a row of out_w
  const bool join = prm.analysis == 1;
  const int sst = (3 * S) | 1;
  // per-warp shared memory
  char *wbase = reinterpret_cast<char *>(fsm) + (size_t)warp * fast_warp_smem_bytes(S);
  double *xy = reinterpret_cast<double *>(wbase);            // [32][S]
  double *st = xy + (size_t)32 * S;                          // [32][sst]  b, v, t per subgroup
  unsigned long long *hasm = (unsigned long long *)(st + (size_t)32 * sst); // [32]
  long long *s_pair = (long long *)(hasm + 32);              // [32] output pair index
  long long *s_m = s_pair + 32;                              // [32] SNP index
  int *s_gene = (int *)(s_m + 32);                           // [32] gene id

  long long my_pair = 0;
  if (fa.from_st) {
    if (lane < tn) {
      my_pair = q0 + lane;
      s_pair[lane] = my_pair;
      hasm[lane] = fa.has_all[q0 + lane];
    }
    for (int i = lane; i < tn * 3 * S; i += 32) {
      const int j = i / (3 * S);
      st[(size_t)j * sst + (i - j * 3 * S)] = fa.st_all[q0 * 3 * S + i];
    }
    __syncwarp();
  } else {
  if (lane < tn) {
    const long long q = q0 + lane;
    int lo = fa.tile_gene[tile]; // gene of the tile's first pair (host-computed); walk forward from it
    while (lo + 1 < fa.n_genes && fa.fast_base[lo + 1] <= q) ++lo;
    const int g = fa.genes[lo];
    const long long off = q - fa.fast_base[lo];
    s_gene[lane] = g;
    s_m[lane] = prm.cis_begin[g] + off;
    my_pair = fa.pair_off[lo] + off;
    s_pair[lane] = my_pair;
    hasm[lane] = 0ull;
  }
  __syncwarp();
  // ---------------- phase A: contraction x . ytil_s
  // the genotype rows of the tile are requested from HBM up front (one L2 prefetch per 128-byte line)
  if (DM && X16) {
    for (int j = 0; j < tn; ++j) {
      const unsigned short *row = fa.x16 + (size_t)s_m[j] * ldn;
      for (int l = lane; l < ((ldn + 63) >> 6); l += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + 64 * l));
    }
  } else
  for (int s0 = 0; s0 < S; s0 += 8) {
    if (s0 > 0 && prm.sub[s0].X == prm.sub[0].X) continue;
    for (int j = 0; j < tn; ++j) {
      const double *row = prm.sub[s0].X + (size_t)s_m[j] * ldn;
      for (int l = lane; l < (ldn >> 4); l += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + 16 * l));
    }
  }
  for (int s0 = 0; s0 < S; s0 += 8) {
    const int sn = min(8, S - s0);
    bool same = true;
    for (int a = 1; a < sn; ++a) same = same && (prm.sub[s0 + a].X == prm.sub[s0].X);
    const FastSub *fsub = fp_->sub + s0;
    if (dbg & 2) {
      for (int i = lane; i < tn * sn; i += 32) xy[(size_t)(i / sn) * S + s0 + i % sn] = 0.1;
    } else if (DM) {
      if (X16) contract_tile_dmma_u16(fa.x16, fa.k2v, fsub, sn, s_m, s_gene, tn, S, ldn, lane, xy + s0);
      else contract_tile_dmma(prm.sub[s0].X, fsub, sn, s_m, s_gene, tn, S, ldn, lane, xy + s0);
    } else if (same) {
      const double *X = prm.sub[s0].X;
      switch (sn) {
      case 1: contract_tile<1>(X, fsub, s_m, s_gene, tn, S, ldn, lane, xy + s0); break;
      case 2: contract_tile<2>(X, fsub, s_m, s_gene, tn, S, ldn, lane, xy + s0); break;
      case 3: contract_tile<3>(X, fsub, s_m, s_gene, tn, S, ldn, lane, xy + s0); break;
      case 4: contract_tile<4>(X, fsub, s_m, s_gene, tn, S, ldn, lane, xy + s0); break;
      case 5: contract_tile<5>(X, fsub, s_m, s_gene, tn, S, ldn, lane, xy + s0); break;
      case 6: contract_tile<6>(X, fsub, s_m, s_gene, tn, S, ldn, lane, xy + s0); break;
      case 7: contract_tile<7>(X, fsub, s_m, s_gene, tn, S, ldn, lane, xy + s0); break;
      default: contract_tile<8>(X, fsub, s_m, s_gene, tn, S, ldn, lane, xy + s0); break;
      }
    } else {
      for (int a = 0; a < sn; ++a) contract_tile<1>(prm.sub[s0 + a].X, fsub + a, s_m, s_gene, tn, S, ldn, lane, xy + s0 + a);
    }
  }
  __syncwarp();
  // ---------------- phase B: lane per (pair, subgroup): summary statistics + standardisation
  for (int it = lane; it < tn * S; it += 32) {
    const int j = it / S, s = it - j * S;
    const long long m = s_m[j];
    const int g = s_gene[j];
    const SubDev &sb = prm.sub[s];
    const FastSub &fs = fp_->sub[s];
    const double *ys = fs.ystat + (size_t)g * 4;
    const bool have = sb.gene_has[g] && sb.snp_has[m] && fs.n > 0;
    PairStat ps;
    ps.pve = ps.sigmahat = ps.betahat = ps.se = ps.pval = nan("");
    ps.b = ps.v = ps.t = nan("");
    if (have) {
      const double *xs = fs.xstat + (size_t)m * 3;
      stats_from_dots(xy[(size_t)j * S + s], xs[0], xs[1], xs[2], ys[0], ys[1], ys[2], fs.n, sb.Q, fs.rankz, fs.tz, fs.tz_nu,
                      fs.tz_wmax, ps);
      atomicOr(&hasm[j], 1ull << s);
    }
    st[(size_t)j * sst + s] = ps.b;
    st[(size_t)j * sst + S + s] = ps.v;
    st[(size_t)j * sst + 2 * S + s] = ps.t;
    const long long pair = s_pair[j];
    if (fa.out_n) fa.out_n[pair * S + s] = have ? fs.n : 0;
    if (fa.out_ss) {
      double *o = fa.out_ss + (pair * S + s) * 5;
      o[0] = ps.pve;
      o[1] = ps.sigmahat;
      o[2] = ps.betahat;
      o[3] = ps.se;
      o[4] = ps.pval;
    }
  }
  __syncwarp();
  if (fa.st_all) { // --bfs all: the standardised statistics of the tile for the second pass (coalesced)
    for (int i = lane; i < tn * 3 * S; i += 32) {
      const int j = i / (3 * S);
      fa.st_all[q0 * 3 * S + i] = st[(size_t)j * sst + (i - j * 3 * S)];
    }
    if (lane < tn) fa.has_all[q0 + lane] = hasm[lane];
  }
  } // (!fa.from_st)
  if (!join || lane >= tn || (dbg & 4)) return;
  const bool st_raw = !(dbg & 1) && fa.out_gen != nullptr; // raw values only on request
  // ---------------- phase C: lane per pair
  const double *stj = st + (size_t)lane * sst;
  const unsigned long long mask = hasm[lane];
  double *og = fa.out_gen + my_pair * 3 * L;
  double *ow = fa.out_w + my_pair * (5 + C);
  LseTab r0, r1, r2;
  r0.init();
  r1.init();
  r2.init();
  const double wL = 1.0 / (double)L;
  for (int u = 0; u < UL; ++u) {
    double den, num, sing;
    consistent_sums_t(stj, S, mask, TP ? gc.uphi[u] : gt.uphi[u], den, num, sing, T); // ONE logarithm per (pair, phi2)
    const int i0 = TP ? (int)gc.ustart[u] : go.ustart[u], i1 = TP ? (int)gc.ustart[u + 1] : go.ustart[u + 1];
    for (int i = i0; i < i1; ++i) {
      const int e = TP ? (int)gc.ent[i] : go.uent[i]; // warp-uniform
      const double v = abf_from_sums_t(den, num, sing, TP ? gc.oma[i] : gt.omaL[e], T);
      if (st_raw) og[e] = v;
      if (e < L) r0.add(v, wL, e == 0, T);
      else if (e < 2 * L) r1.add(v, wL, e == L, T);
      else r2.add(v, wL, e == 2 * L, T);
    }
  }
  const double wgen = r0.result(T);
  ow[0] = wgen;
  ow[1] = r1.result(T);
  ow[2] = r2.result(T);
  if (fa.which == 1) {
    ow[3] = nan("");
    ow[4] = nan("");
    return;
  }
  if (fa.which == 3) return; // (the configurations and both model averages: fast_pair_all_kernel)
  // singleton configurations (CalcAbfsUvlrForSingletons, gene_snp_pair.cpp:422-463) + BMAlite (:552-570)
  double *oc = fa.out_cfg + my_pair * C * K;
  LseTab lite;
  lite.init();
  const double wK = 1.0 / (double)K, wS = 0.5 / (double)S;
  for (int c = 0; c < S; ++c) {
    LseTab rc;
    rc.init();
    const bool has = (mask >> c) & 1ull;
    const double b = stj[c], vv = stj[S + c], tt = stj[2 * S + c];
    for (int k = 0; k < K; ++k) {
      const double v = has ? singleton_value_t(b, vv, tt, TP ? gc.phiS[k] : prm.phi2S[k], TP ? gc.omaS[k] : prm.oma2S[k], T) : 0.0;
      if (st_raw) oc[c * K + k] = v;
      rc.add(v, wK, k == 0, T);
    }
    const double wc = (K > 0) ? rc.result(T) : nan("");
    ow[5 + c] = wc;
    lite.add(wc, wS, c == 0, T);
  }
  lite.add(wgen, 0.5, false, T);
  ow[3] = lite.result(T);
  ow[4] = nan("");
}

} // namespace eqb
