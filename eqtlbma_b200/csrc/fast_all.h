// fast_all.h -- host interface of the `--bfs all` true-pass kernel (fast_all.cu / fast_all_kernel.cuh), used by
// eqtlbma_b200.cu: limits, shared-memory layout and the launcher (its own translation unit: the kernel is instantiated
// three times and compiles in parallel with the rest of the library).
#pragma once

#include <cuda_runtime.h>

#include <cstddef>

namespace eqb {

constexpr int FA_MAXS = 10;   // S <= 10: three parts of <= 4 + 3 + 3 subgroups
constexpr int FA_MAXK = 16;   // K <= 16: at least two pairs per tile
constexpr int FA_CH = 32;     // configurations per chunk
constexpr int FA_SROW = 33;   // staging row stride (doubles): conflict-free in both access directions
constexpr int FA_MAXPPW = 4;  // pairs per tile
constexpr int FA_WARPS = 4, FA_THREADS = FA_WARPS * 32;

struct FaParts {
  int np[3], start[3], off[3], ne; // subgroups per part, first subgroup, first table entry, entries in total
};
__host__ __device__ inline FaParts fa_parts(int S)
{
  FaParts p;
  p.np[0] = (S + 2) / 3;
  p.np[1] = (S - p.np[0] + 1) / 2;
  p.np[2] = S - p.np[0] - p.np[1];
  p.start[0] = 0;
  p.start[1] = p.np[0];
  p.start[2] = p.np[0] + p.np[1];
  p.off[0] = 0;
  p.off[1] = 1 << p.np[0];
  p.off[2] = p.off[1] + (1 << p.np[1]);
  p.ne = p.off[2] + (1 << p.np[2]);
  return p;
}
__host__ __device__ inline int fa_pairs_per_warp(int K) { return (32 / K) < FA_MAXPPW ? (32 / K) : FA_MAXPPW; }
// shared memory of a CTA: tile arrays (doubles: st[4][sst] hasm[4] s_pair[4] genavg[4] part[FA_WARPS][4][3] mcT[4][32]) |
// configuration masks u16[C] (zero-padded to 16 bytes) | tab[ne][3][32] | stg[FA_WARPS][FA_CH][FA_SROW]
__host__ __device__ inline size_t fa_tile_doubles(int S) { return (size_t)4 * ((3 * S) | 1) + 12 + (size_t)FA_WARPS * 12 + 4 * 32; }
__host__ __device__ inline size_t fa_mask_bytes(long long C) { return (size_t)((C * 2 + 15) / 16) * 16; }
__host__ __device__ inline size_t fast_all_smem_bytes(int S, long long C)
{
  return fa_tile_doubles(S) * 8 + fa_mask_bytes(C) + ((size_t)fa_parts(S).ne * 3 * 32 + (size_t)FA_WARPS * FA_CH * FA_SROW) * 8;
}

struct DevParams;
struct FastParams;
struct FastArgs;
struct GridTab;
struct GridConst;

// launches fast_pair_all_kernel<fa_pairs_per_warp(K)> with `grid` persistent CTAs of FA_THREADS threads on `stream`
cudaError_t launch_fast_pair_all(int K, unsigned grid, size_t smem, cudaStream_t stream, const DevParams *d_prm, const FastParams *d_fp,
                                 const FastArgs &fa, const GridTab &gt, const GridConst &gc);

} // namespace eqb
