// fast_all.cu -- launcher of the `--bfs all` true-pass kernel (fast_all_kernel.cuh).
#include "fast_all.h"

#include "fast_all_kernel.cuh"

namespace eqb {

template <int PPW, bool LIN>
static cudaError_t launch_lin(unsigned grid, size_t smem, cudaStream_t stream, const DevParams *d_prm, const FastParams *d_fp,
                              const FastArgs &fa, const GridTab &gt, const GridConst &gc)
{
  cudaError_t e = cudaFuncSetAttribute(fast_pair_all_kernel<PPW, LIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  fast_pair_all_kernel<PPW, LIN><<<grid, FA_THREADS, smem, stream>>>(d_prm, d_fp, fa, gt, gc);
  return cudaGetLastError();
}

// no raw per-configuration values requested: the linear-domain instantiation
template <int PPW>
static cudaError_t launch_one(unsigned grid, size_t smem, cudaStream_t stream, const DevParams *d_prm, const FastParams *d_fp,
                              const FastArgs &fa, const GridTab &gt, const GridConst &gc)
{
  if (fa.out_cfg == nullptr) return launch_lin<PPW, true>(grid, smem, stream, d_prm, d_fp, fa, gt, gc);
  return launch_lin<PPW, false>(grid, smem, stream, d_prm, d_fp, fa, gt, gc);
}

cudaError_t launch_fast_pair_all(int K, unsigned grid, size_t smem, cudaStream_t stream, const DevParams *d_prm, const FastParams *d_fp,
                                 const FastArgs &fa, const GridTab &gt, const GridConst &gc)
{
  const int ppw = fa_pairs_per_warp(K);
  if (ppw == 2) return launch_one<2>(grid, smem, stream, d_prm, d_fp, fa, gt, gc);
  if (ppw == 3) return launch_one<3>(grid, smem, stream, d_prm, d_fp, fa, gt, gc);
  return launch_one<4>(grid, smem, stream, d_prm, d_fp, fa, gt, gc);
}

} // namespace eqb
