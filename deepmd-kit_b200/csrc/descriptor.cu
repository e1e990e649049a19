// Descriptor contraction of se_e2_a for sm_100a:  D = (GR/nnei)^T (GR/nnei)[:, :axis]
// (deepmd/pt/model/descriptor/se_a.py:843-850, deepmd/tf/descriptor/se_a.py:1312-1324) and its
// backward.  GR = the [4][M] tabulate output of one atom; D is [M][axis] (M*axis = 1600 for water).
//
// In the reference this is a batched matmul of a [M x 4] by a [4 x axis] matrix per atom (cuBLAS
// batched GEMM with K = 4: launch- and pointer-bound).  It is a pure streaming op: read 4*M, write
// M*axis values per atom (forward), read M*axis + 4*M, write 4*M (backward) -- HBM-bound, so one
// WARP per atom with the operands staged in shared memory and fully coalesced stores.
#include "common.cuh"

namespace dpb200 {
namespace {

extern __shared__ __align__(16) unsigned char desc_smem[];

template <typename FP>
__global__ void __launch_bounds__(128) k_desc_fwd(FP* __restrict__ D, const FP* __restrict__ X, long long nloc, int M,
                                                  int axis, FP scale) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  FP* xs = reinterpret_cast<FP*>(desc_smem) + (size_t)warp * 4 * M;
  const int nout = M * axis;
  for (long long i = (long long)blockIdx.x * 4 + warp; i < nloc; i += (long long)gridDim.x * 4) {
    const FP* __restrict__ x = X + i * 4 * M;
    __syncwarp();
    for (int e = lane; e < 4 * M; e += 32) xs[e] = x[e] * scale;
    __syncwarp();
    FP* __restrict__ d = D + i * nout;
    for (int e = lane; e < nout; e += 32) {
      const int k1 = e / axis, k2 = e - k1 * axis;
      const FP v = xs[k1] * xs[k2] + xs[M + k1] * xs[M + k2] + xs[2 * M + k1] * xs[2 * M + k2] +
                   xs[3 * M + k1] * xs[3 * M + k2];
      st_cs(d + e, v);
    }
  }
}

// dX[m][k] = scale * ( sum_{k2<axis} dD[k][k2] xs[m][k2]  +  [k<axis] sum_{k1<M} dD[k1][k] xs[m][k1] ),
// xs = X*scale
template <typename FP>
__global__ void __launch_bounds__(128) k_desc_bwd(FP* __restrict__ dX, const FP* __restrict__ dD,
                                                  const FP* __restrict__ X, long long nloc, int M, int axis,
                                                  FP scale) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ld = axis + 1;  // padded row of the staged dD: conflict-free column walks
  FP* xs = reinterpret_cast<FP*>(desc_smem) + (size_t)warp * (4 * M + M * ld);
  FP* g = xs + 4 * M;
  const int nout = M * axis;
  for (long long i = (long long)blockIdx.x * 4 + warp; i < nloc; i += (long long)gridDim.x * 4) {
    const FP* __restrict__ x = X + i * 4 * M;
    const FP* __restrict__ gd = dD + i * nout;
    __syncwarp();
    for (int e = lane; e < 4 * M; e += 32) xs[e] = x[e] * scale;
    for (int e = lane; e < nout; e += 32) {
      const int k1 = e / axis, k2 = e - k1 * axis;
      g[k1 * ld + k2] = __ldcs(gd + e);
    }
    __syncwarp();
    FP* __restrict__ o = dX + i * 4 * M;
    for (int k = lane; k < M; k += 32) {
      FP a0 = 0, a1 = 0, a2 = 0, a3 = 0;
      for (int k2 = 0; k2 < axis; ++k2) {
        const FP w = g[k * ld + k2];
        a0 += w * xs[k2];
        a1 += w * xs[M + k2];
        a2 += w * xs[2 * M + k2];
        a3 += w * xs[3 * M + k2];
      }
      if (k < axis) {
        for (int k1 = 0; k1 < M; ++k1) {
          const FP w = g[k1 * ld + k];
          a0 += w * xs[k1];
          a1 += w * xs[M + k1];
          a2 += w * xs[2 * M + k1];
          a3 += w * xs[3 * M + k1];
        }
      }
      o[k] = a0 * scale;
      o[M + k] = a1 * scale;
      o[2 * M + k] = a2 * scale;
      o[3 * M + k] = a3 * scale;
    }
  }
}

template <typename FP>
int desc_launch(bool bwd, FP* out, const FP* dD, const FP* X, long long nloc, int M, int axis, double scale,
                cudaStream_t st) {
  DPB_REQUIRE(nloc >= 0 && M >= 1 && axis >= 1 && axis <= M, "descriptor: need 1 <= axis <= M");
  if (nloc == 0) return DPB200_OK;
  DPB_REQUIRE(out && X && (!bwd || dD), "descriptor: null pointer");
  const size_t smem = (bwd ? (size_t)(4 * M + M * (axis + 1)) : (size_t)4 * M) * sizeof(FP) * 4;
  DPB_REQUIRE(smem <= 200 * 1024, "descriptor: M*axis too large for shared memory staging");
  int occ = 0;
  long long want = (nloc + 3) / 4;
  if (bwd) {
    auto kern = k_desc_bwd<FP>;
    DPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 128, smem));
    long long cap = (long long)sm_count() * (occ < 1 ? 1 : occ);
    kern<<<(int)(want < cap ? want : cap), 128, smem, st>>>(out, dD, X, nloc, M, axis, (FP)scale);
  } else {
    auto kern = k_desc_fwd<FP>;
    DPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 128, smem));
    long long cap = (long long)sm_count() * (occ < 1 ? 1 : occ);
    kern<<<(int)(want < cap ? want : cap), 128, smem, st>>>(out, X, nloc, M, axis, (FP)scale);
  }
  DPB_CUDA(cudaGetLastError());
  note_launches(1);
  return DPB200_OK;
}

}  // namespace
}  // namespace dpb200

extern "C" {
#define DPB200_DEF_DESC(SUF, FP)                                                                        \
  int dpb200_se_a_descriptor_##SUF(FP* D, const FP* gr, long long nloc, int M, int axis, double scale,  \
                                   dpb200_stream_t stream) {                                            \
    return dpb200::desc_launch<FP>(false, D, nullptr, gr, nloc, M, axis, scale, (cudaStream_t)stream);  \
  }                                                                                                     \
  int dpb200_se_a_descriptor_grad_##SUF(FP* dgr, const FP* dD, const FP* gr, long long nloc, int M,     \
                                        int axis, double scale, dpb200_stream_t stream) {               \
    return dpb200::desc_launch<FP>(true, dgr, dD, gr, nloc, M, axis, scale, (cudaStream_t)stream);      \
  }
DPB200_DEF_DESC(f64, double)
DPB200_DEF_DESC(f32, float)
#undef DPB200_DEF_DESC
}
