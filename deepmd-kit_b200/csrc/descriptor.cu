// Descriptor contraction of se_e2_a for sm_100a:  D = (GR/nnei)^T (GR/nnei)[:, :axis]
// (deepmd/pt/model/descriptor/se_a.py:843-850, deepmd/tf/descriptor/se_a.py:1312-1324) and its
// backward.  GR = the [4][M] tabulate output of one atom; D is [M][axis] (M*axis = 1600 for water).
//
// In the reference this is a batched matmul of a [M x 4] by a [4 x axis] matrix per atom (cuBLAS
// batched GEMM with K = 4: launch- and pointer-bound).  It is a pure streaming op: read 4*M, write
// M*axis values per atom (forward), read M*axis + 4*M, write 4*M (backward) -- HBM-bound, so one
// WARP per atom with the operands staged in shared memory and fully coalesced stores.
#include <cuda_pipeline.h>

#include <cstdlib>

#include "common.cuh"

namespace dpb200 {
namespace {

extern __shared__ __align__(16) unsigned char desc_smem[];

template <typename FP>
__global__ void __launch_bounds__(128) k_desc_fwd(FP* __restrict__ D, const FP* __restrict__ X,
                                                  const int* __restrict__ rows, long long nloc, int M, int axis,
                                                  FP scale) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  FP* xs = reinterpret_cast<FP*>(desc_smem) + (size_t)warp * 4 * M;
  const int nout = M * axis;
  for (long long i = (long long)blockIdx.x * 4 + warp; i < nloc; i += (long long)gridDim.x * 4) {
    const FP* __restrict__ x = X + (rows ? (long long)rows[i] : i) * 4 * M;
    __syncwarp();
    for (int e = lane; e < 4 * M; e += 32) xs[e] = x[e] * scale;
    __syncwarp();
    FP* __restrict__ d = D + i * nout;
    if (32 % axis == 0) {  // the column of a lane never changes, its row advances by 32/axis: no division
      const int k2 = lane % axis, k10 = lane / axis, step = 32 / axis;
#pragma unroll 4
      for (int it = 0; it * 32 + lane < nout; ++it) {
        const int k1 = k10 + it * step;
        const FP v = xs[k1] * xs[k2] + xs[M + k1] * xs[M + k2] + xs[2 * M + k1] * xs[2 * M + k2] +
                     xs[3 * M + k1] * xs[3 * M + k2];
        st_cs(d + it * 32 + lane, v);
      }
    } else {
      for (int e = lane; e < nout; e += 32) {
        const int k1 = e / axis, k2 = e - k1 * axis;
        const FP v = xs[k1] * xs[k2] + xs[M + k1] * xs[M + k2] + xs[2 * M + k1] * xs[2 * M + k2] +
                     xs[3 * M + k1] * xs[3 * M + k2];
        st_cs(d + e, v);
      }
    }
  }
}

// dX[m][k] = scale * ( sum_{k2<axis} dD[k][k2] xs[m][k2]  +  [k<axis] sum_{k1<M} dD[k1][k] xs[m][k1] ),
// xs = X*scale.  Row i of dD belongs to atom rows[i] (or i): the result is written there.
// One warp per atom; the 12.8 KB of dD and the 3.2 KB of X of the NEXT atom stream into the second
// half of a per-warp double buffer with cp.async (LDGSTS, padded rows: conflict-free column walks)
// while the current atom is being contracted, so the kernel runs at HBM speed instead of one
// memory round trip per atom.
template <typename FP>
__global__ void __launch_bounds__(192) k_desc_bwd(FP* __restrict__ dX, const FP* __restrict__ dD,
                                                  const FP* __restrict__ X, const int* __restrict__ rows,
                                                  long long nloc, int M, int axis, FP scale) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ld = axis + 1;
  const int per_buf = 4 * M + M * ld;
  FP* base = reinterpret_cast<FP*>(desc_smem) + (size_t)warp * (2 * per_buf + 4 * axis);
  FP* t2 = base + 2 * per_buf;  // second term, [4][axis]
  const int nout = M * axis;
  const int nw = blockDim.x >> 5;
  const long long stride = (long long)gridDim.x * nw;

  auto issue = [&](long long i, FP* buf) {
    const long long src = rows ? (long long)rows[i] : i;
    const FP* __restrict__ x = X + src * 4 * M;
    const FP* __restrict__ gd = dD + i * nout;
    for (int e = lane; e < 4 * M; e += 32) __pipeline_memcpy_async(buf + e, x + e, sizeof(FP));
    FP* g = buf + 4 * M;
    if (32 % axis == 0) {
      const int k2 = lane % axis, k10 = lane / axis, step = 32 / axis;
#pragma unroll 4
      for (int it = 0; it * 32 + lane < nout; ++it)
        __pipeline_memcpy_async(g + (k10 + it * step) * ld + k2, gd + it * 32 + lane, sizeof(FP));
    } else {
      for (int e = lane; e < nout; e += 32) {
        const int k1 = e / axis, k2 = e - k1 * axis;
        __pipeline_memcpy_async(g + k1 * ld + k2, gd + e, sizeof(FP));
      }
    }
    __pipeline_commit();
  };

  long long i = (long long)blockIdx.x * nw + warp;
  int cur = 0;
  if (i < nloc) issue(i, base);
  while (i < nloc) {
    const long long ni = i + stride;
    if (ni < nloc) {
      issue(ni, base + (cur ^ 1) * per_buf);
      __pipeline_wait_prior(1);
    } else {
      __pipeline_wait_prior(0);
    }
    __syncwarp();
    const FP* xs = base + cur * per_buf;  // unscaled X
    const FP* g = xs + 4 * M;
    // second term: 4*axis outputs, each a dot product over k1 < M
    for (int o = lane; o < 4 * axis; o += 32) {
      const int m = o / axis, k = o - m * axis;
      FP acc = 0;
      for (int k1 = 0; k1 < M; ++k1) acc += g[k1 * ld + k] * xs[m * M + k1];
      t2[o] = acc;
    }
    __syncwarp();
    const long long dst = rows ? (long long)rows[i] : i;
    FP* __restrict__ o = dX + dst * 4 * M;
    const FP s2 = scale * scale;
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
        a0 += t2[k];
        a1 += t2[axis + k];
        a2 += t2[2 * axis + k];
        a3 += t2[3 * axis + k];
      }
      o[k] = a0 * s2;
      o[M + k] = a1 * s2;
      o[2 * M + k] = a2 * s2;
      o[3 * M + k] = a3 * s2;
    }
    __syncwarp();
    cur ^= 1;
    i = ni;
  }
}


// Backward, register-tiled (M <= 32*NC, axis % 4 == 0, axis <= 32): one warp per atom, lane <-> channels
// k = lane + 32c.  Each lane streams ITS OWN rows of dD (axis contiguous values = one 128-byte line in
// fp64) straight from global memory in 4-wide column blocks; the first term is a per-lane dot product,
// the second term's 4x4 partial sums of a column block are reduced over the warp with one butterfly
// reduce-scatter.  No staging of the 12.8 KB dD tile in shared memory, so occupancy is register-bound
// only and the kernel streams at HBM speed (the staged version above ran at ~25 % of it).
template <typename FP>
struct ColBlock {  // columns of dD per step = one 16-byte load
  static constexpr int W = 16 / sizeof(FP);
};
__device__ __forceinline__ void load_blk(const float* q, float (&v)[4]) {
  const float4 a = __ldcs(reinterpret_cast<const float4*>(q));
  v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w;
}
__device__ __forceinline__ void load_blk(const double* q, double (&v)[2]) {
  const double2 a = __ldcs(reinterpret_cast<const double2*>(q));
  v[0] = a.x, v[1] = a.y;
}

template <typename FP, int NC>
__global__ void __launch_bounds__(128, 4) k_desc_bwd_v2(FP* __restrict__ dX, const FP* __restrict__ dD,
                                                        const FP* __restrict__ X, const int* __restrict__ rows,
                                                        long long nloc, int M, int axis, FP scale) {
  constexpr int W = ColBlock<FP>::W;
  constexpr int NP = 4 * W;          // partial sums per column block
  constexpr int SH = W == 4 ? 1 : 2;  // lane l holds the sum of P[l >> SH]
  __shared__ FP sm[4][4 * 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const FP s2 = scale * scale;
  for (long long i = (long long)blockIdx.x * 4 + warp; i < nloc; i += (long long)gridDim.x * 4) {
    const long long src = rows ? (long long)rows[i] : i;
    const FP* __restrict__ x = X + src * 4 * M;
    const FP* __restrict__ g = dD + i * (long long)M * axis;
    FP xo[4][NC], out[4][NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int k = lane + 32 * c;
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        xo[m][c] = k < M ? x[m * M + k] : (FP)0.;
        out[m][c] = (FP)0.;
      }
    }
#pragma unroll 1
    for (int k2 = 0; k2 < axis; k2 += W) {
      FP xb[4][W];
#pragma unroll
      for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int t = 0; t < W; ++t) xb[m][t] = __ldg(x + m * M + k2 + t);
      FP P[NP];
#pragma unroll
      for (int q = 0; q < NP; ++q) P[q] = (FP)0.;
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const int k = lane + 32 * c;
        if (k < M) {
          FP gv[W];
          load_blk(g + (long long)k * axis + k2, gv);
#pragma unroll
          for (int m = 0; m < 4; ++m)
#pragma unroll
            for (int t = 0; t < W; ++t) {
              out[m][c] += gv[t] * xb[m][t];
              P[m * W + t] += gv[t] * xo[m][c];
            }
        }
      }
      const FP tot = reduce_scatter<NP>(P, lane);
      if ((lane & ((1 << SH) - 1)) == 0) {
        const int q = lane >> SH;
        sm[warp][(q / W) * 32 + k2 + (q % W)] = tot;
      }
    }
    __syncwarp();
    if (lane < axis) {
#pragma unroll
      for (int m = 0; m < 4; ++m) out[m][0] += sm[warp][m * 32 + lane];
    }
    FP* __restrict__ o = dX + src * 4 * M;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int k = lane + 32 * c;
      if (k < M) {
#pragma unroll
        for (int m = 0; m < 4; ++m) o[m * M + k] = out[m][c] * s2;
      }
    }
    __syncwarp();
  }
}

// Backward on the FP64 tensor cores (axis == 16, M >= 16; DMMA m8n8k4, fragment ownership: A[r][k] at
// lane 4r+k, B[k][c] at lane 4c+k, C[r][2k..2k+1] at lane 4r+k).  Per atom, with G = dD[i] ([M][16]) and
// X = gr ([4][M]):
//   term 2  T2[4 x 16] = X[4 x M] . G[M x 16]          A = X (rows 4..7 zero), k = 4 channels per step,
//                                                       B = G rows 4j..4j+3, two 8-column halves
//   term 1  T1[4 x M]  = X[:, :16][4 x 16] . G^T[16 x M] A = the first four A fragments of term 2,
//                                                       B = G^T, one C tile per 8 channels
// and dX = scale^2 (T1 + [channel < 16] T2): the C tiles of term 2 own exactly the lanes of the first two
// tiles of term 1, so they seed those accumulators.  ~100 DMMA and ~130 8-byte loads per atom, every
// fetched sector fully used, no shared memory, no shuffles; dD is read once from HBM (term 1 re-reads it
// from L2).  The SIMT versions above remain for other shapes and for fp32.
__device__ __forceinline__ void dmma884d(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

// (FP = float: fp32 operands and results, the products and sums still run on the FP64 tensor cores)
template <typename FP>
__global__ void __launch_bounds__(256) k_desc_bwd_mma(FP* __restrict__ dX, const FP* __restrict__ dD,
                                                      const FP* __restrict__ X, const int* __restrict__ rows,
                                                      long long nloc, int M, double scale) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = lane >> 2, kk = lane & 3;
  const double s2 = scale * scale;
  const int KT = (M + 3) >> 2, NT = (M + 7) >> 3;
  for (long long i = (long long)blockIdx.x * 8 + warp; i < nloc; i += (long long)gridDim.x * 8) {
    const long long src = rows ? (long long)rows[i] : i;
    const FP* __restrict__ x = X + src * 4 * M + (long long)(q & 3) * M;
    const FP* __restrict__ g = dD + i * (long long)M * 16;
    double c00 = 0., c01 = 0., c10 = 0., c11 = 0., d00 = 0., d01 = 0., d10 = 0., d11 = 0.;
    double as[4];
    // term 2 (first four steps unrolled: their A fragments are term 1's)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ch = 4 * j + kk;  // < 16 <= M
      as[j] = q < 4 ? (double)x[ch] : 0.;
      const double b0 = (double)__ldg(g + ch * 16 + q), b1 = (double)__ldg(g + ch * 16 + q + 8);
      if (j & 1) {
        dmma884d(d00, d01, as[j], b0);
        dmma884d(d10, d11, as[j], b1);
      } else {
        dmma884d(c00, c01, as[j], b0);
        dmma884d(c10, c11, as[j], b1);
      }
    }
#pragma unroll 4
    for (int j = 4; j < KT; ++j) {
      const int ch = 4 * j + kk;
      const bool ok = ch < M;
      const int chc = ok ? ch : M - 1;
      const double a = (q < 4 && ok) ? (double)x[chc] : 0.;
      const double b0 = (double)__ldg(g + chc * 16 + q), b1 = (double)__ldg(g + chc * 16 + q + 8);
      if (j & 1) {
        dmma884d(d00, d01, a, b0);
        dmma884d(d10, d11, a, b1);
      } else {
        dmma884d(c00, c01, a, b0);
        dmma884d(c10, c11, a, b1);
      }
    }
    c00 += d00, c01 += d01, c10 += d10, c11 += d11;
    // term 1
    FP* __restrict__ o = dX + src * 4 * M + (long long)(q & 3) * M;
#pragma unroll 2
    for (int t = 0; t < NT; ++t) {
      const int row = 8 * t + q;
      const FP* __restrict__ gr = g + (long long)(row < M ? row : M - 1) * 16 + kk;
      double e0 = t == 0 ? c00 : (t == 1 ? c10 : 0.);
      double e1 = t == 0 ? c01 : (t == 1 ? c11 : 0.);
      const double b0 = (double)__ldg(gr), b1 = (double)__ldg(gr + 4), b2 = (double)__ldg(gr + 8), b3 = (double)__ldg(gr + 12);
      dmma884d(e0, e1, as[0], b0);
      dmma884d(e0, e1, as[1], b1);
      dmma884d(e0, e1, as[2], b2);
      dmma884d(e0, e1, as[3], b3);
      if (q < 4) {
        const int ch = 8 * t + 2 * kk;
        if (ch < M) o[ch] = (FP)(e0 * s2);
        if (ch + 1 < M) o[ch + 1] = (FP)(e1 * s2);
      }
    }
  }
}

// Fused elementwise passes of the fitting MLP (deepmd/pt/model/network/mlp.py: tanh, resnet_dt, skip):
//   forward : a = tanh(z) (kept for the backward, overwrites z); y = a*idt (+ h when the widths match)
//   backward: t = g * idt * (1 - a^2)
// `sp` (fp32 only, may be null): the result is ALSO written as the 3xTF32 left operand of the next GEMM,
// sp[r] = [hi | lo | hi] (3*width floats per row), hi = tf32(v), lo = tf32(v - hi) -- see fitting.cu.
__device__ __forceinline__ void put_split3(float* sp, long long r, int c, int width, float v) {
  unsigned u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  const float hi = __uint_as_float(u);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v - hi));
  float* o = sp + r * 3 * width;
  o[c] = hi;
  o[width + c] = __uint_as_float(u);
  o[2 * width + c] = hi;
}
__device__ __forceinline__ void put_split3(double*, long long, int, int, double) {}

template <typename FP>
__global__ void k_mlp_act_fwd(FP* __restrict__ z_a, FP* __restrict__ y, const FP* __restrict__ h,
                              const FP* __restrict__ idt, long long n, int width, FP* __restrict__ sp) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / width;
    const int c = (int)(e - r * width);
    const FP a = tanh(z_a[e]);
    z_a[e] = a;
    FP v = idt ? a * idt[c] : a;
    if (h) v += h[e];
    y[e] = v;
    if (sp) put_split3(sp, r, c, width, v);
  }
}
template <typename FP>
__global__ void k_mlp_act_bwd(FP* __restrict__ t, const FP* __restrict__ g, long long ldg, const FP* __restrict__ a,
                              const FP* __restrict__ idt, long long n, int width, FP* __restrict__ sp) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / width;
    const int c = (int)(e - r * width);
    const FP av = a[e];
    FP v = g[r * ldg + c] * ((FP)1. - av * av);
    if (idt) v *= idt[c];
    if (t) t[e] = v;
    if (sp) put_split3(sp, r, c, width, v);
  }
}

inline bool desc_mma_enabled() {
  static const bool on = [] {
    const char* e = getenv("DPB200_DESC_MMA");
    return !(e && e[0] == '0');
  }();
  return on;
}

template <typename FP>
int desc_launch(bool bwd, FP* out, const FP* dD, const FP* X, const int* rows, long long nloc, int M, int axis,
                double scale, cudaStream_t st) {
  DPB_REQUIRE(nloc >= 0 && M >= 1 && axis >= 1 && axis <= M, "descriptor: need 1 <= axis <= M");
  if (nloc == 0) return DPB200_OK;
  DPB_REQUIRE(out && X && (!bwd || dD), "descriptor: null pointer");
  const size_t per_warp_bwd = (size_t)(2 * (4 * M + M * (axis + 1)) + 4 * axis) * sizeof(FP);
  int occ = 0;
  long long want = (nloc + 3) / 4;
  const bool v2 = bwd && M <= 128 && axis % 4 == 0 && axis <= 32 &&
                  (reinterpret_cast<uintptr_t>(dD) & 15) == 0 && ((long long)M * axis * sizeof(FP)) % 16 == 0;
  if (bwd && axis == 16 && M >= 16 && desc_mma_enabled()) {
    want = (nloc + 7) / 8;
    auto kern = k_desc_bwd_mma<FP>;
    DPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, 0));
    long long cap = (long long)sm_count() * (occ < 1 ? 1 : occ);
    kern<<<(int)(want < cap ? want : cap), 256, 0, st>>>(out, dD, X, rows, nloc, M, scale);
  } else if (v2) {
    want = (nloc + 3) / 4;
    const int nc = M <= 32 ? 1 : (M <= 64 ? 2 : 4);
#define DPB_DESC_BWD(NC)                                                                                  \
  do {                                                                                                    \
    auto kern = k_desc_bwd_v2<FP, NC>;                                                                    \
    DPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 128, 0));                          \
    long long cap = (long long)sm_count() * (occ < 1 ? 1 : occ);                                          \
    kern<<<(int)(want < cap ? want : cap), 128, 0, st>>>(out, dD, X, rows, nloc, M, axis, (FP)scale);     \
  } while (0)
    if (nc == 1)
      DPB_DESC_BWD(1);
    else if (nc == 2)
      DPB_DESC_BWD(2);
    else
      DPB_DESC_BWD(4);
#undef DPB_DESC_BWD
  } else if (bwd) {
    int nw = (int)((220 * 1024) / per_warp_bwd);
    if (nw > 6) nw = 6;
    DPB_REQUIRE(nw >= 1, "descriptor: M*axis too large for shared memory staging");
    const size_t smem = per_warp_bwd * nw;
    auto kern = k_desc_bwd<FP>;
    DPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, nw * 32, smem));
    want = (nloc + nw - 1) / nw;
    long long cap = (long long)sm_count() * (occ < 1 ? 1 : occ);
    kern<<<(int)(want < cap ? want : cap), nw * 32, smem, st>>>(out, dD, X, rows, nloc, M, axis, (FP)scale);
  } else {
    const size_t smem = (size_t)4 * M * sizeof(FP) * 4;
    DPB_REQUIRE(smem <= 200 * 1024, "descriptor: M too large for shared memory staging");
    auto kern = k_desc_fwd<FP>;
    DPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 128, smem));
    long long cap = (long long)sm_count() * (occ < 1 ? 1 : occ);
    kern<<<(int)(want < cap ? want : cap), 128, smem, st>>>(out, X, rows, nloc, M, axis, (FP)scale);
  }
  DPB_CUDA(cudaGetLastError());
  note_launches(1);
  return DPB200_OK;
}

template <typename FP>
int act_launch(bool bwd, FP* out, FP* z_a, const FP* g, long long ldg, const FP* h, const FP* idt, long long nrow,
               int width, cudaStream_t st, FP* sp = nullptr) {
  DPB_REQUIRE(sp == nullptr || sizeof(FP) == 4, "mlp activation: the TF32 split output exists for fp32 only");
  DPB_REQUIRE(nrow >= 0 && width >= 1, "mlp activation: bad shape");
  const long long n = nrow * width;
  if (n == 0) return DPB200_OK;
  int grid = ceil_div(n, 256);
  const int cap = sm_count() * 16;
  if (grid > cap) grid = cap;
  if (bwd) {
    DPB_REQUIRE((out || sp) && g && z_a, "mlp activation backward: null pointer");
    k_mlp_act_bwd<FP><<<grid, 256, 0, st>>>(out, g, ldg, z_a, idt, n, width, sp);
  } else {
    DPB_REQUIRE(out && z_a, "mlp activation forward: null pointer");
    k_mlp_act_fwd<FP><<<grid, 256, 0, st>>>(z_a, out, h, idt, n, width, sp);
  }
  DPB_CUDA(cudaGetLastError());
  note_launches(1);
  return DPB200_OK;
}

}  // namespace
}  // namespace dpb200

extern "C" {
#define DPB200_DEF_DESC(SUF, FP)                                                                        \
  int dpb200_se_a_descriptor_##SUF(FP* D, const FP* gr, const int* rows, long long nloc, int M,         \
                                   int axis, double scale, dpb200_stream_t stream) {                    \
    return dpb200::desc_launch<FP>(false, D, nullptr, gr, rows, nloc, M, axis, scale,                   \
                                   (cudaStream_t)stream);                                               \
  }                                                                                                     \
  int dpb200_se_a_descriptor_grad_##SUF(FP* dgr, const FP* dD, const FP* gr, const int* rows,           \
                                        long long nloc, int M, int axis, double scale,                  \
                                        dpb200_stream_t stream) {                                       \
    return dpb200::desc_launch<FP>(true, dgr, dD, gr, rows, nloc, M, axis, scale,                       \
                                   (cudaStream_t)stream);                                               \
  }                                                                                                     \
  int dpb200_mlp_tanh_fwd_##SUF(FP* z_a, FP* y, const FP* h, const FP* idt, long long nrow, int width,  \
                                dpb200_stream_t stream) {                                               \
    return dpb200::act_launch<FP>(false, y, z_a, nullptr, 0, h, idt, nrow, width,                       \
                                  (cudaStream_t)stream);                                                \
  }                                                                                                     \
  int dpb200_mlp_tanh_bwd_##SUF(FP* t, const FP* g, long long ldg, const FP* a, const FP* idt,          \
                                long long nrow, int width, dpb200_stream_t stream) {                    \
    return dpb200::act_launch<FP>(true, t, const_cast<FP*>(a), g, ldg, nullptr, idt, nrow, width,       \
                                  (cudaStream_t)stream);                                                \
  }                                                                                                     \
  int dpb200_mlp_tanh_fwd_split_##SUF(FP* z_a, FP* y, const FP* h, const FP* idt, long long nrow,       \
                                      int width, FP* split3, dpb200_stream_t stream) {                  \
    return dpb200::act_launch<FP>(false, y, z_a, nullptr, 0, h, idt, nrow, width,                       \
                                  (cudaStream_t)stream, split3);                                        \
  }                                                                                                     \
  int dpb200_mlp_tanh_bwd_split_##SUF(FP* t, const FP* g, long long ldg, const FP* a, const FP* idt,    \
                                      long long nrow, int width, FP* split3, dpb200_stream_t stream) {  \
    return dpb200::act_launch<FP>(true, t, const_cast<FP*>(a), g, ldg, nullptr, idt, nrow, width,       \
                                  (cudaStream_t)stream, split3);                                        \
  }
DPB200_DEF_DESC(f64, double)
DPB200_DEF_DESC(f32, float)
#undef DPB200_DEF_DESC
}
