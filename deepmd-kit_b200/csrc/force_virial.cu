// prod_force_a / prod_virial_a (and the fused single-pass form) for sm_100a.
//
// Semantics: source/lib/src/prod_force.cc:23-87 and prod_virial.cc:22-69 (CPU):
//   f_ij[d]  = sum_{c<4} net_deriv[i,4j+c] * in_deriv[i,4j+c,d]
//   force[i] -= sum_j f_ij        (over ALL slots, padded or not)
//   force[nlist[i,j]] += f_ij     (slots with nlist >= 0)
//   atom_virial[nlist[i,j], 3*d0+d1] += f_ij[d0] * rij[i,j,d1] ;  virial = sum_j atom_virial
//
// Design (not a port of source/lib/src/gpu/prod_force.cu / prod_virial.cu, which read in_deriv
// twice with one CTA per atom + a (64,3)/(16,9) thread grid doing one scalar atomic each, and
// reduce the virial on 9 CTAs):
//   * one WARP per centre atom, persistent grid; in_deriv (the 12*nnei*F bulk of the traffic) is
//     read ONCE with fully coalesced 16-byte loads, transposed through padded per-warp shared
//     memory (stride 13: conflict-free), then consumed one slot per lane;
//   * force and virial share f_ij, so the fused entry point reads net_deriv / in_deriv once;
//   * the centre-atom term is reduced in the warp (shuffles) and leaves as 3 atomics per atom;
//     neighbour terms are RED.ADD (no return value) straight to L2;
//   * the global virial is accumulated per lane in double across the persistent loop, reduced
//     per CTA and leaves as 9 atomics per CTA (no 9-CTA serial reduction over nall).
#include "common.cuh"

namespace dpb200 {
namespace {

constexpr int kStride = 13;

template <typename FP>
struct FvParams {
  FP* force;        // [nframes][nall][3] or null
  FP* virial;       // [9] or null
  FP* atom_virial;  // [nall][9] or null
  const FP* net_deriv;
  const FP* in_deriv;
  const FP* rij;
  const int* nlist;
  int nloc, nall, nnei;
  long long nrows;    // nframes * nloc
  int center_offset;  // centre atom of row r is r + center_offset (atom-chunked evaluation)
  // optional central pair force (se_atten switch path): slot (r, s) adds -(pair_q * pair_w) * rij to its neighbour
  // (and the opposite to the centre atom), with the same virial bookkeeping as the env-mat path
  const FP* pair_q;  // [nrows][nnei] or null
  const FP* pair_w;  // [nrows][nnei]
};

__device__ __forceinline__ void ld2(const double* q, double& a, double& b) {
  const double2 v = __ldcs(reinterpret_cast<const double2*>(q));
  a = v.x, b = v.y;
}
__device__ __forceinline__ void ld4(const float* q, float& a, float& b, float& c, float& d) {
  const float4 v = __ldcs(reinterpret_cast<const float4*>(q));
  a = v.x, b = v.y, c = v.z, d = v.w;
}

// Stage the 12 derivative components of up to 32 slots (n12 = nvalid*12 elements starting at
// src, 16-byte aligned) into smem as [slot][13].
__device__ __forceinline__ void stage_deriv(double* __restrict__ s, const double* __restrict__ src, int n12, int lane) {
  // 384 doubles = 192 double2 = 6 per lane
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    const int e = 2 * (lane + 32 * q);
    if (e < n12) {
      double a, b;
      ld2(src + e, a, b);
      const int sl = e / 12, c = e - sl * 12;
      s[sl * kStride + c] = a;
      s[sl * kStride + c + 1] = b;
    }
  }
}
__device__ __forceinline__ void stage_deriv(float* __restrict__ s, const float* __restrict__ src, int n12, int lane) {
  // 384 floats = 96 float4 = 3 per lane
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const int e = 4 * (lane + 32 * q);
    if (e < n12) {
      float a, b, c, d;
      ld4(src + e, a, b, c, d);
      const int sl = e / 12, cc = e - sl * 12;  // cc in {0,4,8}: the 4 values stay in one slot
      float* o = s + sl * kStride + cc;
      o[0] = a, o[1] = b, o[2] = c, o[3] = d;
    }
  }
}

template <typename FP, bool FORCE, bool VIRIAL, bool PAIR = false>
__global__ void __launch_bounds__(128) k_force_virial(const __grid_constant__ FvParams<FP> p) {
  __shared__ FP stage_all[4][32 * kStride];
  __shared__ double vred[4][9];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  FP* st = stage_all[warp];
  const int nnei = p.nnei;
  double vs[9];
#pragma unroll
  for (int q = 0; q < 9; ++q) vs[q] = 0.;

  for (long long row = (long long)blockIdx.x * 4 + warp; row < p.nrows; row += (long long)gridDim.x * 4) {
    const long long frame = row / p.nloc;
    const int i = (int)(row - frame * p.nloc) + p.center_offset;
    const FP* __restrict__ nd = p.net_deriv + row * nnei * 4;
    const FP* __restrict__ ed = p.in_deriv + row * nnei * 12;
    const int* __restrict__ nl = p.nlist + row * nnei;
    const FP* __restrict__ rj = VIRIAL ? p.rij + row * nnei * 3 : nullptr;
    FP* __restrict__ fo = FORCE ? p.force + frame * p.nall * 3 : nullptr;
    FP c0 = 0, c1 = 0, c2 = 0;
    for (int s0 = 0; s0 < nnei; s0 += 32) {
      const int nvalid = (nnei - s0) < 32 ? (nnei - s0) : 32;
      __syncwarp();
      stage_deriv(st, ed + (long long)s0 * 12, nvalid * 12, lane);
      __syncwarp();
      const int s = s0 + lane;
      if (lane < nvalid) {
        const typename Vec4<FP>::type g = *reinterpret_cast<const typename Vec4<FP>::type*>(nd + 4 * s);
        const FP* d = st + lane * kStride;
        FP f0 = g.x * d[0] + g.y * d[3] + g.z * d[6] + g.w * d[9];
        FP f1 = g.x * d[1] + g.y * d[4] + g.z * d[7] + g.w * d[10];
        FP f2 = g.x * d[2] + g.y * d[5] + g.z * d[8] + g.w * d[11];
        FP r0 = (FP)0., r1 = (FP)0., r2 = (FP)0.;
        if (PAIR) {  // (compiled out of the plain kernels: loading rij ahead of the list test costs them 10 %)
          r0 = rj[3 * s + 0], r1 = rj[3 * s + 1], r2 = rj[3 * s + 2];
          const FP cf = p.pair_q[row * nnei + s] * p.pair_w[row * nnei + s];
          f0 -= cf * r0, f1 -= cf * r1, f2 -= cf * r2;
        }
        c0 += f0, c1 += f1, c2 += f2;
        const int j = nl[s];
        if (j >= 0) {
          if (FORCE) {
            atomic_add(fo + 3 * (long long)j + 0, f0);
            atomic_add(fo + 3 * (long long)j + 1, f1);
            atomic_add(fo + 3 * (long long)j + 2, f2);
          }
          if (VIRIAL) {
            if (!PAIR) r0 = rj[3 * s + 0], r1 = rj[3 * s + 1], r2 = rj[3 * s + 2];
            const FP t[9] = {f0 * r0, f0 * r1, f0 * r2, f1 * r0, f1 * r1, f1 * r2, f2 * r0, f2 * r1, f2 * r2};
#pragma unroll
            for (int q = 0; q < 9; ++q) vs[q] += (double)t[q];
            if (p.atom_virial) {
              FP* av = p.atom_virial + 9 * (long long)j;
#pragma unroll
              for (int q = 0; q < 9; ++q) atomic_add(av + q, t[q]);
            }
          }
        }
      }
    }
    if (FORCE) {
      c0 = warp_sum(c0), c1 = warp_sum(c1), c2 = warp_sum(c2);
      if (lane < 3) atomic_add(fo + 3 * (long long)i + lane, lane == 0 ? -c0 : (lane == 1 ? -c1 : -c2));
    }
  }
  if (VIRIAL) {
#pragma unroll
    for (int q = 0; q < 9; ++q) vs[q] = warp_sum(vs[q]);
    if (lane == 0) {
#pragma unroll
      for (int q = 0; q < 9; ++q) vred[warp][q] = vs[q];
    }
    __syncthreads();
    if (threadIdx.x < 9) {
      const double t = vred[0][threadIdx.x] + vred[1][threadIdx.x] + vred[2][threadIdx.x] + vred[3][threadIdx.x];
      atomic_add(p.virial + threadIdx.x, (FP)t);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Gradients of prod_force_a / prod_virial_a with respect to net_deriv (the backward of the force / virial
// outputs when a compressed model is trained on forces and virials): deepmd::prod_force_grad_a_gpu
// (source/lib/include/prod_force_grad.h:26-33, CPU semantics src/prod_force_grad.cc:22-77) and
// deepmd::prod_virial_grad_a_gpu (prod_virial_grad.h:26-33, src/prod_virial_grad.cc:21-63).
//   force : gn[r][4k+c] = sum_d ( g[frame][j][d] - g[r][d] ) * ed[r][4k+c][d],  j = nlist[r][k] (j >= nloc -> j % nloc;
//           j < 0: only the centre term)
//   virial: gn[i][4k+c] = sum_{d0,d1} g[d0][d1] * rij[i][k][d1] * ed[i][4k+c][d0]   (0 for j < 0)
// Pure streaming ops (read 12 values, write 1 per element): one thread per output element, the three
// derivative components of consecutive elements are consecutive in memory.
// ------------------------------------------------------------------------------------------
template <typename FP>
__global__ void k_force_grad(FP* __restrict__ gn, const FP* __restrict__ g, const FP* __restrict__ ed,
                             const int* __restrict__ nlist, int nloc, int ngrad, int nnei, long long nrows) {
  // grad holds `ngrad` atoms per frame (the reference op: ngrad == nloc; the adjoint of a scatter into nall
  // atoms: ngrad == nall); neighbour indices beyond it are folded with j % ngrad
  const long long n = nrows * nnei * 4;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const long long rk = e >> 2;  // (row, neighbour slot)
    const long long r = rk / nnei;
    const long long frame = r / nloc;
    int j = nlist[rk];
    if (j >= ngrad) j = j % ngrad;
    const FP* __restrict__ d = ed + e * 3;
    const FP* __restrict__ gi = g + (frame * ngrad + (r - frame * nloc)) * 3;
    FP gx = -gi[0], gy = -gi[1], gz = -gi[2];
    if (j >= 0) {
      const FP* __restrict__ gj = g + (frame * ngrad + j) * 3;
      gx += gj[0], gy += gj[1], gz += gj[2];
    }
    // the reference accumulates centre and neighbour terms separately; the difference is rounding only
    gn[e] = gx * d[0] + gy * d[1] + gz * d[2];
  }
}

template <typename FP>
__global__ void k_virial_grad(FP* __restrict__ gn, const FP* __restrict__ g, const FP* __restrict__ ed,
                              const FP* __restrict__ rij, const int* __restrict__ nlist, int nnei, long long nloc) {
  const long long n = nloc * nnei * 4;
  FP gg[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) gg[k] = g[k];
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const long long ik = e >> 2;
    FP v = (FP)0.;
    if (nlist[ik] >= 0) {
      const FP* __restrict__ d = ed + e * 3;
      const FP* __restrict__ rr = rij + ik * 3;
#pragma unroll
      for (int d0 = 0; d0 < 3; ++d0) v += (gg[d0 * 3] * rr[0] + gg[d0 * 3 + 1] * rr[1] + gg[d0 * 3 + 2] * rr[2]) * d[d0];
    }
    gn[e] = v;
  }
}

template <typename FP>
int launch_fv_grad(bool virial, FP* grad_net, const FP* grad, const FP* in_deriv, const FP* rij, const int* nlist,
                   int nloc, int nnei, int nframes, cudaStream_t st, int ngrad = -1) {
  if (ngrad < 0) ngrad = nloc;
  DPB_REQUIRE(nloc >= 0 && nnei >= 0 && nframes >= 1 && ngrad >= nloc,
              "prod_force/virial_grad: need nloc, nnei >= 0, nframes >= 1, ngrad >= nloc");
  const long long nrows = (long long)nframes * nloc;
  const long long n = nrows * nnei * 4;
  if (n == 0) return DPB200_OK;
  DPB_REQUIRE(grad_net && grad && in_deriv && nlist && (!virial || rij), "prod_force/virial_grad: null pointer");
  int grid = ceil_div(n, 256);
  const int cap = sm_count() * 16;
  if (grid > cap) grid = cap;
  if (virial)
    k_virial_grad<FP><<<grid, 256, 0, st>>>(grad_net, grad, in_deriv, rij, nlist, nnei, nrows);
  else
    k_force_grad<FP><<<grid, 256, 0, st>>>(grad_net, grad, in_deriv, nlist, nloc, ngrad, nnei, nrows);
  DPB_CUDA(cudaGetLastError());
  note_launches(1);
  return DPB200_OK;
}

template <typename FP, bool FORCE, bool VIRIAL>
int launch_fv(FP* force, FP* virial, FP* atom_virial, const FP* net_deriv, const FP* in_deriv,
              const FP* rij, const int* nlist, int nloc, int nall, int nnei, int nframes,
              cudaStream_t st, int center_offset = 0, int accumulate = 0, const FP* pair_q = nullptr,
              const FP* pair_w = nullptr) {
  DPB_REQUIRE(nloc >= 0 && nall >= 0 && nnei >= 0 && nframes >= 1 && center_offset >= 0 &&
                  (long long)center_offset + nloc <= nall,
              "prod_force/virial: need nall >= center_offset + nloc >= 0, nnei >= 0, nframes >= 1");
  if (FORCE) {
    DPB_REQUIRE(force != nullptr || nall == 0, "prod_force_a: force is null");
    if (nall > 0 && !accumulate) DPB_CUDA(cudaMemsetAsync(force, 0, sizeof(FP) * (size_t)nframes * nall * 3, st));
  }
  if (VIRIAL) {
    DPB_REQUIRE(virial != nullptr, "prod_virial_a: virial is null");
    if (!accumulate) {
      DPB_CUDA(cudaMemsetAsync(virial, 0, sizeof(FP) * 9, st));
      if (atom_virial && nall > 0) DPB_CUDA(cudaMemsetAsync(atom_virial, 0, sizeof(FP) * (size_t)nall * 9, st));
    }
  }
  const long long nrows = (long long)nframes * nloc;
  if (nrows == 0 || nnei == 0) return DPB200_OK;
  DPB_REQUIRE(net_deriv && in_deriv && nlist && (!VIRIAL || rij), "prod_force/virial: null input");
  DPB_REQUIRE(((reinterpret_cast<uintptr_t>(net_deriv) | reinterpret_cast<uintptr_t>(in_deriv)) & 15) == 0,
              "prod_force/virial: net_deriv and in_deriv must be 16-byte aligned");
  FvParams<FP> p;
  p.force = force;
  p.virial = virial;
  p.atom_virial = atom_virial;
  p.net_deriv = net_deriv;
  p.in_deriv = in_deriv;
  p.rij = rij;
  p.nlist = nlist;
  p.nloc = nloc;
  p.nall = nall;
  p.nnei = nnei;
  p.nrows = nrows;
  p.center_offset = center_offset;
  DPB_REQUIRE((pair_q == nullptr) == (pair_w == nullptr) && (!pair_q || VIRIAL),
              "prod_force_virial_a_pair: pair_q and pair_w come together (fused force + virial entry only)");
  p.pair_q = pair_q;
  p.pair_w = pair_w;
  int occ = 0;
  long long want = (nrows + 3) / 4;
  if constexpr (FORCE && VIRIAL) {
    if (pair_q) {
      auto kern = k_force_virial<FP, true, true, true>;
      DPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 128, 0));
      long long cap = (long long)sm_count() * (occ < 1 ? 1 : occ);
      kern<<<(int)(want < cap ? want : cap), 128, 0, st>>>(p);
      DPB_CUDA(cudaGetLastError());
      note_launches(1);
      return DPB200_OK;
    }
  }
  auto kern = k_force_virial<FP, FORCE, VIRIAL>;
  DPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 128, 0));
  if (occ < 1) occ = 1;
  long long cap = (long long)sm_count() * occ;
  const int grid = (int)(want < cap ? want : cap);
  kern<<<grid, 128, 0, st>>>(p);
  DPB_CUDA(cudaGetLastError());
  note_launches(1);
  return DPB200_OK;
}

__global__ void k_nlist_map(int* __restrict__ nlist, const int* __restrict__ map, long long n) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int j = nlist[e];
    if (j >= 0) nlist[e] = map[j];
  }
}

}  // namespace
}  // namespace dpb200

extern "C" {

#define DPB200_DEF_FV(SUF, FP)                                                                     \
  int dpb200_prod_force_a_##SUF(FP* force, const FP* net_deriv, const FP* in_deriv,                \
                                const int* nlist, int nloc, int nall, int nnei, int nframes,       \
                                dpb200_stream_t stream) {                                          \
    return dpb200::launch_fv<FP, true, false>(force, nullptr, nullptr, net_deriv, in_deriv,        \
                                              nullptr, nlist, nloc, nall, nnei, nframes,           \
                                              (cudaStream_t)stream);                               \
  }                                                                                                \
  int dpb200_prod_virial_a_##SUF(FP* virial, FP* atom_virial, const FP* net_deriv,                 \
                                 const FP* in_deriv, const FP* rij, const int* nlist, int nloc,    \
                                 int nall, int nnei, dpb200_stream_t stream) {                     \
    return dpb200::launch_fv<FP, false, true>(nullptr, virial, atom_virial, net_deriv, in_deriv,   \
                                              rij, nlist, nloc, nall, nnei, 1,                     \
                                              (cudaStream_t)stream);                               \
  }                                                                                                \
  int dpb200_prod_force_virial_a_##SUF(FP* force, FP* virial, FP* atom_virial,                     \
                                       const FP* net_deriv, const FP* in_deriv, const FP* rij,     \
                                       const int* nlist, int nloc, int nall, int nnei,             \
                                       dpb200_stream_t stream) {                                   \
    return dpb200::launch_fv<FP, true, true>(force, virial, atom_virial, net_deriv, in_deriv,      \
                                             rij, nlist, nloc, nall, nnei, 1,                      \
                                             (cudaStream_t)stream);                                \
  }                                                                                                \
  int dpb200_prod_force_virial_a_ex_##SUF(FP* force, FP* virial, FP* atom_virial,                  \
                                          const FP* net_deriv, const FP* in_deriv, const FP* rij,  \
                                          const int* nlist, int nrows, int center_offset,          \
                                          int nall, int nnei, int accumulate,                      \
                                          dpb200_stream_t stream) {                                \
    return dpb200::launch_fv<FP, true, true>(force, virial, atom_virial, net_deriv, in_deriv,      \
                                             rij, nlist, nrows, nall, nnei, 1,                     \
                                             (cudaStream_t)stream, center_offset, accumulate);     \
  }
DPB200_DEF_FV(f64, double)
DPB200_DEF_FV(f32, float)

/* fused force + virial with an additional central pair force -(pair_q * pair_w) * rij per slot (se_atten: pair_q =
 * dE/d(sw) from the gated table backward, pair_w = sw'(r) / r from dpb200_se_atten_gate_scalars) */
#define DPB200_DEF_FVP(SUF, FP)                                                                    \
  int dpb200_prod_force_virial_a_pair_##SUF(FP* force, FP* virial, FP* atom_virial,                \
                                            const FP* net_deriv, const FP* in_deriv,               \
                                            const FP* rij, const int* nlist, const FP* pair_q,     \
                                            const FP* pair_w, int nloc, int nall, int nnei,        \
                                            dpb200_stream_t stream) {                              \
    return dpb200::launch_fv<FP, true, true>(force, virial, atom_virial, net_deriv, in_deriv,      \
                                             rij, nlist, nloc, nall, nnei, 1,                      \
                                             (cudaStream_t)stream, 0, 0, pair_q, pair_w);          \
  }
DPB200_DEF_FVP(f64, double)
DPB200_DEF_FVP(f32, float)
#undef DPB200_DEF_FVP

#define DPB200_DEF_FVG(SUF, FP)                                                                    \
  int dpb200_prod_force_grad_a_##SUF(FP* grad_net, const FP* grad, const FP* in_deriv,             \
                                     const int* nlist, int nloc, int nnei, int nframes,            \
                                     dpb200_stream_t stream) {                                     \
    return dpb200::launch_fv_grad<FP>(false, grad_net, grad, in_deriv, nullptr, nlist, nloc, nnei, \
                                      nframes, (cudaStream_t)stream);                              \
  }                                                                                                \
  int dpb200_prod_force_grad_a_ex_##SUF(FP* grad_net, const FP* grad, const FP* in_deriv,          \
                                        const int* nlist, int nloc, int ngrad, int nnei,           \
                                        int nframes, dpb200_stream_t stream) {                     \
    return dpb200::launch_fv_grad<FP>(false, grad_net, grad, in_deriv, nullptr, nlist, nloc, nnei, \
                                      nframes, (cudaStream_t)stream, ngrad);                       \
  }                                                                                                \
  int dpb200_prod_virial_grad_a_##SUF(FP* grad_net, const FP* grad, const FP* in_deriv,            \
                                      const FP* rij, const int* nlist, int nloc, int nnei,         \
                                      dpb200_stream_t stream) {                                    \
    return dpb200::launch_fv_grad<FP>(true, grad_net, grad, in_deriv, rij, nlist, nloc, nnei, 1,   \
                                      (cudaStream_t)stream);                                       \
  }
DPB200_DEF_FVG(f64, double)
DPB200_DEF_FVG(f32, float)
#undef DPB200_DEF_FVG
#undef DPB200_DEF_FV

int dpb200_use_nlist_map(int* nlist, const int* nlist_map, int nloc, int nnei, dpb200_stream_t stream) {
  using namespace dpb200;
  DPB_REQUIRE(nloc >= 0 && nnei >= 0, "use_nlist_map: negative size");
  const long long n = (long long)nloc * nnei;
  if (n == 0) return DPB200_OK;
  DPB_REQUIRE(nlist && nlist_map, "use_nlist_map: null pointer");
  int grid = ceil_div(n, 256);
  const int cap = sm_count() * 16;
  if (grid > cap) grid = cap;
  k_nlist_map<<<grid, 256, 0, (cudaStream_t)stream>>>(nlist, nlist_map, n);
  DPB_CUDA(cudaGetLastError());
  note_launches(1);
  return DPB200_OK;
}

}  // extern "C"
