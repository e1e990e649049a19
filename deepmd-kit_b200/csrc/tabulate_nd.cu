// tabulate_fusion_se_a for the higher angular bases, NDESCRPT = 9 / 16 / 25 (source/lib/src/tabulate.cc:162-447
// templates, dispatched at :456-560; GPU counterpart source/lib/src/gpu/tabulate.cu:293-629).  The se_a / se_atten
// hot path (NDESCRPT = 4) lives in tabulate.cu; these kernels complete the operator's contract and are written for
// clarity, not for the roofline: one warp per (atom, 32-channel slice) in the forward and second-order kernels, one
// warp per atom in the backward; the table is read in the reference's own [row][channel][6] layout.
//
// Padding fold (`is_sorted`): as in the reference, the fold test looks at components 1..3 only, whatever NDESCRPT is.
#include <cmath>
#include <type_traits>

#include "common.cuh"

namespace dpb200 {
namespace {

template <typename FP>
struct NdParams {
  const FP* table;
  const FP* em_x;  // [nloc][nnei]
  const FP* em;    // [nloc][nnei][ND]
  const FP* two;   // [nloc][nnei][M] or null
  FP lower, upper, vmax, s0, s1, tail_xx;
  int first, tail_idx;
  int nloc, nnei, M, is_sorted;
  // forward / second order
  FP* out;  // [nloc][ND][M]
  const FP* dz_x;
  const FP* dz_em;
  const FP* dz_two;
  // backward
  const FP* dy;  // [nloc][ND][M]
  FP* dy_dem_x;
  FP* dy_dem;
  FP* dy_dtwo;
};

// tabulate.cc:45-73
template <typename FP>
__device__ __forceinline__ void locate_nd(const NdParams<FP>& p, FP x0, FP& xx, int& idx, FP& delta) {
  delta = (FP)0.;
  if (x0 < p.lower) {
    idx = 0;
    xx = (FP)0.;
    delta = x0 - p.lower;
  } else if (x0 < p.upper) {
    idx = (int)((x0 - p.lower) / p.s0);
    xx = x0 - ((FP)idx * p.s0 + p.lower);
  } else if (x0 < p.vmax) {
    idx = p.first + (int)((x0 - p.upper) / p.s1);
    xx = x0 - ((FP)(idx - p.first) * p.s1 + p.upper);
  } else {
    idx = p.tail_idx;
    xx = p.tail_xx;
    delta = x0 - p.vmax;
  }
}

template <typename FP>
__device__ __forceinline__ void poly_nd(const FP* __restrict__ a, FP x, FP dl, FP& val, FP& grad) {
  const FP a0 = a[0], a1 = a[1], a2 = a[2], a3 = a[3], a4 = a[4], a5 = a[5];
  grad = a1 + ((FP)2. * a2 + ((FP)3. * a3 + ((FP)4. * a4 + (FP)5. * a5 * x) * x) * x) * x;
  val = a0 + (a1 + (a2 + (a3 + (a4 + a5 * x) * x) * x) * x) * x + grad * dl;
}

template <typename FP>
__device__ __forceinline__ bool fold_here(const NdParams<FP>& p, const FP* __restrict__ ll, FP xx, FP last) {
  return p.is_sorted && last == xx && ll[1] == (FP)0. && ll[2] == (FP)0. && ll[3] == (FP)0.;
}

// forward (GG = false) and second order (GG = true): out[i][m][k] = sum_j scale_j (var hh_j[m] + s_j ll_j[m])
template <typename FP, int ND, bool GG>
__global__ void __launch_bounds__(128) k_tab_nd_fwd(const __grid_constant__ NdParams<FP> p) {
  const int lane = threadIdx.x & 31;
  const int slices = (p.M + 31) / 32;
  const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long i = w / slices;
  const int k = (int)(w - i * slices) * 32 + lane;
  if (i >= p.nloc) return;
  const bool live = k < p.M;
  const int kc = live ? k : p.M - 1;
  FP acc[ND];
#pragma unroll
  for (int m = 0; m < ND; ++m) acc[m] = (FP)0.;
  const FP last = p.em_x[i * p.nnei + p.nnei - 1];
  for (int j = 0; j < p.nnei; ++j) {
    const long long pj = i * p.nnei + j;
    const FP* __restrict__ ll = p.em + pj * ND;
    const FP x0 = p.em_x[pj];
    const bool fold = fold_here(p, ll, x0, last);
    FP xx, dl;
    int idx;
    locate_nd(p, x0, xx, idx, dl);
    FP var, vg;
    poly_nd(p.table + ((long long)idx * p.M + kc) * 6, xx, dl, var, vg);
    FP s = (FP)0.;
    if (GG) {
      FP two_grad = (FP)0.;
      if (p.two) {
        const FP t = p.two[pj * p.M + kc];
        two_grad = p.dz_two[pj * p.M + kc] * var;
        var += var * t;
        vg += vg * t;
      }
      s = p.dz_x[pj] * vg + two_grad;
    } else if (p.two) {
      const FP t = p.two[pj * p.M + kc];
      var = var * t + var;
    }
    const FP mult = fold ? (FP)(p.nnei - j) : (FP)1.;
    if (GG) {
      const FP* __restrict__ hh = p.dz_em + pj * ND;
#pragma unroll
      for (int m = 0; m < ND; ++m) acc[m] += mult * (var * hh[m] + s * ll[m]);
    } else {
      const FP scale = fold ? mult * var : var;
#pragma unroll
      for (int m = 0; m < ND; ++m) acc[m] += scale * ll[m];
    }
    if (fold) break;
  }
  if (live) {
#pragma unroll
    for (int m = 0; m < ND; ++m) p.out[(i * ND + m) * (long long)p.M + k] = acc[m];
  }
}

// backward: dy_dem_x[i][j], dy_dem[i][j][m], dy_dtwo[i][j][k]; entries behind the fold stay zero (memset by the host)
template <typename FP, int ND>
__global__ void __launch_bounds__(128) k_tab_nd_grad(const __grid_constant__ NdParams<FP> p) {
  const int lane = threadIdx.x & 31;
  const long long i = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= p.nloc) return;
  const FP last = p.em_x[i * p.nnei + p.nnei - 1];
  const FP* __restrict__ dyi = p.dy + i * ND * (long long)p.M;
  for (int j = 0; j < p.nnei; ++j) {
    const long long pj = i * p.nnei + j;
    const FP* __restrict__ ll = p.em + pj * ND;
    const FP x0 = p.em_x[pj];
    const bool fold = fold_here(p, ll, x0, last);
    const FP mult = fold ? (FP)(p.nnei - j) : (FP)1.;
    FP xx, dl;
    int idx;
    locate_nd(p, x0, xx, idx, dl);
    FP gsum = (FP)0.;
    FP part[ND];
#pragma unroll
    for (int m = 0; m < ND; ++m) part[m] = (FP)0.;
    for (int k = lane; k < p.M; k += 32) {
      FP res, g;
      poly_nd(p.table + ((long long)idx * p.M + k) * 6, xx, dl, res, g);
      const FP res0 = res;
      if (p.two) {
        const FP t = p.two[pj * p.M + k];
        res = res * t + res;
        g += t * g;
      }
      FP dot = (FP)0.;
#pragma unroll
      for (int m = 0; m < ND; ++m) {
        const FP rr = dyi[(long long)m * p.M + k];
        dot += ll[m] * rr;
        part[m] += res * rr;
      }
      gsum += g * dot;
      if (p.two) p.dy_dtwo[pj * p.M + k] = mult * res0 * dot;
    }
    gsum = warp_sum(gsum);
#pragma unroll
    for (int m = 0; m < ND; ++m) part[m] = warp_sum(part[m]);
    if (lane == 0) p.dy_dem_x[pj] = gsum * mult;
#pragma unroll
    for (int m = 0; m < ND; ++m)
      if (lane == (m & 31)) p.dy_dem[pj * ND + m] = part[m] * mult;
    if (fold) break;
  }
}

template <typename FP>
int fill_nd(NdParams<FP>& p, const FP* table, const FP* info, const FP* em_x, const FP* em, const FP* two, int nloc,
            int nnei, int M, int is_sorted) {
  DPB_REQUIRE(table && info && em_x && em, "tabulate (ndescrpt > 4): null pointer (table_info is a HOST pointer)");
  p.table = table;
  p.em_x = em_x;
  p.em = em;
  p.two = two;
  p.lower = info[0];
  p.upper = info[1];
  p.vmax = info[2];
  p.s0 = info[3];
  p.s1 = info[4];
  DPB_REQUIRE(p.s0 > (FP)0. && p.s1 > (FP)0. && p.upper >= p.lower && p.vmax >= p.upper,
              "tabulate: table_info must satisfy lower <= upper <= max and positive strides");
  // tabulate.cc:21-30
  p.first = (int)((p.upper - p.lower) / p.s0);
  const FP edge = std::nextafter(p.vmax, p.lower);
  p.tail_idx = p.first + (int)((edge - p.upper) / p.s1);
  p.tail_xx = p.vmax - ((FP)(p.tail_idx - p.first) * p.s1 + p.upper);
  p.nloc = nloc;
  p.nnei = nnei;
  p.M = M;
  p.is_sorted = is_sorted ? 1 : 0;
  return DPB200_OK;
}

inline bool nd_ok(int nd) { return nd == 9 || nd == 16 || nd == 25; }

template <typename FP, bool GG>
int launch_nd_fwd(FP* out, const FP* table, const FP* info, const FP* em_x, const FP* em, const FP* two,
                  const FP* dz_x, const FP* dz_em, const FP* dz_two, int nloc, int nnei, int M, int is_sorted, int nd,
                  cudaStream_t st) {
  DPB_REQUIRE(nd_ok(nd), "tabulate: ndescrpt must be 4, 9, 16 or 25 (tabulate.cc check_se_a_basis_dimension)");
  DPB_REQUIRE(nloc >= 0 && nnei >= 0 && M >= 0, "tabulate: negative size");
  if (nloc == 0 || M == 0) return DPB200_OK;
  DPB_REQUIRE(out != nullptr, "tabulate: out is null");
  if (nnei == 0) {  // an empty neighbour axis is a valid empty reduction (tabulate.cc:176-181)
    DPB_CUDA(cudaMemsetAsync(out, 0, sizeof(FP) * (size_t)nloc * nd * M, st));
    return DPB200_OK;
  }
  NdParams<FP> p = {};
  int rc = fill_nd(p, table, info, em_x, em, two, nloc, nnei, M, is_sorted);
  if (rc) return rc;
  p.out = out;
  p.dz_x = dz_x;
  p.dz_em = dz_em;
  p.dz_two = dz_two;
  if (GG) {
    DPB_REQUIRE(dz_x && dz_em && (!two || dz_two), "tabulate grad_grad: null cotangent");
  }
  const long long warps = (long long)nloc * ((M + 31) / 32);
  const unsigned grid = (unsigned)((warps + 3) / 4);
  switch (nd) {
    case 9: k_tab_nd_fwd<FP, 9, GG><<<grid, 128, 0, st>>>(p); break;
    case 16: k_tab_nd_fwd<FP, 16, GG><<<grid, 128, 0, st>>>(p); break;
    default: k_tab_nd_fwd<FP, 25, GG><<<grid, 128, 0, st>>>(p); break;
  }
  DPB_CUDA(cudaGetLastError());
  note_launches(1);
  return DPB200_OK;
}

template <typename FP>
int launch_nd_grad(FP* dy_dem_x, FP* dy_dem, FP* dy_dtwo, const FP* table, const FP* info, const FP* em_x,
                   const FP* em, const FP* two, const FP* dy, int nloc, int nnei, int M, int is_sorted, int nd,
                   cudaStream_t st) {
  DPB_REQUIRE(nd_ok(nd), "tabulate grad: ndescrpt must be 4, 9, 16 or 25");
  DPB_REQUIRE(nloc >= 0 && nnei >= 0 && M >= 0, "tabulate grad: negative size");
  if (nloc == 0 || nnei == 0) return DPB200_OK;  // tabulate.cc:251-253: nothing to write
  DPB_REQUIRE(dy_dem_x && dy_dem && dy && (!two || dy_dtwo), "tabulate grad: null pointer");
  DPB_CUDA(cudaMemsetAsync(dy_dem_x, 0, sizeof(FP) * (size_t)nloc * nnei, st));
  DPB_CUDA(cudaMemsetAsync(dy_dem, 0, sizeof(FP) * (size_t)nloc * nnei * nd, st));
  if (two) DPB_CUDA(cudaMemsetAsync(dy_dtwo, 0, sizeof(FP) * (size_t)nloc * nnei * M, st));
  if (M == 0) return DPB200_OK;
  NdParams<FP> p = {};
  int rc = fill_nd(p, table, info, em_x, em, two, nloc, nnei, M, is_sorted);
  if (rc) return rc;
  p.dy = dy;
  p.dy_dem_x = dy_dem_x;
  p.dy_dem = dy_dem;
  p.dy_dtwo = dy_dtwo;
  const unsigned grid = (unsigned)(((long long)nloc + 3) / 4);
  switch (nd) {
    case 9: k_tab_nd_grad<FP, 9><<<grid, 128, 0, st>>>(p); break;
    case 16: k_tab_nd_grad<FP, 16><<<grid, 128, 0, st>>>(p); break;
    default: k_tab_nd_grad<FP, 25><<<grid, 128, 0, st>>>(p); break;
  }
  DPB_CUDA(cudaGetLastError());
  note_launches(1);
  return DPB200_OK;
}

}  // namespace
}  // namespace dpb200

extern "C" {

#define DPB200_DEF_TAB_ND(SUF, FP)                                                                                \
  int dpb200_tabulate_fusion_se_a_nd_##SUF(FP* out, const FP* table, const FP* table_info, const FP* em_x,       \
                                           const FP* em, const FP* two_embed, int nloc, int nnei,                \
                                           int last_layer_size, int is_sorted, int ndescrpt,                     \
                                           dpb200_stream_t stream) {                                             \
    return dpb200::launch_nd_fwd<FP, false>(out, table, table_info, em_x, em, two_embed, nullptr, nullptr,       \
                                            nullptr, nloc, nnei, last_layer_size, is_sorted, ndescrpt,           \
                                            (cudaStream_t)stream);                                               \
  }                                                                                                               \
  int dpb200_tabulate_fusion_se_a_grad_nd_##SUF(FP* dy_dem_x, FP* dy_dem, FP* dy_dtwo, const FP* table,          \
                                                const FP* table_info, const FP* em_x, const FP* em,              \
                                                const FP* two_embed, const FP* dy, int nloc, int nnei,           \
                                                int last_layer_size, int is_sorted, int ndescrpt,                \
                                                dpb200_stream_t stream) {                                        \
    return dpb200::launch_nd_grad<FP>(dy_dem_x, dy_dem, dy_dtwo, table, table_info, em_x, em, two_embed, dy,     \
                                      nloc, nnei, last_layer_size, is_sorted, ndescrpt, (cudaStream_t)stream);   \
  }                                                                                                               \
  int dpb200_tabulate_fusion_se_a_grad_grad_nd_##SUF(FP* dz_dy, const FP* table, const FP* table_info,           \
                                                     const FP* em_x, const FP* em, const FP* two_embed,          \
                                                     const FP* dz_dy_dem_x, const FP* dz_dy_dem,                 \
                                                     const FP* dz_dy_dtwo, int nloc, int nnei,                   \
                                                     int last_layer_size, int is_sorted, int ndescrpt,           \
                                                     dpb200_stream_t stream) {                                   \
    return dpb200::launch_nd_fwd<FP, true>(dz_dy, table, table_info, em_x, em, two_embed, dz_dy_dem_x,           \
                                           dz_dy_dem, dz_dy_dtwo, nloc, nnei, last_layer_size, is_sorted,        \
                                           ndescrpt, (cudaStream_t)stream);                                      \
  }
DPB200_DEF_TAB_ND(f64, double)
DPB200_DEF_TAB_ND(f32, float)
#undef DPB200_DEF_TAB_ND

}  // extern "C"
