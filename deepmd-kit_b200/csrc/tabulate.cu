// tabulate_fusion_se_a / se_atten (forward, first- and second-order backward) for sm_100a.
//
// Semantics: source/lib/src/tabulate.cc:162-447 (CPU), i.e. for every centre atom i
//   out[i,m,k] = sum_j em[i,j,m] * g_k(em_x[i,j]),  g_k = quintic of table row locate_xx(em_x)
// with linear extrapolation outside [lower,max) (tabulate.cc:45-73,122-158), the optional
// se_atten gate g <- g*t + g, and the `is_sorted` fold of the trailing padding (":197-201").
//
// What bounds this op on B200 (measured, profiles/r01_*): neither HBM nor L2 but the L1/shared
// data pipe and the issue slots -- every (neighbour, channel) needs 6 coefficients = 4.8 KB per
// neighbour in fp64 for M=100, ~90 neighbours per atom, against only 9 (forward) / 18 (backward)
// FMAs per coefficient set.  Design:
//   * one WARP per centre atom, lanes over channels (NC channels per lane), persistent loop, no
//     block barrier after the preload.  (Splitting the channel axis over CTAs/clusters to enlarge
//     the on-chip window was measured and lost: per-neighbour bookkeeping and the backward's
//     reduction are then paid once per slice -- profiles/r01_ncu_full_v2_sliced_summary.csv.)
//   * the table is re-laid out once per call as T[row][3][M] of coefficient PAIRS so that one
//     16-byte request per lane is fully dense (the reference layout [row][M][6] costs 2.5x the L1
//     wavefronts); the most crowded rows (the row histogram is extremely skewed: 50 % of all
//     look-ups hit 16 rows because sw(r)/r flattens towards rcut) are kept in shared memory;
//   * neighbours are distance-sorted, so ~35-40 % of them share the previous row: coefficients
//     stay in registers until the row changes;
//   * neighbours are "located" 32 at a time in parallel (exact FP division as the reference) one
//     work item AHEAD of their use (software prefetch of em_x / em hides the HBM latency), records
//     {dx, delta, mult*em[4], row} are staged in per-warp shared memory and read back as broadcasts;
//   * backward: 4 neighbours x 4 components = 16 partial sums per lane are reduced with one
//     butterfly reduce-scatter (16 shuffles) that leaves the 16 contiguous dy_dem values of those
//     neighbours in the even lanes -> coalesced store; dy_dem_x needs 6 more shuffles (the
//     reference does five full warp reductions per neighbour).
// Not a port of source/lib/src/gpu/tabulate.cu (one thread per channel, stride-6 scalar loads, all
// coefficients from global memory, serial padding search by thread 0).
#include <cuda_fp16.h>
#include <cuda_pipeline.h>

#include <cmath>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"


namespace dpb200 {
namespace {

template <typename FP>
struct TabParams {
  const FP* table;  // [nrow][M][6], reference layout
  const FP* T;      // [nrow][3][M][2]: coefficient pairs, channel-contiguous (scratch, built per call)
  FP lower, upper, vmax, s0, s1;
  int first;     // int((upper-lower)/s0)
  int tail_idx;  // row of x >= max
  FP tail_xx;    // max - start of that row
  int nrow;
  const FP* em_x;
  long long ldx_i;
  int ldx_j;
  const FP* em;
  long long ldem_i;
  const FP* two;  // [nloc][nnei][M] or null
  // se_atten strip gate without the materialised two_embed: two_embed[i][j][k] = gate_tt[gate_pair[i][j]][k] * gate_sw[i][j]
  const FP* gate_tt;     // [(ntypes+1)^2][M] type-pair table, or null
  const int* gate_pair;  // [nloc][nnei]
  const FP* gate_sw;     // [nloc][nnei]
  FP* gate_q;            // backward: dE/d(gate_sw) [nloc][nnei]
  int nloc, nnei, M, is_sorted, accumulate, vec_ok;
  int Mc;  // channels per CTA slice (forward: gridDim.y slices; backward: Mc == M)
  int H;   // hot rows kept in shared memory
  int hot_elems;  // H*Mc*6 (+ padding) rounded up to a 16-byte boundary (start of the per-warp records)
  int hot_pad;    // bytes of padding after every hot row (tensor-core forward: 32, see k_tab_fwd_mma)
  int two_ring;   // se_atten: two_embed rows are staged through a per-warp shared-memory ring with cp.async
  int nblk;       // 16-byte blocks per (row, channel): 3 (coefficient pairs) or 2 (compressed, see k_table_relayout_cm)
  const FP* T3;   // compressed mode: the full pair table as well (stride-1 "coarse" rows are never compressed)
  float a5_mul, a5_inv;  // compressed layout: a5 is stored as half(a5 * a5_mul); a5_inv = 1 / a5_mul (powers of two)
  // forward / second order
  FP* out;  // [nloc][4][M]
  const FP* dz_x;
  const FP* dz_em;
  const FP* dz_two;
  // first-order backward
  const FP* dy;  // [nloc][4][M]
  FP* dy_dem_x;
  FP* dy_dem;
  FP* dy_dtwo;
  // descriptor emission fused into the forward's epilogue (k_tab_fwd<..., DESC = true>)
  void* desc;           // mode 1: FP [row][M*axis]; mode 2: split operand of the fitting net's first GEMM
  long long desc_ld;    // elements (of the stored type) per descriptor row
  const int* desc_row;  // descriptor row of atom i (null: i)
  int* row_exp;         // mode 2, fp64: binary exponent of the row's fixed-point scale
  FP desc_scale;        // 1/nnei
  int desc_mode, axis, nslice;
  long long desc_slice_stride;  // mode 2 / 3: elements between two digit slices of a row (>= M*axis; extra columns
                                // are the caller's: se_atten appends the centre type embedding there)
  int desc_min_exp;             // mode 2 / 3: lower bound of the row exponent (so that the caller's extra columns fit)
};

template <typename FP>
struct alignas(16) Rec {
  FP xx, delta;
  FP e[4];
  int idx;
  int mult;
};
template <typename FP>
struct alignas(16) RecGG {  // second order only: dz_dy_dem[4], dz_dy_dem_x of this neighbour
  FP h[4];
  FP zx;
  FP pad_;
};

// tabulate.cc:45-73
template <typename FP>
__device__ __forceinline__ void locate(const TabParams<FP>& p, FP x0, FP& xx, int& idx, FP& delta) {
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

__device__ __forceinline__ void load4(const float* q, bool vec, float (&e)[4]) {
  if (vec) {
    const float4 v = *reinterpret_cast<const float4*>(q);
    e[0] = v.x, e[1] = v.y, e[2] = v.z, e[3] = v.w;
  } else {
    e[0] = q[0], e[1] = q[1], e[2] = q[2], e[3] = q[3];
  }
}
__device__ __forceinline__ void load4(const double* q, bool vec, double (&e)[4]) {
  if (vec) {
    const double2 a = *reinterpret_cast<const double2*>(q);
    const double2 b = *reinterpret_cast<const double2*>(q + 2);
    e[0] = a.x, e[1] = a.y, e[2] = b.x, e[3] = b.y;
  } else {
    e[0] = q[0], e[1] = q[1], e[2] = q[2], e[3] = q[3];
  }
}

template <typename FP>
struct Pair2;
template <>
struct Pair2<double> {
  using type = double2;
};
template <>
struct Pair2<float> {
  using type = float2;
};

// [row][M][6] -> [row][3][M] pairs
template <typename FP>
__global__ void k_table_relayout(FP* __restrict__ T, const FP* __restrict__ table, long long nrow, int M) {
  const long long n = nrow * 6 * (long long)M;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n;
       e += (long long)gridDim.x * blockDim.x) {
    const int h = (int)(e & 1);
    const long long t = e >> 1;
    const int k = (int)(t % M);
    const long long rq = t / M;
    const int q = (int)(rq % 3);
    const long long r = rq / 3;
    T[e] = table[(r * M + k) * 6 + 2 * q + h];
  }
}

// Compressed coefficients (fp64, opt-in by the caller who has validated its table: DPB200_TAB_COMPRESSED_COEF).
// The table kernels are bound by the L1/shared data pipe (48 B of coefficients per evaluation), and for a
// dp-compress table the high-order terms are tiny: a3 x^3 <= 5e-8 |a0|, a4 x^4 <= 2e-10, a5 x^5 <= 5e-13 on the
// stride-0.01 rows.  So a3 and a4 are stored as fp32 and a5 as fp16 in the 16 low mantissa bits of a2 (which keeps
// 36 bits: 1.5e-11 relative on a term that is <= 1e-5 |a0|):  32 B per (row, channel) = two 16-byte blocks
//   block 0 = {a0, a1}   block 1 = {a2 | half(a5), (float a3, float a4)}
// Errors against the fp64 table: < 1e-14 |a0| on the value and < 3e-12 |a1| on the derivative for the water
// table (ops.compressed_coef_flags computes the bound for the actual table and only then sets the flag).
// Unpacking stays off the integer ALU (the kernel is sensitive to it: a bf16 a5 unpacked with one shift, or
// widening fp32 -> fp64 with integer instructions, measured 8 % and 24 % slower than the F2F conversions).
__device__ __forceinline__ double2 pack_cm(double a2, double a3, double a4, double a5, float a5_mul) {
  const unsigned short h5 = __half_as_ushort(__float2half_rn((float)(a5 * (double)a5_mul)));
  const unsigned long long b2 = ((unsigned long long)__double_as_longlong(a2) & ~0xffffull) | (unsigned long long)h5;
  const unsigned long long b34 = (unsigned long long)__float_as_uint((float)a3) |
                                 ((unsigned long long)__float_as_uint((float)a4) << 32);
  return make_double2(__longlong_as_double((long long)b2), __longlong_as_double((long long)b34));
}
__device__ __forceinline__ void unpack_cm(const double2 v, double& a2, float& a3, float& a4, float& a5) {
  const unsigned long long b2 = (unsigned long long)__double_as_longlong(v.x);
  const unsigned long long b34 = (unsigned long long)__double_as_longlong(v.y);
  a2 = v.x;  // the 16 borrowed bits are noise at 2^-36 relative
  a5 = __half2float(__ushort_as_half((unsigned short)(b2 & 0xffffull)));
  a3 = __uint_as_float((unsigned)(b34 & 0xffffffffull));
  a4 = __uint_as_float((unsigned)(b34 >> 32));
}

// [row][M][6] -> [row][2][M] blocks
__global__ void k_table_relayout_cm(double2* __restrict__ T, const double* __restrict__ table, long long nrow, int M,
                                    float a5_mul) {
  const long long n = nrow * (long long)M;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n;
       e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / M;
    const int k = (int)(e - r * M);
    const double* a = table + e * 6;
    T[(r * 2 + 0) * M + k] = make_double2(a[0], a[1]);
    T[(r * 2 + 1) * M + k] = pack_cm(a[2], a[3], a[4], a[5], a5_mul);
  }
}

template <typename FP>
__device__ __forceinline__ FP poly(const FP (&a)[6], FP x);
// fp32 flavour of DPB200_TAB_COMPRESSED_COEF: at fp32 precision the quintic of a dp-compress table IS a cubic
// (a4 x^4 <= 2e-10 |a0|, a5 x^5 <= 5e-13 |a0| on the stride-0.01 rows, 4 a4 x^3 <= 1e-7 |a1|), so the kernels
// stream one float4 {a0, a1, a2, a3} per (row, channel) instead of three float2 pairs and evaluate 3 (value) /
// 5 (value + derivative) FMAs instead of 5 / 9.  Same host-side gate (ops.compressed_coef_flags).
__global__ void k_table_relayout_c32(float4* __restrict__ T, const float* __restrict__ table, long long nrow, int M) {
  const long long n = nrow * (long long)M;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n;
       e += (long long)gridDim.x * blockDim.x) {
    const float* a = table + e * 6;
    T[e] = make_float4(a[0], a[1], a[2], a[3]);
  }
}
template <int NC>
__device__ __forceinline__ void fetch_row_c32(float (&a)[NC][6], const float* __restrict__ hot,
                                              const float* __restrict__ T, int row, int r0, int H, int M,
                                              const int (&ob)[NC]) {
  const unsigned rel = (unsigned)(row - r0);
  const unsigned rb = (unsigned)M * 16u;
  if (rel < (unsigned)H) {
    const char* b0 = reinterpret_cast<const char*>(hot) + rel * rb;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const float4 u = *reinterpret_cast<const float4*>(b0 + 2 * ob[c]);
      a[c][0] = u.x, a[c][1] = u.y, a[c][2] = u.z, a[c][3] = u.w;
    }
  } else {
    const char* b0 = reinterpret_cast<const char*>(T) + (long long)row * rb;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const float4 u = __ldg(reinterpret_cast<const float4*>(b0 + 2 * ob[c]));
      a[c][0] = u.x, a[c][1] = u.y, a[c][2] = u.z, a[c][3] = u.w;
    }
  }
}
template <int NC>
__device__ __forceinline__ void fetch_row_c32(double (&)[NC][6], const double*, const double*, int, int, int, int,
                                              const int (&)[NC]) {}
template <typename FP>
__device__ __forceinline__ FP poly3(const FP (&a)[6], FP x) {
  return a[0] + (a[1] + (a[2] + a[3] * x) * x) * x;
}
template <typename FP>
__device__ __forceinline__ void poly3_both(const FP (&a)[6], FP x, FP& g, FP& gd) {
  const FP b2 = a[2] + a[3] * x;
  const FP b1 = a[1] + b2 * x;
  g = a[0] + b1 * x;
  const FP c2 = b2 + a[3] * x;
  gd = b1 + c2 * x;
}

// fetch_row for the compressed layout in the SIMT forward: a0..a2 go to the fp64 registers, a3..a5 stay fp32
// (8 instead of 12 sixteen-byte requests per lane and row, 9 instead of 12 registers per channel).
template <int NC, int NA>
__device__ __forceinline__ void fetch_row_cmf(double (&a)[NC][6], float (&af_)[NA][3], const double* __restrict__ hot,
                                              const double* __restrict__ T, int row, int r0, int H, int M,
                                              const int (&ob)[NC], float a5_inv) {
  const unsigned rel = (unsigned)(row - r0);
  const unsigned qb = (unsigned)M * 16u;
  if (rel < (unsigned)H) {
    const char* b0 = reinterpret_cast<const char*>(hot) + rel * (2u * qb);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const double2 u = *reinterpret_cast<const double2*>(b0 + ob[c]);
      const double2 v = *reinterpret_cast<const double2*>(b0 + qb + ob[c]);
      a[c][0] = u.x, a[c][1] = u.y;
      float (&af)[3] = af_[NA == NC ? c : 0];
      unpack_cm(v, a[c][2], af[0], af[1], af[2]);
      af[2] *= a5_inv;
    }
  } else {
    const char* b0 = reinterpret_cast<const char*>(T) + (long long)row * (2u * qb);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const double2 u = __ldg(reinterpret_cast<const double2*>(b0 + ob[c]));
      const double2 v = __ldg(reinterpret_cast<const double2*>(b0 + qb + ob[c]));
      a[c][0] = u.x, a[c][1] = u.y;
      float (&af)[3] = af_[NA == NC ? c : 0];
      unpack_cm(v, a[c][2], af[0], af[1], af[2]);
      af[2] *= a5_inv;
    }
  }
}
template <int NC, int NA>
__device__ __forceinline__ void fetch_row_cmf(float (&)[NC][6], float (&)[NA][3], const float*, const float*, int, int,
                                              int, int, const int (&)[NC], float) {}
// value of the quintic with the top two Horner steps on the FP32 pipe
__device__ __forceinline__ double poly_cm(const double (&a)[6], const float (&f)[3], double x, float xf) {
  const float t = fmaf(fmaf(f[2], xf, f[1]), xf, f[0]);
  return a[0] + (a[1] + (a[2] + (double)t * x) * x) * x;
}
__device__ __forceinline__ float poly_cm(const float (&a)[6], const float (&)[3], float x, float) { return poly(a, x); }

// Coefficients of one table row for the NC channels of this lane.  ob[c] = byte offset of the lane's
// channel inside one coefficient-pair block (clamped to channel M-1 so that every lane always loads:
// no divergence, the duplicates cost no extra wavefront).  Source: the shared-memory window
// [r0, r0+H) or the global pair table; the window test is warp-uniform.
template <typename FP, int NC>
__device__ __forceinline__ void fetch_row(FP (&a)[NC][6], const FP* __restrict__ hot, const FP* __restrict__ T,
                                          int row, int r0, int H, int M, const int (&ob)[NC]) {
  using P2 = typename Pair2<FP>::type;
  const unsigned rel = (unsigned)(row - r0);
  const unsigned qb = (unsigned)M * (unsigned)sizeof(P2);
  if (rel < (unsigned)H) {
    const char* b0 = reinterpret_cast<const char*>(hot) + rel * (3u * qb);
    const char* b1 = b0 + qb;
    const char* b2 = b1 + qb;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const P2 u = *reinterpret_cast<const P2*>(b0 + ob[c]);
      const P2 v = *reinterpret_cast<const P2*>(b1 + ob[c]);
      const P2 w = *reinterpret_cast<const P2*>(b2 + ob[c]);
      a[c][0] = u.x, a[c][1] = u.y, a[c][2] = v.x, a[c][3] = v.y, a[c][4] = w.x, a[c][5] = w.y;
    }
  } else {
    const char* b0 = reinterpret_cast<const char*>(T) + (long long)row * (3u * qb);
    const char* b1 = b0 + qb;
    const char* b2 = b1 + qb;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const P2 u = __ldg(reinterpret_cast<const P2*>(b0 + ob[c]));
      const P2 v = __ldg(reinterpret_cast<const P2*>(b1 + ob[c]));
      const P2 w = __ldg(reinterpret_cast<const P2*>(b2 + ob[c]));
      a[c][0] = u.x, a[c][1] = u.y, a[c][2] = v.x, a[c][3] = v.y, a[c][4] = w.x, a[c][5] = w.y;
    }
  }
}

// First row of the hot window: the row of the last slot of atom 0 (padding or the farthest
// neighbour: the crowded end of the table) minus a small margin; every CTA picks the same one.
template <typename FP>
__device__ __forceinline__ int hot_window_start(const TabParams<FP>& p) {
  FP xx, dl;
  int idx;
  locate(p, p.em_x[(long long)(p.nnei - 1) * p.ldx_j], xx, idx, dl);
  int r0 = idx - 4;
  if (r0 > p.nrow - p.H) r0 = p.nrow - p.H;
  if (r0 < 0) r0 = 0;
  return r0;
}

// Cooperative preload of rows [r0, r0+H) of the pair table into shared memory (same layout).
template <typename FP>
__device__ __forceinline__ void preload_hot(FP* __restrict__ hot, const TabParams<FP>& p, int r0) {
  using P2 = typename Pair2<FP>::type;
  const long long n = (long long)p.H * p.nblk * p.M;
  const P2* __restrict__ src = reinterpret_cast<const P2*>(p.T) + (long long)r0 * p.nblk * p.M;
  P2* dst = reinterpret_cast<P2*>(hot);
  if (p.hot_pad == 0) {
    for (long long e = threadIdx.x; e < n; e += blockDim.x) dst[e] = __ldg(src + e);
  } else {
    const int per_row = p.nblk * p.M;
    const int stride = per_row + p.hot_pad / (int)sizeof(P2);
    for (long long e = threadIdx.x; e < n; e += blockDim.x) {
      const int r = (int)(e / per_row);
      dst[(long long)r * stride + (e - (long long)r * per_row)] = __ldg(src + e);
    }
  }
}

// One lane's share of a 32-neighbour chunk, fetched one work item ahead of its use so that the
// HBM/L2 latency of em_x / em overlaps the previous chunk's arithmetic.
template <typename FP, bool GG>
struct Pre {
  FP x;
  FP e[4];
  FP h[4];  // GG only
  FP zx;    // GG only
};

template <typename FP, bool GG>
__device__ __forceinline__ void load_pre(Pre<FP, GG>& q, const TabParams<FP>& p, long long i, int j0, int lane) {
  const int j = j0 + lane;
  q.x = (FP)0.;
#pragma unroll
  for (int m = 0; m < 4; ++m) q.e[m] = (FP)0.;
  if (GG) {
#pragma unroll
    for (int m = 0; m < 4; ++m) q.h[m] = (FP)0.;
    q.zx = (FP)0.;
  }
  if (i < p.nloc && j < p.nnei) {
    q.x = p.em_x[i * p.ldx_i + (long long)j * p.ldx_j];
    load4(p.em + i * p.ldem_i + (long long)j * 4, p.vec_ok != 0, q.e);
    if (GG) {
      load4(p.dz_em + (i * p.nnei + j) * 4, true, q.h);
      q.zx = p.dz_x[i * p.nnei + j];
    }
  }
}

// Locate up to 32 neighbours [j0, j0+32) in parallel and stage their records (em and dz_dy_dem are
// stored pre-multiplied by the fold multiplicity).  Returns how many of them must be processed
// (the fold entry, if any, is the last one).
template <typename FP, bool GG>
__device__ __forceinline__ int stage_chunk(const TabParams<FP>& p, int j0, FP last, const Pre<FP, GG>& q,
                                           Rec<FP>* __restrict__ rec, RecGG<FP>* __restrict__ rgg,
                                           int lane, bool& done, bool& any_delta) {
  const int j = j0 + lane;
  const bool valid = j < p.nnei;
  const bool fold = valid && p.is_sorted && (q.x == last) && q.e[1] == (FP)0. && q.e[2] == (FP)0. && q.e[3] == (FP)0.;
  const unsigned fm = __ballot_sync(kFull, fold);
  const int nvalid = (p.nnei - j0) < 32 ? (p.nnei - j0) : 32;
  const int nproc = fm ? __ffs(fm) : nvalid;
  done = fm != 0u;
  Rec<FP> r;
  locate(p, q.x, r.xx, r.idx, r.delta);
  any_delta = __any_sync(kFull, lane < nproc && r.delta != (FP)0.);
  r.mult = (fm && lane == nproc - 1) ? (p.nnei - j) : 1;
  const FP mult = (FP)r.mult;
#pragma unroll
  for (int m = 0; m < 4; ++m) r.e[m] = q.e[m] * mult;
  RecGG<FP> g2;
  if (GG) {
#pragma unroll
    for (int m = 0; m < 4; ++m) g2.h[m] = q.h[m] * mult;
    g2.zx = q.zx;
    g2.pad_ = (FP)0.;
  }
  __syncwarp();  // previous chunk's readers are done
  rec[lane] = r;
  if (GG) rgg[lane] = g2;
  __syncwarp();
  return nproc;
}

template <typename FP>
__device__ __forceinline__ FP poly(const FP (&a)[6], FP x) {
  return a[0] + (a[1] + (a[2] + (a[3] + (a[4] + a[5] * x) * x) * x) * x) * x;
}
template <typename FP>
__device__ __forceinline__ FP dpoly(const FP (&a)[6], FP x) {
  return a[1] + ((FP)2. * a[2] + ((FP)3. * a[3] + ((FP)4. * a[4] + (FP)5. * a[5] * x) * x) * x) * x;
}
// value and derivative together by synthetic division (Horner twice): 9 FMAs, no multiplies
template <typename FP>
__device__ __forceinline__ void poly_both(const FP (&a)[6], FP x, FP& g, FP& gd) {
  const FP b4 = a[4] + a[5] * x;
  const FP b3 = a[3] + b4 * x;
  const FP b2 = a[2] + b3 * x;
  const FP b1 = a[1] + b2 * x;
  g = a[0] + b1 * x;
  const FP c4 = b4 + a[5] * x;
  const FP c3 = b3 + c4 * x;
  const FP c2 = b2 + c3 * x;
  gd = b1 + c2 * x;
}

// ------------------------------------------------------------------------------------------
// Fused descriptor epilogue: the warp that just finished atom i holds GR = out[i] ([4][M]) in
// registers, so the se_e2_a descriptor D = (s GR)^T (s GR)[:, :axis], s = 1/nnei
// (deepmd/pt/model/descriptor/se_a.py:843-850) is formed here and never makes a round trip
// through HBM as a separate pass.  It is emitted directly in the operand format of the fitting
// net's first GEMM:
//   mode 1  D as FP [row][M*axis]
//   mode 2  fp64: `nslice` balanced base-256 digit slices of the row's fixed-point image, most significant
//           first, int8 [row][nslice][M*axis] + row_exp[row] (split-integer GEMM on the int8
//           tensor cores, error-free products, fp64-grade sums);
//           fp32: TF32 head and tail, float [row][2][M*axis] (3xTF32 GEMM).
//   mode 3  fp32 only: four int8 digit slices [row][4][M*axis] + row_exp[row] (the fp32 model's operand of the
//           same int8 tensor-core GEMMs).
// The first `axis` columns of s*GR are staged in the (idle) per-warp record buffer and read back
// as broadcasts.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float tf32_round(float x) {
  unsigned u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

template <int NC>
__device__ __forceinline__ void desc_store_split(const TabParams<float>& p, const float (&A)[4][NC], float s2,
                                                 const float* __restrict__ stage, long long row, int lane) {
  const int axis = p.axis, M = p.M;
  float* __restrict__ hi = reinterpret_cast<float*>(p.desc) + row * p.desc_ld;
  float* __restrict__ lo = hi + (long long)M * axis;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int k1 = lane + 32 * c;
    if (k1 < M) {
      for (int k2 = 0; k2 < axis; k2 += 4) {
        float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const float4 b = *reinterpret_cast<const float4*>(stage + m * axis + k2);
          v[0] += A[m][c] * b.x, v[1] += A[m][c] * b.y, v[2] += A[m][c] * b.z, v[3] += A[m][c] * b.w;
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) v[t] *= s2;
        float4 h, l;
        h.x = tf32_round(v[0]), h.y = tf32_round(v[1]), h.z = tf32_round(v[2]), h.w = tf32_round(v[3]);
        l.x = tf32_round(v[0] - h.x), l.y = tf32_round(v[1] - h.y), l.z = tf32_round(v[2] - h.z),
        l.w = tf32_round(v[3] - h.w);
        *reinterpret_cast<float4*>(hi + k1 * axis + k2) = h;
        *reinterpret_cast<float4*>(lo + k1 * axis + k2) = l;
      }
    }
  }
}

// mode 3 (fp32): four balanced base-256 digit slices of the row's 32-bit fixed-point image + row exponent -- the
// A operand of the int8 tensor-core fitting GEMMs (csrc/fit_tc.cu, NS = 4: 31 fraction bits below the row maximum).
template <int NC>
__device__ __forceinline__ void desc_store_split_i8(const TabParams<float>& p, const float (&A)[4][NC], float s2,
                                                    const float* __restrict__ stage, long long row, int lane) {
  const int M = p.M;
  float r2 = 0.f;  // Cauchy-Schwarz: |D[k1][k2]| <= max_k |A[:,k]|^2
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const float n2 = A[0][c] * A[0][c] + A[1][c] * A[1][c] + A[2][c] * A[2][c] + A[3][c] * A[3][c];
    r2 = (lane + 32 * c < M && n2 > r2) ? n2 : r2;
  }
  r2 *= s2;
  int e = (int)((__float_as_uint(r2) >> 23) & 0xffu) - 127;
  e = __reduce_max_sync(kFull, e);
  int E = e + 2;  // |D| < 2^(E-1)
  E = E < p.desc_min_exp ? p.desc_min_exp : E;
  E = E < -90 ? -90 : (E > 100 ? 100 : E);
  if (lane == 0) p.row_exp[row] = E;
  const float up = s2 * __uint_as_float((unsigned)(127 + 31 - E) << 23);
  signed char* __restrict__ base = reinterpret_cast<signed char*>(p.desc) + row * p.desc_ld;
  const long long K = p.desc_slice_stride;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int k1 = lane + 32 * c;
    if (k1 < M) {
      unsigned u[16];
#pragma unroll
      for (int t = 0; t < 16; t += 4) {
        float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const float4 b = *reinterpret_cast<const float4*>(stage + m * 16 + t);
          v[0] += A[m][c] * b.x, v[1] += A[m][c] * b.y, v[2] += A[m][c] * b.z, v[3] += A[m][c] * b.w;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) u[t + q] = (unsigned)__float2int_rn(v[q] * up) + 0x80808080u;
      }
#pragma unroll
      for (int s = 0; s < 4; ++s) {  // slice s = byte 3-s of the image, top bit flipped = signed digit
        const unsigned kb = 3u - (unsigned)s;
        const unsigned sel = kb | ((4u + kb) << 4);
        unsigned w[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const unsigned t0 = __byte_perm(u[4 * g], u[4 * g + 1], sel);
          const unsigned t1 = __byte_perm(u[4 * g + 2], u[4 * g + 3], sel);
          w[g] = __byte_perm(t0, t1, 0x5410) ^ 0x80808080u;
        }
        *reinterpret_cast<uint4*>(base + s * K + (long long)k1 * 16) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
  }
}
template <int NC>
__device__ __forceinline__ void desc_store_split_i8(const TabParams<double>&, const double (&)[4][NC], double,
                                                    const double*, long long, int) {}

// The production slice count (6) with everything known at compile time: the generic version below selects bytes
// with run-time selectors and costs ~4300 warp instructions per atom, this one ~1000.  The fixed-point image is
// formed by one FMA against 2^52 + 2^51 + bias (the integer lands in the low mantissa bits, rounded to nearest):
// no F2I, no 64-bit adds.
template <int NS, int NC>
__device__ __forceinline__ void desc_store_split_ns(const TabParams<double>& p, const double (&A)[4][NC], double s2,
                                                    const double* __restrict__ stage, long long row, int lane) {
  static_assert(NS >= 5 && NS <= 6, "the image must reach into the high word and fit in 48 bits");
  const int M = p.M;
  double r2 = 0.;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const double n2 = A[0][c] * A[0][c] + A[1][c] * A[1][c] + A[2][c] * A[2][c] + A[3][c] * A[3][c];
    r2 = (lane + 32 * c < M && n2 > r2) ? n2 : r2;
  }
  r2 *= s2;
  int e = ((__double2hiint(r2) >> 20) & 0x7ff) - 1023;
  e = __reduce_max_sync(kFull, e);
  int E = e + 2;  // |D| < 2^(E-1)
  E = E < p.desc_min_exp ? p.desc_min_exp : E;
  E = E < -900 ? -900 : (E > 900 ? 900 : E);
  if (lane == 0) p.row_exp[row] = E;
  constexpr int P = 7 + 8 * (NS - 1);
  const double up = s2 * __hiloint2double((1023 + P - E) << 20, 0);
  unsigned long long bias = 0;
#pragma unroll
  for (int k = 0; k < NS; ++k) bias = bias * 256ull + 128ull;
  const double magic = 6755399441055744.0 + (double)bias;  // 2^52 + 2^51 + bias, exact
  signed char* __restrict__ base = reinterpret_cast<signed char*>(p.desc) + row * p.desc_ld;
  const long long K = p.desc_slice_stride;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int k1 = lane + 32 * c;
    if (k1 < M) {
      signed char* __restrict__ dst = base + (long long)k1 * 16;
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        unsigned lo[8], hi[8];
#pragma unroll
        for (int t = 0; t < 8; t += 2) {
          double v0 = 0., v1 = 0.;
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            const double2 b = *reinterpret_cast<const double2*>(stage + m * 16 + half * 8 + t);
            v0 += A[m][c] * b.x, v1 += A[m][c] * b.y;
          }
          const double f0 = fma(v0, up, magic), f1 = fma(v1, up, magic);
          lo[t] = (unsigned)__double2loint(f0), hi[t] = (unsigned)__double2hiint(f0);
          lo[t + 1] = (unsigned)__double2loint(f1), hi[t + 1] = (unsigned)__double2hiint(f1);
        }
#pragma unroll
        for (int s = 0; s < NS; ++s) {  // slice s = byte NS-1-s of the image, top bit flipped = signed digit
          constexpr unsigned kSel[4] = {0x40u, 0x51u, 0x62u, 0x73u};
          const int kb = NS - 1 - s;
          const unsigned sel = kSel[kb & 3];
          const unsigned(&src)[8] = kb < 4 ? lo : hi;
          const unsigned a0 = __byte_perm(src[0], src[1], sel), a1 = __byte_perm(src[2], src[3], sel);
          const unsigned a2 = __byte_perm(src[4], src[5], sel), a3 = __byte_perm(src[6], src[7], sel);
          *reinterpret_cast<uint2*>(dst + s * K + half * 8) =
              make_uint2(__byte_perm(a0, a1, 0x5410) ^ 0x80808080u, __byte_perm(a2, a3, 0x5410) ^ 0x80808080u);
        }
      }
    }
  }
}

template <int NC>
__device__ __forceinline__ void desc_store_split(const TabParams<double>& p, const double (&A)[4][NC], double s2,
                                                 const double* __restrict__ stage, long long row, int lane) {
  // axis == 16 (checked on the host): one lane owns the 16 contiguous k2 of each of its channels.
  if (p.nslice == 6) return desc_store_split_ns<6, NC>(p, A, s2, stage, row, lane);
  const int M = p.M, ns = p.nslice;
  // row scale from the Cauchy-Schwarz bound |D[k1][k2]| <= max_k |A[:,k]|^2 (attained on the diagonal)
  double r2 = 0.;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const double n2 = A[0][c] * A[0][c] + A[1][c] * A[1][c] + A[2][c] * A[2][c] + A[3][c] * A[3][c];
    r2 = (lane + 32 * c < M && n2 > r2) ? n2 : r2;
  }
  r2 *= s2;
  int e = ((__double2hiint(r2) >> 20) & 0x7ff) - 1023;
  e = __reduce_max_sync(kFull, e);
  int E = e + 2;  // |D| < 2^(E-1)
  E = E < p.desc_min_exp ? p.desc_min_exp : E;
  E = E < -900 ? -900 : (E > 900 ? 900 : E);
  if (lane == 0) p.row_exp[row] = E;
  const int P = 7 + 8 * (ns - 1);  // fixed-point fraction bits
  const double up = s2 * __hiloint2double((1023 + P - E) << 20, 0);
  // bias that makes every base-256 digit non-negative: byte k of the image = digit_k + 128
  unsigned long long bias = 0;
  for (int k = 0; k < ns; ++k) bias = bias * 256ull + 128ull;
  signed char* __restrict__ base = reinterpret_cast<signed char*>(p.desc) + row * p.desc_ld;
  const long long K = p.desc_slice_stride;
#pragma unroll
  for (int c = 0; c < NC; ++c) {  // (unrolled: a runtime channel index would push acc[][] into local memory)
    const int k1 = lane + 32 * c;
    if (k1 < M) {
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        unsigned lo[8], hi[8];
#pragma unroll
        for (int t = 0; t < 8; t += 2) {
          double v0 = 0., v1 = 0.;
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            const double2 b = *reinterpret_cast<const double2*>(stage + m * 16 + half * 8 + t);
            v0 += A[m][c] * b.x, v1 += A[m][c] * b.y;
          }
          const unsigned long long u0 = (unsigned long long)__double2ll_rn(v0 * up) + bias;
          const unsigned long long u1 = (unsigned long long)__double2ll_rn(v1 * up) + bias;
          lo[t] = (unsigned)u0, hi[t] = (unsigned)(u0 >> 32);
          lo[t + 1] = (unsigned)u1, hi[t + 1] = (unsigned)(u1 >> 32);
        }
        for (int s = 0; s < ns; ++s) {  // slice s = byte ns-1-s of the image, top bit flipped = signed digit
          const int kb = ns - 1 - s;
          const unsigned sel = (unsigned)(kb & 3) | ((unsigned)(4 + (kb & 3)) << 4);
          unsigned w[2];
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const unsigned t0 = kb < 4 ? __byte_perm(lo[4 * g], lo[4 * g + 1], sel) : __byte_perm(hi[4 * g], hi[4 * g + 1], sel);
            const unsigned t1 = kb < 4 ? __byte_perm(lo[4 * g + 2], lo[4 * g + 3], sel)
                                       : __byte_perm(hi[4 * g + 2], hi[4 * g + 3], sel);
            w[g] = __byte_perm(t0, t1, 0x5410) ^ 0x80808080u;
          }
          *reinterpret_cast<uint2*>(base + s * K + (long long)k1 * 16 + half * 8) = make_uint2(w[0], w[1]);
        }
      }
    }
  }
}

template <typename FP, int NC>
__device__ __forceinline__ void desc_epilogue(const TabParams<FP>& p, const FP (&acc)[4][NC], long long i,
                                              int lane, FP* __restrict__ stage) {
  const int axis = p.axis, M = p.M;
  const long long row = p.desc_row ? (long long)p.desc_row[i] : i;
  const FP s2 = p.desc_scale * p.desc_scale;
  __syncwarp();  // the chunk's records are no longer needed
  if (lane < axis) {
#pragma unroll
    for (int m = 0; m < 4; ++m) stage[m * axis + lane] = acc[m][0];
  }
  __syncwarp();
  if (p.desc_mode == 3) {
    desc_store_split_i8<NC>(p, acc, s2, stage, row, lane);
  } else if (p.desc_mode == 2) {
    desc_store_split<NC>(p, acc, s2, stage, row, lane);
  } else {
    FP* __restrict__ d = reinterpret_cast<FP*>(p.desc) + row * p.desc_ld;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int k1 = lane + 32 * c;
      if (k1 < M) {
        for (int k2 = 0; k2 < axis; ++k2) {
          const FP v = acc[0][c] * stage[k2] + acc[1][c] * stage[axis + k2] + acc[2][c] * stage[2 * axis + k2] +
                       acc[3][c] * stage[3 * axis + k2];
          st_cs(d + (long long)k1 * axis + k2, v * s2);
        }
      }
    }
  }
}

// se_atten: two_embed ([nloc][nnei][M], 800 B per neighbour for M = 100 in fp64) is streamed once and is the
// dominant HBM traffic of the gated op.  Its loads sit on the per-neighbour critical path, so the row of the
// neighbour kTwoAhead slots ahead is pulled into L2 while the current one is evaluated (one 32-byte sector per
// lane; no registers, no shared memory): the kernel then waits for L2, not for HBM.
constexpr int kTwoAhead = 8;
// ... and, when the rows are 16-byte aligned, copied asynchronously (cp.async, no registers) into a per-warp ring of
// kRing rows in shared memory kRing-1 neighbours ahead, so that the gate value is a shared-memory read.
constexpr int kRing = 8;  // (a power of two)
template <typename FP>
__device__ __forceinline__ void ring_issue(FP* __restrict__ ring, int Mp, const FP* __restrict__ two, long long row,
                                           int slot_row, int M, int lane) {
  const char* src = reinterpret_cast<const char*>(two + row * (long long)M);
  char* dst = reinterpret_cast<char*>(ring + (slot_row % kRing) * Mp);
  const int pieces = M * (int)sizeof(FP) / 16;
  for (int q = lane; q < pieces; q += 32) __pipeline_memcpy_async(dst + 16 * q, src + 16 * q, 16);
}
template <typename FP>
__device__ __forceinline__ void prefetch_two(const FP* __restrict__ two, long long row, int M, int lane) {
  const int e = lane * (32 / (int)sizeof(FP));
  if (e < M) asm volatile("prefetch.global.L2 [%0];" ::"l"(two + row * (long long)M + e));
}

extern __shared__ __align__(16) unsigned char tab_smem[];

// ------------------------------------------------------------------------------------------
// forward (GG=false) and second-order backward (GG=true): both accumulate a [4][M] tile.
// grid (x: persistent over atoms, y: block of 32*NC channels); block = nw warps.
// smem: hot[H][3][M] pairs | Rec[nw][32] | RecGG[nw][32] (GG only)
// ------------------------------------------------------------------------------------------
// Threads per CTA: the fp64 kernels with the descriptor epilogue run 12 warps with up to 168 registers (no spills;
// measured 6.28 -> 6.11 ms at 332 k atoms), everything else 16 warps at 128 registers (12 warps cost the plain fp64
// and the fp32 forward 8-10 %).
template <typename FP, bool DESC>
constexpr int fwd_threads() {
  return (sizeof(FP) == 8 && DESC) ? 384 : 512;
}

template <typename FP, int NC, bool TWO, bool GG, bool DESC = false, bool CM = false>
__global__ void __launch_bounds__(fwd_threads<FP, DESC>()) k_tab_fwd(const __grid_constant__ TabParams<FP> p) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nw = blockDim.x >> 5;
  FP* hot = reinterpret_cast<FP*>(tab_smem);
  Rec<FP>* rec = reinterpret_cast<Rec<FP>*>(hot + p.hot_elems) + warp * 32;
  RecGG<FP>* rgg = reinterpret_cast<RecGG<FP>*>(reinterpret_cast<Rec<FP>*>(hot + p.hot_elems) + nw * 32) +
                   (GG ? warp * 32 : 0);
  const int Mp = (p.M * (int)sizeof(FP) + 15) / 16 * 16 / (int)sizeof(FP);
  FP* ring = reinterpret_cast<FP*>(reinterpret_cast<RecGG<FP>*>(reinterpret_cast<Rec<FP>*>(hot + p.hot_elems) + nw * 32) +
                                   (GG ? nw * 32 : 0)) +
             warp * kRing * Mp;
  const bool ring_on = TWO && !GG && p.two_ring != 0;
  const int c0 = blockIdx.y * 32 * NC;
  const int r0 = hot_window_start(p);
  preload_hot(hot, p, r0);
  __syncthreads();

  int kc[NC];  // this lane's channels, clamped: lanes beyond M recompute channel M-1 and never store
  int ob[NC];  // their byte offsets inside a coefficient-pair block
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    kc[c] = (c0 + lane + 32 * c < p.M) ? c0 + lane + 32 * c : p.M - 1;
    ob[c] = kc[c] * (int)sizeof(typename Pair2<FP>::type);
  }

  const long long stride = (long long)gridDim.x * nw;
  long long i = (long long)blockIdx.x * nw + warp;
  int j0 = 0;
  Pre<FP, GG> pre;
  load_pre(pre, p, i, 0, lane);
  FP last = i < p.nloc ? p.em_x[i * p.ldx_i + (long long)(p.nnei - 1) * p.ldx_j] : (FP)0.;
  FP acc[4][NC];
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int c = 0; c < NC; ++c)
      acc[m][c] = (p.accumulate && i < p.nloc) ? p.out[(i * 4 + m) * (long long)p.M + kc[c]] : (FP)0.;
  FP a[NC][6];
  float af[CM ? NC : 1][3];  // compressed coefficients: a3, a4, a5 of the cached row stay fp32
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int q = 0; q < 6; ++q) a[c][q] = (FP)0.;
#pragma unroll
  for (int c = 0; c < (CM ? NC : 1); ++c) af[c][0] = af[c][1] = af[c][2] = 0.f;
  int cur_row = -1;  // the coefficient registers stay valid across atoms

  while (i < p.nloc) {
    bool done, any_delta;
    const int nproc = stage_chunk<FP, GG>(p, j0, last, pre, rec, rgg, lane, done, any_delta);
    // next work item, fetched now, consumed after this chunk's arithmetic
    const bool atom_end = done || j0 + 32 >= p.nnei;
    const long long ni = atom_end ? i + stride : i;
    const int nj0 = atom_end ? 0 : j0 + 32;
    load_pre(pre, p, ni, nj0, lane);
    FP nlast = last;
    if (atom_end && ni < p.nloc) nlast = p.em_x[ni * p.ldx_i + (long long)(p.nnei - 1) * p.ldx_j];

    if (!GG && !TWO && !any_delta) {
      // the common case: nothing to gate, nobody outside [lower, max): 9 FMAs per channel, no branches
      // Software pipeline: the coefficients of neighbour jj are consumed by the polynomial, THEN the row of
      // neighbour jj+1 is requested (same registers), and the 16 accumulation FMAs of neighbour jj run while that
      // fetch is in flight; the row index of jj+1 is read one iteration ahead of its comparison.
      auto fetch = [&](int row) {
        cur_row = row;
        if (CM && sizeof(FP) == 4)
          fetch_row_c32<NC>(a, hot, p.T, row, r0, p.H, p.M, ob);
        else if (CM)
          fetch_row_cmf(a, af, hot, p.T, row, r0, p.H, p.M, ob, p.a5_inv);
        else
          fetch_row<FP, NC>(a, hot, p.T, row, r0, p.H, p.M, ob);
      };
      if (nproc > 0 && rec[0].idx != cur_row) fetch(rec[0].idx);  // warp-uniform
#pragma unroll 2
      for (int jj = 0; jj < nproc; ++jj) {
        const Rec<FP>& r = rec[jj];
        const int row_n = rec[jj + 1 < nproc ? jj + 1 : jj].idx;
        const FP xx = r.xx;
        const float xf = (float)xx;
        const FP e0 = r.e[0], e1 = r.e[1], e2 = r.e[2], e3 = r.e[3];
        FP g[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c)
          g[c] = (CM && sizeof(FP) == 4) ? poly3(a[c], xx) : (CM ? poly_cm(a[c], af[CM ? c : 0], xx, xf) : poly(a[c], xx));
        if (row_n != cur_row) fetch(row_n);  // warp-uniform
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          acc[0][c] += e0 * g[c];
          acc[1][c] += e1 * g[c];
          acc[2][c] += e2 * g[c];
          acc[3][c] += e3 * g[c];
        }
      }
    } else if (!GG && TWO && ring_on && !any_delta) {
      // se_atten, nobody outside [lower, max): the gate row of every neighbour arrives through the cp.async ring;
      // same arithmetic as the general path below, without its per-neighbour branches and 64-bit index math
      const FP* __restrict__ two_base = p.two + (i * p.nnei + j0) * (long long)p.M;
      __syncwarp();
      for (int jp = 0; jp < kRing - 1; ++jp) {
        if (jp < nproc) ring_issue(ring, Mp, two_base, jp, jp, p.M, lane);
        __pipeline_commit();
      }
      for (int jj = 0; jj < nproc; ++jj) {
        __syncwarp();  // slot (jj-1) % kRing was read in the previous iteration
        const int jp = jj + kRing - 1;
        if (jp < nproc) ring_issue(ring, Mp, two_base, jp, jp, p.M, lane);
        __pipeline_commit();
        __pipeline_wait_prior(kRing - 1);  // the group of row jj has landed
        __syncwarp();
        const Rec<FP>& r = rec[jj];
        const int row = r.idx;
        if (row != cur_row) {  // warp-uniform
          cur_row = row;
          fetch_row<FP, NC>(a, hot, p.T, row, r0, p.H, p.M, ob);
        }
        const FP xx = r.xx;
        const FP e0 = r.e[0], e1 = r.e[1], e2 = r.e[2], e3 = r.e[3];
        const FP* __restrict__ tw = ring + (jj & (kRing - 1)) * Mp;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          FP g = poly(a[c], xx);
          g = g * tw[kc[c]] + g;
          acc[0][c] += e0 * g;
          acc[1][c] += e1 * g;
          acc[2][c] += e2 * g;
          acc[3][c] += e3 * g;
        }
      }
    } else if (!GG && TWO && p.gate_tt != nullptr && !any_delta) {
      // se_atten with the pair-indexed gate, nobody outside [lower, max): the common case of config 5.  The
      // neighbour's type-pair row index and switch value ride in its record (`delta` is zero for everybody here and
      // `mult` is already folded into e[]), so the loop reads them as shared-memory broadcasts; the (ntypes+1)^2-row
      // pair table stays L1-resident.  Same pipelining as the plain loop above.
      {
        const long long gj = i * p.nnei + j0 + lane;
        FP gsw = (FP)0.;
        int gpr = 0;
        if (lane < nproc) gsw = p.gate_sw[gj], gpr = p.gate_pair[gj];
        rec[lane].delta = gsw;
        rec[lane].mult = gpr;
        __syncwarp();
      }
      auto fetch = [&](int row) {
        cur_row = row;
        if (CM && sizeof(FP) == 4)
          fetch_row_c32<NC>(a, hot, p.T, row, r0, p.H, p.M, ob);
        else if (CM)
          fetch_row_cmf(a, af, hot, p.T, row, r0, p.H, p.M, ob, p.a5_inv);
        else
          fetch_row<FP, NC>(a, hot, p.T, row, r0, p.H, p.M, ob);
      };
      if (nproc > 0 && rec[0].idx != cur_row) fetch(rec[0].idx);  // warp-uniform
#pragma unroll 2
      for (int jj = 0; jj < nproc; ++jj) {
        const Rec<FP>& r = rec[jj];
        const int row_n = rec[jj + 1 < nproc ? jj + 1 : jj].idx;
        const FP xx = r.xx;
        const float xf = (float)xx;
        const FP gs = r.delta;
        const FP* __restrict__ trow = p.gate_tt + (long long)r.mult * p.M;
        const FP e0 = r.e[0], e1 = r.e[1], e2 = r.e[2], e3 = r.e[3];
        FP g[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const FP v = (CM && sizeof(FP) == 4) ? poly3(a[c], xx) : (CM ? poly_cm(a[c], af[CM ? c : 0], xx, xf) : poly(a[c], xx));
          g[c] = v * (__ldg(trow + kc[c]) * gs) + v;
        }
        if (row_n != cur_row) fetch(row_n);  // warp-uniform
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          acc[0][c] += e0 * g[c];
          acc[1][c] += e1 * g[c];
          acc[2][c] += e2 * g[c];
          acc[3][c] += e3 * g[c];
        }
      }
    } else {
      if (ring_on) {
        __syncwarp();
        for (int jp = 0; jp < kRing - 1; ++jp) {
          if (jp < nproc) ring_issue(ring, Mp, p.two, i * p.nnei + j0 + jp, jp, p.M, lane);
          __pipeline_commit();
        }
      } else if (TWO && p.two) {
        for (int jp = 0; jp < kTwoAhead && jp < nproc; ++jp) prefetch_two(p.two, i * p.nnei + j0 + jp, p.M, lane);
      }
      for (int jj = 0; jj < nproc; ++jj) {
        if (ring_on) {
          __syncwarp();  // slot (jj-1) % kRing was read in the previous iteration
          const int jp = jj + kRing - 1;
          if (jp < nproc) ring_issue(ring, Mp, p.two, i * p.nnei + j0 + jp, jp, p.M, lane);
          __pipeline_commit();
          __pipeline_wait_prior(kRing - 1);  // the group of row jj has landed
          __syncwarp();
        } else if (TWO && p.two && jj + kTwoAhead < nproc) {
          prefetch_two(p.two, i * p.nnei + j0 + jj + kTwoAhead, p.M, lane);
        }
        // pair-indexed gate: one table row index and one switch value per neighbour (broadcast loads)
        const FP* __restrict__ gate_row = nullptr;
        FP gate_s = (FP)0.;
        if (TWO && p.gate_tt) {
          const long long gj = i * p.nnei + j0 + jj;
          gate_row = p.gate_tt + (long long)p.gate_pair[gj] * p.M;
          gate_s = p.gate_sw[gj];
        }
        const Rec<FP>& r = rec[jj];
        const int row = r.idx;
        if (row != cur_row) {  // warp-uniform
          cur_row = row;
          if (CM && sizeof(FP) == 4)
            fetch_row_c32<NC>(a, hot, p.T, row, r0, p.H, p.M, ob);  // a4 = a5 = 0 from the initialisation
          else if (CM)
            fetch_row_cmf(a, af, hot, p.T, row, r0, p.H, p.M, ob, p.a5_inv);
          else
            fetch_row<FP, NC>(a, hot, p.T, row, r0, p.H, p.M, ob);
        }
        const FP xx = r.xx;
        const FP dl = r.delta;
        FP e[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) e[m] = r.e[m];
        FP h[4];
        FP zx = (FP)0.;
        if (GG) {
#pragma unroll
          for (int m = 0; m < 4; ++m) h[m] = rgg[jj].h[m];
          zx = rgg[jj].zx;
        }
        const long long two_off = (i * p.nnei + j0 + jj) * (long long)p.M;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          FP g, gd = (FP)0.;
          FP w6[6];  // this (rare) path evaluates in fp64 throughout
#pragma unroll
          for (int k = 0; k < 6; ++k) w6[k] = (CM && sizeof(FP) == 8 && k >= 3) ? (FP)af[CM ? c : 0][k - 3] : a[c][k];
          if (GG) {
            poly_both(w6, xx, g, gd);
            g += gd * dl;
          } else {
            g = poly(w6, xx);
            if (dl != (FP)0.) {
              gd = dpoly(w6, xx);
              g += gd * dl;
            }
          }
          if (GG) {
            FP two_grad = (FP)0.;
            if (TWO) {
              const FP t = p.two[two_off + kc[c]];
              two_grad = p.dz_two[two_off + kc[c]] * g;
              g += g * t;
              gd += gd * t;
            }
            const FP sgl = zx * gd + two_grad;
#pragma unroll
            for (int m = 0; m < 4; ++m) acc[m][c] += g * h[m] + sgl * e[m];
          } else {
            if (TWO) {
              const FP t = gate_row ? gate_row[kc[c]] * gate_s
                                    : (ring_on ? ring[(jj % kRing) * Mp + kc[c]] : p.two[two_off + kc[c]]);
              g = g * t + g;
            }
#pragma unroll
            for (int m = 0; m < 4; ++m) acc[m][c] += e[m] * g;
          }
        }
      }
    }
    if (atom_end) {
      // (accumulate: acc was seeded with the previous contents of out when this atom started)
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        if (c0 + lane + 32 * c < p.M) {
#pragma unroll
          for (int m = 0; m < 4; ++m) p.out[(i * 4 + m) * (long long)p.M + kc[c]] = acc[m][c];
        }
      }
      if (DESC) desc_epilogue<FP, NC>(p, acc, i, lane, reinterpret_cast<FP*>(rec));
#pragma unroll
      for (int c = 0; c < NC; ++c) {
#pragma unroll
        for (int m = 0; m < 4; ++m)
          acc[m][c] = (p.accumulate && ni < p.nloc) ? p.out[(ni * 4 + m) * (long long)p.M + kc[c]] : (FP)0.;
      }
    }
    i = ni;
    j0 = nj0;
    last = nlast;
  }
}

// ------------------------------------------------------------------------------------------
// first-order backward
// ------------------------------------------------------------------------------------------
// v[0..3] per lane -> lane l returns the warp-wide sum of v[l >> 3].
template <typename FP>
__device__ __forceinline__ FP reduce_scatter4(FP (&v)[4], int lane) {
#pragma unroll
  for (int s = 16, n = 4; s >= 8; s >>= 1, n >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int t = 0; t < n / 2; ++t) {
      const FP send = up ? v[t] : v[t + n / 2];
      const FP keep = up ? v[t + n / 2] : v[t];
      v[t] = keep + __shfl_xor_sync(kFull, send, s);
    }
  }
  FP r = v[0];
  r += __shfl_xor_sync(kFull, r, 4);
  r += __shfl_xor_sync(kFull, r, 2);
  r += __shfl_xor_sync(kFull, r, 1);
  return r;
}

// One warp per atom, all channels of the atom in this warp (blocks of 32*NC channels; for
// M <= 32*NC the coefficient registers persist across neighbours and atoms).
// smem: hot[H][3][M] pairs | Rec[nw][32]
template <typename FP, int NC, bool TWO, bool CM = false>
__global__ void __launch_bounds__(384) k_tab_grad(const __grid_constant__ TabParams<FP> p) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nw = blockDim.x >> 5;
  FP* hot = reinterpret_cast<FP*>(tab_smem);
  Rec<FP>* rec = reinterpret_cast<Rec<FP>*>(hot + p.hot_elems) + warp * 32;
  const int M = p.M;
  const int Mp = (M * (int)sizeof(FP) + 15) / 16 * 16 / (int)sizeof(FP);
  FP* ring = reinterpret_cast<FP*>(reinterpret_cast<Rec<FP>*>(hot + p.hot_elems) + nw * 32) + warp * kRing * Mp;
  const bool ring_on = TWO && p.two_ring != 0;
  const int r0 = hot_window_start(p);
  preload_hot(hot, p, r0);
  __syncthreads();
  const bool single = M <= 32 * NC;
  const bool fuse_x = p.dy_dem_x == nullptr;  // em_x IS component 0 of em: add its gradient there

  const long long stride = (long long)gridDim.x * nw;
  long long i = (long long)blockIdx.x * nw + warp;
  int j0 = 0;
  Pre<FP, false> pre;
  load_pre(pre, p, i, 0, lane);
  FP last = i < p.nloc ? p.em_x[i * p.ldx_i + (long long)(p.nnei - 1) * p.ldx_j] : (FP)0.;
  FP dyr[4][NC];
  FP a[NC][6];
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int q = 0; q < 6; ++q) a[c][q] = (FP)0.;
  int cur_row = -1;
  int kc[NC];  // clamped channel of (lane, c) in the current channel block
  int ob[NC];  // its byte offset inside a coefficient-pair block

  while (i < p.nloc) {
    const FP* __restrict__ dyi = p.dy + i * 4 * (long long)M;
    if (j0 == 0 && single) {
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        kc[c] = (lane + 32 * c < M) ? lane + 32 * c : M - 1;
        ob[c] = kc[c] * (int)sizeof(typename Pair2<FP>::type);
#pragma unroll
        for (int m = 0; m < 4; ++m) dyr[m][c] = (lane + 32 * c < M) ? dyi[(long long)m * M + lane + 32 * c] : (FP)0.;
      }
    }
    bool done, any_delta;
    const int nproc = stage_chunk<FP, false>(p, j0, last, pre, rec, nullptr, lane, done, any_delta);
    const bool atom_end = done || j0 + 32 >= p.nnei;
    const long long ni = atom_end ? i + stride : i;
    const int nj0 = atom_end ? 0 : j0 + 32;
    load_pre(pre, p, ni, nj0, lane);
    FP nlast = last;
    if (atom_end && ni < p.nloc) nlast = p.em_x[ni * p.ldx_i + (long long)(p.nnei - 1) * p.ldx_j];

    FP* __restrict__ gem = p.dy_dem + i * p.ldem_i;
    if (ring_on) {
      __syncwarp();
      for (int jp = 0; jp < kRing; ++jp) {
        if (jp < nproc) ring_issue(ring, Mp, p.two, i * p.nnei + j0 + jp, jp, M, lane);
        __pipeline_commit();
      }
    } else if (TWO && p.two) {
      for (int jp = 0; jp < kTwoAhead && jp < nproc; ++jp) prefetch_two(p.two, i * p.nnei + j0 + jp, M, lane);
    }
    for (int b = 0; b < nproc; b += 4) {
      if (ring_on) {
        __pipeline_wait_prior(4);  // 8 + 4*(b/4) groups committed, rows b..b+3 are among the first 4*(b/4 + 1)
        __syncwarp();
      } else if (TWO && p.two) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (b + kTwoAhead + u < nproc) prefetch_two(p.two, i * p.nnei + j0 + b + kTwoAhead + u, M, lane);
      }
      FP v[16];
      FP vx[4];
      FP vq[4] = {(FP)0., (FP)0., (FP)0., (FP)0.};  // gate mode: dE/d(sw) of the four neighbours
      if (!single || (nproc - b) < 4) {  // the fast path below initialises by assignment
#pragma unroll
        for (int t = 0; t < 16; ++t) v[t] = (FP)0.;
#pragma unroll
        for (int t = 0; t < 4; ++t) vx[t] = (FP)0.;
      }
      const bool assign0 = single && (nproc - b) >= 4;
      for (int kb = 0; kb < M; kb += 32 * NC) {
        if (!single) {
          cur_row = -1;
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            const int k = kb + lane + 32 * c;
            kc[c] = k < M ? k : M - 1;
            ob[c] = kc[c] * (int)sizeof(typename Pair2<FP>::type);
#pragma unroll
            for (int m = 0; m < 4; ++m) dyr[m][c] = k < M ? dyi[(long long)m * M + k] : (FP)0.;
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (b + u < nproc) {  // warp-uniform
            const Rec<FP>& rc = rec[b + u];
            if (rc.idx != cur_row) {
              cur_row = rc.idx;
              if (CM)
                fetch_row_c32<NC>(a, hot, p.T, cur_row, r0, p.H, M, ob);
              else
                fetch_row<FP, NC>(a, hot, p.T, cur_row, r0, p.H, M, ob);
            }
            const FP xx = rc.xx;
            const FP dl = rc.delta;
            const FP e0 = rc.e[0], e1 = rc.e[1], e2 = rc.e[2], e3 = rc.e[3];  // pre-multiplied by the fold multiplicity
            const FP* __restrict__ gate_row = nullptr;
            FP gate_s = (FP)0.;
            if (TWO && p.gate_tt) {
              const long long gj = i * p.nnei + j0 + b + u;
              gate_row = p.gate_tt + (long long)p.gate_pair[gj] * M;
              gate_s = p.gate_sw[gj];
            }
#pragma unroll
            for (int c = 0; c < NC; ++c) {
              FP g, gd;
              if (CM)
                poly3_both(a[c], xx, g, gd);
              else
                poly_both(a[c], xx, g, gd);
              g += gd * dl;
              const FP dot = e0 * dyr[0][c] + e1 * dyr[1][c] + e2 * dyr[2][c] + e3 * dyr[3][c];
              if (TWO) {
                const int k = kb + lane + 32 * c;
                if (k < M) {
                  if (gate_row) {  // dE/d(two_embed) is never stored: contract it with the type-pair row right here
                    const FP tt = gate_row[k];
                    vq[u] += g * dot * tt;
                    const FP t = tt * gate_s;
                    g = g * t + g;
                    gd += t * gd;
                  } else {
                    const long long to = (i * p.nnei + j0 + b + u) * (long long)M + k;
                    const FP t = ring_on ? ring[((b + u) % kRing) * Mp + k] : p.two[to];
                    p.dy_dtwo[to] = g * dot;
                    g = g * t + g;
                    gd += t * gd;
                  }
                }
              }
              if (c == 0 && assign0) {  // first channel group of a full batch: start the sums here
                vx[u] = gd * dot;
                v[4 * u + 0] = g * dyr[0][c];
                v[4 * u + 1] = g * dyr[1][c];
                v[4 * u + 2] = g * dyr[2][c];
                v[4 * u + 3] = g * dyr[3][c];
              } else {
                vx[u] += gd * dot;
                v[4 * u + 0] += g * dyr[0][c];
                v[4 * u + 1] += g * dyr[1][c];
                v[4 * u + 2] += g * dyr[2][c];
                v[4 * u + 3] += g * dyr[3][c];
              }
            }
          }
        }
      }
      if (ring_on) {
        __syncwarp();  // rows b..b+3 consumed: their slots take rows b+8..b+11
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int jp = b + kRing + u;
          if (jp < nproc) ring_issue(ring, Mp, p.two, i * p.nnei + j0 + jp, jp, M, lane);
          __pipeline_commit();
        }
      }
      const FP tot = reduce_scatter16(v, lane);   // (u, m) = (lane >> 3, (lane >> 1) & 3)
      const FP totx = reduce_scatter4(vx, lane);  // u = lane >> 3; carries the multiplicity through e
      const int u = lane >> 3;
      if (TWO && p.gate_q) {
        const FP totq = reduce_scatter4(vq, lane);
        if (b + u < nproc && (lane & 7) == 0) p.gate_q[i * p.nnei + j0 + b + u] = totq;
      }
      if (b + u < nproc && (lane & 1) == 0) {
        const FP mult = (FP)rec[b + u].mult;
        const int m = (lane >> 1) & 3;
        const int j = j0 + b + u;
        if (fuse_x) {
          gem[(long long)j * 4 + m] = m == 0 ? tot * mult + totx : tot * mult;
        } else {
          gem[(long long)j * 4 + m] = tot * mult;
          if (m == 0) p.dy_dem_x[i * p.ldx_i + (long long)j * p.ldx_j] = totx;
        }
      }
    }
    if (atom_end) {
      // everything behind the processed slots is zero (tabulate.cc zero-fills the outputs first)
      const int jend = j0 + nproc;
      for (int j = jend + lane; j < p.nnei; j += 32) {
        if (!fuse_x) p.dy_dem_x[i * p.ldx_i + (long long)j * p.ldx_j] = (FP)0.;
#pragma unroll
        for (int m = 0; m < 4; ++m) gem[(long long)j * 4 + m] = (FP)0.;
      }
      if (TWO && p.gate_q) {
        for (int j = jend + lane; j < p.nnei; j += 32) p.gate_q[i * p.nnei + j] = (FP)0.;
      } else if (TWO) {
        for (long long e = (long long)jend * M + lane; e < (long long)p.nnei * M; e += 32)
          p.dy_dtwo[i * p.nnei * (long long)M + e] = (FP)0.;
      }
    }
    i = ni;
    j0 = nj0;
    last = nlast;
  }
}

// ------------------------------------------------------------------------------------------
// FP64 tensor-core variants (plain se_a, M <= 128): the env-mat contraction is a small dense
// product per atom -- forward  out[4 x M] = E^T[4 x n] . G[n x M],  backward
// dy_dem[n x 4] = G[n x M] . dy^T[M x 4]  (and the same with G' for the em_x gradient) -- so it runs
// on the FP64 tensor cores (DMMA m8n8k4) while the FMA pipe only evaluates the quintic: 5 (forward)
// / 9 (backward) FMAs per (neighbour, channel) instead of 9 / 19 on the SIMT path above, every lane
// owns a live (neighbour, channel) pair (no 100-of-128 channel padding), and the backward needs no
// cross-lane reduction at all.  Fragment ownership of mma.m8n8k4.f64: A[r][k] at lane 4r+k,
// B[k][c] at lane 4c+k, C[r][2k..2k+1] at lane 4r+k.
//   forward : A = E^T (rows = the 4 env-mat components, rows 4..7 zero), k = 4 neighbours,
//             B = G (k = neighbour, column = channel 8t + lane/4), one C tile per 8 channels.
//   backward: A = G / G' (row = neighbour 8s + lane/4, k = channel 4t + lane%4),
//             B = dy^T (k = channel, columns 0..3 = component, 4..7 zero), C = [8 neighbours x 4].
// Coefficients are not cached in registers here (every lane evaluates a different row): they come
// from the shared-memory hot window (lanes of one row read one contiguous segment, equal rows are
// broadcast) or, outside the window, from L1/L2.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

template <int NT, bool DESC>
__global__ void __launch_bounds__(512) k_tab_fwd_mma(const __grid_constant__ TabParams<double> p) {
  using FP = double;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nw = blockDim.x >> 5;
  FP* hot = reinterpret_cast<FP*>(tab_smem);
  Rec<FP>* rec = reinterpret_cast<Rec<FP>*>(hot + p.hot_elems) + warp * 32;
  const int r0 = hot_window_start(p);
  preload_hot(hot, p, r0);
  __syncthreads();
  const int M = p.M;
  const int q = lane >> 2, kk = lane & 3;
  const unsigned qb = (unsigned)M * 16u;  // bytes of one coefficient-pair block
  const unsigned hstride = 3u * qb + (unsigned)p.hot_pad;  // padded: the 4 rows of a step hit 4 different bank groups
  const unsigned off_q = (unsigned)q * 16u;
  const unsigned off_last = (unsigned)((8 * (NT - 1) + q < M) ? 8 * (NT - 1) + q : M - 1) * 16u;

  const long long stride = (long long)gridDim.x * nw;
  long long i = (long long)blockIdx.x * nw + warp;
  int j0 = 0;
  Pre<FP, false> pre;
  load_pre(pre, p, i, 0, lane);
  FP last = i < p.nloc ? p.em_x[i * p.ldx_i + (long long)(p.nnei - 1) * p.ldx_j] : (FP)0.;
  FP acc[NT][2];
  auto seed = [&](long long ii) {
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      const int ch = 8 * t + 2 * kk;
      const bool rd = p.accumulate && ii < p.nloc && q < 4;
      acc[t][0] = (rd && ch < M) ? p.out[(ii * 4 + q) * (long long)M + ch] : (FP)0.;
      acc[t][1] = (rd && ch + 1 < M) ? p.out[(ii * 4 + q) * (long long)M + ch + 1] : (FP)0.;
    }
  };
  seed(i);

  while (i < p.nloc) {
    bool done, any_delta;
    const int nproc = stage_chunk<FP, false>(p, j0, last, pre, rec, nullptr, lane, done, any_delta);
    const bool atom_end = done || j0 + 32 >= p.nnei;
    const long long ni = atom_end ? i + stride : i;
    const int nj0 = atom_end ? 0 : j0 + 32;
    load_pre(pre, p, ni, nj0, lane);
    FP nlast = last;
    if (atom_end && ni < p.nloc) nlast = p.em_x[ni * p.ldx_i + (long long)(p.nnei - 1) * p.ldx_j];

    for (int s = 0; s < nproc; s += 4) {
      const int nb = s + kk;
      const bool live = nb < nproc;
      const Rec<FP>& r = rec[live ? nb : nproc - 1];
      const FP xx = r.xx;
      const FP dl = r.delta;
      const FP eq = r.e[q & 3];
      const FP af = (live && q < 4) ? eq : (FP)0.;
      const unsigned rel = (unsigned)(r.idx - r0);
      const bool inwin = rel < (unsigned)p.H;
#define DPB_FWD_TILES(BASE)                                                                   \
  _Pragma("unroll") for (int t = 0; t < NT; ++t) {                                            \
    const unsigned off = (t == NT - 1) ? off_last : (unsigned)t * 128u + off_q;               \
    const double2 u = *reinterpret_cast<const double2*>((BASE) + off);                        \
    const double2 v = *reinterpret_cast<const double2*>((BASE) + qb + off);                   \
    const double2 w = *reinterpret_cast<const double2*>((BASE) + 2u * qb + off);              \
    FP g = u.x + (u.y + (v.x + (v.y + (w.x + w.y * xx) * xx) * xx) * xx) * xx;                \
    if (any_delta && dl != (FP)0.)                                                            \
      g += (u.y + ((FP)2. * v.x + ((FP)3. * v.y + ((FP)4. * w.x + (FP)5. * w.y * xx) * xx) * xx) * xx) * dl; \
    dmma884(acc[t][0], acc[t][1], af, g);                                                     \
  }
      if (__all_sync(kFull, inwin)) {
        const char* b = reinterpret_cast<const char*>(hot) + rel * hstride;
        DPB_FWD_TILES(b)
      } else if (!__any_sync(kFull, inwin)) {
        const char* __restrict__ b = reinterpret_cast<const char*>(p.T) + (long long)r.idx * (3u * qb);
        DPB_FWD_TILES(b)
      } else {
        const char* b = inwin ? reinterpret_cast<const char*>(hot) + rel * hstride
                              : reinterpret_cast<const char*>(p.T) + (long long)r.idx * (3u * qb);
        DPB_FWD_TILES(b)
      }
#undef DPB_FWD_TILES
    }
    if (atom_end) {
      if (q < 4) {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const int ch = 8 * t + 2 * kk;
          if (ch < M) p.out[(i * 4 + q) * (long long)M + ch] = acc[t][0];
          if (ch + 1 < M) p.out[(i * 4 + q) * (long long)M + ch + 1] = acc[t][1];
        }
      }
      if (DESC) {
        // the descriptor epilogue wants channel-per-lane ownership: read the row back (L1/L2 hit)
        __syncwarp();
        FP a4[4][4];
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            const int k = lane + 32 * c;
            a4[m][c] = k < M ? __ldcg(p.out + (i * 4 + m) * (long long)M + k) : (FP)0.;
          }
        desc_epilogue<FP, 4>(p, a4, i, lane, reinterpret_cast<FP*>(rec));
      }
      seed(ni);
    }
    i = ni;
    j0 = nj0;
    last = nlast;
  }
}

// leading dimension (doubles) of the per-warp dy tile of the tensor-core backward: >= 4*KT and = 4 mod 16
__host__ __device__ constexpr int grad_mma_tile_ld(int kt) { return (4 * kt + 11) / 16 * 16 + 4; }

// BSM: the B fragments (dy of the current atom) live in a per-warp shared-memory tile instead of 2*KT
// registers, which lets MAXT / 32 warps (instead of 12) share an SM: the kernel is latency-bound on the
// coefficient fetches of rows outside the hot window, so warps in flight are what it needs.
// GATE: se_atten with the pair-indexed gate (two_embed = tt_full[pair] * sw, never materialised): G and G' are
// scaled by 1 + t, and dE/d(sw) = sum_m e[m] (g tt . dy^T)[m] comes from a third product with A = g * tt.
// KR (BSM only): the B fragments of the first KR channel steps are copied from the shared-memory tile into registers
// once per atom (the kernel has register headroom at 16 warps, and every B load is a wavefront of the L1/shared pipe
// that bounds it).
template <int KT, bool BSM, int MAXT, bool CM = false, bool GATE = false, int KR = 0>
__global__ void __launch_bounds__(MAXT) k_tab_grad_mma(const __grid_constant__ TabParams<double> p) {
  using FP = double;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nw = blockDim.x >> 5;
  FP* hot = reinterpret_cast<FP*>(tab_smem);
  Rec<FP>* rec = reinterpret_cast<Rec<FP>*>(hot + p.hot_elems) + warp * 32;
  const int M = p.M;
  // per-warp dy tile [4][MP]: rows = components (channels >= M zero).  The lanes that own the unused B columns 4..7
  // read the row of column n - 4 (a broadcast of the same words: no predicate, no bank conflict; C columns 4..7
  // are never read).  MP = 4 mod 16 doubles, so the four rows start 8 banks apart and one B load is one wavefront.
  constexpr int MP = grad_mma_tile_ld(KT);
  FP* dyt = reinterpret_cast<FP*>(reinterpret_cast<Rec<FP>*>(hot + p.hot_elems) + nw * 32) + (BSM ? warp * 4 * MP : 0);
  const int r0 = hot_window_start(p);
  preload_hot(hot, p, r0);
  __syncthreads();
  const bool fuse_x = p.dy_dem_x == nullptr;
  const int q = lane >> 2, kk = lane & 3;
  const unsigned qb = (unsigned)M * 16u;
  const unsigned rowb = (CM ? 2u : 3u) * qb;  // bytes per table row
  const unsigned hrowb = rowb + (unsigned)p.hot_pad;  // ... in the shared-memory window (padded, see launch_grad)
  const unsigned off_k = (unsigned)kk * 16u;
  const unsigned off_last = (unsigned)((4 * (KT - 1) + kk < M) ? 4 * (KT - 1) + kk : M - 1) * 16u;

  const long long stride = (long long)gridDim.x * nw;
  long long i = (long long)blockIdx.x * nw + warp;
  int j0 = 0;
  Pre<FP, false> pre;
  load_pre(pre, p, i, 0, lane);
  FP last = i < p.nloc ? p.em_x[i * p.ldx_i + (long long)(p.nnei - 1) * p.ldx_j] : (FP)0.;
  FP bf[BSM ? (KR > 0 ? KR : 1) : KT];  // B fragments: dy[m = q][channel 4t + kk] (q < 4), constant per atom
  const FP* bsrc = dyt + (q & 3) * MP + kk;
  if (BSM) {
    for (int e = lane; e < 4 * MP; e += 32) dyt[e] = (FP)0.;
    __syncwarp();
  }

  while (i < p.nloc) {
    if (j0 == 0) {
      const FP* __restrict__ dyi = p.dy + i * 4 * (long long)M;
      if (BSM) {
        __syncwarp();
        for (int e = lane; e < 4 * M; e += 32) {
          const int m = e / M;
          dyt[m * MP + (e - m * M)] = dyi[e];
        }
        __syncwarp();
        if (KR > 0) {
#pragma unroll
          for (int t = 0; t < (KR < KT ? KR : KT); ++t) bf[t] = bsrc[4 * t];
        }
      } else {
#pragma unroll
        for (int t = 0; t < KT; ++t) {
          const int ch = 4 * t + kk;
          bf[BSM ? 0 : t] = (q < 4 && ch < M) ? dyi[(long long)q * M + ch] : (FP)0.;
        }
      }
    }
    bool done, any_delta;
    const int nproc = stage_chunk<FP, false>(p, j0, last, pre, rec, nullptr, lane, done, any_delta);
    const bool atom_end = done || j0 + 32 >= p.nnei;
    const long long ni = atom_end ? i + stride : i;
    const int nj0 = atom_end ? 0 : j0 + 32;
    load_pre(pre, p, ni, nj0, lane);
    FP nlast = last;
    if (atom_end && ni < p.nloc) nlast = p.em_x[ni * p.ldx_i + (long long)(p.nnei - 1) * p.ldx_j];

    FP* __restrict__ gem = p.dy_dem + i * p.ldem_i;
    if (j0 == 0 && i + stride < p.nloc && lane * 16 < 4 * M) {
      // the next atom's dy tile (4*M doubles, staged into shared memory when that atom starts) into L2 now: the
      // staging stores otherwise wait a full HBM round trip per atom (ncu: 8 % of the kernel's stall samples)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p.dy + (i + stride) * 4 * (long long)M + lane * 16));
    }
    for (int s = 0; s < nproc; s += 8) {
      const int nb = s + q;
      const bool live = nb < nproc;
      const Rec<FP>& r = rec[live ? nb : nproc - 1];
      const FP xx = r.xx;
      const float xf = (float)xx;
      const float xs = xf * p.a5_inv;  // a5 is stored pre-multiplied by a power of two (fp16 range)
      const FP dl = r.delta;
      const unsigned rel = (unsigned)(r.idx - r0);
      const bool inwin = rel < (unsigned)p.H;
      FP c1a = 0., c1b = 0., c2a = 0., c2b = 0., c3a = 0., c3b = 0.;
      FP gsw = 0.;
      const FP* __restrict__ ttrow = nullptr;
      if (GATE) {
        const long long gj = i * p.nnei + j0 + (live ? nb : nproc - 1);
        gsw = p.gate_sw[gj];
        ttrow = p.gate_tt + (long long)p.gate_pair[gj] * M;
      }
#define DPB_GRAD_STEPS(BASE, COMP)                                                            \
  _Pragma("unroll") for (int t = 0; t < KT; ++t) {                                            \
    const unsigned off = (t == KT - 1) ? off_last : (unsigned)t * 64u + off_k;                \
    FP g, gd;                                                                                 \
    if (COMP) {                                                                               \
      /* a3..a5 are fp32 / fp16: the top of both Horner chains runs on the FP32 pipe */       \
      const double2 u = *reinterpret_cast<const double2*>((BASE) + off);                      \
      const double2 v = *reinterpret_cast<const double2*>((BASE) + qb + off);                 \
      double a2;                                                                              \
      float a3, a4, a5;                                                                       \
      unpack_cm(v, a2, a3, a4, a5);                                                           \
      const float f4 = fmaf(a5, xs, a4);                                                      \
      const float f3 = fmaf(f4, xf, a3);                                                      \
      const float e4 = fmaf(a5, xs, f4);                                                      \
      const float e3 = fmaf(e4, xf, f3);                                                      \
      const FP b2 = a2 + (FP)f3 * xx;                                                         \
      const FP b1 = u.y + b2 * xx;                                                            \
      g = u.x + b1 * xx;                                                                      \
      const FP d2 = b2 + (FP)e3 * xx;                                                         \
      gd = b1 + d2 * xx;                                                                      \
    } else {                                                                                  \
      const double2 u = *reinterpret_cast<const double2*>((BASE) + off);                      \
      const double2 v = *reinterpret_cast<const double2*>((BASE) + qb + off);                 \
      const double2 w = *reinterpret_cast<const double2*>((BASE) + 2u * qb + off);            \
      const FP b4 = w.x + w.y * xx;                                                           \
      const FP b3 = v.y + b4 * xx;                                                            \
      const FP b2 = v.x + b3 * xx;                                                            \
      const FP b1 = u.y + b2 * xx;                                                            \
      g = u.x + b1 * xx;                                                                      \
      const FP d4 = b4 + w.y * xx;                                                            \
      const FP d3 = b3 + d4 * xx;                                                             \
      const FP d2 = b2 + d3 * xx;                                                             \
      gd = b1 + d2 * xx;                                                                      \
    }                                                                                         \
    g += gd * dl; /* dl == 0 inside [lower, max): one FMA instead of a select pair */         \
    const FP bt = BSM ? (t < KR ? bf[t < KR ? t : 0] : bsrc[4 * t]) : bf[BSM ? 0 : t];        \
    if (GATE) {                                                                               \
      const FP tt = __ldg(ttrow + (off >> 4)); /* off = 16 * (clamped) channel */              \
      dmma884(c3a, c3b, g * tt, bt);                                                          \
      const FP tg = tt * gsw;                                                                 \
      g = fma(g, tg, g);                                                                      \
      gd = fma(gd, tg, gd);                                                                   \
    }                                                                                         \
    dmma884(c1a, c1b, g, bt);                                                                 \
    dmma884(c2a, c2b, gd, bt);                                                                \
  }
      if (CM && __any_sync(kFull, r.idx >= p.first)) {
        // a stride-1 (extrapolation) row in this step: those are not compressed, take the full table for all lanes
        const char* __restrict__ b = reinterpret_cast<const char*>(p.T3) + (long long)r.idx * (3u * qb);
        DPB_GRAD_STEPS(b, false)
      } else if (__all_sync(kFull, inwin)) {
        const char* b = reinterpret_cast<const char*>(hot) + rel * hrowb;
        DPB_GRAD_STEPS(b, CM)
      } else if (!__any_sync(kFull, inwin)) {
        const char* __restrict__ b = reinterpret_cast<const char*>(p.T) + (long long)r.idx * rowb;
        DPB_GRAD_STEPS(b, CM)
      } else {
        const char* b = inwin ? reinterpret_cast<const char*>(hot) + rel * hrowb
                              : reinterpret_cast<const char*>(p.T) + (long long)r.idx * rowb;
        DPB_GRAD_STEPS(b, CM)
      }
#undef DPB_GRAD_STEPS
      // C row = neighbour q, columns 2kk, 2kk+1 (components; valid for kk < 2)
      const FP ea = r.e[(2 * kk) & 3], eb = r.e[(2 * kk + 1) & 3];  // pre-multiplied by the fold multiplicity
      FP part = kk < 2 ? ea * c2a + eb * c2b : (FP)0.;
      part += __shfl_xor_sync(kFull, part, 1);
      if (GATE) {
        FP pq = kk < 2 ? ea * c3a + eb * c3b : (FP)0.;
        pq += __shfl_xor_sync(kFull, pq, 1);
        if (live && kk == 0) p.gate_q[i * p.nnei + j0 + nb] = pq;
      }
      if (live && kk < 2) {
        const FP mult = (FP)r.mult;
        const int j = j0 + nb;
        FP v0 = c1a * mult;
        const FP v1 = c1b * mult;
        if (kk == 0) {
          if (fuse_x)
            v0 += part;
          else
            p.dy_dem_x[i * p.ldx_i + (long long)j * p.ldx_j] = part;
        }
        gem[(long long)j * 4 + 2 * kk] = v0;
        gem[(long long)j * 4 + 2 * kk + 1] = v1;
      }
    }
    if (atom_end) {
      const int jend = j0 + nproc;
      for (int j = jend + lane; j < p.nnei; j += 32) {
        if (!fuse_x) p.dy_dem_x[i * p.ldx_i + (long long)j * p.ldx_j] = (FP)0.;
#pragma unroll
        for (int m = 0; m < 4; ++m) gem[(long long)j * 4 + m] = (FP)0.;
        if (GATE) p.gate_q[i * p.nnei + j] = (FP)0.;
      }
    }
    i = ni;
    j0 = nj0;
    last = nlast;
  }
}

// fp32 flavour of the tensor-core backward: fp32 operands and results, cubic float4 coefficients (the fp32 form of
// DPB200_TAB_COMPRESSED_COEF: ONE 16-byte request per (neighbour, channel)), the quintic's value and derivative on the
// FP32 pipe (5 FMAs), the contraction on the FP64 tensor cores with fp64 accumulation (the only mma.sync shape
// whose fragment ownership -- one A element per lane -- fits "each lane evaluates one (neighbour, channel)";
// the products are exact in fp64, so this is also more accurate than the fp32 SIMT kernel).
template <int KT>
__global__ void __launch_bounds__(512) k_tab_grad_mma_f32(const __grid_constant__ TabParams<float> p) {
  using FP = float;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nw = blockDim.x >> 5;
  FP* hot = reinterpret_cast<FP*>(tab_smem);
  Rec<FP>* rec = reinterpret_cast<Rec<FP>*>(hot + p.hot_elems) + warp * 32;
  const int M = p.M;
  constexpr int MP = grad_mma_tile_ld(KT);  // [4][MP], B columns 4..7 alias 0..3 (see k_tab_grad_mma)
  double* dyt = reinterpret_cast<double*>(reinterpret_cast<Rec<FP>*>(hot + p.hot_elems) + nw * 32) + warp * 4 * MP;
  const int r0 = hot_window_start(p);
  preload_hot(hot, p, r0);
  __syncthreads();
  const bool fuse_x = p.dy_dem_x == nullptr;
  const int q = lane >> 2, kk = lane & 3;
  const unsigned rowb = (unsigned)M * 16u;  // one float4 per channel
  const unsigned off_k = (unsigned)kk * 16u;
  const unsigned off_last = (unsigned)((4 * (KT - 1) + kk < M) ? 4 * (KT - 1) + kk : M - 1) * 16u;

  const long long stride = (long long)gridDim.x * nw;
  long long i = (long long)blockIdx.x * nw + warp;
  int j0 = 0;
  Pre<FP, false> pre;
  load_pre(pre, p, i, 0, lane);
  FP last = i < p.nloc ? p.em_x[i * p.ldx_i + (long long)(p.nnei - 1) * p.ldx_j] : (FP)0.;
  const double* bsrc = dyt + (q & 3) * MP + kk;
  for (int e = lane; e < 4 * MP; e += 32) dyt[e] = 0.;
  __syncwarp();

  while (i < p.nloc) {
    if (j0 == 0) {
      const FP* __restrict__ dyi = p.dy + i * 4 * (long long)M;
      __syncwarp();
      for (int e = lane; e < 4 * M; e += 32) {
        const int m = e / M;
        dyt[m * MP + (e - m * M)] = (double)dyi[e];
      }
      __syncwarp();
    }
    bool done, any_delta;
    const int nproc = stage_chunk<FP, false>(p, j0, last, pre, rec, nullptr, lane, done, any_delta);
    const bool atom_end = done || j0 + 32 >= p.nnei;
    const long long ni = atom_end ? i + stride : i;
    const int nj0 = atom_end ? 0 : j0 + 32;
    load_pre(pre, p, ni, nj0, lane);
    FP nlast = last;
    if (atom_end && ni < p.nloc) nlast = p.em_x[ni * p.ldx_i + (long long)(p.nnei - 1) * p.ldx_j];

    FP* __restrict__ gem = p.dy_dem + i * p.ldem_i;
    for (int s = 0; s < nproc; s += 8) {
      const int nb = s + q;
      const bool live = nb < nproc;
      const Rec<FP>& r = rec[live ? nb : nproc - 1];
      const FP xx = r.xx;
      const FP dl = r.delta;
      const unsigned rel = (unsigned)(r.idx - r0);
      const bool inwin = rel < (unsigned)p.H;
      double c1a = 0., c1b = 0., c2a = 0., c2b = 0.;
#define DPB_GRAD32_STEPS(BASE)                                                                \
  _Pragma("unroll") for (int t = 0; t < KT; ++t) {                                            \
    const unsigned off = (t == KT - 1) ? off_last : (unsigned)t * 64u + off_k;                \
    const float4 u = *reinterpret_cast<const float4*>((BASE) + off);                          \
    const FP b2 = u.z + u.w * xx;                                                             \
    const FP b1 = u.y + b2 * xx;                                                              \
    FP g = u.x + b1 * xx;                                                                     \
    const FP d2 = b2 + u.w * xx;                                                              \
    const FP gd = b1 + d2 * xx;                                                               \
    g += gd * dl;                                                                             \
    const double bt = bsrc[4 * t];                                                            \
    dmma884(c1a, c1b, (double)g, bt);                                                         \
    dmma884(c2a, c2b, (double)gd, bt);                                                        \
  }
      if (__all_sync(kFull, inwin)) {
        const char* b = reinterpret_cast<const char*>(hot) + rel * rowb;
        DPB_GRAD32_STEPS(b)
      } else if (!__any_sync(kFull, inwin)) {
        const char* __restrict__ b = reinterpret_cast<const char*>(p.T) + (long long)r.idx * rowb;
        DPB_GRAD32_STEPS(b)
      } else {
        const char* b = inwin ? reinterpret_cast<const char*>(hot) + rel * rowb
                              : reinterpret_cast<const char*>(p.T) + (long long)r.idx * rowb;
        DPB_GRAD32_STEPS(b)
      }
#undef DPB_GRAD32_STEPS
      const double ea = (double)r.e[(2 * kk) & 3], eb = (double)r.e[(2 * kk + 1) & 3];
      double part = kk < 2 ? ea * c2a + eb * c2b : 0.;
      part += __shfl_xor_sync(kFull, part, 1);
      if (live && kk < 2) {
        const double mult = (double)r.mult;
        const int j = j0 + nb;
        double v0 = c1a * mult;
        const double v1 = c1b * mult;
        if (kk == 0) {
          if (fuse_x)
            v0 += part;
          else
            p.dy_dem_x[i * p.ldx_i + (long long)j * p.ldx_j] = (FP)part;
        }
        gem[(long long)j * 4 + 2 * kk] = (FP)v0;
        gem[(long long)j * 4 + 2 * kk + 1] = (FP)v1;
      }
    }
    if (atom_end) {
      const int jend = j0 + nproc;
      for (int j = jend + lane; j < p.nnei; j += 32) {
        if (!fuse_x) p.dy_dem_x[i * p.ldx_i + (long long)j * p.ldx_j] = (FP)0.;
#pragma unroll
        for (int m = 0; m < 4; ++m) gem[(long long)j * 4 + m] = (FP)0.;
      }
    }
    i = ni;
    j0 = nj0;
    last = nlast;
  }
}

inline int grad_variant() {
  static const int v = [] {
    const char* e = getenv("DPB200_GRAD_VARIANT");
    return e ? atoi(e) : 1;
  }();
  return v;
}

inline int hot_rows_cap() {
  static const int v = [] {
    const char* e = getenv("DPB200_TAB_HOT_ROWS");
    return e ? atoi(e) : (1 << 30);
  }();
  return v;
}

inline int grad_hot_pad() {
  static const int v = [] {
    const char* e = getenv("DPB200_TAB_GRAD_PAD");
    return e ? atoi(e) : 64;
  }();
  return v;
}

inline int grad_breg() {
  static const int v = [] {
    const char* e = getenv("DPB200_GRAD_BREG");  // 0: all B fragments from shared memory (comparison runs)
    return e ? atoi(e) : 1;
  }();
  return v;
}

inline bool use_mma_fwd() {
  static const bool on = [] {
    const char* e = getenv("DPB200_TAB_MMA_FWD");
    return e && e[0] == '1';
  }();
  return on;
}

inline bool fwd_cm_enabled() {
  static const bool on = [] {
    const char* e = getenv("DPB200_TAB_FWD_CM");
    return !(e && e[0] == '0');
  }();
  return on;
}

inline bool use_mma_path() {
  static const bool on = [] {
    const char* e = getenv("DPB200_TAB_MMA");
    return !(e && e[0] == '0');
  }();
  return on;
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
constexpr size_t kSmemBudget = 225 * 1024;

template <typename FP>
int fill_info(TabParams<FP>& p, const FP* info, int M) {
  DPB_REQUIRE(info != nullptr, "tabulate: table_info is null (it must be a HOST pointer)");
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
  p.nrow = p.tail_idx + 1;
  p.M = M;
  return DPB200_OK;
}

inline bool aligned16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; }

template <typename FP>
int common_args(TabParams<FP>& p, const FP* table, const FP* em_x, long long ldx_i, int ldx_j, const FP* em,
                long long ldem_i, const FP* two, int nloc, int nnei, int is_sorted) {
  DPB_REQUIRE(table != nullptr && em_x != nullptr && em != nullptr, "tabulate: table / em_x / em are null");
  DPB_REQUIRE(aligned16(table), "tabulate: table must be 16-byte aligned");
  p.table = table;
  p.em_x = em_x;
  p.ldx_i = ldx_i;
  p.ldx_j = ldx_j;
  p.em = em;
  p.ldem_i = ldem_i;
  p.two = two;
  p.nloc = nloc;
  p.nnei = nnei;
  p.is_sorted = is_sorted ? 1 : 0;
  p.vec_ok = aligned16(em) && ((ldem_i * sizeof(FP)) % 16 == 0) ? 1 : 0;
  return DPB200_OK;
}

template <typename FP>
int hot_elems_aligned(int H, int Mc) {
  const long long per16 = 16 / sizeof(FP);
  const long long n = (long long)H * Mc * 6;
  return (int)((n + per16 - 1) / per16 * per16);
}

// hot rows (full width) that fit beside the per-warp records
template <typename FP>
int hot_rows(int nrow, int M, size_t other_bytes) {
  const size_t row_bytes = (size_t)M * 6 * sizeof(FP);
  if (other_bytes + row_bytes > kSmemBudget) return 0;
  long long h = (long long)((kSmemBudget - other_bytes) / row_bytes);
  if (h > hot_rows_cap()) h = hot_rows_cap();
  return (int)(h < nrow ? h : nrow);
}

template <typename FP>
int prepare_table(TabParams<FP>& p, FP** scratch, cudaStream_t st) {
  const long long n = (long long)p.nrow * 6 * p.M;
  keep_async_pool();
  DPB_CUDA(cudaMallocAsync((void**)scratch, (size_t)n * sizeof(FP), st));
  int grid = ceil_div(n, 256);
  const int cap = sm_count() * 8;
  if (grid > cap) grid = cap;
  k_table_relayout<FP><<<grid, 256, 0, st>>>(*scratch, p.table, p.nrow, p.M);
  p.T = *scratch;
  return DPB200_OK;
}

inline int prepare_table_cm(TabParams<double>& p, double** scratch, cudaStream_t st) {
  const long long n = (long long)p.nrow * 4 * p.M;  // two 16-byte blocks per (row, channel)
  keep_async_pool();
  DPB_CUDA(cudaMallocAsync((void**)scratch, (size_t)n * sizeof(double), st));
  int grid = ceil_div((long long)p.nrow * p.M, 256);
  const int cap = sm_count() * 8;
  if (grid > cap) grid = cap;
  k_table_relayout_cm<<<grid, 256, 0, st>>>(reinterpret_cast<double2*>(*scratch), p.table, p.nrow, p.M, p.a5_mul);
  p.T = *scratch;
  return DPB200_OK;
}
inline int prepare_table_cm(TabParams<float>& p, float** scratch, cudaStream_t st, int) {
  const long long n = (long long)p.nrow * 4 * p.M;  // one float4 per (row, channel)
  keep_async_pool();
  DPB_CUDA(cudaMallocAsync((void**)scratch, (size_t)n * sizeof(float), st));
  int grid = ceil_div((long long)p.nrow * p.M, 256);
  const int cap = sm_count() * 8;
  if (grid > cap) grid = cap;
  k_table_relayout_c32<<<grid, 256, 0, st>>>(reinterpret_cast<float4*>(*scratch), p.table, p.nrow, p.M);
  p.T = *scratch;
  return DPB200_OK;
}
inline int prepare_table_cm(TabParams<double>& p, double** scratch, cudaStream_t st, int) {
  return prepare_table_cm(p, scratch, st);
}
// full pair table next to the compressed one (for the stride-1 rows)
inline int prepare_table_full(TabParams<double>& p, double** scratch3, cudaStream_t st) {
  const long long n = (long long)p.nrow * 6 * p.M;
  DPB_CUDA(cudaMallocAsync((void**)scratch3, (size_t)n * sizeof(double), st));
  int grid = ceil_div(n, 256);
  const int cap = sm_count() * 8;
  if (grid > cap) grid = cap;
  k_table_relayout<double><<<grid, 256, 0, st>>>(*scratch3, p.table, p.nrow, p.M);
  p.T3 = *scratch3;
  return DPB200_OK;
}
inline int prepare_table_full(TabParams<float>&, float**, cudaStream_t) { return DPB200_ERR_INVALID; }

// bits 8..15 of `flags`: signed power-of-two exponent k, a5 is stored as half(a5 * 2^k)
template <typename FP>
void set_a5_scale(TabParams<FP>& p, int flags) {
  const int k = (int)(signed char)((flags >> 8) & 0xff);
  p.a5_mul = std::ldexp(1.0f, k);
  p.a5_inv = std::ldexp(1.0f, -k);
}

// hot rows / shared-memory sizing for `blocks` coefficient-pair blocks (2 * sizeof(FP) bytes) per (row, channel)
template <typename FP>
void size_hot_window(TabParams<FP>& p, int M, size_t other_bytes, int blocks) {
  const size_t row_bytes = (size_t)M * 2 * sizeof(FP) * blocks + p.hot_pad;  // `blocks` pair blocks per (row, channel)
  long long h = other_bytes + row_bytes > kSmemBudget ? 0 : (long long)((kSmemBudget - other_bytes) / row_bytes);
  if (h > hot_rows_cap()) h = hot_rows_cap();
  p.H = (int)(h < p.nrow ? h : p.nrow);
  p.hot_elems = (int)(((size_t)p.H * row_bytes + 15) / 16 * 16 / sizeof(FP));
}

struct DescArgs {
  void* desc;
  long long desc_ld;
  const int* desc_row;
  int* row_exp;
  double scale;
  int mode, axis, nslice;
  long long slice_stride = 0;  // 0: M * axis
  int min_exp = -100000;
};

template <typename FP>
struct GateArgs {
  const FP* tt;
  const int* pair;
  const FP* sw;
  FP* q;
};

template <typename FP, bool GG>
int launch_fwd(FP* out, const FP* table, const FP* info, const FP* em_x, long long ldx_i, int ldx_j,
               const FP* em, long long ldem_i, const FP* two, const FP* dz_x, const FP* dz_em,
               const FP* dz_two, int nloc, int nnei, int M, int is_sorted, int accumulate,
               cudaStream_t st, const DescArgs* da = nullptr, int flags = 0, const GateArgs<FP>* ga = nullptr) {
  DPB_REQUIRE(nloc >= 0 && nnei >= 0 && M >= 0, "tabulate: negative size");
  if (nloc == 0 || M == 0) return DPB200_OK;
  DPB_REQUIRE(out != nullptr, "tabulate: out is null");
  DPB_REQUIRE(!ga || (!GG && two == nullptr && ga->tt && ga->pair && ga->sw), "tabulate gate: bad arguments");
  // compressed coefficients in the SIMT forward: a3..a5 stay fp32 in the row cache and the top two Horner steps
  // run on the FP32 pipe (widening them to fp64 at fetch time instead was measured 8 % SLOWER than the full table)
  // (not with the pair-indexed gate: the gated loop with fp32 coefficient registers measured 17.1 -> 21.4 ms at
  //  526 848 atoms -- the extra live values push the kernel over its register budget)
  const bool cm = (flags & DPB200_TAB_COMPRESSED_COEF) && !GG && two == nullptr && !ga && !use_mma_fwd() && fwd_cm_enabled();
  if (GG && da) {
    set_error("tabulate+descriptor: plain se_a forward only");
    return DPB200_ERR_INVALID;
  }
  if constexpr (!GG) if (da) {
    const bool plain = two == nullptr && nnei > 0;
    DPB_REQUIRE(plain, "tabulate+descriptor: se_a or pair-gated se_atten forward (no materialised two_embed), nnei > 0");
    DPB_REQUIRE(da->desc != nullptr && (da->mode == 1 || da->mode == 2 || (da->mode == 3 && sizeof(FP) == 4)),
                "tabulate+descriptor: desc is null / bad mode (3 = int8 slices of an fp32 descriptor)");
    DPB_REQUIRE(M <= 128 && da->axis >= 1 && da->axis <= 32 && da->axis <= M,
                "tabulate+descriptor: needs M <= 128 and axis <= min(32, M)");
    DPB_REQUIRE(aligned16(da->desc), "tabulate+descriptor: desc must be 16-byte aligned");
    if (da->mode == 2 && sizeof(FP) == 8) {
      DPB_REQUIRE(da->axis == 16 && da->nslice >= 2 && da->nslice <= 8 && da->row_exp != nullptr &&
                      da->desc_ld % 16 == 0 && da->desc_ld >= (long long)da->nslice * M * 16,
                  "tabulate+descriptor: int8 split needs axis == 16, 2 <= nslice <= 8, row_exp, 16-byte rows");
    } else if (da->mode == 3) {
      DPB_REQUIRE(da->axis == 16 && da->nslice == 4 && da->row_exp != nullptr && da->desc_ld % 16 == 0 &&
                      da->desc_ld >= 4LL * M * 16,
                  "tabulate+descriptor: fp32 int8 split needs axis == 16, nslice == 4, row_exp, 16-byte rows");
    } else if (da->mode == 2) {
      DPB_REQUIRE(da->axis % 4 == 0 && da->desc_ld % 4 == 0 && da->desc_ld >= 2LL * M * da->axis,
                  "tabulate+descriptor: TF32 split needs axis % 4 == 0 and 16-byte rows of >= 2*M*axis floats");
    } else {
      DPB_REQUIRE(da->desc_ld >= (long long)M * da->axis, "tabulate+descriptor: desc_ld < M*axis");
    }
  }
  if (nnei == 0) {  // an empty neighbour axis is a valid empty reduction (tabulate.cc:176-181)
    if (!accumulate) DPB_CUDA(cudaMemsetAsync(out, 0, sizeof(FP) * (size_t)nloc * 4 * M, st));
    return DPB200_OK;
  }
  TabParams<FP> p = {};
  int rc = fill_info(p, info, M);
  if (rc) return rc;
  rc = common_args(p, table, em_x, ldx_i, ldx_j, em, ldem_i, two, nloc, nnei, is_sorted);
  if (rc) return rc;
  p.out = out;
  p.accumulate = accumulate;
  p.dz_x = dz_x;
  p.dz_em = dz_em;
  p.dz_two = dz_two;
  if (ga) p.gate_tt = ga->tt, p.gate_pair = ga->pair, p.gate_sw = ga->sw;
  if (da) {
    p.desc = da->desc;
    p.desc_ld = da->desc_ld;
    p.desc_row = da->desc_row;
    p.row_exp = da->row_exp;
    p.desc_scale = (FP)da->scale;
    p.desc_mode = da->mode;
    p.axis = da->axis;
    p.nslice = da->nslice;
    p.desc_slice_stride = da->slice_stride > 0 ? da->slice_stride : (long long)M * da->axis;
    p.desc_min_exp = da->min_exp;
    DPB_REQUIRE(p.desc_slice_stride >= (long long)M * da->axis && p.desc_slice_stride % 16 == 0 &&
                    (!((da->mode == 2 && sizeof(FP) == 8) || da->mode == 3) ||
                     da->desc_ld >= da->nslice * p.desc_slice_stride),
                "tabulate+descriptor: slice stride must be a multiple of 16, >= M*axis, and nslice strides fit a row");
  }
  if (GG) {
    DPB_REQUIRE(dz_x != nullptr && dz_em != nullptr, "tabulate grad_grad: dz_dy_dem_x / dz_dy_dem are null");
    DPB_REQUIRE(aligned16(dz_em), "tabulate grad_grad: dz_dy_dem must be 16-byte aligned");
    DPB_REQUIRE(two == nullptr || dz_two != nullptr, "tabulate grad_grad: dz_dy_dtwo is null");
  }
  const int nc = M <= 32 ? 1 : (M <= 64 ? 2 : 4);
  const int nw = (da ? fwd_threads<FP, true>() : fwd_threads<FP, false>()) / 32;
  // se_atten: per-warp cp.async ring for two_embed rows (needs 16-byte aligned rows)
  const bool ring = !GG && two != nullptr && ((size_t)M * sizeof(FP)) % 16 == 0 && aligned16(two) && M <= 128;
  p.two_ring = ring ? 1 : 0;
  const size_t rec_bytes = (size_t)nw * 32 * (sizeof(Rec<FP>) + (GG ? sizeof(RecGG<FP>) : 0)) +
                           (ring ? (size_t)nw * kRing * M * sizeof(FP) : 0);
  p.Mc = M;
  p.nblk = cm ? 2 : 3;
  set_a5_scale(p, flags);
  size_hot_window(p, M, rec_bytes, p.nblk);
  const size_t smem = (size_t)p.hot_elems * sizeof(FP) + rec_bytes;
  FP* scratch = nullptr;
  FP* scratch3 = nullptr;
  rc = cm ? prepare_table_cm(p, &scratch, st, 0) : prepare_table(p, &scratch, st);
  if (rc) return rc;
  // (the forward needs the VALUE only: the compressed table is accurate enough on the stride-1 rows too -- the
  //  host-side gate checks that -- so no second table here)
  const bool tw = two != nullptr || ga != nullptr;
  const int nblk = (M + 32 * nc - 1) / (32 * nc);
  long long want = ((long long)nloc + nw - 1) / nw;
  const long long cap = sm_count() / nblk > 0 ? sm_count() / nblk : 1;
  dim3 grid((unsigned)(want < cap ? want : cap), (unsigned)nblk);
  cudaError_t e1 = cudaSuccess;
  bool launched = false;
  if constexpr (std::is_same<FP, double>::value && !GG) {
    const int nt = (M + 7) / 8;
    // Measured on B200 (gpurun_out/tab_variants.log): the forward gathers 4 different rows per quarter-warp and
    // ends up L1/shared-bound (9.3 ms vs 5.7 ms for the SIMT kernel at 332k atoms), so it is opt-in; the backward
    // (2 rows per quarter-warp, no cross-lane reduction) wins 11.5 -> 8.3 ms and is the default.
    if (!tw && use_mma_fwd() && (nt == 4 || nt == 8 || nt == 10 || nt == 13 || nt == 16)) {
      dim3 g1((unsigned)(want < sm_count() ? want : sm_count()));
      {  // padded hot rows (bank-conflict-free 4-row gathers)
        p.hot_pad = 32;
        const size_t row_bytes = (size_t)M * 6 * sizeof(FP) + 32;
        long long h = rec_bytes + row_bytes > kSmemBudget ? 0 : (long long)((kSmemBudget - rec_bytes) / row_bytes);
        if (h > hot_rows_cap()) h = hot_rows_cap();
        p.H = (int)(h < p.nrow ? h : p.nrow);
        p.hot_elems = (int)(((size_t)p.H * row_bytes + 15) / 16 * 16 / sizeof(FP));
      }
      const size_t smem = (size_t)p.hot_elems * sizeof(FP) + rec_bytes;
#define DPB_LAUNCH_FWD_MMA(NT)                                                                  \
  do {                                                                                          \
    if (da) {                                                                                   \
      auto kern = k_tab_fwd_mma<NT, true>;                                                      \
      e1 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
      if (e1 == cudaSuccess) kern<<<g1, nw * 32, smem, st>>>(p);                                \
    } else {                                                                                    \
      auto kern = k_tab_fwd_mma<NT, false>;                                                     \
      e1 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
      if (e1 == cudaSuccess) kern<<<g1, nw * 32, smem, st>>>(p);                                \
    }                                                                                           \
  } while (0)
      switch (nt) {
        case 4: DPB_LAUNCH_FWD_MMA(4); break;
        case 8: DPB_LAUNCH_FWD_MMA(8); break;
        case 10: DPB_LAUNCH_FWD_MMA(10); break;
        case 13: DPB_LAUNCH_FWD_MMA(13); break;
        default: DPB_LAUNCH_FWD_MMA(16); break;
      }
#undef DPB_LAUNCH_FWD_MMA
      launched = true;
    }
  }
#define DPB_LAUNCH_FWD(NC)                                                                      \
  do {                                                                                          \
    if (da && ga) {                                                                             \
      auto kern = k_tab_fwd<FP, NC, true, false, true>;                                         \
      e1 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
      if (e1 == cudaSuccess) kern<<<grid, nw * 32, smem, st>>>(p);                              \
    } else if (da && cm) {                                                                      \
      auto kern = k_tab_fwd<FP, NC, false, false, true, true>;                                  \
      e1 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
      if (e1 == cudaSuccess) kern<<<grid, nw * 32, smem, st>>>(p);                              \
    } else if (cm) {                                                                            \
      auto kern = k_tab_fwd<FP, NC, false, false, false, true>;                                 \
      e1 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
      if (e1 == cudaSuccess) kern<<<grid, nw * 32, smem, st>>>(p);                              \
    } else if (da) {                                                                            \
      auto kern = k_tab_fwd<FP, NC, false, false, true>;                                        \
      e1 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
      if (e1 == cudaSuccess) kern<<<grid, nw * 32, smem, st>>>(p);                              \
    } else if (tw) {                                                                                   \
      auto kern = k_tab_fwd<FP, NC, true, GG>;                                                  \
      e1 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
      if (e1 == cudaSuccess) kern<<<grid, nw * 32, smem, st>>>(p);                              \
    } else {                                                                                    \
      auto kern = k_tab_fwd<FP, NC, false, GG>;                                                 \
      e1 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
      if (e1 == cudaSuccess) kern<<<grid, nw * 32, smem, st>>>(p);                              \
    }                                                                                           \
  } while (0)
  if (launched) {
  } else if (nc == 1)
    DPB_LAUNCH_FWD(1);
  else if (nc == 2)
    DPB_LAUNCH_FWD(2);
  else
    DPB_LAUNCH_FWD(4);
#undef DPB_LAUNCH_FWD
  if (e1 == cudaSuccess) e1 = cudaGetLastError();
  cudaFreeAsync(scratch, st);
  if (scratch3) cudaFreeAsync(scratch3, st);
  DPB_CUDA(e1);
  note_launches(scratch3 ? 3 : 2);
  return DPB200_OK;
}

template <typename FP>
int launch_grad(FP* dy_dem_x, FP* dy_dem, FP* dy_dtwo, const FP* table, const FP* info,
                const FP* em_x, long long ldx_i, int ldx_j, const FP* em, long long ldem_i,
                const FP* two, const FP* dy, int nloc, int nnei, int M, int is_sorted,
                cudaStream_t st, int flags = 0, const GateArgs<FP>* ga = nullptr) {
  DPB_REQUIRE(nloc >= 0 && nnei >= 0 && M >= 0, "tabulate grad: negative size");
  if (nloc == 0 || nnei == 0) return DPB200_OK;  // tabulate.cc: nothing to write
  DPB_REQUIRE(dy_dem != nullptr && dy != nullptr, "tabulate grad: null pointer");
  DPB_REQUIRE(two == nullptr || dy_dtwo != nullptr, "tabulate grad: dy_dtwo is null");
  DPB_REQUIRE(!ga || (two == nullptr && ga->tt && ga->pair && ga->sw && ga->q), "tabulate gate grad: bad arguments");
  TabParams<FP> p = {};
  int rc = fill_info(p, info, M);
  if (rc) return rc;
  rc = common_args(p, table, em_x, ldx_i, ldx_j, em, ldem_i, two, nloc, nnei, is_sorted);
  if (rc) return rc;
  p.dy = dy;
  p.dy_dem_x = dy_dem_x;
  p.dy_dem = dy_dem;
  p.dy_dtwo = dy_dtwo;
  if (ga) p.gate_tt = ga->tt, p.gate_pair = ga->pair, p.gate_sw = ga->sw, p.gate_q = ga->q;
  const int nw = 12;
  const bool tw = two != nullptr || ga != nullptr;
  const bool ring = two != nullptr && ((size_t)M * sizeof(FP)) % 16 == 0 && aligned16(two) && M <= 32 * 4;
  p.two_ring = ring ? 1 : 0;
  const size_t rec_bytes = (size_t)nw * 32 * sizeof(Rec<FP>) + (ring ? (size_t)nw * kRing * M * sizeof(FP) : 0);
  const int kt = (M + 3) / 4;
  // the tensor-core backward serves plain se_a and the pair-indexed gate (not a materialised two_embed)
  const bool mma_ok = std::is_same<FP, double>::value && two == nullptr && use_mma_path() &&
                      (kt == 8 || kt == 16 || kt == 20 || kt == 25 || kt == 32) && (!ga || grad_variant() == 1);
  const bool cm32 = sizeof(FP) == 4 && !tw && (flags & DPB200_TAB_COMPRESSED_COEF) && fwd_cm_enabled();
  const bool cm = cm32 || (mma_ok && (flags & DPB200_TAB_COMPRESSED_COEF) && grad_variant() == 1);
  p.Mc = M;
  p.nblk = cm ? 2 : 3;
  set_a5_scale(p, flags);
  size_hot_window(p, M, rec_bytes, p.nblk);
  const size_t smem = (size_t)p.hot_elems * sizeof(FP) + rec_bytes;
  FP* scratch = nullptr;
  FP* scratch3 = nullptr;
  rc = cm ? prepare_table_cm(p, &scratch, st, 0) : prepare_table(p, &scratch, st);
  if (rc) return rc;
  if (cm && !cm32) {
    rc = prepare_table_full(p, &scratch3, st);
    if (rc) return rc;
  }
  long long want = ((long long)nloc + nw - 1) / nw;
  const long long cap = sm_count();
  const int grid = (int)(want < cap ? want : cap);
  cudaError_t e1 = cudaSuccess;
  bool launched = false;
  if constexpr (std::is_same<FP, float>::value) {
    if (cm32 && use_mma_path() && (kt == 8 || kt == 16 || kt == 20 || kt == 25 || kt == 32)) {
      const int nwv = 16;
      const size_t extra = (size_t)nwv * 4 * grad_mma_tile_ld(kt) * sizeof(double);
      const size_t recv = (size_t)nwv * 32 * sizeof(Rec<FP>);
      size_hot_window(p, M, recv + extra, p.nblk);
      // the double tile behind the records must be 8-byte aligned: hot_elems is a multiple of 16 bytes, Rec is 32 bytes
      const size_t smemv = (size_t)p.hot_elems * sizeof(FP) + recv + extra;
      long long wantv = ((long long)nloc + nwv - 1) / nwv;
      const int gridv = (int)(wantv < cap ? wantv : cap);
#define DPB_LAUNCH_GRAD_MMA32(KT)                                                               \
  do {                                                                                          \
    auto kern = k_tab_grad_mma_f32<KT>;                                                         \
    e1 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemv);   \
    if (e1 == cudaSuccess) kern<<<gridv, nwv * 32, smemv, st>>>(p);                             \
  } while (0)
      switch (kt) {
        case 8: DPB_LAUNCH_GRAD_MMA32(8); break;
        case 16: DPB_LAUNCH_GRAD_MMA32(16); break;
        case 20: DPB_LAUNCH_GRAD_MMA32(20); break;
        case 25: DPB_LAUNCH_GRAD_MMA32(25); break;
        default: DPB_LAUNCH_GRAD_MMA32(32); break;
      }
#undef DPB_LAUNCH_GRAD_MMA32
      launched = true;
    }
  }
  if constexpr (std::is_same<FP, double>::value) {
    if (mma_ok) {
      // variant 0: B fragments in registers, 12 warps; 1: in shared memory, 16 warps; 2: 24 warps
      const int variant = grad_variant();
      const int nwv = variant == 0 ? 12 : (variant == 1 ? 16 : 24);
      const size_t extra = variant == 0 ? 0 : (size_t)nwv * 4 * grad_mma_tile_ld(kt) * sizeof(FP);
      const size_t recv = (size_t)nwv * 32 * sizeof(Rec<FP>);
      // window rows are padded by 64 bytes when their size is a multiple of 128: the two neighbours of a quarter-warp
      // then hit disjoint banks whenever their rows are adjacent (the common case for distance-sorted neighbours)
      p.hot_pad = ((size_t)M * 16 * p.nblk) % 128 == 0 ? grad_hot_pad() : 0;
      size_hot_window(p, M, recv + extra, p.nblk);
      if (cm && p.H > p.first) p.H = p.first;  // the window holds compressed (stride-0) rows only
      const size_t smemv = (size_t)p.hot_elems * sizeof(FP) + recv + extra;
      long long wantv = ((long long)nloc + nwv - 1) / nwv;
      const int gridv = (int)(wantv < cap ? wantv : cap);
#define DPB_LAUNCH_GRAD_MMA(KT)                                                                 \
  do {                                                                                          \
    if (ga && cm) {                                                                             \
      auto kern = k_tab_grad_mma<KT, true, 512, true, true>;                                    \
      e1 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemv); \
      if (e1 == cudaSuccess) kern<<<gridv, nwv * 32, smemv, st>>>(p);                           \
    } else if (ga) {                                                                            \
      auto kern = k_tab_grad_mma<KT, true, 512, false, true>;                                   \
      e1 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemv); \
      if (e1 == cudaSuccess) kern<<<gridv, nwv * 32, smemv, st>>>(p);                           \
    } else if (cm && grad_breg() > 0) {                                                         \
      auto kern = k_tab_grad_mma<KT, true, 512, true, false, 12>;                               \
      e1 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemv); \
      if (e1 == cudaSuccess) kern<<<gridv, nwv * 32, smemv, st>>>(p);                           \
    } else if (cm) {                                                                            \
      auto kern = k_tab_grad_mma<KT, true, 512, true>;                                          \
      e1 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemv); \
      if (e1 == cudaSuccess) kern<<<gridv, nwv * 32, smemv, st>>>(p);                           \
    } else if (variant == 0) {                                                                  \
      auto kern = k_tab_grad_mma<KT, false, 384>;                                               \
      e1 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemv); \
      if (e1 == cudaSuccess) kern<<<gridv, nwv * 32, smemv, st>>>(p);                           \
    } else if (variant == 1) {                                                                  \
      auto kern = k_tab_grad_mma<KT, true, 512>;                                                \
      e1 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemv); \
      if (e1 == cudaSuccess) kern<<<gridv, nwv * 32, smemv, st>>>(p);                           \
    } else {                                                                                    \
      auto kern = k_tab_grad_mma<KT, true, 768>;                                                \
      e1 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemv); \
      if (e1 == cudaSuccess) kern<<<gridv, nwv * 32, smemv, st>>>(p);                           \
    }                                                                                           \
  } while (0)
      switch (kt) {
        case 8: DPB_LAUNCH_GRAD_MMA(8); break;
        case 16: DPB_LAUNCH_GRAD_MMA(16); break;
        case 20: DPB_LAUNCH_GRAD_MMA(20); break;
        case 25: DPB_LAUNCH_GRAD_MMA(25); break;
        default: DPB_LAUNCH_GRAD_MMA(32); break;
      }
#undef DPB_LAUNCH_GRAD_MMA
      launched = true;
    }
  }
#define DPB_LAUNCH_GRAD(NC)                                                                     \
  do {                                                                                          \
    if (cm32) {                                                                                 \
      auto kern = k_tab_grad<FP, NC, false, true>;                                              \
      e1 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
      if (e1 == cudaSuccess) kern<<<grid, nw * 32, smem, st>>>(p);                              \
    } else if (tw) {                                                                                   \
      auto kern = k_tab_grad<FP, NC, true>;                                                     \
      e1 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
      if (e1 == cudaSuccess) kern<<<grid, nw * 32, smem, st>>>(p);                              \
    } else {                                                                                    \
      auto kern = k_tab_grad<FP, NC, false>;                                                    \
      e1 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
      if (e1 == cudaSuccess) kern<<<grid, nw * 32, smem, st>>>(p);                              \
    }                                                                                           \
  } while (0)
  if (launched) {
  } else if (M <= 32)
    DPB_LAUNCH_GRAD(1);
  else if (M <= 64)
    DPB_LAUNCH_GRAD(2);
  else
    DPB_LAUNCH_GRAD(4);
#undef DPB_LAUNCH_GRAD
  if (e1 == cudaSuccess) e1 = cudaGetLastError();
  cudaFreeAsync(scratch, st);
  if (scratch3) cudaFreeAsync(scratch3, st);
  DPB_CUDA(e1);
  note_launches(scratch3 ? 3 : 2);
  return DPB200_OK;
}

}  // namespace
}  // namespace dpb200

extern "C" {

#define DPB200_DEF_TAB(SUF, FP)                                                                    \
  int dpb200_tabulate_fusion_se_a_##SUF(FP* out, const FP* table, const FP* table_info,            \
                                        const FP* em_x, const FP* em, const FP* two_embed,         \
                                        int nloc, int nnei, int last_layer_size, int is_sorted,    \
                                        dpb200_stream_t stream) {                                  \
    return dpb200::launch_fwd<FP, false>(out, table, table_info, em_x, nnei, 1, em,                \
                                         (long long)nnei * 4, two_embed, nullptr, nullptr,         \
                                         nullptr, nloc, nnei, last_layer_size, is_sorted, 0,       \
                                         (cudaStream_t)stream);                                    \
  }                                                                                                \
  int dpb200_tabulate_fusion_se_a_ex_##SUF(FP* out, const FP* table, const FP* table_info,         \
                                           const FP* em_x, long long ldx_i, int ldx_j,             \
                                           const FP* em, long long ldem_i, const FP* two_embed,    \
                                           int nloc, int nnei, int last_layer_size,                \
                                           int is_sorted, int accumulate,                          \
                                           dpb200_stream_t stream) {                               \
    return dpb200::launch_fwd<FP, false>(out, table, table_info, em_x, ldx_i, ldx_j, em, ldem_i,   \
                                         two_embed, nullptr, nullptr, nullptr, nloc, nnei,         \
                                         last_layer_size, is_sorted, accumulate,                   \
                                         (cudaStream_t)stream);                                    \
  }                                                                                                \
  int dpb200_tabulate_fusion_se_a_desc_##SUF(                                                      \
      FP* out, const FP* table, const FP* table_info, const FP* em_x, long long ldx_i, int ldx_j,  \
      const FP* em, long long ldem_i, int nloc, int nnei, int last_layer_size, int is_sorted,      \
      int accumulate, int axis, double scale, const int* desc_row, int desc_mode, void* desc,      \
      long long desc_ld, int nslice, int* row_exp, int flags, dpb200_stream_t stream) {            \
    dpb200::DescArgs da = {desc, desc_ld, desc_row, row_exp, scale, desc_mode, axis, nslice};      \
    return dpb200::launch_fwd<FP, false>(out, table, table_info, em_x, ldx_i, ldx_j, em, ldem_i,   \
                                         nullptr, nullptr, nullptr, nullptr, nloc, nnei,           \
                                         last_layer_size, is_sorted, accumulate,                   \
                                         (cudaStream_t)stream, desc_mode == 0 ? nullptr : &da,     \
                                         flags);                                                   \
  }                                                                                                \
  int dpb200_tabulate_fusion_se_a_grad_fx_##SUF(                                                   \
      FP* dy_dem_x, FP* dy_dem, const FP* table, const FP* table_info, const FP* em_x,             \
      long long ldx_i, int ldx_j, const FP* em, long long ldem_i, const FP* dy, int nloc,          \
      int nnei, int last_layer_size, int is_sorted, int flags, dpb200_stream_t stream) {           \
    return dpb200::launch_grad<FP>(dy_dem_x, dy_dem, nullptr, table, table_info, em_x, ldx_i,      \
                                   ldx_j, em, ldem_i, nullptr, dy, nloc, nnei, last_layer_size,    \
                                   is_sorted, (cudaStream_t)stream, flags);                        \
  }                                                                                                \
  int dpb200_tabulate_fusion_se_a_grad_##SUF(FP* dy_dem_x, FP* dy_dem, FP* dy_dtwo,                \
                                             const FP* table, const FP* table_info,                \
                                             const FP* em_x, const FP* em, const FP* two_embed,    \
                                             const FP* dy, int nloc, int nnei,                     \
                                             int last_layer_size, int is_sorted,                   \
                                             dpb200_stream_t stream) {                             \
    return dpb200::launch_grad<FP>(dy_dem_x, dy_dem, dy_dtwo, table, table_info, em_x, nnei, 1,    \
                                   em, (long long)nnei * 4, two_embed, dy, nloc, nnei,             \
                                   last_layer_size, is_sorted, (cudaStream_t)stream);              \
  }                                                                                                \
  int dpb200_tabulate_fusion_se_a_grad_ex_##SUF(                                                   \
      FP* dy_dem_x, FP* dy_dem, FP* dy_dtwo, const FP* table, const FP* table_info,                \
      const FP* em_x, long long ldx_i, int ldx_j, const FP* em, long long ldem_i,                  \
      const FP* two_embed, const FP* dy, int nloc, int nnei, int last_layer_size, int is_sorted,   \
      dpb200_stream_t stream) {                                                                    \
    return dpb200::launch_grad<FP>(dy_dem_x, dy_dem, dy_dtwo, table, table_info, em_x, ldx_i,      \
                                   ldx_j, em, ldem_i, two_embed, dy, nloc, nnei, last_layer_size,  \
                                   is_sorted, (cudaStream_t)stream);                               \
  }                                                                                                \
  int dpb200_tabulate_fusion_se_atten_gate_##SUF(                                                  \
      FP* out, const FP* table, const FP* table_info, const FP* em_x, const FP* em,                \
      const FP* tt_full, const int* pair, const FP* sw, int nloc, int nnei, int last_layer_size,   \
      int is_sorted, int flags, dpb200_stream_t stream) {                                          \
    dpb200::GateArgs<FP> ga = {tt_full, pair, sw, nullptr};                                        \
    return dpb200::launch_fwd<FP, false>(out, table, table_info, em_x, nnei, 1, em,                \
                                         (long long)nnei * 4, nullptr, nullptr, nullptr, nullptr,  \
                                         nloc, nnei, last_layer_size, is_sorted, 0,                \
                                         (cudaStream_t)stream, nullptr, flags, &ga);               \
  }                                                                                                \
  int dpb200_tabulate_fusion_se_atten_gate_desc_##SUF(                                             \
      FP* out, const FP* table, const FP* table_info, const FP* em_x, const FP* em,                \
      const FP* tt_full, const int* pair, const FP* sw, int nloc, int nnei, int last_layer_size,   \
      int is_sorted, int axis, double scale, const int* desc_row, int desc_mode, void* desc,       \
      long long desc_ld, long long slice_stride, int nslice, int* row_exp, int min_row_exp,        \
      int flags, dpb200_stream_t stream) {                                                         \
    dpb200::GateArgs<FP> ga = {tt_full, pair, sw, nullptr};                                        \
    dpb200::DescArgs da = {desc, desc_ld, desc_row, row_exp, scale, desc_mode, axis, nslice,       \
                           slice_stride, min_row_exp};                                             \
    return dpb200::launch_fwd<FP, false>(out, table, table_info, em_x, nnei, 1, em,                \
                                         (long long)nnei * 4, nullptr, nullptr, nullptr, nullptr,  \
                                         nloc, nnei, last_layer_size, is_sorted, 0,                \
                                         (cudaStream_t)stream, desc_mode == 0 ? nullptr : &da,     \
                                         flags, &ga);                                              \
  }                                                                                                \
  int dpb200_tabulate_fusion_se_atten_gate_grad_##SUF(                                             \
      FP* dy_dem_x, FP* dy_dem, FP* dy_dsw, const FP* table, const FP* table_info,                 \
      const FP* em_x, const FP* em, const FP* tt_full, const int* pair, const FP* sw,              \
      const FP* dy, int nloc, int nnei, int last_layer_size, int is_sorted, int flags,             \
      dpb200_stream_t stream) {                                                                    \
    dpb200::GateArgs<FP> ga = {tt_full, pair, sw, dy_dsw};                                         \
    return dpb200::launch_grad<FP>(dy_dem_x, dy_dem, nullptr, table, table_info, em_x, nnei, 1,    \
                                   em, (long long)nnei * 4, nullptr, dy, nloc, nnei,               \
                                   last_layer_size, is_sorted, (cudaStream_t)stream, flags, &ga);  \
  }                                                                                                \
  int dpb200_tabulate_fusion_se_a_grad_grad_##SUF(                                                 \
      FP* dz_dy, const FP* table, const FP* table_info, const FP* em_x, const FP* em,              \
      const FP* two_embed, const FP* dz_dy_dem_x, const FP* dz_dy_dem, const FP* dz_dy_dtwo,       \
      int nloc, int nnei, int last_layer_size, int is_sorted, dpb200_stream_t stream) {            \
    return dpb200::launch_fwd<FP, true>(dz_dy, table, table_info, em_x, nnei, 1, em,               \
                                        (long long)nnei * 4, two_embed, dz_dy_dem_x, dz_dy_dem,    \
                                        dz_dy_dtwo, nloc, nnei, last_layer_size, is_sorted, 0,     \
                                        (cudaStream_t)stream);                                     \
  }
DPB200_DEF_TAB(f64, double)
DPB200_DEF_TAB(f32, float)
#undef DPB200_DEF_TAB

}  // extern "C"
