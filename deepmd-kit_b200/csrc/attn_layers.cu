// DPA-1 neighbour-gated self-attention layers (deepmd/pt/model/descriptor/se_atten.py:1058-1447:
// NeighborGatedAttention -> NeighborGatedAttentionLayer -> GatedAttentionLayer) around the per-neighbour embedding
// g2 of the strip-mode descriptor (se_atten.py:977-1012).  The dense products of a layer (in_proj, q k^T, A v,
// out_proj and their transposes) are plain GEMMs and are issued by the host side (atten.py) on the library; the
// kernels here are everything between them, fused per stage, forward and hand-derived backward:
//
//   embed        g_s(s), g_s'(s) of every (atom, neighbour) from the `dp compress` quintic table and
//                x0 = g_s (1 + tt_full[pair] sw)                                       (se_atten.py:979-1003)
//   rhat         input_r = normalize(rr[:, 1:4])                                       (se_atten.py:1004-1006)
//   qkv_norm     q, k, v <- x / max(|x|, 1e-12) per row, q additionally * scaling      (:1368-1373)
//   weights      T = (S + shift) sw_i sw_j - shift; P = softmax_j T; A = P sw_i sw_j (rhat_i . rhat_j)   (:1386-1416)
//   res_ln       X' = LayerNorm(X + Y) * gamma + beta                                  (:1288-1291)
//
// and their transposes, which accumulate dE/d(sw) and dE/d(rhat) over the layers.  Padded neighbour slots need no
// special case: their sw is 0, so their row and column of A vanish, while they still count exp(-shift) in every
// softmax denominator, exactly as in the reference.
#include <cmath>

#include "common.cuh"

namespace dpb200 {
namespace {

// ------------------------------------------------------------------------------------------ embed
template <typename FP>
struct EmbedTab {
  const FP* table;  // [rows][M][6], the reference's layout
  FP lower, upper, vmax, s0, s1, tail_xx;
  int first, tail_idx, M;
};

template <typename FP>
int fill_tab(EmbedTab<FP>& p, const FP* table, const FP* info, int M) {
  DPB_REQUIRE(table && info, "se_atten_embed: null table (table_info is a HOST pointer)");
  p.table = table;
  p.lower = info[0];
  p.upper = info[1];
  p.vmax = info[2];
  p.s0 = info[3];
  p.s1 = info[4];
  DPB_REQUIRE(p.s0 > (FP)0. && p.s1 > (FP)0. && p.upper >= p.lower && p.vmax >= p.upper,
              "se_atten_embed: table_info must satisfy lower <= upper <= max and positive strides");
  // source/lib/src/tabulate.cc:21-30, 45-73
  p.first = (int)((p.upper - p.lower) / p.s0);
  const FP edge = std::nextafter(p.vmax, p.lower);
  p.tail_idx = p.first + (int)((edge - p.upper) / p.s1);
  p.tail_xx = p.vmax - ((FP)(p.tail_idx - p.first) * p.s1 + p.upper);
  p.M = M;
  return DPB200_OK;
}

template <typename FP>
__device__ __forceinline__ void locate(const EmbedTab<FP>& p, FP x0, FP& xx, int& idx, FP& delta) {
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

// one warp per (atom, neighbour) row, lanes over the channels
template <typename FP>
__global__ void __launch_bounds__(256) k_embed(const __grid_constant__ EmbedTab<FP> p, FP* __restrict__ x0,
                                               FP* __restrict__ gs, FP* __restrict__ dgs,
                                               const FP* __restrict__ em_x, long long em_stride,
                                               const FP* __restrict__ tt, const int* __restrict__ pair,
                                               const FP* __restrict__ sw, long long rows) {
  const int lane = threadIdx.x & 31;
  const long long wpg = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += wpg) {
    const FP s = em_x[r * em_stride];
    FP xx, dl;
    int idx;
    locate(p, s, xx, idx, dl);
    const FP w = sw[r];
    const FP* __restrict__ trow = tt + (long long)pair[r] * p.M;
    const FP* __restrict__ a = p.table + (long long)idx * p.M * 6;
    for (int c = lane; c < p.M; c += 32) {
      const FP a0 = a[c * 6], a1 = a[c * 6 + 1], a2 = a[c * 6 + 2], a3 = a[c * 6 + 3], a4 = a[c * 6 + 4],
               a5 = a[c * 6 + 5];
      const FP g1 = a1 + ((FP)2. * a2 + ((FP)3. * a3 + ((FP)4. * a4 + (FP)5. * a5 * xx) * xx) * xx) * xx;
      const FP g0 = a0 + (a1 + (a2 + (a3 + (a4 + a5 * xx) * xx) * xx) * xx) * xx + g1 * dl;
      const long long o = r * p.M + c;
      gs[o] = g0;
      dgs[o] = g1;
      x0[o] = g0 * trow[c] * w + g0;
    }
  }
}

// d_s[row] = sum_c dx0 (1 + tt sw) g_s' ;  d_sw[row] += sum_c dx0 g_s tt
template <typename FP>
__global__ void __launch_bounds__(256) k_embed_grad(FP* __restrict__ d_s, long long ds_stride, FP* __restrict__ d_sw,
                                                    const FP* __restrict__ dx0, const FP* __restrict__ gs,
                                                    const FP* __restrict__ dgs, const FP* __restrict__ tt,
                                                    const int* __restrict__ pair, const FP* __restrict__ sw,
                                                    long long rows, int M) {
  const int lane = threadIdx.x & 31;
  const long long wpg = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += wpg) {
    const FP w = sw[r];
    const FP* __restrict__ trow = tt + (long long)pair[r] * M;
    FP as = (FP)0., aw = (FP)0.;
    for (int c = lane; c < M; c += 32) {
      const long long o = r * M + c;
      const FP d = dx0[o], t = trow[c];
      as += d * (t * w + (FP)1.) * dgs[o];
      aw += d * gs[o] * t;
    }
    as = warp_sum(as);
    aw = warp_sum(aw);
    if (lane == 0) {
      d_s[r * ds_stride] += as;
      d_sw[r] += aw;
    }
  }
}

// ------------------------------------------------------------------------------------------ rhat
constexpr double kNormEps = 1e-12;  // torch.nn.functional.normalize default

template <typename FP>
__global__ void k_rhat(FP* __restrict__ rhat, FP* __restrict__ rinv, const FP* __restrict__ em, long long rows) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
    const FP x = em[4 * r + 1], y = em[4 * r + 2], z = em[4 * r + 3];
    const FP n = sqrt(x * x + y * y + z * z);
    const bool tiny = n < (FP)kNormEps;
    const FP inv = (FP)1. / (tiny ? (FP)kNormEps : n);
    rhat[3 * r] = x * inv;
    rhat[3 * r + 1] = y * inv;
    rhat[3 * r + 2] = z * inv;
    rinv[r] = tiny ? -inv : inv;  // the sign records the clamp: no projection term in the derivative then
  }
}

template <typename FP>
__global__ void k_rhat_grad(FP* __restrict__ d_em, const FP* __restrict__ d_rhat, const FP* __restrict__ rhat,
                            const FP* __restrict__ rinv, long long rows) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
    const FP gx = d_rhat[3 * r], gy = d_rhat[3 * r + 1], gz = d_rhat[3 * r + 2];
    const FP inv = rinv[r];
    FP dot = (FP)0.;
    if (inv > (FP)0.) dot = gx * rhat[3 * r] + gy * rhat[3 * r + 1] + gz * rhat[3 * r + 2];
    const FP a = fabs(inv);
    d_em[4 * r + 1] += (gx - rhat[3 * r] * dot) * a;
    d_em[4 * r + 2] += (gy - rhat[3 * r + 1] * dot) * a;
    d_em[4 * r + 3] += (gz - rhat[3 * r + 2] * dot) * a;
  }
}

// ------------------------------------------------------------------------------------------ qkv_norm
// one warp per (row, segment): qkv [rows][3][h] in place; inv [rows][3] (negative = clamped norm)
template <typename FP, bool GRAD>
__global__ void __launch_bounds__(256) k_qkv_norm(FP* __restrict__ qkv, FP* __restrict__ inv_io,
                                                  const FP* __restrict__ yhat, long long rows, int h, FP qscale,
                                                  int normalize) {
  const int lane = threadIdx.x & 31;
  const long long wpg = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long u = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); u < rows * 3; u += wpg) {
    const int seg = (int)(u % 3);
    const FP sc = seg == 0 ? qscale : (FP)1.;
    FP* __restrict__ x = qkv + u * h;
    if (!GRAD) {
      FP inv = (FP)1.;
      if (normalize) {
        FP ss = (FP)0.;
        for (int c = lane; c < h; c += 32) ss += x[c] * x[c];
        const FP n = sqrt(warp_sum(ss));
        const bool tiny = n < (FP)kNormEps;
        inv = (FP)1. / (tiny ? (FP)kNormEps : n);
        if (lane == 0) inv_io[u] = tiny ? -inv : inv;
      }
      for (int c = lane; c < h; c += 32) x[c] *= inv * sc;
    } else {
      // x holds d/d(stored y), stored y = sc * yhat: d/dx = sc * (g - yhat (yhat . g)) * inv
      const FP* __restrict__ y = yhat + u * h;
      if (!normalize) {
        for (int c = lane; c < h; c += 32) x[c] *= sc;
        continue;
      }
      const FP inv = inv_io[u];
      FP dot = (FP)0.;
      if (inv > (FP)0.) {
        for (int c = lane; c < h; c += 32) dot += x[c] * y[c];
        dot = warp_sum(dot) / sc;  // y = sc * yhat
      }
      const FP a = fabs(inv) * sc;
      for (int c = lane; c < h; c += 32) x[c] = (x[c] - (y[c] / sc) * dot) * a;
    }
  }
}

// ------------------------------------------------------------------------------------------ weights
// Lanes over the neighbours j, NJ per lane (NJ = 4 serves nnei <= 128, NJ = 8 nnei <= 256).
//
// Slabs with trailing empty slots may be evaluated on the first n < n_full slots only (the host keeps one empty slot
// as the representative of all of them): an empty column has T = -shift whatever S is, so the n_full - n omitted ones
// add (n_full - n) exp(-shift) to every softmax denominator and nothing else.
constexpr int kMaxJ = 8;

// one warp per (atom, i)
template <typename FP, int NJ>
__global__ void __launch_bounds__(256) k_attn_weights(FP* __restrict__ P, FP* __restrict__ A, const FP* __restrict__ S,
                                                      const FP* __restrict__ sw, const FP* __restrict__ rhat,
                                                      long long natoms, int n, int n_full, FP shift, int dotr) {
  const int lane = threadIdx.x & 31;
  const long long wpg = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long u = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); u < natoms * n; u += wpg) {
    const long long atom = u / n;
    const FP* __restrict__ swa = sw + atom * n;
    const FP* __restrict__ ra = rhat + atom * n * 3;
    const FP swi = sw[u];
    const FP rx = rhat[3 * u], ry = rhat[3 * u + 1], rz = rhat[3 * u + 2];
    const FP* __restrict__ srow = S + u * n;
    FP t[NJ], w[NJ];
    FP mx = n_full > n ? -shift : -INFINITY;
#pragma unroll
    for (int k = 0; k < NJ; ++k) {
      const int j = lane + 32 * k;
      if (j < n) {
        w[k] = swi * swa[j];
        t[k] = (srow[j] + shift) * w[k] - shift;
        mx = fmax(mx, t[k]);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(kFull, mx, o));
    FP sum = (FP)0.;
#pragma unroll
    for (int k = 0; k < NJ; ++k) {
      const int j = lane + 32 * k;
      if (j < n) {
        t[k] = exp(t[k] - mx);
        sum += t[k];
      }
    }
    sum = warp_sum(sum);
    if (n_full > n) sum += (FP)(n_full - n) * exp(-shift - mx);
    const FP rs = (FP)1. / sum;
#pragma unroll
    for (int k = 0; k < NJ; ++k) {
      const int j = lane + 32 * k;
      if (j < n) {
        const FP pj = t[k] * rs;
        FP ww = w[k];
        if (dotr) ww *= rx * ra[3 * j] + ry * ra[3 * j + 1] + rz * ra[3 * j + 2];
        st_cs(P + u * n + j, pj);
        A[u * n + j] = pj * ww;
      }
    }
  }
}

// one CTA per atom, one warp per row i; dS may alias dA.  Column sums (the j side of d sw_i sw_j and d rhat_i . rhat_j)
// stay in registers over the rows of a warp and meet the row sums in shared memory once per warp.
// d_sw [natoms][n] and d_rhat [natoms][n][3] are accumulated into.
template <typename FP, int NJ>
__global__ void __launch_bounds__(128, NJ == 4 ? 3 : 1) k_attn_weights_grad(FP* __restrict__ dS, FP* __restrict__ d_sw,
                                                           FP* __restrict__ d_rhat, const FP* __restrict__ dA,
                                                           const FP* __restrict__ P, const FP* __restrict__ S,
                                                           const FP* __restrict__ sw, const FP* __restrict__ rhat,
                                                           long long natoms, int n, FP shift, int dotr) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FP* s_sw = reinterpret_cast<FP*>(smem_raw);  // [n]
  FP* s_r = s_sw + n;                          // [n][3]
  FP* c_sw = s_r + 3 * n;                      // [n]     accumulators of this atom
  FP* c_r = c_sw + n;                          // [n][3]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (long long atom = blockIdx.x; atom < natoms; atom += gridDim.x) {
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
      s_sw[j] = sw[atom * n + j];
      c_sw[j] = (FP)0.;
    }
    for (int j = threadIdx.x; j < 3 * n; j += blockDim.x) {
      s_r[j] = rhat[atom * n * 3 + j];
      c_r[j] = (FP)0.;
    }
    __syncthreads();
    FP csw[NJ], crx[NJ], cry[NJ], crz[NJ], swj[NJ], rjx[NJ], rjy[NJ], rjz[NJ];
#pragma unroll
    for (int k = 0; k < NJ; ++k) {
      const int j = lane + 32 * k;
      const bool in = j < n;
      csw[k] = crx[k] = cry[k] = crz[k] = (FP)0.;
      swj[k] = in ? s_sw[j] : (FP)0.;
      rjx[k] = in ? s_r[3 * j] : (FP)0.;
      rjy[k] = in ? s_r[3 * j + 1] : (FP)0.;
      rjz[k] = in ? s_r[3 * j + 2] : (FP)0.;
    }
    for (int i = warp; i < n; i += nwarp) {
      const long long u = atom * n + i;
      const FP swi = s_sw[i];
      const FP rx = s_r[3 * i], ry = s_r[3 * i + 1], rz = s_r[3 * i + 2];
      FP da[NJ], pp[NJ], ss[NJ];
#pragma unroll
      for (int k = 0; k < NJ; ++k) {
        const int j = lane + 32 * k;
        if (j < n) {
          da[k] = __ldcs(dA + u * n + j);
          pp[k] = __ldcs(P + u * n + j);
          ss[k] = __ldcs(S + u * n + j);
        } else {
          da[k] = pp[k] = ss[k] = (FP)0.;
        }
      }
      FP dot = (FP)0.;
#pragma unroll
      for (int k = 0; k < NJ; ++k) {
        const FP rr = dotr ? rx * rjx[k] + ry * rjy[k] + rz * rjz[k] : (FP)1.;
        dot += da[k] * swi * swj[k] * rr * pp[k];  // sum_j dP_ij P_ij
      }
      dot = warp_sum(dot);
      FP row_sw = (FP)0., row_rx = (FP)0., row_ry = (FP)0., row_rz = (FP)0.;
#pragma unroll
      for (int k = 0; k < NJ; ++k) {
        const int j = lane + 32 * k;
        const FP ww = swi * swj[k];
        const FP rr = dotr ? rx * rjx[k] + ry * rjy[k] + rz * rjz[k] : (FP)1.;
        const FP dT = pp[k] * (da[k] * ww * rr - dot);
        const FP dw = da[k] * pp[k] * rr + dT * (ss[k] + shift);  // d / d(sw_i sw_j)
        if (j < n) dS[u * n + j] = dT * ww;
        row_sw += dw * swj[k];
        csw[k] += dw * swi;
        if (dotr) {
          const FP dR = da[k] * pp[k] * ww;
          row_rx += dR * rjx[k];
          row_ry += dR * rjy[k];
          row_rz += dR * rjz[k];
          crx[k] += dR * rx;
          cry[k] += dR * ry;
          crz[k] += dR * rz;
        }
      }
      row_sw = warp_sum(row_sw);
      if (dotr) {
        row_rx = warp_sum(row_rx);
        row_ry = warp_sum(row_ry);
        row_rz = warp_sum(row_rz);
      }
      if (lane == 0) {
        atomicAdd(&c_sw[i], row_sw);
        if (dotr) {
          atomicAdd(&c_r[3 * i], row_rx);
          atomicAdd(&c_r[3 * i + 1], row_ry);
          atomicAdd(&c_r[3 * i + 2], row_rz);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < NJ; ++k) {
      const int j = lane + 32 * k;
      if (j < n) {
        atomicAdd(&c_sw[j], csw[k]);
        if (dotr) {
          atomicAdd(&c_r[3 * j], crx[k]);
          atomicAdd(&c_r[3 * j + 1], cry[k]);
          atomicAdd(&c_r[3 * j + 2], crz[k]);
        }
      }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += blockDim.x) d_sw[atom * n + j] += c_sw[j];
    if (dotr)
      for (int j = threadIdx.x; j < 3 * n; j += blockDim.x) d_rhat[atom * n * 3 + j] += c_r[j];
  }
}

// ------------------------------------------------------------------------------------------ res_ln
// one warp per row: z = x + y; zhat = (z - mean) * rstd (biased variance); out = zhat * gamma + beta; y <- zhat
template <typename FP>
__global__ void __launch_bounds__(256) k_res_ln(FP* __restrict__ out, FP* __restrict__ y_zhat, FP* __restrict__ rstd,
                                                const FP* __restrict__ x, const FP* __restrict__ gamma,
                                                const FP* __restrict__ beta, long long rows, int C, FP eps) {
  const int lane = threadIdx.x & 31;
  const long long wpg = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += wpg) {
    FP s = (FP)0.;
    for (int c = lane; c < C; c += 32) {
      const FP z = x[r * C + c] + y_zhat[r * C + c];
      y_zhat[r * C + c] = z;
      s += z;
    }
    const FP mean = warp_sum(s) / (FP)C;
    FP v = (FP)0.;
    for (int c = lane; c < C; c += 32) {
      const FP d = y_zhat[r * C + c] - mean;
      v += d * d;
    }
    const FP rs = (FP)1. / sqrt(warp_sum(v) / (FP)C + eps);
    if (lane == 0) rstd[r] = rs;
    for (int c = lane; c < C; c += 32) {
      const FP zh = (y_zhat[r * C + c] - mean) * rs;
      y_zhat[r * C + c] = zh;
      out[r * C + c] = zh * gamma[c] + beta[c];
    }
  }
}

// dz = rstd (g - mean(g) - zhat mean(g zhat)), g = dout * gamma
template <typename FP>
__global__ void __launch_bounds__(256) k_res_ln_grad(FP* __restrict__ dz, const FP* __restrict__ dout,
                                                     const FP* __restrict__ zhat, const FP* __restrict__ rstd,
                                                     const FP* __restrict__ gamma, long long rows, int C) {
  const int lane = threadIdx.x & 31;
  const long long wpg = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += wpg) {
    FP s1 = (FP)0., s2 = (FP)0.;
    for (int c = lane; c < C; c += 32) {
      const FP g = dout[r * C + c] * gamma[c];
      s1 += g;
      s2 += g * zhat[r * C + c];
    }
    s1 = warp_sum(s1) / (FP)C;
    s2 = warp_sum(s2) / (FP)C;
    const FP rs = rstd[r];
    for (int c = lane; c < C; c += 32) {
      const FP g = dout[r * C + c] * gamma[c];
      dz[r * C + c] = rs * (g - s1 - zhat[r * C + c] * s2);
    }
  }
}

inline unsigned warp_grid(long long warps, int warps_per_cta) {
  long long grid = (warps + warps_per_cta - 1) / warps_per_cta;
  const long long cap = (long long)sm_count() * 32;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  return (unsigned)grid;
}

template <typename FP>
int launch_embed(FP* x0, FP* gs, FP* dgs, const FP* table, const FP* info, const FP* em_x, long long em_stride,
                 const FP* tt, const int* pair, const FP* sw, long long rows, int M, cudaStream_t st) {
  DPB_REQUIRE(rows >= 0 && M >= 0 && em_stride >= 1, "se_atten_embed: bad sizes");
  if (rows == 0 || M == 0) return DPB200_OK;
  DPB_REQUIRE(x0 && gs && dgs && em_x && tt && pair && sw, "se_atten_embed: null pointer");
  EmbedTab<FP> p = {};
  int rc = fill_tab(p, table, info, M);
  if (rc) return rc;
  k_embed<FP><<<warp_grid(rows, 8), 256, 0, st>>>(p, x0, gs, dgs, em_x, em_stride, tt, pair, sw, rows);
  DPB_CUDA(cudaGetLastError());
  note_launches(1);
  return DPB200_OK;
}

template <typename FP>
int launch_embed_grad(FP* d_s, long long ds_stride, FP* d_sw, const FP* dx0, const FP* gs, const FP* dgs, const FP* tt,
                      const int* pair, const FP* sw, long long rows, int M, cudaStream_t st) {
  DPB_REQUIRE(rows >= 0 && M >= 0 && ds_stride >= 1, "se_atten_embed_grad: bad sizes");
  if (rows == 0 || M == 0) return DPB200_OK;
  DPB_REQUIRE(d_s && d_sw && dx0 && gs && dgs && tt && pair && sw, "se_atten_embed_grad: null pointer");
  k_embed_grad<FP><<<warp_grid(rows, 8), 256, 0, st>>>(d_s, ds_stride, d_sw, dx0, gs, dgs, tt, pair, sw, rows, M);
  DPB_CUDA(cudaGetLastError());
  note_launches(1);
  return DPB200_OK;
}

template <typename FP>
int launch_rhat(FP* rhat, FP* rinv, const FP* em, long long rows, cudaStream_t st) {
  DPB_REQUIRE(rows >= 0, "se_atten_rhat: bad sizes");
  if (rows == 0) return DPB200_OK;
  DPB_REQUIRE(rhat && rinv && em, "se_atten_rhat: null pointer");
  k_rhat<FP><<<warp_grid(rows, 256), 256, 0, st>>>(rhat, rinv, em, rows);
  DPB_CUDA(cudaGetLastError());
  note_launches(1);
  return DPB200_OK;
}

template <typename FP>
int launch_rhat_grad(FP* d_em, const FP* d_rhat, const FP* rhat, const FP* rinv, long long rows, cudaStream_t st) {
  DPB_REQUIRE(rows >= 0, "se_atten_rhat_grad: bad sizes");
  if (rows == 0) return DPB200_OK;
  DPB_REQUIRE(d_em && d_rhat && rhat && rinv, "se_atten_rhat_grad: null pointer");
  k_rhat_grad<FP><<<warp_grid(rows, 256), 256, 0, st>>>(d_em, d_rhat, rhat, rinv, rows);
  DPB_CUDA(cudaGetLastError());
  note_launches(1);
  return DPB200_OK;
}

template <typename FP, bool GRAD>
int launch_qkv_norm(FP* qkv, FP* inv, const FP* yhat, long long rows, int h, double qscale, int normalize,
                    cudaStream_t st) {
  DPB_REQUIRE(rows >= 0 && h >= 1 && qscale != 0., "attn_qkv_normalize: bad sizes");
  if (rows == 0) return DPB200_OK;
  DPB_REQUIRE(qkv && (!normalize || inv) && (!GRAD || yhat), "attn_qkv_normalize: null pointer");
  k_qkv_norm<FP, GRAD><<<warp_grid(rows * 3, 8), 256, 0, st>>>(qkv, inv, yhat, rows, h, (FP)qscale, normalize);
  DPB_CUDA(cudaGetLastError());
  note_launches(1);
  return DPB200_OK;
}

template <typename FP>
int launch_weights(FP* P, FP* A, const FP* S, const FP* sw, const FP* rhat, long long natoms, int n, int n_full,
                   double shift, int dotr, cudaStream_t st) {
  DPB_REQUIRE(natoms >= 0 && n >= 1 && n <= 32 * kMaxJ && n_full >= n, "attn_weights: nnei must be in 1..256, nnei_full >= nnei");
  if (natoms == 0) return DPB200_OK;
  DPB_REQUIRE(P && A && S && sw && rhat, "attn_weights: null pointer");
  const unsigned grid = warp_grid(natoms * n, 8);
  if (n <= 128)
    k_attn_weights<FP, 4><<<grid, 256, 0, st>>>(P, A, S, sw, rhat, natoms, n, n_full, (FP)shift, dotr);
  else
    k_attn_weights<FP, 8><<<grid, 256, 0, st>>>(P, A, S, sw, rhat, natoms, n, n_full, (FP)shift, dotr);
  DPB_CUDA(cudaGetLastError());
  note_launches(1);
  return DPB200_OK;
}

template <typename FP>
int launch_weights_grad(FP* dS, FP* d_sw, FP* d_rhat, const FP* dA, const FP* P, const FP* S, const FP* sw,
                        const FP* rhat, long long natoms, int n, double shift, int dotr, cudaStream_t st) {
  DPB_REQUIRE(natoms >= 0 && n >= 1 && n <= 32 * kMaxJ, "attn_weights_grad: nnei must be in 1..256");
  if (natoms == 0) return DPB200_OK;
  DPB_REQUIRE(dS && d_sw && d_rhat && dA && P && S && sw && rhat, "attn_weights_grad: null pointer");
  long long grid = natoms;
  const long long cap = (long long)sm_count() * 12;
  if (grid > cap) grid = cap;
  const size_t smem = sizeof(FP) * 8 * (size_t)n;
  if (n <= 128)
    k_attn_weights_grad<FP, 4><<<(unsigned)grid, 128, smem, st>>>(dS, d_sw, d_rhat, dA, P, S, sw, rhat, natoms, n,
                                                                  (FP)shift, dotr);
  else
    k_attn_weights_grad<FP, 8><<<(unsigned)grid, 128, smem, st>>>(dS, d_sw, d_rhat, dA, P, S, sw, rhat, natoms, n,
                                                                  (FP)shift, dotr);
  DPB_CUDA(cudaGetLastError());
  note_launches(1);
  return DPB200_OK;
}

template <typename FP>
int launch_res_ln(FP* out, FP* y_zhat, FP* rstd, const FP* x, const FP* gamma, const FP* beta, long long rows, int C,
                  double eps, cudaStream_t st) {
  DPB_REQUIRE(rows >= 0 && C >= 1 && eps >= 0., "attn_residual_layernorm: bad sizes");
  if (rows == 0) return DPB200_OK;
  DPB_REQUIRE(out && y_zhat && rstd && x && gamma && beta, "attn_residual_layernorm: null pointer");
  k_res_ln<FP><<<warp_grid(rows, 8), 256, 0, st>>>(out, y_zhat, rstd, x, gamma, beta, rows, C, (FP)eps);
  DPB_CUDA(cudaGetLastError());
  note_launches(1);
  return DPB200_OK;
}

template <typename FP>
int launch_res_ln_grad(FP* dz, const FP* dout, const FP* zhat, const FP* rstd, const FP* gamma, long long rows, int C,
                       cudaStream_t st) {
  DPB_REQUIRE(rows >= 0 && C >= 1, "attn_residual_layernorm_grad: bad sizes");
  if (rows == 0) return DPB200_OK;
  DPB_REQUIRE(dz && dout && zhat && rstd && gamma, "attn_residual_layernorm_grad: null pointer");
  k_res_ln_grad<FP><<<warp_grid(rows, 8), 256, 0, st>>>(dz, dout, zhat, rstd, gamma, rows, C);
  DPB_CUDA(cudaGetLastError());
  note_launches(1);
  return DPB200_OK;
}

}  // namespace
}  // namespace dpb200

extern "C" {

#define DPB200_DEF_ATTN(SUF, FP)                                                                                    \
  int dpb200_se_atten_embed_##SUF(FP* x0, FP* gs, FP* dgs, const FP* table, const FP* table_info, const FP* em_x,   \
                                  long long em_x_stride, const FP* tt_full, const int* pair, const FP* sw,          \
                                  long long rows, int last_layer_size, dpb200_stream_t stream) {                    \
    return dpb200::launch_embed<FP>(x0, gs, dgs, table, table_info, em_x, em_x_stride, tt_full, pair, sw, rows,     \
                                    last_layer_size, (cudaStream_t)stream);                                         \
  }                                                                                                                 \
  int dpb200_se_atten_embed_grad_##SUF(FP* d_em_x, long long d_em_x_stride, FP* d_sw, const FP* dx0, const FP* gs,  \
                                       const FP* dgs, const FP* tt_full, const int* pair, const FP* sw,             \
                                       long long rows, int last_layer_size, dpb200_stream_t stream) {               \
    return dpb200::launch_embed_grad<FP>(d_em_x, d_em_x_stride, d_sw, dx0, gs, dgs, tt_full, pair, sw, rows,        \
                                         last_layer_size, (cudaStream_t)stream);                                    \
  }                                                                                                                 \
  int dpb200_se_atten_rhat_##SUF(FP* rhat, FP* rinv, const FP* em, long long rows, dpb200_stream_t stream) {        \
    return dpb200::launch_rhat<FP>(rhat, rinv, em, rows, (cudaStream_t)stream);                                     \
  }                                                                                                                 \
  int dpb200_se_atten_rhat_grad_##SUF(FP* d_em, const FP* d_rhat, const FP* rhat, const FP* rinv, long long rows,   \
                                      dpb200_stream_t stream) {                                                     \
    return dpb200::launch_rhat_grad<FP>(d_em, d_rhat, rhat, rinv, rows, (cudaStream_t)stream);                      \
  }                                                                                                                 \
  int dpb200_attn_qkv_normalize_##SUF(FP* qkv, FP* inv_norm, long long rows, int hidden, double q_scale,            \
                                      int normalize, dpb200_stream_t stream) {                                      \
    return dpb200::launch_qkv_norm<FP, false>(qkv, inv_norm, nullptr, rows, hidden, q_scale, normalize,             \
                                              (cudaStream_t)stream);                                                \
  }                                                                                                                 \
  int dpb200_attn_qkv_normalize_grad_##SUF(FP* d_qkv, const FP* qkv_hat, const FP* inv_norm, long long rows,        \
                                           int hidden, double q_scale, int normalize, dpb200_stream_t stream) {     \
    return dpb200::launch_qkv_norm<FP, true>(d_qkv, const_cast<FP*>(inv_norm), qkv_hat, rows, hidden, q_scale,      \
                                             normalize, (cudaStream_t)stream);                                      \
  }                                                                                                                 \
  int dpb200_attn_weights_##SUF(FP* P, FP* A, const FP* S, const FP* sw, const FP* rhat, long long natoms,          \
                                int nnei, int nnei_full, double shift, int dotr, dpb200_stream_t stream) {          \
    return dpb200::launch_weights<FP>(P, A, S, sw, rhat, natoms, nnei, nnei_full, shift, dotr,                      \
                                      (cudaStream_t)stream);                                                        \
  }                                                                                                                 \
  int dpb200_attn_weights_grad_##SUF(FP* dS, FP* d_sw, FP* d_rhat, const FP* dA, const FP* P, const FP* S,          \
                                     const FP* sw, const FP* rhat, long long natoms, int nnei, double shift,        \
                                     int dotr, dpb200_stream_t stream) {                                            \
    return dpb200::launch_weights_grad<FP>(dS, d_sw, d_rhat, dA, P, S, sw, rhat, natoms, nnei, shift, dotr,         \
                                           (cudaStream_t)stream);                                                   \
  }                                                                                                                 \
  int dpb200_attn_residual_layernorm_##SUF(FP* out, FP* y_zhat, FP* rstd, const FP* x, const FP* gamma,             \
                                           const FP* beta, long long rows, int width, double eps,                   \
                                           dpb200_stream_t stream) {                                                \
    return dpb200::launch_res_ln<FP>(out, y_zhat, rstd, x, gamma, beta, rows, width, eps, (cudaStream_t)stream);    \
  }                                                                                                                 \
  int dpb200_attn_residual_layernorm_grad_##SUF(FP* dz, const FP* dout, const FP* zhat, const FP* rstd,             \
                                                const FP* gamma, long long rows, int width,                         \
                                                dpb200_stream_t stream) {                                           \
    return dpb200::launch_res_ln_grad<FP>(dz, dout, zhat, rstd, gamma, rows, width, (cudaStream_t)stream);          \
  }
DPB200_DEF_ATTN(f64, double)
DPB200_DEF_ATTN(f32, float)
#undef DPB200_DEF_ATTN

}  // extern "C"
