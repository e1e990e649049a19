// prod_env_mat_a for sm_100a: neighbour formatting (type/distance sort, per-type truncation),
// environment matrix, its derivative, rij and avg/std normalisation in ONE kernel.
//
// Semantics follow the reference CPU path, which is what "bit-exact neighbour list" is pinned
// to:
//   format   : source/lib/src/fmt_nlist.cc:98-143   (float rr2 from float-cast coordinates,
//              inclusive float cutoff, order (type, rr2, index), per-type truncation)
//   env-mat  : source/lib/src/env_mat.cc:123-221, include/switcher.h:61-84
//   normalise: source/lib/src/prod_env_mat.cc:104-121
//
// Design (not a port of source/lib/src/gpu/prod_env_mat.cu, which uses one CTA per atom, a CUB
// 64-bit block radix sort through a 16 B/candidate global key buffer, and stride-12 scalar
// stores):
//   * one WARP per centre atom, persistent grid (k x 148 CTAs);
//   * candidates are gathered from a packed {x,y,z,type} float4 array (one 16 B L2 hit per
//     candidate instead of four loads), filtered with the reference's float arithmetic
//     (explicit __fmul_rn/__fadd_rn: no FMA contraction), ballot-compacted into shared memory
//     as 64-bit keys  [type:7 | float_bits(rr2):31 | index:26];
//   * the sort is a rank computation over the SURVIVORS only (~90 of ~214 for water):
//     every lane counts the keys smaller than its own through shared-memory broadcasts,
//     which directly yields the output slot sec[t] + rank_in_type — no key movement at all;
//   * env-mat values are produced per 32-slot chunk aligned to each type section, staged
//     through padded shared memory and written with fully coalesced streaming stores
//     (the 12-wide derivative rows are the bulk of the 19*nnei*F bytes per atom).
#include "common.cuh"

namespace dpb200 {
namespace {

typedef unsigned long long u64;

constexpr int kTypeShift = 57;
constexpr int kIdxBits = 26;
constexpr u64 kIdxMask = (1ull << kIdxBits) - 1;
constexpr int kDvStride = 13;  // 12 derivative components + 1 pad word: conflict-free staging

template <typename FP>
struct EnvParams {
  FP* em;
  FP* em_deriv;
  FP* rij;
  int* nlist;
  const FP* coord;
  const float4* packed;  // {float(x), float(y), float(z), bits(f_type)} per atom
  const int* type;
  const int* ilist;
  const int* numneigh;
  const int* const* firstneigh;
  const int* rows;
  long long row_stride;
  const FP* avg;
  const FP* inv_std;
  const FP* pad;
  int nloc, nall, nnei, ntypes, max_nbor;
  long long nrows;
  float rcut2;
  float rmin, rmax;
  int keys_cap;   // u64 entries per warp
  int nnei_pad;   // ints per warp for the formatted row
  int warp_bytes;
  int sec[DPB200_MAX_TYPES + 1];
};

template <typename FP>
__global__ void k_pack_coord(float4* __restrict__ packed, const FP* __restrict__ coord,
                             const int* __restrict__ ftype, long long n) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 v;
  v.x = (float)coord[3 * i + 0];
  v.y = (float)coord[3 * i + 1];
  v.z = (float)coord[3 * i + 2];
  v.w = __int_as_float(ftype[i]);
  packed[i] = v;
}

// inv_std = 1/std ; pad = (0 - avg)/std  (value of an empty slot, prod_env_mat.cc:106-108)
template <typename FP>
__global__ void k_norm_tables(FP* __restrict__ inv_std, FP* __restrict__ pad,
                              const FP* __restrict__ avg, const FP* __restrict__ std_, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  inv_std[i] = (FP)1. / std_[i];
  pad[i] = ((FP)0. - avg[i]) / std_[i];
}

__device__ __forceinline__ void store4(float* o, float a, float b, float c, float d) {
  __stcs(reinterpret_cast<float4*>(o), make_float4(a, b, c, d));
}
__device__ __forceinline__ void store4(double* o, double a, double b, double c, double d) {
  __stcs(reinterpret_cast<double2*>(o), make_double2(a, b));
  __stcs(reinterpret_cast<double2*>(o) + 1, make_double2(c, d));
}

__device__ __forceinline__ float inv_sqrt(float x) { return rsqrtf(x); }
__device__ __forceinline__ double inv_sqrt(double x) { return rsqrt(x); }

// Rank of up to R*32 keys (lane owns keys kb + r*32 + lane) among all survivors.
template <int R>
__device__ __forceinline__ void rank_and_place(const u64* __restrict__ keys, int nsurv, int kb,
                                               int lane, const int* __restrict__ tstart,
                                               const int* __restrict__ sec, int* __restrict__ slot_j) {
  u64 mine[R];
  int rk[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int k = kb + r * 32 + lane;
    mine[r] = k < nsurv ? keys[k] : 0ull;  // 0 is smaller than nothing: rank stays 0
    rk[r] = 0;
  }
  const ulonglong2* kv = reinterpret_cast<const ulonglong2*>(keys);
  const int npair = (nsurv + 1) >> 1;  // keys[nsurv] holds a ~0 sentinel
#pragma unroll 2
  for (int m = 0; m < npair; ++m) {
    const ulonglong2 kk = kv[m];  // same address in all lanes: one broadcast wavefront
#pragma unroll
    for (int r = 0; r < R; ++r) {
      rk[r] += (kk.x < mine[r]) ? 1 : 0;
      rk[r] += (kk.y < mine[r]) ? 1 : 0;
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int k = kb + r * 32 + lane;
    if (k < nsurv) {
      const int t = (int)(mine[r] >> kTypeShift);
      const int rin = rk[r] - tstart[t];
      const int s0 = sec[t];
      if (rin < sec[t + 1] - s0) slot_j[s0 + rin] = (int)(mine[r] & kIdxMask);
    }
  }
}

template <typename FP, bool kEnv>
// (resident CTAs per SM: measured on B200 at 332 k atoms -- fp64 3 / 4 / 5 / 6 CTAs: 2.37 / 2.09 / 2.18 / 2.99 ms (the
//  fifth CTA costs 160 bytes of spills per thread), fp32 4 / 5: 1.71 / 1.63 ms)
__global__ void __launch_bounds__(128, (sizeof(FP) == 8 ? 4 : 5)) k_env_mat_a(const __grid_constant__ EnvParams<FP> p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  unsigned char* base = smem_raw + (size_t)warp * p.warp_bytes;
  u64* keys = reinterpret_cast<u64*>(base);
  FP* dv_s = reinterpret_cast<FP*>(keys + p.keys_cap);
  FP* rij_s = dv_s + 32 * kDvStride;
  int* slot_j = reinterpret_cast<int*>(rij_s + 96);
  int* tstart = slot_j + p.nnei_pad;
  const int nnei = p.nnei;
  const int ntypes = p.ntypes;
  const unsigned lt_mask = (1u << lane) - 1u;

  for (long long row = (long long)blockIdx.x * wpb + warp; row < p.nrows;
       row += (long long)gridDim.x * wpb) {
    const long long frame = row / p.nloc;
    const int i_loc = p.ilist ? p.ilist[row] : (int)(row - frame * p.nloc);
    const long long out_row = frame * p.nloc + i_loc;
    const long long fbase = frame * p.nall;
    int n = p.numneigh[row];
    n = n < p.max_nbor ? n : p.max_nbor;
    const int* __restrict__ rowp = p.firstneigh ? p.firstneigh[row] : p.rows + row * p.row_stride;
    const float4 ci = p.packed[fbase + i_loc];

    // ---- A: gather, float cutoff test, ballot-compact the survivors' keys ----------------
    for (int s = lane; s < nnei; s += 32) slot_j[s] = -1;
    int nsurv = 0;
    // 8 x 32 candidates per pass: all index loads first, then all coordinate gathers, so that one
    // pass costs two memory round trips instead of sixteen
    for (int b0 = 0; b0 < n; b0 += 256) {
      int jv[8];
      float4 cv[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int k = b0 + 32 * q + lane;
        jv[q] = k < n ? rowp[k] : -1;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (jv[q] >= 0) cv[q] = p.packed[fbase + jv[q]];
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (b0 + 32 * q >= n) break;  // warp-uniform
        bool ok = false;
        u64 key = 0;
        if (jv[q] >= 0) {
          const int j = jv[q];
          const float4 cj = cv[q];
          const int tj = __float_as_int(cj.w);
          // (float)rj - (float)ri, then dx*dx + dy*dy + dz*dz left to right, no contraction
          const float dx = __fsub_rn(cj.x, ci.x);
          const float dy = __fsub_rn(cj.y, ci.y);
          const float dz = __fsub_rn(cj.z, ci.z);
          const float rr2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
          ok = (tj >= 0) && (tj < ntypes) && (rr2 <= p.rcut2);  // a type outside sec[] has no section: skipped
          key = ((u64)(unsigned)tj << kTypeShift) | ((u64)__float_as_uint(rr2) << kIdxBits) | (u64)(unsigned)j;
        }
        const unsigned m = __ballot_sync(kFull, ok);
        if (ok) keys[nsurv + __popc(m & lt_mask)] = key;
        nsurv += __popc(m);
      }
    }
    if (lane < 2) keys[nsurv + lane] = ~0ull;  // sentinels for the paired rank loop
    __syncwarp();

    // ---- B: first rank of every type among the survivors --------------------------------
    for (int t = 0; t <= ntypes; ++t) {
      int c = 0;
      for (int b = 0; b < nsurv; b += 32) {
        const int k = b + lane;
        const bool lower = (k < nsurv) && ((int)(keys[k] >> kTypeShift) < t);
        c += __popc(__ballot_sync(kFull, lower));
      }
      if (lane == 0) tstart[t] = c;
    }
    __syncwarp();

    // ---- C: rank => slot (sec[t] + rank within type), truncated at sel[t] ---------------
    for (int kb = 0; kb < nsurv; kb += 128) {
      const int left = nsurv - kb;
      if (left > 96) {
        rank_and_place<4>(keys, nsurv, kb, lane, tstart, p.sec, slot_j);
      } else if (left > 64) {
        rank_and_place<3>(keys, nsurv, kb, lane, tstart, p.sec, slot_j);
      } else if (left > 32) {
        rank_and_place<2>(keys, nsurv, kb, lane, tstart, p.sec, slot_j);
      } else {
        rank_and_place<1>(keys, nsurv, kb, lane, tstart, p.sec, slot_j);
      }
    }
    __syncwarp();

    int* __restrict__ nl_row = p.nlist + out_row * nnei;
    for (int s = lane; s < nnei; s += 32) st_cs(nl_row + s, slot_j[s]);
    if (!kEnv) {
      __syncwarp();
      continue;
    }

    // ---- D: environment matrix, section by section in 32-slot chunks ---------------------
    const int ti = p.type[fbase + i_loc];
    const bool live = ti >= 0;  // a virtual centre atom yields zero em / em_deriv
    const FP* __restrict__ cf = p.coord + fbase * 3;
    const FP xi = cf[3 * (long long)i_loc + 0];
    const FP yi = cf[3 * (long long)i_loc + 1];
    const FP zi = cf[3 * (long long)i_loc + 2];
    const long long tb = (long long)(live ? ti : 0) * nnei * 4;
    const FP* __restrict__ avg_t = p.avg + tb;
    const FP* __restrict__ istd_t = p.inv_std + tb;
    const FP* __restrict__ pad_t = p.pad + tb;
    FP* __restrict__ em_row = p.em + out_row * nnei * 4;
    FP* __restrict__ dv_row = p.em_deriv + out_row * nnei * 12;
    FP* __restrict__ rij_row = p.rij + out_row * nnei * 3;
    const FP rmin = (FP)p.rmin;
    const FP span = (FP)(p.rmax - p.rmin);  // float difference, then promoted (switcher.h:70)
    const FP du = (FP)1. / span;

    for (int t = 0; t < ntypes; ++t) {
      const int s0 = p.sec[t];
      const int s1 = p.sec[t + 1];
      int nreal = tstart[t + 1] - tstart[t];
      nreal = nreal < s1 - s0 ? nreal : s1 - s0;
      for (int c0 = s0; c0 < s1; c0 += 32) {
        const int slot = c0 + lane;
        const bool inb = slot < s1;
        const int nvalid = (s1 - c0) < 32 ? (s1 - c0) : 32;
        if (c0 >= s0 + nreal) {
          // chunk of empty slots only: em = -avg/std, everything else zero
          if (inb) {
            typename Vec4<FP>::type v;
            if (live) {
              v = *reinterpret_cast<const typename Vec4<FP>::type*>(pad_t + 4 * slot);
            } else {
              v.x = v.y = v.z = v.w = (FP)0.;
            }
            store4(em_row + 4 * (long long)slot, v.x, v.y, v.z, v.w);
          }
          for (int e = lane; e < nvalid * 12; e += 32) st_cs(dv_row + (long long)c0 * 12 + e, (FP)0.);
          for (int e = lane; e < nvalid * 3; e += 32) st_cs(rij_row + (long long)c0 * 3 + e, (FP)0.);
          continue;
        }
        const int j = inb ? slot_j[slot] : -1;
        FP v0 = 0, v1 = 0, v2 = 0, v3 = 0;
        FP d[12];
#pragma unroll
        for (int c = 0; c < 12; ++c) d[c] = (FP)0.;
        FP rx = 0, ry = 0, rz = 0;
        if (j >= 0) {
          rx = cf[3 * (long long)j + 0] - xi;
          ry = cf[3 * (long long)j + 1] - yi;
          rz = cf[3 * (long long)j + 2] - zi;
          const FP nr2 = rx * rx + ry * ry + rz * rz;
          const FP inr = inv_sqrt(nr2);
          const FP nr = nr2 * inr;
          const FP inr2 = inr * inr;
          const FP inr4 = inr2 * inr2;
          const FP inr3 = inr4 * nr;
          FP sw, dsw;
          if (nr < rmin) {
            sw = (FP)1.;
            dsw = (FP)0.;
          } else if (nr < (FP)p.rmax) {
            const FP uu = (nr - rmin) / span;
            const FP q = (FP)-6. * uu * uu + (FP)15. * uu - (FP)10.;
            const FP u2 = uu * uu;
            const FP u3 = u2 * uu;
            sw = u3 * q + (FP)1.;
            dsw = ((FP)3. * u2 * q + u3 * ((FP)-12. * uu + (FP)15.)) * du;
          } else {
            sw = (FP)0.;
            dsw = (FP)0.;
          }
          const FP a0 = inr;         // 1/r
          const FP a1 = rx * inr2;   // x/r^2
          const FP a2 = ry * inr2;
          const FP a3 = rz * inr2;
          const FP g = dsw * inr;    // dsw / r
          const FP rr[3] = {rx, ry, rz};
          const FP aa[4] = {a0, a1, a2, a3};
          const FP two_inr4 = (FP)2. * inr4;
#pragma unroll
          for (int dd = 0; dd < 3; ++dd) d[dd] = rr[dd] * inr3 * sw - a0 * g * rr[dd];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
#pragma unroll
            for (int dd = 0; dd < 3; ++dd) {
              FP lead = two_inr4 * rr[c] * rr[dd];
              if (c == dd) lead -= inr2;
              d[3 + 3 * c + dd] = lead * sw - aa[1 + c] * g * rr[dd];
            }
          }
          v0 = a0 * sw;
          v1 = a1 * sw;
          v2 = a2 * sw;
          v3 = a3 * sw;
        }
        if (inb) {
          FP e0, e1, e2, e3;
          if (!live) {
            e0 = e1 = e2 = e3 = (FP)0.;
#pragma unroll
            for (int c = 0; c < 12; ++c) d[c] = (FP)0.;
          } else if (j >= 0) {
            const typename Vec4<FP>::type av = *reinterpret_cast<const typename Vec4<FP>::type*>(avg_t + 4 * slot);
            const typename Vec4<FP>::type is = *reinterpret_cast<const typename Vec4<FP>::type*>(istd_t + 4 * slot);
            e0 = (v0 - av.x) * is.x;
            e1 = (v1 - av.y) * is.y;
            e2 = (v2 - av.z) * is.z;
            e3 = (v3 - av.w) * is.w;
#pragma unroll
            for (int dd = 0; dd < 3; ++dd) {
              d[dd] *= is.x;
              d[3 + dd] *= is.y;
              d[6 + dd] *= is.z;
              d[9 + dd] *= is.w;
            }
          } else {
            const typename Vec4<FP>::type pv = *reinterpret_cast<const typename Vec4<FP>::type*>(pad_t + 4 * slot);
            e0 = pv.x;
            e1 = pv.y;
            e2 = pv.z;
            e3 = pv.w;
          }
          store4(em_row + 4 * (long long)slot, e0, e1, e2, e3);
        }
#pragma unroll
        for (int c = 0; c < 12; ++c) dv_s[lane * kDvStride + c] = d[c];
        rij_s[lane * 3 + 0] = rx;
        rij_s[lane * 3 + 1] = ry;
        rij_s[lane * 3 + 2] = rz;
        __syncwarp();
        FP* dvo = dv_row + (long long)c0 * 12;
#pragma unroll
        for (int q = 0; q < 12; ++q) {
          const int e = q * 32 + lane;
          if (e < nvalid * 12) {
            const int sl = e / 12;
            st_cs(dvo + e, dv_s[sl * kDvStride + (e - sl * 12)]);
          }
        }
        FP* ro = rij_row + (long long)c0 * 3;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const int e = q * 32 + lane;
          if (e < nvalid * 3) st_cs(ro + e, rij_s[e]);
        }
        __syncwarp();
      }
    }
  }
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct EnvWorkspace {
  size_t packed_off, istd_off, pad_off, total;
};

EnvWorkspace env_workspace(int ntypes, int nnei, long long natoms_all, int fp_bytes) {
  EnvWorkspace w;
  w.packed_off = 0;
  size_t cur = align_up((size_t)natoms_all * sizeof(float4), 256);
  w.istd_off = cur;
  cur = align_up(cur + (size_t)ntypes * nnei * 4 * fp_bytes, 256);
  w.pad_off = cur;
  cur = align_up(cur + (size_t)ntypes * nnei * 4 * fp_bytes, 256);
  w.total = cur;
  return w;
}

template <typename FP, bool kEnv>
int launch_env(FP* em, FP* em_deriv, FP* rij, int* nlist, const FP* coord, const int* type,
               const int* f_type, const int* ilist, const int* numneigh,
               const int* const* firstneigh, const int* rows, int row_stride, int max_nbor_size,
               const FP* avg, const FP* std_, int nloc, int nall, int nframes, float rcut,
               float rcut_smth, const int* sec, int nsec, void* workspace, size_t workspace_bytes,
               cudaStream_t stream, int ntypes_center = 0) {
  DPB_REQUIRE(nsec >= 2 && nsec - 1 <= DPB200_MAX_TYPES, "prod_env_mat_a: 1 <= ntypes <= 128 required");
  DPB_REQUIRE(nloc >= 0 && nall >= nloc && nframes >= 1, "prod_env_mat_a: need nall >= nloc >= 0, nframes >= 1");
  DPB_REQUIRE(nall <= DPB200_MAX_NALL, "prod_env_mat_a: nall exceeds 2^26 (index bits of the sort key)");
  DPB_REQUIRE(max_nbor_size >= 0 && max_nbor_size <= DPB200_MAX_NBOR_SIZE,
              "prod_env_mat_a: neighbour rows wider than 4096 are not supported");
  for (int t = 0; t + 1 < nsec; ++t) DPB_REQUIRE(sec[t + 1] >= sec[t] && sec[0] == 0, "prod_env_mat_a: sec must be a non-decreasing prefix sum starting at 0");
  const int ntypes = nsec - 1;
  const int nnei = sec[nsec - 1];
  const long long nrows = (long long)nframes * nloc;
  if (nrows == 0 || nnei == 0) return DPB200_OK;
  DPB_REQUIRE(numneigh != nullptr && (firstneigh != nullptr || rows != nullptr), "prod_env_mat_a: neighbour list pointers are null");
  // rows of avg / std = centre-atom types; more than the sections when the list is formatted with a coarser
  // f_type (se_atten: ONE distance-ordered section, statistics per real centre type)
  const int ntc = ntypes_center > 0 ? ntypes_center : ntypes;
  const EnvWorkspace w = env_workspace(ntc, nnei, (long long)nframes * nall, sizeof(FP));
  DPB_REQUIRE(workspace != nullptr && workspace_bytes >= w.total, "prod_env_mat_a: workspace too small (see dpb200_prod_env_mat_a_workspace_bytes)");
  DPB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "prod_env_mat_a: workspace must be 256-byte aligned");

  unsigned char* ws = static_cast<unsigned char*>(workspace);
  float4* packed = reinterpret_cast<float4*>(ws + w.packed_off);
  FP* inv_std = reinterpret_cast<FP*>(ws + w.istd_off);
  FP* pad = reinterpret_cast<FP*>(ws + w.pad_off);
  const long long natoms_all = (long long)nframes * nall;
  k_pack_coord<FP><<<ceil_div(natoms_all, 256), 256, 0, stream>>>(packed, coord, f_type ? f_type : type, natoms_all);
  if (kEnv) {
    DPB_REQUIRE(avg != nullptr && std_ != nullptr, "prod_env_mat_a: avg/std are null");
    DPB_REQUIRE(((reinterpret_cast<uintptr_t>(avg) | reinterpret_cast<uintptr_t>(em)) & 15) == 0,
                "prod_env_mat_a: avg and em must be 16-byte aligned");
    const int ntab = ntc * nnei * 4;
    k_norm_tables<FP><<<ceil_div(ntab, 256), 256, 0, stream>>>(inv_std, pad, avg, std_, ntab);
  }

  EnvParams<FP> p;
  p.em = em;
  p.em_deriv = em_deriv;
  p.rij = rij;
  p.nlist = nlist;
  p.coord = coord;
  p.packed = packed;
  p.type = type;
  p.ilist = ilist;
  p.numneigh = numneigh;
  p.firstneigh = firstneigh;
  p.rows = rows;
  p.row_stride = row_stride;
  p.avg = avg;
  p.inv_std = inv_std;
  p.pad = pad;
  p.nloc = nloc;
  p.nall = nall;
  p.nnei = nnei;
  p.ntypes = ntypes;
  p.max_nbor = max_nbor_size;
  p.nrows = nrows;
  p.rcut2 = rcut * rcut;  // float product, fmt_nlist.cc:112
  p.rmin = rcut_smth;
  p.rmax = rcut;
  p.keys_cap = (int)align_up((size_t)max_nbor_size + 2, 2);
  p.nnei_pad = (int)align_up((size_t)nnei, 4);
  for (int t = 0; t < nsec; ++t) p.sec[t] = sec[t];
  for (int t = nsec; t <= DPB200_MAX_TYPES; ++t) p.sec[t] = nnei;
  size_t wb = (size_t)p.keys_cap * 8 + (size_t)(32 * kDvStride + 96) * sizeof(FP) +
              (size_t)(p.nnei_pad + ntypes + 1) * 4;
  wb = align_up(wb, 16);
  p.warp_bytes = (int)wb;
  const int wpb = 4;
  const size_t smem = wb * wpb;
  DPB_REQUIRE(smem <= 227 * 1024, "prod_env_mat_a: nnei / max_nbor_size too large for shared memory");
  auto kern = k_env_mat_a<FP, kEnv>;
  DPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  DPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, wpb * 32, smem));
  if (occ < 1) occ = 1;
  long long want = (nrows + wpb - 1) / wpb;
  long long cap = (long long)sm_count() * occ;
  int grid = (int)(want < cap ? want : cap);
  kern<<<grid, wpb * 32, smem, stream>>>(p);
  DPB_CUDA(cudaGetLastError());
  note_launches(kEnv ? 3 : 2);
  return DPB200_OK;
}

}  // namespace
}  // namespace dpb200

extern "C" {

size_t dpb200_prod_env_mat_a_workspace_bytes(int ntypes, int nnei, int nall, int nframes, int fp_bytes) {
  return dpb200::env_workspace(ntypes, nnei, (long long)nframes * nall, fp_bytes).total;
}

#define DPB200_DEF_ENV(SUF, FP)                                                                    \
  int dpb200_prod_env_mat_a_##SUF(FP* em, FP* em_deriv, FP* rij, int* nlist, const FP* coord,      \
                                  const int* type, const int* f_type, const int* ilist,            \
                                  const int* numneigh, const int* const* firstneigh,               \
                                  const int* rows, int row_stride, int max_nbor_size,              \
                                  const FP* avg, const FP* std, int nloc, int nall, int nframes,   \
                                  float rcut, float rcut_smth, const int* sec, int nsec,           \
                                  void* workspace, size_t workspace_bytes,                         \
                                  dpb200_stream_t stream) {                                        \
    return dpb200::launch_env<FP, true>(em, em_deriv, rij, nlist, coord, type, f_type, ilist,      \
                                        numneigh, firstneigh, rows, row_stride, max_nbor_size,     \
                                        avg, std, nloc, nall, nframes, rcut, rcut_smth, sec, nsec, \
                                        workspace, workspace_bytes, (cudaStream_t)stream);         \
  }                                                                                                \
  int dpb200_prod_env_mat_a_ex_##SUF(FP* em, FP* em_deriv, FP* rij, int* nlist, const FP* coord,   \
                                     const int* type, const int* f_type, const int* ilist,         \
                                     const int* numneigh, const int* const* firstneigh,            \
                                     const int* rows, int row_stride, int max_nbor_size,           \
                                     const FP* avg, const FP* std, int ntypes_center, int nloc,    \
                                     int nall, int nframes, float rcut, float rcut_smth,           \
                                     const int* sec, int nsec, void* workspace,                    \
                                     size_t workspace_bytes, dpb200_stream_t stream) {             \
    return dpb200::launch_env<FP, true>(em, em_deriv, rij, nlist, coord, type, f_type, ilist,      \
                                        numneigh, firstneigh, rows, row_stride, max_nbor_size,     \
                                        avg, std, nloc, nall, nframes, rcut, rcut_smth, sec, nsec, \
                                        workspace, workspace_bytes, (cudaStream_t)stream,          \
                                        ntypes_center);                                            \
  }                                                                                                \
  int dpb200_format_nlist_##SUF(int* nlist, const FP* coord, const int* type, const int* ilist,    \
                                const int* numneigh, const int* const* firstneigh,                 \
                                const int* rows, int row_stride, int max_nbor_size, int nloc,      \
                                int nall, int nframes, float rcut, const int* sec, int nsec,       \
                                void* workspace, size_t workspace_bytes,                           \
                                dpb200_stream_t stream) {                                          \
    return dpb200::launch_env<FP, false>(nullptr, nullptr, nullptr, nlist, coord, type, nullptr,   \
                                         ilist, numneigh, firstneigh, rows, row_stride,            \
                                         max_nbor_size, nullptr, nullptr, nloc, nall, nframes,     \
                                         rcut, 0.f, sec, nsec, workspace, workspace_bytes,         \
                                         (cudaStream_t)stream);                                    \
  }
DPB200_DEF_ENV(f64, double)
DPB200_DEF_ENV(f32, float)
#undef DPB200_DEF_ENV

}  // extern "C"
