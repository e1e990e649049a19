// Neighbour-list front end for sm_100a: coordinate normalisation, periodic ghost images and a
// cell-list raw neighbour build.
//
// Semantics follow the reference CPU functions
//   normalize_coord_cpu : source/lib/src/coord.cc:13-28  (region.cc:84-124 for the two products)
//   copy_coord_cpu      : source/lib/src/coord.cc:30-66 -> neighbor_list.cc:747-847 (cell-granular
//                         ghost shell, double arithmetic), compute_cell_info coord.cc:68-108
//   build_nlist_cpu     : source/lib/src/neighbor_list.cc:875-928 (strict FPTYPE cutoff, rows in
//                         ascending neighbour index)
// The reference GPU versions (source/lib/src/gpu/coord.cu, neighbor_list.cu) are O(ncell*nloc)
// and O(nloc*nall) with 2*nloc*nall*4 bytes of scratch; nothing of them is reused.  Here:
//   * ghosts: one thread per local atom counts its images from its cell index, one exclusive
//     scan, one thread per atom writes them (order: owner atom, then image shift);
//   * raw list: atoms are binned into cells of edge >= rcut over the bounding box of the extended
//     system and copied in cell order ({x,y,z,index} records, coalesced scans), one WARP per centre
//     atom walks the 27 surrounding cells, ballot-compacts the hits into shared memory and
//     rank-sorts them so that every row comes out in ascending index (deterministic, and equal
//     to build_nlist_cpu element by element).
// All floating-point expressions that decide membership are written with explicit round-to-nearest
// intrinsics so that no FMA contraction can change a decision with respect to the CPU build.
#include <cmath>

#include "common.cuh"

namespace dpb200 {
namespace {

// ---------------------------------------------------------------- small utilities ---------
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }

constexpr int kScanTile = 2048;  // 256 threads x 8

// tile-local exclusive scan; tile totals go to sums[blockIdx.x]
__global__ void __launch_bounds__(256) k_scan_tiles(int* __restrict__ out, const int* __restrict__ in,
                                                    int* __restrict__ sums, long long n) {
  __shared__ int wsum[8];
  const long long base = (long long)blockIdx.x * kScanTile + threadIdx.x * 8;
  int v[8];
  int t = 0;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    v[q] = (base + q < n) ? in[base + q] : 0;
    t += v[q];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = t;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(kFull, inc, o);
    if (lane >= o) inc += u;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  int woff = 0;
  for (int w = 0; w < warp; ++w) woff += wsum[w];
  int run = woff + inc - t;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    if (base + q < n) out[base + q] = run;
    run += v[q];
  }
  if (threadIdx.x == 255) sums[blockIdx.x] = woff + inc;
}

__global__ void __launch_bounds__(256) k_scan_add(int* __restrict__ out, const int* __restrict__ offs, long long n) {
  const int add = offs[blockIdx.x];
  const long long base = (long long)blockIdx.x * kScanTile + threadIdx.x * 8;
#pragma unroll
  for (int q = 0; q < 8; ++q)
    if (base + q < n) out[base + q] += add;
}

size_t scan_tmp_ints(long long n) {
  size_t tot = 0;
  while (n > 1) {
    n = (n + kScanTile - 1) / kScanTile;
    tot += (size_t)n + 1;
  }
  return tot + 2;
}

// out[0..n) = exclusive prefix sum of in[0..n); out[n] = total when with_total (out has n+1 slots:
// the caller passes in[n] == 0 and n+1 as length instead).
int exclusive_scan(int* out, const int* in, long long n, int* tmp, cudaStream_t st) {
  if (n <= 0) return DPB200_OK;
  const long long nt = (n + kScanTile - 1) / kScanTile;
  k_scan_tiles<<<(unsigned)nt, 256, 0, st>>>(out, in, tmp, n);
  note_launches(nt > 1 ? 2 : 1);
  if (nt > 1) {
    int* next = tmp + nt;
    int rc = exclusive_scan(tmp, tmp, nt, next, st);
    if (rc) return rc;
    k_scan_add<<<(unsigned)nt, 256, 0, st>>>(out, tmp, n);
  }
  DPB_CUDA(cudaGetLastError());
  return DPB200_OK;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---------------------------------------------------------------- normalize_coord ---------
template <typename FP>
struct Box {
  FP b[9];    // rows = cell vectors
  FP rec[9];  // reciprocal (region.cc:50-81)
};

template <typename FP>
void make_rec(FP* rec, const FP* b) {  // host; region.cc:50-81 operation order
  FP vol = b[0] * (b[4] * b[8] - b[7] * b[5]) - b[1] * (b[3] * b[8] - b[6] * b[5]) + b[2] * (b[3] * b[7] - b[6] * b[4]);
  vol = vol < 0 ? -vol : vol;
  const FP vi = (FP)1. / vol;
  rec[0] = (b[4] * b[8] - b[7] * b[5]) * vi;
  rec[4] = (b[0] * b[8] - b[6] * b[2]) * vi;
  rec[8] = (b[0] * b[4] - b[3] * b[1]) * vi;
  rec[1] = (-b[3] * b[8] + b[6] * b[5]) * vi;
  rec[2] = (b[3] * b[7] - b[6] * b[4]) * vi;
  rec[3] = (-b[1] * b[8] + b[7] * b[2]) * vi;
  rec[5] = (-b[0] * b[7] + b[6] * b[1]) * vi;
  rec[6] = (b[1] * b[5] - b[4] * b[2]) * vi;
  rec[7] = (-b[0] * b[5] + b[3] * b[2]) * vi;
}

template <typename FP>
__global__ void k_normalize(FP* __restrict__ coord, int natom, const __grid_constant__ Box<FP> bx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= natom) return;
  FP* r = coord + 3 * (long long)i;
  const FP r0 = r[0], r1 = r[1], r2 = r[2];
  FP s[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    FP v = add_rn(add_rn(mul_rn(r0, bx.rec[3 * d]), mul_rn(r1, bx.rec[3 * d + 1])), mul_rn(r2, bx.rec[3 * d + 2]));
    v = fmod(v, (FP)1.);
    if (v < (FP)0.) v = add_rn(v, (FP)1.);
    s[d] = v;
  }
#pragma unroll
  for (int d = 0; d < 3; ++d)
    r[d] = add_rn(add_rn(mul_rn(s[0], bx.b[d]), mul_rn(s[1], bx.b[3 + d])), mul_rn(s[2], bx.b[6 + d]));
}

// ---------------------------------------------------------------- copy_coord ---------------
struct CellInfo {
  double box[9], rec[9];
  int ncell[3], ng[3];
};

// host; SimulationRegion_Impl.h:364-372,427-484 + coord.cc:68-108
void make_cell_info(CellInfo& ci, const double* b, float rcut) {
  for (int k = 0; k < 9; ++k) ci.box[k] = b[k];
  double vol = b[0] * (b[4] * b[8] - b[7] * b[5]) - b[1] * (b[3] * b[8] - b[6] * b[5]) + b[2] * (b[3] * b[7] - b[6] * b[4]);
  vol = std::fabs(vol);
  const double vi = 1. / vol;
  double* rec = ci.rec;
  rec[0] = (b[4] * b[8] - b[7] * b[5]) * vi;
  rec[4] = (b[0] * b[8] - b[6] * b[2]) * vi;
  rec[8] = (b[0] * b[4] - b[3] * b[1]) * vi;
  rec[1] = (-b[3] * b[8] + b[6] * b[5]) * vi;
  rec[2] = (b[3] * b[7] - b[6] * b[4]) * vi;
  rec[3] = (-b[1] * b[8] + b[7] * b[2]) * vi;
  rec[5] = (-b[0] * b[7] + b[6] * b[1]) * vi;
  rec[6] = (b[1] * b[5] - b[4] * b[2]) * vi;
  rec[7] = (-b[0] * b[5] + b[3] * b[2]) * vi;
  const double* r[3] = {b, b + 3, b + 6};
  const double rc = rcut;
  for (int d = 0; d < 3; ++d) {
    const double* u = r[(d + 1) % 3];
    const double* v = r[(d + 2) % 3];
    const double c[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
    const double face = vol * (1. / std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]));
    ci.ncell[d] = (int)(face / rc);
    if (ci.ncell[d] == 0) ci.ncell[d] = 1;
    const double cs = face / ci.ncell[d];
    ci.ng[d] = (int)(rc / cs) + 1;
  }
}

__device__ __forceinline__ int floor_div(int a, int b) {  // b > 0
  int q = a / b;
  if ((a % b != 0) && (a < 0)) --q;
  return q;
}

// cell of a local atom along d (build_clist, neighbor_list.cc:43-148) and the range of image
// shifts s for which the extended cell c - s*ncell lies in [-ng, ncell+ng).
template <typename FP>
__device__ __forceinline__ void image_ranges(const CellInfo& ci, const FP* __restrict__ r, int (&smin)[3], int (&smax)[3]) {
  const double p0 = (double)r[0], p1 = (double)r[1], p2 = (double)r[2];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const double s = __dadd_rn(__dadd_rn(__dmul_rn(p0, ci.rec[3 * d]), __dmul_rn(p1, ci.rec[3 * d + 1])), __dmul_rn(p2, ci.rec[3 * d + 2]));
    const int nc = ci.ncell[d];
    const double cs = 1. / nc;
    int c = (int)(s / cs);
    if (s < 0.) c--;
    if (c < 0) c = 0;
    if (c >= nc) c = nc - 1;
    const int g = ci.ng[d];
    smax[d] = floor_div(c + g, nc);
    smin[d] = floor_div(c - nc - g, nc) + 1;
  }
}

template <typename FP>
__global__ void k_ghost_count(int* __restrict__ cnt, const FP* __restrict__ coord, int nloc,
                              const __grid_constant__ CellInfo ci) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > nloc) return;
  if (i == nloc) {
    cnt[i] = 0;
    return;
  }
  int smin[3], smax[3];
  image_ranges(ci, coord + 3 * (long long)i, smin, smax);
  cnt[i] = (smax[0] - smin[0] + 1) * (smax[1] - smin[1] + 1) * (smax[2] - smin[2] + 1) - 1;
}

template <typename FP>
__global__ void k_ghost_fill(FP* __restrict__ out_c, int* __restrict__ out_t, int* __restrict__ mapping,
                             const int* __restrict__ off, const FP* __restrict__ coord,
                             const int* __restrict__ type, int nloc, const __grid_constant__ CellInfo ci) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nloc) return;
  const FP* r = coord + 3 * (long long)i;
  const int t = type[i];
  out_c[3 * (long long)i + 0] = r[0];
  out_c[3 * (long long)i + 1] = r[1];
  out_c[3 * (long long)i + 2] = r[2];
  out_t[i] = t;
  mapping[i] = i;
  int smin[3], smax[3];
  image_ranges(ci, r, smin, smax);
  long long w = (long long)nloc + off[i];
  // descending shift = ascending extended cell, the x-major walk of the reference
  for (int a = smax[0]; a >= smin[0]; --a)
    for (int b = smax[1]; b >= smin[1]; --b)
      for (int c = smax[2]; c >= smin[2]; --c) {
        if (a == 0 && b == 0 && c == 0) continue;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const double sv = __dadd_rn(__dadd_rn(__dmul_rn((double)a, ci.box[d]), __dmul_rn((double)b, ci.box[3 + d])), __dmul_rn((double)c, ci.box[6 + d]));
          out_c[3 * w + d] = (FP)__dsub_rn((double)r[d], sv);
        }
        out_t[w] = t;
        mapping[w] = i;
        ++w;
      }
}

// ---------------------------------------------------------------- build_nlist --------------
struct GridDev {  // filled on the device (bounding box), read by the later kernels
  double lo[3], inv[3];
  int n[3];
  int ncell;
};

__device__ __forceinline__ unsigned long long ord(double v) {
  const unsigned long long u = (unsigned long long)__double_as_longlong(v);
  return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double unord(unsigned long long u) {
  u = (u & 0x8000000000000000ull) ? (u & 0x7fffffffffffffffull) : ~u;
  return __longlong_as_double((long long)u);
}

__global__ void k_bbox_init(unsigned long long* mm) {
  if (threadIdx.x < 3) mm[threadIdx.x] = ~0ull;       // min
  else if (threadIdx.x < 6) mm[threadIdx.x] = 0ull;   // max
}

template <typename FP>
__global__ void __launch_bounds__(256) k_bbox(unsigned long long* __restrict__ mm, const FP* __restrict__ coord, long long nall) {
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nall; i += (long long)gridDim.x * blockDim.x) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const double v = (double)coord[3 * i + d];
      lo[d] = fmin(lo[d], v);
      hi[d] = fmax(hi[d], v);
    }
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[d] = fmin(lo[d], __shfl_xor_sync(kFull, lo[d], o));
      hi[d] = fmax(hi[d], __shfl_xor_sync(kFull, hi[d], o));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      atomicMin(mm + d, ord(lo[d]));
      atomicMax(mm + 3 + d, ord(hi[d]));
    }
  }
}

// cell edge >= rcut (so that the 27 surrounding cells cover the cutoff sphere); total cell count
// bounded by max_cells.
__global__ void k_grid_setup(GridDev* g, const unsigned long long* mm, double rcut, int max_cells) {
  if (threadIdx.x != 0) return;
  double ext[3];
  for (int d = 0; d < 3; ++d) {
    g->lo[d] = unord(mm[d]);
    ext[d] = unord(mm[3 + d]) - g->lo[d];
    if (!(ext[d] >= 0.)) ext[d] = 0.;
  }
  double cs = rcut * 1.0000001 + 1e-12;
  for (int iter = 0; iter < 64; ++iter) {
    long long tot = 1;
    for (int d = 0; d < 3; ++d) {
      int n = (int)(ext[d] / cs) + 1;
      g->n[d] = n;
      tot *= n;
    }
    if (tot <= max_cells) break;
    cs *= 1.26;
  }
  for (int d = 0; d < 3; ++d) {
    // width w = (ext + eps)/n >= cs*(n-1)/n ... enforce w >= rcut explicitly
    double w = (ext[d] + 1e-9 + cs * 1e-6) / g->n[d];
    while (w < rcut * 1.0000001 && g->n[d] > 1) {
      g->n[d] -= 1;
      w = (ext[d] + 1e-9 + cs * 1e-6) / g->n[d];
    }
    g->inv[d] = 1.0 / w;
  }
  g->ncell = g->n[0] * g->n[1] * g->n[2];
}

template <typename FP>
__device__ __forceinline__ int cell_of(const GridDev& g, const FP* __restrict__ r, int (&c)[3]) {
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    int v = (int)(((double)r[d] - g.lo[d]) * g.inv[d]);
    v = v < 0 ? 0 : v;
    v = v >= g.n[d] ? g.n[d] - 1 : v;
    c[d] = v;
  }
  return (c[0] * g.n[1] + c[1]) * g.n[2] + c[2];
}

template <typename FP>
__global__ void k_cell_count(int* __restrict__ cell_id, int* __restrict__ count, const FP* __restrict__ coord,
                             long long nall, const GridDev* __restrict__ g) {
  const GridDev gg = *g;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nall; i += (long long)gridDim.x * blockDim.x) {
    int c[3];
    const int id = cell_of(gg, coord + 3 * i, c);
    cell_id[i] = id;
    atomicAdd(count + id, 1);
  }
}

template <typename FP>
struct alignas(16) SortedAtom {
  FP x, y, z;
  int idx;
  int type;
};

template <typename FP>
__global__ void k_cell_fill(SortedAtom<FP>* __restrict__ sorted, int* __restrict__ cursor,
                            const int* __restrict__ cell_id, const FP* __restrict__ coord,
                            const int* __restrict__ type, long long nall) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nall; i += (long long)gridDim.x * blockDim.x) {
    const int pos = atomicAdd(cursor + cell_id[i], 1);
    SortedAtom<FP> a;
    a.x = coord[3 * i], a.y = coord[3 * i + 1], a.z = coord[3 * i + 2];
    a.idx = (int)i;
    a.type = type ? type[i] : 0;
    sorted[pos] = a;
  }
}

template <typename FP>
struct BuildParams {
  int* numneigh;
  int* rows;
  int* max_list;  // device scalar
  const FP* coord;
  const int* type;
  const SortedAtom<FP>* sorted;
  const int* cell_start;  // [ncell+1] exclusive scan of counts
  const GridDev* grid;
  int nloc, mem_size, cap;
  FP rcut2;
};

template <typename FP>
__global__ void __launch_bounds__(128) k_build_rows(const __grid_constant__ BuildParams<FP> p) {
  extern __shared__ int cand_all[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int* cand = cand_all + (size_t)warp * p.cap;
  const GridDev g = *p.grid;
  const unsigned lt = (1u << lane) - 1u;
  int wmax = 0;
  for (int i = blockIdx.x * 4 + warp; i < p.nloc; i += gridDim.x * 4) {
    const FP xi = p.coord[3 * (long long)i], yi = p.coord[3 * (long long)i + 1], zi = p.coord[3 * (long long)i + 2];
    int cnt = 0;
    const bool live = !(p.type && p.type[i] < 0);
    if (live) {
      const FP ri[3] = {xi, yi, zi};
      int c[3];
      cell_of(g, ri, c);
      for (int a = c[0] - 1; a <= c[0] + 1; ++a) {
        if (a < 0 || a >= g.n[0]) continue;
        for (int b = c[1] - 1; b <= c[1] + 1; ++b) {
          if (b < 0 || b >= g.n[1]) continue;
          // the three z-cells are contiguous in the cell-sorted array
          const int z0 = c[2] - 1 < 0 ? 0 : c[2] - 1;
          const int z1 = c[2] + 1 >= g.n[2] ? g.n[2] - 1 : c[2] + 1;
          const int base = (a * g.n[1] + b) * g.n[2];
          const int m0 = p.cell_start[base + z0], m1 = p.cell_start[base + z1 + 1];
          for (int m = m0; m < m1; m += 32) {
            bool ok = false;
            int j = -1;
            if (m + lane < m1) {
              const SortedAtom<FP> s = p.sorted[m + lane];
              j = s.idx;
              const FP dx = sub_rn(xi, s.x), dy = sub_rn(yi, s.y), dz = sub_rn(zi, s.z);
              const FP d2 = add_rn(add_rn(mul_rn(dx, dx), mul_rn(dy, dy)), mul_rn(dz, dz));
              ok = (j != i) && (s.type >= 0) && (d2 < p.rcut2);
            }
            const unsigned mk = __ballot_sync(kFull, ok);
            if (ok) {
              const int w = cnt + __popc(mk & lt);
              if (w < p.cap) cand[w] = j;
            }
            cnt += __popc(mk);
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) p.numneigh[i] = cnt;
    wmax = cnt > wmax ? cnt : wmax;
    if (cnt <= p.mem_size && cnt <= p.cap) {
      // rank sort: indices are distinct, rank = number of smaller ones
      int* __restrict__ row = p.rows + (long long)i * p.mem_size;
      for (int k0 = 0; k0 < cnt; k0 += 128) {
        int mine[4], rk[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int k = k0 + 32 * r + lane;
          mine[r] = k < cnt ? cand[k] : -1;
          rk[r] = 0;
        }
        for (int m = 0; m < cnt; ++m) {
          const int v = cand[m];
#pragma unroll
          for (int r = 0; r < 4; ++r) rk[r] += (v < mine[r]) ? 1 : 0;
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
          if (mine[r] >= 0) row[rk[r]] = mine[r];
      }
    }
    __syncwarp();
  }
  wmax = max(wmax, __shfl_xor_sync(kFull, wmax, 16));
  if (lane == 0 && wmax > 0) atomicMax(p.max_list, wmax);
}

struct BuildWs {
  size_t mm, grid, maxl, cell_id, count, start, sorted, scan_tmp, total;
  int max_cells;
};

BuildWs build_ws(long long nall, int fp_bytes) {
  BuildWs w;
  // cells: at most ~ one per 2 atoms, at least 4096
  long long mc = nall / 2 + 4096;
  if (mc > (1 << 24)) mc = 1 << 24;
  w.max_cells = (int)mc;
  size_t cur = 0;
  w.mm = cur;
  cur = align_up(cur + 6 * 8, 256);
  w.grid = cur;
  cur = align_up(cur + sizeof(GridDev), 256);
  w.maxl = cur;
  cur = align_up(cur + 16, 256);
  w.cell_id = cur;
  cur = align_up(cur + (size_t)nall * 4, 256);
  w.count = cur;
  cur = align_up(cur + ((size_t)mc + 2) * 4, 256);
  w.start = cur;
  cur = align_up(cur + ((size_t)mc + 2) * 4, 256);
  w.sorted = cur;
  cur = align_up(cur + (size_t)nall * (fp_bytes == 8 ? 32 : 32), 256);
  w.scan_tmp = cur;
  cur = align_up(cur + scan_tmp_ints(mc + 2) * 4, 256);
  w.total = cur;
  return w;
}

int grid_1d(long long n, int threads, int per_sm) {
  long long want = (n + threads - 1) / threads;
  long long cap = (long long)sm_count() * per_sm;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

template <typename FP>
int do_normalize(FP* coord, int natom, const FP* boxt, cudaStream_t st) {
  DPB_REQUIRE(natom >= 0, "normalize_coord: negative natom");
  if (natom == 0) return DPB200_OK;
  DPB_REQUIRE(coord && boxt, "normalize_coord: null pointer");
  Box<FP> bx;
  for (int k = 0; k < 9; ++k) bx.b[k] = boxt[k];
  make_rec(bx.rec, boxt);
  k_normalize<FP><<<ceil_div(natom, 256), 256, 0, st>>>(coord, natom, bx);
  DPB_CUDA(cudaGetLastError());
  note_launches(1);
  return DPB200_OK;
}

template <typename FP>
int do_copy_coord(FP* out_c, int* out_t, int* mapping, int* nall_out, const FP* in_c, const int* in_t,
                  int nloc, int mem_nall, float rcut, const int* ncell, const int* ngcell, const FP* boxt,
                  void* workspace, size_t ws_bytes, cudaStream_t st) {
  DPB_REQUIRE(nloc >= 0 && mem_nall >= 0 && nall_out, "copy_coord: bad size arguments");
  DPB_REQUIRE((rcut > 0.f || (ncell && ngcell)) && boxt, "copy_coord: rcut must be positive and boxt non-null");
  if (nloc == 0) {
    *nall_out = 0;
    return DPB200_OK;
  }
  DPB_REQUIRE(in_c && in_t && out_c && out_t && mapping, "copy_coord: null pointer");
  DPB_REQUIRE(workspace && ws_bytes >= dpb200_copy_coord_workspace_bytes(nloc), "copy_coord: workspace too small");
  CellInfo ci;
  double b[9];
  for (int k = 0; k < 9; ++k) b[k] = (double)boxt[k];
  make_cell_info(ci, b, rcut > 0.f ? rcut : 1.f);
  if (ncell && ngcell) {  // cell grid handed over by the caller (compute_cell_info, coord.cc:68-108)
    for (int d = 0; d < 3; ++d) {
      DPB_REQUIRE(ncell[d] >= 1 && ngcell[d] >= 0, "copy_coord: invalid cell grid");
      ci.ncell[d] = ncell[d];
      ci.ng[d] = ngcell[d];
    }
  }
  int* cnt = static_cast<int*>(workspace);
  int* off = cnt + align_up((size_t)nloc + 1, 64);
  int* tmp = off + align_up((size_t)nloc + 1, 64);
  k_ghost_count<FP><<<ceil_div((long long)nloc + 1, 256), 256, 0, st>>>(cnt, in_c, nloc, ci);
  note_launches(1);
  int rc = exclusive_scan(off, cnt, (long long)nloc + 1, tmp, st);
  if (rc) return rc;
  int total = 0;
  DPB_CUDA(cudaMemcpyAsync(&total, off + nloc, sizeof(int), cudaMemcpyDeviceToHost, st));
  DPB_CUDA(cudaStreamSynchronize(st));
  const long long nall = (long long)nloc + total;
  DPB_REQUIRE(nall <= 0x7fffffffll, "copy_coord: nall overflows int");
  *nall_out = (int)nall;
  if (nall > mem_nall) return 1;
  k_ghost_fill<FP><<<ceil_div(nloc, 128), 128, 0, st>>>(out_c, out_t, mapping, off, in_c, in_t, nloc, ci);
  DPB_CUDA(cudaGetLastError());
  note_launches(1);
  return DPB200_OK;
}

template <typename FP>
int do_build_nlist(int* numneigh, int* rows, int* max_list_size, const FP* coord, int nloc, int nall,
                   int mem_size, float rcut, const int* type, void* workspace, size_t ws_bytes,
                   cudaStream_t st) {
  DPB_REQUIRE(nloc >= 0 && nall >= nloc && mem_size >= 0 && max_list_size, "build_nlist: bad size arguments");
  *max_list_size = 0;
  if (nloc == 0) return DPB200_OK;
  DPB_REQUIRE(rcut > 0.f, "build_nlist: rcut must be positive");
  DPB_REQUIRE(mem_size <= 8192, "build_nlist: mem_size (row capacity) above 8192 is not supported");
  DPB_REQUIRE(numneigh && rows && coord, "build_nlist: null pointer");
  const BuildWs w = build_ws(nall, sizeof(FP));
  DPB_REQUIRE(workspace && ws_bytes >= w.total, "build_nlist: workspace too small");
  DPB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "build_nlist: workspace must be 256-byte aligned");
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  unsigned long long* mm = reinterpret_cast<unsigned long long*>(ws + w.mm);
  GridDev* grid = reinterpret_cast<GridDev*>(ws + w.grid);
  int* maxl = reinterpret_cast<int*>(ws + w.maxl);
  int* cell_id = reinterpret_cast<int*>(ws + w.cell_id);
  int* count = reinterpret_cast<int*>(ws + w.count);
  int* start = reinterpret_cast<int*>(ws + w.start);
  SortedAtom<FP>* sorted = reinterpret_cast<SortedAtom<FP>*>(ws + w.sorted);
  int* scan_tmp = reinterpret_cast<int*>(ws + w.scan_tmp);

  k_bbox_init<<<1, 32, 0, st>>>(mm);
  k_bbox<FP><<<grid_1d(nall, 256, 8), 256, 0, st>>>(mm, coord, nall);
  k_grid_setup<<<1, 32, 0, st>>>(grid, mm, (double)rcut, w.max_cells);
  DPB_CUDA(cudaMemsetAsync(count, 0, ((size_t)w.max_cells + 2) * 4, st));
  DPB_CUDA(cudaMemsetAsync(maxl, 0, 4, st));
  k_cell_count<FP><<<grid_1d(nall, 256, 8), 256, 0, st>>>(cell_id, count, coord, nall, grid);
  int rc = exclusive_scan(start, count, (long long)w.max_cells + 1, scan_tmp, st);
  if (rc) return rc;
  // cursor = copy of start (count buffer is reused)
  DPB_CUDA(cudaMemcpyAsync(count, start, ((size_t)w.max_cells + 1) * 4, cudaMemcpyDeviceToDevice, st));
  k_cell_fill<FP><<<grid_1d(nall, 256, 8), 256, 0, st>>>(sorted, count, cell_id, coord, type, nall);

  BuildParams<FP> p;
  p.numneigh = numneigh;
  p.rows = rows;
  p.max_list = maxl;
  p.coord = coord;
  p.type = type;
  p.sorted = sorted;
  p.cell_start = start;
  p.grid = grid;
  p.nloc = nloc;
  p.mem_size = mem_size;
  int cap = mem_size < 64 ? 64 : mem_size;
  if (cap > 8192) cap = 8192;
  p.cap = cap;
  p.rcut2 = (FP)(rcut * rcut);  // float product, then converted (neighbor_list.cc:887)
  const size_t smem = (size_t)cap * 4 * 4;
  auto kern = k_build_rows<FP>;
  DPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  DPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 128, smem));
  if (occ < 1) occ = 1;
  long long want = ((long long)nloc + 3) / 4;
  long long capb = (long long)sm_count() * occ;
  kern<<<(int)(want < capb ? want : capb), 128, smem, st>>>(p);
  DPB_CUDA(cudaGetLastError());
  note_launches(6);  // bbox_init, bbox, grid_setup, cell_count, cell_fill, build_rows (+ scans above)
  int mx = 0;
  DPB_CUDA(cudaMemcpyAsync(&mx, maxl, sizeof(int), cudaMemcpyDeviceToHost, st));
  DPB_CUDA(cudaStreamSynchronize(st));
  *max_list_size = mx;
  return mx > mem_size ? 1 : DPB200_OK;
}

}  // namespace
}  // namespace dpb200

namespace dpb200 {
namespace {
// Dense rows [nrows][row_stride] -> the caller-owned rows of an InputNlist (firstneigh[i] = device pointer of row
// i, neighbor_list.h:20-57); ilist[i] = i % nloc as build_nlist of source/lib/src/gpu/neighbor_list.cu:78-93.
__global__ void k_scatter_rows(int* const* __restrict__ firstneigh, int* __restrict__ ilist,
                               const int* __restrict__ rows, int row_stride, const int* __restrict__ numneigh,
                               int nrows, int nloc) {
  const int lane = threadIdx.x & 31;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < nrows; r += nw) {
    int* __restrict__ dst = firstneigh[r];
    const int* __restrict__ src = rows + (long long)r * row_stride;
    const int n = numneigh[r];
    for (int j = lane; j < n; j += 32) dst[j] = src[j];
    if (lane == 0 && ilist) ilist[r] = r % nloc;
  }
}
}  // namespace
}  // namespace dpb200

extern "C" {

size_t dpb200_copy_coord_workspace_bytes(int nloc) {
  using namespace dpb200;
  const size_t n = align_up((size_t)(nloc > 0 ? nloc : 0) + 1, 64);
  return (2 * n + scan_tmp_ints((long long)nloc + 1) + 64) * sizeof(int);
}

size_t dpb200_build_nlist_workspace_bytes(int nall) { return dpb200::build_ws(nall > 0 ? nall : 0, 8).total; }

#define DPB200_DEF_NL(SUF, FP)                                                                     \
  int dpb200_normalize_coord_##SUF(FP* coord, int natom, const FP* boxt, dpb200_stream_t stream) { \
    return dpb200::do_normalize<FP>(coord, natom, boxt, (cudaStream_t)stream);                     \
  }                                                                                                \
  int dpb200_copy_coord_##SUF(FP* out_c, int* out_t, int* mapping, int* nall, const FP* in_c,      \
                              const int* in_t, int nloc, int mem_nall, float rcut, const FP* boxt, \
                              void* workspace, size_t workspace_bytes, dpb200_stream_t stream) {   \
    return dpb200::do_copy_coord<FP>(out_c, out_t, mapping, nall, in_c, in_t, nloc, mem_nall,      \
                                     rcut, nullptr, nullptr, boxt, workspace, workspace_bytes,     \
                                     (cudaStream_t)stream);                                        \
  }                                                                                                \
  int dpb200_copy_coord_cells_##SUF(FP* out_c, int* out_t, int* mapping, int* nall,                \
                                    const FP* in_c, const int* in_t, int nloc, int mem_nall,       \
                                    const int* ncell, const int* ngcell, const FP* boxt,           \
                                    void* workspace, size_t workspace_bytes,                       \
                                    dpb200_stream_t stream) {                                      \
    return dpb200::do_copy_coord<FP>(out_c, out_t, mapping, nall, in_c, in_t, nloc, mem_nall,      \
                                     0.f, ncell, ngcell, boxt, workspace, workspace_bytes,         \
                                     (cudaStream_t)stream);                                        \
  }                                                                                                \
  int dpb200_build_nlist_##SUF(int* numneigh, int* rows, int* max_list_size, const FP* coord,      \
                               int nloc, int nall, int mem_size, float rcut, const int* type,      \
                               void* workspace, size_t workspace_bytes, dpb200_stream_t stream) {  \
    return dpb200::do_build_nlist<FP>(numneigh, rows, max_list_size, coord, nloc, nall, mem_size,  \
                                      rcut, type, workspace, workspace_bytes,                      \
                                      (cudaStream_t)stream);                                       \
  }
DPB200_DEF_NL(f64, double)
DPB200_DEF_NL(f32, float)
#undef DPB200_DEF_NL

int dpb200_scatter_nlist_rows(int* const* firstneigh, int* ilist, const int* rows, int row_stride,
                              const int* numneigh, int nrows, int nloc, dpb200_stream_t stream) {
  using namespace dpb200;
  DPB_REQUIRE(nrows >= 0 && row_stride >= 0 && nloc >= 1, "scatter_nlist_rows: bad shape");
  if (nrows == 0) return DPB200_OK;
  DPB_REQUIRE(firstneigh && rows && numneigh, "scatter_nlist_rows: null pointer");
  int grid = ceil_div(nrows, 8);
  const int cap = sm_count() * 8;
  if (grid > cap) grid = cap;
  k_scatter_rows<<<grid, 256, 0, (cudaStream_t)stream>>>(firstneigh, ilist, rows, row_stride, numneigh, nrows, nloc);
  DPB_CUDA(cudaGetLastError());
  note_launches(1);
  return DPB200_OK;
}

}  // extern "C"
