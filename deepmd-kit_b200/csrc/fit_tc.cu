// Fitting-net GEMMs on the Blackwell tensor cores (tcgen05.mma kind::i8, accumulators in TMEM, operands
// staged by TMA) -- SURVEY 8f-1; algebra deepmd/pt/model/network/mlp.py (tanh, resnet_dt, skip connection),
// composition deepmd/pt/model/descriptor/se_a.py:843-850 + the energy fitting net; the reference's fused fp32
// analogue is source/op/pt/graph_fitting.cu:40-90,230-340.
//
// fp64 parity (1e-10) rules out a reduced-precision product, and B200's FP64 tensor pipe peaks at ~33-40
// TFLOP/s, so every GEMM of the net is made EXACT on the int8 tensor cores by operand splitting
// (csrc/fitting.cu has the digit convention):
//     x = 2^Ex sum_i X_i 2^(-7-8i),   w = 2^Ew sum_j W_j 2^(-7-8j),    X_i, W_j balanced base-256 digits (int8)
//     x.w = 2^(Ex+Ew-14) sum_d 2^(-8d) S_d,     S_d = sum_{i+j=d} X_i.W_j   (int32, error free),  d < NS
// (NS = 6: 47 fraction bits per operand; the dropped orders d >= NS are below 2^-48 of the row * column scale)
// One CTA owns a [128 rows x NT columns] output tile and keeps ALL NS order accumulators S_d of the tile in
// tensor memory (NS*NT <= 512 columns of 128 lanes x 32 bit); the NS(NS+1)/2 slice products of one K-chunk are
// tcgen05.mma instructions issued by one thread against the same staged operand tiles, so the operands are
// read from shared memory once per chunk and the int32 partial products never exist outside TMEM.  The
// epilogue reads the accumulators back with tcgen05.ld (one thread = one row), recombines them in fp64 and
// applies the layer's elementwise chain in the same pass:
//     EPI_FWD   z = x.w + b, t = tanh(z), y = t*idt (+ y_prev)      -> t, y (fp64) and y as int8 slices (the next
//               layer's A operand, fixed exponent: |y| is bounded by sum |idt|)
//     EPI_BWD   g = dz.W^T (+ g_next | + w_head), dz' = g * idt * (1 - t^2)   -> g, dz' (fp64)
//     EPI_PLAIN out = x.w  (row-major fp64: dE/dD handed to the descriptor backward)
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue
// (warp w reads TMEM lanes 32*(w%4)..).  Persistent grid over tiles, column tile fastest so that concurrently
// running CTAs share the A rows in L2.
//
// fp64 intermediates of the net (t, y, g, dz) use a row-blocked column-major layout
//     element (r, c) at ((r / 128) * N + c) * 128 + r % 128
// so that "one thread = one row" global accesses are coalesced (32 consecutive rows = 256 contiguous bytes).
#include <cuda.h>
#include <cudaTypedefs.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace dpb200 {
namespace {

constexpr int kTileM = 128;
constexpr int kChunkK = 64;   // bytes of K per operand unit = one SWIZZLE_64B row
constexpr int kStages = 5;    // operand ring of HALF chunks: slices [0, NS/2) or [NS/2, NS) of one K-chunk
constexpr int kEpiWarps = 20; // 5 column groups of 16 x 4 TMEM lane quadrants
#ifndef DPB_EPI_BATCH
#define DPB_EPI_BATCH 8
#endif
constexpr int kEpiBatch = DPB_EPI_BATCH;  // columns whose fp64 inputs are fetched together in the epilogue
constexpr int kThreads = 32 * (2 + kEpiWarps);

enum { EPI_FWD = 0, EPI_BWD = 1, EPI_PLAIN = 2 };

struct GemmParams {
  long long n;        // rows
  int N;              // output columns
  int nk;             // K chunks of 64
  int n_tiles;        // column tiles
  long long m_blocks; // row blocks of 128
  const int* row_exp; // per-row exponent of the A operand, or null -> row_exp_fixed
  int row_exp_fixed;
  const double* colv;       // [N][4] per output column: {2^(col_exp - 14), add, mul, 0}
                            //   FWD: add = bias, mul = idt (1 without resnet_dt)
                            //   BWD: add = head weight (0 unless the layer above is the head), mul = idt of the layer below
  const double* skip;       // blocked [n][N] or null (FWD: y_prev; BWD: g of the layer above)
  const double* t_in;       // blocked [n][N]  (BWD: tanh values of the layer below)
  double* out0;             // FWD: t (blocked); BWD: g (blocked, nullable); PLAIN: row-major
  double* out1;             // FWD: y (blocked); BWD: dz (blocked)
  signed char* slices_out;  // FWD: y as [n][NS][Kp_out] int8 or null
  long long ld_slices;
  int Kp_out;
  int out_exp;
  long long ld_out;  // PLAIN: leading dimension
  int wide_store;    // PLAIN: rows are 32-byte aligned (256-bit stores)
  int out_f32;       // PLAIN: out0 is a float matrix (the fp32 model's dE/dD), 16-byte aligned rows
  long long* dbg;    // optional per-role cycle counters of every CTA ([grid][16]; profiling only)
};

// ---------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// A legitimate wait lasts microseconds; a pipeline bug must fail loudly instead of hanging the device.
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) {
      printf("dpb200 fit_gemm: mbarrier wait timed out (block %d thread %d barrier %u parity %u)\n", (int)blockIdx.x,
             (int)threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (!mbar_try_wait(bar, parity)) mbar_wait_slow(bar, parity);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::
          "r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
// multicast variant: the box lands at the same CTA-relative offset of every CTA in `mask` and signals the barrier
// at the same offset there
__device__ __forceinline__ void tma_load_3d_mc(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2,
                                               uint32_t bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, "
      "%3, %4}], [%5], %6;" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar), "h"(mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// arrive on the barrier at the same offset in every CTA of `mask` once the MMAs issued so far have completed
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, int8 x int8 -> int32
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 16 consecutive 32-bit columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor of a K-major tile [rows][64 B] written by TMA with SWIZZLE_64B
// (cute::UMMA::SmemDescriptor: start address, LBO (unused for swizzled K-major), SBO = 8 rows * 64 B,
// version 1, layout type 4).
__device__ __forceinline__ uint64_t smem_desc_sw64(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}

// One elected lane of a converged warp (the compiler then knows the region is single-threaded and feeds the
// uniform-register operands of UTCIMMA without a per-instruction waterfall loop).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
// descriptor of the tile `byte_off` bytes after the one described by `base_lo` (same layout; < 256 KB of smem)
__device__ __forceinline__ uint64_t desc_at(uint32_t base_lo, uint32_t byte_off) {
  return ((uint64_t)0x80004020u << 32) | (uint64_t)(base_lo + (byte_off >> 4));
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): D = s32, A = B = signed int8, both K-major, M x N.
__host__ __device__ constexpr uint32_t idesc_i8(int M, int N) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ double pow2i(int e) { return __hiloint2double((1023 + e) << 20, 0); }

// NS slices of 16 fp64 values -> one 16-byte store per slice.  Fixed-point image u = round(v * up) + bias with
// bias = sum_k 128 * 256^k: byte k of u is digit_k + 128, so the signed digit is that byte with its top bit flipped.
// Slice s (most significant first) holds byte NS-1-s of the 16 values.
template <int NS>
__device__ __forceinline__ void store_slices16(signed char* __restrict__ base, long long slice_stride,
                                               const double (&v)[16], double up) {
  static_assert(NS <= 6, "the fixed-point image must fit in 48 bits (it is read from the mantissa of a double)");
  unsigned long long bias = 0;
#pragma unroll
  for (int k = 0; k < NS; ++k) bias = bias * 256ull + 128ull;
  // one FMA against 2^52 + 2^51 + bias leaves the rounded image in the low mantissa bits (|image| < 2^48): no F2I
  const double magic = 6755399441055744.0 + (double)bias;
  unsigned lo[16], hi[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const double f = fma(v[j], up, magic);
    lo[j] = (unsigned)__double2loint(f);
    hi[j] = (unsigned)__double2hiint(f);
  }
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    const int kb = NS - 1 - s;  // byte of the image
    unsigned w[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const unsigned sel = (unsigned)(kb & 3) | ((unsigned)(4 + (kb & 3)) << 4);
      const unsigned t0 = kb < 4 ? __byte_perm(lo[4 * g], lo[4 * g + 1], sel) : __byte_perm(hi[4 * g], hi[4 * g + 1], sel);
      const unsigned t1 =
          kb < 4 ? __byte_perm(lo[4 * g + 2], lo[4 * g + 3], sel) : __byte_perm(hi[4 * g + 2], hi[4 * g + 3], sel);
      w[g] = __byte_perm(t0, t1, 0x5410) ^ 0x80808080u;
    }
    *reinterpret_cast<uint4*>(base + s * slice_stride) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

extern __shared__ __align__(16) unsigned char fit_smem_raw[];

template <int NS, int NT>
struct SmemLayout {
  static constexpr int NH = NS / 2;                 // slices per half chunk
  static constexpr int kABytes = kTileM * kChunkK;  // one slice of one A chunk
  static constexpr int kBBytes = NT * kChunkK;
  static constexpr int kAHalf = NH * kABytes;       // [NH][128][64]
  static constexpr int kBHalf = NH * kBBytes;       // [NH][NT][64]: consecutive slices = consecutive row groups
  static constexpr int kStageBytes = kAHalf + kBHalf;
  static constexpr int kBarOff = kStages * kStageBytes;
  static constexpr int kColvOff = kBarOff + 512;       // per epilogue warp: colv of its 16 columns (512 B)
  static constexpr int kTotal = kColvOff + kEpiWarps * 512 + 1024;  // + slack to align the tiles to 1024 B
};

// int32 -> double without the conversion unit: 2^52 + 2^31 + a has the bits {0x43300000, a ^ 0x80000000}
__device__ __forceinline__ double i2d(int a) {
  return __hiloint2double(0x43300000, a ^ (int)0x80000000) - 4503601774854144.0;
}

// Branch-free tanh, absolute error ~1e-16 (the callers need absolute, not relative, accuracy: t feeds
// y = t*idt + skip and 1 - t^2):  u = exp(-2|x|), tanh = sign(x) (1 - u) / (1 + u).
__device__ __forceinline__ double tanh_bf(double x) {
  const double a = fmin(fabs(x), 20.0);
  const double t = -2.0 * a;
  double kf = fma(t, 1.4426950408889634, 6755399441055744.0);
  const int ki = __double2loint(kf);
  kf -= 6755399441055744.0;
  double r = fma(kf, -6.93147180369123816490e-01, t);
  r = fma(kf, -1.90821492927058770002e-10, r);
  double p = 2.505210838544172e-08;              // 1/11!  (|r| <= ln2/2: truncation 6e-15 of u)
  p = fma(p, r, 2.755731922398589e-07);          // 1/10!
  p = fma(p, r, 2.7557319223985893e-06);         // 1/9!
  p = fma(p, r, 2.48015873015873e-05);           // 1/8!
  p = fma(p, r, 1.984126984126984e-04);          // 1/7!
  p = fma(p, r, 1.388888888888889e-03);          // 1/6!
  p = fma(p, r, 8.333333333333333e-03);          // 1/5!
  p = fma(p, r, 4.1666666666666664e-02);         // 1/4!
  p = fma(p, r, 1.6666666666666666e-01);         // 1/3!
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const double u = p * __hiloint2double((1023 + ki) << 20, 0);
  const double num = 1.0 - u, den = 1.0 + u;  // den in [1, 2]
  double rc;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rc) : "d"(den));
  rc = fma(fma(-den, rc, 1.0), rc, rc);
  rc = fma(fma(-den, rc, 1.0), rc, rc);
  return copysign(num * rc, x);
}

__device__ __forceinline__ void st_256(double* p, double a, double b, double c, double d) {
  asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

__device__ __forceinline__ double4 ldg_256(const double* p) {
  double4 v;
  asm("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
  return v;
}

// CN = CTAs per cluster.  The CN CTAs of a cluster work on CN neighbouring column tiles of the same row block in
// lockstep and share the A operand: every CTA fetches 1/CN of each A stage and TMA multicasts it into the shared
// memory of all of them (CN = 3: one slice each; CN = 4: 32 rows of every slice each).  The kernel is bound by
// the L2 -> SM path (A 48 KB + B 30 KB per K-chunk against ~1700 cycles of MMA): multicast removes (CN-1)/CN of
// the A traffic.
template <int NS, int NT, int EPI, int CN>
__global__ void __launch_bounds__(kThreads, 1)
    k_fit_gemm(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ GemmParams p) {
  using L = SmemLayout<NS, NT>;
  constexpr int NH = L::NH;
  static_assert(CN == 1 || CN == NH || CN == 4, "A is shared by slice (CN = NS/2) or by 32-row quarters (CN = 4)");
  static_assert(NS == 4 || NS == 6, "built for 6 (fp64 model) and 4 (fp32 model) operand slices");
  constexpr uint16_t kMaskAll = (uint16_t)((1u << CN) - 1u);
  const uint32_t crank = CN > 1 ? cluster_ctarank() : 0u;
  static_assert(NS % 2 == 0 && NS * NT <= 512 && NT % 16 == 0 && NH * NT <= 256,
                "accumulators must fit in tensor memory, one MMA spans up to NS/2 slices of B");
  static_assert(NT / 16 * 4 == kEpiWarps, "one epilogue warp per (lane quadrant, 16-column group)");
  static_assert(L::kStageBytes % 1024 == 0 && L::kAHalf % 1024 == 0, "tiles must stay 1024-byte aligned");
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  unsigned char* fit_smem = fit_smem_raw + ((1024u - (smem_u32(fit_smem_raw) & 1023u)) & 1023u);
  const uint32_t smem0 = smem_u32(fit_smem);
  const uint32_t bar0 = smem0 + L::kBarOff;
  // barriers: full[kStages], empty[kStages], tmem_full, tmem_empty ; then the TMEM base address
  const uint32_t bar_full = bar0, bar_empty = bar0 + 8 * kStages, bar_tfull = bar0 + 16 * kStages,
                 bar_tempty = bar_tfull + 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(fit_smem + L::kBarOff + 16 * kStages + 16);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, CN);  // released by the MMA warps of all CTAs that receive the multicast
    }
    mbar_init(bar_tfull, 1);
    mbar_init(bar_tempty, 32 * kEpiWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  if (CN > 1)
    cluster_sync_all();  // peers' barriers are initialised before any multicast or remote arrive can reach them
  else
    __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // work units of a cluster: (row block, group of CN column tiles); CTA `crank` takes column tile group*CN + crank
  const int groups = p.n_tiles / CN;
  const long long total = p.m_blocks * groups;
  const long long unit0 = blockIdx.x / CN, unit_step = gridDim.x / CN;
  const int nk = p.nk;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    // One stage = half a K-chunk: A slices [h*NH, (h+1)*NH) as [NH][128][64 B] and the same slices of B as
    // [NH][NT][64 B], one bulk tensor copy each.
    if (lane == 0) {
      uint32_t it = 0;
      long long w_empty = 0;
      const long long t_begin = clock64();
      for (long long tile = unit0; tile < total; tile += unit_step) {
        const int mb = (int)(tile / groups);
        const int nb = (int)(tile - (long long)mb * groups) * CN + (int)crank;
        for (int kc = 0; kc < nk; ++kc) {
#pragma unroll
          for (int h = 0; h < 2; ++h, ++it) {
            const uint32_t stage = it % kStages;
            const long long c0 = p.dbg ? clock64() : 0;
            mbar_wait(bar_empty + 8 * stage, ((it / kStages) & 1u) ^ 1u);
            if (p.dbg) w_empty += clock64() - c0;
            const uint32_t fb = bar_full + 8 * stage;
            mbar_expect_tx(fb, (uint32_t)L::kStageBytes);
            const uint32_t sa = smem0 + stage * L::kStageBytes;
            if (CN == 1) {
              tma_load_3d(sa, &tmA, kc * kChunkK, mb * kTileM, h * NH, fb);
            } else if (CN == NH) {  // my slice of the half chunk, to everybody
              tma_load_3d_mc(sa + crank * L::kABytes, &tmA, kc * kChunkK, mb * kTileM, h * NH + (int)crank, fb, kMaskAll);
            } else {  // my 32 rows of every slice, to everybody
#pragma unroll
              for (int sl = 0; sl < NH; ++sl)
                tma_load_3d_mc(sa + sl * L::kABytes + crank * (kTileM / 4) * kChunkK, &tmA, kc * kChunkK,
                               mb * kTileM + (int)crank * (kTileM / 4), h * NH + sl, fb, kMaskAll);
            }
            tma_load_3d(sa + L::kAHalf, &tmB, kc * kChunkK, nb * NT, h * NH, fb);
          }
        }
      }
      if (p.dbg) {
        p.dbg[blockIdx.x * 16 + 0] = clock64() - t_begin;
        p.dbg[blockIdx.x * 16 + 1] = w_empty;
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    // The product A_i . B_j belongs to order i + j, whose accumulator is the column block (i + j) * NT of tensor
    // memory.  Slices of B that are consecutive in shared memory are consecutive row groups of ONE K-major tile, so
    // A_i . [B_j; B_j+1; ..] is a single tcgen05.mma of N = NT * count whose result lands in the consecutive
    // order blocks i + j, i + j + 1, ...: 9 instructions per K-step instead of 21, and the A tile is read from
    // shared memory 9 times instead of 21 (the N = NT products are bound by the shared-memory read of A).
    const uint32_t lo0 = (uint32_t)smem_desc_sw64(smem0);
    const bool leader = elect_one();
    uint32_t it = 0, acc_phase = 0;
    long long w_tempty = 0, w_full = 0;
    const long long t_begin = clock64();
    for (long long tile = unit0; tile < total; tile += unit_step) {
      long long c0 = p.dbg ? clock64() : 0;
      mbar_wait(bar_tempty, acc_phase ^ 1);
      if (p.dbg) w_tempty += clock64() - c0;
      tc_fence_after();
      for (int kc = 0; kc < nk; ++kc, it += 2) {
        const uint32_t s0 = it % kStages, s1 = (it + 1) % kStages;
        const uint32_t o0 = s0 * L::kStageBytes, o1 = s1 * L::kStageBytes;
        c0 = p.dbg ? clock64() : 0;
        mbar_wait(bar_full + 8 * s0, (it / kStages) & 1u);
        // the first chunk of a tile needs both halves at once: A_0 . B_hi must be the first write of the upper orders
        if (kc == 0) mbar_wait(bar_full + 8 * s1, ((it + 1) / kStages) & 1u);
        if (p.dbg) w_full += clock64() - c0;
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int k = 0; k < kChunkK / 32; ++k) {
            const uint32_t acc = (kc == 0 && k == 0) ? 0u : 1u;
            // A_0 . B_lo -> orders 0..NH-1 ; A_0 . B_hi -> orders NH..NS-1 (first chunk only, see above)
            mma_i8(tmem_base, desc_at(lo0, o0 + k * 32), desc_at(lo0, o0 + L::kAHalf + k * 32), idesc_i8(kTileM, NH * NT),
                   acc);
            if (kc == 0)
              mma_i8(tmem_base + NH * NT, desc_at(lo0, o0 + k * 32), desc_at(lo0, o1 + L::kAHalf + k * 32),
                     idesc_i8(kTileM, NH * NT), acc);
#pragma unroll
            for (int i = 1; i < NH; ++i)  // A_i . B_lo -> orders i..i+NH-1
              mma_i8(tmem_base + i * NT, desc_at(lo0, o0 + i * L::kABytes + k * 32),
                     desc_at(lo0, o0 + L::kAHalf + k * 32), idesc_i8(kTileM, NH * NT), 1u);
          }
        }
        __syncwarp();
        if (kc != 0) {
          c0 = p.dbg ? clock64() : 0;
          mbar_wait(bar_full + 8 * s1, ((it + 1) / kStages) & 1u);
          if (p.dbg) w_full += clock64() - c0;
          tc_fence_after();
        }
        if (leader) {
#pragma unroll
          for (int k = 0; k < kChunkK / 32; ++k) {
            if (kc != 0)
              mma_i8(tmem_base + NH * NT, desc_at(lo0, o0 + k * 32), desc_at(lo0, o1 + L::kAHalf + k * 32),
                     idesc_i8(kTileM, NH * NT), 1u);
#pragma unroll
            for (int i = 1; i < NH; ++i)  // A_i . B_hi[0 .. NH-i) -> orders NH+i..NS-1
              mma_i8(tmem_base + (NH + i) * NT, desc_at(lo0, o0 + i * L::kABytes + k * 32),
                     desc_at(lo0, o1 + L::kAHalf + k * 32), idesc_i8(kTileM, (NH - i) * NT), 1u);
#pragma unroll
            for (int i = NH; i < NS; ++i)  // A_i (upper half) . B_lo[0 .. NS-i) -> orders i..NS-1
              mma_i8(tmem_base + i * NT, desc_at(lo0, o1 + (i - NH) * L::kABytes + k * 32),
                     desc_at(lo0, o0 + L::kAHalf + k * 32), idesc_i8(kTileM, (NS - i) * NT), 1u);
          }
          if (CN == 1) {
            tc_commit(bar_empty + 8 * s0);
            tc_commit(bar_empty + 8 * s1);
          } else {
            tc_commit_mc(bar_empty + 8 * s0, kMaskAll);
            tc_commit_mc(bar_empty + 8 * s1, kMaskAll);
          }
          if (kc == nk - 1) tc_commit(bar_tfull);
        }
        __syncwarp();
      }
      acc_phase ^= 1;
    }
    if (p.dbg && lane == 0) {
      p.dbg[blockIdx.x * 16 + 2] = clock64() - t_begin;
      p.dbg[blockIdx.x * 16 + 3] = w_tempty;
      p.dbg[blockIdx.x * 16 + 4] = w_full;
    }
  } else {
    // ---------------------------------------------------------------- epilogue (thread = row, 16 columns)
    const int quad = warp & 3;
    const int cg = (warp - 2) >> 2;
    const int cc = cg * 16;
    const int row_in_tile = quad * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)cc;
    double* cvs = reinterpret_cast<double*>(fit_smem + L::kColvOff + (warp - 2) * 512);
    uint32_t acc_phase = 0;
    long long w_tfull = 0, t_drain = 0, t_math = 0;
    const long long t_begin = clock64();
    for (long long tile = unit0; tile < total; tile += unit_step) {
      const int mb = (int)(tile / groups);
      const int nb = (int)(tile - (long long)mb * groups) * CN + (int)crank;
      const long long r = (long long)mb * kTileM + row_in_tile;
      const bool row_ok = r < p.n;
      const int c0 = nb * NT + cc;
      const bool live = c0 < p.N;  // (warp-uniform)
      // blocked offset of (r, c0): ((mb * N + c) * 128 + row_in_tile)
      const long long boff = ((long long)mb * p.N + c0) * kTileM + row_in_tile;
      const double rs = pow2i(live && row_ok && p.row_exp ? __ldg(p.row_exp + r) : p.row_exp_fixed);
      // the fp64 inputs of the elementwise chain are streamed from HBM: pull them into L2 while the MMAs run
      if (live && EPI != EPI_PLAIN) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (p.skip) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.skip + boff + (long long)j * kTileM));
          if (EPI == EPI_BWD) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.t_in + boff + (long long)j * kTileM));
        }
      }
      // per-column constants of this warp's 16 columns -> shared memory (read back as broadcasts in the math below;
      // a global load per element would put an L2 round trip on every element's critical path)
      __syncwarp();
      if (live && lane < 16) {
        const double4 c4 = ldg_256(p.colv + 4 * (long long)(c0 + lane));
        *reinterpret_cast<double2*>(cvs + 4 * lane) = make_double2(c4.x, c4.y);
        *reinterpret_cast<double2*>(cvs + 4 * lane + 2) = make_double2(c4.z, c4.w);
      }
      __syncwarp();
      const long long e0 = p.dbg ? clock64() : 0;
      mbar_wait(bar_tfull, acc_phase);
      const long long e1 = p.dbg ? clock64() : 0;
      tc_fence_after();
      double v[16];
      if (live) {
        int a[16];
        tmem_ld16(lane_addr + (uint32_t)((NS - 1) * NT), a);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = i2d(a[j]);
#pragma unroll
        for (int d = NS - 2; d >= 0; --d) {
          tmem_ld16(lane_addr + (uint32_t)(d * NT), a);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = fma(v[j], 0.00390625, i2d(a[j]));
        }
      }
      // the accumulators are in registers: the next tile's products may overwrite tensor memory now
      tc_fence_before();
      mbar_arrive(bar_tempty);
      acc_phase ^= 1;
      const long long e2 = p.dbg ? clock64() : 0;
      if (p.dbg) w_tfull += e1 - e0, t_drain += e2 - e1;
      if (!live) continue;
      // N is a multiple of 16 (host check) and the blocked matrices are padded to whole row blocks, so only the
      // row-major outputs (slices, PLAIN) need the row predicate.  Inputs are fetched 8 columns at a time.
      if (EPI == EPI_FWD) {
        const double* __restrict__ sk = p.skip ? p.skip + boff : nullptr;
        double* __restrict__ o0 = p.out0 + boff;
        double* __restrict__ o1 = p.out1 + boff;
#pragma unroll
        for (int h = 0; h < 16; h += kEpiBatch) {
          double xin[kEpiBatch];
#pragma unroll
          for (int j = 0; j < kEpiBatch; ++j) xin[j] = sk ? __ldcs(sk + (h + j) * kTileM) : 0.;
#pragma unroll
          for (int j = 0; j < kEpiBatch; ++j) {
            const double2 c01 = *reinterpret_cast<const double2*>(cvs + 4 * (h + j));
            const double2 c23 = *reinterpret_cast<const double2*>(cvs + 4 * (h + j) + 2);
            const double4 c4 = make_double4(c01.x, c01.y, c23.x, c23.y);
            const double t = tanh_bf(fma(v[h + j], rs * c4.x, c4.y));
            const double yy = fma(t, c4.z, xin[j]);
            v[h + j] = yy;
            __stcs(o0 + (h + j) * kTileM, t);
            __stcs(o1 + (h + j) * kTileM, yy);
          }
        }
        if (row_ok && p.slices_out) {
          const double up = pow2i(7 + 8 * (NS - 1) - p.out_exp);
          store_slices16<NS>(p.slices_out + r * p.ld_slices + c0, p.Kp_out, v, up);
        }
      } else if (EPI == EPI_BWD) {
        const double* __restrict__ sk = p.skip ? p.skip + boff : nullptr;
        const double* __restrict__ ti = p.t_in + boff;
        double* __restrict__ o0 = p.out0 ? p.out0 + boff : nullptr;
        double* __restrict__ o1 = p.out1 + boff;
#pragma unroll
        for (int h = 0; h < 16; h += kEpiBatch) {
          double xin[kEpiBatch], tin[kEpiBatch];
#pragma unroll
          for (int j = 0; j < kEpiBatch; ++j) {
            xin[j] = sk ? __ldcs(sk + (h + j) * kTileM) : 0.;
            tin[j] = __ldcs(ti + (h + j) * kTileM);
          }
#pragma unroll
          for (int j = 0; j < kEpiBatch; ++j) {
            const double2 c01 = *reinterpret_cast<const double2*>(cvs + 4 * (h + j));
            const double2 c23 = *reinterpret_cast<const double2*>(cvs + 4 * (h + j) + 2);
            const double4 c4 = make_double4(c01.x, c01.y, c23.x, c23.y);
            const double g = fma(v[h + j], rs * c4.x, xin[j] + c4.y);
            const double dz = g * c4.z * fma(-tin[j], tin[j], 1.0);
            if (o0) __stcs(o0 + (h + j) * kTileM, g);
            __stcs(o1 + (h + j) * kTileM, dz);
          }
        }
      } else {
        if (row_ok) {
          double* __restrict__ o = p.out0 + (p.out_f32 ? 0 : r * p.ld_out + c0);
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] *= rs * cvs[4 * j];
          if (p.out_f32) {
            float* __restrict__ of = reinterpret_cast<float*>(p.out0) + r * p.ld_out + c0;
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              __stcs(reinterpret_cast<float4*>(of + j),
                     make_float4((float)v[j], (float)v[j + 1], (float)v[j + 2], (float)v[j + 3]));
          } else if (p.wide_store) {  // 32-byte aligned rows: one full sector per store
#pragma unroll
            for (int j = 0; j < 16; j += 4) st_256(o + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) __stcs(o + j, v[j]);
          }
        }
      }
      if (p.dbg) t_math += clock64() - e2;
    }
    if (p.dbg && lane == 0 && (warp == 2 || warp == 21)) {
      long long* d = p.dbg + blockIdx.x * 16 + (warp == 2 ? 5 : 9);
      d[0] = clock64() - t_begin;
      d[1] = w_tfull;
      d[2] = t_drain;
      d[3] = t_math;
    }
  }

  tc_fence_before();
  if (CN > 1)
    cluster_sync_all();  // no CTA leaves while a peer may still multicast into it or arrive on its barriers
  else
    __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------------
// Row slicing of a blocked fp64 matrix into the int8 A-operand format with a per-row exponent, one thread per
// row.  HEAD: the matrix is t of the last hidden layer; the values sliced are dz = w_head * idt * (1 - t^2)
// (the seed of the backward chain) and the atomic energy e = y . w_head + b_head is produced on the way.
// ------------------------------------------------------------------------------------------------------
// One block = 32 rows x (N / 16) warps: warp w owns columns [16 w, 16 w + 16), lane = row.  Every load instruction of
// a warp reads 256 contiguous bytes of the blocked layout; the 16 values stay in registers while the row maximum is
// reduced across the warps through shared memory, so the matrix is read exactly once.
template <int NS, bool HEAD>
__global__ void __launch_bounds__(1024) k_fit_slice(signed char* __restrict__ out, long long ld_out, int Kp,
                                                    int* __restrict__ row_exp, const double* __restrict__ x,
                                                    const double* __restrict__ y, const double* __restrict__ w_head,
                                                    const double* __restrict__ idt, double b_head,
                                                    double* __restrict__ e_out, long long n, int N) {
  __shared__ int s_exp[32];
  __shared__ double s_e[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long r = (long long)blockIdx.x * 32 + lane;
  const int c0 = warp * 16;
  if (warp == 0) {
    s_exp[lane] = -2000;
    s_e[lane] = b_head;
  }
  __syncthreads();
  // (r / 128) * N * 128 + c * 128 + r % 128
  const long long boff = (r >> 7) * (long long)N * kTileM + (long long)c0 * kTileM + (r & (kTileM - 1));
  double v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __ldcs(x + boff + (long long)j * kTileM);
  if (HEAD) {
    double yv[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) yv[j] = __ldcs(y + boff + (long long)j * kTileM);
    double e = 0.;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const double w = __ldg(w_head + c0 + j);
      e = fma(yv[j], w, e);
      double d = w * fma(-v[j], v[j], 1.0);
      if (idt) d *= __ldg(idt + c0 + j);
      v[j] = d;
    }
    atomicAdd(&s_e[lane], e);
  }
  double m = 0.;
#pragma unroll
  for (int j = 0; j < 16; ++j) m = fmax(m, fabs(v[j]));
  atomicMax(&s_exp[lane], ((__double2hiint(m) >> 20) & 0x7ff) - 1023 + 2);
  __syncthreads();
  if (r >= n) return;
  int E = s_exp[lane];
  E = E < -900 ? -900 : (E > 900 ? 900 : E);
  if (warp == 0) {
    row_exp[r] = E;
    if (HEAD) e_out[r] = s_e[lane];
  }
  store_slices16<NS>(out + r * ld_out + c0, Kp, v, pow2i(7 + 8 * (NS - 1) - E));
}

// Extra operand columns that are a table look-up per row (se_atten: the centre atom's type embedding behind the
// descriptor): columns [col0, col0 + width) of every digit slice of row r get the digits of src[idx[r]][c] at the
// row's exponent row_exp[r] (written earlier by the producer of the other columns, which must have bounded it from
// below so that these values fit: |v| < 2^(E-1)).
template <int NS>
__global__ void k_fit_slice_cols(signed char* __restrict__ out, long long ld_out, long long slice_stride, int col0,
                                 int width, const int* __restrict__ row_exp, const double* __restrict__ src,
                                 int src_ld, const int* __restrict__ idx, long long n) {
  unsigned long long bias = 0;
#pragma unroll
  for (int k = 0; k < NS; ++k) bias = bias * 256ull + 128ull;
  const double magic = 6755399441055744.0 + (double)bias;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n * width;
       e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / width;
    const int c = (int)(e - r * width);
    const int E = row_exp[r];
    const long long sr = idx ? (long long)idx[r] : r;
    const double f = fma(src[sr * src_ld + c], pow2i(7 + 8 * (NS - 1) - E), magic);
    const unsigned long long u =
        ((unsigned long long)(unsigned)__double2hiint(f) << 32) | (unsigned long long)(unsigned)__double2loint(f);
    signed char* __restrict__ o = out + r * ld_out + col0 + c;
#pragma unroll
    for (int s = 0; s < NS; ++s) o[s * slice_stride] = (signed char)(((u >> (8 * (NS - 1 - s))) & 0xffu) ^ 0x80u);
  }
}

// row-major [n][N] (leading dimension ld) <-> blocked layout (tests and the non-tensor-core callers)
__global__ void k_fit_to_blocked(double* __restrict__ dst, const double* __restrict__ src, long long ld, long long n,
                                 int N, int to_blocked) {
  const long long tot = ((n + kTileM - 1) / kTileM) * kTileM * N;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < tot; e += (long long)gridDim.x * blockDim.x) {
    // e enumerates the blocked order: fastest = row in block
    const long long blk = e / ((long long)N * kTileM);
    const long long rem = e - blk * (long long)N * kTileM;
    const int c = (int)(rem / kTileM);
    const long long r = blk * kTileM + (rem - (long long)c * kTileM);
    if (r < n) {
      if (to_blocked)
        dst[e] = src[r * ld + c];
      else
        dst[r * ld + c] = src[e];
    }
  }
}

PFN_cuTensorMapEncodeTiled tensor_map_encoder() {
  static std::mutex mu;
  static PFN_cuTensorMapEncodeTiled fn = nullptr;
  std::lock_guard<std::mutex> lk(mu);
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(ptr);
  }
  return fn;
}

// 3-D byte tensor {d0 (contiguous), d1, d2} with byte strides s1, s2 and a {64, b1, b2} box, SWIZZLE_64B.
int make_map(CUtensorMap* map, const void* base, unsigned long long d0, unsigned long long d1, unsigned long long d2,
             unsigned long long s1, unsigned long long s2, unsigned b1, unsigned b2) {
  PFN_cuTensorMapEncodeTiled enc = tensor_map_encoder();
  if (!enc) {
    set_error("fit_gemm: cuTensorMapEncodeTiled is not available from the CUDA driver");
    return DPB200_ERR_CUDA;
  }
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {s1, s2};
  cuuint32_t box[3] = {(cuuint32_t)kChunkK, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult rc = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    set_error("fit_gemm: cuTensorMapEncodeTiled failed with code " + std::to_string((int)rc));
    return DPB200_ERR_INVALID;
  }
  return DPB200_OK;
}

template <int NS, int NT, int EPI, int CN>
int launch_gemm(const CUtensorMap& ma, const CUtensorMap& mb, const GemmParams& p, cudaStream_t st) {
  using L = SmemLayout<NS, NT>;
  auto kern = k_fit_gemm<NS, NT, EPI, CN>;
  static int max_clusters = 0;  // clusters of CN CTAs that can be resident at once (one CTA per SM)
  if (max_clusters == 0) {
    DPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    if (CN > 1) {
      cudaLaunchConfig_t q = {};
      q.gridDim = dim3((unsigned)(sm_count() / CN * CN));
      q.blockDim = dim3(kThreads);
      q.dynamicSmemBytes = L::kTotal;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = CN;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      q.attrs = at;
      q.numAttrs = 1;
      int n = 0;
      DPB_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &q));
      max_clusters = n > 0 ? n : 1;
    } else {
      max_clusters = sm_count();
    }
  }
  const long long total = p.m_blocks * (p.n_tiles / CN);
  const int clusters = (int)(total < max_clusters ? total : max_clusters);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(clusters * CN));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = L::kTotal;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CN;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CN > 1 ? 1 : 0;
  DPB_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, mb, p));
  note_launches(1);
  return DPB200_OK;
}

template <int NS, int NT, int CN>
int launch_mode(int mode, const CUtensorMap& ma, const CUtensorMap& mb, const GemmParams& p, cudaStream_t st) {
  if (mode == EPI_FWD) return launch_gemm<NS, NT, EPI_FWD, CN>(ma, mb, p, st);
  if (mode == EPI_BWD) return launch_gemm<NS, NT, EPI_BWD, CN>(ma, mb, p, st);
  return launch_gemm<NS, NT, EPI_PLAIN, CN>(ma, mb, p, st);
}

}  // namespace
}  // namespace dpb200

namespace dpb200 {
namespace {
long long* g_fit_dbg = nullptr;

template <int NS>
int fit_gemm_impl(int mode, long long nrow, int N, int K, int nslice, const signed char* a_slices,
                           long long a_slice_stride, long long a_row_stride, const int* row_exp, int row_exp_fixed,
                           const signed char* b_slices, int b_k_stride, const double* colv, const double* skip,
                           const double* t_in, double* out0, double* out1, long long ld_out, signed char* slices_out,
                           long long ld_slices, int kp_out, int out_exp, dpb200_stream_t stream) {
  const bool out_f32 = mode == 3;  // plain product stored as float
  if (out_f32) mode = EPI_PLAIN;
  DPB_REQUIRE(nrow >= 0 && N >= 16 && N % 16 == 0 && K >= 1, "fit_gemm: N must be a positive multiple of 16");
  if (nrow == 0) return DPB200_OK;
  DPB_REQUIRE(a_slices && b_slices && colv && (out0 || mode == EPI_BWD) && ((uintptr_t)colv & 31) == 0,
              "fit_gemm: null pointer (colv must be 32-byte aligned)");
  DPB_REQUIRE(K % 16 == 0 && a_slice_stride % 16 == 0 && a_row_stride % 16 == 0 && b_k_stride % 64 == 0 &&
                  b_k_stride >= K && ((uintptr_t)a_slices & 15) == 0 && ((uintptr_t)b_slices & 15) == 0,
              "fit_gemm: operands must be 16-byte aligned with K a multiple of 16 and the weight rows padded to 64");
  DPB_REQUIRE(mode != EPI_FWD || out1, "fit_gemm: forward needs both outputs");
  DPB_REQUIRE(mode != EPI_BWD || (t_in && out1), "fit_gemm: backward needs t_in and the dz output");
  DPB_REQUIRE(!slices_out || (kp_out % 16 == 0 && kp_out >= N && ld_slices >= (long long)nslice * kp_out &&
                              ((uintptr_t)slices_out & 15) == 0 && ld_slices % 16 == 0),
              "fit_gemm: bad slice output layout");
  constexpr int NT = 80;
  GemmParams p;
  p.n = nrow;
  p.N = N;
  p.nk = (K + kChunkK - 1) / kChunkK;
  p.n_tiles = (N + NT - 1) / NT;
  p.m_blocks = (nrow + kTileM - 1) / kTileM;
  p.row_exp = row_exp;
  p.row_exp_fixed = row_exp_fixed;
  p.colv = colv;
  p.skip = skip;
  p.t_in = t_in;
  p.out0 = out0;
  p.out1 = out1;
  p.slices_out = slices_out;
  p.ld_slices = ld_slices;
  p.Kp_out = kp_out;
  p.out_exp = out_exp;
  p.ld_out = ld_out;
  p.dbg = g_fit_dbg;
  p.wide_store = (mode == EPI_PLAIN && ld_out % 4 == 0 && ((uintptr_t)out0 & 31) == 0) ? 1 : 0;
  p.out_f32 = out_f32 ? 1 : 0;
  DPB_REQUIRE(!out_f32 || (ld_out % 4 == 0 && ((uintptr_t)out0 & 15) == 0), "fit_gemm: float output rows must be 16-byte aligned");
  CUtensorMap ma, mb;
  // Cluster width: the CTAs of a cluster share A by TMA multicast (csrc kernel comment).  DPB200_FIT_CLUSTER=1
  // switches the sharing off (comparison runs).
  static const int cluster_env = [] {
    const char* e = getenv("DPB200_FIT_CLUSTER");
    return e ? atoi(e) : 0;
  }();
  int cn = 1;
  if (p.n_tiles % 4 == 0 && nrow >= 4 * kTileM) cn = 4;
  else if (NS == 6 && p.n_tiles % 3 == 0 && nrow >= 3 * kTileM) cn = 3;
  if (cluster_env == 1 || (cluster_env > 1 && p.n_tiles % cluster_env != 0)) cn = 1;
  else if ((cluster_env == 3 && NS == 6) || cluster_env == 4) cn = cluster_env;
  // A: [nrow][nslice][K] bytes seen as {K, row, slice} (byte strides given); the box lands in shared memory as
  // [slice][row][64]; rows / K beyond the tensor are zero-filled by TMA
  const unsigned a_rows = cn == 4 ? kTileM / 4 : kTileM, a_slices_box = cn == 1 ? NS / 2 : 1;
  int rc = make_map(&ma, a_slices, (unsigned long long)K, (unsigned long long)nrow, (unsigned long long)nslice,
                    (unsigned long long)a_row_stride, (unsigned long long)a_slice_stride, a_rows, a_slices_box);
  if (rc != DPB200_OK) return rc;
  // B: [nslice][N][b_k_stride] bytes (weights transposed: row = output column, K contiguous, zero padded)
  rc = make_map(&mb, b_slices, (unsigned long long)b_k_stride, (unsigned long long)N, (unsigned long long)nslice,
                (unsigned long long)b_k_stride, (unsigned long long)b_k_stride * N, NT, NS / 2);
  if (rc != DPB200_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (cn == 4) return launch_mode<NS, NT, 4>(mode, ma, mb, p, st);
  if constexpr (NS == 6) {
    if (cn == 3) return launch_mode<NS, NT, 3>(mode, ma, mb, p, st);
  }
  return launch_mode<NS, NT, 1>(mode, ma, mb, p, st);
}

}  // namespace
}  // namespace dpb200

extern "C" {

/* profiling hook: device buffer of [grid][16] cycle counters filled by the next fit_gemm launches (NULL: off) */
void dpb200_fit_gemm_debug(long long* dbg) { dpb200::g_fit_dbg = dbg; }

int dpb200_fit_gemm_i8_f64(int mode, long long nrow, int N, int K, int nslice, const signed char* a_slices,
                           long long a_slice_stride, long long a_row_stride, const int* row_exp, int row_exp_fixed,
                           const signed char* b_slices, int b_k_stride, const double* colv, const double* skip,
                           const double* t_in, double* out0, double* out1, long long ld_out, signed char* slices_out,
                           long long ld_slices, int kp_out, int out_exp, dpb200_stream_t stream) {
  using namespace dpb200;
  DPB_REQUIRE(mode >= 0 && mode <= 3,
              "fit_gemm: mode must be 0 (forward), 1 (backward), 2 (plain) or 3 (plain, float output)");
  DPB_REQUIRE(nslice == 6 || nslice == 4, "fit_gemm: built for 6 (fp64 model) or 4 (fp32 model) operand slices");
  if (nslice == 4)
    return fit_gemm_impl<4>(mode, nrow, N, K, nslice, a_slices, a_slice_stride, a_row_stride, row_exp, row_exp_fixed,
                            b_slices, b_k_stride, colv, skip, t_in, out0, out1, ld_out, slices_out, ld_slices, kp_out,
                            out_exp, stream);
  return fit_gemm_impl<6>(mode, nrow, N, K, nslice, a_slices, a_slice_stride, a_row_stride, row_exp, row_exp_fixed,
                          b_slices, b_k_stride, colv, skip, t_in, out0, out1, ld_out, slices_out, ld_slices, kp_out,
                          out_exp, stream);
}

int dpb200_fit_slice_rows_f64(signed char* out, long long ld_out, int kp, int* row_exp, const double* x,
                              long long nrow, int N, int nslice, dpb200_stream_t stream) {
  using namespace dpb200;
  DPB_REQUIRE(nslice == 6 || nslice == 4, "fit_slice_rows: built for 6 or 4 operand slices");
  DPB_REQUIRE(nrow >= 0 && N >= 16 && N % 16 == 0 && N <= 512 && kp == N && ld_out >= (long long)nslice * kp &&
                  ld_out % 16 == 0,
              "fit_slice_rows: N must be a multiple of 16 up to 512 and kp == N");
  if (nrow == 0) return DPB200_OK;
  DPB_REQUIRE(out && row_exp && x && ((uintptr_t)out & 15) == 0, "fit_slice_rows: null or unaligned pointer");
  if (nslice == 6)
    k_fit_slice<6, false><<<(unsigned)((nrow + 31) / 32), 32 * (N / 16), 0, (cudaStream_t)stream>>>(
        out, ld_out, kp, row_exp, x, nullptr, nullptr, nullptr, 0., nullptr, nrow, N);
  else
    k_fit_slice<4, false><<<(unsigned)((nrow + 31) / 32), 32 * (N / 16), 0, (cudaStream_t)stream>>>(
        out, ld_out, kp, row_exp, x, nullptr, nullptr, nullptr, 0., nullptr, nrow, N);
  DPB_CUDA(cudaGetLastError());
  note_launches(1);
  return DPB200_OK;
}

int dpb200_fit_head_f64(double* e_out, signed char* out, long long ld_out, int kp, int* row_exp, const double* t,
                        const double* y, const double* w_head, const double* idt, double b_head, long long nrow,
                        int N, int nslice, dpb200_stream_t stream) {
  using namespace dpb200;
  DPB_REQUIRE(nslice == 6 || nslice == 4, "fit_head: built for 6 or 4 operand slices");
  DPB_REQUIRE(nrow >= 0 && N >= 16 && N % 16 == 0 && N <= 512 && kp == N && ld_out >= (long long)nslice * kp &&
                  ld_out % 16 == 0,
              "fit_head: N must be a multiple of 16 up to 512 and kp == N");
  if (nrow == 0) return DPB200_OK;
  DPB_REQUIRE(e_out && out && row_exp && t && y && w_head && ((uintptr_t)out & 15) == 0,
              "fit_head: null or unaligned pointer");
  if (nslice == 6)
    k_fit_slice<6, true><<<(unsigned)((nrow + 31) / 32), 32 * (N / 16), 0, (cudaStream_t)stream>>>(
        out, ld_out, kp, row_exp, t, y, w_head, idt, b_head, e_out, nrow, N);
  else
    k_fit_slice<4, true><<<(unsigned)((nrow + 31) / 32), 32 * (N / 16), 0, (cudaStream_t)stream>>>(
        out, ld_out, kp, row_exp, t, y, w_head, idt, b_head, e_out, nrow, N);
  DPB_CUDA(cudaGetLastError());
  note_launches(1);
  return DPB200_OK;
}

int dpb200_fit_slice_cols_f64(signed char* out, long long ld_out, long long slice_stride, int col0, int width,
                              int nslice, const int* row_exp, const double* src, int src_ld, const int* idx,
                              long long nrow, dpb200_stream_t stream) {
  using namespace dpb200;
  DPB_REQUIRE(nslice == 6 || nslice == 4, "fit_slice_cols: built for 6 or 4 operand slices");
  DPB_REQUIRE(nrow >= 0 && width >= 0 && col0 >= 0 && col0 + width <= slice_stride &&
                  ld_out >= (long long)nslice * slice_stride && src_ld >= width,
              "fit_slice_cols: the columns must lie inside one slice of the row");
  if (nrow == 0 || width == 0) return DPB200_OK;
  DPB_REQUIRE(out && row_exp && src, "fit_slice_cols: null pointer");
  long long grid = (nrow * width + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (grid > cap) grid = cap;
  if (nslice == 6)
    k_fit_slice_cols<6><<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(out, ld_out, slice_stride, col0, width, row_exp,
                                                                          src, src_ld, idx, nrow);
  else
    k_fit_slice_cols<4><<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(out, ld_out, slice_stride, col0, width, row_exp,
                                                                          src, src_ld, idx, nrow);
  DPB_CUDA(cudaGetLastError());
  note_launches(1);
  return DPB200_OK;
}

int dpb200_fit_blocked_f64(double* dst, const double* src, long long ld, long long nrow, int N, int to_blocked,
                           dpb200_stream_t stream) {
  using namespace dpb200;
  DPB_REQUIRE(nrow >= 0 && N >= 1 && ld >= N, "fit_blocked: bad shape");
  if (nrow == 0) return DPB200_OK;
  DPB_REQUIRE(dst && src, "fit_blocked: null pointer");
  const long long tot = ((nrow + kTileM - 1) / kTileM) * kTileM * N;
  int grid = ceil_div(tot, 256);
  const int cap = sm_count() * 16;
  if (grid > cap) grid = cap;
  k_fit_to_blocked<<<grid, 256, 0, (cudaStream_t)stream>>>(dst, src, ld, nrow, N, to_blocked);
  DPB_CUDA(cudaGetLastError());
  note_launches(1);
  return DPB200_OK;
}

}  // extern "C"
