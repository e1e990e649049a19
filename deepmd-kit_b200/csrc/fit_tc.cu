// Fitting-net GEMMs on the Blackwell tensor cores (tcgen05.mma kind::i8, accumulators in TMEM, operands
// staged by TMA) -- SURVEY 8f-1; algebra deepmd/pt/model/network/mlp.py (tanh, resnet_dt, skip connection),
// composition deepmd/pt/model/descriptor/se_a.py:843-850 + the energy fitting net; the reference's fused fp32
// analogue is source/op/pt/graph_fitting.cu:40-90,230-340.
//
// fp64 parity (1e-10) rules out a reduced-precision product, and B200's FP64 tensor pipe peaks at ~33-40
// TFLOP/s, so every GEMM of the net is made EXACT on the int8 tensor cores by operand splitting
// (csrc/fitting.cu has the digit convention):
//     x = 2^Ex sum_i X_i 2^(-6-7i),   w = 2^Ew sum_j W_j 2^(-6-7j),    X_i, W_j signed 7-bit digits
//     x.w = 2^(Ex+Ew-12) sum_d 2^(-7d) S_d,     S_d = sum_{i+j=d} X_i.W_j   (int32, error free),  d < NS
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
#include <mutex>

#include "common.cuh"

namespace dpb200 {
namespace {

constexpr int kTileM = 128;
constexpr int kChunkK = 64;  // bytes of K per pipeline stage = one SWIZZLE_64B row
constexpr int kStages = 2;
constexpr int kThreads = 192;

enum { EPI_FWD = 0, EPI_BWD = 1, EPI_PLAIN = 2 };

struct GemmParams {
  long long n;        // rows
  int N;              // output columns
  int nk;             // K chunks of 64
  int n_tiles;        // column tiles
  long long m_blocks; // row blocks of 128
  const int* row_exp; // per-row exponent of the A operand, or null -> row_exp_fixed
  int row_exp_fixed;
  const double* col_scale;  // [N] 2^(col_exp - 12)
  const double* bias;       // [N]             (FWD)
  const double* idt;        // [N] or null     (FWD: of this layer; BWD: of the layer below)
  const double* skip;       // blocked [n][N] or null (FWD: y_prev; BWD: g of the layer above)
  const double* skip_vec;   // [N] or null     (BWD of the last hidden layer: g = head weights)
  const double* t_in;       // blocked [n][N]  (BWD: tanh values of the layer below)
  double* out0;             // FWD: t (blocked); BWD: g (blocked, nullable); PLAIN: row-major
  double* out1;             // FWD: y (blocked); BWD: dz (blocked)
  signed char* slices_out;  // FWD: y as [n][NS][Kp_out] int8 or null
  long long ld_slices;
  int Kp_out;
  int out_exp;
  long long ld_out;  // PLAIN: leading dimension
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
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
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

// Instruction descriptor (cute::UMMA::InstrDescriptor): D = s32, A = B = signed int8, both K-major, M x N.
__host__ __device__ constexpr uint32_t idesc_i8(int M, int N) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ double pow2i(int e) { return __hiloint2double((1023 + e) << 20, 0); }

// NS slices of 16 fp64 values (fixed-point image q = round(v * 2^(P - E)) + bias) -> one 16-byte store per slice
template <int NS>
__device__ __forceinline__ void store_slices16(signed char* __restrict__ base, long long slice_stride,
                                               const double (&v)[16], double up) {
  long long bias = 0;
#pragma unroll
  for (int k = 0; k < NS; ++k) bias = bias * 128 + 64;
  unsigned long long q[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) q[j] = (unsigned long long)(__double2ll_rn(v[j] * up) + bias);
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    const int sh = 7 * (NS - 1 - s);
    unsigned w[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const unsigned d0 = (unsigned)(q[4 * g + 0] >> sh) & 127u;
      const unsigned d1 = (unsigned)(q[4 * g + 1] >> sh) & 127u;
      const unsigned d2 = (unsigned)(q[4 * g + 2] >> sh) & 127u;
      const unsigned d3 = (unsigned)(q[4 * g + 3] >> sh) & 127u;
      const unsigned pk = d0 | (d1 << 8) | (d2 << 16) | (d3 << 24);
      w[g] = ((pk | 0x80808080u) - 0x40404040u) ^ 0x80808080u;  // per byte: digit' - 64
    }
    *reinterpret_cast<uint4*>(base + s * slice_stride) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

extern __shared__ __align__(16) unsigned char fit_smem_raw[];

template <int NS, int NT>
struct SmemLayout {
  static constexpr int kABytes = kTileM * kChunkK;  // one slice of the A chunk
  static constexpr int kBBytes = NT * kChunkK;
  static constexpr int kStageBytes = NS * (kABytes + kBBytes);
  static constexpr int kBarOff = kStages * kStageBytes;
  static constexpr int kTotal = kBarOff + 128 + 1024;  // + slack to align the tiles to 1024 B
};

template <int NS, int NT, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
    k_fit_gemm(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ GemmParams p) {
  using L = SmemLayout<NS, NT>;
  static_assert(NS * NT <= 512 && NT % 16 == 0 && NT <= 256, "accumulators must fit in tensor memory");
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
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_tfull, 1);
    mbar_init(bar_tempty, 128);
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
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const long long total = p.m_blocks * p.n_tiles;
  const int nk = p.nk;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int mb = (int)(tile / p.n_tiles);
        const int nb = (int)(tile - (long long)mb * p.n_tiles);
        for (int kc = 0; kc < nk; ++kc) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t fb = bar_full + 8 * stage;
          mbar_expect_tx(fb, (uint32_t)L::kStageBytes);
          const uint32_t sa = smem0 + stage * L::kStageBytes;
          const uint32_t sb = sa + NS * L::kABytes;
#pragma unroll
          for (int s = 0; s < NS; ++s) {
            tma_load_3d(sa + s * L::kABytes, &tmA, kc * kChunkK, s, mb * kTileM, fb);
            tma_load_3d(sb + s * L::kBBytes, &tmB, kc * kChunkK, nb * NT, s, fb);
          }
          if (++stage == kStages) stage = 0, phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc = idesc_i8(kTileM, NT);
    int stage = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (long long tile = blockIdx.x; tile < total; tile += gridDim.x) {
      mbar_wait(bar_tempty, acc_phase ^ 1);
      tc_fence_after();
      uint32_t started = 0;
      for (int kc = 0; kc < nk; ++kc) {
        mbar_wait(bar_full + 8 * stage, phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem0 + stage * L::kStageBytes;
          const uint32_t sb = sa + NS * L::kABytes;
#pragma unroll
          for (int k = 0; k < kChunkK / 32; ++k) {
#pragma unroll
            for (int i = 0; i < NS; ++i) {
              const uint64_t ad = smem_desc_sw64(sa + i * L::kABytes + k * 32);
#pragma unroll
              for (int j = 0; j < NS - i; ++j) {
                const uint64_t bd = smem_desc_sw64(sb + j * L::kBBytes + k * 32);
                const int d = i + j;
                mma_i8(tmem_base + d * NT, ad, bd, idesc, (started >> d) & 1u);
                started |= 1u << d;
              }
            }
          }
          tc_commit(bar_empty + 8 * stage);
          if (kc == nk - 1) tc_commit(bar_tfull);
        }
        __syncwarp();
        if (++stage == kStages) stage = 0, phase ^= 1;
      }
      acc_phase ^= 1;
    }
  } else {
    // ---------------------------------------------------------------- epilogue (thread = row)
    const int quad = warp & 3;
    const int row_in_tile = quad * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    uint32_t acc_phase = 0;
    for (long long tile = blockIdx.x; tile < total; tile += gridDim.x) {
      const int mb = (int)(tile / p.n_tiles);
      const int nb = (int)(tile - (long long)mb * p.n_tiles);
      const long long r = (long long)mb * kTileM + row_in_tile;
      const bool row_ok = r < p.n;
      const double rs = pow2i(row_ok && p.row_exp ? p.row_exp[r] : p.row_exp_fixed);
      mbar_wait(bar_tfull, acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < NT; cc += 16) {
        const int c0 = nb * NT + cc;
        if (c0 >= p.N) break;  // (uniform)
        double v[16];
        {
          int a[16];
          tmem_ld16(lane_addr + (uint32_t)((NS - 1) * NT + cc), a);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = (double)a[j];
#pragma unroll
          for (int d = NS - 2; d >= 0; --d) {
            tmem_ld16(lane_addr + (uint32_t)(d * NT + cc), a);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = v[j] * 0.0078125 + (double)a[j];
          }
        }
        // blocked offset of (r, c0): ((mb * N + c) * 128 + row_in_tile)
        const long long boff = ((long long)mb * p.N + c0) * kTileM + row_in_tile;
        if (EPI == EPI_FWD) {
          double y[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int c = c0 + j;
            const bool ok = c < p.N;
            const int cs = ok ? c : p.N - 1;
            const double z = v[j] * rs * __ldg(p.col_scale + cs) + __ldg(p.bias + cs);
            const double t = tanh(z);
            double yy = p.idt ? t * __ldg(p.idt + cs) : t;
            if (p.skip && row_ok && ok) yy += p.skip[boff + (long long)j * kTileM];
            y[j] = ok ? yy : 0.;
            if (row_ok && ok) {
              p.out0[boff + (long long)j * kTileM] = t;
              p.out1[boff + (long long)j * kTileM] = yy;
            }
          }
          if (p.slices_out && row_ok) {
            // (columns >= N inside the padded K of the next layer are written as zero digits)
            const double up = pow2i(6 + 7 * (NS - 1) - p.out_exp);
            if (c0 + 16 <= p.Kp_out) store_slices16<NS>(p.slices_out + r * p.ld_slices + c0, p.Kp_out, y, up);
          }
        } else if (EPI == EPI_BWD) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int c = c0 + j;
            if (c < p.N && row_ok) {
              double g = v[j] * rs * __ldg(p.col_scale + c);
              if (p.skip) g += p.skip[boff + (long long)j * kTileM];
              if (p.skip_vec) g += __ldg(p.skip_vec + c);
              const double t = p.t_in[boff + (long long)j * kTileM];
              double dz = g * (1. - t * t);
              if (p.idt) dz *= __ldg(p.idt + c);
              if (p.out0) p.out0[boff + (long long)j * kTileM] = g;
              p.out1[boff + (long long)j * kTileM] = dz;
            }
          }
        } else {
          if (row_ok) {
            double* __restrict__ o = p.out0 + r * p.ld_out + c0;
            if (c0 + 16 <= p.N && (p.ld_out & 1) == 0) {
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                const double a0 = v[j] * rs * __ldg(p.col_scale + c0 + j);
                const double a1 = v[j + 1] * rs * __ldg(p.col_scale + c0 + j + 1);
                __stcs(reinterpret_cast<double2*>(o + j), make_double2(a0, a1));
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (c0 + j < p.N) o[j] = v[j] * rs * __ldg(p.col_scale + c0 + j);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar_tempty);
      acc_phase ^= 1;
    }
  }

  tc_fence_before();
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
template <int NS, bool HEAD>
__global__ void __launch_bounds__(128) k_fit_slice(signed char* __restrict__ out, long long ld_out, int Kp,
                                                   int* __restrict__ row_exp, const double* __restrict__ x,
                                                   const double* __restrict__ y, const double* __restrict__ w_head,
                                                   const double* __restrict__ idt, double b_head,
                                                   double* __restrict__ e_out, long long n, int N) {
  const long long mb = blockIdx.x;
  const long long r = mb * kTileM + threadIdx.x;
  if (r >= n) return;
  const double* __restrict__ xb = x + mb * (long long)N * kTileM + threadIdx.x;
  double m = 0., e = b_head;
  for (int c = 0; c < N; ++c) {
    double v = xb[(long long)c * kTileM];
    if (HEAD) {
      const double w = __ldg(w_head + c);
      e += y[mb * (long long)N * kTileM + (long long)c * kTileM + threadIdx.x] * w;
      v = w * (1. - v * v);
      if (idt) v *= __ldg(idt + c);
    }
    m = fmax(m, fabs(v));
  }
  if (HEAD) e_out[r] = e;
  int E = ((__double2hiint(m) >> 20) & 0x7ff) - 1023 + 2;
  E = E < -900 ? -900 : (E > 900 ? 900 : E);
  row_exp[r] = E;
  const double up = pow2i(6 + 7 * (NS - 1) - E);
  signed char* __restrict__ o = out + r * ld_out;
  for (int c0 = 0; c0 < Kp; c0 += 16) {
    double v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int c = c0 + j;
      double t = 0.;
      if (c < N) {
        t = xb[(long long)c * kTileM];
        if (HEAD) {
          t = __ldg(w_head + c) * (1. - t * t);
          if (idt) t *= __ldg(idt + c);
        }
      }
      v[j] = t;
    }
    store_slices16<NS>(o + c0, Kp, v, up);
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

template <int NS, int NT, int EPI>
int launch_gemm(const CUtensorMap& ma, const CUtensorMap& mb, const GemmParams& p, cudaStream_t st) {
  using L = SmemLayout<NS, NT>;
  static bool attr_done = false;
  if (!attr_done) {
    DPB_CUDA(cudaFuncSetAttribute(k_fit_gemm<NS, NT, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    attr_done = true;
  }
  const long long total = p.m_blocks * p.n_tiles;
  const int grid = (int)(total < sm_count() ? total : sm_count());
  k_fit_gemm<NS, NT, EPI><<<grid, kThreads, L::kTotal, st>>>(ma, mb, p);
  DPB_CUDA(cudaGetLastError());
  note_launches(1);
  return DPB200_OK;
}

}  // namespace
}  // namespace dpb200

extern "C" {

int dpb200_fit_gemm_i8_f64(int mode, long long nrow, int N, int K, int nslice, const signed char* a_slices,
                           long long a_slice_stride, long long a_row_stride, const int* row_exp, int row_exp_fixed,
                           const signed char* b_slices, int b_k_stride, const double* col_scale, const double* bias,
                           const double* idt, const double* skip, const double* skip_vec, const double* t_in,
                           double* out0, double* out1, long long ld_out, signed char* slices_out,
                           long long ld_slices, int kp_out, int out_exp, dpb200_stream_t stream) {
  using namespace dpb200;
  DPB_REQUIRE(mode >= 0 && mode <= 2, "fit_gemm: mode must be 0 (forward), 1 (backward) or 2 (plain)");
  DPB_REQUIRE(nslice == 6, "fit_gemm: only 6 operand slices are built");
  DPB_REQUIRE(nrow >= 0 && N >= 1 && K >= 1, "fit_gemm: bad shape");
  if (nrow == 0) return DPB200_OK;
  DPB_REQUIRE(a_slices && b_slices && col_scale && (out0 || mode == EPI_BWD), "fit_gemm: null pointer");
  DPB_REQUIRE(K % 16 == 0 && a_slice_stride % 16 == 0 && a_row_stride % 16 == 0 && b_k_stride % 64 == 0 &&
                  b_k_stride >= K && ((uintptr_t)a_slices & 15) == 0 && ((uintptr_t)b_slices & 15) == 0,
              "fit_gemm: operands must be 16-byte aligned with K a multiple of 16 and the weight rows padded to 64");
  DPB_REQUIRE(mode != EPI_FWD || (bias && out1), "fit_gemm: forward needs bias and both outputs");
  DPB_REQUIRE(mode != EPI_BWD || (t_in && out1), "fit_gemm: backward needs t_in and the dz output");
  DPB_REQUIRE(!slices_out || (kp_out % 16 == 0 && kp_out >= N && ld_slices >= (long long)nslice * kp_out &&
                              ((uintptr_t)slices_out & 15) == 0 && ld_slices % 16 == 0),
              "fit_gemm: bad slice output layout");
  constexpr int NS = 6, NT = 80;
  GemmParams p;
  p.n = nrow;
  p.N = N;
  p.nk = (K + kChunkK - 1) / kChunkK;
  p.n_tiles = (N + NT - 1) / NT;
  p.m_blocks = (nrow + kTileM - 1) / kTileM;
  p.row_exp = row_exp;
  p.row_exp_fixed = row_exp_fixed;
  p.col_scale = col_scale;
  p.bias = bias;
  p.idt = idt;
  p.skip = skip;
  p.skip_vec = skip_vec;
  p.t_in = t_in;
  p.out0 = out0;
  p.out1 = out1;
  p.slices_out = slices_out;
  p.ld_slices = ld_slices;
  p.Kp_out = kp_out;
  p.out_exp = out_exp;
  p.ld_out = ld_out;
  CUtensorMap ma, mb;
  // A: [nrow][nslice][K] bytes (slice stride, row stride given); K beyond the tensor is zero-filled by TMA
  int rc = make_map(&ma, a_slices, (unsigned long long)K, (unsigned long long)nslice, (unsigned long long)nrow,
                    (unsigned long long)a_slice_stride, (unsigned long long)a_row_stride, 1, kTileM);
  if (rc != DPB200_OK) return rc;
  // B: [nslice][N][b_k_stride] bytes (weights transposed: row = output column, K contiguous, zero padded)
  rc = make_map(&mb, b_slices, (unsigned long long)b_k_stride, (unsigned long long)N, (unsigned long long)nslice,
                (unsigned long long)b_k_stride, (unsigned long long)b_k_stride * N, NT, 1);
  if (rc != DPB200_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == EPI_FWD) return launch_gemm<NS, NT, EPI_FWD>(ma, mb, p, st);
  if (mode == EPI_BWD) return launch_gemm<NS, NT, EPI_BWD>(ma, mb, p, st);
  return launch_gemm<NS, NT, EPI_PLAIN>(ma, mb, p, st);
}

int dpb200_fit_slice_rows_f64(signed char* out, long long ld_out, int kp, int* row_exp, const double* x,
                              long long nrow, int N, int nslice, dpb200_stream_t stream) {
  using namespace dpb200;
  DPB_REQUIRE(nslice == 6, "fit_slice_rows: only 6 operand slices are built");
  DPB_REQUIRE(nrow >= 0 && N >= 1 && kp % 16 == 0 && kp >= N && ld_out >= (long long)nslice * kp && ld_out % 16 == 0,
              "fit_slice_rows: bad shape");
  if (nrow == 0) return DPB200_OK;
  DPB_REQUIRE(out && row_exp && x && ((uintptr_t)out & 15) == 0, "fit_slice_rows: null or unaligned pointer");
  k_fit_slice<6, false><<<(unsigned)((nrow + kTileM - 1) / kTileM), kTileM, 0, (cudaStream_t)stream>>>(
      out, ld_out, kp, row_exp, x, nullptr, nullptr, nullptr, 0., nullptr, nrow, N);
  DPB_CUDA(cudaGetLastError());
  note_launches(1);
  return DPB200_OK;
}

int dpb200_fit_head_f64(double* e_out, signed char* out, long long ld_out, int kp, int* row_exp, const double* t,
                        const double* y, const double* w_head, const double* idt, double b_head, long long nrow,
                        int N, int nslice, dpb200_stream_t stream) {
  using namespace dpb200;
  DPB_REQUIRE(nslice == 6, "fit_head: only 6 operand slices are built");
  DPB_REQUIRE(nrow >= 0 && N >= 1 && kp % 16 == 0 && kp >= N && ld_out >= (long long)nslice * kp && ld_out % 16 == 0,
              "fit_head: bad shape");
  if (nrow == 0) return DPB200_OK;
  DPB_REQUIRE(e_out && out && row_exp && t && y && w_head && ((uintptr_t)out & 15) == 0,
              "fit_head: null or unaligned pointer");
  k_fit_slice<6, true><<<(unsigned)((nrow + kTileM - 1) / kTileM), kTileM, 0, (cudaStream_t)stream>>>(
      out, ld_out, kp, row_exp, t, y, w_head, idt, b_head, e_out, nrow, N);
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
