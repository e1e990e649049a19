// Ghost-atom halo exchange helpers for the spatially decomposed multi-GPU path (sm_100a).
//
// The reference leaves this to LAMMPS (Comm::forward_comm / reverse_comm over MPI, host-staged
// unless CUDA-aware: source/lmp/pair_deepmd.cpp:482-488,1069-1093).  Here the wire transfer is one
// grouped NCCL send/recv issued from the host side (deepmd-kit_b200/domain.py); these kernels are
// the two device-side ends of it:
//   pack        : sendbuf[k] = coord[sendlist[k]] + shift[k]   (periodic image shift per entry)
//   unpack_add  : force[sendlist[k]] += recvbuf[k]             (ghost forces back to their owners;
//                 an atom can sit in several send directions, hence RED.ADD)
// Both are pure HBM gathers/scatters: one thread per (entry, component), grid sized to the entries.
#include "common.cuh"

namespace dpb200 {
namespace {

template <typename FP>
__global__ void k_halo_pack(FP* __restrict__ out, const FP* __restrict__ coord, const int* __restrict__ list,
                            const FP* __restrict__ shift, long long n3) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n3; e += (long long)gridDim.x * blockDim.x) {
    const long long k = e / 3;
    const int d = (int)(e - 3 * k);
    out[e] = coord[3 * (long long)list[k] + d] + shift[e];
  }
}

template <typename FP>
__global__ void k_halo_unpack_add(FP* __restrict__ force, const FP* __restrict__ in, const int* __restrict__ list,
                                  long long n3) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n3; e += (long long)gridDim.x * blockDim.x) {
    const long long k = e / 3;
    const int d = (int)(e - 3 * k);
    atomic_add(force + 3 * (long long)list[k] + d, in[e]);
  }
}

template <typename FP>
int halo_pack(FP* out, const FP* coord, const int* list, const FP* shift, int n, cudaStream_t st) {
  DPB_REQUIRE(n >= 0, "halo_pack: negative count");
  if (n == 0) return DPB200_OK;
  DPB_REQUIRE(out && coord && list && shift, "halo_pack: null pointer");
  const long long n3 = 3ll * n;
  int grid = ceil_div(n3, 256);
  const int cap = sm_count() * 16;
  if (grid > cap) grid = cap;
  k_halo_pack<FP><<<grid, 256, 0, st>>>(out, coord, list, shift, n3);
  DPB_CUDA(cudaGetLastError());
  note_launches(1);
  return DPB200_OK;
}

template <typename FP>
int halo_unpack_add(FP* force, const FP* in, const int* list, int n, cudaStream_t st) {
  DPB_REQUIRE(n >= 0, "halo_unpack_add: negative count");
  if (n == 0) return DPB200_OK;
  DPB_REQUIRE(force && in && list, "halo_unpack_add: null pointer");
  const long long n3 = 3ll * n;
  int grid = ceil_div(n3, 256);
  const int cap = sm_count() * 16;
  if (grid > cap) grid = cap;
  k_halo_unpack_add<FP><<<grid, 256, 0, st>>>(force, in, list, n3);
  DPB_CUDA(cudaGetLastError());
  note_launches(1);
  return DPB200_OK;
}

}  // namespace
}  // namespace dpb200

extern "C" {
#define DPB200_DEF_HALO(SUF, FP)                                                                       \
  int dpb200_halo_pack_##SUF(FP* sendbuf, const FP* coord, const int* sendlist, const FP* shift, int n, \
                             dpb200_stream_t stream) {                                                 \
    return dpb200::halo_pack<FP>(sendbuf, coord, sendlist, shift, n, (cudaStream_t)stream);            \
  }                                                                                                    \
  int dpb200_halo_unpack_add_##SUF(FP* force, const FP* recvbuf, const int* sendlist, int n,           \
                                   dpb200_stream_t stream) {                                           \
    return dpb200::halo_unpack_add<FP>(force, recvbuf, sendlist, n, (cudaStream_t)stream);             \
  }
DPB200_DEF_HALO(f64, double)
DPB200_DEF_HALO(f32, float)
#undef DPB200_DEF_HALO
}
