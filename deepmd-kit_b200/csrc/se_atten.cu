// se_atten (DPA-1, strip mode, smooth type embedding, attn_layer 0) glue between prod_env_mat_a and the pair-indexed
// gate of the table kernels: deepmd/pt/model/descriptor/se_atten.py:916-926 (`tebd_idx` = centre type * (ntypes + 1) +
// neighbour type, padding = the extra type `ntypes`), :979 (`gg_t = tt_full[tebd_idx]`), :981-983 (`gg_t * sw` with the
// switch of the environment matrix, source/lib/include/switcher.h:61-84).  One pass over the formatted list writes,
// per (centre, slot): the type-pair row, the switch value, and sw'(r) / r -- the factor that turns dE/d(sw), produced
// by the gated backward, into the pair force along r_ij (dpb200_prod_force_virial_a_pair).
#include "common.cuh"

namespace dpb200 {
namespace {

template <typename FP>
__global__ void k_gate_scalars(int* __restrict__ pair, FP* __restrict__ sw_out, FP* __restrict__ dswr_out,
                               const int* __restrict__ nlist, const int* __restrict__ type, const FP* __restrict__ rij,
                               long long nloc, int nnei, int ntypes, FP rmin, FP rmax) {
  const long long n = nloc * nnei;
  const FP span = rmax - rmin;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const long long i = e / nnei;
    const int j = nlist[e];
    int ti = type[i];
    if (ti < 0 || ti > ntypes) ti = ntypes;
    int tj = ntypes;
    FP sw = (FP)0., dswr = (FP)0.;
    if (j >= 0) {
      tj = type[j];
      if (tj < 0 || tj > ntypes) tj = ntypes;
      const FP x = rij[3 * e], y = rij[3 * e + 1], z = rij[3 * e + 2];
      const FP r = sqrt(x * x + y * y + z * z);
      if (r < rmin) {
        sw = (FP)1.;
      } else if (r < rmax) {
        const FP uu = (r - rmin) / span;
        const FP q = (FP)-6. * uu * uu + (FP)15. * uu - (FP)10.;
        const FP u3 = uu * uu * uu;
        sw = u3 * q + (FP)1.;
        const FP dsw = ((FP)3. * uu * uu * q + u3 * ((FP)-12. * uu + (FP)15.)) / span;
        dswr = r > (FP)0. ? dsw / r : (FP)0.;
      }
    }
    pair[e] = ti * (ntypes + 1) + tj;
    sw_out[e] = sw;
    dswr_out[e] = dswr;
  }
}

template <typename FP>
int launch_gate_scalars(int* pair, FP* sw, FP* dswr, const int* nlist, const int* type, const FP* rij, int nloc,
                        int nnei, int ntypes, float rcut_smth, float rcut, cudaStream_t st) {
  DPB_REQUIRE(nloc >= 0 && nnei >= 0 && ntypes >= 1, "se_atten_gate_scalars: bad sizes");
  const long long n = (long long)nloc * nnei;
  if (n == 0) return DPB200_OK;
  DPB_REQUIRE(pair && sw && dswr && nlist && type && rij, "se_atten_gate_scalars: null pointer");
  long long grid = (n + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (grid > cap) grid = cap;
  k_gate_scalars<FP><<<(unsigned)grid, 256, 0, st>>>(pair, sw, dswr, nlist, type, rij, nloc, nnei, ntypes,
                                                     (FP)rcut_smth, (FP)rcut);
  DPB_CUDA(cudaGetLastError());
  note_launches(1);
  return DPB200_OK;
}

}  // namespace
}  // namespace dpb200

extern "C" {

int dpb200_se_atten_gate_scalars_f64(int* pair, double* sw, double* dsw_over_r, const int* nlist, const int* type,
                                     const double* rij, int nloc, int nnei, int ntypes, float rcut_smth, float rcut,
                                     dpb200_stream_t stream) {
  return dpb200::launch_gate_scalars<double>(pair, sw, dsw_over_r, nlist, type, rij, nloc, nnei, ntypes, rcut_smth,
                                             rcut, (cudaStream_t)stream);
}
int dpb200_se_atten_gate_scalars_f32(int* pair, float* sw, float* dsw_over_r, const int* nlist, const int* type,
                                     const float* rij, int nloc, int nnei, int ntypes, float rcut_smth, float rcut,
                                     dpb200_stream_t stream) {
  return dpb200::launch_gate_scalars<float>(pair, sw, dsw_over_r, nlist, type, rij, nloc, nnei, ntypes, rcut_smth,
                                            rcut, (cudaStream_t)stream);
}

}  // extern "C"
