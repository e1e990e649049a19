// Reference-side binding: the C++ `deepmd::*_gpu` symbols of libdeepmd_op_cuda
// (source/lib/src/gpu/*.cu in the reference) re-exported on top of the dpb200 C ABI.
//
// This file is compiled AGAINST THE REFERENCE'S OWN HEADERS (-I $DEEPMD/source/lib/include
// -DGOOGLE_CUDA), so the declarations, default arguments, `InputNlist` / `Region` layouts and the
// exception types are the caller's, not copies.  Linking the result as `libdeepmd_op_cuda.so` in
// place of the reference's makes `source/op/tf/*_multi_device.cc`, `source/op/pt/
// tabulate_multi_device.cc`, DeepPot and the LAMMPS pair style run on the sm_100a kernels unchanged.
//
// Behavioural contract kept from the reference wrappers:
//  * synchronous: work is issued on the legacy default stream and the stream is synchronised
//    before returning (the reference brackets every kernel with cudaDeviceSynchronize);
//  * outputs need not be pre-zeroed;
//  * errors surface as deepmd::deepmd_exception / deepmd_exception_oom /
//    deepmd_exception_nlist_capacity (source/lib/include/errors.h:10-35).
// Scope: the se_a / se_atten path (ndescrpt == 4 hot path; 9 / 16 / 25 through csrc/tabulate_nd.cu).  Other descriptors' symbols (se_r, se_t, ...)
// are intentionally not provided here.
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "coord.h"
#include "dpb200.h"
#include "errors.h"
#include "fmt_nlist.h"
#include "neighbor_list.h"
#include "prod_env_mat.h"
#include "prod_force.h"
#include "prod_force_grad.h"
#include "prod_virial.h"
#include "prod_virial_grad.h"
#include "region.h"
#include "tabulate.h"

#define DPB_EXPORT __attribute__((visibility("default")))

namespace {

void check(int rc, const char* what) {
  if (rc >= 0) return;
  const std::string msg = std::string(what) + ": " + dpb200_last_error();
  if (rc == DPB200_ERR_OOM) throw deepmd::deepmd_exception_oom(msg);
  if (rc == DPB200_ERR_NLIST_CAPACITY) throw deepmd::deepmd_exception_nlist_capacity(msg);
  throw deepmd::deepmd_exception(msg);
}

void sync_default_stream(const char* what) {
  const cudaError_t e = cudaStreamSynchronize(0);
  if (e != cudaSuccess) {
    throw deepmd::deepmd_exception(std::string(what) + ": CUDA error " + cudaGetErrorString(e));
  }
}

// caller scratch (array_longlong: nloc*max_nbor_size*2 words by contract) or a temporary
struct Workspace {
  void* ptr = nullptr;
  bool owned = false;
  // Uses the caller's scratch when it is large enough once aligned up to 256 bytes (no device allocation, hence no
  // implicit device synchronisation per call); allocates otherwise.
  Workspace(void* scratch, size_t scratch_bytes, size_t need) {
    if (scratch != nullptr) {
      const uintptr_t base = reinterpret_cast<uintptr_t>(scratch);
      const uintptr_t aligned = (base + 255) & ~uintptr_t(255);
      if (aligned - base <= scratch_bytes && scratch_bytes - (aligned - base) >= need) {
        ptr = reinterpret_cast<void*>(aligned);
        return;
      }
    }
    if (cudaMalloc(&ptr, need) != cudaSuccess) {
      cudaGetLastError();
      throw deepmd::deepmd_exception_oom("dpb200 workspace allocation failed");
    }
    owned = true;
  }
  ~Workspace() {
    if (owned) cudaFree(ptr);
  }
};

template <typename FP>
struct Fn;
template <>
struct Fn<double> {
  static constexpr auto env = dpb200_prod_env_mat_a_f64;
  static constexpr auto fmt = dpb200_format_nlist_f64;
  static constexpr auto tab = dpb200_tabulate_fusion_se_a_f64;
  static constexpr auto tab_grad = dpb200_tabulate_fusion_se_a_grad_f64;
  static constexpr auto tab_gg = dpb200_tabulate_fusion_se_a_grad_grad_f64;
  static constexpr auto tab_nd = dpb200_tabulate_fusion_se_a_nd_f64;
  static constexpr auto tab_grad_nd = dpb200_tabulate_fusion_se_a_grad_nd_f64;
  static constexpr auto tab_gg_nd = dpb200_tabulate_fusion_se_a_grad_grad_nd_f64;
  static constexpr auto force = dpb200_prod_force_a_f64;
  static constexpr auto virial = dpb200_prod_virial_a_f64;
  static constexpr auto force_grad = dpb200_prod_force_grad_a_f64;
  static constexpr auto virial_grad = dpb200_prod_virial_grad_a_f64;
  static constexpr auto normalize = dpb200_normalize_coord_f64;
  static constexpr auto copy_coord = dpb200_copy_coord_cells_f64;
  static constexpr auto build = dpb200_build_nlist_f64;
};
template <>
struct Fn<float> {
  static constexpr auto env = dpb200_prod_env_mat_a_f32;
  static constexpr auto fmt = dpb200_format_nlist_f32;
  static constexpr auto tab = dpb200_tabulate_fusion_se_a_f32;
  static constexpr auto tab_grad = dpb200_tabulate_fusion_se_a_grad_f32;
  static constexpr auto tab_gg = dpb200_tabulate_fusion_se_a_grad_grad_f32;
  static constexpr auto tab_nd = dpb200_tabulate_fusion_se_a_nd_f32;
  static constexpr auto tab_grad_nd = dpb200_tabulate_fusion_se_a_grad_nd_f32;
  static constexpr auto tab_gg_nd = dpb200_tabulate_fusion_se_a_grad_grad_nd_f32;
  static constexpr auto force = dpb200_prod_force_a_f32;
  static constexpr auto virial = dpb200_prod_virial_a_f32;
  static constexpr auto force_grad = dpb200_prod_force_grad_a_f32;
  static constexpr auto virial_grad = dpb200_prod_virial_grad_a_f32;
  static constexpr auto normalize = dpb200_normalize_coord_f32;
  static constexpr auto copy_coord = dpb200_copy_coord_cells_f32;
  static constexpr auto build = dpb200_build_nlist_f32;
};

// tabulate.h is_supported_se_a_basis_dimension: 4 (se_a / se_atten hot path), 9, 16, 25 (csrc/tabulate_nd.cu)
void need_basis_dimension(int ndescrpt) {
  if (ndescrpt != 4 && ndescrpt != 9 && ndescrpt != 16 && ndescrpt != 25) {
    throw deepmd::deepmd_exception("The environment basis dimension must be 4, 9, 16 or 25, got " +
                                   std::to_string(ndescrpt));
  }
}

}  // namespace

namespace deepmd {

template <typename FPTYPE>
DPB_EXPORT void prod_env_mat_a_gpu(FPTYPE* em, FPTYPE* em_deriv, FPTYPE* rij, int* nlist, const FPTYPE* coord,
                                   const int* type, const InputNlist& gpu_inlist, int* /*array_int*/,
                                   unsigned long long* array_longlong, const int max_nbor_size, const FPTYPE* avg,
                                   const FPTYPE* std, const int nloc, const int nall, const int nframes,
                                   const float rcut, const float rcut_smth, const std::vector<int> sec,
                                   const int* f_type) {
  const int nsec = (int)sec.size();
  const int nnei = nsec ? sec.back() : 0;
  const size_t need = dpb200_prod_env_mat_a_workspace_bytes(nsec - 1, nnei, nall, nframes, sizeof(FPTYPE));
  Workspace ws(array_longlong, (size_t)nframes * nloc * max_nbor_size * 2 * sizeof(unsigned long long), need);
  check(Fn<FPTYPE>::env(em, em_deriv, rij, nlist, coord, type, f_type, gpu_inlist.ilist, gpu_inlist.numneigh,
                        gpu_inlist.firstneigh, nullptr, 0, max_nbor_size, avg, std, nloc, nall, nframes, rcut,
                        rcut_smth, sec.data(), nsec, ws.ptr, need, nullptr),
        "prod_env_mat_a_gpu");
  sync_default_stream("prod_env_mat_a_gpu");
}

template <typename FPTYPE>
DPB_EXPORT void format_nbor_list_gpu(int* nlist, const FPTYPE* coord, const int* type,
                                     const deepmd::InputNlist& gpu_inlist, int* /*array_int*/,
                                     uint_64* array_longlong, const int max_nbor_size, const int nloc, const int nall,
                                     const int nframes, const float rcut, const std::vector<int> sec) {
  const int nsec = (int)sec.size();
  const int nnei = nsec ? sec.back() : 0;
  const size_t need = dpb200_prod_env_mat_a_workspace_bytes(nsec - 1, nnei, nall, nframes, sizeof(FPTYPE));
  Workspace ws(array_longlong, (size_t)nframes * nloc * max_nbor_size * 2 * sizeof(unsigned long long), need);
  check(Fn<FPTYPE>::fmt(nlist, coord, type, gpu_inlist.ilist, gpu_inlist.numneigh, gpu_inlist.firstneigh, nullptr, 0,
                        max_nbor_size, nloc, nall, nframes, rcut, sec.data(), nsec, ws.ptr, need, nullptr),
        "format_nbor_list_gpu");
  sync_default_stream("format_nbor_list_gpu");
}

template <typename FPTYPE>
DPB_EXPORT void tabulate_fusion_se_a_gpu(FPTYPE* out, const FPTYPE* table, const FPTYPE* table_info,
                                         const FPTYPE* em_x, const FPTYPE* em, const FPTYPE* two_embed, const int nloc,
                                         const int nnei, const int last_layer_size, const bool is_sorted,
                                         const int ndescrpt) {
  need_basis_dimension(ndescrpt);
  if (ndescrpt != 4)
    check(Fn<FPTYPE>::tab_nd(out, table, table_info, em_x, em, two_embed, nloc, nnei, last_layer_size, is_sorted,
                             ndescrpt, nullptr),
          "tabulate_fusion_se_a_gpu");
  else
    check(Fn<FPTYPE>::tab(out, table, table_info, em_x, em, two_embed, nloc, nnei, last_layer_size, is_sorted, nullptr),
          "tabulate_fusion_se_a_gpu");
  sync_default_stream("tabulate_fusion_se_a_gpu");
}

template <typename FPTYPE>
DPB_EXPORT void tabulate_fusion_se_a_grad_gpu(FPTYPE* dy_dem_x, FPTYPE* dy_dem, FPTYPE* dy_dtwo, const FPTYPE* table,
                                              const FPTYPE* table_info, const FPTYPE* em_x, const FPTYPE* em,
                                              const FPTYPE* two_embed, const FPTYPE* dy, const int nloc,
                                              const int nnei, const int last_layer_size, const bool is_sorted,
                                              const int ndescrpt) {
  need_basis_dimension(ndescrpt);
  if (ndescrpt != 4)
    check(Fn<FPTYPE>::tab_grad_nd(dy_dem_x, dy_dem, dy_dtwo, table, table_info, em_x, em, two_embed, dy, nloc, nnei,
                                  last_layer_size, is_sorted, ndescrpt, nullptr),
          "tabulate_fusion_se_a_grad_gpu");
  else
    check(Fn<FPTYPE>::tab_grad(dy_dem_x, dy_dem, dy_dtwo, table, table_info, em_x, em, two_embed, dy, nloc, nnei,
                               last_layer_size, is_sorted, nullptr),
          "tabulate_fusion_se_a_grad_gpu");
  sync_default_stream("tabulate_fusion_se_a_grad_gpu");
}

template <typename FPTYPE>
DPB_EXPORT void tabulate_fusion_se_a_grad_grad_gpu(FPTYPE* dz_dy, const FPTYPE* table, const FPTYPE* table_info,
                                                   const FPTYPE* em_x, const FPTYPE* em, const FPTYPE* two_embed,
                                                   const FPTYPE* dz_dy_dem_x, const FPTYPE* dz_dy_dem,
                                                   const FPTYPE* dz_dy_dtwo, const int nloc, const int nnei,
                                                   const int last_layer_size, const bool is_sorted,
                                                   const int ndescrpt) {
  need_basis_dimension(ndescrpt);
  if (ndescrpt != 4)
    check(Fn<FPTYPE>::tab_gg_nd(dz_dy, table, table_info, em_x, em, two_embed, dz_dy_dem_x, dz_dy_dem, dz_dy_dtwo,
                                nloc, nnei, last_layer_size, is_sorted, ndescrpt, nullptr),
          "tabulate_fusion_se_a_grad_grad_gpu");
  else
    check(Fn<FPTYPE>::tab_gg(dz_dy, table, table_info, em_x, em, two_embed, dz_dy_dem_x, dz_dy_dem, dz_dy_dtwo, nloc,
                             nnei, last_layer_size, is_sorted, nullptr),
          "tabulate_fusion_se_a_grad_grad_gpu");
  sync_default_stream("tabulate_fusion_se_a_grad_grad_gpu");
}

template <typename FPTYPE>
DPB_EXPORT void prod_force_a_gpu(FPTYPE* force, const FPTYPE* net_deriv, const FPTYPE* in_deriv, const int* nlist,
                                 const int nloc, const int nall, const int nnei, const int nframes) {
  check(Fn<FPTYPE>::force(force, net_deriv, in_deriv, nlist, nloc, nall, nnei, nframes, nullptr), "prod_force_a_gpu");
  sync_default_stream("prod_force_a_gpu");
}

template <typename FPTYPE>
DPB_EXPORT void prod_virial_a_gpu(FPTYPE* virial, FPTYPE* atom_virial, const FPTYPE* net_deriv,
                                  const FPTYPE* env_deriv, const FPTYPE* rij, const int* nlist, const int nloc,
                                  const int nall, const int nnei) {
  check(Fn<FPTYPE>::virial(virial, atom_virial, net_deriv, env_deriv, rij, nlist, nloc, nall, nnei, nullptr),
        "prod_virial_a_gpu");
  sync_default_stream("prod_virial_a_gpu");
}

// source/lib/include/prod_force_grad.h:26-33, prod_virial_grad.h:26-33
template <typename FPTYPE>
DPB_EXPORT void prod_force_grad_a_gpu(FPTYPE* grad_net, const FPTYPE* grad, const FPTYPE* env_deriv, const int* nlist,
                                      const int nloc, const int nnei, const int nframes) {
  check(Fn<FPTYPE>::force_grad(grad_net, grad, env_deriv, nlist, nloc, nnei, nframes, nullptr), "prod_force_grad_a_gpu");
  sync_default_stream("prod_force_grad_a_gpu");
}

template <typename FPTYPE>
DPB_EXPORT void prod_virial_grad_a_gpu(FPTYPE* grad_net, const FPTYPE* grad, const FPTYPE* env_deriv,
                                       const FPTYPE* rij, const int* nlist, const int nloc, const int nnei) {
  check(Fn<FPTYPE>::virial_grad(grad_net, grad, env_deriv, rij, nlist, nloc, nnei, nullptr), "prod_virial_grad_a_gpu");
  sync_default_stream("prod_virial_grad_a_gpu");
}

DPB_EXPORT void use_nlist_map(int* nlist, const int* nlist_map, const int nloc, const int nnei) {
  check(dpb200_use_nlist_map(nlist, nlist_map, nloc, nnei, nullptr), "use_nlist_map");
  sync_default_stream("use_nlist_map");
}

// Region<FPTYPE> handed to the *_gpu functions holds DEVICE pointers (coord.cu:334-346).
template <typename FPTYPE>
static void fetch_box(FPTYPE (&b)[9], const deepmd::Region<FPTYPE>& region) {
  if (cudaMemcpy(b, region.boxt, sizeof(FPTYPE) * 9, cudaMemcpyDeviceToHost) != cudaSuccess) {
    cudaGetLastError();
    throw deepmd::deepmd_exception("dpb200: cannot read the simulation cell from the device");
  }
}

template <typename FPTYPE>
DPB_EXPORT void normalize_coord_gpu(FPTYPE* coord, const int natom, const deepmd::Region<FPTYPE>& region) {
  FPTYPE b[9];
  fetch_box(b, region);
  check(Fn<FPTYPE>::normalize(coord, natom, b, nullptr), "normalize_coord_gpu");
  sync_default_stream("normalize_coord_gpu");
}

template <typename FPTYPE>
DPB_EXPORT int copy_coord_gpu(FPTYPE* out_c, int* out_t, int* mapping, int* nall, int* /*int_data*/,
                              const FPTYPE* in_c, const int* in_t, const int& nloc, const int& mem_nall,
                              const int& /*loc_cellnum*/, const int& /*total_cellnum*/, const int* cell_info,
                              const deepmd::Region<FPTYPE>& region) {
  FPTYPE b[9];
  fetch_box(b, region);
  // the cutoff is not an argument of the reference function: the caller hands over the cell grid it built with
  // compute_cell_info (coord.cc:68-108): ncell = cell_info[3..5], ngcell = cell_info[12..14].  cell_info is a
  // DEVICE array (prod_env_mat_multi_device.cc:2437-2450 uploads it; coord.cu reads it inside kernels only).
  int ci[23];
  if (cudaMemcpy(ci, cell_info, sizeof(ci), cudaMemcpyDeviceToHost) != cudaSuccess) {
    cudaGetLastError();
    throw deepmd::deepmd_exception("dpb200 copy_coord_gpu: cannot read cell_info from the device");
  }
  const size_t need = dpb200_copy_coord_workspace_bytes(nloc);
  Workspace ws(nullptr, 0, need);
  const int rc = Fn<FPTYPE>::copy_coord(out_c, out_t, mapping, nall, in_c, in_t, nloc, mem_nall, ci + 3, ci + 12, b,
                                        ws.ptr, need, nullptr);
  check(rc, "copy_coord_gpu");
  sync_default_stream("copy_coord_gpu");
  return rc;
}

template <typename FPTYPE>
DPB_EXPORT int build_nlist_gpu(InputNlist& nlist, int* max_list_size, int* nlist_data, const FPTYPE* c_cpy,
                               const int& nloc, const int& nall, const int& mem_size, const float& rcut,
                               const int& nframes, const int* type) {
  // Contract of source/lib/src/gpu/neighbor_list.cu:197-250: row r = frame * nloc + atom; its neighbours (ascending
  // index) go to the caller-owned row firstneigh[r], numneigh[r] and ilist[r] = atom are filled, nlist_data
  // (2 * nrows * mem_size ints) is scratch, a row capacity below nall is refused with 1.
  if (mem_size < nall) {
    return 1;
  }
  const long long nrows = (long long)nframes * nloc;
  const size_t need = dpb200_build_nlist_workspace_bytes(nall);
  Workspace ws(nullptr, 0, need);
  int max_nei = 0;
  for (int f = 0; f < nframes; ++f) {
    int frame_max = 0;
    const int rc = Fn<FPTYPE>::build(nlist.numneigh + (long long)f * nloc, nlist_data + (long long)f * nloc * mem_size,
                                     &frame_max, c_cpy + (long long)f * nall * 3, nloc, nall, mem_size, rcut,
                                     type ? type + (long long)f * nall : nullptr, ws.ptr, need, nullptr);
    check(rc, "build_nlist_gpu");
    if (rc != 0) {
      sync_default_stream("build_nlist_gpu");
      return rc;
    }
    max_nei = frame_max > max_nei ? frame_max : max_nei;
  }
  check(dpb200_scatter_nlist_rows(nlist.firstneigh, nlist.ilist, nlist_data, mem_size, nlist.numneigh, (int)nrows, nloc,
                                  nullptr),
        "build_nlist_gpu");
  nlist.inum = (int)nrows;
  *max_list_size = max_nei;
  sync_default_stream("build_nlist_gpu");
  return 0;
}

#define DPB_INSTANTIATE(FP)                                                                                          \
  template void prod_env_mat_a_gpu<FP>(FP*, FP*, FP*, int*, const FP*, const int*, const InputNlist&, int*,           \
                                       unsigned long long*, const int, const FP*, const FP*, const int, const int,   \
                                       const int, const float, const float, const std::vector<int>, const int*);     \
  template void format_nbor_list_gpu<FP>(int*, const FP*, const int*, const deepmd::InputNlist&, int*, uint_64*,      \
                                         const int, const int, const int, const int, const float,                    \
                                         const std::vector<int>);                                                    \
  template void tabulate_fusion_se_a_gpu<FP>(FP*, const FP*, const FP*, const FP*, const FP*, const FP*, const int,   \
                                             const int, const int, const bool, const int);                           \
  template void tabulate_fusion_se_a_grad_gpu<FP>(FP*, FP*, FP*, const FP*, const FP*, const FP*, const FP*,          \
                                                  const FP*, const FP*, const int, const int, const int, const bool, \
                                                  const int);                                                        \
  template void tabulate_fusion_se_a_grad_grad_gpu<FP>(FP*, const FP*, const FP*, const FP*, const FP*, const FP*,    \
                                                       const FP*, const FP*, const FP*, const int, const int,        \
                                                       const int, const bool, const int);                            \
  template void prod_force_a_gpu<FP>(FP*, const FP*, const FP*, const int*, const int, const int, const int,          \
                                     const int);                                                                     \
  template void prod_virial_a_gpu<FP>(FP*, FP*, const FP*, const FP*, const FP*, const int*, const int, const int,    \
                                      const int);                                                                    \
  template void prod_force_grad_a_gpu<FP>(FP*, const FP*, const FP*, const int*, const int, const int, const int);    \
  template void prod_virial_grad_a_gpu<FP>(FP*, const FP*, const FP*, const FP*, const int*, const int, const int);   \
  template void normalize_coord_gpu<FP>(FP*, const int, const deepmd::Region<FP>&);                                   \
  template int copy_coord_gpu<FP>(FP*, int*, int*, int*, int*, const FP*, const int*, const int&, const int&,         \
                                  const int&, const int&, const int*, const deepmd::Region<FP>&);                    \
  template int build_nlist_gpu<FP>(InputNlist&, int*, int*, const FP*, const int&, const int&, const int&,            \
                                   const float&, const int&, const int*);
DPB_INSTANTIATE(double)
DPB_INSTANTIATE(float)
#undef DPB_INSTANTIATE

}  // namespace deepmd
