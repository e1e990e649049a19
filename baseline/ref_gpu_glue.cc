// ON-BOX GPU COMPARATOR ONLY (see baseline/Makefile).  C-callable glue over the UNMODIFIED reference CUDA library:
// each entry point adapts flat device arrays to the C++ signatures of
//   source/lib/include/prod_env_mat.h:91-110  (prod_env_mat_a_gpu)
//   source/lib/include/tabulate.h:175-202     (tabulate_fusion_se_a_gpu, tabulate_fusion_se_a_grad_gpu)
//   source/lib/include/prod_force.h:71-79     (prod_force_a_gpu)
//   source/lib/include/prod_virial.h:30-39    (prod_virial_a_gpu)
// No arithmetic happens here.
#include <string>
#include <vector>

#include "neighbor_list.h"
#include "prod_env_mat.h"
#include "prod_force.h"
#include "prod_virial.h"
#include "tabulate.h"

namespace {
thread_local std::string g_err;
template <typename F>
int guarded(F&& f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}
}  // namespace

extern "C" {
__attribute__((visibility("default"))) const char* refgpu_last_error() { return g_err.c_str(); }

// ilist / numneigh / firstneigh: DEVICE arrays (firstneigh = device array of device row pointers), as the reference's
// op layer hands them over after convert_nlist_gpu_device
__attribute__((visibility("default"))) int refgpu_prod_env_mat_a_f64(
    double* em, double* em_deriv, double* rij, int* nlist, const double* coord, const int* type, int* ilist,
    int* numneigh, int** firstneigh, int* array_int, unsigned long long* array_longlong, int max_nbor_size,
    const double* avg, const double* std_, int nloc, int nall, float rcut, float rcut_smth, const int* sec, int nsec) {
  return guarded([&] {
    deepmd::InputNlist inlist(nloc, ilist, numneigh, firstneigh);
    deepmd::prod_env_mat_a_gpu<double>(em, em_deriv, rij, nlist, coord, type, inlist, array_int, array_longlong,
                                       max_nbor_size, avg, std_, nloc, nall, 1, rcut, rcut_smth,
                                       std::vector<int>(sec, sec + nsec));
  });
}
__attribute__((visibility("default"))) int refgpu_tabulate_fusion_se_a_f64(double* out, const double* table,
                                                                           const double* info_host, const double* em_x,
                                                                           const double* em, int nloc, int nnei, int M) {
  return guarded([&] { deepmd::tabulate_fusion_se_a_gpu<double>(out, table, info_host, em_x, em, nullptr, nloc, nnei, M); });
}
__attribute__((visibility("default"))) int refgpu_tabulate_fusion_se_a_grad_f64(double* dy_dem_x, double* dy_dem,
                                                                                const double* table,
                                                                                const double* info_host,
                                                                                const double* em_x, const double* em,
                                                                                const double* dy, int nloc, int nnei,
                                                                                int M) {
  return guarded([&] {
    deepmd::tabulate_fusion_se_a_grad_gpu<double>(dy_dem_x, dy_dem, nullptr, table, info_host, em_x, em, nullptr, dy, nloc,
                                                  nnei, M);
  });
}
__attribute__((visibility("default"))) int refgpu_prod_force_a_f64(double* force, const double* net_deriv,
                                                                   const double* in_deriv, const int* nlist, int nloc,
                                                                   int nall, int nnei) {
  return guarded([&] { deepmd::prod_force_a_gpu<double>(force, net_deriv, in_deriv, nlist, nloc, nall, nnei, 1); });
}
__attribute__((visibility("default"))) int refgpu_prod_virial_a_f64(double* virial, double* atom_virial,
                                                                    const double* net_deriv, const double* in_deriv,
                                                                    const double* rij, const int* nlist, int nloc,
                                                                    int nall, int nnei) {
  return guarded([&] { deepmd::prod_virial_a_gpu<double>(virial, atom_virial, net_deriv, in_deriv, rij, nlist, nloc, nall, nnei); });
}
}
