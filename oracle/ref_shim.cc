// TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// C-callable glue over the *unmodified* reference CPU library.  This file is
// compiled together with the reference's own sources where they lie under
// /root/reference/source/lib/src (recipe: oracle/Makefile, output:
// oracle/_ref/libdeepmd_ref.so).  Nothing here implements any arithmetic: each
// entry point only adapts flat C arrays to the C++ signatures declared in
//   source/lib/include/prod_env_mat.h:10-27   (prod_env_mat_a_cpu)
//   source/lib/include/fmt_nlist.h:11-19,86-93 (format_nlist_cpu, format_nlist_i_cpu)
//   source/lib/include/env_mat.h              (env_mat_a_cpu)
//   source/lib/include/tabulate.h:28-68       (tabulate_fusion_se_a{,_grad,_grad_grad}_cpu)
//   source/lib/include/prod_force.h:19-27     (prod_force_a_cpu)
//   source/lib/include/prod_virial.h:6-15     (prod_virial_a_cpu)
//   source/lib/include/prod_force_grad.h:8-15, prod_virial_grad.h:8-15 (prod_{force,virial}_grad_a_cpu)
//   source/lib/include/neighbor_list.h:166-176,301-352 (build_nlist_cpu, legacy build_nlist / copy_coord)
//   source/lib/include/coord.h:9-46           (normalize_coord_cpu, copy_coord_cpu, compute_cell_info)
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "SimulationRegion.h"
#include "coord.h"
#include "env_mat.h"
#include "fmt_nlist.h"
#include "neighbor_list.h"
#include "prod_env_mat.h"
#include "prod_force.h"
#include "prod_force_grad.h"
#include "prod_virial.h"
#include "prod_virial_grad.h"
#include "region.h"
#include "tabulate.h"

namespace {

thread_local std::string g_err;

// Raw neighbour rows arrive as CSR: row r = neigh[offsets[r] .. offsets[r+1]).
struct CsrList {
  std::vector<int> ilist, numneigh;
  std::vector<int*> first;
  deepmd::InputNlist view;
  CsrList(int inum, const int* ilist_in, const int64_t* offsets, const int* neigh)
      : ilist(inum), numneigh(inum), first(inum) {
    for (int r = 0; r < inum; ++r) {
      ilist[r] = ilist_in ? ilist_in[r] : r;
      numneigh[r] = static_cast<int>(offsets[r + 1] - offsets[r]);
      first[r] = const_cast<int*>(neigh) + offsets[r];
    }
    view = deepmd::InputNlist(inum, ilist.data(), numneigh.data(), first.data());
  }
};

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

template <typename FP>
int prod_env_mat_a(FP* em, FP* em_deriv, FP* rij, int* nlist, const FP* coord,
                   const int* type, const int* f_type, int inum, const int* ilist,
                   const int64_t* offsets, const int* neigh, const FP* avg,
                   const FP* std_, int nloc, int nall, int nframes, float rcut,
                   float rcut_smth, const int* sec, int nsec) {
  return guarded([&] {
    CsrList l(inum, ilist, offsets, neigh);
    int mx = deepmd::max_numneigh(l.view);
    deepmd::prod_env_mat_a_cpu<FP>(em, em_deriv, rij, nlist, coord, type, l.view, mx,
                                   avg, std_, nloc, nall, nframes, rcut, rcut_smth,
                                   std::vector<int>(sec, sec + nsec), f_type);
  });
}

template <typename FP>
int format_nlist(int* nlist, int* overflow, const FP* coord, const int* type,
                 int inum, const int* ilist, const int64_t* offsets,
                 const int* neigh, int nall, float rcut, const int* sec, int nsec) {
  return guarded([&] {
    std::vector<FP> posi(coord, coord + static_cast<size_t>(nall) * 3);
    std::vector<int> ty(type, type + nall), secv(sec, sec + nsec), fmt;
    const int nnei = secv.back();
    for (int r = 0; r < inum; ++r) {
      const int i = ilist ? ilist[r] : r;
      std::vector<int> row(neigh + offsets[r], neigh + offsets[r + 1]);
      int ret = format_nlist_i_cpu<FP>(fmt, posi, ty, i, row, rcut, secv);
      if (overflow) overflow[i] = ret;
      std::memcpy(nlist + static_cast<size_t>(i) * nnei, fmt.data(), sizeof(int) * nnei);
    }
  });
}

template <typename FP>
int env_mat_a_rows(FP* em, FP* em_deriv, FP* rij, const FP* coord, const int* type,
                   const int* fmt_nlist, int nloc, int nall, float rmin, float rmax,
                   const int* sec, int nsec) {
  return guarded([&] {
    std::vector<FP> posi(coord, coord + static_cast<size_t>(nall) * 3);
    std::vector<int> ty(type, type + nall), secv(sec, sec + nsec);
    const int nnei = secv.back();
    std::vector<FP> a, b, c;
    for (int i = 0; i < nloc; ++i) {
      std::vector<int> row(fmt_nlist + static_cast<size_t>(i) * nnei,
                           fmt_nlist + static_cast<size_t>(i + 1) * nnei);
      deepmd::env_mat_a_cpu<FP>(a, b, c, posi, ty, i, row, secv, rmin, rmax);
      std::memcpy(em + static_cast<size_t>(i) * nnei * 4, a.data(), sizeof(FP) * nnei * 4);
      std::memcpy(em_deriv + static_cast<size_t>(i) * nnei * 12, b.data(), sizeof(FP) * nnei * 12);
      std::memcpy(rij + static_cast<size_t>(i) * nnei * 3, c.data(), sizeof(FP) * nnei * 3);
    }
  });
}

template <typename FP>
int build_nlist_bruteforce(int* numneigh, int* rows, int* max_list_size,
                           const FP* coord, int nloc, int nall, int mem_size,
                           float rcut, const int* type) {
  int status = 0;
  int rc = guarded([&] {
    std::vector<int> ilist(nloc);
    std::vector<int*> first(nloc);
    for (int i = 0; i < nloc; ++i) first[i] = rows + static_cast<size_t>(i) * mem_size;
    deepmd::InputNlist l(nloc, ilist.data(), numneigh, first.data());
    status = deepmd::build_nlist_cpu<FP>(l, max_list_size, coord, nloc, nall, mem_size,
                                         rcut, 1, type);
  });
  return rc ? rc : status;
}

template <typename FP>
int normalize_coord(FP* coord, int natom, const FP* boxt) {
  return guarded([&] {
    deepmd::Region<FP> region;
    deepmd::init_region_cpu<FP>(region, boxt);
    deepmd::normalize_coord_cpu<FP>(coord, natom, region);
  });
}

template <typename FP>
int copy_coord_new(FP* out_c, int* out_t, int* mapping, int* nall, const FP* in_c,
                   const int* in_t, int nloc, int mem_nall, float rcut, const FP* boxt) {
  int status = 0;
  int rc = guarded([&] {
    deepmd::Region<FP> region;
    deepmd::init_region_cpu<FP>(region, boxt);
    status = deepmd::copy_coord_cpu<FP>(out_c, out_t, mapping, nall, in_c, in_t, nloc,
                                        mem_nall, rcut, region);
  });
  return rc ? rc : status;
}

}  // namespace

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

#define INST_FP(SUF, FP)                                                                   \
  int ref_prod_env_mat_a_##SUF(FP* em, FP* em_deriv, FP* rij, int* nlist, const FP* coord, \
                               const int* type, const int* f_type, int inum,               \
                               const int* ilist, const int64_t* offsets, const int* neigh, \
                               const FP* avg, const FP* std_, int nloc, int nall,          \
                               int nframes, float rcut, float rcut_smth, const int* sec,   \
                               int nsec) {                                                 \
    return prod_env_mat_a<FP>(em, em_deriv, rij, nlist, coord, type, f_type, inum, ilist,  \
                              offsets, neigh, avg, std_, nloc, nall, nframes, rcut,        \
                              rcut_smth, sec, nsec);                                       \
  }                                                                                        \
  int ref_format_nlist_##SUF(int* nlist, int* overflow, const FP* coord, const int* type,  \
                             int inum, const int* ilist, const int64_t* offsets,           \
                             const int* neigh, int nall, float rcut, const int* sec,       \
                             int nsec) {                                                   \
    return format_nlist<FP>(nlist, overflow, coord, type, inum, ilist, offsets, neigh,     \
                            nall, rcut, sec, nsec);                                        \
  }                                                                                        \
  int ref_env_mat_a_##SUF(FP* em, FP* em_deriv, FP* rij, const FP* coord, const int* type, \
                          const int* fmt_nlist, int nloc, int nall, float rmin,            \
                          float rmax, const int* sec, int nsec) {                          \
    return env_mat_a_rows<FP>(em, em_deriv, rij, coord, type, fmt_nlist, nloc, nall, rmin, \
                              rmax, sec, nsec);                                            \
  }                                                                                        \
  int ref_tabulate_fusion_se_a_##SUF(FP* out, const FP* table, const FP* info,             \
                                     const FP* em_x, const FP* em, const FP* two_embed,    \
                                     int nloc, int nnei, int M, int is_sorted) {           \
    return guarded([&] {                                                                   \
      deepmd::tabulate_fusion_se_a_cpu<FP>(out, table, info, em_x, em, two_embed, nloc,    \
                                           nnei, M, is_sorted != 0);                       \
    });                                                                                    \
  }                                                                                        \
  int ref_tabulate_fusion_se_a_grad_##SUF(FP* dy_dem_x, FP* dy_dem, FP* dy_dtwo,           \
                                          const FP* table, const FP* info,                 \
                                          const FP* em_x, const FP* em,                    \
                                          const FP* two_embed, const FP* dy, int nloc,     \
                                          int nnei, int M, int is_sorted) {                \
    return guarded([&] {                                                                   \
      deepmd::tabulate_fusion_se_a_grad_cpu<FP>(dy_dem_x, dy_dem, dy_dtwo, table, info,    \
                                                em_x, em, two_embed, dy, nloc, nnei, M,    \
                                                is_sorted != 0);                           \
    });                                                                                    \
  }                                                                                        \
  int ref_tabulate_fusion_se_a_grad_grad_##SUF(                                            \
      FP* dz_dy, const FP* table, const FP* info, const FP* em_x, const FP* em,            \
      const FP* two_embed, const FP* dz_dy_dem_x, const FP* dz_dy_dem,                     \
      const FP* dz_dy_dtwo, int nloc, int nnei, int M, int is_sorted) {                    \
    return guarded([&] {                                                                   \
      deepmd::tabulate_fusion_se_a_grad_grad_cpu<FP>(dz_dy, table, info, em_x, em,         \
                                                     two_embed, dz_dy_dem_x, dz_dy_dem,    \
                                                     dz_dy_dtwo, nloc, nnei, M,            \
                                                     is_sorted != 0);                      \
    });                                                                                    \
  }                                                                                        \
  int ref_tabulate_fusion_se_a_nd_##SUF(FP* out, const FP* table, const FP* info,          \
                                        const FP* em_x, const FP* em, const FP* two_embed, \
                                        int nloc, int nnei, int M, int is_sorted, int nd) {\
    return guarded([&] {                                                                   \
      deepmd::tabulate_fusion_se_a_cpu<FP>(out, table, info, em_x, em, two_embed, nloc,    \
                                           nnei, M, is_sorted != 0, nd);                   \
    });                                                                                    \
  }                                                                                        \
  int ref_tabulate_fusion_se_a_grad_nd_##SUF(FP* dy_dem_x, FP* dy_dem, FP* dy_dtwo,        \
                                             const FP* table, const FP* info,              \
                                             const FP* em_x, const FP* em,                 \
                                             const FP* two_embed, const FP* dy, int nloc,  \
                                             int nnei, int M, int is_sorted, int nd) {     \
    return guarded([&] {                                                                   \
      deepmd::tabulate_fusion_se_a_grad_cpu<FP>(dy_dem_x, dy_dem, dy_dtwo, table, info,    \
                                                em_x, em, two_embed, dy, nloc, nnei, M,    \
                                                is_sorted != 0, nd);                       \
    });                                                                                    \
  }                                                                                        \
  int ref_tabulate_fusion_se_a_grad_grad_nd_##SUF(                                         \
      FP* dz_dy, const FP* table, const FP* info, const FP* em_x, const FP* em,            \
      const FP* two_embed, const FP* dz_dy_dem_x, const FP* dz_dy_dem,                     \
      const FP* dz_dy_dtwo, int nloc, int nnei, int M, int is_sorted, int nd) {            \
    return guarded([&] {                                                                   \
      deepmd::tabulate_fusion_se_a_grad_grad_cpu<FP>(dz_dy, table, info, em_x, em,         \
                                                     two_embed, dz_dy_dem_x, dz_dy_dem,    \
                                                     dz_dy_dtwo, nloc, nnei, M,            \
                                                     is_sorted != 0, nd);                  \
    });                                                                                    \
  }                                                                                        \
  int ref_prod_force_a_##SUF(FP* force, const FP* net_deriv, const FP* in_deriv,           \
                             const int* nlist, int nloc, int nall, int nnei,               \
                             int nframes) {                                                \
    return guarded([&] {                                                                   \
      deepmd::prod_force_a_cpu<FP>(force, net_deriv, in_deriv, nlist, nloc, nall, nnei,    \
                                   nframes);                                               \
    });                                                                                    \
  }                                                                                        \
  int ref_prod_virial_a_##SUF(FP* virial, FP* atom_virial, const FP* net_deriv,            \
                              const FP* in_deriv, const FP* rij, const int* nlist,         \
                              int nloc, int nall, int nnei) {                              \
    return guarded([&] {                                                                   \
      deepmd::prod_virial_a_cpu<FP>(virial, atom_virial, net_deriv, in_deriv, rij, nlist,  \
                                    nloc, nall, nnei);                                     \
    });                                                                                    \
  }                                                                                        \
  int ref_prod_force_grad_a_##SUF(FP* grad_net, const FP* grad, const FP* in_deriv,        \
                                  const int* nlist, int nloc, int nnei, int nframes) {     \
    return guarded([&] {                                                                   \
      deepmd::prod_force_grad_a_cpu<FP>(grad_net, grad, in_deriv, nlist, nloc, nnei,       \
                                        nframes);                                          \
    });                                                                                    \
  }                                                                                        \
  int ref_prod_virial_grad_a_##SUF(FP* grad_net, const FP* grad, const FP* in_deriv,       \
                                   const FP* rij, const int* nlist, int nloc, int nnei) {  \
    return guarded([&] {                                                                   \
      deepmd::prod_virial_grad_a_cpu<FP>(grad_net, grad, in_deriv, rij, nlist, nloc,       \
                                         nnei);                                            \
    });                                                                                    \
  }                                                                                        \
  int ref_build_nlist_cpu_##SUF(int* numneigh, int* rows, int* max_list_size,              \
                                const FP* coord, int nloc, int nall, int mem_size,         \
                                float rcut, const int* type) {                             \
    return build_nlist_bruteforce<FP>(numneigh, rows, max_list_size, coord, nloc, nall,    \
                                      mem_size, rcut, type);                               \
  }                                                                                        \
  int ref_normalize_coord_##SUF(FP* coord, int natom, const FP* boxt) {                    \
    return normalize_coord<FP>(coord, natom, boxt);                                        \
  }                                                                                        \
  int ref_copy_coord_##SUF(FP* out_c, int* out_t, int* mapping, int* nall, const FP* in_c, \
                           const int* in_t, int nloc, int mem_nall, float rcut,            \
                           const FP* boxt) {                                               \
    return copy_coord_new<FP>(out_c, out_t, mapping, nall, in_c, in_t, nloc, mem_nall,     \
                              rcut, boxt);                                                 \
  }

INST_FP(f64, double)
INST_FP(f32, float)
#undef INST_FP

// compute_cell_info (coord.h:38-45): 23 ints.
int ref_compute_cell_info(int* cell_info, float rcut, const double* boxt) {
  return guarded([&] {
    deepmd::Region<double> region;
    deepmd::init_region_cpu<double>(region, boxt);
    deepmd::compute_cell_info<double>(cell_info, rcut, region);
  });
}

// Legacy fixture generators used by the reference's own lib tests
// (source/lib/tests/test_env_mat_a.cc:176-194): periodic image copy followed by
// the extended-grid cell-list build.  Outputs are returned through a handle so
// that the caller can size its buffers.
struct LegacySystem {
  std::vector<double> posi_cpy;
  std::vector<int> atype_cpy, mapping, ncell, ngcell;
  std::vector<std::vector<int>> nlist;
};

void* ref_legacy_copy_and_build(const double* posi, const int* atype, int nloc,
                                const double* box9, double rc_copy, double rc_list) {
  auto* sys = new LegacySystem();
  int rc = guarded([&] {
    SimulationRegion<double> region;
    region.reinitBox(box9);
    std::vector<double> p(posi, posi + static_cast<size_t>(nloc) * 3);
    std::vector<int> t(atype, atype + nloc);
    copy_coord(sys->posi_cpy, sys->atype_cpy, sys->mapping, sys->ncell, sys->ngcell, p, t,
               rc_copy, region);
    std::vector<int> nat_stt(3, 0), ext_stt(3), ext_end(3);
    for (int d = 0; d < 3; ++d) {
      ext_stt[d] = -sys->ngcell[d];
      ext_end[d] = sys->ncell[d] + sys->ngcell[d];
    }
    std::vector<std::vector<int>> nlist_r;
    build_nlist(sys->nlist, nlist_r, sys->posi_cpy, nloc, rc_list, rc_list, nat_stt,
                sys->ncell, ext_stt, ext_end, region, sys->ncell);
  });
  if (rc) {
    delete sys;
    return nullptr;
  }
  return sys;
}

int ref_legacy_nall(void* h) { return static_cast<LegacySystem*>(h)->atype_cpy.size(); }

int64_t ref_legacy_nnz(void* h) {
  int64_t n = 0;
  for (auto& r : static_cast<LegacySystem*>(h)->nlist) n += r.size();
  return n;
}

void ref_legacy_fetch(void* h, double* posi_cpy, int* atype_cpy, int* mapping, int* ncell3,
                      int* ngcell3, int64_t* offsets, int* neigh) {
  auto* s = static_cast<LegacySystem*>(h);
  std::memcpy(posi_cpy, s->posi_cpy.data(), sizeof(double) * s->posi_cpy.size());
  std::memcpy(atype_cpy, s->atype_cpy.data(), sizeof(int) * s->atype_cpy.size());
  std::memcpy(mapping, s->mapping.data(), sizeof(int) * s->mapping.size());
  for (int d = 0; d < 3; ++d) {
    ncell3[d] = s->ncell[d];
    ngcell3[d] = s->ngcell[d];
  }
  int64_t off = 0;
  for (size_t i = 0; i < s->nlist.size(); ++i) {
    offsets[i] = off;
    std::memcpy(neigh + off, s->nlist[i].data(), sizeof(int) * s->nlist[i].size());
    off += s->nlist[i].size();
  }
  offsets[s->nlist.size()] = off;
}

void ref_legacy_free(void* h) { delete static_cast<LegacySystem*>(h); }

}  // extern "C"
