// Test driver for lib/libdeepmd_op_cuda.so (the reference-side C++ binding): calls the reference's
// own declarations deepmd::prod_env_mat_a_gpu / tabulate_fusion_se_a_gpu / _grad_gpu /
// prod_force_a_gpu / prod_virial_a_gpu / prod_force_grad_a_gpu / prod_virial_grad_a_gpu exactly as source/op/tf/*_multi_device.cc would, on inputs read
// from a flat binary file, and writes the outputs back.  Compiled against the reference headers
// (tests/shim/build.sh); the pytest side compares the outputs with the CPU oracle.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "coord.h"
#include "neighbor_list.h"
#include "prod_env_mat.h"
#include "prod_force.h"
#include "prod_force_grad.h"
#include "prod_virial.h"
#include "prod_virial_grad.h"
#include "region.h"
#include "tabulate.h"

#define CK(x)                                                                  \
  do {                                                                         \
    cudaError_t e = (x);                                                       \
    if (e != cudaSuccess) {                                                    \
      fprintf(stderr, "CUDA %s at %d\n", cudaGetErrorString(e), __LINE__);     \
      exit(2);                                                                 \
    }                                                                          \
  } while (0)

template <typename T>
std::vector<T> rd(FILE* f, size_t n) {
  std::vector<T> v(n);
  if (n && fread(v.data(), sizeof(T), n, f) != n) {
    fprintf(stderr, "short read\n");
    exit(3);
  }
  return v;
}
template <typename T>
T* up(const std::vector<T>& v) {
  T* d = nullptr;
  CK(cudaMalloc((void**)&d, sizeof(T) * (v.size() ? v.size() : 1)));
  CK(cudaMemcpy(d, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice));
  return d;
}
template <typename T>
void down(FILE* f, const T* d, size_t n) {
  std::vector<T> v(n);
  CK(cudaMemcpy(v.data(), d, sizeof(T) * n, cudaMemcpyDeviceToHost));
  fwrite(v.data(), sizeof(T), n, f);
}

// Second mode (`shim_driver nlist in out`): the neighbour front end exactly as _norm_copy_coord_gpu and
// _build_nlist_gpu of source/op/tf/prod_env_mat_multi_device.cc:2399-2600 drive it -- Region and cell_info live in
// DEVICE memory, the rows are written into the caller-owned jlist through firstneigh, ind_data is separate scratch.
// (init_region_cpu / compute_cell_info are host functions of the reference CPU library: oracle/_ref.)
static int nlist_mode(const char* in, const char* out) {
  FILE* fi = fopen(in, "rb");
  FILE* fo = fopen(out, "wb");
  if (!fi || !fo) return 1;
  auto hd = rd<int>(fi, 3);  // nloc, mem_cpy, nframes
  const int nloc = hd[0], mem_cpy = hd[1], nframes = hd[2];
  auto rc = rd<float>(fi, 1);
  auto box = rd<double>(fi, 9);
  auto coord = rd<double>(fi, (size_t)nloc * 3);
  auto type = rd<int>(fi, nloc);
  deepmd::Region<double> region;
  deepmd::init_region_cpu(region, box.data());
  std::vector<double> box_info(18);
  for (int k = 0; k < 9; ++k) box_info[k] = region.boxt[k], box_info[9 + k] = region.rec_boxt[k];
  std::vector<int> cell_info(23);
  deepmd::compute_cell_info(cell_info.data(), rc[0], region);
  const int loc_cellnum = cell_info[21], total_cellnum = cell_info[22];
  double* box_dev = up(box_info);
  std::vector<int> ints(23 + (size_t)nloc * 3 + loc_cellnum + (size_t)total_cellnum * 3 * 2 + loc_cellnum + 1 +
                        total_cellnum + 1 + nloc, 0);
  for (int k = 0; k < 23; ++k) ints[k] = cell_info[k];
  int* cell_info_dev = up(ints);
  deepmd::Region<double> region_dev(box_dev, box_dev + 9);
  double* tmp_coord = up(coord);
  int* d_type = up(type);
  double* coord_cpy;
  int* type_cpy;
  CK(cudaMalloc((void**)&coord_cpy, sizeof(double) * (size_t)mem_cpy * 3));
  CK(cudaMalloc((void**)&type_cpy, sizeof(int) * (size_t)mem_cpy * 2));
  int* idx_mapping = type_cpy + mem_cpy;
  int nall = nloc, ret_small = -1, ret = -1, ret_cap = -1, max_nnei = 0;
  std::vector<int> numneigh_h, ilist_h, jlist_h;
  int mem_nnei = 0;
  try {
    deepmd::normalize_coord_gpu(tmp_coord, nloc, region_dev);
    // a copy buffer that is too small must be reported with 1, not written past
    ret_small = deepmd::copy_coord_gpu(coord_cpy, type_cpy, idx_mapping, &nall, cell_info_dev + 23, tmp_coord, d_type, nloc,
                                       nloc + 1, loc_cellnum, total_cellnum, cell_info_dev, region_dev);
    ret = deepmd::copy_coord_gpu(coord_cpy, type_cpy, idx_mapping, &nall, cell_info_dev + 23, tmp_coord, d_type, nloc,
                                 mem_cpy, loc_cellnum, total_cellnum, cell_info_dev, region_dev);
    if (ret != 0) {
      fprintf(stderr, "copy_coord_gpu returned %d\n", ret);
      return 5;
    }
    // nframes identical frames of the extended system
    const long long nrows = (long long)nframes * nloc;
    double* c_frames;
    int* t_frames;
    CK(cudaMalloc((void**)&c_frames, sizeof(double) * (size_t)nframes * nall * 3));
    CK(cudaMalloc((void**)&t_frames, sizeof(int) * (size_t)nframes * nall));
    for (int f = 0; f < nframes; ++f) {
      CK(cudaMemcpy(c_frames + (size_t)f * nall * 3, coord_cpy, sizeof(double) * nall * 3, cudaMemcpyDeviceToDevice));
      CK(cudaMemcpy(t_frames + (size_t)f * nall, type_cpy, sizeof(int) * nall, cudaMemcpyDeviceToDevice));
    }
    mem_nnei = nall;
    int *ilist, *numneigh, *jlist;
    int** firstneigh;
    CK(cudaMalloc((void**)&ilist, sizeof(int) * nrows * 2));
    numneigh = ilist + nrows;
    CK(cudaMalloc((void**)&jlist, sizeof(int) * (size_t)nrows * mem_nnei * 3));
    CK(cudaMemset(jlist, 0xff, sizeof(int) * (size_t)nrows * mem_nnei * 3));
    int* ind_data = jlist + (size_t)nrows * mem_nnei;
    CK(cudaMalloc((void**)&firstneigh, sizeof(int*) * nrows));
    std::vector<int*> first(nrows);
    for (long long i = 0; i < nrows; ++i) first[i] = jlist + i * mem_nnei;
    CK(cudaMemcpy(firstneigh, first.data(), sizeof(int*) * nrows, cudaMemcpyHostToDevice));
    deepmd::InputNlist inlist((int)nrows, ilist, numneigh, firstneigh);
    ret_cap = deepmd::build_nlist_gpu(inlist, &max_nnei, ind_data, c_frames, nloc, nall, nall - 1, rc[0], nframes, t_frames);
    ret = deepmd::build_nlist_gpu(inlist, &max_nnei, ind_data, c_frames, nloc, nall, mem_nnei, rc[0], nframes, t_frames);
    if (ret != 0) {
      fprintf(stderr, "build_nlist_gpu returned %d\n", ret);
      return 5;
    }
    // the caller's firstneigh table must be untouched and the rows must be in jlist
    std::vector<int*> first_after(nrows);
    CK(cudaMemcpy(first_after.data(), firstneigh, sizeof(int*) * nrows, cudaMemcpyDeviceToHost));
    for (long long i = 0; i < nrows; ++i)
      if (first_after[i] != first[i]) {
        fprintf(stderr, "firstneigh[%lld] was overwritten\n", i);
        return 6;
      }
    numneigh_h.resize(nrows);
    ilist_h.resize(nrows);
    jlist_h.resize((size_t)nrows * mem_nnei);
    CK(cudaMemcpy(numneigh_h.data(), numneigh, sizeof(int) * nrows, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ilist_h.data(), ilist, sizeof(int) * nrows, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(jlist_h.data(), jlist, sizeof(int) * jlist_h.size(), cudaMemcpyDeviceToHost));
  } catch (const std::exception& e) {
    fprintf(stderr, "exception: %s\n", e.what());
    return 4;
  }
  int head[6] = {nall, ret_small, ret_cap, max_nnei, mem_nnei, nframes};
  fwrite(head, sizeof(int), 6, fo);
  down(fo, tmp_coord, (size_t)nloc * 3);
  down(fo, coord_cpy, (size_t)nall * 3);
  down(fo, type_cpy, (size_t)nall);
  down(fo, idx_mapping, (size_t)nall);
  fwrite(ilist_h.data(), sizeof(int), ilist_h.size(), fo);
  fwrite(numneigh_h.data(), sizeof(int), numneigh_h.size(), fo);
  // rows, compacted: numneigh entries each
  for (size_t i = 0; i < numneigh_h.size(); ++i) fwrite(jlist_h.data() + i * mem_nnei, sizeof(int), numneigh_h[i], fo);
  fclose(fo);
  printf("SHIM_DRIVER_OK\n");
  return 0;
}

// Third mode (`shim_driver tabnd in out`): tabulate_fusion_se_a{,_grad,_grad_grad}_gpu with ndescrpt = 9 / 16 / 25
// through the reference's own C++ signatures (source/lib/include/tabulate.h:175-218, the trailing `ndescrpt`).
static int tabnd_mode(const char* in, const char* out) {
  FILE* fi = fopen(in, "rb");
  FILE* fo = fopen(out, "wb");
  if (!fi || !fo) return 1;
  auto hd = rd<int>(fi, 6);  // nloc nnei M nd nspline is_sorted
  const int nloc = hd[0], nnei = hd[1], M = hd[2], nd = hd[3], nspline = hd[4];
  const bool is_sorted = hd[5] != 0;
  auto table = rd<double>(fi, (size_t)nspline * M * 6);
  auto info = rd<double>(fi, 6);
  auto em_x = rd<double>(fi, (size_t)nloc * nnei);
  auto em = rd<double>(fi, (size_t)nloc * nnei * nd);
  auto dy = rd<double>(fi, (size_t)nloc * nd * M);
  auto dzx = rd<double>(fi, (size_t)nloc * nnei);
  auto dzem = rd<double>(fi, (size_t)nloc * nnei * nd);
  double *d_table = up(table), *d_x = up(em_x), *d_em = up(em), *d_dy = up(dy), *d_dzx = up(dzx), *d_dzem = up(dzem);
  double *desc, *gx, *gem, *gg;
  CK(cudaMalloc((void**)&desc, sizeof(double) * nloc * nd * M));
  CK(cudaMalloc((void**)&gx, sizeof(double) * nloc * nnei));
  CK(cudaMalloc((void**)&gem, sizeof(double) * nloc * nnei * nd));
  CK(cudaMalloc((void**)&gg, sizeof(double) * nloc * nd * M));
  try {
    deepmd::tabulate_fusion_se_a_gpu<double>(desc, d_table, info.data(), d_x, d_em, nullptr, nloc, nnei, M, is_sorted, nd);
    deepmd::tabulate_fusion_se_a_grad_gpu<double>(gx, gem, nullptr, d_table, info.data(), d_x, d_em, nullptr, d_dy, nloc,
                                                  nnei, M, is_sorted, nd);
    deepmd::tabulate_fusion_se_a_grad_grad_gpu<double>(gg, d_table, info.data(), d_x, d_em, nullptr, d_dzx, d_dzem,
                                                       nullptr, nloc, nnei, M, is_sorted, nd);
    bool refused = false;
    try {  // 5 is not a supported basis dimension
      deepmd::tabulate_fusion_se_a_gpu<double>(desc, d_table, info.data(), d_x, d_em, nullptr, 0, nnei, M, is_sorted, 5);
    } catch (const std::exception&) {
      refused = true;
    }
    if (!refused) {
      fprintf(stderr, "ndescrpt = 5 was not refused\n");
      return 5;
    }
  } catch (const std::exception& e) {
    fprintf(stderr, "exception: %s\n", e.what());
    return 4;
  }
  down(fo, desc, (size_t)nloc * nd * M);
  down(fo, gx, (size_t)nloc * nnei);
  down(fo, gem, (size_t)nloc * nnei * nd);
  down(fo, gg, (size_t)nloc * nd * M);
  fclose(fo);
  printf("SHIM_DRIVER_OK\n");
  return 0;
}

int main(int argc, char** argv) {
  if (argc >= 4 && std::string(argv[1]) == "nlist") return nlist_mode(argv[2], argv[3]);
  if (argc >= 4 && std::string(argv[1]) == "tabnd") return tabnd_mode(argv[2], argv[3]);
  if (argc < 3) return 1;
  FILE* fi = fopen(argv[1], "rb");
  FILE* fo = fopen(argv[2], "wb");
  if (!fi || !fo) return 1;
  // header: nloc nall nnei ntypes max_nbor nspline M
  auto hd = rd<int>(fi, 7);
  const int nloc = hd[0], nall = hd[1], nnei = hd[2], ntypes = hd[3], max_nbor = hd[4], nspline = hd[5], M = hd[6];
  auto sec = rd<int>(fi, ntypes + 1);
  auto rc = rd<float>(fi, 2);  // rcut, rcut_smth
  auto coord = rd<double>(fi, (size_t)nall * 3);
  auto type = rd<int>(fi, nall);
  auto numneigh = rd<int>(fi, nloc);
  auto rows = rd<int>(fi, (size_t)nloc * max_nbor);
  auto avg = rd<double>(fi, (size_t)ntypes * nnei * 4);
  auto std_ = rd<double>(fi, (size_t)ntypes * nnei * 4);
  auto table = rd<double>(fi, (size_t)nspline * M * 6);
  auto info = rd<double>(fi, 6);
  auto net_deriv = rd<double>(fi, (size_t)nloc * nnei * 4);
  auto dy = rd<double>(fi, (size_t)nloc * 4 * M);

  double *d_coord = up(coord), *d_avg = up(avg), *d_std = up(std_), *d_table = up(table), *d_nd = up(net_deriv),
         *d_dy = up(dy);
  int *d_type = up(type), *d_numneigh = up(numneigh), *d_rows = up(rows);
  std::vector<int> il(nloc);
  std::vector<int*> first(nloc);
  for (int i = 0; i < nloc; ++i) {
    il[i] = i;
    first[i] = d_rows + (size_t)i * max_nbor;
  }
  int* d_ilist = up(il);
  int** d_first = nullptr;
  CK(cudaMalloc((void**)&d_first, sizeof(int*) * nloc));
  CK(cudaMemcpy(d_first, first.data(), sizeof(int*) * nloc, cudaMemcpyHostToDevice));
  deepmd::InputNlist gpu_inlist(nloc, d_ilist, d_numneigh, d_first);

  double *em, *dv, *rij, *desc, *gx, *gem, *force, *virial, *atom_virial;
  int *nlist, *array_int;
  unsigned long long* array_ll;
  CK(cudaMalloc((void**)&em, sizeof(double) * nloc * nnei * 4));
  CK(cudaMalloc((void**)&dv, sizeof(double) * nloc * nnei * 12));
  CK(cudaMalloc((void**)&rij, sizeof(double) * nloc * nnei * 3));
  CK(cudaMalloc((void**)&nlist, sizeof(int) * nloc * nnei));
  CK(cudaMalloc((void**)&array_int, sizeof(int) * (sec.size() + (size_t)nloc * sec.size() + nloc)));
  CK(cudaMalloc((void**)&array_ll, sizeof(unsigned long long) * (size_t)nloc * max_nbor * 2));
  CK(cudaMalloc((void**)&desc, sizeof(double) * nloc * 4 * M));
  CK(cudaMalloc((void**)&gx, sizeof(double) * nloc * nnei));
  CK(cudaMalloc((void**)&gem, sizeof(double) * nloc * nnei * 4));
  CK(cudaMalloc((void**)&force, sizeof(double) * nall * 3));
  CK(cudaMalloc((void**)&virial, sizeof(double) * 9));
  CK(cudaMalloc((void**)&atom_virial, sizeof(double) * nall * 9));
  double *gn_f, *gn_v;
  CK(cudaMalloc((void**)&gn_f, sizeof(double) * nloc * nnei * 4));
  CK(cudaMalloc((void**)&gn_v, sizeof(double) * nloc * nnei * 4));

  try {
    deepmd::prod_env_mat_a_gpu<double>(em, dv, rij, nlist, d_coord, d_type, gpu_inlist, array_int, array_ll, max_nbor,
                                       d_avg, d_std, nloc, nall, 1, rc[0], rc[1], sec);
    // one table over the whole env-mat (em_x = component 0): the op-level call of tabulate_multi_device.cc
    std::vector<double> h_em((size_t)nloc * nnei * 4), h_x((size_t)nloc * nnei);
    CK(cudaMemcpy(h_em.data(), em, sizeof(double) * h_em.size(), cudaMemcpyDeviceToHost));
    for (size_t k = 0; k < h_x.size(); ++k) h_x[k] = h_em[4 * k];
    double* d_x = up(h_x);
    deepmd::tabulate_fusion_se_a_gpu<double>(desc, d_table, info.data(), d_x, em, nullptr, nloc, nnei, M);
    deepmd::tabulate_fusion_se_a_grad_gpu<double>(gx, gem, nullptr, d_table, info.data(), d_x, em, nullptr, d_dy, nloc,
                                                  nnei, M);
    deepmd::prod_force_a_gpu<double>(force, d_nd, dv, nlist, nloc, nall, nnei, 1);
    deepmd::prod_virial_a_gpu<double>(virial, atom_virial, d_nd, dv, rij, nlist, nloc, nall, nnei);
    // gradients w.r.t. net_deriv, fed with the first nloc force rows / the virial just computed
    deepmd::prod_force_grad_a_gpu<double>(gn_f, force, dv, nlist, nloc, nnei, 1);
    deepmd::prod_virial_grad_a_gpu<double>(gn_v, virial, dv, rij, nlist, nloc, nnei);
  } catch (const std::exception& e) {
    fprintf(stderr, "exception: %s\n", e.what());
    return 4;
  }
  down(fo, em, (size_t)nloc * nnei * 4);
  down(fo, dv, (size_t)nloc * nnei * 12);
  down(fo, rij, (size_t)nloc * nnei * 3);
  down(fo, nlist, (size_t)nloc * nnei);
  down(fo, desc, (size_t)nloc * 4 * M);
  down(fo, gx, (size_t)nloc * nnei);
  down(fo, gem, (size_t)nloc * nnei * 4);
  down(fo, force, (size_t)nall * 3);
  down(fo, virial, 9);
  down(fo, atom_virial, (size_t)nall * 9);
  down(fo, gn_f, (size_t)nloc * nnei * 4);
  down(fo, gn_v, (size_t)nloc * nnei * 4);
  fclose(fo);
  printf("SHIM_DRIVER_OK\n");
  return 0;
}
