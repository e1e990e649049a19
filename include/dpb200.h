/* dpb200 — C ABI of the B200-native compressed se_e2_a / se_atten force-evaluation hot path.
 *
 * This is the drop-in boundary: plain device pointers, sizes and a CUDA
 * stream; no torch / TF types.  Every entry point names the reference interface it
 * replaces (paths relative to /root/reference).  The reference-side bindings (C++
 * `deepmd::*_gpu` shim, torch ops) that a maintainer adds on top are shown in
 * INTEGRATION.md and live in deepmd-kit_b200/csrc/deepmd_gpu_shim.cc and
 * deepmd-kit_b200/ops.py.
 *
 * Conventions
 *  - `_f64` / `_f32` suffix = FPTYPE of the reference template instantiation.
 *  - All array arguments are DEVICE pointers unless marked [host].
 *  - Outputs need not be pre-zeroed (the reference wrappers memset them; so do we).
 *  - Work is enqueued on `stream` and NOT synchronised (the reference synchronises the whole
 *    device around every kernel; callers that rely on that must synchronise the stream).
 *  - Return value: DPB200_OK or a negative DPB200_ERR_*; dpb200_last_error() gives the
 *    thread-local message.  There is no CPU fallback: without a CUDA device every compute
 *    entry point returns DPB200_ERR_CUDA.
 */
#ifndef DPB200_H_
#define DPB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

#define DPB200_OK 0
#define DPB200_ERR_INVALID (-1)        /* bad argument (maps to deepmd::deepmd_exception) */
#define DPB200_ERR_CUDA (-2)           /* CUDA runtime error */
#define DPB200_ERR_OOM (-3)            /* maps to deepmd::deepmd_exception_oom (errors.h:17-22) */
#define DPB200_ERR_NLIST_CAPACITY (-4) /* maps to deepmd_exception_nlist_capacity (errors.h:30-33) */

#define DPB200_TAB_COMPRESSED_COEF 1   /* flags bit of the dpb200_tabulate_fusion_se_a_desc / _grad_fx entry points */

#define DPB200_MAX_NBOR_SIZE 4096 /* GPU_MAX_NBOR_SIZE, source/lib/include/gpu_cuda.h:20 */
#define DPB200_MAX_TYPES 128      /* 7 type bits in the sort key (same limit as prod_env_mat.cu:83-104) */
#define DPB200_MAX_NALL (1 << 26) /* 26 index bits in the sort key (reference: 1<<24) */

typedef struct CUstream_st* dpb200_stream_t; /* == cudaStream_t */

const char* dpb200_last_error(void);
int dpb200_abi_version(void);
/* Number of dpb200 kernels enqueued by this process so far (all threads). */
long long dpb200_launch_count(void);
/* Bookkeeping only: kernels re-issued by replaying a CUDA graph that captured dpb200 calls. */
void dpb200_count_replayed_launches(long long n);
/* Measured peak FMA rate (TFLOP/s, 2 flops per FMA) of the FP64 / FP32 pipe of the current device:
 * the roofline denominator of the tabulate kernels (MEASURED_PEAKS.json has no such entry). */
int dpb200_fma_peak_f64(double* tflops /*host out*/, dpb200_stream_t stream);
int dpb200_fma_peak_f32(double* tflops /*host out*/, dpb200_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * prod_env_mat_a : neighbour formatting + environment matrix + normalisation, one launch.
 * Replaces deepmd::prod_env_mat_a_gpu  (source/lib/include/prod_env_mat.h:91-110,
 * source/lib/src/gpu/prod_env_mat.cu:681-726); results follow the CPU semantics of
 * prod_env_mat_a_cpu (source/lib/src/prod_env_mat.cc:14-132): nlist bit-exact with
 * format_nlist_i_cpu (source/lib/src/fmt_nlist.cc:98-143).
 *
 * Raw neighbour rows: row r (r < nframes*nloc) holds numneigh[r] indices of centre atom
 * ilist[r] (ilist==NULL: centre r).  Either `firstneigh` (device array of device row
 * pointers, the InputNlist layout of neighbor_list.h:20-57 after convert_nlist_gpu_device)
 * or, when firstneigh==NULL, a dense block `rows` with `row_stride` ints per row.
 * max_nbor_size: upper bound of numneigh (row capacity), <= DPB200_MAX_NBOR_SIZE.
 * f_type may be NULL (= type).  sec [host] has nsec = ntypes+1 entries.
 * workspace: >= dpb200_prod_env_mat_a_workspace_bytes(...) bytes, 256-byte aligned.
 * ------------------------------------------------------------------------------------- */
size_t dpb200_prod_env_mat_a_workspace_bytes(int ntypes, int nnei, int nall, int nframes, int fp_bytes);

#define DPB200_DECL_ENV(SUF, FP)                                                                  \
  int dpb200_prod_env_mat_a_##SUF(FP* em, FP* em_deriv, FP* rij, int* nlist, const FP* coord,     \
                                  const int* type, const int* f_type, const int* ilist,           \
                                  const int* numneigh, const int* const* firstneigh,              \
                                  const int* rows, int row_stride, int max_nbor_size,             \
                                  const FP* avg, const FP* std, int nloc, int nall, int nframes,  \
                                  float rcut, float rcut_smth, const int* sec, int nsec,          \
                                  void* workspace, size_t workspace_bytes, dpb200_stream_t stream); \
  /* same with avg / std holding `ntypes_center` rows (centre-atom types) independent of the number of  \
   * sections: se_atten formats ONE distance-ordered section (f_type = 0) but normalises per real centre  \
   * type (deepmd/pt/model/descriptor/se_atten.py; prod_env_mat_a_cpu indexes avg by type[i] the same way). \
   * workspace: dpb200_prod_env_mat_a_workspace_bytes(ntypes_center, ...). */                          \
  int dpb200_prod_env_mat_a_ex_##SUF(FP* em, FP* em_deriv, FP* rij, int* nlist, const FP* coord,  \
                                     const int* type, const int* f_type, const int* ilist,        \
                                     const int* numneigh, const int* const* firstneigh,           \
                                     const int* rows, int row_stride, int max_nbor_size,          \
                                     const FP* avg, const FP* std, int ntypes_center, int nloc,   \
                                     int nall, int nframes, float rcut, float rcut_smth,          \
                                     const int* sec, int nsec, void* workspace,                   \
                                     size_t workspace_bytes, dpb200_stream_t stream);             \
  /* format only: replaces deepmd::format_nbor_list_gpu (source/lib/include/fmt_nlist.h:22-34) */ \
  int dpb200_format_nlist_##SUF(int* nlist, const FP* coord, const int* type, const int* ilist,   \
                                const int* numneigh, const int* const* firstneigh,                \
                                const int* rows, int row_stride, int max_nbor_size, int nloc,     \
                                int nall, int nframes, float rcut, const int* sec, int nsec,      \
                                void* workspace, size_t workspace_bytes, dpb200_stream_t stream);
DPB200_DECL_ENV(f64, double)
DPB200_DECL_ENV(f32, float)
#undef DPB200_DECL_ENV

/* ---------------------------------------------------------------------------------------
 * tabulate_fusion_se_a (+ se_atten when two_embed != NULL) forward / backward / 2nd order.
 * Replace deepmd::tabulate_fusion_se_a_gpu, _grad_gpu, _grad_grad_gpu
 * (source/lib/include/tabulate.h:175-218; source/lib/src/gpu/tabulate.cu:1219-1370), CPU
 * semantics source/lib/src/tabulate.cc:162-447.  table: [nspline][M][6] (coefficients
 * innermost); table_info [host]: lower, upper, max, stride0, stride1, (check_freq).
 * em_x [nloc*nnei], em [nloc*nnei*4], two_embed [nloc*nnei*M] or NULL, out/dy [nloc*4*M].
 *
 * The `_ex` forms take element strides so a caller can pass one type-section of the full
 * env-mat without copying (deepmd/pt/model/descriptor/se_a.py:810-831 slices and copies):
 *   em_x[i,j]   at em_x + i*ldx_i + j*ldx_j ;  em[i,j,0:4] at em + i*ldem_i + j*4
 *   dy_dem_x / dy_dem use the same strides as em_x / em;  grad_ex with dy_dem_x == NULL adds
 *   the em_x gradient into dy_dem[i,j,0] (em_x being component 0 of em, as in se_a.py:818-820);
 *   accumulate!=0 adds into `out` instead of overwriting it (sum over type sections).
 * ------------------------------------------------------------------------------------- */
#define DPB200_DECL_TAB(SUF, FP)                                                                   \
  int dpb200_tabulate_fusion_se_a_##SUF(FP* out, const FP* table, const FP* table_info,            \
                                        const FP* em_x, const FP* em, const FP* two_embed,         \
                                        int nloc, int nnei, int last_layer_size, int is_sorted,    \
                                        dpb200_stream_t stream);                                   \
  int dpb200_tabulate_fusion_se_a_grad_##SUF(FP* dy_dem_x, FP* dy_dem, FP* dy_dtwo,                \
                                             const FP* table, const FP* table_info,                \
                                             const FP* em_x, const FP* em, const FP* two_embed,    \
                                             const FP* dy, int nloc, int nnei,                     \
                                             int last_layer_size, int is_sorted,                   \
                                             dpb200_stream_t stream);                              \
  /* se_atten strip gate WITHOUT the materialised two_embed [nloc*nnei][M] (96 KB per atom in fp64):     \
   * two_embed[i][j][k] = tt_full[pair[i][j]][k] * sw[i][j]  (deepmd/pt/model/descriptor/se_atten.py:979-985: \
   * gg_t = tt_full[tebd_idx] * sw) is formed inside the kernel; the backward returns                     \
   * dy_dsw[i][j] = sum_k dy_dtwo[i][j][k] * tt_full[pair][k] instead of dy_dtwo.  Same results as the      \
   * reference-schema entry points above fed with the materialised tensor.  dy_dem_x may be NULL: em_x is  \
   * then component 0 of em and its gradient is added into dy_dem[..][0].  flags: 0 or the               \
   * DPB200_TAB_COMPRESSED_COEF word of the se_a_desc entry point below (caller-validated table).  fp64  \
   * backward: FP64 tensor cores (three m8n8k4 products per step: G(1+t), G'(1+t), G tt). */             \
  int dpb200_tabulate_fusion_se_atten_gate_##SUF(                                                  \
      FP* out, const FP* table, const FP* table_info, const FP* em_x, const FP* em,                \
      const FP* tt_full, const int* pair, const FP* sw, int nloc, int nnei, int last_layer_size,   \
      int is_sorted, int flags, dpb200_stream_t stream);                                           \
  /* the gated forward with the descriptor epilogue of the se_a_desc entry point (same desc_mode values);        \
   * slice_stride (0: M*axis) = elements between two digit slices of a row, min_row_exp = lower bound of the    \
   * row exponent: the caller appends its own columns (centre type embedding) with dpb200_fit_slice_cols. */     \
  int dpb200_tabulate_fusion_se_atten_gate_desc_##SUF(                                             \
      FP* out, const FP* table, const FP* table_info, const FP* em_x, const FP* em,                \
      const FP* tt_full, const int* pair, const FP* sw, int nloc, int nnei, int last_layer_size,   \
      int is_sorted, int axis, double scale, const int* desc_row /*nullable*/, int desc_mode,      \
      void* desc, long long desc_ld, long long slice_stride, int nslice, int* row_exp,             \
      int min_row_exp, int flags, dpb200_stream_t stream);                                         \
  int dpb200_tabulate_fusion_se_atten_gate_grad_##SUF(                                             \
      FP* dy_dem_x /*nullable*/, FP* dy_dem, FP* dy_dsw, const FP* table, const FP* table_info,    \
      const FP* em_x, const FP* em, const FP* tt_full, const int* pair, const FP* sw,              \
      const FP* dy, int nloc, int nnei, int last_layer_size, int is_sorted, int flags,             \
      dpb200_stream_t stream);                                                                     \
  int dpb200_tabulate_fusion_se_a_grad_grad_##SUF(                                                 \
      FP* dz_dy, const FP* table, const FP* table_info, const FP* em_x, const FP* em,              \
      const FP* two_embed, const FP* dz_dy_dem_x, const FP* dz_dy_dem, const FP* dz_dy_dtwo,       \
      int nloc, int nnei, int last_layer_size, int is_sorted, dpb200_stream_t stream);             \
  int dpb200_tabulate_fusion_se_a_ex_##SUF(FP* out, const FP* table, const FP* table_info,         \
                                           const FP* em_x, long long ldx_i, int ldx_j,             \
                                           const FP* em, long long ldem_i, const FP* two_embed,    \
                                           int nloc, int nnei, int last_layer_size,                \
                                           int is_sorted, int accumulate, dpb200_stream_t stream); \
  int dpb200_tabulate_fusion_se_a_grad_ex_##SUF(                                                   \
      FP* dy_dem_x, FP* dy_dem, FP* dy_dtwo, const FP* table, const FP* table_info,                \
      const FP* em_x, long long ldx_i, int ldx_j, const FP* em, long long ldem_i,                  \
      const FP* two_embed, const FP* dy, int nloc, int nnei, int last_layer_size, int is_sorted,   \
      dpb200_stream_t stream);                                                                     \
  /* `_ex` forward of the LAST type section with the se_e2_a descriptor contraction fused into its \
   * epilogue (deepmd/pt/model/descriptor/se_a.py:843-850): after out[i] is complete,              \
   * D = (scale*out[i])^T (scale*out[i])[:, :axis] is written to row desc_row[i] (NULL: i) of      \
   * `desc` (row stride desc_ld elements), in the operand format of the fitting net's first GEMM:  \
   *   desc_mode 0: no descriptor (plain `_ex` forward, but honouring `flags`);                     \
   *   desc_mode 1: FP [M*axis];                                                                    \
   *   desc_mode 2, f64: int8 [nslice][M*axis] balanced base-256 digit slices (most significant first) of     \
   *     D * 2^-row_exp[row] (needs axis == 16);  f32: float [2][M*axis] = TF32 head | tail;         \
   *   desc_mode 3, f32 only: int8 [4][M*axis] digit slices of the 32-bit fixed-point image of       \
   *     D * 2^-row_exp[row] (axis == 16, nslice == 4): operand of dpb200_fit_gemm_i8 with nslice 4.  \
   * M <= 128, axis <= 32, desc 16-byte aligned.                                                   \
   * flags: DPB200_TAB_COMPRESSED_COEF (f64 only, ignored for f32) = the caller has checked that    \
   * for THIS table a3, a4 may be rounded to fp32, a5 to fp16 and a2 to 36 mantissa bits (true for  \
   * dp-compress tables with stride 0.01: error < 1e-12 relative; see csrc/tabulate.cu pack_cm).    \
   * Bits 8..15 of flags: signed exponent k, a5 is stored as half(a5 * 2^k) (choose k so that       \
   * max|a5| * 2^k ~ 2^14).  The backward then streams 32 instead of 48 bytes per coefficient set  \
   * (the forward ignores the flag: measured slower).  The reference-facing                         \
   * entry points above never do this. */                                                          \
  int dpb200_tabulate_fusion_se_a_desc_##SUF(                                                      \
      FP* out, const FP* table, const FP* table_info, const FP* em_x, long long ldx_i, int ldx_j,  \
      const FP* em, long long ldem_i, int nloc, int nnei, int last_layer_size, int is_sorted,      \
      int accumulate, int axis, double scale, const int* desc_row /*nullable*/, int desc_mode,     \
      void* desc, long long desc_ld, int nslice, int* row_exp /*f64 mode 2*/, int flags,           \
      dpb200_stream_t stream);                                                                     \
  /* grad_ex (plain se_a, dy_dem_x may be NULL) with `flags` */                                    \
  int dpb200_tabulate_fusion_se_a_grad_fx_##SUF(                                                   \
      FP* dy_dem_x, FP* dy_dem, const FP* table, const FP* table_info, const FP* em_x,             \
      long long ldx_i, int ldx_j, const FP* em, long long ldem_i, const FP* dy, int nloc,          \
      int nnei, int last_layer_size, int is_sorted, int flags, dpb200_stream_t stream);
DPB200_DECL_TAB(f64, double)
DPB200_DECL_TAB(f32, float)
#undef DPB200_DECL_TAB

/* ---------------------------------------------------------------------------------------
 * prod_force_a / prod_virial_a : reverse scatter of dE/d(env-mat) to forces and virial.
 * Replace deepmd::prod_force_a_gpu (source/lib/include/prod_force.h:71-79,
 * source/lib/src/gpu/prod_force.cu:104-131) and deepmd::prod_virial_a_gpu
 * (source/lib/include/prod_virial.h:30-39, source/lib/src/gpu/prod_virial.cu:106-134);
 * CPU semantics source/lib/src/prod_force.cc:23-87, prod_virial.cc:22-69.
 * force [nframes*nall*3], virial [9], atom_virial [nall*9] (may be NULL in the fused form).
 * dpb200_prod_force_virial_a does both in one pass over net_deriv / in_deriv.
 * ------------------------------------------------------------------------------------- */
#define DPB200_DECL_FV(SUF, FP)                                                                    \
  int dpb200_prod_force_a_##SUF(FP* force, const FP* net_deriv, const FP* in_deriv,                \
                                const int* nlist, int nloc, int nall, int nnei, int nframes,       \
                                dpb200_stream_t stream);                                           \
  int dpb200_prod_virial_a_##SUF(FP* virial, FP* atom_virial, const FP* net_deriv,                 \
                                 const FP* in_deriv, const FP* rij, const int* nlist, int nloc,    \
                                 int nall, int nnei, dpb200_stream_t stream);                      \
  int dpb200_prod_force_virial_a_##SUF(FP* force, FP* virial, FP* atom_virial,                     \
                                       const FP* net_deriv, const FP* in_deriv, const FP* rij,     \
                                       const int* nlist, int nloc, int nall, int nnei,             \
                                       dpb200_stream_t stream);                                    \
  /* atom-chunked form: rows r = 0..nrows-1 belong to centre atoms center_offset + r; with           \
   * accumulate != 0 the outputs are added to instead of zeroed first (sum over chunks) */          \
  int dpb200_prod_force_virial_a_ex_##SUF(FP* force, FP* virial, FP* atom_virial,                  \
                                          const FP* net_deriv, const FP* in_deriv, const FP* rij,  \
                                          const int* nlist, int nrows, int center_offset,          \
                                          int nall, int nnei, int accumulate,                      \
                                          dpb200_stream_t stream);
DPB200_DECL_FV(f64, double)
DPB200_DECL_FV(f32, float)
#undef DPB200_DECL_FV

/* Gradients of the two scatters with respect to net_deriv (backward of force / virial when a compressed
 * model is trained).  Replace deepmd::prod_force_grad_a_gpu (source/lib/include/prod_force_grad.h:26-33;
 * CPU semantics source/lib/src/prod_force_grad.cc:22-77: grad [nframes*nloc*3], neighbour indices >= nloc are
 * folded with j % nloc) and deepmd::prod_virial_grad_a_gpu (prod_virial_grad.h:26-33; prod_virial_grad.cc:21-63:
 * grad [9], one frame).  grad_net [nframes*nloc*nnei*4] is fully written. */
#define DPB200_DECL_FVG(SUF, FP)                                                                   \
  int dpb200_prod_force_grad_a_##SUF(FP* grad_net, const FP* grad, const FP* in_deriv,             \
                                     const int* nlist, int nloc, int nnei, int nframes,            \
                                     dpb200_stream_t stream);                                      \
  /* same with grad holding `ngrad` >= nloc atoms per frame (fold modulus ngrad): with ngrad = nall   \
   * this is the exact adjoint of dpb200_prod_force_a for forces on ghost atoms */                  \
  int dpb200_prod_force_grad_a_ex_##SUF(FP* grad_net, const FP* grad, const FP* in_deriv,          \
                                        const int* nlist, int nloc, int ngrad, int nnei,           \
                                        int nframes, dpb200_stream_t stream);                      \
  int dpb200_prod_virial_grad_a_##SUF(FP* grad_net, const FP* grad, const FP* in_deriv,            \
                                      const FP* rij, const int* nlist, int nloc, int nnei,         \
                                      dpb200_stream_t stream);
DPB200_DECL_FVG(f64, double)
DPB200_DECL_FVG(f32, float)
#undef DPB200_DECL_FVG

/* ---------------------------------------------------------------------------------------
 * Neighbour-list front end (cell list; the reference GPU path is O(nloc*nall)).
 *  normalize_coord : deepmd::normalize_coord_gpu (source/lib/include/coord.h:47-55)
 *  copy_coord      : deepmd::copy_coord_gpu      (coord.h:57-85): periodic ghost images within
 *                    rcut; local atoms first; returns 1 and the needed *nall when nall>mem_nall.
 *                    Ghost ORDER is unspecified (the reference tests sort before comparing,
 *                    source/lib/tests/test_coord.cc:137-165); we emit (owner cell, image) order.
 *  build_nlist     : deepmd::build_nlist_gpu     (source/lib/include/neighbor_list.h:256-266):
 *                    all j != i with |ri-rj|^2 < rcut^2 (strict, FPTYPE), type<0 excluded; rows
 *                    written to a dense [nloc][mem_size] block in ascending j (the order of
 *                    build_nlist_cpu, neighbor_list.cc:875-928); returns 1 and *max_list_size
 *                    when a row needs more than mem_size.
 * boxt [host] is the 3x3 row-major cell.  workspace sizes via the *_workspace_bytes calls.
 * ------------------------------------------------------------------------------------- */
size_t dpb200_copy_coord_workspace_bytes(int nloc);
size_t dpb200_build_nlist_workspace_bytes(int nall);

#define DPB200_DECL_NL(SUF, FP)                                                                    \
  int dpb200_normalize_coord_##SUF(FP* coord, int natom, const FP* boxt, dpb200_stream_t stream);  \
  int dpb200_copy_coord_##SUF(FP* out_c, int* out_t, int* mapping, int* nall /*host out*/,         \
                              const FP* in_c, const int* in_t, int nloc, int mem_nall,             \
                              float rcut, const FP* boxt, void* workspace,                         \
                              size_t workspace_bytes, dpb200_stream_t stream);                     \
  /* same, with the cell grid of compute_cell_info (coord.cc:68-108) handed over: ncell[3] =        \
   * cell_info[3..5], ngcell[3] = cell_info[12..14] -- the form copy_coord_gpu receives */          \
  int dpb200_copy_coord_cells_##SUF(FP* out_c, int* out_t, int* mapping, int* nall /*host out*/,   \
                                    const FP* in_c, const int* in_t, int nloc, int mem_nall,       \
                                    const int* ncell, const int* ngcell, const FP* boxt,           \
                                    void* workspace, size_t workspace_bytes,                       \
                                    dpb200_stream_t stream);                                       \
  int dpb200_build_nlist_##SUF(int* numneigh, int* rows, int* max_list_size /*host out*/,          \
                               const FP* coord, int nloc, int nall, int mem_size, float rcut,      \
                               const int* type, void* workspace, size_t workspace_bytes,           \
                               dpb200_stream_t stream);
DPB200_DECL_NL(f64, double)
DPB200_DECL_NL(f32, float)
#undef DPB200_DECL_NL
/* Copy dense rows [nrows][row_stride] into the caller-owned rows of an InputNlist (firstneigh: device array of
 * device row pointers, neighbor_list.h:20-57) and fill ilist[i] = i % nloc (nullable) -- what fill_nlist /
 * build_nlist of source/lib/src/gpu/neighbor_list.cu:78-128 leave behind for the caller of build_nlist_gpu. */
int dpb200_scatter_nlist_rows(int* const* firstneigh, int* ilist /*nullable*/, const int* rows, int row_stride,
                              const int* numneigh, int nrows, int nloc, dpb200_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Ghost halo for the spatially decomposed multi-GPU path.  Replaces what LAMMPS' Comm does around
 * PairDeepMD::compute (source/lmp/pair_deepmd.cpp:482-488 reverse_comm, :1069-1093 pack/unpack):
 *   halo_pack       sendbuf[k,:] = coord[sendlist[k],:] + shift[k,:]   (k < n; shift = image vector)
 *   halo_unpack_add force[sendlist[k],:] += recvbuf[k,:]               (ghost forces to owners)
 * The transfer itself is a grouped NCCL send/recv issued by the host side.
 * ------------------------------------------------------------------------------------- */
#define DPB200_DECL_HALO(SUF, FP)                                                                      \
  int dpb200_halo_pack_##SUF(FP* sendbuf, const FP* coord, const int* sendlist, const FP* shift, int n, \
                             dpb200_stream_t stream);                                                  \
  int dpb200_halo_unpack_add_##SUF(FP* force, const FP* recvbuf, const int* sendlist, int n,           \
                                   dpb200_stream_t stream);
DPB200_DECL_HALO(f64, double)
DPB200_DECL_HALO(f32, float)
#undef DPB200_DECL_HALO

/* ---------------------------------------------------------------------------------------
 * se_e2_a descriptor contraction  D[i] = (gr[r]*scale)^T (gr[r]*scale)[:, :axis]  and its backward
 * (deepmd/pt/model/descriptor/se_a.py:843-850: xyz_scatter /= nnei; matmul(xyz_scatter_1,
 * xyz_scatter_2)).  gr [*][4][M] = summed tabulate output, scale = 1/nnei, D [nloc][M*axis];
 * rows (may be NULL = identity): row i of D / dD belongs to atom r = rows[i] of gr / dgr, so the
 * per-type gather of the fitting net and the scatter of its gradient need no extra pass.
 * mlp_tanh_fwd / _bwd: the elementwise part of one fitting-net layer (deepmd/pt/model/network/
 * mlp.py: y = tanh(z)*idt + h; t = g*idt*(1 - tanh^2)), z_a is overwritten by tanh(z).
 * ------------------------------------------------------------------------------------- */
#define DPB200_DECL_DESC(SUF, FP)                                                                       \
  int dpb200_se_a_descriptor_##SUF(FP* D, const FP* gr, const int* rows, long long nloc, int M,         \
                                   int axis, double scale, dpb200_stream_t stream);                     \
  int dpb200_se_a_descriptor_grad_##SUF(FP* dgr, const FP* dD, const FP* gr, const int* rows,           \
                                        long long nloc, int M, int axis, double scale,                  \
                                        dpb200_stream_t stream);                                        \
  int dpb200_mlp_tanh_fwd_##SUF(FP* z_a, FP* y, const FP* h /*nullable*/, const FP* idt /*nullable*/,   \
                                long long nrow, int width, dpb200_stream_t stream);                     \
  int dpb200_mlp_tanh_bwd_##SUF(FP* t, const FP* g, long long ldg, const FP* a,                         \
                                const FP* idt /*nullable*/, long long nrow, int width,                  \
                                dpb200_stream_t stream);                                                \
  /* same, the result additionally (t may be NULL in _bwd_split) written as the 3xTF32 left operand   \
   * [hi | lo | hi] of the next GEMM (split3: [nrow][3*width]; fp32 only, must be NULL for f64) */      \
  int dpb200_mlp_tanh_fwd_split_##SUF(FP* z_a, FP* y, const FP* h, const FP* idt, long long nrow,       \
                                      int width, FP* split3, dpb200_stream_t stream);                   \
  int dpb200_mlp_tanh_bwd_split_##SUF(FP* t, const FP* g, long long ldg, const FP* a, const FP* idt,    \
                                      long long nrow, int width, FP* split3, dpb200_stream_t stream);
DPB200_DECL_DESC(f64, double)
DPB200_DECL_DESC(f32, float)
#undef DPB200_DECL_DESC

/* ---------------------------------------------------------------------------------------
 * Split operands for the fitting net's GEMMs on the tensor cores (csrc/fitting.cu): the GEMMs are
 * library calls; these kernels make them fp64 / fp32 accurate.
 *  split_i8_rows   : x [nrow][width] (row stride ldx) -> out int8 [nrow][nslice][width] (row stride
 *                    ld_out bytes), row_exp[nrow]:  x = 2^row_exp * sum_s out[s] 2^(-7-8s)
 *  split_i8_combine: z = 2^(row_exp[r]+col_exp[c]-14) * sum_d acc[d][r][c] 2^(-8d) + bias[c], acc[d] =
 *                    int32 [nrow][width] at acc + d*acc_stride = sum_{i+j=d} X_i.W_j;  activation != 0:
 *                    a_out = tanh(z), y_out = a*idt + h (mlp.py layer);  == 0: a_out = z.
 *  split_tf32      : x -> [hi | lo (| hi)] with hi = tf32(x), lo = tf32(x - hi), copies = 2 | 3.
 * ------------------------------------------------------------------------------------- */
int dpb200_split_i8_rows_f64(signed char* out, long long ld_out, int* row_exp, const double* x, long long ldx,
                             long long nrow, int width, int nslice, dpb200_stream_t stream);
int dpb200_split_i8_combine_f64(double* a_out, double* y_out, const int* acc, long long acc_stride, int nslice,
                                const int* row_exp, const int* col_exp, const double* bias /*nullable*/,
                                const double* idt /*nullable*/, const double* h /*nullable*/, long long nrow,
                                int width, int activation, dpb200_stream_t stream);
int dpb200_split_tf32_f32(float* out, long long ld_out, const float* x, long long ldx, long long nrow, int width,
                          int copies, dpb200_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * se_atten (DPA-1 strip mode, attn_layer 0) glue around the pair-indexed gate (csrc/se_atten.cu,
 * csrc/force_virial.cu).  Reference algebra: deepmd/pt/model/descriptor/se_atten.py:916-926 (tebd_idx), :979-985
 * (gg_t = tt_full[tebd_idx] * sw, tabulate_fusion_se_atten); the force through the switch is autograd there.
 *  se_atten_gate_scalars : per (centre, slot) of the formatted list `nlist` (extended indices, -1 = empty):
 *      pair = type_i * (ntypes + 1) + type_j (empty slot or negative type -> the padding type `ntypes`),
 *      sw = spline5 switch of |rij| (switcher.h:61-84; 0 for empty slots), dsw_over_r = sw'(r) / r.
 *  prod_force_virial_a_pair : dpb200_prod_force_virial_a plus the central pair force
 *      -(pair_q * pair_w) * rij on the neighbour of every slot (opposite on the centre), included in the virial
 *      and the atomic virial: pair_q = dE/d(sw) (dpb200_tabulate_fusion_se_atten_gate_grad), pair_w = dsw_over_r.
 * ------------------------------------------------------------------------------------- */
#define DPB200_DECL_ATTEN_GLUE(SUF, FP)                                                                          \
  int dpb200_se_atten_gate_scalars_##SUF(int* pair, FP* sw, FP* dsw_over_r, const int* nlist, const int* type,  \
                                         const FP* rij, int nloc, int nnei, int ntypes, float rcut_smth,        \
                                         float rcut, dpb200_stream_t stream);                                   \
  int dpb200_prod_force_virial_a_pair_##SUF(FP* force, FP* virial, FP* atom_virial /*nullable*/,                \
                                            const FP* net_deriv, const FP* in_deriv, const FP* rij,             \
                                            const int* nlist, const FP* pair_q, const FP* pair_w, int nloc,     \
                                            int nall, int nnei, dpb200_stream_t stream);
DPB200_DECL_ATTEN_GLUE(f64, double)
DPB200_DECL_ATTEN_GLUE(f32, float)
#undef DPB200_DECL_ATTEN_GLUE

/* ---------------------------------------------------------------------------------------
 * DPA-1 attention layers, attn_layer > 0 (csrc/attn_layers.cu).  Replaces what the reference runs
 * as torch modules + autograd: deepmd/pt/model/descriptor/se_atten.py:977-1012 (strip-mode g2 = g_s (1 + gg_t sw),
 * input_r) and :1058-1447 (NeighborGatedAttention / NeighborGatedAttentionLayer / GatedAttentionLayer).  The dense
 * products of a layer are plain GEMMs issued by the caller; these entry points are the stages between them, forward
 * and backward.  All tensors are row-major device arrays; `rows` = atoms * nnei.
 *  embed      : g_s, g_s' [rows, M] from the `dp compress` table at s = em_x[row * em_x_stride] (table_info is a
 *               HOST pointer, 6 numbers) and x0 = g_s (1 + tt_full[pair] sw).
 *  embed_grad : d_em_x[row * stride] += sum_c dx0 (1 + tt sw) g_s';  d_sw[row] += sum_c dx0 g_s tt.
 *  rhat       : rhat [rows, 3] = normalize(em[row, 1:4]) (eps 1e-12), rinv [rows] = 1 / norm (negative when clamped).
 *  rhat_grad  : d_em[row, 1:4] += (d_rhat - rhat (rhat . d_rhat)) * rinv.
 *  qkv_normalize      : qkv [rows, 3, hidden] in place: each of q, k, v <- x / max(|x|, 1e-12) (skipped when
 *               normalize = 0), q additionally * q_scale; inv_norm [rows, 3].
 *  qkv_normalize_grad : d_qkv in place from the stored (normalised, q-scaled) qkv_hat and inv_norm.
 *  weights    : S [natoms, nnei, nnei] = q k^T ->  T = (S + shift) sw_i sw_j - shift, P = softmax_j T,
 *               A = P sw_i sw_j (rhat_i . rhat_j if dotr).  nnei_full >= nnei: the slab was cut to its first nnei
 *               slots and the nnei_full - nnei omitted (empty, trailing) slots are accounted for in the softmax
 *               denominators, each with exp(-shift).
 *  weights_grad : dS (may alias dA) from dA; d_sw [natoms, nnei] and d_rhat [natoms, nnei, 3] are ACCUMULATED into.
 *  residual_layernorm : z = x + y; zhat = (z - mean) / sqrt(var + eps) (biased variance) overwrites y;
 *               out = zhat gamma + beta; rstd [rows].
 *  residual_layernorm_grad : dz = rstd (g - mean g - zhat mean(g zhat)), g = dout gamma.
 * ------------------------------------------------------------------------------------- */
#define DPB200_DECL_ATTN(SUF, FP)                                                                                  \
  int dpb200_se_atten_embed_##SUF(FP* x0, FP* gs, FP* dgs, const FP* table, const FP* table_info, const FP* em_x,  \
                                  long long em_x_stride, const FP* tt_full, const int* pair, const FP* sw,         \
                                  long long rows, int last_layer_size, dpb200_stream_t stream);                    \
  int dpb200_se_atten_embed_grad_##SUF(FP* d_em_x, long long d_em_x_stride, FP* d_sw, const FP* dx0, const FP* gs, \
                                       const FP* dgs, const FP* tt_full, const int* pair, const FP* sw,            \
                                       long long rows, int last_layer_size, dpb200_stream_t stream);               \
  int dpb200_se_atten_rhat_##SUF(FP* rhat, FP* rinv, const FP* em, long long rows, dpb200_stream_t stream);        \
  int dpb200_se_atten_rhat_grad_##SUF(FP* d_em, const FP* d_rhat, const FP* rhat, const FP* rinv, long long rows,  \
                                      dpb200_stream_t stream);                                                     \
  int dpb200_attn_qkv_normalize_##SUF(FP* qkv, FP* inv_norm, long long rows, int hidden, double q_scale,           \
                                      int normalize, dpb200_stream_t stream);                                      \
  int dpb200_attn_qkv_normalize_grad_##SUF(FP* d_qkv, const FP* qkv_hat, const FP* inv_norm, long long rows,       \
                                           int hidden, double q_scale, int normalize, dpb200_stream_t stream);     \
  int dpb200_attn_weights_##SUF(FP* P, FP* A, const FP* S, const FP* sw, const FP* rhat, long long natoms,         \
                                int nnei, int nnei_full, double shift, int dotr, dpb200_stream_t stream);          \
  int dpb200_attn_weights_grad_##SUF(FP* dS, FP* d_sw, FP* d_rhat, const FP* dA, const FP* P, const FP* S,         \
                                     const FP* sw, const FP* rhat, long long natoms, int nnei, double shift,       \
                                     int dotr, dpb200_stream_t stream);                                            \
  int dpb200_attn_residual_layernorm_##SUF(FP* out, FP* y_zhat, FP* rstd, const FP* x, const FP* gamma,            \
                                           const FP* beta, long long rows, int width, double eps,                  \
                                           dpb200_stream_t stream);                                                \
  int dpb200_attn_residual_layernorm_grad_##SUF(FP* dz, const FP* dout, const FP* zhat, const FP* rstd,            \
                                                const FP* gamma, long long rows, int width,                        \
                                                dpb200_stream_t stream);
DPB200_DECL_ATTN(f64, double)
DPB200_DECL_ATTN(f32, float)
#undef DPB200_DECL_ATTN

/* ---------------------------------------------------------------------------------------
 * tabulate_fusion_se_a for the higher angular bases, ndescrpt = 9 / 16 / 25 (csrc/tabulate_nd.cu).
 * Replaces deepmd::tabulate_fusion_se_a{,_grad,_grad_grad}_{cpu,gpu} called with ndescrpt != 4
 * (source/lib/include/tabulate.h:28-72, dispatch source/lib/src/tabulate.cc:456-560; caller
 * source/op/pt/tabulate_multi_device.cc:101-117, ndescrpt = em.size(2)).  Same argument meaning as the
 * reference: em [nloc][nnei][ndescrpt], out / dy / dz_dy [nloc][ndescrpt][last_layer_size], two_embed and its
 * cotangents [nloc][nnei][last_layer_size] or NULL; table_info is a HOST pointer.  ndescrpt == 4 is served by
 * the entry points above (the hot path); any other value is DPB200_ERR_INVALID as in the reference's
 * check_se_a_basis_dimension.
 * ------------------------------------------------------------------------------------- */
#define DPB200_DECL_TAB_ND(SUF, FP)                                                                               \
  int dpb200_tabulate_fusion_se_a_nd_##SUF(FP* out, const FP* table, const FP* table_info, const FP* em_x,       \
                                           const FP* em, const FP* two_embed /*nullable*/, int nloc, int nnei,   \
                                           int last_layer_size, int is_sorted, int ndescrpt,                     \
                                           dpb200_stream_t stream);                                              \
  int dpb200_tabulate_fusion_se_a_grad_nd_##SUF(FP* dy_dem_x, FP* dy_dem, FP* dy_dtwo /*nullable*/,              \
                                                const FP* table, const FP* table_info, const FP* em_x,           \
                                                const FP* em, const FP* two_embed /*nullable*/, const FP* dy,    \
                                                int nloc, int nnei, int last_layer_size, int is_sorted,          \
                                                int ndescrpt, dpb200_stream_t stream);                           \
  int dpb200_tabulate_fusion_se_a_grad_grad_nd_##SUF(FP* dz_dy, const FP* table, const FP* table_info,           \
                                                     const FP* em_x, const FP* em,                               \
                                                     const FP* two_embed /*nullable*/, const FP* dz_dy_dem_x,    \
                                                     const FP* dz_dy_dem, const FP* dz_dy_dtwo /*nullable*/,     \
                                                     int nloc, int nnei, int last_layer_size, int is_sorted,     \
                                                     int ndescrpt, dpb200_stream_t stream);
DPB200_DECL_TAB_ND(f64, double)
DPB200_DECL_TAB_ND(f32, float)
#undef DPB200_DECL_TAB_ND

/* ---------------------------------------------------------------------------------------
 * Fitting net on the tcgen05 tensor cores (csrc/fit_tc.cu).  Replaces the library GEMMs of the
 * energy fitting net (deepmd/pt/model/network/mlp.py layers: y = tanh(x.W + b) * idt (+ x); reference's fused
 * fp32 analogue source/op/pt/graph_fitting.cu:40-90,230-340) with error-free int8 split products whose order
 * accumulators stay in tensor memory, the layer's elementwise chain fused into the epilogue.
 *  fit_gemm_i8 : C[nrow][N] = A.B^T, A = int8 slices [nrow][nslice][K] (byte strides a_slice_stride /
 *                a_row_stride; per-row exponent row_exp[r] or row_exp_fixed when row_exp == NULL), B = int8
 *                slices [nslice][N][b_k_stride] (K contiguous, zero padded to a multiple of 64), N % 16 == 0.
 *                colv [N][4] (32-byte aligned) = per output column {2^(col_exp-14), add, mul, 0}.
 *                fp64 matrices other than `out0` of mode 2 use the row-blocked layout: element (r, c) at
 *                ((r/128)*N + c)*128 + r%128, allocated for whole blocks of 128 rows.
 *                mode 0  t = tanh(C + add), y = t*mul (+ skip): out0 = t, out1 = y, slices_out (nullable) = y as
 *                        [nrow][nslice][kp_out] int8 with the fixed exponent out_exp   (add = bias, mul = idt);
 *                mode 1  g = C + add (+ skip), dz = g*mul*(1 - t_in^2): out0 = g (nullable), out1 = dz
 *                        (add = head weights or 0, mul = idt of the layer below);
 *                mode 2  out0[r*ld_out + c] = C  (row-major);
 *                mode 3  the same with `out0` pointing to a FLOAT matrix (16-byte aligned rows): dE/dD of an fp32
 *                        model.
 *                nslice = 6 (fp64 model: 47 fraction bits per operand) or 4 (fp32 model: 31 bits; the A operand of
 *                the first layer is then the mode-3 descriptor of dpb200_tabulate_fusion_se_a_desc_f32).
 *  fit_slice_rows : blocked fp64 [nrow][N] -> int8 slices [nrow][nslice][kp] + row_exp (A operand of mode 1 / 2).
 *  fit_head    : e[r] = y[r,:].w_head + b_head and the backward seed dz = w_head*idt*(1 - t^2) as slices.
 *  fit_blocked : row-major <-> blocked conversion (tests, callers that keep row-major activations).
 * ------------------------------------------------------------------------------------- */
/* columns [col0, col0 + width) of every digit slice of row r := digits of src[idx[r] (NULL: r)][c] at the exponent
 * row_exp[r] already chosen by the producer of the other columns (se_atten: the centre type embedding appended to
 * the descriptor written by dpb200_tabulate_fusion_se_atten_gate_desc with min_row_exp bounding these values). */
int dpb200_fit_slice_cols_f64(signed char* out, long long ld_out, long long slice_stride, int col0, int width,
                              int nslice, const int* row_exp, const double* src, int src_ld,
                              const int* idx /*nullable*/, long long nrow, dpb200_stream_t stream);
int dpb200_fit_gemm_i8_f64(int mode, long long nrow, int N, int K, int nslice, const signed char* a_slices,
                           long long a_slice_stride, long long a_row_stride, const int* row_exp /*nullable*/,
                           int row_exp_fixed, const signed char* b_slices, int b_k_stride, const double* colv,
                           const double* skip /*nullable*/, const double* t_in, double* out0, double* out1,
                           long long ld_out, signed char* slices_out /*nullable*/, long long ld_slices, int kp_out,
                           int out_exp, dpb200_stream_t stream);
int dpb200_fit_slice_rows_f64(signed char* out, long long ld_out, int kp, int* row_exp, const double* x,
                              long long nrow, int N, int nslice, dpb200_stream_t stream);
int dpb200_fit_head_f64(double* e_out, signed char* out, long long ld_out, int kp, int* row_exp, const double* t,
                        const double* y, const double* w_head, const double* idt /*nullable*/, double b_head,
                        long long nrow, int N, int nslice, dpb200_stream_t stream);
int dpb200_fit_blocked_f64(double* dst, const double* src, long long ld, long long nrow, int N, int to_blocked,
                           dpb200_stream_t stream);
/* Profiling hook: the following fit_gemm launches write per-role cycle counters ([grid][16] long long, device
 * memory: TMA producer / MMA issuer / two epilogue warps: total, barrier waits, drain, math); NULL switches it off. */
void dpb200_fit_gemm_debug(long long* counters);

/* use_nlist_map (neighbor_list.h:219-222): nlist[k] = map[nlist[k]] for entries >= 0. */
int dpb200_use_nlist_map(int* nlist, const int* nlist_map, int nloc, int nnei, dpb200_stream_t stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* DPB200_H_ */
