/* TEST INFRASTRUCTURE ONLY — CPU restatement ("oracle") of the reference's compressed
 * se_e2_a / se_atten force-evaluation hot path.
 *
 * Who may use this: tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs, as the CHECKER only.  The product (deepmd-kit_b200/) never
 * imports, links or executes anything under oracle/.
 *
 * Parity status: PINNED.  tests/test_oracle_cpu.py checks every function here against
 * (1) the literal golden vectors of the reference's own unit tests (tests/golden/*.json,
 * extracted by tests/golden/make_golden.py) and (2) the unmodified reference library
 * compiled from /root/reference (oracle/_ref/libdeepmd_ref.so, recipe in oracle/Makefile).
 *
 * The arithmetic lives in dp_oracle_fp.inc, instantiated below for double and float.
 * Build: gcc -O2 -fopenmp -std=c11 -ffp-contract=off -shared (see Makefile). */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int type;
  float d;
  int idx;
} nb_t;

/* (type, dist, index) lexicographic — source/lib/src/fmt_nlist.cc:20-25. */
static int nb_cmp(const void* pa, const void* pb) {
  const nb_t* a = (const nb_t*)pa;
  const nb_t* b = (const nb_t*)pb;
  if (a->type != b->type) return a->type < b->type ? -1 : 1;
  if (a->d != b->d) return a->d < b->d ? -1 : 1;
  if (a->idx != b->idx) return a->idx < b->idx ? -1 : 1;
  return 0;
}

static int int_cmp(const void* pa, const void* pb) {
  const int a = *(const int*)pa, b = *(const int*)pb;
  return (a > b) - (a < b);
}

/* Reciprocal box and face distances in double — include/SimulationRegion_Impl.h:364-372
 * (toFaceDistance), :427-484 (computeVolume / computeRecBox). */
static void box_setup_d(const double* b, double* rec, double* face) {
  double vol = b[0] * (b[4] * b[8] - b[7] * b[5]) - b[1] * (b[3] * b[8] - b[6] * b[5]) +
               b[2] * (b[3] * b[7] - b[6] * b[4]);
  vol = fabs(vol);
  const double vi = 1. / vol;
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
  for (int d = 0; d < 3; ++d) {
    const double* u = r[(d + 1) % 3];
    const double* v = r[(d + 2) % 3];
    const double c[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2],
                         u[0] * v[1] - u[1] * v[0]};
    face[d] = vol * (1. / sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]));
  }
}

/* Image shift that brings an extended-cell index back into [0, ncell) —
 * source/lib/src/neighbor_list.cc:730-745. */
static int pbc_shift_i(int idx, int ncell) {
  int s = 0;
  while (idx + s * ncell < 0) ++s;
  while (idx + s * ncell >= ncell) --s;
  return s;
}

/* compute_cell_info — source/lib/src/coord.cc:68-108 (23 ints). */
int dpo_compute_cell_info(int* ci, float rcut, const double* boxt) {
  double rec[9], face[3];
  box_setup_d(boxt, rec, face);
  for (int d = 0; d < 3; ++d) {
    ci[d] = 0;
    ci[3 + d] = (int)(face[d] / rcut);
    if (ci[3 + d] == 0) ci[3 + d] = 1;
    const double cs = face[d] / ci[3 + d];
    ci[12 + d] = (int)(rcut / cs) + 1;
    ci[6 + d] = -ci[12 + d];
    ci[9 + d] = ci[3 + d] + ci[12 + d];
    ci[15 + d] = ci[12 + d];
    ci[18 + d] = (int)(rcut / cs);
    if (ci[18 + d] * cs < rcut) ci[18 + d] += 1;
  }
  ci[21] = ci[3] * ci[4] * ci[5];
  if (ci[21] <= 0) return -1;
  ci[22] = (2 * ci[12] + ci[3]) * (2 * ci[13] + ci[4]) * (2 * ci[14] + ci[5]);
  return 0;
}

#define FP double
#define SUF f64
#define SQRT sqrt
#define FMOD fmod
#define NEXTAFTER nextafter
#include "dp_oracle_fp.inc"
#undef FP
#undef SUF
#undef SQRT
#undef FMOD
#undef NEXTAFTER

#define FP float
#define SUF f32
#define SQRT sqrtf
#define FMOD fmodf
#define NEXTAFTER nextafterf
#include "dp_oracle_fp.inc"
#undef FP
#undef SUF
#undef SQRT
#undef FMOD
#undef NEXTAFTER
