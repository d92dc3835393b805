/* CPU restatement (C + OpenMP) of the reference's QAOA energy path in the exact-statevector
 * limit  --  TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/qaoa_oracle.py for the numpy twin).
 *
 * It exists so that (a) the GPU parity tests have a checker at register sizes where numpy needs minutes or does not
 * fit (n = 26 ... 32: BASELINE configs 3 and 4 at FULL size), and (b) bench.py can time "the reference's CPU
 * implementation of the path" on the GPU box's host cores: the real reference delegates all amplitude arithmetic to
 * qiskit-aer, which is not installable here (no network; setup.py:22-24 pins only lower bounds), so this file restates
 * what Aer's statevector simulator is asked to do by qaoa/qaoa.py:418-520:
 *     init  ->  [ diag(exp(+i gamma cost(x)))  ->  mixer(beta) ] x p  ->  <cost>, <cost^2>, min, max
 *   cost   : GraphProblem.cost for k = 2 (qaoa/problems/graph_problem.py:152-176), QUBO.qubo_cost
 *            (qaoa/problems/qubo_problem.py:101-108; Portfolio / ExactCover through their Q, c, b)
 *   mixers : X (x_mixer.py:35), XMultiAngle (x_multiangle_mixer.py:40-50), XY (xy_mixer.py:42-79, ordered pairs),
 *            Grover (grover_mixer.py:43-82) over |+> or |D^n_k>
 *   init   : |+> (plus_initialstate.py:19-25), |D^n_k> (dicke_initialstate.py:23-56)
 * Everything runs in the TRUE frame with plain complex arithmetic -- none of the engine's devices (S frame, lifted
 * butterflies, symmetric half register, tables of rotated phases) appear here.
 *
 * Two forms of the X-type layer: `plain` = one sweep over the 2^n amplitudes per gate (what an unfused simulator does)
 * and `blocked` = the same gates grouped per cache block (14 low qubits per block, then <= 5 qubits per sweep: the
 * grouping a gate-fusing simulator such as Aer applies); identical results up to the rounding of the same operations
 * in the same order per amplitude pair (tests/test_oracle.py pins one against the other and both against numpy).
 * Cost diagonal: f64, or u8 (cost = c0 + delta * k, exact for integer-weighted MaxCut) so that n = 32 complex64 needs
 * 32 + 4 GiB of host memory.
 * PARITY: energies are "parity unpinned" against Aer itself; this file is pinned against oracle/qaoa_oracle.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { double re, im; } c128;
typedef struct { float re, im; } c64;

enum { QREF_DIAG_F64 = 0, QREF_DIAG_U8 = 1 };
enum { QREF_INIT_PLUS = 0, QREF_INIT_DICKE = 1 };
enum { QREF_MIXER_X = 0, QREF_MIXER_XMULTI = 1, QREF_MIXER_XY = 2, QREF_MIXER_GROVER = 3 };

typedef struct {
  int kind;          /* QREF_DIAG_* */
  const void* data;  /* 2^n doubles, or 2^n bytes with cost = c0 + delta * byte */
  double c0, delta;
} qref_diag;

typedef struct {
  int n;
  int init_kind, init_k;
  int mixer;
  int sub_kind, sub_k;      /* Grover: |s> */
  const int32_t* pairs;     /* XY: (q0, q1) in application order */
  int n_pairs, trotter;
  int blocked;              /* X-type layers: cache-blocked sweeps (1) or one sweep per gate (0) */
} qref_circuit;

int qref_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* launchers such as torchrun export OMP_NUM_THREADS=1; the CPU baseline wants every host core */
void qref_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* cost diagonal for unweighted/weighted MaxCut: cost(x) = sum_e w_e [x_u != x_v] */
void qref_maxcut_diag(int n, int n_edges, const int32_t* u, const int32_t* v, const double* w, double* diag) {
  const int64_t N = (int64_t)1 << n;
#pragma omp parallel for schedule(static)
  for (int64_t x = 0; x < N; x++) {
    double s = 0.0;
    for (int e = 0; e < n_edges; e++) s += (((x >> u[e]) ^ (x >> v[e])) & 1) ? w[e] : 0.0;
    diag[x] = s;
  }
}

/* the same cost for integer weights with total <= 255, one byte per basis state (returns 1 if it does not fit) */
int qref_maxcut_diag_u8(int n, int n_edges, const int32_t* u, const int32_t* v, const int32_t* w, uint8_t* diag) {
  int64_t tot = 0;
  for (int e = 0; e < n_edges; e++) { if (w[e] < 0) return 1; tot += w[e]; }
  if (tot > 255) return 1;
  const int64_t N = (int64_t)1 << n;
#pragma omp parallel for schedule(static)
  for (int64_t x = 0; x < N; x++) {
    int s = 0;
    for (int e = 0; e < n_edges; e++) s += (((x >> u[e]) ^ (x >> v[e])) & 1) ? w[e] : 0;
    diag[x] = (uint8_t)s;
  }
  return 0;
}

/* cost diagonal of a QUBO: cost(x) = -(x^T Q x + c^T x + b)   (qaoa/problems/qubo_problem.py:101-108) */
void qref_qubo_diag(int n, const double* Q, const double* c, double b, double* diag) {
  const int64_t N = (int64_t)1 << n;
#pragma omp parallel for schedule(static)
  for (int64_t x = 0; x < N; x++) {
    double f = b;
    for (int i = 0; i < n; i++) {
      if (!((x >> i) & 1)) continue;
      double row = (c ? c[i] : 0.0);
      for (int j = 0; j < n; j++) if ((x >> j) & 1) row += Q[(size_t)i * n + j];
      f += row;
    }
    diag[x] = -f;
  }
}

/* The same QUBO cost in O(1) per basis state for large n (n = 30: 2^30 x 30^2 / 4 terms otherwise): split the index
 * into a low half of nl bits and a high half; cost = -(f_lo[x_lo] + f_hi[x_hi] + sum_{j in lo} v_j(x_hi) x_j) with
 * v_j(x_hi) = sum_{i in x_hi} (Q_ij + Q_ji), and the inner sum built incrementally over x_lo.  Rounding differs from
 * qref_qubo_diag by a few ulp (different summation order); tests/test_oracle.py pins one against the other. */
int qref_qubo_diag_fast(int n, const double* Q, const double* c, double b, double* diag) {
  if (n < 2) { qref_qubo_diag(n, Q, c, b, diag); return 0; }
  const int nl = n / 2, nh = n - nl;
  const int64_t NL = (int64_t)1 << nl, NH = (int64_t)1 << nh;
  double* flo = (double*)malloc(sizeof(double) * (size_t)NL);
  double* fhi = (double*)malloc(sizeof(double) * (size_t)NH);
  if (!flo || !fhi) { free(flo); free(fhi); return 1; }
  /* f over a subset of qubits [q0, q0 + m): sum_i x_i (c_i + sum_j Q_ij x_j) restricted to the subset */
  for (int half = 0; half < 2; half++) {
    const int q0 = half ? nl : 0, m = half ? nh : nl;
    double* f = half ? fhi : flo;
    const int64_t M = (int64_t)1 << m;
#pragma omp parallel for schedule(static)
    for (int64_t x = 0; x < M; x++) {
      double s = 0.0;
      for (int i = 0; i < m; i++) {
        if (!((x >> i) & 1)) continue;
        double row = c ? c[q0 + i] : 0.0;
        for (int j = 0; j < m; j++) if ((x >> j) & 1) row += Q[(size_t)(q0 + i) * n + q0 + j];
        s += row;
      }
      f[x] = s;
    }
  }
  int fail = 0;
#pragma omp parallel
  {
    double* cross = (double*)malloc(sizeof(double) * (size_t)NL);
    if (!cross) {
#pragma omp atomic write
      fail = 1;
    }
#pragma omp barrier
    if (!fail) {
#pragma omp for schedule(static)
      for (int64_t xh = 0; xh < NH; xh++) {
        double v[64];
        for (int j = 0; j < nl; j++) {
          double t = 0.0;
          for (int i = 0; i < nh; i++)
            if ((xh >> i) & 1) t += Q[(size_t)(nl + i) * n + j] + Q[(size_t)j * n + nl + i];
          v[j] = t;
        }
        cross[0] = 0.0;
        double* out = diag + (xh << nl);
        out[0] = -(b + fhi[xh]);
        for (int64_t xl = 1; xl < NL; xl++) {
          cross[xl] = cross[xl & (xl - 1)] + v[__builtin_ctzll((unsigned long long)xl)];
          out[xl] = -(b + flo[xl] + fhi[xh] + cross[xl]);
        }
      }
    }
    free(cross);
  }
  free(flo);
  free(fhi);
  return fail;
}

#define QREF_CAT2(a, b) a##b
#define QREF_CAT(a, b) QREF_CAT2(a, b)

#define RT float
#define CT c64
#define FN(name) QREF_CAT(name, _c64)
#include "qaoa_ref_kernels.h"
#undef RT
#undef CT
#undef FN

#define RT double
#define CT c128
#define FN(name) QREF_CAT(name, _c128)
#include "qaoa_ref_kernels.h"
#undef RT
#undef CT
#undef FN

/* General entry.  prec: 0 complex64, 1 complex128.  angles: the reference's flat layout
 * [(gamma_d, beta_d[0..n_beta-1]) for d < p] (qaoa/qaoa.py:937-990), n_beta = n for XMultiAngle, else 1.
 * out5 = {E[cost], E[cost^2], min over support, max over support, norm}.  Returns 0 on success, 1 = out of memory,
 * 2 = bad argument. */
int qref_run(int n, int prec, int diag_kind, const void* diag, double c0, double delta, int init_kind, int init_k,
             int mixer, int sub_kind, int sub_k, const int32_t* pairs, int n_pairs, int trotter, int blocked, int p,
             const double* angles, double* out5) {
  if (n < 1 || n > 40 || p < 1 || !diag || !angles || !out5) return 2;
  if (mixer == QREF_MIXER_XY && (!pairs || n_pairs < 1)) return 2;
  qref_diag D = {diag_kind, diag, c0, delta};
  qref_circuit C = {n, init_kind, init_k, mixer, sub_kind, sub_k, pairs, n_pairs, trotter < 1 ? 1 : trotter, blocked};
  return prec == 0 ? run_c64(&C, &D, p, angles, out5, NULL) : run_c128(&C, &D, p, angles, out5, NULL);
}

/* ---- the first version's entry points (MaxCut / QUBO diagonal in f64, |+>, X or multi-angle X, one sweep per gate) */
#define DEFINE_ENERGY(NAME, CT, SUF, PREC)                                                                        \
  /* out[4] = {E[cost], E[cost^2], min over support, max over support}; returns 0 on success */                 \
  int NAME##_nb(int n, const double* diag, int p, int n_beta, const double* angles, double* out, CT* work) {      \
    qref_diag D = {QREF_DIAG_F64, diag, 0.0, 1.0};                                                                \
    qref_circuit C = {n, QREF_INIT_PLUS, 0, n_beta == 1 ? QREF_MIXER_X : QREF_MIXER_XMULTI, 0, 0, NULL, 0, 1, 0}; \
    double o5[5];                                                                                                 \
    const int rc = run##SUF(&C, &D, p, angles, o5, work);                                                         \
    if (rc) return rc;                                                                                            \
    out[0] = o5[0]; out[1] = o5[1]; out[2] = o5[2]; out[3] = o5[3];                                               \
    return 0;                                                                                                     \
  }                                                                                                               \
  int NAME(int n, const double* diag, int p, const double* angles, double* out, CT* work) {                       \
    return NAME##_nb(n, diag, p, 1, angles, out, work);                                                           \
  }

DEFINE_ENERGY(qref_energy_c128, c128, _c128, 1)
DEFINE_ENERGY(qref_energy_c64, c64, _c64, 0)
