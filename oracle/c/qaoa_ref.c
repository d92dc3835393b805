/* CPU restatement (C + OpenMP) of the reference's QAOA energy path in the exact-statevector
 * limit  --  TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/qaoa_oracle.py for the numpy twin).
 *
 * It exists so that bench.py can time "the reference's CPU implementation of the path" on the GPU
 * box's host cores: the real reference delegates all amplitude arithmetic to qiskit-aer, which is
 * not installable here (no network; setup.py:22-24 pins only lower bounds), so this file restates
 * what Aer's statevector simulator is asked to do by qaoa/qaoa.py:418-520 for MaxCut + X mixer:
 *     |+>^n  ->  [ diag(exp(+i gamma cost(x)))  ->  RX(-2 beta) on every qubit ] x p  ->  <cost>
 * with cost(x) = GraphProblem.cost (qaoa/problems/graph_problem.py:152-176) for k = 2.
 * One sweep over the 2^n amplitudes per gate, OpenMP over amplitudes -- the same parallelisation
 * strategy a CPU statevector simulator uses.  PARITY: energies are "parity unpinned" against
 * Aer itself; this file is pinned against oracle/qaoa_oracle.py in tests/test_oracle.py.
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

/* angles: the reference's flat layout [(gamma_d, beta_d[0..n_beta-1]) for d < p], n_beta = 1 (X mixer,
 * x_mixer.py:35) or n (XMultiAngle, x_multiangle_mixer.py:40-50: parameter index = qubit index) */
#define DEFINE_ENERGY(NAME, CT, RT)                                                              \
  /* out[4] = {E[cost], E[cost^2], min over support, max over support}; returns 0 on success */  \
  int NAME##_nb(int n, const double* diag, int p, int n_beta, const double* angles, double* out, CT* work) { \
    const int64_t N = (int64_t)1 << n;                                                           \
    CT* psi = work ? work : (CT*)malloc(sizeof(CT) * (size_t)N);                                 \
    if (!psi) return 1;                                                                          \
    const RT a0 = (RT)pow(2.0, -0.5 * n);                                                        \
    _Pragma("omp parallel for schedule(static)")                                                 \
    for (int64_t x = 0; x < N; x++) { psi[x].re = a0; psi[x].im = 0; }                           \
    for (int d = 0; d < p; d++) {                                                                \
      const double gamma = angles[(size_t)d * (1 + n_beta)];                                     \
      const double* betas = angles + (size_t)d * (1 + n_beta) + 1;                               \
      _Pragma("omp parallel for schedule(static)")                                               \
      for (int64_t x = 0; x < N; x++) { /* phase separator, validation.py:43 sign */            \
        const RT c = (RT)cos(gamma * diag[x]), s = (RT)sin(gamma * diag[x]);                     \
        const RT re = psi[x].re, im = psi[x].im;                                                 \
        psi[x].re = re * c - im * s;                                                             \
        psi[x].im = re * s + im * c;                                                             \
      }                                                                                          \
      for (int q = 0; q < n; q++) { /* RX(-2 beta) = [[c, i s], [i s, c]], x_mixer.py:35 */      \
        const double beta = betas[n_beta == 1 ? 0 : q];                                          \
        const RT cb = (RT)cos(beta), sb = (RT)sin(beta);                                         \
        const int64_t bit = (int64_t)1 << q;                                                     \
        _Pragma("omp parallel for schedule(static)")                                             \
        for (int64_t i = 0; i < N / 2; i++) {                                                    \
          const int64_t x0 = ((i >> q) << (q + 1)) | (i & (bit - 1)), x1 = x0 | bit;             \
          const RT ar = psi[x0].re, ai = psi[x0].im, br = psi[x1].re, bi = psi[x1].im;           \
          psi[x0].re = cb * ar - sb * bi;                                                        \
          psi[x0].im = cb * ai + sb * br;                                                        \
          psi[x1].re = cb * br - sb * ai;                                                        \
          psi[x1].im = cb * bi + sb * ar;                                                        \
        }                                                                                        \
      }                                                                                          \
    }                                                                                            \
    double e = 0, e2 = 0, mn = 1e300, mx = -1e300;                                               \
    _Pragma("omp parallel for schedule(static) reduction(+:e,e2) reduction(min:mn) reduction(max:mx)") \
    for (int64_t x = 0; x < N; x++) {                                                            \
      const double pr = (double)psi[x].re * psi[x].re + (double)psi[x].im * psi[x].im;           \
      e += pr * diag[x];                                                                         \
      e2 += pr * diag[x] * diag[x];                                                              \
      if (pr > 0) { if (diag[x] < mn) mn = diag[x]; if (diag[x] > mx) mx = diag[x]; }            \
    }                                                                                            \
    out[0] = e; out[1] = e2; out[2] = mn; out[3] = mx;                                           \
    if (!work) free(psi);                                                                        \
    return 0;                                                                                    \
  }                                                                                              \
  int NAME(int n, const double* diag, int p, const double* angles, double* out, CT* work) {      \
    return NAME##_nb(n, diag, p, 1, angles, out, work);                                          \
  }

DEFINE_ENERGY(qref_energy_c128, c128, double)
DEFINE_ENERGY(qref_energy_c64, c64, float)
