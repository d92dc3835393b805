/* Per-precision kernels of the C oracle (included twice by qaoa_ref.c with RT / CT / FN defined).
 * TEST / BASELINE INFRASTRUCTURE ONLY.  Every function restates a piece of the reference path in the exact
 * statevector limit; citations are to /root/reference (qaoa/...).
 *
 * All state arithmetic is in the TRUE frame (no S frame here): the oracle must not share the engine's tricks.
 */

/* |+>^n (plus_initialstate.py:19-25) or |D^n_k> (dicke_initialstate.py:23-56: uniform over weight-k strings) */
static void FN(init_state)(CT* psi, int n, int init_kind, int k) {
  const int64_t N = (int64_t)1 << n;
  double amp;
  if (init_kind == QREF_INIT_PLUS) amp = pow(2.0, -0.5 * n);
  else amp = exp(-0.5 * (lgamma(n + 1.0) - lgamma(k + 1.0) - lgamma(n - k + 1.0)));
  const RT a = (RT)amp;
#pragma omp parallel for schedule(static)
  for (int64_t x = 0; x < N; x++) {
    psi[x].re = (init_kind == QREF_INIT_PLUS || __builtin_popcountll((unsigned long long)x) == k) ? a : (RT)0;
    psi[x].im = 0;
  }
}

/* phase separator psi(x) *= exp(+i gamma cost(x)) on [x0, x1)  (sign: qaoa/utils/validation.py:43;
 * qubo_problem.py:111-150, graph_problem.py:79-116 up to a global phase).  u8 diagonals use a 256-entry table
 * (cos, sin evaluated in double, like the plain path). */
static void FN(phase_range)(CT* psi, int64_t x0, int64_t x1, double gamma, const qref_diag* D, const RT* tab) {
  if (D->kind == QREF_DIAG_U8) {
    const uint8_t* d = (const uint8_t*)D->data;
    for (int64_t x = x0; x < x1; x++) {
      const RT c = tab[2 * d[x]], s = tab[2 * d[x] + 1];
      const RT re = psi[x].re, im = psi[x].im;
      psi[x].re = re * c - im * s;
      psi[x].im = re * s + im * c;
    }
  } else {
    const double* d = (const double*)D->data;
    for (int64_t x = x0; x < x1; x++) {
      const RT c = (RT)cos(gamma * d[x]), s = (RT)sin(gamma * d[x]);
      const RT re = psi[x].re, im = psi[x].im;
      psi[x].re = re * c - im * s;
      psi[x].im = re * s + im * c;
    }
  }
}

/* RX(-2 beta) = exp(+i beta X) = [[c, i s], [i s, c]] (x_mixer.py:35) between two amplitudes */
#define QREF_RX(A, B, cb, sb)                                   \
  do {                                                          \
    const RT ar = (A).re, ai = (A).im, br = (B).re, bi = (B).im; \
    (A).re = cb * ar - sb * bi;                                 \
    (A).im = cb * ai + sb * br;                                 \
    (B).re = cb * br - sb * ai;                                 \
    (B).im = cb * bi + sb * ar;                                 \
  } while (0)

/* One X / XMultiAngle layer, cache blocked.  Gates on different qubits commute and the phase separator is diagonal,
 * so the layer may be executed block by block: (1) phase + the LB lowest qubits on contiguous blocks of 2^LB
 * amplitudes (one sweep), (2) the remaining qubits in groups of <= 5: 2^k rows of a contiguous chunk at a time (one
 * sweep per group) -- the same grouping a gate-fusing CPU simulator applies (Aer fuses up to 5 qubits per sweep). */
static void FN(x_layer_blocked)(CT* psi, int n, double gamma, const double* betas, int n_beta, const qref_diag* D) {
  const int64_t N = (int64_t)1 << n;
  const int LB = n < 14 ? n : 14;
  RT tab[512];
  if (D->kind == QREF_DIAG_U8)
    for (int k = 0; k < 256; k++) {
      tab[2 * k] = (RT)cos(gamma * (D->c0 + D->delta * k));
      tab[2 * k + 1] = (RT)sin(gamma * (D->c0 + D->delta * k));
    }
  RT cb[64], sb[64];
  for (int q = 0; q < n; q++) {
    const double b = betas[n_beta == 1 ? 0 : q];
    cb[q] = (RT)cos(b);
    sb[q] = (RT)sin(b);
  }
  const int64_t nblk = N >> LB, BL = (int64_t)1 << LB;
#pragma omp parallel for schedule(static)
  for (int64_t blk = 0; blk < nblk; blk++) {
    CT* p = psi + blk * BL;
    FN(phase_range)(psi, blk * BL, (blk + 1) * BL, gamma, D, tab);
    for (int q = 0; q < LB; q++) {
      const int64_t bit = (int64_t)1 << q;
      const RT c = cb[q], s = sb[q];
      for (int64_t b0 = 0; b0 < BL; b0 += 2 * bit)
        for (int64_t i = 0; i < bit; i++) QREF_RX(p[b0 + i], p[b0 + bit + i], c, s);
    }
  }
  const int nh = n - LB;
  if (nh <= 0) return;
  const int ngroups = (nh + 4) / 5;
  int q0 = LB;
  for (int g = 0; g < ngroups; g++) {
    const int k = (nh - (q0 - LB) + (ngroups - g) - 1) / (ngroups - g);   /* spread the qubits evenly */
    const int64_t CH = 512;                                               /* amplitudes of a row handled at a time */
    const int64_t row = (int64_t)1 << q0;                                 /* stride between rows */
    const int64_t nouter = N >> (q0 + k);                                 /* values of the bits above the group */
    const int64_t nch = row / CH;
#pragma omp parallel for schedule(static) collapse(2)
    for (int64_t o = 0; o < nouter; o++)
      for (int64_t ch = 0; ch < nch; ch++) {
        CT* base = psi + (o << (q0 + k)) + ch * CH;
        for (int j = 0; j < k; j++) {
          const RT c = cb[q0 + j], s = sb[q0 + j];
          const int64_t jb = (int64_t)1 << j;
          for (int64_t r = 0; r < ((int64_t)1 << k); r++) {
            if (r & jb) continue;
            CT* a = base + r * row;
            CT* b = base + (r | jb) * row;
            for (int64_t i = 0; i < CH; i++) QREF_RX(a[i], b[i], c, s);
          }
        }
      }
    q0 += k;
  }
}

/* the same layer, one sweep per gate (the unfused form the first version of this oracle had; pins the blocked one) */
static void FN(x_layer_plain)(CT* psi, int n, double gamma, const double* betas, int n_beta, const qref_diag* D) {
  const int64_t N = (int64_t)1 << n;
  RT tab[512];
  if (D->kind == QREF_DIAG_U8)
    for (int k = 0; k < 256; k++) {
      tab[2 * k] = (RT)cos(gamma * (D->c0 + D->delta * k));
      tab[2 * k + 1] = (RT)sin(gamma * (D->c0 + D->delta * k));
    }
  const int64_t CHK = 1 << 12;
#pragma omp parallel for schedule(static)
  for (int64_t c0 = 0; c0 < N; c0 += CHK) FN(phase_range)(psi, c0, c0 + CHK < N ? c0 + CHK : N, gamma, D, tab);
  for (int q = 0; q < n; q++) {
    const double beta = betas[n_beta == 1 ? 0 : q];
    const RT c = (RT)cos(beta), s = (RT)sin(beta);
    const int64_t bit = (int64_t)1 << q;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N / 2; i++) {
      const int64_t x0 = ((i >> q) << (q + 1)) | (i & (bit - 1));
      QREF_RX(psi[x0], psi[x0 | bit], c, s);
    }
  }
}

static void FN(phase_all)(CT* psi, int n, double gamma, const qref_diag* D) {
  const int64_t N = (int64_t)1 << n;
  RT tab[512];
  if (D->kind == QREF_DIAG_U8)
    for (int k = 0; k < 256; k++) {
      tab[2 * k] = (RT)cos(gamma * (D->c0 + D->delta * k));
      tab[2 * k + 1] = (RT)sin(gamma * (D->c0 + D->delta * k));
    }
  const int64_t CHK = 1 << 12;
#pragma omp parallel for schedule(static)
  for (int64_t c0 = 0; c0 < N; c0 += CHK) FN(phase_range)(psi, c0, c0 + CHK < N ? c0 + CHK : N, gamma, D, tab);
}

/* Grover mixer  psi <- psi + (e^{-i beta} - 1) <s|psi> s   (grover_mixer.py:43-82), |s> = |+>^n or |D^n_k> */
static void FN(grover)(CT* psi, int n, int sub_kind, int sub_k, double beta) {
  const int64_t N = (int64_t)1 << n;
  double amp;
  if (sub_kind == QREF_INIT_PLUS) amp = pow(2.0, -0.5 * n);
  else amp = exp(-0.5 * (lgamma(n + 1.0) - lgamma(sub_k + 1.0) - lgamma(n - sub_k + 1.0)));
  double dr = 0.0, di = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : dr, di)
  for (int64_t x = 0; x < N; x++)
    if (sub_kind == QREF_INIT_PLUS || __builtin_popcountll((unsigned long long)x) == sub_k) {
      dr += psi[x].re;
      di += psi[x].im;
    }
  dr *= amp;
  di *= amp;
  const double kr = cos(beta) - 1.0, ki = -sin(beta);
  const RT fr = (RT)((kr * dr - ki * di) * amp), fi = (RT)((kr * di + ki * dr) * amp);
#pragma omp parallel for schedule(static)
  for (int64_t x = 0; x < N; x++)
    if (sub_kind == QREF_INIT_PLUS || __builtin_popcountll((unsigned long long)x) == sub_k) {
      psi[x].re += fr;
      psi[x].im += fi;
    }
}

/* XXPlusYYGate(beta / 2) on (q0, q1) (xy_mixer.py:54-56): on span{|q0=1,q1=0>, |q0=0,q1=1>} the rotation
 * [[cos(beta/4), -i sin(beta/4)], [-i sin(beta/4), cos(beta/4)]]; `ang` = beta / 4 (per trotter step) */
static void FN(xy_gate)(CT* psi, int n, int q0, int q1, double ang) {
  const int64_t N = (int64_t)1 << n;
  const RT c = (RT)cos(ang), s = (RT)sin(ang);
  const int lo = q0 < q1 ? q0 : q1, hi = q0 < q1 ? q1 : q0;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < N / 4; i++) {
    const int64_t t = ((i >> lo) << (lo + 1)) | (i & (((int64_t)1 << lo) - 1));
    const int64_t base = ((t >> hi) << (hi + 1)) | (t & (((int64_t)1 << hi) - 1));
    const int64_t xa = base | ((int64_t)1 << q0), xb = base | ((int64_t)1 << q1);
    const RT ar = psi[xa].re, ai = psi[xa].im, br = psi[xb].re, bi = psi[xb].im;
    psi[xa].re = c * ar + s * bi;   /* a' = c a - i s b */
    psi[xa].im = c * ai - s * br;
    psi[xb].re = c * br + s * ai;   /* b' = -i s a + c b */
    psi[xb].im = c * bi - s * ar;
  }
}

/* exact analogue of Statistic (statistic.py:57-142): E[c], E[c^2], min / max over the support, norm */
static void FN(stats)(const CT* psi, int n, const qref_diag* D, double* out5) {
  const int64_t N = (int64_t)1 << n;
  double e0 = 0, e = 0, e2 = 0, mn = 1e300, mx = -1e300;
#pragma omp parallel for schedule(static) reduction(+ : e0, e, e2) reduction(min : mn) reduction(max : mx)
  for (int64_t x = 0; x < N; x++) {
    const double pr = (double)psi[x].re * psi[x].re + (double)psi[x].im * psi[x].im;
    const double c = D->kind == QREF_DIAG_U8 ? D->c0 + D->delta * ((const uint8_t*)D->data)[x] : ((const double*)D->data)[x];
    e0 += pr;
    e += pr * c;
    e2 += pr * c * c;
    if (pr > 0) { if (c < mn) mn = c; if (c > mx) mx = c; }
  }
  out5[0] = e / e0; out5[1] = e2 / e0; out5[2] = mn; out5[3] = mx; out5[4] = e0;
}

/* init -> [phase(gamma_d) -> mixer(beta_d) / trotter x trotter] x p -> stats   (qaoa/qaoa.py:418-520 layer order) */
static int FN(run)(const qref_circuit* C, const qref_diag* D, int p, const double* angles, double* out5, CT* keep) {
  const int n = C->n;
  const int64_t N = (int64_t)1 << n;
  CT* psi = keep ? keep : (CT*)malloc(sizeof(CT) * (size_t)N);
  if (!psi) return 1;
  FN(init_state)(psi, n, C->init_kind, C->init_k);
  const int n_beta = C->mixer == QREF_MIXER_XMULTI ? n : 1;
  for (int d = 0; d < p; d++) {
    const double gamma = angles[(size_t)d * (1 + n_beta)];
    const double* betas = angles + (size_t)d * (1 + n_beta) + 1;
    if (C->mixer == QREF_MIXER_X || C->mixer == QREF_MIXER_XMULTI) {
      if (C->blocked) FN(x_layer_blocked)(psi, n, gamma, betas, n_beta, D);
      else FN(x_layer_plain)(psi, n, gamma, betas, n_beta, D);
    } else if (C->mixer == QREF_MIXER_GROVER) {
      FN(phase_all)(psi, n, gamma, D);
      FN(grover)(psi, n, C->sub_kind, C->sub_k, betas[0]);   /* invariant under trotterisation */
    } else {
      FN(phase_all)(psi, n, gamma, D);
      for (int rep = 0; rep < C->trotter; rep++)             /* qaoa/qaoa.py:494-503: beta / trotter, repeated */
        for (int e = 0; e < C->n_pairs; e++)
          FN(xy_gate)(psi, n, C->pairs[2 * e], C->pairs[2 * e + 1], 0.25 * betas[0] / C->trotter);
    }
  }
  FN(stats)(psi, n, D, out5);
  if (!keep) free(psi);
  return 0;
}
