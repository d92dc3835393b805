// Cost-diagonal generators, the single-CTA fused kernel for small registers, the element-wise
// Grover / XY passes and the small helper kernels (phase tables, partial-sum finalisation).
#pragma once
#include "common.cuh"

// ============================================================================================
// cost-diagonal generators (natural index order)
// ============================================================================================
// Replaces GraphProblem.cost (reference qaoa/problems/graph_problem.py:152-176) evaluated for
// every basis index.  lut[] maps the b-bit integer (bit j = qubit v*b+j) to a colour id; the host
// builds it from the reference's string tables (maxkcut_binary_powertwo.py:111-142,
// maxkcut_binary_fullH.py:119-192).  Nodes >= n_free carry the fixed colour (fix_one_node).
struct MaxKCutDev {
  int n_edges;
  const int* eu;
  const int* ev;
  const double* ew;   // weights as given
  const int* ewi;     // integer weights (valid when integer path is used)
  int bits;
  int n_free;
  int fixed_colour;
  unsigned char lut[8];
};

__device__ __forceinline__ int node_colour(const MaxKCutDev& D, unsigned long long x, int v) {
  if (v >= D.n_free) return D.fixed_colour;
  return D.lut[(x >> (v * D.bits)) & ((1u << D.bits) - 1u)];
}

// x0: first basis index of the slice this engine holds (sharded registers; 0 otherwise)
// tw_mask != 0: twisted shard coordinates (tile_kernel.cuh): local indices with bit 0 set belong to the basis index
// whose rank bits (at tw_shift) are complemented.
__global__ void gen_maxkcut_kernel(MaxKCutDev D, size_t N, size_t x0, DiagInfo info, void* out, int tw_shift = 0,
                                   unsigned tw_mask = 0) {
  for (size_t xl = blockIdx.x * (size_t)blockDim.x + threadIdx.x; xl < N; xl += (size_t)gridDim.x * blockDim.x) {
    size_t x = x0 + xl;
    if (tw_mask != 0 && (xl & 1)) x ^= (size_t)tw_mask << tw_shift;
    if (info.kind == DIAG_U8) {
      int s = 0;
      for (int e = 0; e < D.n_edges; e++) {
        int cu = node_colour(D, x, __ldg(D.eu + e)), cv = node_colour(D, x, __ldg(D.ev + e));
        s += (cu != cv) ? __ldg(D.ewi + e) : 0;
      }
      ((uint8_t*)out)[xl] = (uint8_t)(s - (int)llrint(info.c0));
    } else {
      double s = 0.0;
      for (int e = 0; e < D.n_edges; e++) {
        int cu = node_colour(D, x, __ldg(D.eu + e)), cv = node_colour(D, x, __ldg(D.ev + e));
        s += (cu != cv) ? __ldg(D.ew + e) : 0.0;   // same edge order as the reference's sum()
      }
      if (info.kind == DIAG_U32FX) ((uint32_t*)out)[xl] = (uint32_t)llrint((s - info.c0) / info.delta);
      else ((double*)out)[xl] = s;
    }
  }
}

// Replaces QUBO.qubo_cost (qubo_problem.py:101-108): cost = -(x^T Q x + c^T x + b).
// d[i] = Q_ii + c_i ; M[i*n+j] = Q_ij + Q_ji for j < i (host-prepared).
__global__ void gen_qubo_kernel(int n, const double* __restrict__ d, const double* __restrict__ M,
                                double b, size_t N, size_t x0, DiagInfo info, void* out) {
  extern __shared__ double sm[];
  double* sd = sm;
  double* sM = sm + n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) sd[i] = d[i];
  for (int i = threadIdx.x; i < n * n; i += blockDim.x) sM[i] = M[i];
  __syncthreads();
  for (size_t xl = blockIdx.x * (size_t)blockDim.x + threadIdx.x; xl < N; xl += (size_t)gridDim.x * blockDim.x) {
    double f = b;
    unsigned long long rem = x0 + xl;
    while (rem) {
      int i = 63 - __clzll((long long)rem);
      rem &= ~(1ull << i);
      double row = sd[i];
      unsigned long long r2 = rem;  // bits j < i that are set
      while (r2) {
        int j = 63 - __clzll((long long)r2);
        r2 &= ~(1ull << j);
        row += sM[i * n + j];
      }
      f += row;
    }
    double cost = -f;
    if (info.kind == DIAG_U8) ((uint8_t*)out)[xl] = (uint8_t)llrint((cost - info.c0) / info.delta);
    else if (info.kind == DIAG_U32FX) ((uint32_t*)out)[xl] = (uint32_t)llrint((cost - info.c0) / info.delta);
    else ((double*)out)[xl] = cost;
  }
}

// host-provided diagonal (escape hatch): quantise a chunk of doubles
__global__ void quantise_diag_kernel(const double* __restrict__ src, size_t count, size_t offset,
                                     DiagInfo info, void* out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
    double c = src[i];
    size_t x = offset + i;
    if (info.kind == DIAG_U8) ((uint8_t*)out)[x] = (uint8_t)llrint((c - info.c0) / info.delta);
    else if (info.kind == DIAG_U32FX) ((uint32_t*)out)[x] = (uint32_t)llrint((c - info.c0) / info.delta);
    else ((double*)out)[x] = c;
  }
}

__global__ void read_diag_kernel(const void* diag, DiagInfo info, size_t offset, size_t count, double* out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
    out[i] = diag_decode(info, diag, offset + i);
}

// min / max of the diagonal (computeMinMaxCosts, base_problem.py:121-134, without the Python loop)
// cost(x) == cost(~x) for every x?  (raw stored values; flag |= 1 on the first mismatch) -- decides whether the
// tile path may store the symmetric half of the register only
template <typename T>
__global__ void diag_symmetry_kernel(const T* __restrict__ diag, size_t N, int* flag) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  int bad = 0;
  for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < N / 2; x += stride)
    if (diag[x] != diag[N - 1 - x]) bad = 1;
  if (bad) atomicOr(flag, 1);
}

__global__ void diag_minmax_kernel(const void* diag, DiagInfo info, size_t N, double* partials) {
  __shared__ double red[32 * 5];
  double v[5] = {0, 0, 0, 1e300, -1e300};
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < N; i += (size_t)gridDim.x * blockDim.x) {
    double c = diag_decode(info, diag, i);
    v[3] = fmin(v[3], c);
    v[4] = fmax(v[4], c);
  }
  block_reduce5(v, red);
  if (threadIdx.x == 0) { partials[blockIdx.x * 2] = v[3]; partials[blockIdx.x * 2 + 1] = v[4]; }
}

// ============================================================================================
// phase tables and result finalisation
// ============================================================================================
// table[t][k] = sigma_t * exp(i gamma_t (c0 + delta k)),  k = 0..255
template <typename real, typename cplx>
__global__ void build_tables_kernel(const double* __restrict__ gam_sig, double c0, double delta, cplx* out) {
  const int t = blockIdx.x, k = threadIdx.x;
  const double gamma = gam_sig[2 * t], sigma = gam_sig[2 * t + 1];
  double s, c;
  sincos(gamma * (c0 + delta * (double)k), &s, &c);
  cplx r;
  r.x = (real)(sigma * c);
  r.y = (real)(sigma * s);
  out[(size_t)t * 256 + k] = r;
}

// out[point] = {E[cost], E[cost^2], min, max, norm}; partials hold {S0, S1, S2, kmin, kmax} with
// cost = c0 + delta * k.  Fixed summation order => run-to-run reproducible.
__global__ void finalize_kernel(const double* __restrict__ partials, int ntiles, DiagInfo info, double* out) {
  __shared__ double red[32 * 5];
  const double* p = partials + (size_t)blockIdx.x * ntiles * 5;
  double v[5] = {0, 0, 0, 1e300, -1e300};
  for (int i = threadIdx.x; i < ntiles; i += blockDim.x) {
    v[0] += p[i * 5 + 0];
    v[1] += p[i * 5 + 1];
    v[2] += p[i * 5 + 2];
    v[3] = fmin(v[3], p[i * 5 + 3]);
    v[4] = fmax(v[4], p[i * 5 + 4]);
  }
  block_reduce5(v, red);
  if (threadIdx.x == 0) {
    const double c0 = info.kind == DIAG_F64 ? 0.0 : info.c0, dl = info.kind == DIAG_F64 ? 1.0 : info.delta;
    const double m1 = v[1] / v[0], m2 = v[2] / v[0];
    double* o = out + (size_t)blockIdx.x * 5;
    o[0] = c0 + dl * m1;
    o[1] = c0 * c0 + 2.0 * c0 * dl * m1 + dl * dl * m2;
    double lo = c0 + dl * v[3], hi = c0 + dl * v[4];
    o[2] = fmin(lo, hi);
    o[3] = fmax(lo, hi);
    o[4] = v[0];
  }
}

// Sharded registers: this rank's share of the sums in TRUE cost units, to be combined across ranks
// (sum, sum, sum, min, max):  out[point] = {sum p, sum p c, sum p c^2, min c, max c}.
__global__ void finalize_sums_kernel(const double* __restrict__ partials, int nparts, DiagInfo info, double* out) {
  __shared__ double red[32 * 5];
  const double* p = partials + (size_t)blockIdx.x * nparts * 5;
  double v[5] = {0, 0, 0, 1e300, -1e300};
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) {
    v[0] += p[i * 5 + 0];
    v[1] += p[i * 5 + 1];
    v[2] += p[i * 5 + 2];
    v[3] = fmin(v[3], p[i * 5 + 3]);
    v[4] = fmax(v[4], p[i * 5 + 4]);
  }
  block_reduce5(v, red);
  if (threadIdx.x == 0) {
    const double c0 = info.kind == DIAG_F64 ? 0.0 : info.c0, dl = info.kind == DIAG_F64 ? 1.0 : info.delta;
    double* o = out + (size_t)blockIdx.x * 5;
    o[0] = v[0];
    o[1] = c0 * v[0] + dl * v[1];
    o[2] = c0 * c0 * v[0] + 2.0 * c0 * dl * v[1] + dl * dl * v[2];
    const double lo = c0 + dl * v[3], hi = c0 + dl * v[4];
    o[3] = v[3] > 1e299 ? 1e300 : fmin(lo, hi);
    o[4] = v[4] < -1e299 ? -1e300 : fmax(lo, hi);
  }
}

// ============================================================================================
// initial / Grover reference states in the S frame
// ============================================================================================
struct StateGen {
  int kind;       // INIT_PLUS / INIT_DICKE / INIT_STATEVECTOR
  int k;          // Dicke weight
  double amp;     // 2^{-n/2} or C(n,k)^{-1/2}
  const void* vec;  // S-frame vector for INIT_STATEVECTOR (same element type as the state)
  int pc0;        // sharded registers: popcount of the rank bits (the basis index is x0 + local index)
};

template <typename real>
__device__ __forceinline__ void gen_amp(const StateGen& g, size_t x, real& re, real& im) {
  if (g.kind == INIT_STATEVECTOR) {
    re = ((const real*)g.vec)[2 * x];
    im = ((const real*)g.vec)[2 * x + 1];
    return;
  }
  if (g.kind == INIT_UNIFORM) {   // compact feasible-subspace register: every entry carries the same amplitude
    re = (real)g.amp;
    im = (real)0;
    return;
  }
  int pc = __popcll((unsigned long long)x) + g.pc0;
  re = (real)g.amp;
  im = (real)0;
  if (g.kind == INIT_DICKE && pc != g.k) re = (real)0;
  mul_i_pow(re, im, pc);
}

// psi~(x) = i^popc(x) psi(x)   (in place, natural order, interleaved re/im)
template <typename real>
__global__ void to_sframe_kernel(real* st, size_t N) {
  for (size_t x = blockIdx.x * (size_t)blockDim.x + threadIdx.x; x < N; x += (size_t)gridDim.x * blockDim.x) {
    real re = st[2 * x], im = st[2 * x + 1];
    mul_i_pow(re, im, __popcll((unsigned long long)x));
    st[2 * x] = re;
    st[2 * x + 1] = im;
  }
}

// X gates between two layers (reference qaoa/qaoa.py:505-509, flip=True), natural order, S frame on both sides:
//   psi'(x) = psi(x ^ m)   <=>   psi~'(x) = i^(popc(x) - popc(x ^ m)) psi~(x ^ m) = (-1)^popc(x & m) i^(-popc(m)) psi~(x ^ m).
// `src` is a KEPT state: the true amplitude is src(y) * scale * (-1)^popc(y & zmask) (what qaoa_read_statevector undoes).
template <typename real>
__global__ void flip_continue_kernel(const real* __restrict__ src, real* __restrict__ dst, size_t N,
                                     unsigned long long m, unsigned long long zmask, real scale, int rot) {
  for (size_t x = blockIdx.x * (size_t)blockDim.x + threadIdx.x; x < N; x += (size_t)gridDim.x * blockDim.x) {
    const size_t y = x ^ (size_t)m;
    real re = src[2 * y], im = src[2 * y + 1];
    const int odd = (__popcll((unsigned long long)y & zmask) + __popcll((unsigned long long)x & m)) & 1;
    const real f = odd ? -scale : scale;
    re *= f;
    im *= f;
    mul_i_pow(re, im, rot);
    dst[2 * x] = re;
    dst[2 * x + 1] = im;
  }
}

// ============================================================================================
// single-CTA fused kernel: whole circuit for n <= 14 (complex64) / 13 (complex128) in shared memory
// ============================================================================================
struct SmallDesc {
  int n;
  int p;
  int mixer;       // MixerKind
  int n_beta;      // parameters per layer for the mixer (1 or n)
  int n_init;      // leading initial-state parameters of the angle vector (PlusParameterized phases)
  int n_gamma;     // parameters per layer of the phase separator: 1, or the number of cost terms
  const double* terms;   // n_gamma > 1: term diagonals [n_gamma][2^n]; the phase is exp(i sum_k gamma_k terms_k)
  int n_pairs;     // XY
  int trotter;     // XY only (X and Grover are invariant under trotterisation)
  int write_state; // store the final S-frame state to global memory
  StateGen init;
  StateGen sub;    // Grover |s>
  NodeVec node;    // MIXER_NODE_GROVER: |s_node>
};

// One node register of the tensorised Grover mixer: for every group of 2^b amplitudes that differ only in the bits
// [shift, shift + b):  psi_j += (e^{-i beta} - 1) <s_node|psi_group> s_node[j]   (k = e^{-i beta} - 1 = kr + i ki).
// `ps`: interleaved (re, im), S frame; works on shared or global memory.
template <typename real>
__device__ __forceinline__ void node_grover_group(real* ps, size_t base, int shift, const NodeVec& nv, double kr, double ki) {
  const int K = 1 << nv.bits;
  double dr = 0.0, di = 0.0;
  for (int j = 0; j < K; j++) {
    const size_t x = base | ((size_t)j << shift);
    const double re = ps[2 * x], im = ps[2 * x + 1];
    dr += nv.re[j] * re + nv.im[j] * im;   // conj(s) * psi
    di += nv.re[j] * im - nv.im[j] * re;
  }
  const double fr = kr * dr - ki * di, fi = kr * di + ki * dr;
  for (int j = 0; j < K; j++) {
    const size_t x = base | ((size_t)j << shift);
    ps[2 * x] += (real)(fr * nv.re[j] - fi * nv.im[j]);
    ps[2 * x + 1] += (real)(fr * nv.im[j] + fi * nv.re[j]);
  }
}

// large registers: one sweep of the state per node register
template <typename real>
__global__ void node_grover_kernel(real* __restrict__ state, size_t N, int shift, NodeVec nv, double kr, double ki) {
  const size_t groups = N >> nv.bits;
  const size_t low = ((size_t)1 << shift) - 1;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < groups; i += (size_t)gridDim.x * blockDim.x) {
    const size_t base = ((i & ~low) << nv.bits) | (i & low);
    node_grover_group<real>(state, base, shift, nv, kr, ki);
  }
}

template <typename real>
__global__ void __launch_bounds__(512, 1)
small_fused_kernel(SmallDesc D, const int* __restrict__ pairs, const double* __restrict__ angles,
                   int angles_stride, const void* __restrict__ diag, DiagInfo info,
                   real* __restrict__ state_out, double* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  real* ps = reinterpret_cast<real*>(smem_raw);  // interleaved re, im
  __shared__ double red[16 * 5];
  __shared__ double dotsh[2];
  const int n = D.n;
  const size_t N = (size_t)1 << n;
  const int tid = threadIdx.x, nt = blockDim.x;
  const double* ang0 = angles + (size_t)blockIdx.x * angles_stride;
  const double* ang = ang0 + D.n_init;   // flat layout: initial-state parameters first (qaoa/qaoa.py:937-990)
  const int ng = D.n_gamma, per = ng + D.n_beta;

  for (size_t x = tid; x < N; x += nt) {
    real re, im;
    gen_amp<real>(D.init, x, re, im);
    if (D.n_init > 0) {
      // PlusParameterized (plus_parameterized_initialstate.py:44-64): RZ(phi_q) = diag(e^{-i phi/2}, e^{+i phi/2})
      // on qubit q < n_init after the Hadamards
      double theta = 0.0;
      for (int q = 0; q < D.n_init; q++) theta += ang0[q] * (((x >> q) & 1) ? 0.5 : -0.5);
      double sn, cs;
      sincos(theta, &sn, &cs);
      const real r2 = re * (real)cs - im * (real)sn;
      im = re * (real)sn + im * (real)cs;
      re = r2;
    }
    ps[2 * x] = re;
    ps[2 * x + 1] = im;
  }
  __syncthreads();

  for (int d = 0; d < D.p; d++) {
    const double gamma = ang[d * per];
    // phase separator  e^{+i gamma cost(x)}   (sign: qaoa/utils/validation.py:43); with several cost terms
    // (one gamma per edge / edge orbit, maxcut_free_problem.py:66-73, maxcut_orbit_problem.py:139-144)
    // the angle is sum_k gamma_k term_k(x)
    for (size_t x = tid; x < N; x += nt) {
      double s, c;
      double theta;
      if (ng > 1) {
        theta = 0.0;
        for (int k = 0; k < ng; k++) theta += ang[d * per + k] * D.terms[(size_t)k * N + x];
      } else theta = gamma * diag_decode(info, diag, x);
      sincos(theta, &s, &c);
      real re = ps[2 * x], im = ps[2 * x + 1];
      ps[2 * x] = re * (real)c - im * (real)s;
      ps[2 * x + 1] = re * (real)s + im * (real)c;
    }
    __syncthreads();
    if (D.mixer == MIXER_X || D.mixer == MIXER_XMULTI) {
      for (int q = 0; q < n; q++) {
        const double beta = ang[d * per + ng + (D.n_beta == 1 ? 0 : q)];
        double sd, cd;
        sincos(beta, &sd, &cd);
        const real c = (real)cd, s = (real)sd;
        const size_t bit = (size_t)1 << q;
        for (size_t i = tid; i < N / 2; i += nt) {
          size_t x0 = ((i >> q) << (q + 1)) | (i & (bit - 1));
          size_t x1 = x0 | bit;
          real ar = ps[2 * x0], ai = ps[2 * x0 + 1], br = ps[2 * x1], bi = ps[2 * x1 + 1];
          ps[2 * x0] = c * ar + s * br;   // S frame: real rotation [[c, s], [-s, c]]
          ps[2 * x0 + 1] = c * ai + s * bi;
          ps[2 * x1] = c * br - s * ar;
          ps[2 * x1 + 1] = c * bi - s * ai;
        }
        __syncthreads();
      }
    } else if (D.mixer == MIXER_XY) {
      // XXPlusYYGate(0.5 beta): rotation by beta/4 on span{|01>,|10>} (xy_mixer.py:54-56), in order
      const double beta = ang[d * per + ng] / D.trotter;
      double sd, cd;
      sincos(0.25 * beta, &sd, &cd);
      const real c = (real)cd, s = (real)sd;
      for (int rep = 0; rep < D.trotter; rep++) {
        for (int e = 0; e < D.n_pairs; e++) {
          const int q0 = pairs[2 * e], q1 = pairs[2 * e + 1];
          const int lo = q0 < q1 ? q0 : q1, hi = q0 < q1 ? q1 : q0;
          for (size_t i = tid; i < N / 4; i += nt) {
            size_t t = ((i >> lo) << (lo + 1)) | (i & (((size_t)1 << lo) - 1));
            size_t base = ((t >> hi) << (hi + 1)) | (t & (((size_t)1 << hi) - 1));
            size_t xa = base | ((size_t)1 << q0), xb = base | ((size_t)1 << q1);
            real ar = ps[2 * xa], ai = ps[2 * xa + 1], br = ps[2 * xb], bi = ps[2 * xb + 1];
            // a' = c a - i s b ; b' = -i s a + c b
            ps[2 * xa] = c * ar + s * bi;
            ps[2 * xa + 1] = c * ai - s * br;
            ps[2 * xb] = c * br + s * ai;
            ps[2 * xb + 1] = c * bi - s * ar;
          }
          __syncthreads();
        }
      }
    } else if (D.mixer == MIXER_NODE_GROVER) {
      // Grover mixer per node register, all nodes sharing beta (maxkcut_grover_mixer.py:116-124 with grover_mixer.py:43-82)
      const double beta = ang[d * per + ng];
      double sb, cb;
      sincos(beta, &sb, &cb);
      const int b = D.node.bits;
      for (int v = 0; v < n / b; v++) {
        const int shift = v * b;
        const size_t low = ((size_t)1 << shift) - 1;
        for (size_t i = tid; i < (N >> b); i += nt)
          node_grover_group<real>(ps, ((i & ~low) << b) | (i & low), shift, D.node, cb - 1.0, -sb);
        __syncthreads();
      }
    } else {  // MIXER_GROVER:  psi += (e^{-i beta} - 1) <s|psi> s   (grover_mixer.py:43-82)
      const double beta = ang[d * per + ng];
      double v[5] = {0, 0, 0, 1e300, -1e300};
      for (size_t x = tid; x < N; x += nt) {
        real sr, si;
        gen_amp<real>(D.sub, x, sr, si);
        double re = ps[2 * x], im = ps[2 * x + 1];
        v[0] += (double)sr * re + (double)si * im;   // conj(s) * psi
        v[1] += (double)sr * im - (double)si * re;
      }
      block_reduce5(v, red);
      if (tid == 0) { dotsh[0] = v[0]; dotsh[1] = v[1]; }
      __syncthreads();
      double sb, cb;
      sincos(beta, &sb, &cb);
      const double kr = cb - 1.0, ki = -sb;
      const double fr = kr * dotsh[0] - ki * dotsh[1], fi = kr * dotsh[1] + ki * dotsh[0];
      for (size_t x = tid; x < N; x += nt) {
        real sr, si;
        gen_amp<real>(D.sub, x, sr, si);
        ps[2 * x] += (real)(fr * sr - fi * si);
        ps[2 * x + 1] += (real)(fr * si + fi * sr);
      }
      __syncthreads();
    }
  }

  double v[5] = {0, 0, 0, 1e300, -1e300};
  for (size_t x = tid; x < N; x += nt) {
    double re = ps[2 * x], im = ps[2 * x + 1];
    double pr = re * re + im * im;
    double c = diag_decode(info, diag, x);
    v[0] += pr;
    v[1] += pr * c;
    v[2] += pr * c * c;
    if (pr > 0.0) { v[3] = fmin(v[3], c); v[4] = fmax(v[4], c); }
    if (D.write_state) {
      state_out[2 * x] = (real)re;
      state_out[2 * x + 1] = (real)im;
    }
  }
  block_reduce5(v, red);
  if (tid == 0) {
    double* o = out + (size_t)blockIdx.x * 5;
    o[0] = v[1] / v[0];
    o[1] = v[2] / v[0];
    o[2] = v[3];
    o[3] = v[4];
    o[4] = v[0];
  }
}

// ============================================================================================
// element-wise passes for large registers: Grover mixer and XY mixer (natural order, S frame)
// ============================================================================================
template <typename real> struct Real2;
template <> struct Real2<float> { typedef float2 type; };
template <> struct Real2<double> { typedef double2 type; };

enum EwFlags : int {
  EW_INIT = 1, EW_AXPY = 2, EW_PHASE = 4, EW_DOT = 8, EW_REDUCE = 16, EW_STORE = 32
};

// One sweep:  v = init|load ; v += coefdot * s ; v *= e^{i gamma c} ; dot += conj(s) v ; stats ; store
// partial layout per block: {dot_re, dot_im, -, -, -} (EW_DOT) or {S0,S1,S2,min,max} (EW_REDUCE)
template <typename real>
__global__ void __launch_bounds__(256)
elementwise_pass_kernel(real* __restrict__ st, size_t N, int flags, StateGen init, StateGen sub,
                        const double* __restrict__ coefdot, double gamma, const void* __restrict__ diag,
                        DiagInfo info, double* __restrict__ part_dot, double* __restrict__ part_red) {
  __shared__ double red[8 * 5];
  double fr = 0.0, fi = 0.0;
  if (flags & EW_AXPY) { fr = coefdot[0]; fi = coefdot[1]; }
  double dv[5] = {0, 0, 0, 1e300, -1e300};
  double rv[5] = {0, 0, 0, 1e300, -1e300};
  // four elements per thread and iteration, loads first: enough bytes in flight to cover the DRAM latency
  typedef typename Real2<real>::type real2;
  const real2* st2 = reinterpret_cast<const real2*>(st);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t x0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x; x0 < N; x0 += 4 * stride) {
    real2 in[4];
    if (!(flags & EW_INIT)) {
#pragma unroll
      for (int k = 0; k < 4; k++) if (x0 + k * stride < N) in[k] = st2[x0 + k * stride];
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
    const size_t x = x0 + k * stride;
    if (x >= N) break;
    real re, im;
    if (flags & EW_INIT) gen_amp<real>(init, x, re, im);
    else { re = in[k].x; im = in[k].y; }
    real sr = 0, si = 0;
    if (flags & (EW_AXPY | EW_DOT)) gen_amp<real>(sub, x, sr, si);
    if (flags & EW_AXPY) {
      re += (real)(fr * sr - fi * si);
      im += (real)(fr * si + fi * sr);
    }
    double c = 0.0;
    if (flags & (EW_PHASE | EW_REDUCE)) c = diag_decode(info, diag, x);
    if (flags & EW_PHASE) {
      real s, co;
      if (sizeof(real) == 4) {   // range reduction in double (turns), sine / cosine in single precision
        const double turns = gamma * 0.15915494309189533576888 * c;
        float sf, cf;
        sincospif(2.0f * (float)(turns - rint(turns)), &sf, &cf);
        s = (real)sf; co = (real)cf;
      } else {
        double sd, cd;
        sincos(gamma * c, &sd, &cd);
        s = (real)sd; co = (real)cd;
      }
      real r2 = re * co - im * s;
      im = re * s + im * co;
      re = r2;
    }
    if (flags & EW_DOT) {
      dv[0] += (double)sr * re + (double)si * im;
      dv[1] += (double)sr * im - (double)si * re;
    }
    if (flags & EW_REDUCE) {
      double pr = (double)re * re + (double)im * im;
      rv[0] += pr;
      rv[1] += pr * c;
      rv[2] += pr * c * c;
      if (pr > 0.0) { rv[3] = fmin(rv[3], c); rv[4] = fmax(rv[4], c); }
    }
    if (flags & EW_STORE) { st[2 * x] = re; st[2 * x + 1] = im; }
    }
  }
  if (flags & EW_DOT) {
    block_reduce5(dv, red);
    if (threadIdx.x == 0) { part_dot[blockIdx.x * 2] = dv[0]; part_dot[blockIdx.x * 2 + 1] = dv[1]; }
  }
  if (flags & EW_REDUCE) {
    block_reduce5(rv, red);
    if (threadIdx.x == 0)
      for (int i = 0; i < 5; i++) part_red[blockIdx.x * 5 + i] = rv[i];
  }
}

// ---- complex64 Grover passes, specialised (flags, diagonal kind, |s> generator as template parameters) ----------------
// The generic kernel above spends its issue slots on double-precision range reduction, per-element kind switches and
// 8-byte accesses (ncu: 24-36 % of the HBM rate at n = 30).  Here: 16-byte accesses (two amplitudes per thread and
// access, four accesses in flight), the phase angle reduced in 64-bit FIXED POINT (turns = frac(B + k * A) from the raw
// diagonal integer k: two integer multiply-adds, |error| <= 2^-33 turns) followed by the single-precision sincospif, the
// dot product <s|v> accumulated as (-i)^popc(x) v in floats per batch of eight elements and in doubles across batches,
// |s> = amp * i^popc(x) [popc == k] applied by sign / swap instead of a complex multiply.
template <int GK>
__device__ __forceinline__ bool gk_support(const StateGen& g, int pc) {
  return GK == INIT_DICKE ? pc == g.k : true;
}

// FLAGS: compile-time EwFlags; DK: DIAG_U8 / DIAG_U32FX; GK: INIT_PLUS / INIT_DICKE / INIT_UNIFORM (|s> AND the initial
// state: the Grover mixers of the reference are built over their own initial state, grover_mixer.py:43-82).
template <int FLAGS, int DK, int GK>
__global__ void __launch_bounds__(256)
grover_pass_c64_kernel(float2* __restrict__ st, size_t N, StateGen gen, const double* __restrict__ coefdot, double gamma,
                       const void* __restrict__ diag, DiagInfo info, double* __restrict__ part_dot,
                       double* __restrict__ part_red) {
  __shared__ double red[8 * 5];
  const float amp = (float)gen.amp;
  float far = 0.f, fai = 0.f;                        // coefdot * amp
  if (FLAGS & EW_AXPY) { far = (float)(coefdot[0] * gen.amp); fai = (float)(coefdot[1] * gen.amp); }
  const PhaseFx fx = make_phase_fx(gamma, info);
  double dv[5] = {0, 0, 0, 1e300, -1e300};
  double rv[5] = {0, 0, 0, 1e300, -1e300};
  constexpr int U = 4;
  const size_t npairs = N >> 1;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  float4* st4 = reinterpret_cast<float4*>(st);

  auto one = [&](size_t x, float& re, float& im, unsigned k, float& dr, float& di) {
    const int pc = GK == INIT_UNIFORM ? 0 : __popcll((unsigned long long)x) + gen.pc0;
    const bool on = gk_support<GK>(gen, pc);
    if (FLAGS & EW_INIT) {
      re = on ? amp : 0.f;
      im = 0.f;
      if (GK != INIT_UNIFORM) mul_i_pow(re, im, pc);
    }
    if ((FLAGS & EW_AXPY) && on) {
      float ar = far, ai = fai;
      if (GK != INIT_UNIFORM) mul_i_pow(ar, ai, pc);
      re += ar;
      im += ai;
    }
    if (FLAGS & EW_PHASE) {
      float s, c;
      phase_fx_sincos(fx, k, s, c);
      const float r2 = re * c - im * s;
      im = re * s + im * c;
      re = r2;
    }
    if ((FLAGS & EW_DOT) && on) {                    // conj(i^pc) v = (-i)^pc v
      float tr = re, ti = im;
      if (GK != INIT_UNIFORM) mul_i_pow(tr, ti, -pc);
      dr += tr;
      di += ti;
    }
    if (FLAGS & EW_REDUCE) {
      const double c = info.c0 + info.delta * (double)k;
      const double pr = (double)re * re + (double)im * im;
      rv[0] += pr;
      rv[1] += pr * c;
      rv[2] += pr * c * c;
      if (pr > 0.0) { rv[3] = fmin(rv[3], c); rv[4] = fmax(rv[4], c); }
    }
  };
  auto diag_pair = [&](size_t i, unsigned& k0, unsigned& k1) {
    if (DK == DIAG_U8) {
      const unsigned short w = __ldg(reinterpret_cast<const unsigned short*>(diag) + i);
      k0 = w & 0xffu; k1 = w >> 8;
    } else {
      const uint2 w = __ldg(reinterpret_cast<const uint2*>(diag) + i);
      k0 = w.x; k1 = w.y;
    }
  };

  for (size_t i0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i0 < npairs; i0 += U * stride) {
    float4 v[U];
    unsigned k0[U], k1[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const size_t i = i0 + u * stride;
      if (i < npairs) {
        if (!(FLAGS & EW_INIT)) v[u] = st4[i];
        if (FLAGS & (EW_PHASE | EW_REDUCE)) diag_pair(i, k0[u], k1[u]);
      }
    }
    float dr = 0.f, di = 0.f;
#pragma unroll
    for (int u = 0; u < U; u++) {
      const size_t i = i0 + u * stride;
      if (i >= npairs) break;
      one(2 * i, v[u].x, v[u].y, k0[u], dr, di);
      one(2 * i + 1, v[u].z, v[u].w, k1[u], dr, di);
      if (FLAGS & EW_STORE) st4[i] = v[u];
    }
    if (FLAGS & EW_DOT) { dv[0] += (double)dr; dv[1] += (double)di; }
  }
  if ((N & 1) && blockIdx.x == 0 && threadIdx.x == 0) {   // odd compact registers: the last entry
    const size_t x = N - 1;
    float re = 0.f, im = 0.f, dr = 0.f, di = 0.f;
    unsigned k = 0;
    if (!(FLAGS & EW_INIT)) { re = st[x].x; im = st[x].y; }
    if (FLAGS & (EW_PHASE | EW_REDUCE))
      k = DK == DIAG_U8 ? (unsigned)reinterpret_cast<const unsigned char*>(diag)[x] : reinterpret_cast<const unsigned*>(diag)[x];
    one(x, re, im, k, dr, di);
    if (FLAGS & EW_STORE) st[x] = make_float2(re, im);
    if (FLAGS & EW_DOT) { dv[0] += (double)dr; dv[1] += (double)di; }
  }
  if (FLAGS & EW_DOT) {
    dv[0] *= gen.amp;
    dv[1] *= gen.amp;
    block_reduce5(dv, red);
    if (threadIdx.x == 0) { part_dot[blockIdx.x * 2] = dv[0]; part_dot[blockIdx.x * 2 + 1] = dv[1]; }
  }
  if (FLAGS & EW_REDUCE) {
    if (FLAGS & EW_DOT) __syncthreads();
    block_reduce5(rv, red);
    if (threadIdx.x == 0)
      for (int i = 0; i < 5; i++) part_red[blockIdx.x * 5 + i] = rv[i];
  }
}

// coefdot = (e^{-i beta} - 1) * sum(partials)     (single block, fixed order)
__global__ void grover_dot_finalize_kernel(const double* __restrict__ part, int nparts, double beta, double* coefdot) {
  __shared__ double red[32 * 5];
  double v[5] = {0, 0, 0, 1e300, -1e300};
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) { v[0] += part[2 * i]; v[1] += part[2 * i + 1]; }
  block_reduce5(v, red);
  if (threadIdx.x == 0) {
    double sb, cb;
    sincos(beta, &sb, &cb);
    const double kr = cb - 1.0, ki = -sb;
    coefdot[0] = kr * v[0] - ki * v[1];
    coefdot[1] = kr * v[1] + ki * v[0];
  }
}

// raw-sum variant of finalize for the element-wise path (cost values are already decoded)
__global__ void finalize_raw_kernel(const double* __restrict__ partials, int nparts, double* out) {
  __shared__ double red[32 * 5];
  double v[5] = {0, 0, 0, 1e300, -1e300};
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) {
    v[0] += partials[i * 5 + 0];
    v[1] += partials[i * 5 + 1];
    v[2] += partials[i * 5 + 2];
    v[3] = fmin(v[3], partials[i * 5 + 3]);
    v[4] = fmax(v[4], partials[i * 5 + 4]);
  }
  block_reduce5(v, red);
  if (threadIdx.x == 0) {
    out[0] = v[1] / v[0];
    out[1] = v[2] / v[0];
    out[2] = v[3];
    out[3] = v[4];
    out[4] = v[0];
  }
}

// one XY gate on qubits (q0, q1) over the whole register: rotation by `ang` on span{|01>,|10>}
template <typename real>
__global__ void __launch_bounds__(256)
xy_gate_kernel(real* __restrict__ st, size_t N, int q0, int q1, double cd, double sd) {
  const real c = (real)cd, s = (real)sd;
  const int lo = q0 < q1 ? q0 : q1, hi = q0 < q1 ? q1 : q0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < N / 4; i += (size_t)gridDim.x * blockDim.x) {
    size_t t = ((i >> lo) << (lo + 1)) | (i & (((size_t)1 << lo) - 1));
    size_t base = ((t >> hi) << (hi + 1)) | (t & (((size_t)1 << hi) - 1));
    size_t xa = base | ((size_t)1 << q0), xb = base | ((size_t)1 << q1);
    real ar = st[2 * xa], ai = st[2 * xa + 1], br = st[2 * xb], bi = st[2 * xb + 1];
    st[2 * xa] = c * ar + s * bi;
    st[2 * xa + 1] = c * ai - s * br;
    st[2 * xb] = c * br + s * ai;
    st[2 * xb + 1] = c * bi - s * ar;
  }
}

// ============================================================================================
// XY mixer at large n: shared-memory tiles, many gates per sweep
// ============================================================================================
// One CTA owns a tile of 2^tb amplitudes (64 KB: tb = 13 complex64 / 12 complex128; three CTAs per SM)
// gathered from the c lowest amplitude bits (256-byte rows) plus tb - c arbitrary higher bits, applies an
// ORDERED run of XY gates whose qubits all lie inside the tile (reference xy_mixer.py:42-79: the pairs do
// not commute, order is semantics) and writes the tile back -- a chain of 8 gates costs one sweep instead
// of eight.  The phase separator is fused into the first pass of a layer, |D^n_k> / |+> generation into
// the very first pass, the statistics into the very last one.
constexpr int XY_MAX_GATES = 64;
constexpr int XY_MAX_TILE_BITS = 13;
struct XYTilePass {
  int tb;                             // tile amplitude bits
  int pos[XY_MAX_TILE_BITS];          // amplitude-bit position of each tile-local bit (ascending)
  int nfree;
  int freepos[QAOA_MAX_QUBITS];       // positions of the other bits (ascending)
  int ngates;
  unsigned char gi[XY_MAX_GATES], gj[XY_MAX_GATES];   // tile-local bits of (q0, q1), application order
  int flags;                          // EW_INIT / EW_PHASE / EW_REDUCE / EW_STORE
  double gamma, cd, sd;               // phase angle; cos / sin of the gate angle beta / 4 / trotter
};

// shared-memory index swizzle of the XY tiles: bits 4 and 5 are folded onto the low four bits so that the
// pair accesses of a chain gate on tile bits (k, k+1) -- consecutive lanes skip those two bits -- spread
// over all banks for every k (and for the ring closure (tb-1, 0)); plain consecutive accesses stay conflict free
__device__ __forceinline__ size_t xy_swz(size_t e) { return e ^ (((e >> 4) & 1) * 5) ^ (((e >> 5) & 1) * 10); }

// offset contributed by the upper tile bits of element tid + 512 k (k is a compile-time constant at every
// use: the sum lives in uniform registers)
template <int TB>
__device__ __forceinline__ size_t xy_off_k(const XYTilePass& P, int k) {
  size_t o = 0;
#pragma unroll
  for (int b = 9; b < TB; b++) o += (size_t)((k >> (b - 9)) & 1) << P.pos[b];
  return o;
}

template <typename real, int TB>
__global__ void __launch_bounds__(512, 2)
xy_tile_kernel(const __grid_constant__ XYTilePass P, real* __restrict__ st, size_t ntiles, StateGen init,
               const void* __restrict__ diag, DiagInfo info, double* __restrict__ part_red) {
  extern __shared__ __align__(16) unsigned char xy_smem[];
  typedef typename Real2<real>::type real2;
  real2* ps = reinterpret_cast<real2*>(xy_smem);
  __shared__ double red[16 * 5];
  const int tid = threadIdx.x;
  constexpr size_t T = (size_t)1 << TB;
  constexpr int PER = (int)(T / 512);   // elements per thread: all loads of a tile are issued before the first use
  real2* st2 = reinterpret_cast<real2*>(st);
  const real c = (real)P.cd, s = (real)P.sd;
  // element e = tid + 512 k: the thread's 9 low tile bits are fixed, k supplies the upper ones
  size_t off_t = 0;
#pragma unroll
  for (int b = 0; b < 9; b++) off_t += (size_t)((tid >> b) & 1) << P.pos[b];
  constexpr int CH = PER < 8 ? PER : 8;   // loads in flight per thread
  double rv[5] = {0, 0, 0, 1e300, -1e300};
  for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    size_t base = 0;
    for (int i = 0; i < P.nfree; i++) base |= (size_t)((tile >> i) & 1) << P.freepos[i];
#pragma unroll
    for (int k0 = 0; k0 < PER; k0 += CH) {
    real2 in[CH];
    if (!(P.flags & EW_INIT)) {
#pragma unroll
      for (int k = 0; k < CH; k++) in[k] = st2[base + off_t + xy_off_k<TB>(P, k0 + k)];
    }
#pragma unroll
    for (int kk = 0; kk < CH; kk++) {
      const int k = k0 + kk;
      const size_t e = tid + 512 * (size_t)k;
      const size_t x = base + off_t + xy_off_k<TB>(P, k);
      real re, im;
      if (P.flags & EW_INIT) gen_amp<real>(init, x, re, im);
      else { re = in[kk].x; im = in[kk].y; }
      if (P.flags & EW_PHASE) {
        const double cst = diag_decode(info, diag, x);
        real sn, co;
        if (sizeof(real) == 4) {
          const double turns = P.gamma * 0.15915494309189533576888 * cst;
          float sf, cf;
          sincospif(2.0f * (float)(turns - rint(turns)), &sf, &cf);
          sn = (real)sf; co = (real)cf;
        } else {
          double sd2, cd2;
          sincos(P.gamma * cst, &sd2, &cd2);
          sn = (real)sd2; co = (real)cd2;
        }
        const real r2 = re * co - im * sn;
        im = re * sn + im * co;
        re = r2;
      }
      real2 o; o.x = re; o.y = im;
      ps[xy_swz(e)] = o;
    }
    }
    __syncthreads();
    for (int g = 0; g < P.ngates; g++) {
      const int i = P.gi[g], j = P.gj[g];
      const int lo = i < j ? i : j, hi = i < j ? j : i;
      for (size_t t = tid; t < T / 4; t += 512) {
        size_t u = ((t >> lo) << (lo + 1)) | (t & (((size_t)1 << lo) - 1));
        size_t b0 = ((u >> hi) << (hi + 1)) | (u & (((size_t)1 << hi) - 1));
        const size_t xa = xy_swz(b0 | ((size_t)1 << i)), xb = xy_swz(b0 | ((size_t)1 << j));
        const real2 a = ps[xa], b = ps[xb];
        // a' = c a - i s b ; b' = -i s a + c b   (XXPlusYYGate(beta/2) on span{|01>,|10>})
        real2 na, nb;
        na.x = c * a.x + s * b.y;
        na.y = c * a.y - s * b.x;
        nb.x = c * b.x + s * a.y;
        nb.y = c * b.y - s * a.x;
        ps[xa] = na;
        ps[xb] = nb;
      }
      __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < PER; k++) {
      const size_t e = tid + 512 * (size_t)k;
      const size_t x = base + off_t + xy_off_k<TB>(P, k);
      const real2 o = ps[xy_swz(e)];
      const real re = o.x, im = o.y;
      if (P.flags & EW_REDUCE) {
        const double pr = (double)re * re + (double)im * im;
        const double cst = diag_decode(info, diag, x);
        rv[0] += pr;
        rv[1] += pr * cst;
        rv[2] += pr * cst * cst;
        if (pr > 0.0) { rv[3] = fmin(rv[3], cst); rv[4] = fmax(rv[4], cst); }
      }
      if (P.flags & EW_STORE) st2[x] = o;
    }
    __syncthreads();
  }
  if (P.flags & EW_REDUCE) {
    block_reduce5(rv, red);
    if (tid == 0)
      for (int i = 0; i < 5; i++) part_red[blockIdx.x * 5 + i] = rv[i];
  }
}

// ============================================================================================
// sampling from |psi|^2 and the distribution of cost values (kept state, natural order)
// ============================================================================================
// Replaces backend.run(shots=...) + get_counts of QAOA.hist (reference qaoa/qaoa.py:1134-1172): the
// engine draws basis indices by inverse-CDF search.  Stage 1: probability mass of every chunk of
// 2^chunk_log2 amplitudes; the host turns the caller's uniform numbers into (chunk, residual mass)
// pairs; stage 2: one block per shot walks its chunk.  Deterministic for given uniform numbers.
template <typename real>
__global__ void __launch_bounds__(256)
chunk_prob_kernel(const real* __restrict__ st, size_t N, int chunk_log2, double* __restrict__ sums) {
  __shared__ double red[8 * 5];
  const size_t lo = (size_t)blockIdx.x << chunk_log2;
  const size_t hi = lo + ((size_t)1 << chunk_log2) < N ? lo + ((size_t)1 << chunk_log2) : N;
  double v[5] = {0, 0, 0, 1e300, -1e300};
  for (size_t x = lo + threadIdx.x; x < hi; x += blockDim.x) {
    const double re = st[2 * x], im = st[2 * x + 1];
    v[0] += re * re + im * im;
  }
  block_reduce5(v, red);
  if (threadIdx.x == 0) sums[blockIdx.x] = v[0];
}

template <typename real>
__global__ void __launch_bounds__(256)
sample_locate_kernel(const real* __restrict__ st, size_t N, int chunk_log2, const unsigned* __restrict__ chunk_of,
                     const double* __restrict__ resid, unsigned long long* __restrict__ out) {
  __shared__ double ssum[256];
  __shared__ int found;
  const int shot = blockIdx.x, tid = threadIdx.x;
  const size_t lo = (size_t)chunk_of[shot] << chunk_log2;
  const size_t hi = lo + ((size_t)1 << chunk_log2) < N ? lo + ((size_t)1 << chunk_log2) : N;
  const size_t len = hi - lo, per = (len + 255) / 256;
  const size_t a = lo + per * tid < hi ? lo + per * tid : hi;
  const size_t b = a + per < hi ? a + per : hi;
  double s = 0.0;
  for (size_t x = a; x < b; x++) {
    const double re = st[2 * x], im = st[2 * x + 1];
    s += re * re + im * im;
  }
  ssum[tid] = s;
  if (tid == 0) found = 0;
  __syncthreads();
  double prefix = 0.0;
  for (int i = 0; i < tid; i++) prefix += ssum[i];
  const double r = resid[shot];
  if (s > 0.0 && prefix <= r && r < prefix + s) {
    double run = prefix;
    size_t pick = b - 1;
    for (size_t x = a; x < b; x++) {
      const double re = st[2 * x], im = st[2 * x + 1];
      run += re * re + im * im;
      if (run > r) { pick = x; break; }
    }
    out[shot] = (unsigned long long)pick;
    found = 1;
  }
  __syncthreads();
  if (!found && tid == 0) {   // rounding pushed r to the chunk's end: the last amplitude with mass
    size_t pick = lo;
    for (size_t x = hi; x-- > lo;) {
      const double re = st[2 * x], im = st[2 * x + 1];
      if (re * re + im * im > 0.0) { pick = x; break; }
    }
    out[shot] = (unsigned long long)pick;
  }
}

// P(cost = c0 + delta k), k < 256, of the kept state (U8 diagonals): partial[block][256].
// Gives the exact CVaR (reference statistic.py:132-142 in the shots -> infinity limit) and Var.
template <typename real>
__global__ void __launch_bounds__(256)
cost_hist_kernel(const real* __restrict__ st, const uint8_t* __restrict__ diag, size_t N, double* __restrict__ partial) {
  __shared__ double bins[256];
  bins[threadIdx.x] = 0.0;
  __syncthreads();
  for (size_t x = blockIdx.x * (size_t)blockDim.x + threadIdx.x; x < N; x += (size_t)gridDim.x * blockDim.x) {
    const double re = st[2 * x], im = st[2 * x + 1];
    const double p = re * re + im * im;
    if (p > 0.0) atomicAdd(&bins[diag[x]], p);
  }
  __syncthreads();
  partial[(size_t)blockIdx.x * 256 + threadIdx.x] = bins[threadIdx.x];
}

// ---- exact CVaR of ANY cost representation: weighted quantile select on the device ----------------------------------
// CVaR_alpha = mean of the cost over the best alpha of the probability mass (statistic.py:132-142 in the shots ->
// infinity limit: the reference sorts the sampled costs and averages the top round(alpha * shots)).  The threshold
// cost is found digit by digit over an order-preserving integer key of the cost (U8: the byte, U32FX: the 32-bit
// integer, F64: the sign-flipped bit pattern), most significant digit first: one sweep per 8-bit digit histograms the
// probability mass and the mass-weighted cost of the amplitudes whose key matches the prefix found so far.
__device__ __forceinline__ unsigned long long cost_key(const DiagInfo& info, const void* diag, size_t x, double& c) {
  if (info.kind == DIAG_U8) { const unsigned k = ((const uint8_t*)diag)[x]; c = info.c0 + info.delta * (double)k; return k; }
  if (info.kind == DIAG_U32FX) { const unsigned k = ((const uint32_t*)diag)[x]; c = info.c0 + info.delta * (double)k; return k; }
  c = ((const double*)diag)[x];
  const unsigned long long b = (unsigned long long)__double_as_longlong(c);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

// partial: [gridDim.x][512] = {mass[256], mass x cost[256]} of the digit at `shift` among keys with (key >> (shift+8)) == prefix
template <typename real>
__global__ void __launch_bounds__(256)
cvar_digit_kernel(const real* __restrict__ st, const void* __restrict__ diag, DiagInfo info, size_t N,
                  unsigned long long prefix, int shift, int check_prefix, double* __restrict__ partial) {
  __shared__ double bins[512];
  bins[threadIdx.x] = 0.0;
  bins[256 + threadIdx.x] = 0.0;
  __syncthreads();
  for (size_t x = blockIdx.x * (size_t)blockDim.x + threadIdx.x; x < N; x += (size_t)gridDim.x * blockDim.x) {
    const double re = st[2 * x], im = st[2 * x + 1];
    const double p = re * re + im * im;
    if (!(p > 0.0)) continue;
    double c;
    const unsigned long long key = cost_key(info, diag, x, c);
    if (check_prefix && ((key >> shift) >> 8) != prefix) continue;
    const unsigned d = (unsigned)(key >> shift) & 255u;
    atomicAdd(&bins[d], p);
    atomicAdd(&bins[256 + d], p * c);
  }
  __syncthreads();
  partial[(size_t)blockIdx.x * 512 + threadIdx.x] = bins[threadIdx.x];
  partial[(size_t)blockIdx.x * 512 + 256 + threadIdx.x] = bins[256 + threadIdx.x];
}

// out[j] = sum over blocks of partial[b][j]   (fixed order)
__global__ void sum_partials_kernel(const double* __restrict__ partial, int nblocks, int width, double* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= width) return;
  double s = 0.0;
  for (int b = 0; b < nblocks; b++) s += partial[(size_t)b * width + j];
  out[j] = s;
}

// Basis indices of the support of the kept state whose cost equals `target` (the extreme the reduction
// reported): the shots -> infinity limit of Statistic.get_max_sols / get_min_sols (statistic.py:69-82).
template <typename real>
__global__ void __launch_bounds__(256)
collect_cost_equal_kernel(const real* __restrict__ st, const void* __restrict__ diag, DiagInfo info, size_t N,
                          double target, unsigned cap, unsigned long long* __restrict__ out, unsigned* __restrict__ count) {
  for (size_t x = blockIdx.x * (size_t)blockDim.x + threadIdx.x; x < N; x += (size_t)gridDim.x * blockDim.x) {
    if (diag_decode(info, diag, x) != target) continue;
    const double re = st[2 * x], im = st[2 * x + 1];
    if (re * re + im * im > 0.0) {
      const unsigned i = atomicAdd(count, 1u);
      if (i < cap) out[i] = (unsigned long long)x;
    }
  }
}

// ============================================================================================
// feasible-subspace compression: Dicke initial state + Grover mixer over the same Dicke state
// ============================================================================================
// |D^n_k> and the Grover reflection about it never leave the span of the C(n,k) weight-k basis states
// (dicke_initialstate.py:38-56, grover_mixer.py:43-82), so the register is stored compactly: entry i is the
// i-th weight-k bit string in the combinatorial number system (rank = sum_j C(b_j, j+1) over its set bits
// b_0 < b_1 < ...).  This kernel gathers the cost of every feasible string into the compact diagonal
// (raw representation, same DiagInfo).  binom[b * (k+1) + j] = C(b, j).
template <int VALBYTES>
__global__ void gather_feasible_kernel(const unsigned char* __restrict__ diag, unsigned char* __restrict__ out, size_t M,
                                       int n, int k, const unsigned long long* __restrict__ binom) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < M; i += (size_t)gridDim.x * blockDim.x) {
    unsigned long long r = i, x = 0;
    int j = k;
    for (int b = n - 1; b >= 0 && j > 0; b--) {
      const unsigned long long cbj = binom[b * (k + 1) + j];
      if (r >= cbj) { x |= 1ull << b; r -= cbj; j--; }
    }
#pragma unroll
    for (int q = 0; q < VALBYTES; q++) out[i * VALBYTES + q] = diag[x * VALBYTES + q];
  }
}

// ---- XY mixer in the feasible subspace ------------------------------------------------------------
// XY gates conserve the Hamming weight (xy_mixer.py:42-79 on |D^n_k>, dicke_initialstate.py:38-56), so with a Dicke
// initial state the register never leaves the C(n,k) weight-k strings: it is stored compactly, entry i = the i-th
// weight-k string in colex order (rank(x) = sum over set bits, the j-th from below at position b, of C(b, j) -- the
// order gather_feasible_kernel uses).  `strings[i]` holds the string itself.
template <typename S>
__global__ void unrank_feasible_kernel(S* __restrict__ strings, size_t M, int n, int k,
                                       const unsigned long long* __restrict__ binom) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < M; i += (size_t)gridDim.x * blockDim.x) {
    unsigned long long r = i, x = 0;
    int j = k;
    for (int b = n - 1; b >= 0 && j > 0; b--) {
      const unsigned long long cbj = binom[b * (k + 1) + j];
      if (r >= cbj) { x |= 1ull << b; r -= cbj; j--; }
    }
    strings[i] = (S)x;
  }
}

// One XY gate on (q0, q1) over the compact register: the thread of a string with x[q0] = 1, x[q1] = 0 finds the rank of
// its partner (bits q0, q1 swapped) -- only the set bits between the two positions change their contribution to the
// rank -- and rotates the pair:  a' = c a - i s b,  b' = c b - i s a  (the S frame leaves XY gates unchanged).
template <typename real, typename S>
__global__ void __launch_bounds__(256)
xy_compact_gate_kernel(typename Real2<real>::type* __restrict__ st, const S* __restrict__ strings, size_t M, int q0, int q1,
                       real c, real s, const unsigned long long* __restrict__ binom, int n, int k) {
  extern __shared__ unsigned long long sb[];
  const int k1 = k + 1;
  for (int t = threadIdx.x; t < (n + 1) * k1; t += blockDim.x) sb[t] = binom[t];
  __syncthreads();
  const int lo = q0 < q1 ? q0 : q1, hi = q0 < q1 ? q1 : q0;
  const unsigned long long range = ((hi == 63 ? 0ull : (1ull << (hi + 1))) - 1ull) & ~((1ull << lo) - 1ull);
  typedef typename Real2<real>::type real2;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < M; i += (size_t)gridDim.x * blockDim.x) {
    const unsigned long long x = (unsigned long long)strings[i];
    if (!((x >> q0) & 1ull) || ((x >> q1) & 1ull)) continue;
    const unsigned long long xp = x ^ (1ull << q0) ^ (1ull << q1);
    const int j0 = __popcll(x & ((1ull << lo) - 1ull));
    long long delta = 0;
    unsigned long long m = x & range;
    int j = j0;
    while (m) { const int b = __ffsll((long long)m) - 1; m &= m - 1; j++; delta -= (long long)sb[b * k1 + j]; }
    m = xp & range;
    j = j0;
    while (m) { const int b = __ffsll((long long)m) - 1; m &= m - 1; j++; delta += (long long)sb[b * k1 + j]; }
    const size_t ip = (size_t)((long long)i + delta);
    const real2 a = st[i], b2 = st[ip];
    real2 an, bn;
    an.x = c * a.x + s * b2.y;
    an.y = c * a.y - s * b2.x;
    bn.x = c * b2.x + s * a.y;
    bn.y = c * b2.y - s * a.x;
    st[i] = an;
    st[ip] = bn;
  }
}

// ---- per-term gammas and parameterised |+> beyond one CTA -----------------------------------------
// MaxCutFree / MaxCutOrbit give every edge (orbit) its own gamma (maxcut_free_problem.py:31-99,
// maxcut_orbit_problem.py:51-181): the phase separator of layer d is exp(i sum_e g_e [colour(u_e) != colour(v_e)]) with
// g_e = gamma_{d, term(e)} * w_e -- no single cost diagonal times a scalar, so no phase table.  Large registers run such
// circuits LAYER BY LAYER (engine.cu run_layered_point): this sweep turns the state kept by the previous layer's mixer
// passes (true amplitude = src * scale) -- or, for the first layer, the generated initial state with PlusParameterized's
// RZ phases (plus_parameterized_initialstate.py:44-64) -- into the next segment's initial state, multiplied by the
// layer's edge phases evaluated on the fly from the edge list; the mixer passes then run with gamma = 0.
template <typename real>
__global__ void __launch_bounds__(256)
layer_phase_kernel(const real* __restrict__ src, real* __restrict__ dst, size_t N, StateGen init, real scale,
                   MaxKCutDev D /* ew = g_e of this layer; n_edges = 0: no edge phase */, int n_init,
                   const double* __restrict__ init_phi) {
  for (size_t x = blockIdx.x * (size_t)blockDim.x + threadIdx.x; x < N; x += (size_t)gridDim.x * blockDim.x) {
    real re, im;
    double theta = 0.0;
    if (src) { re = src[2 * x] * scale; im = src[2 * x + 1] * scale; }
    else {
      gen_amp<real>(init, x, re, im);
      for (int q = 0; q < n_init; q++) theta += __ldg(init_phi + q) * (((x >> q) & 1) ? 0.5 : -0.5);
    }
    for (int e = 0; e < D.n_edges; e++) {
      const int cu = node_colour(D, x, __ldg(D.eu + e)), cv = node_colour(D, x, __ldg(D.ev + e));
      theta += (cu != cv) ? __ldg(D.ew + e) : 0.0;
    }
    double turns = theta * 0.15915494309189533576888;
    turns -= rint(turns);
    real s, c;
    if (sizeof(real) == 4) { float sf, cf; sincospif(2.0f * (float)turns, &sf, &cf); s = (real)sf; c = (real)cf; }
    else { double sd, cd; sincospi(2.0 * turns, &sd, &cd); s = (real)sd; c = (real)cd; }
    dst[2 * x] = re * c - im * s;
    dst[2 * x + 1] = re * s + im * c;
  }
}
