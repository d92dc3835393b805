// Shared definitions for the B200 QAOA statevector engine.
//
// Conventions:
//  * The device state is ALWAYS kept in the "S frame":  psi~(x) = i^popcount(x) * psi(x).
//    In that frame the X mixer exp(+i beta X) (reference qaoa/mixers/x_mixer.py:35) is the REAL
//    rotation [[c, s], [-s, c]] acting identically on the real and the imaginary part, diagonal
//    operators (the phase separator) are unchanged, and |psi~|^2 = |psi|^2.
//  * complex64 state : element = float2 (re, im), element index e == basis index x.
//    complex128 state: element = double, element index e = 2*x + part (part 0 = re, 1 = im);
//    "element bit 0" is then the part bit and qubit q is element bit q+1.
//  * bit i of x  <->  qubit i  <->  string[i] of the reference's cost(string) (qaoa/qaoa.py:605-609).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define QAOA_MAX_QUBITS 40

enum DiagKind : int {
  DIAG_NONE = 0,
  DIAG_U8 = 1,     // cost(x) = c0 + delta * u8      (exact for small integer costs; phase by 256-entry table)
  DIAG_U32FX = 2,  // cost(x) = c0 + delta * u32     (complex64 runs of real-valued costs)
  DIAG_F64 = 3     // cost(x) = f64                  (complex128 runs of real-valued costs)
};

enum InitKind : int { INIT_PLUS = 0, INIT_DICKE = 1, INIT_STATEVECTOR = 2, INIT_UNIFORM = 4 /* internal */ };
enum MixerKind : int { MIXER_X = 0, MIXER_XMULTI = 1, MIXER_XY = 2, MIXER_GROVER = 3, MIXER_NODE_GROVER = 4 };

// Tensor product of Grover mixers over b-qubit node registers (MaxKCutGrover(tensorized=True), reference
// qaoa/mixers/maxkcut_grover_mixer.py:116-124): the node state |s_node>, 2^b <= 8 amplitudes, in the S frame.
struct NodeVec {
  int bits;
  double re[8], im[8];
};

struct DiagInfo {
  int kind;
  double c0;
  double delta;
};

__host__ __device__ inline int diag_value_bytes(int kind) {
  return kind == DIAG_U8 ? 1 : (kind == DIAG_U32FX ? 4 : 8);
}

// multiply (re, im) by i^k
template <typename T>
__device__ __forceinline__ void mul_i_pow(T& re, T& im, int k) {
  k &= 3;
  T r = re, m = im;
  if (k == 1) { re = -m; im = r; }
  else if (k == 2) { re = -r; im = -m; }
  else if (k == 3) { re = m; im = -r; }
}

__device__ __forceinline__ double diag_decode(const DiagInfo& d, const void* p, size_t x) {
  if (d.kind == DIAG_U8) return d.c0 + d.delta * (double)((const uint8_t*)p)[x];
  if (d.kind == DIAG_U32FX) return d.c0 + d.delta * (double)((const uint32_t*)p)[x];
  return ((const double*)p)[x];
}

// Phase angle of a raw diagonal integer k in 64-bit FIXED POINT: turns = frac(B + k * A), |error| <= 2^-33 turns for
// k < 2^32 -- two integer multiply-adds instead of a double-precision multiply, round and subtract per amplitude.
struct PhaseFx {
  unsigned long long A, B;   // turns per diagonal unit / turns at k = 0, as 0.64 fixed-point fractions
};
__host__ __device__ inline unsigned long long turns_to_fx(double t) {
  t -= floor(t);                                   // [0, 1)
  const double v = t * 18446744073709551616.0;     // 2^64: exact scaling
  return v >= 18446744073709549568.0 ? 0xfffffffffffff800ull : (unsigned long long)v;
}
__host__ __device__ inline PhaseFx make_phase_fx(double gamma, const DiagInfo& d) {
  const double inv2pi = 0.15915494309189533576888;
  PhaseFx f;
  f.A = turns_to_fx(gamma * d.delta * inv2pi);
  f.B = turns_to_fx(gamma * d.c0 * inv2pi);
  return f;
}
__device__ __forceinline__ void phase_fx_sincos(const PhaseFx& f, unsigned k, float& s, float& c) {
  const unsigned long long t = f.B + f.A * (unsigned long long)k;     // frac(turns) * 2^64, wraps mod 1 turn
  const float half_turns = (float)(int)(t >> 32) * 4.656612873077392578125e-10f;   // 2 * signed turns in [-1, 1)
  sincospif(half_turns, &s, &c);
}

// same reduction, hardware sine / cosine (MUFU; absolute error ~2^-21 on [-pi, pi]): the tile passes
__device__ __forceinline__ void phase_fx_sincos_fast(const PhaseFx& f, unsigned k, float& s, float& c) {
  const unsigned long long t = f.B + f.A * (unsigned long long)k;
  __sincosf((float)(int)(t >> 32) * 1.4629180792671596e-9f, &s, &c);   // 2 pi / 2^32
}

// deterministic block reduction of NV doubles: sums for the first NS, then one min, one max
// scratch must hold (blockDim.x/32) * 5 doubles.
__device__ __forceinline__ void block_reduce5(double v[5], double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
    v[1] += __shfl_xor_sync(0xffffffffu, v[1], o);
    v[2] += __shfl_xor_sync(0xffffffffu, v[2], o);
    v[3] = fmin(v[3], __shfl_xor_sync(0xffffffffu, v[3], o));
    v[4] = fmax(v[4], __shfl_xor_sync(0xffffffffu, v[4], o));
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 5; i++) scratch[warp * 5 + i] = v[i];
  }
  __syncthreads();
  if (warp == 0) {
    double w[5] = {0.0, 0.0, 0.0, 1e300, -1e300};
    if (lane < nw) {
#pragma unroll
      for (int i = 0; i < 5; i++) w[i] = scratch[lane * 5 + i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      w[0] += __shfl_xor_sync(0xffffffffu, w[0], o);
      w[1] += __shfl_xor_sync(0xffffffffu, w[1], o);
      w[2] += __shfl_xor_sync(0xffffffffu, w[2], o);
      w[3] = fmin(w[3], __shfl_xor_sync(0xffffffffu, w[3], o));
      w[4] = fmax(w[4], __shfl_xor_sync(0xffffffffu, w[4], o));
    }
#pragma unroll
    for (int i = 0; i < 5; i++) v[i] = w[i];
  }
  __syncthreads();
}
