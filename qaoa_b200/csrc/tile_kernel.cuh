// Register-tiled QAOA pass kernel for sm_100a.
//
// One CTA (512 threads) owns a tile of 2^14 state ELEMENTS (8 bytes each: a complex64 amplitude,
// or one double of a complex128 amplitude) held entirely in registers, 32 per thread.  A tile is
// the set of elements whose index differs only in 14 chosen "tile bits" pos[0..13] (pos[0]=0,
// pos[0..3] are the 4 lowest element bits so that every global access is a full 128-byte line).
// The kernel interprets a short per-pass program (TilePass::ops) whose steps are
//     LOAD / INIT   : fill the registers (coalesced 128-bit loads, or generate |+> / Dicke in place)
//     PHASE         : psi(x) *= sigma * exp(i gamma cost(x))      (phase separator, fused)
//     ROUND         : X-mixer butterflies on the 5 tile bits that are register-resident right now
//     WT / BT       : warp-local / block-wide transposes through shared memory that rotate which
//                     tile bits are register-resident (frames A -> A' -> B -> B' -> A ...)
//     STORE / REDUCE: write back, or fuse the <cost>, <cost^2>, min, max reduction and skip the write
//
// Butterflies use the unnormalised "lifted" rotation: in the S frame the mixer on one qubit is
// c*[[1,t],[-t,1]] (t = tan beta), i.e. ONE fused multiply-add per amplitude (FFMA2 on the packed
// (re,im) pair for complex64, DFMA for complex128).  The per-layer scale prod(c) is folded into the
// next layer's phase multiply (sigma) and finally into the reduction scale, so it costs nothing.
// For |tan beta| > 1 the host switches to the swapped form (rotation by beta -+ pi/2 composed with
// a swap) and tracks the resulting sign flips (pending Z) itself; the kernel only sees two signed
// coefficients and a swap flag per register slot.
#pragma once
#include "common.cuh"

constexpr int TILE_BITS = 14;
constexpr int TILE_ELEMS = 1 << TILE_BITS;
constexpr int TILE_THREADS = 512;
constexpr int TILE_MAX_OPS = 64;
constexpr int TILE_MAX_PHASES = 16;
constexpr int TILE_ROUND_DOUBLES = 11;  // 5 x (tau1, tau2) + flags
constexpr int TILE_EXCH_BYTES = TILE_ELEMS * 8;

enum TileOpType : int {
  TOP_LOAD = 1, TOP_STORE = 2, TOP_INIT = 3, TOP_PHASE = 4,
  TOP_ROUND = 5, TOP_WT = 6, TOP_BT = 7, TOP_REDUCE = 8
};

struct TileOp {
  int16_t type;
  int16_t frame;  // 0 = A, 1 = B   (LOAD/STORE/INIT/PHASE/REDUCE)
  int32_t arg;    // ROUND: round index; PHASE/REDUCE: phase slot index
};

struct TilePass {
  int nops;
  TileOp ops[TILE_MAX_OPS];
  int pos[TILE_BITS];  // element-bit position of each tile-local bit (ascending, pos[0] == 0)
  int nfree;
  int freepos[QAOA_MAX_QUBITS + 1];  // element-bit positions of the non-tile bits (ascending)
  int qshift;                        // 0: complex64, 1: complex128 (element bit 0 = re/im)
  int init_kind;
  int dicke_k;
  double init_amp;
  int nrounds;
  int nphases;
  int params_stride;                    // doubles per point
  const void* streams[TILE_MAX_PHASES]; // permuted cost-diagonal stream per phase slot
  int table_id[TILE_MAX_PHASES];        // index of the 256-entry phase table (U8 mode)
};

template <typename Elem> struct ElemTraits;
template <> struct ElemTraits<float2> {
  typedef float real;
  typedef float2 cplx;
  static constexpr int NAMP = 32;
  static constexpr bool IS_C64 = true;
};
template <> struct ElemTraits<double> {
  typedef double real;
  typedef double2 cplx;
  static constexpr int NAMP = 16;
  static constexpr bool IS_C64 = false;
};

// ---- element arithmetic -------------------------------------------------------------------
__device__ __forceinline__ float2 efma(float t, float2 b, float2 a) {
  return __ffma2_rn(make_float2(t, t), b, a);
}
__device__ __forceinline__ double efma(double t, double b, double a) { return fma(t, b, a); }

__device__ __forceinline__ void load16(const float2* p, float2& a, float2& b) {
  float4 t = *reinterpret_cast<const float4*>(p);
  a = make_float2(t.x, t.y);
  b = make_float2(t.z, t.w);
}
__device__ __forceinline__ void load16(const double* p, double& a, double& b) {
  double2 t = *reinterpret_cast<const double2*>(p);
  a = t.x;
  b = t.y;
}
__device__ __forceinline__ void store16(float2* p, float2 a, float2 b) {
  *reinterpret_cast<float4*>(p) = make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void store16(double* p, double a, double b) {
  *reinterpret_cast<double2*>(p) = make_double2(a, b);
}

// ---- addressing ---------------------------------------------------------------------------
struct TileAddr {
  size_t e0;      // element index of (this thread, register 0) in the given frame
  size_t roff[4]; // element offsets of the 4 upper register bits
};

__device__ __forceinline__ size_t tile_base(const TilePass& P, unsigned tile) {
  size_t base = 0;
  for (int i = 0; i < P.nfree; i++) base |= (size_t)((tile >> i) & 1u) << P.freepos[i];
  return base;
}

__device__ __forceinline__ TileAddr tile_addr(const TilePass& P, size_t base, int frame, int warp, int lane) {
  TileAddr a;
  size_t e = base;
#pragma unroll
  for (int j = 0; j < 5; j++) e += (size_t)((lane >> j) & 1) << P.pos[1 + j];
  const int wofs = frame == 0 ? 6 : 10, rofs = frame == 0 ? 10 : 6;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    e += (size_t)((warp >> j) & 1) << P.pos[wofs + j];
    a.roff[j] = (size_t)1 << P.pos[rofs + j];
  }
  a.e0 = e;
  return a;
}

__device__ __forceinline__ size_t reg_off(const TileAddr& a, int q) {
  return ((q & 1) ? a.roff[0] : 0) + ((q & 2) ? a.roff[1] : 0) + ((q & 4) ? a.roff[2] : 0) +
         ((q & 8) ? a.roff[3] : 0);
}

// position (in 16-byte units) of chunk j of this thread's values inside a diagonal stream
__device__ __forceinline__ size_t stream_chunk(unsigned tile, int warp, int lane, int nchunks, int j) {
  return (((size_t)tile * 16 + warp) * nchunks + j) * 32 + lane;
}

// ---- transposes ---------------------------------------------------------------------------
// warp-local: (lane l, reg r) -> (lane r, reg l); XOR-swizzled 16-byte chunks, conflict free.
template <typename Elem>
__device__ __forceinline__ void warp_transpose(Elem (&v)[32], Elem* exch, int warp, int lane) {
  Elem* base = exch + warp * 1024;
  __syncwarp();
#pragma unroll
  for (int q = 0; q < 16; q++) {
    int chunk = lane * 16 + (q ^ (lane & 15));
    store16(base + chunk * 2, v[2 * q], v[2 * q + 1]);
  }
  __syncwarp();
#pragma unroll
  for (int r = 0; r < 32; r++) {
    int chunk = r * 16 + ((lane >> 1) ^ (r & 15));
    v[r] = base[chunk * 2 + (lane & 1)];
  }
}

// block-wide: (warp w, lane l, reg r) -> (warp l>>1, lane r, reg (l&1)|(w<<1)).
template <typename Elem>
__device__ __forceinline__ void block_transpose(Elem (&v)[32], Elem* exch, int warp, int lane) {
  __syncthreads();
  {
    const int hi = lane >> 1, half = lane & 1;
    const int ubase = 32 * warp + 512 * hi;
#pragma unroll
    for (int r = 0; r < 32; r++) {
      int u = (ubase + r) ^ (hi & 7);
      exch[2 * u + half] = v[r];
    }
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 16; q++) {
    int u = (lane + 32 * q + 512 * warp) ^ (warp & 7);
    load16(exch + 2 * u, v[2 * q], v[2 * q + 1]);
  }
}

// ---- butterflies --------------------------------------------------------------------------
template <typename Elem>
__device__ __forceinline__ void do_round(Elem (&v)[32], const double* __restrict__ rp) {
  typedef typename ElemTraits<Elem>::real real;
  const unsigned long long fl = (unsigned long long)__double_as_longlong(__ldg(rp + 10));
#pragma unroll
  for (int s = 0; s < 5; s++) {
    if ((fl >> s) & 1ull) {
      const real t1 = (real)__ldg(rp + 2 * s), t2 = (real)__ldg(rp + 2 * s + 1);
      if ((fl >> (8 + s)) & 1ull) {
#pragma unroll
        for (int i = 0; i < 32; i++) {
          if (!(i & (1 << s))) {
            Elem a = v[i], b = v[i | (1 << s)];
            v[i] = efma(t1, a, b);
            v[i | (1 << s)] = efma(t2, b, a);
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; i++) {
          if (!(i & (1 << s))) {
            Elem a = v[i], b = v[i | (1 << s)];
            v[i] = efma(t1, b, a);
            v[i | (1 << s)] = efma(t2, a, b);
          }
        }
      }
    }
  }
}

// ---- amplitude access helpers (frames A and B only: register bit 0 is tile bit 0) ----------
__device__ __forceinline__ void amp_get(const float2 (&v)[32], int a, float& re, float& im) {
  re = v[a].x; im = v[a].y;
}
__device__ __forceinline__ void amp_set(float2 (&v)[32], int a, float re, float im) {
  v[a] = make_float2(re, im);
}
__device__ __forceinline__ void amp_get(const double (&v)[32], int a, double& re, double& im) {
  re = v[2 * a]; im = v[2 * a + 1];
}
__device__ __forceinline__ void amp_set(double (&v)[32], int a, double re, double im) {
  v[2 * a] = re; v[2 * a + 1] = im;
}

__device__ __forceinline__ void phase_sincos(double turns, float& s, float& c) {
  double fr = turns - rint(turns);
  sincospif(2.0f * (float)fr, &s, &c);
}
__device__ __forceinline__ void phase_sincos(double turns, double& s, double& c) {
  double fr = turns - rint(turns);
  sincospi(2.0 * fr, &s, &c);
}

// Loads this thread's NAMP diagonal values (as raw integer k for U8/U32FX, as double for F64).
// kd[a] = value such that cost = c0 + delta * kd (F64: c0 = 0, delta = 1).
template <int NAMP, typename F>
__device__ __forceinline__ void for_each_diag(const void* stream, int kind, unsigned tile, int warp,
                                              int lane, F&& f) {
  const uint4* s4 = reinterpret_cast<const uint4*>(stream);
  if (kind == DIAG_U8) {
#pragma unroll
    for (int j = 0; j < NAMP / 16; j++) {
      uint4 w = __ldg(s4 + stream_chunk(tile, warp, lane, NAMP / 16, j));
      unsigned ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int b = 0; b < 16; b++) f(j * 16 + b, (ww[b >> 2] >> (8 * (b & 3))) & 0xffu, 0.0);
    }
  } else if (kind == DIAG_U32FX) {
#pragma unroll
    for (int j = 0; j < NAMP / 4; j++) {
      uint4 w = __ldg(s4 + stream_chunk(tile, warp, lane, NAMP / 4, j));
      f(j * 4 + 0, w.x, 0.0);
      f(j * 4 + 1, w.y, 0.0);
      f(j * 4 + 2, w.z, 0.0);
      f(j * 4 + 3, w.w, 0.0);
    }
  } else {
#pragma unroll
    for (int j = 0; j < NAMP / 2; j++) {
      uint4 w = __ldg(s4 + stream_chunk(tile, warp, lane, NAMP / 2, j));
      f(j * 2 + 0, 0u, __hiloint2double((int)w.y, (int)w.x));
      f(j * 2 + 1, 0u, __hiloint2double((int)w.w, (int)w.z));
    }
  }
}

// ---- the pass kernel ----------------------------------------------------------------------
template <typename Elem>
__global__ void __launch_bounds__(TILE_THREADS, 1)
tile_pass_kernel(const __grid_constant__ TilePass P, Elem* __restrict__ state, size_t point_stride,
                 const double* __restrict__ params, const void* __restrict__ tables,
                 int tables_per_point, DiagInfo dinfo, double* __restrict__ partials) {
  typedef typename ElemTraits<Elem>::real real;
  typedef typename ElemTraits<Elem>::cplx cplx;
  constexpr int NAMP = ElemTraits<Elem>::NAMP;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  Elem* exch = reinterpret_cast<Elem*>(smem_raw);
  cplx* tab = reinterpret_cast<cplx*>(smem_raw + TILE_EXCH_BYTES);             // 256 entries
  double* red = reinterpret_cast<double*>(smem_raw + TILE_EXCH_BYTES + 4096);  // 16*5 doubles

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned tile = blockIdx.x;
  const unsigned point = blockIdx.y;
  state += (size_t)point * point_stride;
  const double* pp = params + (size_t)point * P.params_stride;
  const double* phase_pp = pp + P.nrounds * TILE_ROUND_DOUBLES;
  const size_t base = tile_base(P, tile);

  Elem v[32];

  for (int io = 0; io < P.nops; io++) {
    const TileOp op = P.ops[io];
    switch (op.type) {
      case TOP_LOAD: {
        TileAddr ad = tile_addr(P, base, op.frame, warp, lane);
#pragma unroll
        for (int q = 0; q < 16; q++) load16(state + ad.e0 + reg_off(ad, q), v[2 * q], v[2 * q + 1]);
      } break;
      case TOP_STORE: {
        TileAddr ad = tile_addr(P, base, op.frame, warp, lane);
#pragma unroll
        for (int q = 0; q < 16; q++) store16(state + ad.e0 + reg_off(ad, q), v[2 * q], v[2 * q + 1]);
      } break;
      case TOP_INIT: {
        TileAddr ad = tile_addr(P, base, op.frame, warp, lane);
        const real A = (real)P.init_amp;
#pragma unroll
        for (int a = 0; a < NAMP; a++) {
          // amplitude a of this thread: complex64 -> register a; complex128 -> registers 2a, 2a+1
          const int q = ElemTraits<Elem>::IS_C64 ? (a >> 1) : a;
          size_t e = ad.e0 + reg_off(ad, q) + (ElemTraits<Elem>::IS_C64 ? (a & 1) : 0);
          unsigned long long x = (unsigned long long)(e >> P.qshift);
          int pc = __popcll(x);
          real re = A, im = (real)0;
          if (P.init_kind == INIT_DICKE && pc != P.dicke_k) re = (real)0;
          mul_i_pow(re, im, pc);
          amp_set(v, a, re, im);
        }
      } break;
      case TOP_PHASE: {
        const int slot = op.arg;
        const double gamma = __ldg(phase_pp + 2 * slot), sigma = __ldg(phase_pp + 2 * slot + 1);
        if (dinfo.kind == DIAG_U8) {
          __syncthreads();
          if (threadIdx.x < 256) {
            const cplx* t = reinterpret_cast<const cplx*>(tables) +
                            ((size_t)point * tables_per_point + P.table_id[slot]) * 256;
            tab[threadIdx.x] = t[threadIdx.x];
          }
          __syncthreads();
          for_each_diag<NAMP>(P.streams[slot], DIAG_U8, tile, warp, lane,
                              [&](int a, unsigned k, double) {
                                cplx ph = tab[k];
                                real re, im;
                                amp_get(v, a, re, im);
                                amp_set(v, a, re * ph.x - im * ph.y, re * ph.y + im * ph.x);
                              });
        } else {
          const double inv2pi = 0.15915494309189533576888;
          const double g0 = gamma * dinfo.c0 * inv2pi, gd = gamma * dinfo.delta * inv2pi;
          const real sg = (real)sigma;
          for_each_diag<NAMP>(P.streams[slot], dinfo.kind, tile, warp, lane,
                              [&](int a, unsigned k, double cd) {
                                double turns = dinfo.kind == DIAG_U32FX ? fma((double)k, gd, g0)
                                                                        : fma(cd, gd, g0);
                                real s, c;
                                phase_sincos(turns, s, c);
                                s *= sg; c *= sg;
                                real re, im;
                                amp_get(v, a, re, im);
                                amp_set(v, a, re * c - im * s, re * s + im * c);
                              });
        }
      } break;
      case TOP_ROUND:
        do_round<Elem>(v, pp + op.arg * TILE_ROUND_DOUBLES);
        break;
      case TOP_WT:
        warp_transpose<Elem>(v, exch, warp, lane);
        break;
      case TOP_BT:
        block_transpose<Elem>(v, exch, warp, lane);
        break;
      case TOP_REDUCE: {
        const int slot = op.arg;
        const double scale = __ldg(phase_pp + 2 * P.nphases);  // red_scale follows the phase params
        real s0 = 0, s1 = 0, s2 = 0;
        double kmin = 1e300, kmax = -1e300;
        for_each_diag<NAMP>(P.streams[slot], dinfo.kind, tile, warp, lane,
                            [&](int a, unsigned k, double cd) {
                              real re, im;
                              amp_get(v, a, re, im);
                              real p = re * re + im * im;
                              real kv = dinfo.kind == DIAG_F64 ? (real)cd : (real)k;
                              double kdv = dinfo.kind == DIAG_F64 ? cd : (double)k;
                              s0 += p;
                              s1 = fma(p, kv, s1);
                              s2 = fma(p * kv, kv, s2);
                              if (p > (real)0) { kmin = fmin(kmin, kdv); kmax = fmax(kmax, kdv); }
                            });
        double r5[5] = {(double)s0 * scale, (double)s1 * scale, (double)s2 * scale, kmin, kmax};
        block_reduce5(r5, red);
        if (threadIdx.x == 0) {
          double* o = partials + ((size_t)point * gridDim.x + tile) * 5;
#pragma unroll
          for (int i = 0; i < 5; i++) o[i] = r5[i];
        }
      } break;
      default:
        break;
    }
  }
}

// Builds one permuted diagonal stream: stream[(tile, warp, chunk, lane, i)] = diag[x(tile,warp,lane,amp)]
// using exactly the addressing of the pass kernel (same helper functions).
template <int VALBYTES>
__global__ void make_stream_kernel(const __grid_constant__ TilePass P, int frame, int namp,
                                   const unsigned char* __restrict__ diag,
                                   unsigned char* __restrict__ stream) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned tile = blockIdx.x;
  const size_t base = tile_base(P, tile);
  TileAddr ad = tile_addr(P, base, frame, warp, lane);
  constexpr int V = 16 / VALBYTES;
  const int nchunks = namp / V;
  for (int a = 0; a < namp; a++) {
    // amplitude a -> register index: complex64: r = a (q = a>>1, low bit a&1); complex128: q = a
    int q = (namp == 32) ? (a >> 1) : a;
    size_t e = ad.e0 + reg_off(ad, q) + ((namp == 32) ? (a & 1) : 0);
    size_t x = e >> P.qshift;
    size_t dst = (stream_chunk(tile, warp, lane, nchunks, a / V) * V + (a % V)) * VALBYTES;
    const unsigned char* src = diag + x * VALBYTES;
#pragma unroll
    for (int b = 0; b < VALBYTES; b++) stream[dst + b] = src[b];
  }
}
