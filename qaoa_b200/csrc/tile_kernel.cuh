// Register-tiled QAOA pass kernels for sm_100a.
//
// One CTA (512 threads) owns a tile of 2^14 state ELEMENTS (8 bytes each: a complex64 amplitude,
// or one double of a complex128 amplitude) held entirely in registers, 32 per thread.  A tile is
// the set of elements whose index differs only in 14 chosen "tile bits" pos[0..13] (pos[0]=0,
// pos[0..3] are the 4 lowest element bits so that every global access is a full 128-byte line).
// A pass is a short program of steps:
//     LOAD / INIT   : fill the registers (coalesced 128-bit loads, or generate |+> / Dicke in place)
//     PHASE         : psi(x) *= m * exp(i gamma cost(x))      (phase separator, fused; m = scale owed)
//     ROUND         : X-mixer butterflies on the 5 tile bits that are register-resident right now
//     WT / BT       : warp-local / block-wide transposes through shared memory that rotate which
//                     tile bits are register-resident (frames A -> A' -> B -> B' -> A ...)
//     STORE / REDUCE: write back, or fuse the <cost>, <cost^2>, min, max reduction and skip the write
//
// Two kernels execute such programs:
//   * tile_prog_kernel<Elem, Prog>: the program is a COMPILE-TIME constant, the whole pass is
//     straight-line code (free register renaming, no per-step dispatch) -- used for the four
//     canonical passes FIRST / INTERIOR / BOUNDARY / FINAL that make up every large evaluation;
//   * tile_pass_kernel<Elem>: interprets TilePass::ops at run time -- fallback for irregular passes
//     (several layers in one pass at small n, keep_state, mixed butterfly forms).
//
// Butterflies use the unnormalised "lifted" rotation: in the S frame the mixer on one qubit is
// c*[[1,t],[-t,1]] (t = tan beta), i.e. ONE fused multiply-add per amplitude (FFMA2 on the packed
// (re,im) pair for complex64, DFMA for complex128).  The per-layer scale prod(c) is split into a
// pre-shrink folded into the SAME layer's phase multiply and a post factor folded into the NEXT
// phase multiply / the reduction, so it costs nothing and keeps fp32 inside its exponent range.
// Near beta = pi/2 (|t| huge) the host selects the "cot form" s*[[u,1],[-1,u]] (u = cot beta) per
// register slot; passes containing such slots run on the interpreter kernel.
#pragma once
#include "common.cuh"

constexpr int TILE_BITS = 14;
constexpr int TILE_ELEMS = 1 << TILE_BITS;
constexpr int TILE_THREADS = 512;
constexpr int TILE_MAX_OPS = 64;
constexpr int TILE_MAX_PHASES = 16;
constexpr int TILE_ROUND_REALS = 12;  // k1[5], k2[5], form mask, pad
constexpr int TILE_EXCH_BYTES = TILE_ELEMS * 8;
constexpr int TILE_SMEM_BYTES = TILE_EXCH_BYTES + 4096 + 16 * 5 * 8;
// persistent kernels additionally stage TILE_STAGE_PAIRS of the 16 register pairs of the NEXT tile
constexpr int TILE_STAGE_PAIRS = 11;
constexpr int TILE_STAGE_OFFSET = TILE_EXCH_BYTES + 4096 + 1024;
constexpr int TILE_SMEM_BYTES_STAGED = TILE_STAGE_OFFSET + TILE_STAGE_PAIRS * TILE_THREADS * 16;

enum TileOpType : int {
  TOP_LOAD = 1, TOP_STORE = 2, TOP_INIT = 3, TOP_PHASE = 4,
  TOP_ROUND = 5, TOP_WT = 6, TOP_BT = 7, TOP_REDUCE = 8
};

struct TileOp {
  int16_t type;
  int16_t frame;  // 0 = A, 1 = B   (LOAD/STORE/INIT/PHASE/REDUCE)
  int16_t arg;    // ROUND: round index; PHASE/REDUCE: phase slot index
  int16_t mask;   // ROUND: register slots compiled in (canonical programs)
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
  int tau_stride;                       // reals per point  (nrounds * TILE_ROUND_REALS)
  int phs_stride;                       // doubles per point (2 * nphases + 1)
  int prefetch_dist;                    // tiles ahead whose lines are pulled into L2 (0 = off)
  unsigned ntiles;                      // tiles per point
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
__device__ __forceinline__ float2 eneg(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ double eneg(double a) { return -a; }

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

// element offset contributed by this thread's lane and warp bits in the given frame
__device__ __forceinline__ size_t thread_off(const TilePass& P, int frame, int warp, int lane) {
  size_t e = 0;
#pragma unroll
  for (int j = 0; j < 5; j++) e += (size_t)((lane >> j) & 1) << P.pos[1 + j];
  const int wofs = frame == 0 ? 6 : 10;
#pragma unroll
  for (int j = 0; j < 4; j++) e += (size_t)((warp >> j) & 1) << P.pos[wofs + j];
  return e;
}

__device__ __forceinline__ TileAddr tile_addr_off(const TilePass& P, size_t base, int frame, size_t thr) {
  TileAddr a;
  const int rofs = frame == 0 ? 10 : 6;
#pragma unroll
  for (int j = 0; j < 4; j++) a.roff[j] = (size_t)1 << P.pos[rofs + j];
  a.e0 = base + thr;
  return a;
}

__device__ __forceinline__ TileAddr tile_addr(const TilePass& P, size_t base, int frame, int warp, int lane) {
  return tile_addr_off(P, base, frame, thread_off(P, frame, warp, lane));
}

__device__ __forceinline__ size_t reg_off(const TileAddr& a, int q) {
  return ((q & 1) ? a.roff[0] : 0) + ((q & 2) ? a.roff[1] : 0) + ((q & 4) ? a.roff[2] : 0) +
         ((q & 8) ? a.roff[3] : 0);
}

// position (in 16-byte units) of chunk j of this thread's values inside a diagonal stream
__device__ __forceinline__ size_t stream_chunk(unsigned tile, int warp, int lane, int nchunks, int j) {
  return (((size_t)tile * 16 + warp) * nchunks + j) * 32 + lane;
}

// ---- per-thread context -------------------------------------------------------------------
template <typename Elem>
struct Ctx {
  typedef typename ElemTraits<Elem>::real real;
  typedef typename ElemTraits<Elem>::cplx cplx;
  const TilePass* P;
  Elem* state;
  const real* tau;     // this point's round coefficients
  const double* phs;   // this point's (gamma, mult) pairs followed by the reduction scale
  const void* tables;
  int tables_per_point;
  DiagInfo dinfo;
  double* partials;
  unsigned tile, point;
  unsigned ntiles;
  unsigned long long id, total, stride;   // persistent kernels: linear (point, tile) index walk
  Elem* state0;
  const real* tau0;
  const double* phs0;
  size_t point_stride;
  int warp, lane;
  size_t base;
  unsigned char* smem;
  // hoisted shared-memory byte addresses (see warp_transpose / block_transpose)
  unsigned wt_st, wt_ld, bt_st, bt_ld;
  size_t thr[2];   // hoisted per-thread element offsets for frames A and B
};

template <typename Elem>
__device__ __forceinline__ void ctx_smem_setup(Ctx<Elem>& c) {
  const unsigned lane = c.lane, warp = c.warp;
  // warp transpose: write chunk (16 B) index lane*16 + (q ^ (lane&15)); read chunk r*16 + ((lane>>1)^(r&15))
  c.wt_st = warp * 8192u + lane * 256u;          // ^ ((q ^ (lane&15)) << 4), folded below
  c.wt_st ^= (lane & 15u) << 4;
  c.wt_ld = warp * 8192u + (lane & 1u) * 8u;
  c.wt_ld ^= (lane >> 1) << 4;
  // block transpose: write element (warp,lane,r) at chunk (r + 32 warp + 512 (lane>>1)) ^ ((lane>>1)&7)
  const unsigned hi = lane >> 1;
  c.bt_st = (32u * warp + 512u * hi) * 16u + (lane & 1u) * 8u;
  c.bt_st ^= (hi & 7u) << 4;
  // read chunk (lane + 32 q + 512 warp) ^ (warp & 7)
  c.bt_ld = ((lane ^ (warp & 7u)) + 512u * warp) * 16u;
  c.thr[0] = thread_off(*c.P, 0, c.warp, c.lane);
  c.thr[1] = thread_off(*c.P, 1, c.warp, c.lane);
}

// ---- transposes ---------------------------------------------------------------------------
// warp-local: (lane l, reg r) -> (lane r, reg l); XOR-swizzled 16-byte chunks, conflict free.
template <typename Elem>
__device__ __forceinline__ void warp_transpose(const Ctx<Elem>& c, Elem (&v)[32]) {
  __syncwarp();
#pragma unroll
  for (int q = 0; q < 16; q++)
    store16(reinterpret_cast<Elem*>(c.smem + (c.wt_st ^ (unsigned)(q << 4))), v[2 * q], v[2 * q + 1]);
  __syncwarp();
#pragma unroll
  for (int r = 0; r < 32; r++)
    v[r] = *reinterpret_cast<const Elem*>(c.smem + (c.wt_ld ^ (unsigned)((r & 15) << 4)) + r * 256);
}

// block-wide: (warp w, lane l, reg r) -> (warp l>>1, lane r, reg (l&1)|(w<<1)).
template <typename Elem>
__device__ __forceinline__ void block_transpose(const Ctx<Elem>& c, Elem (&v)[32]) {
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 32; r++)
    *reinterpret_cast<Elem*>(c.smem + (c.bt_st ^ (unsigned)((r & 7) << 4)) + (r >> 3) * 128) = v[r];
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 16; q++)
    load16(reinterpret_cast<const Elem*>(c.smem + c.bt_ld + q * 512), v[2 * q], v[2 * q + 1]);
}

// ---- butterflies --------------------------------------------------------------------------
// canonical (straight-line) round: normal form only, slots selected at compile time
template <typename Elem, int MASK>
__device__ __forceinline__ void round_static(const Ctx<Elem>& c, Elem (&v)[32], int round) {
  typedef typename ElemTraits<Elem>::real real;
  const real* rp = c.tau + round * TILE_ROUND_REALS;
#pragma unroll
  for (int s = 0; s < 5; s++) {
    if (MASK & (1 << s)) {
      const real t = __ldg(rp + s);
#pragma unroll
      for (int i = 0; i < 32; i++) {
        if (!(i & (1 << s))) {
          Elem a = v[i], b = v[i | (1 << s)];
          v[i] = efma(t, b, a);
          v[i | (1 << s)] = efma(-t, a, b);
        }
      }
    }
  }
}

// interpreter round: per-slot form, idle slots skipped
template <typename Elem>
__device__ __forceinline__ void round_dynamic(const Ctx<Elem>& c, Elem (&v)[32], int round) {
  typedef typename ElemTraits<Elem>::real real;
  const real* rp = c.tau + round * TILE_ROUND_REALS;
  const int forms = (int)__ldg(rp + 10);
#pragma unroll
  for (int s = 0; s < 5; s++) {
    const real k1 = __ldg(rp + s), k2 = __ldg(rp + 5 + s);
    if ((forms >> s) & 1) {  // cot form: a' = k1 a + b ; b' = k2 b - a
#pragma unroll
      for (int i = 0; i < 32; i++) {
        if (!(i & (1 << s))) {
          Elem a = v[i], b = v[i | (1 << s)];
          v[i] = efma(k1, a, b);
          v[i | (1 << s)] = efma(k2, b, eneg(a));
        }
      }
    } else if (k1 != (real)0 || k2 != (real)0) {  // normal form: a' = a + k1 b ; b' = b + k2 a
#pragma unroll
      for (int i = 0; i < 32; i++) {
        if (!(i & (1 << s))) {
          Elem a = v[i], b = v[i | (1 << s)];
          v[i] = efma(k1, b, a);
          v[i | (1 << s)] = efma(k2, a, b);
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

// Visits this thread's NAMP diagonal values (raw integer k for U8/U32FX, double for F64).
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

// ---- the steps ----------------------------------------------------------------------------
template <typename Elem>
__device__ __forceinline__ void op_load(const Ctx<Elem>& c, Elem (&v)[32], int frame) {
  TileAddr ad = tile_addr_off(*c.P, c.base, frame, c.thr[frame]);
  const Elem* p = c.state + ad.e0;
#pragma unroll
  for (int q = 0; q < 16; q++) load16(p + reg_off(ad, q), v[2 * q], v[2 * q + 1]);
}

// Pulls the 1024 128-byte rows of a LATER tile (the one the next CTA scheduled on this SM will most
// likely own) from HBM into L2, so that its loads overlap this tile's arithmetic.  One CTA per SM
// cannot overlap its own load and compute phases; the L2 (126 MB) is the staging buffer instead.
template <typename Elem>
__device__ __forceinline__ void prefetch_later_tile(const Ctx<Elem>& c) {
  const TilePass& P = *c.P;
  if (P.prefetch_dist <= 0) return;
  const unsigned long long id = c.id + (unsigned)P.prefetch_dist;
  if (id >= c.total) return;
  const unsigned ntile = (unsigned)(id % c.ntiles), npoint = (unsigned)(id / c.ntiles);
  const size_t nbase = tile_base(P, ntile);
  const Elem* st = c.state0 + (size_t)npoint * c.point_stride;
#pragma unroll
  for (int k = 0; k < 2; k++) {
    const unsigned row = threadIdx.x + k * TILE_THREADS;  // 10 bits: tile-local bits 4..13
    size_t e = nbase;
#pragma unroll
    for (int j = 0; j < 10; j++) e += (size_t)((row >> j) & 1u) << P.pos[4 + j];
    asm volatile("prefetch.global.L2 [%0];" ::"l"(st + e));
  }
}

template <typename Elem>
__device__ __forceinline__ void op_store(const Ctx<Elem>& c, Elem (&v)[32], int frame) {
  TileAddr ad = tile_addr_off(*c.P, c.base, frame, c.thr[frame]);
  Elem* p = c.state + ad.e0;
#pragma unroll
  for (int q = 0; q < 16; q++) store16(p + reg_off(ad, q), v[2 * q], v[2 * q + 1]);
}

template <typename Elem>
__device__ __forceinline__ void op_init(const Ctx<Elem>& c, Elem (&v)[32], int frame) {
  typedef typename ElemTraits<Elem>::real real;
  constexpr int NAMP = ElemTraits<Elem>::NAMP;
  const TilePass& P = *c.P;
  TileAddr ad = tile_addr_off(P, c.base, frame, c.thr[frame]);
  const real A = (real)P.init_amp;
  const int pc0 = __popcll((unsigned long long)(ad.e0 >> P.qshift));
#pragma unroll
  for (int a = 0; a < NAMP; a++) {
    // tile bits are distinct element bits: popcount(x) = popcount(thread part) + popcount(reg part)
    const int regbits = ElemTraits<Elem>::IS_C64 ? a : (a << 1);
    const int pc = pc0 + __popc(regbits >> P.qshift);
    real re = A, im = (real)0;
    if (P.init_kind == INIT_DICKE && pc != P.dicke_k) re = (real)0;
    mul_i_pow(re, im, pc);
    amp_set(v, a, re, im);
  }
}

template <typename Elem, int DK>
__device__ __forceinline__ void op_phase(const Ctx<Elem>& c, Elem (&v)[32], int slot) {
  typedef typename ElemTraits<Elem>::real real;
  typedef typename ElemTraits<Elem>::cplx cplx;
  constexpr int NAMP = ElemTraits<Elem>::NAMP;
  const TilePass& P = *c.P;
  const double gamma = __ldg(c.phs + 2 * slot), mult = __ldg(c.phs + 2 * slot + 1);
  if (DK == DIAG_U8 || (DK == 0 && c.dinfo.kind == DIAG_U8)) {
    cplx* tab = reinterpret_cast<cplx*>(c.smem + TILE_EXCH_BYTES);
    __syncthreads();
    if (threadIdx.x < 256) {
      const cplx* t = reinterpret_cast<const cplx*>(c.tables) +
                      ((size_t)c.point * c.tables_per_point + P.table_id[slot]) * 256;
      tab[threadIdx.x] = t[threadIdx.x];
    }
    __syncthreads();
    for_each_diag<NAMP>(P.streams[slot], DIAG_U8, c.tile, c.warp, c.lane, [&](int a, unsigned k, double) {
      cplx ph = tab[k];
      real re, im;
      amp_get(v, a, re, im);
      amp_set(v, a, re * ph.x - im * ph.y, re * ph.y + im * ph.x);
    });
  } else if (DK == 0) {
    const double inv2pi = 0.15915494309189533576888;
    const double g0 = gamma * c.dinfo.c0 * inv2pi, gd = gamma * c.dinfo.delta * inv2pi;
    const real sg = (real)mult;
    const int kind = c.dinfo.kind;
    for_each_diag<NAMP>(P.streams[slot], kind, c.tile, c.warp, c.lane, [&](int a, unsigned k, double cd) {
      double turns = kind == DIAG_U32FX ? fma((double)k, gd, g0) : fma(cd, gd, g0);
      real s, co;
      phase_sincos(turns, s, co);
      s *= sg;
      co *= sg;
      real re, im;
      amp_get(v, a, re, im);
      amp_set(v, a, re * co - im * s, re * s + im * co);
    });
  }
}

template <typename Elem, int DK>
__device__ __forceinline__ void op_reduce(const Ctx<Elem>& c, Elem (&v)[32], int slot) {
  typedef typename ElemTraits<Elem>::real real;
  constexpr int NAMP = ElemTraits<Elem>::NAMP;
  const TilePass& P = *c.P;
  const real post = (real)__ldg(c.phs + 2 * P.nphases);  // scale still owed by the last layer
  real s0 = 0, s1 = 0, s2 = 0;
  double kmin = 1e300, kmax = -1e300;
  const int kind = DK == DIAG_U8 ? (int)DIAG_U8 : c.dinfo.kind;
  for_each_diag<NAMP>(P.streams[slot], kind, c.tile, c.warp, c.lane, [&](int a, unsigned k, double cd) {
    real re, im;
    amp_get(v, a, re, im);
    re *= post;
    im *= post;
    real p = re * re + im * im;
    real kv = kind == DIAG_F64 ? (real)cd : (real)k;
    double kdv = kind == DIAG_F64 ? cd : (double)k;
    s0 += p;
    s1 = fma(p, kv, s1);
    s2 = fma(p * kv, kv, s2);
    if (p > (real)0) { kmin = fmin(kmin, kdv); kmax = fmax(kmax, kdv); }
  });
  double r5[5] = {(double)s0, (double)s1, (double)s2, kmin, kmax};
  block_reduce5(r5, reinterpret_cast<double*>(c.smem + TILE_EXCH_BYTES + 4096));
  if (threadIdx.x == 0) {
    double* o = c.partials + ((size_t)c.point * c.ntiles + c.tile) * 5;
#pragma unroll
    for (int i = 0; i < 5; i++) o[i] = r5[i];
  }
}

template <typename Elem>
__device__ __forceinline__ void ctx_set_tile(Ctx<Elem>& c, unsigned long long id) {
  const TilePass& P = *c.P;
  c.id = id;
  c.tile = (unsigned)(id % c.ntiles);
  c.point = (unsigned)(id / c.ntiles);
  c.state = c.state0 + (size_t)c.point * c.point_stride;
  c.tau = c.tau0 + (size_t)c.point * P.tau_stride;
  c.phs = c.phs0 + (size_t)c.point * P.phs_stride;
  c.base = tile_base(P, c.tile);
}

template <typename Elem>
__device__ __forceinline__ void ctx_init(Ctx<Elem>& c, const TilePass& P, Elem* state, size_t point_stride,
                                         const typename ElemTraits<Elem>::real* tau, const double* phs,
                                         const void* tables, int tables_per_point, const DiagInfo& dinfo,
                                         double* partials, unsigned char* smem, unsigned long long total) {
  c.P = &P;
  c.lane = threadIdx.x & 31;
  c.warp = threadIdx.x >> 5;
  c.ntiles = P.ntiles;
  c.total = total;
  c.stride = gridDim.x;
  c.state0 = state;
  c.tau0 = tau;
  c.phs0 = phs;
  c.point_stride = point_stride;
  c.tables = tables;
  c.tables_per_point = tables_per_point;
  c.dinfo = dinfo;
  c.partials = partials;
  c.smem = smem;
  ctx_smem_setup(c);
  ctx_set_tile(c, blockIdx.x);
}

// ---- staged (software-pipelined) loads for the persistent kernels ---------------------------
__device__ __forceinline__ void cp_async16(unsigned smem_addr, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Asynchronous copies of register pairs [Q0, Q1) of tile `id` into this thread's private slots.
// Pairs < TILE_STAGE_PAIRS live in the dedicated staging area ("early" group, issued right after the
// current tile has been pulled into registers); the remaining pairs use the first 40 KB of the
// exchange buffer ("late" group, issued once the tile's last transpose is done with that buffer).
template <typename Elem, int Q0, int Q1>
__device__ __forceinline__ void stage_issue(const Ctx<Elem>& c, unsigned long long id, int frame) {
  if (id < c.total) {
    const TilePass& P = *c.P;
    const unsigned ntile = (unsigned)(id % c.ntiles), npoint = (unsigned)(id / c.ntiles);
    TileAddr ad = tile_addr_off(P, tile_base(P, ntile), frame, c.thr[frame]);
    const Elem* p = c.state0 + (size_t)npoint * c.point_stride + ad.e0;
    const unsigned sb = (unsigned)__cvta_generic_to_shared(c.smem);
    const unsigned early = sb + TILE_STAGE_OFFSET + threadIdx.x * 16u;
    // late slots sit inside the OWNING warp's 8 KB transpose region, so the only writers that can
    // touch them before they are consumed are the warp's own (program-ordered) transposes
    const unsigned late = sb + c.warp * 8192u + c.lane * 16u;
#pragma unroll
    for (int q = Q0; q < Q1; q++) {
      const unsigned dst = q < TILE_STAGE_PAIRS ? early + q * (TILE_THREADS * 16) : late + (q - TILE_STAGE_PAIRS) * 512;
      cp_async16(dst, p + reg_off(ad, q));
    }
  }
  cp_async_commit();
}

// persistent-kernel load: every register pair comes from the slots filled during the previous tile
template <typename Elem>
__device__ __forceinline__ void op_load_staged(const Ctx<Elem>& c, Elem (&v)[32], int frame) {
  cp_async_wait_all();
  const unsigned char* early = c.smem + TILE_STAGE_OFFSET + threadIdx.x * 16;
  const unsigned char* late = c.smem + c.warp * 8192 + c.lane * 16;
#pragma unroll
  for (int q = 0; q < 16; q++) {
    const unsigned char* src = q < TILE_STAGE_PAIRS ? early + q * (TILE_THREADS * 16) : late + (q - TILE_STAGE_PAIRS) * 512;
    load16(reinterpret_cast<const Elem*>(src), v[2 * q], v[2 * q + 1]);
  }
  __syncwarp();
  stage_issue<Elem, 0, TILE_STAGE_PAIRS>(c, c.id + c.stride, frame);
}

// ---- interpreter kernel -------------------------------------------------------------------
template <typename Elem>
__global__ void __launch_bounds__(TILE_THREADS, 1)
tile_pass_kernel(const __grid_constant__ TilePass P, Elem* __restrict__ state, size_t point_stride,
                 const typename ElemTraits<Elem>::real* __restrict__ tau, const double* __restrict__ phs,
                 const void* __restrict__ tables, int tables_per_point, DiagInfo dinfo,
                 double* __restrict__ partials) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  Ctx<Elem> c;
  ctx_init<Elem>(c, P, state, point_stride, tau, phs, tables, tables_per_point, dinfo, partials, smem_raw,
                 (unsigned long long)gridDim.x);
  Elem v[32];
  TileOp op = P.ops[0];
  for (int io = 0; io < P.nops; io++) {
    const TileOp cur = op;
    if (io + 1 < P.nops) op = P.ops[io + 1];  // fetch the next step early (constant-bank latency)
    switch (cur.type) {
      case TOP_LOAD: op_load<Elem>(c, v, cur.frame); prefetch_later_tile<Elem>(c); break;
      case TOP_STORE: op_store<Elem>(c, v, cur.frame); break;
      case TOP_INIT: op_init<Elem>(c, v, cur.frame); break;
      case TOP_PHASE: op_phase<Elem, 0>(c, v, cur.arg); break;
      case TOP_ROUND: round_dynamic<Elem>(c, v, cur.arg); break;
      case TOP_WT: warp_transpose<Elem>(c, v); break;
      case TOP_BT: block_transpose<Elem>(c, v); break;
      case TOP_REDUCE: op_reduce<Elem, 0>(c, v, cur.arg); break;
      default: break;
    }
  }
}

// ---- compile-time programs ----------------------------------------------------------------
struct COp {
  int type, frame, arg, mask;
};

constexpr int M_ALL = 0x1f, M_HI4 = 0x1e, M_TOP2 = 0x18;

// full cover of all 14 tile bits starting in frame F (ends in frame F^1)
//   R(all) WT R(all) BT R(hi4)
// high cover of the 10 upper tile bits:  R(hi4) WT R(top2) BT R(hi4)
struct ProgFirstInit {  // |+>/Dicke generated in registers, phase 1, full layer-1 cover of the tile
  static constexpr int last_bt() { int k = -1; for (int i = 0; i < N; i++) if (op(i).type == TOP_BT) k = i; return k; }
  static constexpr int N = 8;
  static constexpr COp op(int i) {
    constexpr COp o[N] = {{TOP_INIT, 0, 0, 0}, {TOP_PHASE, 0, 0, 0}, {TOP_ROUND, 0, 0, M_ALL}, {TOP_WT, 0, 0, 0},
                          {TOP_ROUND, 0, 1, M_ALL}, {TOP_BT, 0, 0, 0}, {TOP_ROUND, 0, 2, M_HI4}, {TOP_STORE, 1, 0, 0}};
    return o[i];
  }
};
struct ProgFirstLoad {
  static constexpr int last_bt() { int k = -1; for (int i = 0; i < N; i++) if (op(i).type == TOP_BT) k = i; return k; }
  static constexpr int N = 8;
  static constexpr COp op(int i) {
    constexpr COp o[N] = {{TOP_LOAD, 0, 0, 0}, {TOP_PHASE, 0, 0, 0}, {TOP_ROUND, 0, 0, M_ALL}, {TOP_WT, 0, 0, 0},
                          {TOP_ROUND, 0, 1, M_ALL}, {TOP_BT, 0, 0, 0}, {TOP_ROUND, 0, 2, M_HI4}, {TOP_STORE, 1, 0, 0}};
    return o[i];
  }
};
struct ProgInterior {
  static constexpr int last_bt() { int k = -1; for (int i = 0; i < N; i++) if (op(i).type == TOP_BT) k = i; return k; }
  static constexpr int N = 7;
  static constexpr COp op(int i) {
    constexpr COp o[N] = {{TOP_LOAD, 0, 0, 0}, {TOP_ROUND, 0, 0, M_HI4}, {TOP_WT, 0, 0, 0}, {TOP_ROUND, 0, 1, M_TOP2},
                          {TOP_BT, 0, 0, 0}, {TOP_ROUND, 0, 2, M_HI4}, {TOP_STORE, 1, 0, 0}};
    return o[i];
  }
};
struct ProgBoundary {
  static constexpr int last_bt() { int k = -1; for (int i = 0; i < N; i++) if (op(i).type == TOP_BT) k = i; return k; }
  static constexpr int N = 13;
  static constexpr COp op(int i) {
    constexpr COp o[N] = {{TOP_LOAD, 0, 0, 0}, {TOP_ROUND, 0, 0, M_HI4}, {TOP_WT, 0, 0, 0}, {TOP_ROUND, 0, 1, M_TOP2},
                          {TOP_BT, 0, 0, 0}, {TOP_ROUND, 0, 2, M_HI4}, {TOP_PHASE, 1, 0, 0}, {TOP_ROUND, 0, 3, M_ALL},
                          {TOP_WT, 0, 0, 0}, {TOP_ROUND, 0, 4, M_ALL}, {TOP_BT, 0, 0, 0}, {TOP_ROUND, 0, 5, M_HI4},
                          {TOP_STORE, 0, 0, 0}};
    return o[i];
  }
};
struct ProgFinal {
  static constexpr int last_bt() { int k = -1; for (int i = 0; i < N; i++) if (op(i).type == TOP_BT) k = i; return k; }
  static constexpr int N = 7;
  static constexpr COp op(int i) {
    constexpr COp o[N] = {{TOP_LOAD, 0, 0, 0}, {TOP_ROUND, 0, 0, M_HI4}, {TOP_WT, 0, 0, 0}, {TOP_ROUND, 0, 1, M_TOP2},
                          {TOP_BT, 0, 0, 0}, {TOP_ROUND, 0, 2, M_HI4}, {TOP_REDUCE, 1, 0, 0}};
    return o[i];
  }
};

template <typename Elem, typename Prog, int DK, int I>
struct ProgExec {
  static __device__ __forceinline__ void run(const Ctx<Elem>& c, Elem (&v)[32]) {
    if constexpr (I < Prog::N) {
      constexpr COp op = Prog::op(I);
      if constexpr (op.type == TOP_LOAD) { op_load_staged<Elem>(c, v, op.frame); prefetch_later_tile<Elem>(c); }
      else if constexpr (op.type == TOP_STORE) op_store<Elem>(c, v, op.frame);
      else if constexpr (op.type == TOP_INIT) op_init<Elem>(c, v, op.frame);
      else if constexpr (op.type == TOP_PHASE) op_phase<Elem, DK>(c, v, op.arg);
      else if constexpr (op.type == TOP_ROUND) {
        // complex128: register slot 0 of frames A/B is the re/im bit, never a qubit
        round_static<Elem, op.mask>(c, v, op.arg);
      }
      else if constexpr (op.type == TOP_WT) warp_transpose<Elem>(c, v);
      else if constexpr (op.type == TOP_BT) {
        block_transpose<Elem>(c, v);
        if constexpr (Prog::op(0).type == TOP_LOAD && I == Prog::last_bt()) {
          __syncthreads();  // every warp has read its share of the exchange buffer
          stage_issue<Elem, TILE_STAGE_PAIRS, 16>(c, c.id + c.stride, Prog::op(0).frame);
        }
      }
      else if constexpr (op.type == TOP_REDUCE) op_reduce<Elem, DK>(c, v, op.arg);
      ProgExec<Elem, Prog, DK, I + 1>::run(c, v);
    }
  }
};

template <typename Elem, typename Prog, int DK>
__global__ void __launch_bounds__(TILE_THREADS, 1)
tile_prog_kernel(const __grid_constant__ TilePass P, Elem* __restrict__ state, size_t point_stride,
                 const typename ElemTraits<Elem>::real* __restrict__ tau, const double* __restrict__ phs,
                 const void* __restrict__ tables, int tables_per_point, DiagInfo dinfo,
                 double* __restrict__ partials, unsigned n_points) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  Ctx<Elem> c;
  const unsigned long long total = (unsigned long long)P.ntiles * n_points;
  ctx_init<Elem>(c, P, state, point_stride, tau, phs, tables, tables_per_point, dinfo, partials, smem_raw, total);
  constexpr COp first = Prog::op(0);
  if constexpr (first.type == TOP_LOAD) {
    stage_issue<Elem, 0, TILE_STAGE_PAIRS>(c, blockIdx.x, first.frame);
    stage_issue<Elem, TILE_STAGE_PAIRS, 16>(c, blockIdx.x, first.frame);
  }
  Elem v[32];
  for (unsigned long long id = blockIdx.x; id < total; id += gridDim.x) {
    ctx_set_tile(c, id);
    ProgExec<Elem, Prog, DK, 0>::run(c, v);
  }
}

// host-side: does a planned pass match a compiled program exactly?
template <typename Prog>
inline bool prog_matches(const TilePass& tp) {
  if (tp.nops != Prog::N) return false;
  for (int i = 0; i < Prog::N; i++) {
    const COp o = Prog::op(i);
    if (tp.ops[i].type != o.type) return false;
    if (o.type == TOP_ROUND) {
      if (tp.ops[i].arg != o.arg || tp.ops[i].mask != o.mask) return false;
    } else if (o.type == TOP_PHASE || o.type == TOP_REDUCE) {
      if (tp.ops[i].arg != o.arg || tp.ops[i].frame != o.frame) return false;
    } else if (o.type == TOP_LOAD || o.type == TOP_STORE || o.type == TOP_INIT) {
      if (tp.ops[i].frame != o.frame) return false;
    }
  }
  return true;
}

// Builds one permuted diagonal stream: stream[(tile, warp, chunk, lane, i)] = diag[x(tile,warp,lane,amp)]
// using exactly the addressing of the pass kernels (same helper functions).
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
  for (int j = 0; j < nchunks; j++) {
    __align__(16) unsigned char buf[16];
    for (int i = 0; i < V; i++) {
      const int a = j * V + i;
      // amplitude a -> register pair q: complex64: q = a>>1 with low bit a&1; complex128: q = a
      const int q = (namp == 32) ? (a >> 1) : a;
      const size_t e = ad.e0 + reg_off(ad, q) + ((namp == 32) ? (a & 1) : 0);
      const unsigned char* src = diag + (e >> P.qshift) * VALBYTES;
#pragma unroll
      for (int b = 0; b < VALBYTES; b++) buf[i * VALBYTES + b] = src[b];
    }
    reinterpret_cast<uint4*>(stream)[stream_chunk(tile, warp, lane, nchunks, j)] = *reinterpret_cast<uint4*>(buf);
  }
}
