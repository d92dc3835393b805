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
#define QAOA_MAX_RANKS 8
constexpr int TILE_ROUND_REALS = 12;  // k1[5], k2[5], form mask, pad
constexpr int TILE_EXCH_BYTES = TILE_ELEMS * 8;
constexpr int TILE_SMEM_BYTES = TILE_EXCH_BYTES + 4096 + 16 * 5 * 8;   // interpreter kernel
constexpr int TILE_SMEM_MAX = 232448;                                   // 227 KB per CTA on sm_100a
constexpr int TILE_CANON_ROUNDS = 8;                                    // rounds a compiled program may hold

enum TileOpType : int {
  TOP_LOAD = 1, TOP_STORE = 2, TOP_INIT = 3, TOP_PHASE = 4,
  TOP_ROUND = 5, TOP_WT = 6, TOP_BT = 7, TOP_REDUCE = 8,
  TOP_XT = 9,    // block transpose frame A -> frame C
  TOP_XTI = 10   // block transpose frame C -> frame A
};
// Frames (which tile-local bit sits where; t0..t13 = tile bits, t0 = element bit 0):
//   A: reg {t0, t10..t13}  lane {t1..t5}      warp {t6..t9}      (global access: 16 B per thread)
//   B: reg {t0, t6..t9}    lane {t1..t5}      warp {t10..t13}    (global access: 16 B per thread)
//   C: reg {t5, t6..t9}    lane {t0, t1..t4}  warp {t10..t13}    (register-only frame, complex64)
//   after WT (from A or B): reg {t1..t5}
// Tiles with 4 low bits (t0..t3 low, t4..t13 high) use two more frames for their high-bit passes:
//   A4 (3): reg {t9, t10..t13}  lane {t0..t3, t4}  warp {t5..t8}  (global access: 8 B per thread, 128-byte rows)
//   C4 (4): reg {t4, t5..t8}    lane {t9, t0..t3}  warp {t10..t13} (register-only; A4 <-> C4 is the XT / XTI transpose)

struct TileOp {
  int16_t type;
  int16_t frame;  // 0 = A, 1 = B, 2 = C, 3 = A4, 4 = C4   (LOAD/STORE/INIT/PHASE/REDUCE)
  int16_t arg;    // ROUND: round index; PHASE/REDUCE: phase slot index
  int16_t mask;   // ROUND: register slots compiled in (canonical programs)
};

struct TilePass {
  int nops;
  TileOp ops[TILE_MAX_OPS];
  int pos[TILE_BITS];  // element-bit position of each tile-local bit (ascending, pos[0] == 0)
  int nfree;
  int freepos[QAOA_MAX_QUBITS + 1];  // element-bit positions of the non-tile bits (ascending)
  // the non-tile bits always form at most two contiguous runs; the kernels address tiles through them
  int fr_lo_bits, fr_lo_shift, fr_hi_shift;
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
  // Sharded registers (one engine per GPU, the register split on its top qubits): element bits >=
  // shard_bits select the owning rank.  A pass whose tile contains those bits is a CROSS pass: its
  // tile ids are global (this rank owns [tile0, tile0 + ntiles)) and every access goes through the
  // peer table (NVLink peer memory); all other passes see only the local shard.
  int shard_bits;                       // element bits per shard (0: not sharded)
  int rank_popc;                        // popcount of this rank's index (initial-state generation)
  unsigned tile0;
  void* peers[QAOA_MAX_RANKS];          // state base of every rank
  const void* dpeers[QAOA_MAX_RANKS];   // cost-diagonal base of every rank (stream construction)
  // Symmetric (Z2) registers, see "symmetric registers" below: zbit = element bit of the VIRTUAL top
  // qubit (0 = off); it is tile bit 13 of the low tile shape and of no other shape.
  int zbit;
  int zsign;                            // i^(2-n) = +1 / -1 (n even)
  int znreal;                           // stored qubits of the WHOLE register (n - 1)
  // Sharded symmetric registers ("twisted" coordinates, see below): rank bit j of a stored index is
  // x'[EL + j] XOR x'[0], so that the mirror image of a shard is the shard itself.
  int zdelta;                           // g - 2 popc(rank): popcount owed by elements with index bit 0 set (0: untwisted)
  int ztw;                              // 1: sigma does not alternate with index bit 0 (g odd)
  int xgbits;                           // cross pass of a twisted register: g (its rounds are twisted rounds), else 0
  // element offsets of the upper register bits, precomputed by the host (tilepass_set_offsets): frames A / B
  // (register bits 1..4) and A4 (register bits 0..4) -- constant-bank operands instead of per-tile shifts
  unsigned long long roffe[2][4];
  unsigned long long roffe8[5];
  // Chunked launches (sharded registers: the cross pass of one chunk overlaps the local passes of another, engine.cu
  // "chunk pipeline"): this launch walks only the tiles of chunk ch_c of 2^ch_m (ch_m <= 3).  A chunk is the set of local indices
  // with  z[v_i] XOR z[u] = bit i of ch_c  for ch_m bit pairs -- invariant under the mirror pairing of a symmetric
  // register, and u, v_i are bits no chunk-closed pass mixes.  The launch enumerates j in [0, ntiles) with ntiles =
  // (tiles of the shape) >> ch_m; the tile index is j with the v_i bits inserted (positions ch_pv, ascending, in tile-
  // index coordinates) and set to ch_c XOR the u bit (position ch_pu).  ch_part0: first partial-result slot (REDUCE).
  int ch_m;
  int ch_c;
  int ch_pu;
  int ch_pv[3];
  int ch_part0;
};

__device__ __forceinline__ unsigned chunk_tile(const TilePass& P, unsigned j) {
  if (P.ch_m == 0) return j;
  unsigned t = j;
#pragma unroll
  for (int i = 0; i < 3; i++)
    if (i < P.ch_m) { const int pv = P.ch_pv[i]; t = ((t >> pv) << (pv + 1)) | (t & ((1u << pv) - 1u)); }
  const unsigned ub = (t >> P.ch_pu) & 1u;
#pragma unroll
  for (int i = 0; i < 3; i++)
    if (i < P.ch_m) t |= ((((unsigned)P.ch_c >> i) & 1u) ^ ub) << P.ch_pv[i];
  return t;
}

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

// Shared-memory layout of the compiled-program (persistent) kernels:
//   [exchange buffer 128 KB][phase table 256 cplx][reduction scratch][round coefficients of the
//   current point][staging area: STAGE_PAIRS of the 16 register pairs of the NEXT tile]
template <typename Elem>
struct TileSmem {
  typedef typename ElemTraits<Elem>::real real;
  typedef typename ElemTraits<Elem>::cplx cplx;
  static constexpr int TABLE_OFF = TILE_EXCH_BYTES;
  static constexpr int TABLE_BYTES = 256 * (int)sizeof(cplx);
  static constexpr int SCRATCH_OFF = TABLE_OFF + TABLE_BYTES;
  static constexpr int SCRATCH_BYTES = 16 * 5 * 8;
  static constexpr int TAU_OFF = SCRATCH_OFF + SCRATCH_BYTES;
  static constexpr int TAU_BYTES = TILE_CANON_ROUNDS * TILE_ROUND_REALS * (int)sizeof(real);
  static constexpr int STAGE_OFF = TAU_OFF + TAU_BYTES;
  static constexpr int STAGE_PAIRS = (TILE_SMEM_MAX - STAGE_OFF) / (TILE_THREADS * 16);   // 12 (c64) / 11 (c128)
  static constexpr int TOTAL = STAGE_OFF + STAGE_PAIRS * TILE_THREADS * 16;
  static_assert(STAGE_OFF % 16 == 0, "staging area must be 16-byte aligned");
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
  return ((size_t)(tile & ((1u << P.fr_lo_bits) - 1u)) << P.fr_lo_shift) |
         ((size_t)(tile >> P.fr_lo_bits) << P.fr_hi_shift);
}

// element index -> pointer.  CROSS passes translate through the peer table (one point only).
template <int CROSS, typename Elem>
__device__ __forceinline__ Elem* eptr(const TilePass& P, Elem* point_base, size_t e) {
  if constexpr (CROSS == 1)
    return reinterpret_cast<Elem*>(P.peers[e >> P.shard_bits]) + (e & (((size_t)1 << P.shard_bits) - 1));
  else
    return point_base + e;
}

// host: derive the two-run description from freepos (false if the free bits need more than two runs)
inline bool tilepass_set_free_runs(TilePass& tp) {
  tp.fr_lo_bits = 0; tp.fr_lo_shift = 0; tp.fr_hi_shift = 0;
  if (tp.nfree == 0) return true;
  int i = 1;
  while (i < tp.nfree && tp.freepos[i] == tp.freepos[i - 1] + 1) i++;
  tp.fr_lo_bits = i;
  tp.fr_lo_shift = tp.freepos[0];
  if (i == tp.nfree) return true;
  tp.fr_hi_shift = tp.freepos[i];
  for (int j = i + 1; j < tp.nfree; j++) if (tp.freepos[j] != tp.freepos[j - 1] + 1) return false;
  return true;
}

// tile-local bit held by each lane / warp / register bit of a frame (see the frame table above)
__host__ __device__ inline int frame_reg_bit(int frame, int j);
// host: fill the precomputed register-bit offsets from pos (call whenever pos is final)
inline void tilepass_set_offsets(TilePass& tp) {
  for (int j = 0; j < 4; j++) {
    tp.roffe[0][j] = 1ull << tp.pos[10 + j];
    tp.roffe[1][j] = 1ull << tp.pos[6 + j];
  }
  for (int j = 0; j < 5; j++) tp.roffe8[j] = 1ull << tp.pos[frame_reg_bit(3, j)];
}
__host__ __device__ inline int frame_lane_bit(int frame, int j) {
  if (frame == 2 || frame == 3) return j;
  if (frame == 4) return j == 0 ? 9 : j - 1;
  return 1 + j;
}
__host__ __device__ inline int frame_warp_bit(int frame, int j) {
  return (frame == 0 ? 6 : (frame == 3 ? 5 : 10)) + j;
}
__host__ __device__ inline int frame_reg_bit(int frame, int j) {
  if (frame == 2) return 5 + j;
  if (frame == 3) return 9 + j;
  if (frame == 4) return 4 + j;
  return j == 0 ? 0 : (frame == 0 ? 9 : 5) + j;
}

// element offset contributed by this thread's lane and warp bits in the given frame
__device__ __forceinline__ size_t thread_off(const TilePass& P, int frame, int warp, int lane) {
  size_t e = 0;
#pragma unroll
  for (int j = 0; j < 5; j++) e += (size_t)((lane >> j) & 1) << P.pos[frame_lane_bit(frame, j)];
#pragma unroll
  for (int j = 0; j < 4; j++) e += (size_t)((warp >> j) & 1) << P.pos[frame_warp_bit(frame, j)];
  return e;
}

// element offset of register `r` (0..31) of a frame
__device__ __forceinline__ size_t reg_elem_off(const TilePass& P, int frame, int r) {
  size_t e = 0;
#pragma unroll
  for (int j = 0; j < 5; j++) e += (size_t)((r >> j) & 1) << P.pos[frame_reg_bit(frame, j)];
  return e;
}

// frames A and B only (register bit 0 = element bit 0, 16-byte accesses)
__device__ __forceinline__ TileAddr tile_addr_off(const TilePass& P, size_t base, int frame, size_t thr) {
  TileAddr a;
#pragma unroll
  for (int j = 0; j < 4; j++) a.roff[j] = (size_t)P.roffe[frame == 0 ? 0 : 1][j];
  a.e0 = base + thr;
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

// ---- symmetric (Z2) registers ---------------------------------------------------------------
// When cost(x) = cost(~x), the mixer is a sum of X_q and the initial state is |+>^n (MaxCut), the state
// keeps psi(x) = psi(~x) through every layer: only the half with the top qubit = 0 is stored
// (2^(n-1) amplitudes, index x' = low n-1 bits).  The mixer on the top qubit pairs x' with ~x':
//     psi~'(x') = c * (psi~(x') + t * sigma(x') * psi~(~x')),   sigma(x') = i^(2-n) * (-1)^popc(x')   (n even)
// (S frame of the stored half; sigma(~x') = -sigma(x'), so the pair (x', ~x') sees the usual lifted rotation
// with t -> t * sigma).  The LOW tile shape carries this pairing as a VIRTUAL tile bit t13 = element bit
// `zbit`: its tiles are 13 real low bits, a tile = one contiguous run of 2^13 elements plus its mirror image
// (the run at the complemented base, walked backwards).  A virtual element index e with bit zbit set means
// "the stored element at fold(e)", fold = clear zbit, complement every qubit bit below it.  t13 is a register
// bit of frame A (slot 4) -- the butterfly is register local there -- and a warp bit of frames B and C.
// Every access of the mirror image is as coalesced as the plain one (lanes run backwards); a 16-byte pair of
// complex64 elements arrives with its halves swapped.
// Inside the mirror image a virtual tile bit = 0 is a REAL qubit bit = 1, so the two elements of every ordinary
// butterfly change roles there: mirrored elements (register bit 4 in frame A, lane bit 4 after its warp
// transpose, warp bit 3 in frames B and C) take t -> -t in every round of the low tile.
//
// Sharded symmetric registers (complex64): a plain split of the stored half on its top g bits would put the mirror
// image of rank r's shard on rank ~r.  Instead the shards use TWISTED coordinates: stored index x' <-> (z, r) with
// z = the EL low bits of x' and rank bit r_j = x'[EL + j] XOR x'[0].  Complementing x' complements z and leaves r
// alone, so every shard is a symmetric register of its own (virtual bit = local element bit EL) and the low-tile
// passes never leave the GPU.  What changes (the S frame is still i^popc(x')):
//   * popc(x') = popc(z) + popc(r) + z_0 * zdelta, zdelta = g - 2 popc(r)   (INIT; sigma loses its z_0 factor for odd g)
//   * qubit 0 pairs (z_0 = 0, r) with (z_0 = 1, ~r): it is mixed in the CROSS passes, where z_0 and the rank bits
//     are all register bits of frame A (round_xtw_*), and is idle in the low tile;
//   * a rank qubit j is 0 where r_j = z_0: its butterflies take t -> -t in registers with z_0 = 1.
// The cost diagonal is generated in the same coordinates (gen_maxkcut_kernel, tw_mask).
struct Z2Addr {
  size_t b0, b1;    // pair base with the register's t13 clear / set (frame A); frame B: b0 only
  size_t roff[4];   // frame A: offsets of t10..t12; frame B: offsets of t6..t9, negated in mirrored threads
  bool mirrored;    // frame B: this thread's t13 (a warp bit) is set
};

__device__ __forceinline__ Z2Addr z2_addr(const TilePass& P, size_t base, int frame, size_t thr) {
  Z2Addr a;
  const size_t zpair = (((size_t)1 << P.zbit) - 1) & ~(size_t)1;   // complement mask of a 16-byte pair base
  if (frame == 0) {
#pragma unroll
    for (int j = 0; j < 3; j++) a.roff[j] = (size_t)P.roffe[0][j];
    a.roff[3] = 0;
    a.b0 = base + thr;
    a.b1 = a.b0 ^ zpair;   // the offsets' bits are set here: the mirror image walks DOWN from it
    a.mirrored = false;
  } else {
    const size_t v = base + thr;
    a.mirrored = ((v >> P.zbit) & 1) != 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const size_t r = (size_t)P.roffe[1][j];
      a.roff[j] = a.mirrored ? (size_t)0 - r : r;
    }
    a.b0 = a.mirrored ? (v ^ (((size_t)1 << P.zbit) | zpair)) : v;
    a.b1 = a.b0;
  }
  return a;
}

// stored element index of the 16-byte pair q (registers 2q, 2q+1) of this thread
__device__ __forceinline__ size_t z2_pair_index(const Z2Addr& a, int frame, int q) {
  if (frame == 0) {
    const size_t off = ((q & 1) ? a.roff[0] : 0) + ((q & 2) ? a.roff[1] : 0) + ((q & 4) ? a.roff[2] : 0);
    return (q & 8) ? a.b1 - off : a.b0 + off;
  }
  return a.b0 + ((q & 1) ? a.roff[0] : 0) + ((q & 2) ? a.roff[1] : 0) + ((q & 4) ? a.roff[2] : 0) +
         ((q & 8) ? a.roff[3] : 0);
}

// complex64 only: the two elements of a mirrored pair sit in reverse order (complex128: bit 0 is re / im)
template <typename Elem>
__device__ __forceinline__ bool z2_swapped(const Z2Addr& a, int frame, int q) {
  if (!ElemTraits<Elem>::IS_C64) return false;
  return frame == 0 ? (q & 8) != 0 : a.mirrored;
}

// virtual element index -> stored element index (stream construction; qshift low bits are not qubits)
__host__ __device__ inline size_t z2_fold(size_t e, int zbit, int qshift) {
  if (zbit > 0 && ((e >> zbit) & 1)) e ^= (((size_t)2 << zbit) - 1) & ~(((size_t)1 << qshift) - 1);
  return e;
}

// ---- per-thread context -------------------------------------------------------------------
template <typename Elem>
struct Ctx {
  typedef typename ElemTraits<Elem>::real real;
  typedef typename ElemTraits<Elem>::cplx cplx;
  const TilePass* P;
  Elem* state;
  const real* tau;     // this point's round coefficients (global memory)
  const double* phs;   // this point's (gamma, mult) pairs followed by the reduction scale
  const void* tables;
  int tables_per_point;
  DiagInfo dinfo;
  double* partials;
  unsigned tile, point;
  unsigned ntiles, tmask;
  int tshift;                             // ntiles = 1 << tshift
  unsigned long long id, total, stride;   // persistent kernels: linear (point, tile) index walk
  Elem* state0;
  const real* tau0;
  const double* phs0;
  size_t point_stride;
  int warp, lane;
  size_t base;
  unsigned char* smem;
  unsigned char* tab;   // 256-entry phase table in shared memory
  double* scratch;      // block-reduction scratch in shared memory
  // hoisted shared-memory byte addresses (see warp_transpose / block_transpose)
  unsigned wt_st, wt_ld, bt_st, bt_ld;
  size_t thr[5];   // hoisted per-thread element offsets for frames A, B, C, A4, C4
  real zsig;       // symmetric registers: sigma of (this thread, register 0) in frame A, see round_static
  real zsigo;      // ... of register 1 (= -zsig unless the shard is twisted with an odd number of rank bits)
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
  c.thr[2] = thread_off(*c.P, 2, c.warp, c.lane);
  c.thr[3] = thread_off(*c.P, 3, c.warp, c.lane);
  c.thr[4] = thread_off(*c.P, 4, c.warp, c.lane);
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

// frame A -> frame C: (warp w, lane l, reg r) -> (warp r>>1, lane ((l&15)<<1)|(r&1), reg (w<<1)|(l>>4)).
// Element (w, l, r) lives at byte (r>>1)*8192 + (32 w + l)*16 + (r&1)*8: every warp-wide access of
// both sides covers contiguous bytes, so no swizzle is needed and all offsets are immediates.
template <typename Elem>
__device__ __forceinline__ void xt_a_to_c(const Ctx<Elem>& c, Elem (&v)[32]) {
  __syncthreads();
  unsigned char* st = c.smem + threadIdx.x * 16u;
#pragma unroll
  for (int q = 0; q < 16; q++) store16(reinterpret_cast<Elem*>(st + q * 8192), v[2 * q], v[2 * q + 1]);
  __syncthreads();
  const unsigned char* ld = c.smem + c.warp * 8192u + c.lane * 8u;
#pragma unroll
  for (int r = 0; r < 32; r++) v[r] = *reinterpret_cast<const Elem*>(ld + r * 256);
}

// frame C -> frame A (inverse of the above, same layout)
template <typename Elem>
__device__ __forceinline__ void xt_c_to_a(const Ctx<Elem>& c, Elem (&v)[32]) {
  __syncthreads();
  unsigned char* st = c.smem + c.warp * 8192u + c.lane * 8u;
#pragma unroll
  for (int r = 0; r < 32; r++) *reinterpret_cast<Elem*>(st + r * 256) = v[r];
  __syncthreads();
  const unsigned char* ld = c.smem + threadIdx.x * 16u;
#pragma unroll
  for (int q = 0; q < 16; q++) load16(reinterpret_cast<const Elem*>(ld + q * 8192), v[2 * q], v[2 * q + 1]);
}

// ---- butterflies --------------------------------------------------------------------------
// canonical (straight-line) round: normal form only, slots selected at compile time; the
// coefficients of the current point sit in shared memory (see point_setup)
template <typename Elem, int MASK, bool ZS = false, bool ZR = false, bool ZT = false>
__device__ __forceinline__ void round_static(const typename ElemTraits<Elem>::real* tau_s, Elem (&v)[32], int round,
                                             typename ElemTraits<Elem>::real zsig = 0,
                                             typename ElemTraits<Elem>::real msign = 1,
                                             typename ElemTraits<Elem>::real zsigo = 0) {
  typedef typename ElemTraits<Elem>::real real;
  const real* rp = tau_s + round * TILE_ROUND_REALS;
#pragma unroll
  for (int s = 0; s < 5; s++) {
    if (ZS && s == 4 && (MASK & 16)) {
      // symmetric register, frame A: slot 4 pairs an element with its mirror image, t -> t * sigma(x');
      // sigma alternates with the parity of the amplitude's index bits held in the register number
      // (zsigo: sigma of register 1 = -zsig, except in twisted shards with an odd number of rank bits)
      const real tz = rp[4] * zsig, tzo = ElemTraits<Elem>::IS_C64 ? rp[4] * zsigo : tz;
#pragma unroll
      for (int i = 0; i < 16; i++) {
        const real tb = (ElemTraits<Elem>::IS_C64 && (i & 1)) ? tzo : tz;
        const real tt = (__builtin_popcount(i >> 1) & 1) ? -tb : tb;
        Elem a = v[i], b = v[i | 16];
        v[i] = efma(tt, b, a);
        v[i | 16] = efma(-tt, a, b);
      }
    } else if (MASK & (1 << s)) {
      // ZT: this thread holds mirrored elements (msign = -1);  ZR: registers 16..31 do (frame A)
      const real t = ZT ? rp[s] * msign : rp[s];
#pragma unroll
      for (int i = 0; i < 32; i++) {
        if (!(i & (1 << s))) {
          const real tt = (ZR && (i & 16)) ? -t : t;
          Elem a = v[i], b = v[i | (1 << s)];
          v[i] = efma(tt, b, a);
          v[i | (1 << s)] = efma(-tt, a, b);
        }
      }
    }
  }
}

__device__ __forceinline__ float2 esign(float s, float2 a) { return make_float2(s * a.x, s * a.y); }
__device__ __forceinline__ double esign(double s, double a) { return s * a; }

// sigma of register i (frame A, i < 16) relative to the thread: alternates with the popcount of the amplitude's index
// bits held in the register number (twisted shards with odd g: not with index bit 0)
template <typename Elem>
__device__ __forceinline__ typename ElemTraits<Elem>::real z2_sigma_reg(const Ctx<Elem>& c, int i) {
  typedef typename ElemTraits<Elem>::real real;
  const real b = (ElemTraits<Elem>::IS_C64 && (i & 1)) ? c.zsigo : c.zsig;
  return (__builtin_popcount(i >> 1) & 1) ? -b : b;
}

// Symmetric registers (Z): zslot = slot 4 of this round is the mirror pairing (frame A; see round_static): the
// mirrored partner enters and leaves the butterfly multiplied by sigma = +-1, which covers both forms.  Mirrored
// elements of ordinary slots (zr: registers 16..31; otherwise the whole thread when msign = -1) have the roles of
// a and b exchanged, which is the same butterfly with b negated on the way in and out.
template <typename Elem, bool Z = false>
__device__ __forceinline__ void round_dynamic(const Ctx<Elem>& c, Elem (&v)[32], int round, bool zslot = false,
                                              bool zr = false, typename ElemTraits<Elem>::real msign = 1) {
  typedef typename ElemTraits<Elem>::real real;
  const real* rp = c.tau + round * TILE_ROUND_REALS;
  const int forms = (int)__ldg(rp + 10);
#pragma unroll
  for (int s = 0; s < 5; s++) {
    const real k1 = __ldg(rp + s), k2 = __ldg(rp + 5 + s);
    const bool vslot = Z && s == 4 && zslot;
    if (Z) {
#pragma unroll
      for (int i = 0; i < 32; i++) {
        if (!(i & (1 << s))) {
          real sg = vslot ? z2_sigma_reg<Elem>(c, i)
                          : (zr ? ((i & 16) ? (real)-1 : (real)1) : msign);
          v[i | (1 << s)] = esign(sg, v[i | (1 << s)]);
        }
      }
    }
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
    if (Z) {
#pragma unroll
      for (int i = 0; i < 32; i++) {
        if (!(i & (1 << s))) {
          real sg = vslot ? z2_sigma_reg<Elem>(c, i)
                          : (zr ? ((i & 16) ? (real)-1 : (real)1) : msign);
          v[i | (1 << s)] = esign(sg, v[i | (1 << s)]);
        }
      }
    }
  }
}

// ---- cross passes of a twisted (sharded symmetric) register, frame A: register bit 0 is z_0, the top G register
// bits are the rank bits.  Rank qubit butterflies take t -> -t where z_0 = 1; qubit 0 pairs register i (z_0 = 0)
// with register i ^ (1 | rank bits); the remaining slots are ordinary local qubits.  Coefficients: slot s of the
// round as usual (slot 0 = qubit 0).
template <typename Elem, int G>
__device__ __forceinline__ void round_xtw_static(const typename ElemTraits<Elem>::real* tau_s, Elem (&v)[32], int round) {
  typedef typename ElemTraits<Elem>::real real;
  const real* rp = tau_s + round * TILE_ROUND_REALS;
  constexpr int RM = (0x1f << (5 - G)) & 0x1f;
#pragma unroll
  for (int s = 1; s < 5; s++) {   // slots below 5 - G are ordinary local qubits (the planner may place some here)
    const real t = rp[s];
#pragma unroll
    for (int i = 0; i < 32; i++) {
      if (!(i & (1 << s))) {
        const real tt = (s >= 5 - G && (i & 1)) ? -t : t;
        Elem a = v[i], b = v[i | (1 << s)];
        v[i] = efma(tt, b, a);
        v[i | (1 << s)] = efma(-tt, a, b);
      }
    }
  }
  const real t0 = rp[0];
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    Elem a = v[i], b = v[i ^ (1 | RM)];
    v[i] = efma(t0, b, a);
    v[i ^ (1 | RM)] = efma(-t0, a, b);
  }
}

template <typename Elem, int G>
__device__ __forceinline__ void round_xtw_dynamic(const Ctx<Elem>& c, Elem (&v)[32], int round) {
  typedef typename ElemTraits<Elem>::real real;
  const real* rp = c.tau + round * TILE_ROUND_REALS;
  const int forms = (int)__ldg(rp + 10);
  constexpr int RM = (0x1f << (5 - G)) & 0x1f;
  auto bfly = [&](Elem& a, Elem& b, real k1, real k2, bool cot) {
    Elem x = a, y = b;
    if (cot) { a = efma(k1, x, y); b = efma(k2, y, eneg(x)); }       // a' = k1 a + b ; b' = k2 b - a
    else { a = efma(k1, y, x); b = efma(k2, x, y); }                 // a' = a + k1 b ; b' = b + k2 a
  };
#pragma unroll
  for (int s = 1; s < 5; s++) {   // slots below 5 - G are ordinary local qubits
    const real k1 = __ldg(rp + s), k2 = __ldg(rp + 5 + s);
    const bool cot = (forms >> s) & 1;
    if (!cot && k1 == (real)0 && k2 == (real)0) continue;
#pragma unroll
    for (int i = 0; i < 32; i++) {
      if (!(i & (1 << s))) {
        if (s >= 5 - G && (i & 1)) bfly(v[i | (1 << s)], v[i], k1, k2, cot);   // z_0 = 1: the element with the rank bit SET has the qubit = 0
        else bfly(v[i], v[i | (1 << s)], k1, k2, cot);
      }
    }
  }
  {
    const real k1 = __ldg(rp), k2 = __ldg(rp + 5);
    const bool cot = forms & 1;
    if (cot || k1 != (real)0 || k2 != (real)0) {
#pragma unroll
      for (int i = 0; i < 32; i += 2) bfly(v[i], v[i ^ (1 | RM)], k1, k2, cot);
    }
  }
}

// ---- amplitude access helpers ---------------------------------------------------------------
// complex64: register a IS amplitude a (every frame).  complex128: amplitude a = registers
// (2a, 2a+1), frames A and B only (register bit 0 is the re/im bit there).
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

// complex64: range reduction in double (turns), then the hardware sine / cosine (MUFU): absolute error
// ~2^-21 on [-pi, pi], two orders of magnitude inside the 1e-4 energy tolerance of the complex64 path
__device__ __forceinline__ void phase_sincos(double turns, float& s, float& c) {
  double fr = turns - rint(turns);
  __sincosf(6.283185307179586f * (float)fr, &s, &c);
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
template <typename Elem, int CROSS>
__device__ __forceinline__ void op_load(const Ctx<Elem>& c, Elem (&v)[32], int frame) {
  if constexpr (CROSS == 2) {
    const Z2Addr za = z2_addr(*c.P, c.base, frame, c.thr[frame]);
#pragma unroll
    for (int q = 0; q < 16; q++) {
      Elem x, y;
      load16(c.state + z2_pair_index(za, frame, q), x, y);
      const bool sw = z2_swapped<Elem>(za, frame, q);
      v[2 * q] = sw ? y : x;
      v[2 * q + 1] = sw ? x : y;
    }
    return;
  }
  TileAddr ad = tile_addr_off(*c.P, c.base, frame, c.thr[frame]);
#pragma unroll
  for (int q = 0; q < 16; q++)
    load16(eptr<CROSS, Elem>(*c.P, c.state, ad.e0 + reg_off(ad, q)), v[2 * q], v[2 * q + 1]);
}

// Pulls the 1024 128-byte rows of a LATER tile (the one the next CTA scheduled on this SM will most
// likely own) from HBM into L2, so that its loads overlap this tile's arithmetic.
template <typename Elem>
__device__ __forceinline__ void prefetch_later_tile(const Ctx<Elem>& c) {
  const TilePass& P = *c.P;
  if (P.prefetch_dist <= 0 || P.shard_bits != 0 || P.zbit != 0) return;
  const unsigned long long id = c.id + (unsigned)P.prefetch_dist;
  if (id >= c.total) return;
  const unsigned ntile = chunk_tile(*c.P, (unsigned)(id & c.tmask)), npoint = (unsigned)(id >> c.tshift);
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

template <typename Elem, int CROSS>
__device__ __forceinline__ void op_store(const Ctx<Elem>& c, Elem (&v)[32], int frame) {
  if constexpr (CROSS == 2) {
    const Z2Addr za = z2_addr(*c.P, c.base, frame, c.thr[frame]);
#pragma unroll
    for (int q = 0; q < 16; q++) {
      const bool sw = z2_swapped<Elem>(za, frame, q);
      store16(c.state + z2_pair_index(za, frame, q), sw ? v[2 * q + 1] : v[2 * q], sw ? v[2 * q] : v[2 * q + 1]);
    }
    return;
  }
  TileAddr ad = tile_addr_off(*c.P, c.base, frame, c.thr[frame]);
#pragma unroll
  for (int q = 0; q < 16; q++)
    store16(eptr<CROSS, Elem>(*c.P, c.state, ad.e0 + reg_off(ad, q)), v[2 * q], v[2 * q + 1]);
}

// frame A4: every register is one 8-byte element of its own 128-byte row
template <typename Elem, int CROSS>
__device__ __forceinline__ void op_load8(const Ctx<Elem>& c, Elem (&v)[32]) {
  const TilePass& P = *c.P;
  const size_t e0 = c.base + c.thr[3];
  size_t ro[5];
#pragma unroll
  for (int j = 0; j < 5; j++) ro[j] = (size_t)P.roffe8[j];
#pragma unroll
  for (int r = 0; r < 32; r++) {
    size_t e = e0;
#pragma unroll
    for (int j = 0; j < 5; j++) if (r & (1 << j)) e += ro[j];
    v[r] = *eptr<CROSS, Elem>(P, c.state, e);
  }
}

template <typename Elem, int CROSS>
__device__ __forceinline__ void op_store8(const Ctx<Elem>& c, Elem (&v)[32]) {
  const TilePass& P = *c.P;
  const size_t e0 = c.base + c.thr[3];
  size_t ro[5];
#pragma unroll
  for (int j = 0; j < 5; j++) ro[j] = (size_t)P.roffe8[j];
#pragma unroll
  for (int r = 0; r < 32; r++) {
    size_t e = e0;
#pragma unroll
    for (int j = 0; j < 5; j++) if (r & (1 << j)) e += ro[j];
    *eptr<CROSS, Elem>(P, c.state, e) = v[r];
  }
}

template <typename Elem>
__device__ __forceinline__ void op_init(const Ctx<Elem>& c, Elem (&v)[32], int frame) {
  typedef typename ElemTraits<Elem>::real real;
  constexpr int NAMP = ElemTraits<Elem>::NAMP;
  const TilePass& P = *c.P;
  TileAddr ad = tile_addr_off(P, c.base, frame, c.thr[frame]);
  const real A = (real)P.init_amp;
  const int pc0 = __popcll((unsigned long long)(ad.e0 >> P.qshift)) + P.rank_popc;
#pragma unroll
  for (int a = 0; a < NAMP; a++) {
    // tile bits are distinct element bits: popcount(x) = popcount(thread part) + popcount(reg part)
    const int regbits = ElemTraits<Elem>::IS_C64 ? a : (a << 1);
    // (twisted shards, complex64: elements with index bit 0 set owe zdelta more, see "twisted coordinates")
    const int tw = (regbits & 1) ? P.zdelta : 0;
    int pc = pc0 + __popc(regbits >> P.qshift) + tw;
    // symmetric register (INIT runs in frame A): register bit 4 is the virtual bit, the stored index of
    // such an element is the complement of its n-1 real qubit bits
    if (P.zbit != 0 && (regbits & 16)) pc = P.znreal - (pc0 + __popc((regbits & 15) >> P.qshift)) - tw;
    real re = A, im = (real)0;
    if (P.init_kind == INIT_DICKE && pc != P.dicke_k) re = (real)0;
    mul_i_pow(re, im, pc);
    amp_set(v, a, re, im);
  }
}

// phase separator, generic form: the stream is read where it is needed, U8 tables are (re)loaded
// into shared memory on every call
template <typename Elem, int DK>
__device__ __forceinline__ void op_phase(const Ctx<Elem>& c, Elem (&v)[32], int slot) {
  typedef typename ElemTraits<Elem>::real real;
  typedef typename ElemTraits<Elem>::cplx cplx;
  constexpr int NAMP = ElemTraits<Elem>::NAMP;
  const TilePass& P = *c.P;
  const double gamma = __ldg(c.phs + 2 * slot), mult = __ldg(c.phs + 2 * slot + 1);
  if (DK == DIAG_U8 || (DK == 0 && c.dinfo.kind == DIAG_U8)) {
    cplx* tab = reinterpret_cast<cplx*>(c.tab);
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
    const PhaseFx fx = make_phase_fx(gamma, c.dinfo);
    for_each_diag<NAMP>(P.streams[slot], kind, c.tile, c.warp, c.lane, [&](int a, unsigned k, double cd) {
      real s, co;
      if (sizeof(real) == 4 && kind == DIAG_U32FX) {     // complex64 + fixed-point diagonal: integer range reduction
        float sf, cf;
        phase_fx_sincos_fast(fx, k, sf, cf);
        s = (real)sf; co = (real)cf;
      } else {
        double turns = kind == DIAG_U32FX ? fma((double)k, gd, g0) : fma(cd, gd, g0);
        phase_sincos(turns, s, co);
      }
      s *= sg;
      co *= sg;
      real re, im;
      amp_get(v, a, re, im);
      amp_set(v, a, re * co - im * s, re * s + im * co);
    });
  }
}

// phase separator, compiled programs with a U8 diagonal: the thread's stream words were fetched at
// the start of the tile (dq) and the table of the current point is resident in shared memory
template <typename Elem>
__device__ __forceinline__ void op_phase_u8_regs(const Ctx<Elem>& c, Elem (&v)[32], const uint4 (&dq)[2]) {
  typedef typename ElemTraits<Elem>::real real;
  typedef typename ElemTraits<Elem>::cplx cplx;
  constexpr int NAMP = ElemTraits<Elem>::NAMP;
  const cplx* tab = reinterpret_cast<const cplx*>(c.tab);
#pragma unroll
  for (int j = 0; j < NAMP / 16; j++) {
    const unsigned ww[4] = {dq[j].x, dq[j].y, dq[j].z, dq[j].w};
#pragma unroll
    for (int b = 0; b < 16; b++) {
      const unsigned k = (ww[b >> 2] >> (8 * (b & 3))) & 0xffu;
      const cplx ph = tab[k];
      real re, im;
      amp_get(v, j * 16 + b, re, im);
      amp_set(v, j * 16 + b, re * ph.x - im * ph.y, re * ph.y + im * ph.x);
    }
  }
}

// |+> generation fused with the first phase separator (compiled FIRST pass, U8 diagonal):
//   psi~(x) = i^popc(x) * A * m * e^{i gamma c(x)}  =  tab4[popc(x) & 3][k(x)],
// four rotated copies of the phase table sit in the (otherwise unused) staging area, so an amplitude
// costs one byte extract and one shared-memory load.
template <typename Elem, bool Z2>
__device__ __forceinline__ void op_init_phase_plus(const Ctx<Elem>& c, Elem (&v)[32], const uint4 (&dq)[2]) {
  typedef typename ElemTraits<Elem>::real real;
  typedef typename ElemTraits<Elem>::cplx cplx;
  constexpr int NAMP = ElemTraits<Elem>::NAMP;
  const TilePass& P = *c.P;
  const size_t e0 = c.base + c.thr[0];
  const int pc0 = __popcll((unsigned long long)(e0 >> P.qshift)) + P.rank_popc;
  const unsigned char* tab4 = c.smem + TileSmem<Elem>::STAGE_OFF;
  constexpr unsigned TB = 256 * (unsigned)sizeof(cplx);   // bytes per rotated table; tables are addressed by byte offset
  unsigned tb[4];
#pragma unroll
  for (int j = 0; j < 4; j++) tb[j] = ((pc0 + j) & 3) * TB;
  // symmetric register: amplitudes whose top index bit (the virtual tile bit) is set are stored at the
  // complemented index: popc = (n-1) - popc(thread part) - popc(low register part)
  constexpr int VB = NAMP / 2;
  unsigned tm[4];
#pragma unroll
  for (int j = 0; j < 4; j++) tm[j] = Z2 ? ((P.znreal - pc0 - j) & 3) * TB : 0u;
  // twisted shards (complex64): amplitudes with index bit 0 set owe zdelta more (mirrored ones: less)
  constexpr bool TW = Z2 && ElemTraits<Elem>::IS_C64;
  unsigned tbo[4], tmo[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    tbo[j] = TW ? ((pc0 + P.zdelta + j) & 3) * TB : 0u;
    tmo[j] = TW ? ((P.znreal - pc0 - P.zdelta - j) & 3) * TB : 0u;
  }
#pragma unroll
  for (int j = 0; j < NAMP / 16; j++) {
    const unsigned ww[4] = {dq[j].x, dq[j].y, dq[j].z, dq[j].w};
#pragma unroll
    for (int b = 0; b < 16; b++) {
      const int a = j * 16 + b;
      const unsigned k = (ww[b >> 2] >> (8 * (b & 3))) & 0xffu;
      const bool odd = TW && (a & 1);
      const unsigned tsel = (Z2 && (a & VB)) ? (odd ? tmo : tm)[__builtin_popcount(a & (VB - 1)) & 3]
                                             : (odd ? tbo : tb)[__builtin_popcount(a) & 3];
      const cplx ph = *reinterpret_cast<const cplx*>(tab4 + tsel + k * (unsigned)sizeof(cplx));
      amp_set(v, a, (real)ph.x, (real)ph.y);
    }
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
  block_reduce5(r5, c.scratch);
  if (threadIdx.x == 0) {
    double* o = c.partials + ((size_t)c.point * c.ntiles + c.tile) * 5;
#pragma unroll
    for (int i = 0; i < 5; i++) o[i] = r5[i];
  }
}

// Fused <cost>, <cost^2>, min, max reduction of the compiled programs.  Every thread keeps double
// accumulators across ALL tiles its CTA processes for a point (acc = {S0, S1, S2, kmin, kmax}); the
// block-wide tree runs once per point (flush_acc), not once per tile.  Tile-to-CTA assignment is
// static, so the summation order is fixed and results are reproducible run to run.
// U8 diagonal: stream words in registers, integer min/max over whole words (the exact
// "support only" handling runs only when a zero amplitude shows up).
// (float)(byte b of w) without the conversion pipe: build 2^23 + k with one byte permute, subtract 2^23
__device__ __forceinline__ float byte_to_float(unsigned w, int b) {
  return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650 | b)) - 8388608.0f;
}

// Amplitudes enter UNSCALED (the scale still owed by the last layer, `post`, is <= 1 in magnitude, so
// the stored values are the larger ones: no underflow is introduced); the sums are scaled by post^2
// once per tile in double precision.
template <typename Elem>
__device__ __forceinline__ void op_reduce_u8_regs(const Ctx<Elem>& c, Elem (&v)[32], const uint4 (&dq)[2],
                                                  typename ElemTraits<Elem>::real post, double (&acc)[5]) {
  typedef typename ElemTraits<Elem>::real real;
  constexpr int NAMP = ElemTraits<Elem>::NAMP;
  unsigned mn4 = 0xffffffffu, mx4 = 0u;
  double kmin, kmax;
  bool has_zero;
  double t0, t1, t2;
  if constexpr (ElemTraits<Elem>::IS_C64) {
    float2 s0 = make_float2(0.f, 0.f), s1 = s0, s2 = s0;   // two interleaved accumulator lanes (FFMA2)
    float pmin = 1e30f;
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const unsigned ww[4] = {dq[j].x, dq[j].y, dq[j].z, dq[j].w};
#pragma unroll
      for (int w = 0; w < 4; w++) { mn4 = __vminu4(mn4, ww[w]); mx4 = __vmaxu4(mx4, ww[w]); }
#pragma unroll
      for (int b = 0; b < 16; b += 2) {
        const float2 va = v[j * 16 + b], vb = v[j * 16 + b + 1];
        const float2 p2 = make_float2(fmaf(va.x, va.x, va.y * va.y), fmaf(vb.x, vb.x, vb.y * vb.y));
        const float2 k2 = make_float2(byte_to_float(ww[b >> 2], b & 3), byte_to_float(ww[(b + 1) >> 2], (b + 1) & 3));
        const float2 pk = __fmul2_rn(p2, k2);
        s0 = __fadd2_rn(s0, p2);
        s1 = __fadd2_rn(s1, pk);
        s2 = __ffma2_rn(pk, k2, s2);
        pmin = fminf(pmin, fminf(p2.x, p2.y));
      }
    }
    t0 = (double)s0.x + (double)s0.y;
    t1 = (double)s1.x + (double)s1.y;
    t2 = (double)s2.x + (double)s2.y;
    has_zero = !(pmin > 0.f);
  } else {
    real s0 = 0, s1 = 0, s2 = 0, pmin = (real)1e300;
    const unsigned ww[4] = {dq[0].x, dq[0].y, dq[0].z, dq[0].w};
#pragma unroll
    for (int w = 0; w < 4; w++) { mn4 = __vminu4(mn4, ww[w]); mx4 = __vmaxu4(mx4, ww[w]); }
#pragma unroll
    for (int b = 0; b < NAMP; b++) {
      real re, im;
      amp_get(v, b, re, im);
      const real p = re * re + im * im;
      const real kv = (real)((ww[b >> 2] >> (8 * (b & 3))) & 0xffu);
      s0 += p;
      s1 = fma(p, kv, s1);
      s2 = fma(p * kv, kv, s2);
      pmin = p < pmin ? p : pmin;
    }
    t0 = s0; t1 = s1; t2 = s2;
    has_zero = !(pmin > (real)0);
  }
  const unsigned kmn = min(min(mn4 & 0xffu, (mn4 >> 8) & 0xffu), min((mn4 >> 16) & 0xffu, mn4 >> 24));
  const unsigned kmx = max(max(mx4 & 0xffu, (mx4 >> 8) & 0xffu), max((mx4 >> 16) & 0xffu, mx4 >> 24));
  kmin = (double)kmn;
  kmax = (double)kmx;
  if (has_zero) {   // some amplitude is exactly zero: min / max over the support only
    kmin = 1e300; kmax = -1e300;
#pragma unroll
    for (int j = 0; j < NAMP / 16; j++) {
      const unsigned ww[4] = {dq[j].x, dq[j].y, dq[j].z, dq[j].w};
#pragma unroll
      for (int b = 0; b < 16; b++) {
        const unsigned k = (ww[b >> 2] >> (8 * (b & 3))) & 0xffu;
        real re, im;
        amp_get(v, j * 16 + b, re, im);
        if (re * re + im * im > (real)0) { kmin = fmin(kmin, (double)k); kmax = fmax(kmax, (double)k); }
      }
    }
  }
  const double sc = (double)post * (double)post;
  acc[0] += sc * t0;
  acc[1] += sc * t1;
  acc[2] += sc * t2;
  acc[3] = fmin(acc[3], kmin);
  acc[4] = fmax(acc[4], kmax);
}

// same, any diagonal kind, stream read in place
template <typename Elem, int DK>
__device__ __forceinline__ void op_reduce_acc(const Ctx<Elem>& c, Elem (&v)[32], int slot, double (&acc)[5]) {
  typedef typename ElemTraits<Elem>::real real;
  constexpr int NAMP = ElemTraits<Elem>::NAMP;
  const TilePass& P = *c.P;
  const real post = (real)__ldg(c.phs + 2 * P.nphases);
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
  acc[0] += (double)s0;
  acc[1] += (double)s1;
  acc[2] += (double)s2;
  acc[3] = fmin(acc[3], kmin);
  acc[4] = fmax(acc[4], kmax);
}

// block tree over the per-thread accumulators -> partials[point][cta], then reset
template <typename Elem>
__device__ __forceinline__ void flush_acc(const Ctx<Elem>& c, double (&acc)[5], unsigned point) {
  block_reduce5(acc, c.scratch);
  if (threadIdx.x == 0) {
    double* o = c.partials + ((size_t)point * gridDim.x + blockIdx.x + c.P->ch_part0) * 5;
#pragma unroll
    for (int i = 0; i < 5; i++) o[i] = acc[i];
  }
  acc[0] = 0.0; acc[1] = 0.0; acc[2] = 0.0; acc[3] = 1e300; acc[4] = -1e300;
}

// ---- complex128 in frame C ----------------------------------------------------------------------
// There the re/im bit (t0) is a LANE bit: register r of an even lane is the real part, of the odd neighbour
// the imaginary part of the same amplitude; the streams of frame C carry one value per REGISTER (32 per thread,
// both lanes of a pair see the same value).  The phase multiply fetches the partner with one shuffle; the
// statistics need no pairing at all (|psi|^2 c = re^2 c + im^2 c).
// Visits the thread's 32 diagonal values: from the prefetched words (U8, compiled programs) or from the stream.
template <typename F>
__device__ __forceinline__ void c128c_each(const Ctx<double>& c, int slot, int kind, const uint4* dq, F&& f) {
  if (dq) {
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const unsigned ww[4] = {dq[j].x, dq[j].y, dq[j].z, dq[j].w};
#pragma unroll
      for (int b = 0; b < 16; b++) f(j * 16 + b, (ww[b >> 2] >> (8 * (b & 3))) & 0xffu, 0.0);
    }
  } else {
    for_each_diag<32>(c.P->streams[slot], kind, c.tile, c.warp, c.lane, f);
  }
}

// TABLE_READY: the phase table of this slot is resident in shared memory (compiled programs)
template <bool TABLE_READY>
__device__ __forceinline__ void op_phase_c128c(const Ctx<double>& c, double (&v)[32], int slot, const uint4* dq) {
  const TilePass& P = *c.P;
  const int kind = c.dinfo.kind;
  const double sgn = (c.lane & 1) ? 1.0 : -1.0;   // re' = re c - im s ; im' = im c + re s
  if (kind == DIAG_U8) {
    double2* tab = reinterpret_cast<double2*>(c.tab);
    if (!TABLE_READY) {
      __syncthreads();
      if (threadIdx.x < 256) {
        const double2* t = reinterpret_cast<const double2*>(c.tables) +
                           ((size_t)c.point * c.tables_per_point + P.table_id[slot]) * 256;
        tab[threadIdx.x] = t[threadIdx.x];
      }
      __syncthreads();
    }
    c128c_each(c, slot, kind, dq, [&](int r, unsigned k, double) {
      const double2 ph = tab[k];
      const double other = __shfl_xor_sync(0xffffffffu, v[r], 1);
      v[r] = fma(sgn * other, ph.y, v[r] * ph.x);
    });
  } else {
    const double gamma = __ldg(c.phs + 2 * slot), mult = __ldg(c.phs + 2 * slot + 1);
    const double inv2pi = 0.15915494309189533576888;
    const double g0 = gamma * c.dinfo.c0 * inv2pi, gd = gamma * c.dinfo.delta * inv2pi;
    c128c_each(c, slot, kind, nullptr, [&](int r, unsigned k, double cd) {
      const double turns = kind == DIAG_U32FX ? fma((double)k, gd, g0) : fma(cd, gd, g0);
      double sn, co;
      phase_sincos(turns, sn, co);
      const double other = __shfl_xor_sync(0xffffffffu, v[r], 1);
      v[r] = mult * fma(sgn * other, sn, v[r] * co);
    });
  }
}

__device__ __forceinline__ void op_reduce_c128c(const Ctx<double>& c, double (&v)[32], int slot, const uint4* dq,
                                                double post, double (&acc)[5]) {
  const int kind = c.dinfo.kind;
  double s0 = 0, s1 = 0, s2 = 0, kmin = 1e300, kmax = -1e300;
  c128c_each(c, slot, kind, dq, [&](int r, unsigned k, double cd) {
    const double p = v[r] * v[r];
    const double kv = kind == DIAG_F64 ? cd : (double)k;
    s0 += p;
    s1 = fma(p, kv, s1);
    s2 = fma(p * kv, kv, s2);
    if (p > 0.0) { kmin = fmin(kmin, kv); kmax = fmax(kmax, kv); }
  });
  const double sc = post * post;
  acc[0] += sc * s0;
  acc[1] += sc * s1;
  acc[2] += sc * s2;
  acc[3] = fmin(acc[3], kmin);
  acc[4] = fmax(acc[4], kmax);
}

template <typename Elem>
__device__ __forceinline__ void ctx_set_tile(Ctx<Elem>& c, unsigned long long id) {
  const TilePass& P = *c.P;
  c.id = id;
  c.tile = chunk_tile(P, (unsigned)(id & c.tmask));   // ntiles is a power of two
  c.point = (unsigned)(id >> c.tshift);
  c.state = c.state0 + (size_t)c.point * c.point_stride;
  c.tau = c.tau0 + (size_t)c.point * P.tau_stride;
  c.phs = c.phs0 + (size_t)c.point * P.phs_stride;
  c.base = tile_base(P, c.tile + P.tile0);
  if (P.zbit != 0) {   // sigma of (this thread, register 0) in frame A
    typedef typename ElemTraits<Elem>::real real;
    const int par = (__popcll((unsigned long long)((c.base + c.thr[0]) >> P.qshift)) + P.rank_popc) & 1;
    c.zsig = (real)((par ? -1 : 1) * P.zsign);
    c.zsigo = P.ztw ? c.zsig : -c.zsig;
  }
}


template <typename Elem>
__device__ __forceinline__ void ctx_init(Ctx<Elem>& c, const TilePass& P, Elem* state, size_t point_stride,
                                         const typename ElemTraits<Elem>::real* tau, const double* phs,
                                         const void* tables, int tables_per_point, const DiagInfo& dinfo,
                                         double* partials, unsigned char* smem, int tab_off, int scratch_off,
                                         unsigned long long total) {
  c.P = &P;
  c.lane = threadIdx.x & 31;
  c.warp = threadIdx.x >> 5;
  c.ntiles = P.ntiles;
  c.tshift = 31 - __clz((int)P.ntiles);
  c.tmask = P.ntiles - 1u;
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
  c.tab = smem + tab_off;
  c.scratch = reinterpret_cast<double*>(smem + scratch_off);
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
// Pairs < STAGE_PAIRS live in the dedicated staging area ("early" group, issued right after the
// current tile has been pulled into registers); the remaining pairs use the start of the exchange
// buffer ("late" group, issued once the tile's last transpose is done with that buffer).  With
// L2_REST the lines of the pairs >= Q1 are pulled into L2 so that the late group is an L2 hit.
template <typename Elem, int Q0, int Q1, bool L2_REST, int CROSS>
__device__ __forceinline__ void stage_issue(const Ctx<Elem>& c, unsigned long long id, int frame) {
  constexpr int SP = TileSmem<Elem>::STAGE_PAIRS;
  if (id < c.total) {
    const TilePass& P = *c.P;
    const unsigned ntile = chunk_tile(*c.P, (unsigned)(id & c.tmask)), npoint = (unsigned)(id >> c.tshift);
    Elem* pb = c.state0 + (size_t)npoint * c.point_stride;
    const unsigned sb = (unsigned)__cvta_generic_to_shared(c.smem);
    const unsigned early = sb + TileSmem<Elem>::STAGE_OFF + threadIdx.x * 16u;
    // late slots sit inside the OWNING warp's 8 KB region of the exchange buffer; every block-wide
    // transpose starts with a barrier, so they are consumed before anybody else writes there
    const unsigned late = sb + c.warp * 8192u + c.lane * 16u;
    if constexpr (CROSS == 2) {   // symmetric register: mirrored pairs land with their halves swapped (op_load_staged)
      const Z2Addr za = z2_addr(P, tile_base(P, ntile + P.tile0), frame, c.thr[frame]);
#pragma unroll
      for (int q = Q0; q < Q1; q++) {
        const unsigned dst = q < SP ? early + q * (TILE_THREADS * 16) : late + (q - SP) * 512;
        cp_async16(dst, pb + z2_pair_index(za, frame, q));
      }
      if (L2_REST) {
#pragma unroll
        for (int q = Q1; q < 16; q++) asm volatile("prefetch.global.L2 [%0];" ::"l"(pb + z2_pair_index(za, frame, q)));
      }
    } else {
    TileAddr ad = tile_addr_off(P, tile_base(P, ntile + P.tile0), frame, c.thr[frame]);
#pragma unroll
    for (int q = Q0; q < Q1; q++) {
      const unsigned dst = q < SP ? early + q * (TILE_THREADS * 16) : late + (q - SP) * 512;
      cp_async16(dst, eptr<CROSS, Elem>(P, pb, ad.e0 + reg_off(ad, q)));
    }
    if (L2_REST && CROSS != 1) {   // (peer lines bypass the local L2)
#pragma unroll
      for (int q = Q1; q < 16; q++) asm volatile("prefetch.global.L2 [%0];" ::"l"(pb + ad.e0 + reg_off(ad, q)));
    }
    }
  }
  cp_async_commit();
}

__device__ __forceinline__ void cp_async8(unsigned smem_addr, const void* g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_addr), "l"(g) : "memory");
}

// frame A4 variant: register pair q = registers (2q, 2q+1) = two 8-byte elements of different rows,
// staged side by side in the same 16-byte slot the A / B frames use
template <typename Elem, int Q0, int Q1, bool L2_REST, int CROSS>
__device__ __forceinline__ void stage_issue8(const Ctx<Elem>& c, unsigned long long id) {
  constexpr int SP = TileSmem<Elem>::STAGE_PAIRS;
  if (id < c.total) {
    const TilePass& P = *c.P;
    const unsigned ntile = chunk_tile(*c.P, (unsigned)(id & c.tmask)), npoint = (unsigned)(id >> c.tshift);
    const size_t e0 = tile_base(P, ntile + P.tile0) + c.thr[3];
    size_t ro[5];
#pragma unroll
    for (int j = 0; j < 5; j++) ro[j] = (size_t)P.roffe8[j];
    Elem* pb = c.state0 + (size_t)npoint * c.point_stride;
    const unsigned sb = (unsigned)__cvta_generic_to_shared(c.smem);
    const unsigned early = sb + TileSmem<Elem>::STAGE_OFF + threadIdx.x * 16u;
    const unsigned late = sb + c.warp * 8192u + c.lane * 16u;
#pragma unroll
    for (int r = 2 * Q0; r < 2 * Q1; r++) {
      const int q = r >> 1;
      size_t e = e0;
#pragma unroll
      for (int j = 0; j < 5; j++) if (r & (1 << j)) e += ro[j];
      const unsigned dst = (q < SP ? early + q * (TILE_THREADS * 16) : late + (q - SP) * 512) + (r & 1) * 8;
      cp_async8(dst, eptr<CROSS, Elem>(P, pb, e));
    }
    if (L2_REST && CROSS != 1) {
#pragma unroll
      for (int r = 2 * Q1; r < 32; r++) {
        size_t e = e0;
#pragma unroll
        for (int j = 0; j < 5; j++) if (r & (1 << j)) e += ro[j];
        asm volatile("prefetch.global.L2 [%0];" ::"l"(pb + e));
      }
    }
  }
  cp_async_commit();
}

// persistent-kernel load: every register pair comes from the slots filled during the previous tile
template <typename Elem, int CROSS>
__device__ __forceinline__ void op_load_staged(const Ctx<Elem>& c, Elem (&v)[32], int frame) {
  constexpr int SP = TileSmem<Elem>::STAGE_PAIRS;
  cp_async_wait_all();
  const unsigned char* early = c.smem + TileSmem<Elem>::STAGE_OFF + threadIdx.x * 16;
  const unsigned char* late = c.smem + c.warp * 8192 + c.lane * 16;
#pragma unroll
  for (int q = 0; q < 16; q++) {
    const unsigned char* src = q < SP ? early + q * (TILE_THREADS * 16) : late + (q - SP) * 512;
    if constexpr (CROSS == 2 && ElemTraits<Elem>::IS_C64) {
      const bool sw = frame == 0 ? (q & 8) != 0 : (c.warp & 8) != 0;   // see z2_swapped
      Elem x, y;
      load16(reinterpret_cast<const Elem*>(src), x, y);
      v[2 * q] = sw ? y : x;
      v[2 * q + 1] = sw ? x : y;
    } else {
      load16(reinterpret_cast<const Elem*>(src), v[2 * q], v[2 * q + 1]);
    }
  }
  __syncwarp();
  if (frame == 3) stage_issue8<Elem, 0, SP, true, CROSS>(c, c.id + c.stride);
  else stage_issue<Elem, 0, SP, true, CROSS>(c, c.id + c.stride, frame);
}

// ---- interpreter kernel -------------------------------------------------------------------
template <typename Elem, int CROSS>
__global__ void __launch_bounds__(TILE_THREADS, 1)
tile_pass_kernel(const __grid_constant__ TilePass P, Elem* __restrict__ state, size_t point_stride,
                 const typename ElemTraits<Elem>::real* __restrict__ tau, const double* __restrict__ phs,
                 const void* __restrict__ tables, int tables_per_point, DiagInfo dinfo,
                 double* __restrict__ partials) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  typedef typename ElemTraits<Elem>::real real_t;
  Ctx<Elem> c;
  ctx_init<Elem>(c, P, state, point_stride, tau, phs, tables, tables_per_point, dinfo, partials, smem_raw,
                 TILE_EXCH_BYTES, TILE_EXCH_BYTES + 4096, (unsigned long long)gridDim.x);
  Elem v[32];
  TileOp op = P.ops[0];
  int fcur = op.frame;   // frame tracking (symmetric registers: which elements are mirrored)
  bool wt = false;
  for (int io = 0; io < P.nops; io++) {
    const TileOp cur = op;
    if (io + 1 < P.nops) op = P.ops[io + 1];  // fetch the next step early (constant-bank latency)
    switch (cur.type) {
      case TOP_LOAD:
        if (cur.frame == 3) op_load8<Elem, CROSS>(c, v); else op_load<Elem, CROSS>(c, v, cur.frame);
        prefetch_later_tile<Elem>(c);
        break;
      case TOP_STORE:
        if (cur.frame == 3) op_store8<Elem, CROSS>(c, v); else op_store<Elem, CROSS>(c, v, cur.frame);
        break;
      case TOP_INIT: op_init<Elem>(c, v, cur.frame); break;
      case TOP_PHASE:
        if constexpr (!ElemTraits<Elem>::IS_C64) {
          if (cur.frame == 2) { op_phase_c128c<false>(c, v, cur.arg, nullptr); break; }
        }
        op_phase<Elem, 0>(c, v, cur.arg);
        break;
      case TOP_ROUND:
        if constexpr (CROSS == 2) {
          const bool zr = fcur == 0 && !wt;
          const bool mir = (fcur == 0 ? (c.lane & 16) : (c.warp & 8)) != 0;
          round_dynamic<Elem, true>(c, v, cur.arg, zr && cur.frame == 1, zr, (!zr && mir) ? (real_t)-1 : (real_t)1);
        } else if constexpr (CROSS == 1) {
          const int xg = P.xgbits;
          if (xg == 0) round_dynamic<Elem>(c, v, cur.arg);
          else if (xg == 1) round_xtw_dynamic<Elem, 1>(c, v, cur.arg);
          else if (xg == 2) round_xtw_dynamic<Elem, 2>(c, v, cur.arg);
          else round_xtw_dynamic<Elem, 3>(c, v, cur.arg);
        } else {
          round_dynamic<Elem>(c, v, cur.arg);
        }
        break;
      case TOP_WT: warp_transpose<Elem>(c, v); wt = true; break;
      case TOP_BT: block_transpose<Elem>(c, v); fcur ^= 1; wt = false; break;
      case TOP_XT: xt_a_to_c<Elem>(c, v); fcur = 2; wt = false; break;
      case TOP_XTI: xt_c_to_a<Elem>(c, v); fcur = 0; wt = false; break;
      case TOP_REDUCE:
        if constexpr (!ElemTraits<Elem>::IS_C64) {
          if (cur.frame == 2) {   // per-tile partials, like op_reduce
            double acc[5] = {0.0, 0.0, 0.0, 1e300, -1e300};
            op_reduce_c128c(c, v, cur.arg, nullptr, __ldg(c.phs + 2 * P.nphases), acc);
            block_reduce5(acc, c.scratch);
            if (threadIdx.x == 0) {
              double* o = c.partials + ((size_t)c.point * c.ntiles + c.tile) * 5;
              for (int i = 0; i < 5; i++) o[i] = acc[i];
            }
            break;
          }
        }
        op_reduce<Elem, 0>(c, v, cur.arg);
        break;
      default: break;
    }
  }
}

// ---- compile-time programs ----------------------------------------------------------------
struct COp {
  int type, frame, arg, mask;
};

constexpr int M_ALL = 0x1f, M_HI4 = 0x1e, M_TOP2 = 0x18;

// full cover of all 14 tile bits starting in frame F (ends in frame F^1):  R(all) WT R(all) BT R(hi4)
// high cover of the 10 upper tile bits:                                    R(hi4) WT R(top2) BT R(hi4)
// high cover of the 9 upper tile bits through frame C:                     R(hi4) XT R(all)   [XTI R(hi4)]
#define PROG_COMMON                                                                                  \
  static constexpr int find_last(int t0, int t1, int t2, int t3) {                                   \
    int k = -1;                                                                                      \
    for (int i = 0; i < N; i++)                                                                      \
      if (op(i).type == t0 || op(i).type == t1 || op(i).type == t2 || op(i).type == t3) k = i;       \
    return k;                                                                                        \
  }                                                                                                  \
  static constexpr int last_transpose() { return find_last(TOP_BT, TOP_XT, TOP_XTI, TOP_BT); }       \
  /* frame the registers are in right before step I (+8: after a warp transpose) */                                \
  static constexpr int frame_before(int I) {                                                         \
    int f = op(0).frame;                                                                             \
    bool wt = false;                                                                                 \
    for (int i = 0; i < I && i < N; i++) {                                                           \
      const int t = op(i).type;                                                                      \
      if (t == TOP_WT) wt = true;                                                                    \
      else if (t == TOP_BT) { f ^= 1; wt = false; }                                                  \
      else if (t == TOP_XT) f = 2;                                                                   \
      else if (t == TOP_XTI) f = 0;                                                                  \
    }                                                                                                \
    return f + (wt ? 8 : 0);                                                                         \
  }                                                                                                  \
  static constexpr int stream_op() { return find_last(TOP_PHASE, TOP_REDUCE, TOP_PHASE, TOP_PHASE); } \
  static constexpr int n_stream_ops() {                                                              \
    int k = 0;                                                                                       \
    for (int i = 0; i < N; i++) if (op(i).type == TOP_PHASE || op(i).type == TOP_REDUCE) k++;        \
    return k;                                                                                        \
  }

struct ProgFirstInit {  // |+>/Dicke generated in registers, phase 1, full layer-1 cover of the tile
  static constexpr int N = 8;
  static constexpr COp op(int i) {
    constexpr COp o[N] = {{TOP_INIT, 0, 0, 0}, {TOP_PHASE, 0, 0, 0}, {TOP_ROUND, 0, 0, M_ALL}, {TOP_WT, 0, 0, 0},
                          {TOP_ROUND, 0, 1, M_ALL}, {TOP_BT, 0, 0, 0}, {TOP_ROUND, 0, 2, M_HI4}, {TOP_STORE, 1, 0, 0}};
    return o[i];
  }
  PROG_COMMON
};
struct ProgFirstLoad {
  static constexpr int N = 8;
  static constexpr COp op(int i) {
    constexpr COp o[N] = {{TOP_LOAD, 0, 0, 0}, {TOP_PHASE, 0, 0, 0}, {TOP_ROUND, 0, 0, M_ALL}, {TOP_WT, 0, 0, 0},
                          {TOP_ROUND, 0, 1, M_ALL}, {TOP_BT, 0, 0, 0}, {TOP_ROUND, 0, 2, M_HI4}, {TOP_STORE, 1, 0, 0}};
    return o[i];
  }
  PROG_COMMON
};
struct ProgInterior {
  static constexpr int N = 7;
  static constexpr COp op(int i) {
    constexpr COp o[N] = {{TOP_LOAD, 0, 0, 0}, {TOP_ROUND, 0, 0, M_HI4}, {TOP_WT, 0, 0, 0}, {TOP_ROUND, 0, 1, M_TOP2},
                          {TOP_BT, 0, 0, 0}, {TOP_ROUND, 0, 2, M_HI4}, {TOP_STORE, 1, 0, 0}};
    return o[i];
  }
  PROG_COMMON
};
struct ProgBoundary {
  static constexpr int N = 13;
  static constexpr COp op(int i) {
    constexpr COp o[N] = {{TOP_LOAD, 0, 0, 0}, {TOP_ROUND, 0, 0, M_HI4}, {TOP_WT, 0, 0, 0}, {TOP_ROUND, 0, 1, M_TOP2},
                          {TOP_BT, 0, 0, 0}, {TOP_ROUND, 0, 2, M_HI4}, {TOP_PHASE, 1, 0, 0}, {TOP_ROUND, 0, 3, M_ALL},
                          {TOP_WT, 0, 0, 0}, {TOP_ROUND, 0, 4, M_ALL}, {TOP_BT, 0, 0, 0}, {TOP_ROUND, 0, 5, M_HI4},
                          {TOP_STORE, 0, 0, 0}};
    return o[i];
  }
  PROG_COMMON
};
struct ProgFinal {
  static constexpr int N = 7;
  static constexpr COp op(int i) {
    constexpr COp o[N] = {{TOP_LOAD, 0, 0, 0}, {TOP_ROUND, 0, 0, M_HI4}, {TOP_WT, 0, 0, 0}, {TOP_ROUND, 0, 1, M_TOP2},
                          {TOP_BT, 0, 0, 0}, {TOP_ROUND, 0, 2, M_HI4}, {TOP_REDUCE, 1, 0, 0}};
    return o[i];
  }
  PROG_COMMON
};
// ---- balanced schedule (complex64, 5 low bits, >= 3 tile shapes): per layer one MID pass over the
// low tile (14 qubits) and one BOUND2 pass over a high tile (9 qubits of layer d, phase d+1, the
// same 9 qubits of layer d+1), two transposes each; FINAL2 needs a single transpose.
struct ProgMid {
  static constexpr int N = 7;
  static constexpr COp op(int i) {
    constexpr COp o[N] = {{TOP_LOAD, 0, 0, 0}, {TOP_ROUND, 0, 0, M_ALL}, {TOP_WT, 0, 0, 0}, {TOP_ROUND, 0, 1, M_ALL},
                          {TOP_BT, 0, 0, 0}, {TOP_ROUND, 0, 2, M_HI4}, {TOP_STORE, 1, 0, 0}};
    return o[i];
  }
  PROG_COMMON
};
struct ProgBound2 {
  static constexpr int N = 9;
  static constexpr COp op(int i) {
    constexpr COp o[N] = {{TOP_LOAD, 0, 0, 0}, {TOP_ROUND, 0, 0, M_HI4}, {TOP_XT, 0, 0, 0}, {TOP_ROUND, 0, 1, M_ALL},
                          {TOP_PHASE, 2, 0, 0}, {TOP_ROUND, 0, 2, M_ALL}, {TOP_XTI, 0, 0, 0}, {TOP_ROUND, 0, 3, M_HI4},
                          {TOP_STORE, 0, 0, 0}};
    return o[i];
  }
  PROG_COMMON
};
struct ProgFinal2 {
  static constexpr int N = 5;
  static constexpr COp op(int i) {
    constexpr COp o[N] = {{TOP_LOAD, 0, 0, 0}, {TOP_ROUND, 0, 0, M_HI4}, {TOP_XT, 0, 0, 0}, {TOP_ROUND, 0, 1, M_ALL},
                          {TOP_REDUCE, 2, 0, 0}};
    return o[i];
  }
  PROG_COMMON
};

// balanced schedule for tiles with 4 low bits: the 10 high tile bits split 5 + 5 over frames A4 / C4
struct ProgBound2L4 {
  static constexpr int N = 9;
  static constexpr COp op(int i) {
    constexpr COp o[N] = {{TOP_LOAD, 3, 0, 0}, {TOP_ROUND, 0, 0, M_ALL}, {TOP_XT, 0, 0, 0}, {TOP_ROUND, 0, 1, M_ALL},
                          {TOP_PHASE, 4, 0, 0}, {TOP_ROUND, 0, 2, M_ALL}, {TOP_XTI, 0, 0, 0}, {TOP_ROUND, 0, 3, M_ALL},
                          {TOP_STORE, 3, 0, 0}};
    return o[i];
  }
  PROG_COMMON
};
struct ProgFinal2L4 {
  static constexpr int N = 5;
  static constexpr COp op(int i) {
    constexpr COp o[N] = {{TOP_LOAD, 3, 0, 0}, {TOP_ROUND, 0, 0, M_ALL}, {TOP_XT, 0, 0, 0}, {TOP_ROUND, 0, 1, M_ALL},
                          {TOP_REDUCE, 4, 0, 0}};
    return o[i];
  }
  PROG_COMMON
};

// ---- cross shape of a sharded register (low 14 - g element bits + the g rank bits): the only qubits
// left to mix are the rank bits, which are register bits of frame A -- no transposes at all.
struct ProgBoundX {   // rank qubits of layer d, phase d+1, rank qubits of layer d+1
  static constexpr int N = 5;
  static constexpr COp op(int i) {
    constexpr COp o[N] = {{TOP_LOAD, 0, 0, 0}, {TOP_ROUND, 0, 0, M_HI4}, {TOP_PHASE, 0, 0, 0}, {TOP_ROUND, 0, 1, M_HI4},
                          {TOP_STORE, 0, 0, 0}};
    return o[i];
  }
  PROG_COMMON
};
struct ProgFinalX {   // rank qubits of the last layer, fused reduction
  static constexpr int N = 3;
  static constexpr COp op(int i) {
    constexpr COp o[N] = {{TOP_LOAD, 0, 0, 0}, {TOP_ROUND, 0, 0, M_HI4}, {TOP_REDUCE, 0, 0, 0}};
    return o[i];
  }
  PROG_COMMON
};

// Per-point residents of the compiled-program kernels: the round coefficients (and the reduction
// scale, parked in the pad slot 11 of round 0) and, for U8 diagonals, the phase table of the
// program's PHASE step.  Called whenever the CTA moves on to a new point.
template <typename Elem, typename Prog, int DK>
__device__ __forceinline__ void point_setup(const Ctx<Elem>& c) {
  typedef typename ElemTraits<Elem>::real real;
  typedef typename ElemTraits<Elem>::cplx cplx;
  const TilePass& P = *c.P;
  real* tau_s = reinterpret_cast<real*>(c.smem + TileSmem<Elem>::TAU_OFF);
  __syncthreads();
  if ((int)threadIdx.x < P.tau_stride) {
    real val = __ldg(c.tau + threadIdx.x);
    if (threadIdx.x == 11) val = (real)__ldg(c.phs + 2 * P.nphases);
    tau_s[threadIdx.x] = val;
  }
  constexpr int so = Prog::stream_op();
  if constexpr (DK == DIAG_U8 && so >= 0) {
    if constexpr (Prog::op(so >= 0 ? so : 0).type == TOP_PHASE) {
      const cplx* t = reinterpret_cast<const cplx*>(c.tables) +
                      ((size_t)c.point * c.tables_per_point + P.table_id[Prog::op(so).arg]) * 256;
      if (threadIdx.x >= 256) reinterpret_cast<cplx*>(c.tab)[threadIdx.x - 256] = t[threadIdx.x - 256];
      if constexpr (Prog::op(0).type == TOP_INIT) {
        if (P.init_kind == INIT_PLUS) {   // tab4[r][k] = i^r * A * table[k]  (see op_init_phase_plus)
          cplx* tab4 = reinterpret_cast<cplx*>(c.smem + TileSmem<Elem>::STAGE_OFF);
          const real A = (real)P.init_amp;
          for (int i = threadIdx.x; i < 1024; i += TILE_THREADS) {
            const cplx ph = t[i & 255];
            real re = A * ph.x, im = A * ph.y;
            mul_i_pow(re, im, i >> 8);
            cplx o; o.x = re; o.y = im;
            tab4[i] = o;
          }
        }
      }
    }
  }
  __syncthreads();
}

template <typename Elem, typename Prog, int DK, int CROSS, int I>
struct ProgExec {
  static __device__ __forceinline__ void run(const Ctx<Elem>& c, Elem (&v)[32], const uint4 (&dq)[2], double (&acc)[5]) {
    typedef typename ElemTraits<Elem>::real real;
    if constexpr (I < Prog::N) {
      constexpr COp op = Prog::op(I);
      constexpr bool REGS = DK == DIAG_U8 && Prog::n_stream_ops() == 1;   // stream words prefetched in dq
      const real* tau_s = reinterpret_cast<const real*>(c.smem + TileSmem<Elem>::TAU_OFF);
      if constexpr (op.type == TOP_LOAD) {
        op_load_staged<Elem, CROSS>(c, v, op.frame);
        if constexpr (Prog::last_transpose() < 0) {  // the exchange buffer is never used: stage the rest right away
          if constexpr (op.frame == 3) stage_issue8<Elem, TileSmem<Elem>::STAGE_PAIRS, 16, false, CROSS>(c, c.id + c.stride);
          else stage_issue<Elem, TileSmem<Elem>::STAGE_PAIRS, 16, false, CROSS>(c, c.id + c.stride, op.frame);
        }
        prefetch_later_tile<Elem>(c);
      }
      else if constexpr (op.type == TOP_STORE) {
        if constexpr (op.frame == 3) op_store8<Elem, CROSS>(c, v);
        else op_store<Elem, CROSS>(c, v, op.frame);
      }
      else if constexpr (op.type == TOP_INIT) {
        if constexpr (REGS && Prog::op(I + 1 < Prog::N ? I + 1 : I).type == TOP_PHASE) {
          if (c.P->init_kind == INIT_PLUS) op_init_phase_plus<Elem, CROSS == 2>(c, v, dq);   // phase fused in
          else op_init<Elem>(c, v, op.frame);
        } else op_init<Elem>(c, v, op.frame);
      }
      else if constexpr (op.type == TOP_PHASE && !ElemTraits<Elem>::IS_C64 && op.frame == 2) {
        op_phase_c128c<DK == DIAG_U8>(c, v, op.arg, REGS ? dq : nullptr);
      }
      else if constexpr (op.type == TOP_PHASE) {
        if constexpr (REGS) {
          constexpr bool after_init = I > 0 && Prog::op(I > 0 ? I - 1 : 0).type == TOP_INIT;
          if (!(after_init && c.P->init_kind == INIT_PLUS)) op_phase_u8_regs<Elem>(c, v, dq);
        }
        else op_phase<Elem, DK>(c, v, op.arg);
      }
      else if constexpr (op.type == TOP_ROUND) {
        if constexpr (CROSS == 2) {   // low tile of a symmetric register
          constexpr int fb = Prog::frame_before(I);
          constexpr bool ZR = fb == 0;                    // frame A: registers 16..31 are mirrored
          constexpr bool ZS = ZR && (op.mask & 16) != 0;  // ... and slot 4 is the mirror pairing
          real ms = (real)1;
          if constexpr (!ZR) ms = ((fb == 8 ? (c.lane & 16) : (c.warp & 8)) != 0) ? (real)-1 : (real)1;
          round_static<Elem, op.mask, ZS, ZR, !ZR>(tau_s, v, op.arg, c.zsig, ms, c.zsigo);
        } else if constexpr (CROSS == 1) {
          const int xg = c.P->xgbits;   // twisted shards: qubit 0 and sign-twisted rank qubits (uniform branch)
          if (xg == 0) round_static<Elem, op.mask>(tau_s, v, op.arg);
          else if (xg == 1) round_xtw_static<Elem, 1>(tau_s, v, op.arg);
          else if (xg == 2) round_xtw_static<Elem, 2>(tau_s, v, op.arg);
          else round_xtw_static<Elem, 3>(tau_s, v, op.arg);
        } else {
          round_static<Elem, op.mask>(tau_s, v, op.arg);
        }
      }
      else if constexpr (op.type == TOP_WT) warp_transpose<Elem>(c, v);
      else if constexpr (op.type == TOP_BT || op.type == TOP_XT || op.type == TOP_XTI) {
        if constexpr (op.type == TOP_BT) block_transpose<Elem>(c, v);
        else if constexpr (op.type == TOP_XT) xt_a_to_c<Elem>(c, v);
        else xt_c_to_a<Elem>(c, v);
        if constexpr (Prog::op(0).type == TOP_LOAD && I == Prog::last_transpose()) {
          __syncthreads();  // every warp has read its share of the exchange buffer
          if constexpr (Prog::op(0).frame == 3) stage_issue8<Elem, TileSmem<Elem>::STAGE_PAIRS, 16, false, CROSS>(c, c.id + c.stride);
          else stage_issue<Elem, TileSmem<Elem>::STAGE_PAIRS, 16, false, CROSS>(c, c.id + c.stride, Prog::op(0).frame);
        }
      }
      else if constexpr (op.type == TOP_REDUCE && !ElemTraits<Elem>::IS_C64 && op.frame == 2) {
        op_reduce_c128c(c, v, op.arg, REGS ? dq : nullptr, (double)tau_s[11], acc);
      }
      else if constexpr (op.type == TOP_REDUCE) {
        if constexpr (REGS) op_reduce_u8_regs<Elem>(c, v, dq, tau_s[11], acc);
        else op_reduce_acc<Elem, DK>(c, v, op.arg, acc);
      }
      ProgExec<Elem, Prog, DK, CROSS, I + 1>::run(c, v, dq, acc);
    }
  }
};

template <typename Elem, typename Prog, int DK, int CROSS>
__global__ void __launch_bounds__(TILE_THREADS, 1)
tile_prog_kernel(const __grid_constant__ TilePass P, Elem* __restrict__ state, size_t point_stride,
                 const typename ElemTraits<Elem>::real* __restrict__ tau, const double* __restrict__ phs,
                 const void* __restrict__ tables, int tables_per_point, DiagInfo dinfo,
                 double* __restrict__ partials, unsigned n_points) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  typedef TileSmem<Elem> L;
  constexpr int NAMP = ElemTraits<Elem>::NAMP;
  Ctx<Elem> c;
  const unsigned long long total = (unsigned long long)P.ntiles * n_points;
  ctx_init<Elem>(c, P, state, point_stride, tau, phs, tables, tables_per_point, dinfo, partials, smem_raw,
                 L::TABLE_OFF, L::SCRATCH_OFF, total);
  constexpr COp first = Prog::op(0);
  if constexpr (first.type == TOP_LOAD) {
    if constexpr (first.frame == 3) {
      stage_issue8<Elem, 0, L::STAGE_PAIRS, false, CROSS>(c, blockIdx.x);
      stage_issue8<Elem, L::STAGE_PAIRS, 16, false, CROSS>(c, blockIdx.x);
    } else {
      stage_issue<Elem, 0, L::STAGE_PAIRS, false, CROSS>(c, blockIdx.x, first.frame);
      stage_issue<Elem, L::STAGE_PAIRS, 16, false, CROSS>(c, blockIdx.x, first.frame);
    }
  }
  constexpr bool HAS_RED = Prog::op(Prog::N - 1).type == TOP_REDUCE;
  double acc[5] = {0.0, 0.0, 0.0, 1e300, -1e300};
  if constexpr (HAS_RED) {
    // partials[point][cta]: neutral entries for the points this CTA never touches
    for (unsigned pt = threadIdx.x; pt < n_points; pt += TILE_THREADS) {
      double* o = partials + ((size_t)pt * gridDim.x + blockIdx.x + P.ch_part0) * 5;
      o[0] = 0.0; o[1] = 0.0; o[2] = 0.0; o[3] = 1e300; o[4] = -1e300;
    }
  }
  // Passes that generate the state (no LOAD) consume their diagonal words first thing in a tile: stage
  // the NEXT tile's words in shared memory (double-buffered, behind the rotated tables) so that the
  // DRAM latency of the stream is off the critical path as well.
  constexpr bool DQ_SMEM = DK == DIAG_U8 && Prog::n_stream_ops() == 1 && Prog::op(0).type == TOP_INIT;
  constexpr int DQ_OFF = L::STAGE_OFF + 4 * L::TABLE_BYTES;
  auto dq_issue = [&](unsigned long long id, int buf) {
    if (id < total) {
      const unsigned t = chunk_tile(P, (unsigned)(id & c.tmask));
      const uint4* s4 = reinterpret_cast<const uint4*>(P.streams[Prog::op(Prog::stream_op() >= 0 ? Prog::stream_op() : 0).arg]);
      const unsigned sb = (unsigned)__cvta_generic_to_shared(smem_raw) + DQ_OFF + buf * (TILE_THREADS * 32) + threadIdx.x * 16u;
#pragma unroll
      for (int j = 0; j < NAMP / 16; j++)
        cp_async16(sb + j * (TILE_THREADS * 16), s4 + stream_chunk(t, c.warp, c.lane, NAMP / 16, j));
    }
    cp_async_commit();
  };
  if constexpr (DQ_SMEM) dq_issue(blockIdx.x, 0);
  Elem v[32];
  unsigned loaded_point = 0xffffffffu;
  int it = 0;
  for (unsigned long long id = blockIdx.x; id < total; id += gridDim.x, it ^= 1) {
    ctx_set_tile(c, id);
    if (c.point != loaded_point) {
      if constexpr (HAS_RED) {
        if (loaded_point != 0xffffffffu) flush_acc<Elem>(c, acc, loaded_point);
      }
      point_setup<Elem, Prog, DK>(c);
      loaded_point = c.point;
    }
    uint4 dq[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
    constexpr int so = Prog::stream_op();
    if constexpr (DQ_SMEM) {
      cp_async_wait_all();
      const unsigned char* src = smem_raw + DQ_OFF + it * (TILE_THREADS * 32) + threadIdx.x * 16;
#pragma unroll
      for (int j = 0; j < NAMP / 16; j++) dq[j] = *reinterpret_cast<const uint4*>(src + j * (TILE_THREADS * 16));
      dq_issue(id + gridDim.x, it ^ 1);
    } else if constexpr (DK == DIAG_U8 && Prog::n_stream_ops() == 1) {
      // this tile's slice of the permuted diagonal: fetched now, consumed by the PHASE / REDUCE step
      // (frame C streams carry one value per register, also for complex128)
      constexpr int SAMP = Prog::op(so >= 0 ? so : 0).frame == 2 ? 32 : NAMP;
      const uint4* s4 = reinterpret_cast<const uint4*>(P.streams[Prog::op(so >= 0 ? so : 0).arg]);
#pragma unroll
      for (int j = 0; j < SAMP / 16; j++) dq[j] = __ldg(s4 + stream_chunk(c.tile, c.warp, c.lane, SAMP / 16, j));
    }
    if constexpr (DK == 0 && ElemTraits<Elem>::IS_C64) {
      // wide diagonal values (U32FX: 128 bytes per thread and stream) are read where the PHASE / REDUCE step needs
      // them: pull the NEXT tile's slices into L2 now (one 128-byte line per lane; a warp's slice is contiguous), so
      // that those loads cost an L2 hit instead of a DRAM round trip in the middle of the tile
      const unsigned long long nid = id + gridDim.x;
      if (nid < total) {
        const unsigned nt = chunk_tile(P, (unsigned)(nid & c.tmask));
        const int wbytes = NAMP * diag_value_bytes(c.dinfo.kind) * 32;
        for (int i = 0; i < P.nops; i++) {
          if (P.ops[i].type != TOP_PHASE && P.ops[i].type != TOP_REDUCE) continue;
          const char* base = reinterpret_cast<const char*>(P.streams[P.ops[i].arg]) + ((size_t)nt * 16 + c.warp) * wbytes;
          for (int off = c.lane * 128; off < wbytes; off += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
        }
      }
    }
    ProgExec<Elem, Prog, DK, CROSS, 0>::run(c, v, dq, acc);
  }
  if constexpr (HAS_RED) {
    if (loaded_point != 0xffffffffu) flush_acc<Elem>(c, acc, loaded_point);
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
// using exactly the addressing of the pass kernels (same helper functions).  namp == 32: amplitude a is
// register a (complex64, every frame); namp == 16: amplitude a is the register pair (2a, 2a+1)
// (complex128, frames A and B).
template <int VALBYTES>
__global__ void make_stream_kernel(const __grid_constant__ TilePass P, int frame, int namp, int cross,
                                   const unsigned char* __restrict__ diag,
                                   unsigned char* __restrict__ stream) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned tile = blockIdx.x;
  const size_t e0 = tile_base(P, tile + P.tile0) + thread_off(P, frame, warp, lane);
  constexpr int V = 16 / VALBYTES;
  const int nchunks = namp / V;
  for (int j = 0; j < nchunks; j++) {
    __align__(16) unsigned char buf[16];
    for (int i = 0; i < V; i++) {
      const int a = j * V + i;
      const size_t e = z2_fold(e0 + reg_elem_off(P, frame, namp == 32 ? a : 2 * a), P.zbit, P.qshift);
      const size_t x = e >> P.qshift;   // amplitude index (global in a cross pass)
      const int abits = P.shard_bits - P.qshift;
      const unsigned char* src = cross
          ? reinterpret_cast<const unsigned char*>(P.dpeers[x >> abits]) + (x & (((size_t)1 << abits) - 1)) * VALBYTES
          : diag + x * VALBYTES;
#pragma unroll
      for (int b = 0; b < VALBYTES; b++) buf[i * VALBYTES + b] = src[b];
    }
    reinterpret_cast<uint4*>(stream)[stream_chunk(tile, warp, lane, nchunks, j)] = *reinterpret_cast<uint4*>(buf);
  }
}
