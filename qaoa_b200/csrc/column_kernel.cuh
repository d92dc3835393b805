// Column kernels: the high-tile passes of the balanced schedule (BOUND2, FINAL2) with DECOUPLED warps.
//
// STATUS (round 2): OFF BY DEFAULT (engine option "column_kernels").  Parity: tests/test_gpu_zzz_column_kernels.py
// test_column_kernels_match_the_tile_kernels (energies equal to the tile kernels' to 1e-9, repeated, and to the C oracle).
// Speed on a B200 (n = 32, p = 5, symmetric register, bench.py --column 1 / 0 on the same box): BOUND2 7.25 ms against
// 7.33 ms for tile_prog_kernel<ProgBound2>, FINAL2 3.89 against 3.98 ms, but the tile passes that FOLLOW a column kernel
// run 2 % slower (MID 6.00 against 5.86 ms), so the evaluation is 68.5-68.7 ms either way -- parity, not a gain: both
// designs keep the shared-memory pipe ~70 % busy (profiles/r03i_column_kernels_n28.md); with 227 KB of shared memory there
// are three 64 KB stages for two groups, so a buffer's TMA round trip (store, wait for the read, reload) has half an item.
// Lesson kept in the code: the first version produced deviations of 1e-7 ... 1e-6 or a launch failure in about one run in
// three at 2^28 amplitudes and above.  Cause: the state is written through the GENERIC proxy by the kernel that ran
// before and read here through the ASYNC proxy (TMA), and vice versa for the kernel that follows -- a kernel boundary
// does not order the two proxies; fence.proxy.async before the first TMA load and after the last TMA store does.
//
// Why.  In a high tile (5 low element bits + 9 high bits) the low bits are spectators: the tile is 32 independent
// columns of 2^9 amplitudes.  tile_prog_kernel gives every thread 32 amplitudes of 16 different rows and moves the
// 9 high qubits through the registers with two BLOCK-wide transposes, i.e. four barriers per tile during which the
// FMA pipe, the shared-memory pipe and the memory system take turns (BOUND2 = 8.5 us per tile against
// 6.2 us of HBM time).  Here a WARP owns whole columns -- 2 x 512 amplitudes (element bit 0 and all 9 high bits) in
// its 32 x 32 registers -- so the only exchanges left are lane <-> register transposes INSIDE the warp, and the sixteen
// warps of an SM drift apart: one warp's butterflies run under another warp's transposes and a third one's loads.
//
// Data movement is the Blackwell way: a half tile (512 rows x 128 bytes) is ONE 5-D tensor-map box pair brought in by
// the TMA unit (cp.async.bulk.tensor, SWIZZLE_128B, completion on an mbarrier) and written back by it; no thread computes
// a global address.  Three 64 KB stage buffers rotate over the half tiles of a CTA; two groups of eight warps take the
// half tiles in turn, so that one group computes while the other group's buffer is in flight.  The store of a half tile
// is issued by an elected thread of its group, which then waits for the TMA to have READ the buffer and refills it with
// the half tile three items ahead -- the only cross-group signal is the buffer's "full" mbarrier (one per buffer and
// consuming group: a parity wait is only safe for a waiter that has seen every earlier phase of its barrier).
//
// Frames (t0..t13 = tile bits as in tile_kernel.cuh; rows r = t5..t13, chunk c = t1..t3, half h = t4, e = t0):
//   FA: reg {e, r5..r8 = t10..t13}   lane {r0..r4 = t5..t9}       -- 16-byte accesses, the registers of frame A
//   FB: reg {r0..r4 = t5..t9}        lane {e, r5..r8 = t10..t13}  --  8-byte accesses, the registers of frame C
// so the round coefficients, phase tables and frame-C diagonal streams of the existing BOUND2 / FINAL2 plans are used
// unchanged.  The transposes go through the warp's own column of the stage buffer, in place: the TMA swizzle
// (16-byte chunk index XOR row & 7) makes the FA accesses conflict free; while the column is in transit its row r sits
// in the slot of row rho(r) = r ^ ((r >> 3 ^ r >> 6) & 7), which makes the FB accesses conflict free as well.
#pragma once
#include <cuda.h>

#include "tile_kernel.cuh"

constexpr int COL_STAGE_BYTES = 65536;   // 512 rows x 128 bytes
constexpr int COL_NSTAGE = 3;
constexpr int COL_TAB_OFF = COL_NSTAGE * COL_STAGE_BYTES;             // 256-entry phase table (float2)
constexpr int COL_TAU_OFF = COL_TAB_OFF + 2048;                       // round coefficients (floats)
constexpr int COL_SCRATCH_OFF = COL_TAU_OFF + TILE_CANON_ROUNDS * TILE_ROUND_REALS * 4;
constexpr int COL_BAR_OFF = COL_SCRATCH_OFF + 16 * 5 * 8;
constexpr int COL_SMEM_BYTES = COL_BAR_OFF + 64;

struct ColGeom {
  int mid_bits;              // free element bits between bit 5 and the first high bit (tile index, low run)
  unsigned long long tiles;  // tiles of this launch
  unsigned tile0;            // first tile (chunked launches use the generic remap instead; here 0)
};

__device__ __forceinline__ unsigned col_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void col_mbar_init(unsigned bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void col_mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void col_mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "COL_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra COL_DONE;\n"
      "bra COL_WAIT;\n"
      "COL_DONE:\n"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void col_tma_load(unsigned dst, const CUtensorMap* map, unsigned bar, int c0, int c1, int c2, int c3,
                                             int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void col_tma_store(const CUtensorMap* map, unsigned src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(map), "r"(src),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void col_group_sync(int g) { asm volatile("bar.sync %0, 256;" ::"r"(1 + g) : "memory"); }

// byte offset (inside a stage buffer) of the 16-byte chunk c of the slot of row r, TMA SWIZZLE_128B layout
__device__ __forceinline__ unsigned col_slot(unsigned r, unsigned c) { return r * 128u + ((c ^ (r & 7u)) << 4); }
__device__ __forceinline__ unsigned col_rho(unsigned r) { return r ^ (((r >> 3) ^ (r >> 6)) & 7u); }

// The slot of the in-transit row rho(r) splits into a per-thread part and a compile-time part joined by ONE xor:
//   FA (r = q << 5 | lane):  rho(r) & 31 = lane ^ fl ^ fq,  fl = (lane >> 3) & 7,  fq = ((q << 2) ^ (q >> 1)) & 7
//       slot = (q << 12) + (baseA ^ KA(q)),  baseA = (lane ^ fl) << 7 | (c ^ ((lane ^ fl) & 7)) << 4,  KA = fq << 7 | fq << 4
//   FB (r = hi << 5 | i):    rho(r) & 31 = Ci ^ fh,  Ci = i ^ (i >> 3),  fh = ((hi << 2) ^ (hi >> 1)) & 7
//       slot = baseB ^ KB(i),  baseB = hi << 12 | fh << 7 | ((c ^ fh) & 7) << 4 | e << 3,  KB = Ci << 7 | (Ci & 7) << 4
struct ColAddr {
  unsigned baseA, baseB;
};
__device__ __forceinline__ ColAddr col_addr(unsigned c, int lane) {
  ColAddr a;
  const unsigned l = (unsigned)lane, fl = (l >> 3) & 7u, lx = l ^ fl;
  a.baseA = (lx << 7) | ((c ^ (lx & 7u)) << 4);
  const unsigned e = l & 1u, hi = l >> 1, fh = ((hi << 2) ^ (hi >> 1)) & 7u;
  a.baseB = (hi << 12) | (fh << 7) | (((c ^ fh) & 7u) << 4) | (e << 3);
  return a;
}
__device__ __forceinline__ constexpr unsigned col_ka(int q) { return ((((unsigned)q << 2) ^ ((unsigned)q >> 1)) & 7u) * 0x90u; }
__device__ __forceinline__ constexpr unsigned col_kb(int i) {
  return ((((unsigned)i ^ ((unsigned)i >> 3)) & 31u) << 7) | ((((unsigned)i ^ ((unsigned)i >> 3)) & 7u) << 4);
}

// FA -> FB: registers {e, r5..r8} / lanes {r0..r4}  ->  registers {r0..r4} / lanes {e, r5..r8}
__device__ __forceinline__ void col_fa_to_fb(unsigned char* buf, const ColAddr& A, float2 (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 16; q++)
    *reinterpret_cast<float4*>(buf + (q << 12) + (A.baseA ^ col_ka(q))) = make_float4(v[2 * q].x, v[2 * q].y, v[2 * q + 1].x, v[2 * q + 1].y);
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 32; i++) v[i] = *reinterpret_cast<const float2*>(buf + (A.baseB ^ col_kb(i)));
  __syncwarp();
}
// FB -> FA
__device__ __forceinline__ void col_fb_to_fa(unsigned char* buf, const ColAddr& A, float2 (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 32; i++) *reinterpret_cast<float2*>(buf + (A.baseB ^ col_kb(i))) = v[i];
  __syncwarp();
#pragma unroll
  for (int q = 0; q < 16; q++) {
    const float4 w = *reinterpret_cast<const float4*>(buf + (q << 12) + (A.baseA ^ col_ka(q)));
    v[2 * q] = make_float2(w.x, w.y);
    v[2 * q + 1] = make_float2(w.z, w.w);
  }
  __syncwarp();
}

// Prog: ProgBound2 (LOAD R XT R PHASE R XTI R STORE) or ProgFinal2 (LOAD R XT R REDUCE); complex64, U8 diagonal, one point.
template <typename Prog>
__global__ void __launch_bounds__(TILE_THREADS, 1)
column_prog_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ TilePass P, ColGeom G,
                   const float* __restrict__ tau, const double* __restrict__ phs, const void* __restrict__ tables,
                   int tables_per_point, double* __restrict__ partials) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  constexpr bool HAS_RED = Prog::op(Prog::N - 1).type == TOP_REDUCE;
  constexpr int so = Prog::stream_op();
  const int tid = threadIdx.x, lane = tid & 31, g = tid >> 8, wi = (tid >> 5) & 7;
  const unsigned c = (unsigned)wi;                 // chunk of the half row: element bits 1..3
  float* tau_s = reinterpret_cast<float*>(smem_raw + COL_TAU_OFF);
  const unsigned bar0 = col_smem_u32(smem_raw + COL_BAR_OFF);
  // residents of the (single) point: round coefficients, reduction scale in pad slot 11, phase table
  if (tid < P.tau_stride) {
    float val = __ldg(tau + tid);
    if (tid == 11) val = (float)__ldg(phs + 2 * P.nphases);
    tau_s[tid] = val;
  }
  if constexpr (so >= 0) {
    if constexpr (Prog::op(so >= 0 ? so : 0).type == TOP_PHASE) {
      const float2* t = reinterpret_cast<const float2*>(tables) + (size_t)P.table_id[Prog::op(so).arg] * 256;
      if (tid >= 256) reinterpret_cast<float2*>(smem_raw + COL_TAB_OFF)[tid - 256] = t[tid - 256];
    }
  }
  if (tid == 0) {
    for (int b = 0; b < 2 * COL_NSTAGE; b++) col_mbar_init(bar0 + 8 * b, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  Ctx<float2> cx;
  cx.P = &P;
  cx.lane = lane;
  cx.warp = tid >> 5;
  cx.tab = smem_raw + COL_TAB_OFF;
  cx.scratch = reinterpret_cast<double*>(smem_raw + COL_SCRATCH_OFF);
  cx.partials = partials;
  cx.point = 0;
  cx.smem = smem_raw;

  const unsigned long long my_tiles = G.tiles > blockIdx.x ? (G.tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const unsigned long long n_items = 2 * my_tiles;
  const unsigned mid_mask = (1u << G.mid_bits) - 1u;
  auto tile_of = [&](unsigned long long J) { return chunk_tile(P, (unsigned)(blockIdx.x + (J >> 1) * gridDim.x)); };
  auto issue_load = [&](unsigned long long J) {   // one thread
    if (J >= n_items) return;
    const unsigned t = tile_of(J), b = (unsigned)(J % COL_NSTAGE);
    const unsigned dst = col_smem_u32(smem_raw) + b * COL_STAGE_BYTES;
    // one "full" mbarrier per (buffer, consuming group): a buffer serves the two groups alternately, and a parity wait
    // is only safe for a waiter that has seen every earlier phase of its barrier
    const unsigned bar = bar0 + 8 * (2 * b + (unsigned)(J & 1));
    col_mbar_expect_tx(bar, COL_STAGE_BYTES);
    col_tma_load(dst, &tmap, bar, 0, (int)(J & 1), (int)(t & mid_mask), 0, (int)(t >> G.mid_bits));
    col_tma_load(dst + 32768u, &tmap, bar, 0, (int)(J & 1), (int)(t & mid_mask), 256, (int)(t >> G.mid_bits));
  };
  if (tid == 0) {
    // the state was written through the generic proxy (by the kernel before this one); the TMA reads it through the async proxy
    asm volatile("fence.proxy.async;" ::: "memory");
    for (int J = 0; J < COL_NSTAGE; J++) issue_load(J);
  }

  double acc[5] = {0.0, 0.0, 0.0, 1e300, -1e300};
  if constexpr (HAS_RED) {
    if (tid == 0) {
      double* o = partials + ((size_t)blockIdx.x + P.ch_part0) * 5;
      o[0] = 0.0; o[1] = 0.0; o[2] = 0.0; o[3] = 1e300; o[4] = -1e300;
    }
  }
  float2 v[32];
  const ColAddr CA = col_addr(c, lane);
  const unsigned io_off = col_slot((unsigned)lane, c);     // slot of row (q << 5 | lane) = (q << 12) + io_off
  for (unsigned long long J = g; J < n_items; J += 2) {
    const unsigned b = (unsigned)(J % COL_NSTAGE);
    unsigned char* buf = smem_raw + b * COL_STAGE_BYTES;
    const unsigned t = tile_of(J);
    // the tile's slice of the frame-C diagonal stream: the 32 bytes of the OLD kernel's thread (warp = t10..t13, lane =
    // {e, t1..t3, t4}) are exactly this thread's 32 registers in frame FB
    uint4 dq[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
    const uint4* dqp = nullptr;
    if constexpr (so >= 0) {
      const uint4* s4 = reinterpret_cast<const uint4*>(P.streams[Prog::op(so).arg]);
      const int ow = lane >> 1, ol = (lane & 1) | (wi << 1) | ((int)(J & 1) << 4);
      dqp = s4 + stream_chunk(t, ow, ol, 2, 0);      // chunk 1 sits 32 uint4 further on
      // pulled into L2 now, loaded right before the step that needs it (two registers less across the rounds)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(dqp));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(dqp + 32));
    }
    col_mbar_wait(bar0 + 8 * (2 * b + (unsigned)g), (unsigned)((J / (2 * COL_NSTAGE)) & 1));
    // LOAD (frame FA)
#pragma unroll
    for (int q = 0; q < 16; q++) {
      const float4 w = *reinterpret_cast<const float4*>(buf + (q << 12) + io_off);
      v[2 * q] = make_float2(w.x, w.y);
      v[2 * q + 1] = make_float2(w.z, w.w);
    }
    __syncwarp();
    // the program (its frames A / C are FA / FB here)
    constexpr int N = Prog::N;
#pragma unroll
    for (int I = 1; I < N; I++) {
      // (compile-time dispatch: Prog::op is constexpr and the loop is fully unrolled)
      const COp op = Prog::op(I);
      if (op.type == TOP_ROUND) {
        if (op.mask == M_ALL) round_static<float2, M_ALL>(tau_s, v, op.arg);
        else round_static<float2, M_HI4>(tau_s, v, op.arg);
      } else if (op.type == TOP_XT) {
        col_fa_to_fb(buf, CA, v);
        if constexpr (HAS_RED) {
          // a reducing program is done with the buffer here: refill it now, under the rest of the item's arithmetic
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          col_group_sync(g);
          if ((tid & 255) == 0) issue_load(J + COL_NSTAGE);
        }
      }
      else if (op.type == TOP_XTI) col_fb_to_fa(buf, CA, v);
      else if (op.type == TOP_PHASE) {
        dq[0] = __ldg(dqp); dq[1] = __ldg(dqp + 32);
        op_phase_u8_regs<float2>(cx, v, dq);
      } else if (op.type == TOP_REDUCE) {
        dq[0] = __ldg(dqp); dq[1] = __ldg(dqp + 32);
        op_reduce_u8_regs<float2>(cx, v, dq, tau_s[11], acc);
      }
      else if (op.type == TOP_STORE) {
#pragma unroll
        for (int q = 0; q < 16; q++)
          *reinterpret_cast<float4*>(buf + (q << 12) + io_off) = make_float4(v[2 * q].x, v[2 * q].y, v[2 * q + 1].x, v[2 * q + 1].y);
      }
    }
    // hand the buffer on: store it, then refill it with the half tile three items ahead
    if constexpr (!HAS_RED) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes before the TMA reads the buffer
      col_group_sync(g);
      if ((tid & 255) == 0) {
        const unsigned src = col_smem_u32(buf);
        col_tma_store(&tmap, src, 0, (int)(J & 1), (int)(t & mid_mask), 0, (int)(t >> G.mid_bits));
        col_tma_store(&tmap, src + 32768u, 0, (int)(J & 1), (int)(t & mid_mask), 256, (int)(t >> G.mid_bits));
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        issue_load(J + COL_NSTAGE);
      }
    }
  }
  if constexpr (!HAS_RED) {
    if ((tid & 255) == 0) {
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      asm volatile("fence.proxy.async;" ::: "memory");     // ... and the next kernel reads it through the generic proxy
      __threadfence();
    }
  }
  if constexpr (HAS_RED) {
    __syncthreads();
    flush_acc<float2>(cx, acc, 0);
  }
}
