// Host side of libqaoa_b200.so: pass planner, parameter lowering and the C ABI (include/qaoa_b200.h).
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "../../include/qaoa_b200.h"
#include "common.cuh"
#include "kernels_misc.cuh"
#include "column_kernel.cuh"
#include "tile_kernel.cuh"

namespace {

std::string g_global_error;

#define CUDA_OK(call)                                                                       \
  do {                                                                                      \
    cudaError_t _e = (call);                                                                \
    if (_e != cudaSuccess) {                                                                \
      char _b[512];                                                                         \
      snprintf(_b, sizeof(_b), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e),      \
               __FILE__, __LINE__);                                                         \
      throw std::string(_b);                                                                \
    }                                                                                       \
  } while (0)

// scratch device allocation of a set-up call: released on every exit path (a failed CUDA call throws)
struct DevTmp {
  void* p = nullptr;
  DevTmp() {}
  explicit DevTmp(size_t bytes) { alloc(bytes); }
  void alloc(size_t bytes) {
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 16);
    if (e != cudaSuccess) { p = nullptr; throw std::string("cudaMalloc of a temporary buffer failed: ") + cudaGetErrorString(e); }
  }
  ~DevTmp() { if (p) cudaFree(p); }
  DevTmp(const DevTmp&) = delete;
  DevTmp& operator=(const DevTmp&) = delete;
  template <typename T> T* as() const { return (T*)p; }
};
[[noreturn]] void fail(const std::string& m) { throw std::string(m); }

// ------------------------------------------------------------------------------------------
// tile plan
// ------------------------------------------------------------------------------------------
struct RoundSpec {
  int ebit[5];   // element bit handled by each register slot, -1 = slot idle
  int layer;     // 0-based layer index
};
struct PhaseSpec {
  int layer;     // 0-based layer whose gamma is applied (sigma of layer-1 is folded in)
  int shape;
  int frame;
};
enum ProgId { PROG_GENERIC = 0, PROG_FIRST_INIT, PROG_FIRST_LOAD, PROG_INTERIOR, PROG_BOUNDARY, PROG_FINAL,
              PROG_MID, PROG_BOUND2, PROG_FINAL2, PROG_BOUNDX, PROG_FINALX, PROG_BOUND2L4, PROG_FINAL2L4, PROG_COUNT };
struct PlanPass {
  TilePass tp;
  std::vector<RoundSpec> rounds;
  std::vector<PhaseSpec> phases;
  bool has_reduce = false;
  int reduce_shape = 0, reduce_frame = 0;
  size_t tau_offset = 0;  // reals, per point block
  size_t phs_offset = 0;  // doubles, per point block
  int prog = PROG_GENERIC;
  bool cross = false;
  bool zshape = false;    // pass over the low tile shape of a symmetric register (virtual tile bit 13)
  int shape_index = 0;    // Plan::shapes entry this pass tiles over
  bool chunk_closed = false;   // the pass never couples two chunks of the chunk pipeline (plan_chunks)
};
struct Shape {
  int pos[TILE_BITS];
  std::vector<int> freepos;
  bool cross = false;   // contains element bits owned by other ranks (sharded registers)
  bool zshape = false;  // low tile of a symmetric register: pos[13] is the virtual top qubit
};
struct Plan {
  int p = 0;
  std::vector<Shape> shapes;
  std::vector<PlanPass> passes;
  size_t tau_per_point = 0;  // reals over all passes
  size_t phs_per_point = 0;  // doubles over all passes
  bool z2 = false;           // symmetric register: 2^(n-1) stored amplitudes (tile_kernel.cuh)
  int kx = 0;                // sharded registers: local bits carried by the cross tile (build_plan)
  // chunk pipeline (sharded registers, plan_chunks): 2^ch_m chunks, chunk bits u / v_i (element bits), the one local
  // shape that mixes them (every pass over another shape is chunk-closed)
  int ch_m = 0, ch_u = -1, ch_v[3] = {-1, -1, -1}, ch_open_shape = -1;
};

struct Engine {
  int n = 0, dtype = 0, device = 0;
  int qshift = 0, E = 0;     // element bits of the WHOLE register
  size_t N = 0;              // basis states held by THIS engine (2^n, or 2^(n - gbits) when sharded)
  // sharding: the register is split on its top gbits qubits over world = 2^gbits engines
  int rank = 0, world = 1, gbits = 0;
  int El = 0;                // element bits of the local shard (E - gbits)
  size_t x0 = 0;             // first basis index of the local shard
  void* peer_state[QAOA_MAX_RANKS] = {nullptr};
  void* peer_diag[QAOA_MAX_RANKS] = {nullptr};
  bool peer_state_ipc[QAOA_MAX_RANKS] = {false}, peer_diag_ipc[QAOA_MAX_RANKS] = {false};
  bool peers_state_set = false, peers_diag_set = false;
  size_t amp_bytes = 8;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;      // side stream of the chunk pipeline (cross passes)
  // overrides of the launch in flight (tile_launch_pass_chunk): stream, persistent CTAs, chunk
  cudaStream_t cur_stream = nullptr;
  int cur_ctas = 0, cur_ch_c = -1;
  int column_kernels = 0;              // option "column_kernels": BOUND2 / FINAL2 on the warp-decoupled TMA kernels
  int column_sync = 0;                 // debug: 1 = stream sync before a column kernel, 2 = after, 3 = both
  std::map<std::pair<const void*, int>, CUtensorMap> tmaps;   // (state pointer, first high bit) -> tensor map
  int chunking = 1;                    // option "chunk_pipeline"
  int chunk_bits = 3;                  // option "chunk_bits": at most 2^chunk_bits chunks

  // state storage: `state_points` consecutive registers
  void* state = nullptr;
  size_t state_points = 0;

  // cost
  DiagInfo dinfo{DIAG_NONE, 0.0, 1.0};
  void* diag = nullptr;
  std::map<std::pair<int, int>, void*> streams;  // (shape, frame) -> permuted diagonal

  // per-term gammas (MaxCutFree / MaxCutOrbit): n_gamma term diagonals, single-CTA path only
  int n_gamma = 1;
  double* d_terms = nullptr;
  // the same terms as an edge list (layered evaluation of registers beyond one CTA, run_layered_point)
  std::vector<int> term_u, term_v, term_id;
  std::vector<double> term_w;
  int term_bits = 1, term_n_free = 0, term_fixed_colour = 0;
  unsigned char term_lut[8] = {0};
  int *d_term_u = nullptr, *d_term_v = nullptr;
  double* d_layer_w = nullptr;      // g_e of the layer being prepared
  double* d_init_phi = nullptr;     // PlusParameterized phases of the point being evaluated
  void* layer_vec = nullptr;        // the segment's initial state (N amplitudes)
  // feasible-subspace compression (Dicke + Grover over the same Dicke state): compact cost diagonal of the
  // C(n,k) weight-k strings, built lazily; compact_k < 0: not built
  void* diag_compact = nullptr;
  void* strings_compact = nullptr;          // XY in the feasible subspace: the weight-k strings (u32 for n <= 32, else u64)
  unsigned long long* binom_compact = nullptr;   // C(b, j), (n+1) x (k+1), on the device
  size_t n_compact = 0;
  int compact_k = -1;
  int compact_enabled = 1;
  int n_init = 0;            // PlusParameterized: leading phase parameters of the angle vector (single-CTA path only)

  // mixer / init
  int mixer = MIXER_X;
  std::vector<int> pairs;
  int* d_pairs = nullptr;
  int trotter = 1;
  StateGen init{INIT_PLUS, 0, 1.0, nullptr};
  StateGen sub{INIT_PLUS, 0, 1.0, nullptr};
  void* init_vec = nullptr;
  void* sub_vec = nullptr;
  NodeVec node{0, {0}, {0}};   // MIXER_NODE_GROVER

  // plans
  std::map<int, Plan> plans;

  // options
  int small_max_qubits = -1;  // -1: default by dtype
  int keep_state = 0;
  int batch_points = 0;       // 0: auto
  size_t batch_bytes = (size_t)16 << 30;

  // scratch
  double* d_params = nullptr; size_t params_cap = 0;   // [tau block (real)][phs block (double)]
  double* h_params = nullptr; size_t hparams_cap = 0;
  // second pinned parameter block + one event per block: the host lowers the angles of chunk i+1 while chunk i runs
  double* h_params2 = nullptr; size_t hparams2_cap = 0;
  cudaEvent_t ev_params[2] = {nullptr, nullptr};
  int param_slot = 0;
  int force_generic = 0;
  int layer_budget = 1;       // multi-angle mixers: normal form whenever the LAYER's total growth fits (fill_point_params)
  int prefetch_dist = 0;     // extra L2 prefetch distance in tiles (0 = off)
  int persistent_ctas = 148; // CTAs of the persistent pass kernels (one per SM)
  int low_bits = 0;          // 0: automatic choice between 4 and 5 low tile bits
  int schedule = 0;          // 0: automatic, 1: classic (interior + boundary), 2: balanced (mid + bound2)
  int z2_enabled = 1;        // option "symmetric": store half the register when the problem has the global bit-flip symmetry
  int cost_symmetric = -1;   // cost(x) == cost(~x) for every x: -1 unknown, 0 no, 1 yes (checked on the device)
  // Sharded symmetric register (option "shard_symmetric", tile_kernel.cuh "twisted coordinates"): the engine holds
  // 2^(n-1-gbits) STORED amplitudes; N, El and x0 describe that stored shard, E still counts the virtual top qubit.
  int ztw_mode = 0;
  int cross_spare = 1;       // option "cross_spare": let the cross tile carry spare local bits (build_plan)
  size_t state_bytes = 0;
  double* d_partials = nullptr; size_t partials_cap = 0;
  double* d_out = nullptr; size_t out_cap = 0;
  double* h_out = nullptr; size_t hout_cap = 0;
  void* d_tables = nullptr; size_t tables_cap = 0;
  double* d_gamsig = nullptr; size_t gamsig_cap = 0;
  double* d_coefdot = nullptr;
  double* d_angles = nullptr; size_t angles_cap = 0;

  // bookkeeping of the last evaluation (for read_statevector)
  std::vector<int> last_z;
  int last_gsign = 1;
  double last_scale = 1.0;
  bool last_valid = false;
  int last_path = 0;  // 0 small/elementwise (plain S frame), 1 tile

  // context of the evaluation in flight (pass-by-pass execution)
  struct EvalCtx {
    Plan* pl = nullptr;
    int Bc = 0;
    size_t tau_bytes = 0;
    std::vector<char> all_normal;
    std::vector<size_t> tbase, pbase;
    int nparts = 0;
  } ev;

  double launches = 0;
  double prog_launches = 0;
  double bytes_alg = 0;
  // optional CUDA-event timing of every pass launch, accumulated per program id
  int time_passes = 0;
  double pass_ms[PROG_COUNT] = {0}, pass_count[PROG_COUNT] = {0}, pass_bytes[PROG_COUNT] = {0};
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool;
  struct PendingEv { int prog; size_t idx; double bytes; int chunk; int side; };
  std::vector<PendingEv> ev_pending;
  // timeline of the last harvested evaluation (option time_passes): start / end of every launch in ms after the first
  // launch's start, with its program, chunk (-1: whole pass) and stream (0 main, 1 side) -- see qaoa_get_timeline
  struct TimelineEntry { int prog, chunk, side; float t0, t1; };
  std::vector<TimelineEntry> timeline;
  double last_device_ms = 0;
  cudaEvent_t ev_call0 = nullptr, ev_call1 = nullptr;
  bool ev_call_recorded = false;
  std::string err;

  int small_max() const {
    if (small_max_qubits >= 0) return small_max_qubits;
    return dtype == QAOA_DTYPE_C64 ? 13 : 12;
  }
};

template <typename T>
void ensure(T*& p, size_t& cap, size_t need_bytes, bool pinned_host = false) {
  if (cap >= need_bytes && p) return;
  if (p) {
    if (pinned_host) cudaFreeHost(p); else cudaFree(p);
    p = nullptr;
  }
  size_t nb = std::max(need_bytes, (size_t)4096);
  if (pinned_host) CUDA_OK(cudaMallocHost((void**)&p, nb));
  else CUDA_OK(cudaMalloc((void**)&p, nb));
  cap = nb;
}

// Host <-> device copies of the set-up paths, ordered on the ENGINE's stream.  (The stream is non-blocking: a
// plain cudaMemcpy from pageable memory may return while its DMA is still in flight, and a kernel launched on
// the engine's stream right after it would not wait for it.)
void h2d(Engine* h, void* dst, const void* src, size_t bytes) {
  if (!bytes) return;
  CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
}
void d2h(Engine* h, void* dst, const void* src, size_t bytes) {
  if (!bytes) return;
  CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
}

void ensure_state(Engine* h, size_t points, size_t amps_per_point = 0) {
  if (!amps_per_point) amps_per_point = h->N;
  const size_t need = amps_per_point * h->amp_bytes * points;
  if (h->state && h->state_bytes >= need) return;
  if (h->state) { cudaFree(h->state); h->state = nullptr; h->state_bytes = 0; }
  CUDA_OK(cudaMalloc(&h->state, need));
  h->state_bytes = need;
  h->state_points = points;
}

// forget the peer pointers of the state (which = 0) or the cost diagonal (which = 1)
void close_peers(Engine* h, int which) {
  void** ptrs = which == 0 ? h->peer_state : h->peer_diag;
  bool* ipc = which == 0 ? h->peer_state_ipc : h->peer_diag_ipc;
  for (int r = 0; r < QAOA_MAX_RANKS; r++) {
    if (ptrs[r] && ipc[r]) cudaIpcCloseMemHandle(ptrs[r]);
    ptrs[r] = nullptr;
    ipc[r] = false;
  }
  if (which == 0) h->peers_state_set = false; else h->peers_diag_set = false;
}

void free_streams(Engine* h) {
  for (auto& kv : h->streams) cudaFree(kv.second);
  h->streams.clear();
}

void invalidate(Engine* h) {
  h->plans.clear();
  h->last_valid = false;
}

int grid_for(size_t N, int threads);

void drop_compact(Engine* h) {
  if (h->diag_compact) { cudaFree(h->diag_compact); h->diag_compact = nullptr; }
  if (h->strings_compact) { cudaFree(h->strings_compact); h->strings_compact = nullptr; }
  if (h->binom_compact) { cudaFree(h->binom_compact); h->binom_compact = nullptr; }
  h->compact_k = -1;
  h->n_compact = 0;
}

// compact diagonal for weight k (see gather_feasible_kernel); false if compression does not apply / pay off
bool ensure_compact(Engine* h, int k) {
  if (!h->compact_enabled || h->keep_state || h->world > 1 || h->n > 62) return false;
  if (h->compact_k == k) return true;
  drop_compact(h);
  std::vector<unsigned long long> binom((size_t)(h->n + 1) * (k + 1), 0ull);
  for (int b = 0; b <= h->n; b++)
    for (int j = 0; j <= k; j++) {
      unsigned long long v = j == 0 ? 1ull : (b == 0 ? 0ull : binom[(size_t)(b - 1) * (k + 1) + j] + binom[(size_t)(b - 1) * (k + 1) + j - 1]);
      binom[(size_t)b * (k + 1) + j] = v;
    }
  const unsigned long long M = binom[(size_t)h->n * (k + 1) + k];
  if (M == 0 || M > h->N / 4) return false;   // little to gain
  const int vb = diag_value_bytes(h->dinfo.kind);
  CUDA_OK(cudaMalloc((void**)&h->binom_compact, binom.size() * 8));
  unsigned long long* d_binom = h->binom_compact;
  h2d(h, d_binom, binom.data(), binom.size() * 8);
  CUDA_OK(cudaMalloc(&h->diag_compact, (size_t)M * vb));
  const int grid = grid_for((size_t)M, 256);
  if (vb == 1) gather_feasible_kernel<1><<<grid, 256, 0, h->stream>>>((const unsigned char*)h->diag, (unsigned char*)h->diag_compact, (size_t)M, h->n, k, d_binom);
  else if (vb == 4) gather_feasible_kernel<4><<<grid, 256, 0, h->stream>>>((const unsigned char*)h->diag, (unsigned char*)h->diag_compact, (size_t)M, h->n, k, d_binom);
  else gather_feasible_kernel<8><<<grid, 256, 0, h->stream>>>((const unsigned char*)h->diag, (unsigned char*)h->diag_compact, (size_t)M, h->n, k, d_binom);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaStreamSynchronize(h->stream));
  h->launches += 1;
  h->n_compact = (size_t)M;
  h->compact_k = k;
  return true;
}

// the weight-k strings themselves (XY mixer in the feasible subspace), built lazily next to the compact diagonal
bool ensure_compact_strings(Engine* h) {
  if (h->compact_k < 0) return false;
  if (h->strings_compact) return true;
  const size_t M = h->n_compact;
  const int grid = grid_for(M, 256);
  if (h->n <= 32) {
    CUDA_OK(cudaMalloc(&h->strings_compact, M * 4));
    unrank_feasible_kernel<unsigned><<<grid, 256, 0, h->stream>>>((unsigned*)h->strings_compact, M, h->n, h->compact_k, h->binom_compact);
  } else {
    CUDA_OK(cudaMalloc(&h->strings_compact, M * 8));
    unrank_feasible_kernel<unsigned long long><<<grid, 256, 0, h->stream>>>((unsigned long long*)h->strings_compact, M, h->n,
                                                                          h->compact_k, h->binom_compact);
  }
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaStreamSynchronize(h->stream));
  h->launches += 1;
  return true;
}

// Symmetric (Z2) registers (tile_kernel.cuh): MaxCut-like costs with cost(x) = cost(~x), X / multi-angle X
// mixer, |+> initial state, even n.  The tile path then stores the half register with the top qubit = 0.
bool cost_is_symmetric(Engine* h);
bool use_z2(Engine* h) {
  if (h->ztw_mode) {   // chosen by the caller at creation: everything else is rejected, never re-routed
    if (h->mixer != MIXER_X && h->mixer != MIXER_XMULTI) fail("sharded symmetric register: X / multi-angle X mixers only");
    if (h->init.kind != INIT_PLUS || h->n_gamma != 1 || h->n_init != 0) fail("sharded symmetric register: |+> initial state only");
    if (h->cost_symmetric != 1) fail("sharded symmetric register: the cost function is not known to be invariant under the global bit flip");
    if (h->keep_state || h->schedule == 1) fail("sharded symmetric register: keep_state / the classic schedule are not supported");
    return true;
  }
  if (!h->z2_enabled || h->world > 1 || h->keep_state || (h->n & 1)) return false;
  if (h->mixer != MIXER_X && h->mixer != MIXER_XMULTI) return false;
  if (h->init.kind != INIT_PLUS || h->n_gamma != 1 || h->n_init != 0) return false;
  if (h->E - 1 < TILE_BITS + 1) return false;          // the half register still needs a high tile shape
  return cost_is_symmetric(h);
}
inline int plan_el(const Engine* h, const Plan& pl) { return (pl.z2 && !h->ztw_mode) ? h->El - 1 : h->El; }      // stored element bits
inline size_t plan_amps(const Engine* h, const Plan& pl) { return (pl.z2 && !h->ztw_mode) ? h->N >> 1 : h->N; }  // stored amplitudes
inline unsigned plan_ntiles(const Engine* h, const Plan& pl) { return 1u << (plan_el(h, pl) - TILE_BITS); }

bool cost_is_symmetric(Engine* h) {
  if (h->cost_symmetric >= 0) return h->cost_symmetric == 1;
  if (h->dinfo.kind == DIAG_NONE || !h->diag || h->world > 1) return false;   // (sharded: known by construction only)
  DevTmp t_flag(sizeof(int));
  int* d_flag = t_flag.as<int>();
  CUDA_OK(cudaMemsetAsync(d_flag, 0, sizeof(int), h->stream));
  const int grid = grid_for(h->N / 2, 256);
  if (h->dinfo.kind == DIAG_U8) diag_symmetry_kernel<uint8_t><<<grid, 256, 0, h->stream>>>((const uint8_t*)h->diag, h->N, d_flag);
  else if (h->dinfo.kind == DIAG_U32FX) diag_symmetry_kernel<uint32_t><<<grid, 256, 0, h->stream>>>((const uint32_t*)h->diag, h->N, d_flag);
  else diag_symmetry_kernel<unsigned long long><<<grid, 256, 0, h->stream>>>((const unsigned long long*)h->diag, h->N, d_flag);
  CUDA_OK(cudaGetLastError());
  h->launches += 1;
  int flag = 1;
  d2h(h, &flag, d_flag, sizeof(int));
  h->cost_symmetric = flag ? 0 : 1;
  return h->cost_symmetric == 1;
}

// sharding fields of a pass descriptor (see TilePass)
void shard_fill(const Engine* h, TilePass& tp, bool cross) {
  tp.shard_bits = h->world > 1 ? h->El : 0;
  tp.rank_popc = __builtin_popcount((unsigned)h->rank);
  tp.tile0 = 0;
  tp.zdelta = 0; tp.ztw = 0; tp.xgbits = 0;
  if (h->ztw_mode) {
    tp.zdelta = h->gbits - 2 * tp.rank_popc;
    tp.ztw = h->gbits & 1;
    tp.xgbits = cross ? h->gbits : 0;
  }
  if (h->world > 1) {
    // tiles enumerate LOCAL free bits only (a cross tile has the rank bits among its tile bits, any
    // other tile lives entirely inside the local shard)
    int nf = 0;
    for (int j = 0; j < tp.nfree; j++) if (tp.freepos[j] < h->El) tp.freepos[nf++] = tp.freepos[j];
    tp.nfree = nf;
    if (!tilepass_set_free_runs(tp)) fail("tile shape with more than two runs of free bits");
    if (nf != h->El - (cross ? TILE_BITS - h->gbits : TILE_BITS)) fail("sharded tile shape straddles the shard boundary");
  }
  if (cross) {
    if (!h->peers_state_set) fail("sharded engine: peer state pointers have not been set");
    tp.tile0 = (unsigned)h->rank << (h->El - TILE_BITS);
    for (int r = 0; r < h->world; r++) { tp.peers[r] = h->peer_state[r]; tp.dpeers[r] = h->peer_diag[r]; }
  }
}

int grid_for(size_t N, int threads) {
  size_t b = (N + threads - 1) / threads;
  return (int)std::min<size_t>(b, 148 * 16);
}

// ------------------------------------------------------------------------------------------
// planner: which 14 element bits each pass tiles over, and the op program of each pass
// ------------------------------------------------------------------------------------------
// register-resident tile-local bits for (frame, sub):  sub 0 = frame proper, sub 1 = after WT
void reg_bits(int frame, int sub, int out[5]) {
  if (sub == 1) { for (int i = 0; i < 5; i++) out[i] = 1 + i; return; }
  out[0] = 0;
  for (int i = 0; i < 4; i++) out[1 + i] = (frame == 0 ? 10 : 6) + i;
}

// Balanced schedule (complex64, 5 low bits, >= 3 tile shapes).  Shape 0 is the low tile (element
// bits 0..13), shapes 1.. are the high tiles (5 low bits + 9 high bits).  Per layer d:
//   MID    on shape 0            : all 14 qubits of the low tile                       (2 transposes)
//   INTERIOR on all but one high : the tile's high qubits                              (2 transposes)
//   BOUND2 on the last high tile : its high qubits of layer d, phase d+1, the same
//                                  qubits of layer d+1                                 (2 transposes)
// against the classic INTERIOR (2 transposes) + BOUNDARY (4 transposes) pair.  The last layer ends
// with FINAL2 (1 transpose, fused reduction).  Same number of HBM sweeps, a third less shared-memory
// traffic and evenly split arithmetic.
// pass descriptor fields of the low tile of a symmetric register: the virtual top qubit (logical element bit E - 1
// in Shape::pos[13]) is addressed as the element bit right above the stored (local) register
void set_zshape(const Engine* h, const Plan& pl, TilePass& tp) {
  tp.zbit = plan_el(h, pl);
  tp.pos[TILE_BITS - 1] = tp.zbit;
  tp.zsign = ((h->n - 2) / 2) % 2 ? -1 : 1;
  tp.znreal = h->n - 1;
}

void build_balanced(Engine* h, int p, Plan& pl) {
  const int E = h->E, qs = h->qshift, m = (int)pl.shapes.size();
  std::vector<std::vector<char>> done(p, std::vector<char>(E, 0));
  std::vector<std::vector<char>> gdone(p + 1, std::vector<char>(m, 0));
  const bool load_first = (h->init.kind == INIT_STATEVECTOR);
  // tile-local bit held by register slot s:  kind 0..4 = frames A, B, C, A4, C4;  kind 5 = after WT
  auto reg_tile_bit = [](int kind, int s) {
    if (kind == 5) return 1 + s;
    return frame_reg_bit(kind, s);
  };
  const bool low4 = pl.shapes[m > 1 ? 1 : 0].pos[4] != 4;   // high tiles carry 4 low bits
  struct Builder {
    PlanPass pp;
    const Shape* sh;
  };
  auto begin = [&](int shape) {
    Builder b;
    b.sh = &pl.shapes[shape];
    b.pp.cross = b.sh->cross;
    b.pp.shape_index = shape;
    memset(&b.pp.tp, 0, sizeof(TilePass));
    TilePass& tp = b.pp.tp;
    for (int j = 0; j < TILE_BITS; j++) tp.pos[j] = b.sh->pos[j];
    tp.nfree = (int)b.sh->freepos.size();
    for (int j = 0; j < tp.nfree; j++) tp.freepos[j] = b.sh->freepos[j];
    if (!tilepass_set_free_runs(tp)) fail("tile shape with more than two runs of free bits");
    tp.qshift = qs;
    tp.init_kind = h->init.kind;
    tp.dicke_k = h->init.k;
    tp.init_amp = h->init.amp * (pl.z2 ? sqrt(2.0) : 1.0);   // the stored half carries the whole norm
    b.pp.zshape = b.sh->zshape;
    if (b.sh->zshape) set_zshape(h, pl, tp);
    return b;
  };
  auto push = [&](Builder& b, int type, int fr, int arg, int mask) {
    TilePass& tp = b.pp.tp;
    if (tp.nops >= TILE_MAX_OPS) fail("tile pass program overflow");
    tp.ops[tp.nops].type = (int16_t)type;
    tp.ops[tp.nops].frame = (int16_t)fr;
    tp.ops[tp.nops].arg = (int16_t)arg;
    tp.ops[tp.nops].mask = (int16_t)mask;
    tp.nops++;
  };
  auto round = [&](Builder& b, int layer, int kind, int mask) {
    RoundSpec rs;
    rs.layer = layer;
    for (int s = 0; s < 5; s++) {
      rs.ebit[s] = -1;
      if (!((mask >> s) & 1)) continue;
      const int eb = b.sh->pos[reg_tile_bit(kind, s)];
      if (h->ztw_mode && eb == qs) continue;   // twisted shards: qubit 0 is mixed in the cross passes only
      if (eb >= qs && !done[layer][eb]) { rs.ebit[s] = eb; done[layer][eb] = 1; }
    }
    if (h->ztw_mode && b.sh->cross && kind == 0 && !done[layer][qs]) { rs.ebit[0] = qs; done[layer][qs] = 1; }
    // frame field of a ROUND: 1 = slot 4 is the mirror pairing of a symmetric register (frame A of the low tile)
    push(b, TOP_ROUND, (b.sh->zshape && kind == 0 && ((mask >> 4) & 1)) ? 1 : 0, (int)b.pp.rounds.size(), mask);
    b.pp.rounds.push_back(rs);
  };
  auto phase = [&](Builder& b, int layer, int shape, int frame) {
    push(b, TOP_PHASE, frame, (int)b.pp.phases.size(), 0);
    b.pp.phases.push_back(PhaseSpec{layer, shape, frame});
  };
  auto finish = [&](Builder& b) {
    PlanPass& pp = b.pp;
    TilePass& tp = pp.tp;
    tp.nrounds = (int)pp.rounds.size();
    tp.nphases = (int)pp.phases.size();
    tp.tau_stride = tp.nrounds * TILE_ROUND_REALS;
    tp.phs_stride = 2 * tp.nphases + 1;
    for (int j = 0; j < tp.nphases; j++) tp.table_id[j] = pp.phases[j].layer;
    pp.tau_offset = pl.tau_per_point;
    pp.phs_offset = pl.phs_per_point;
    pl.tau_per_point += tp.tau_stride;
    pl.phs_per_point += tp.phs_stride;
    pp.prog = PROG_GENERIC;
    if (prog_matches<ProgFirstInit>(tp)) pp.prog = PROG_FIRST_INIT;
    else if (prog_matches<ProgFirstLoad>(tp)) pp.prog = PROG_FIRST_LOAD;
    else if (prog_matches<ProgInterior>(tp)) pp.prog = PROG_INTERIOR;
    else if (prog_matches<ProgMid>(tp)) pp.prog = PROG_MID;
    else if (prog_matches<ProgBound2>(tp)) pp.prog = PROG_BOUND2;
    else if (prog_matches<ProgFinal2>(tp)) pp.prog = PROG_FINAL2;
    else if (prog_matches<ProgBoundX>(tp)) pp.prog = PROG_BOUNDX;
    else if (prog_matches<ProgFinalX>(tp)) pp.prog = PROG_FINALX;
    else if (prog_matches<ProgBound2L4>(tp)) pp.prog = PROG_BOUND2L4;
    else if (prog_matches<ProgFinal2L4>(tp)) pp.prog = PROG_FINAL2L4;
    if (pp.prog == PROG_GENERIC) fail("balanced planner produced a pass without a compiled program");
    pl.passes.push_back(pp);
  };
  auto full_cover = [&](Builder& b, int layer) {   // R(all) WT R(all) BT R(hi4): frame A -> B
    round(b, layer, 0, M_ALL);
    push(b, TOP_WT, 0, 0, 0);
    round(b, layer, 5, M_ALL);
    push(b, TOP_BT, 0, 0, 0);
    round(b, layer, 1, M_HI4);
  };
  {  // FIRST: initial state, phase 0, the low tile of layer 0
    Builder b = begin(0);
    push(b, load_first ? TOP_LOAD : TOP_INIT, 0, 0, 0);
    phase(b, 0, 0, 0);
    full_cover(b, 0);
    push(b, TOP_STORE, 1, 0, 0);
    finish(b);
    gdone[0][0] = 1;
  }
  for (int d = 0; d < p; d++) {
    std::vector<int> rem;
    for (int g = 1; g < m; g++) if (!gdone[d][g]) rem.push_back(g);
    if (rem.empty()) fail("balanced planner: no high tile left for the layer boundary");
    for (size_t i = 0; i + 1 < rem.size(); i++) {   // INTERIOR passes
      Builder b = begin(rem[i]);
      push(b, TOP_LOAD, 0, 0, 0);
      round(b, d, 0, M_HI4);
      push(b, TOP_WT, 0, 0, 0);
      round(b, d, 5, M_TOP2);
      push(b, TOP_BT, 0, 0, 0);
      round(b, d, 1, M_HI4);
      push(b, TOP_STORE, 1, 0, 0);
      finish(b);
      gdone[d][rem[i]] = 1;
    }
    const int g = rem.back();
    Builder b = begin(g);
    const bool xs = b.sh->cross;   // cross shape: only the rank bits (register bits of frame A) are left
    const bool l4 = low4 && !xs;   // 4 low bits: frames A4 / C4, five qubits per round
    const int fio = l4 ? 3 : 0, fmid = xs ? 0 : (l4 ? 4 : 2);
    push(b, TOP_LOAD, fio, 0, 0);
    round(b, d, fio, l4 ? M_ALL : M_HI4);
    if (!xs) {
      push(b, TOP_XT, 0, 0, 0);
      round(b, d, fmid, M_ALL);
    }
    gdone[d][g] = 1;
    for (int e = qs; e < E; e++) if (!done[d][e]) fail("balanced planner: layer left incomplete");
    if (d == p - 1) {
      b.pp.has_reduce = true;
      b.pp.reduce_shape = g;
      b.pp.reduce_frame = fmid;
      push(b, TOP_REDUCE, fmid, (int)b.pp.phases.size(), 0);
      finish(b);
      break;
    }
    phase(b, d + 1, g, fmid);
    if (!xs) {
      round(b, d + 1, fmid, M_ALL);
      push(b, TOP_XTI, 0, 0, 0);
    }
    round(b, d + 1, fio, l4 ? M_ALL : M_HI4);
    push(b, TOP_STORE, fio, 0, 0);
    finish(b);
    gdone[d + 1][g] = 1;
    Builder mb = begin(0);   // MID: the low tile of layer d+1
    push(mb, TOP_LOAD, 0, 0, 0);
    full_cover(mb, d + 1);
    push(mb, TOP_STORE, 1, 0, 0);
    finish(mb);
    gdone[d + 1][0] = 1;
  }
}

Plan build_plan(Engine* h, int p) {
  Plan pl;
  pl.p = p;
  const int E = h->E, qs = h->qshift;
  if (E < TILE_BITS) fail("tile path needs at least 14 element bits");
  // c low bits ride along in every tile (rows of 2^c * 8 bytes).  256-byte rows (c = 5) run at full
  // DRAM speed at every stride, 128-byte rows lose ~15 % at the largest strides; use c = 5 whenever it
  // does not cost an extra pass per layer.
  // Sharded registers: the local shapes tile the El bits of the shard exactly like an unsharded
  // register of that size; the rank bits get a shape of their own (appended below).
  pl.z2 = use_z2(h);
  const int EL = plan_el(h, pl);
  // Sharded registers, balanced schedule: the cross tile has register slots to spare in frame A (t10 .. t13 minus
  // the rank bits), and a qubit that lives there is mixed for BOTH layers of every cross pass -- it needs no other
  // pass.  Move the top kx local bits into the cross tile when that lets the local shapes keep 5 low bits
  // (256-byte rows) without an extra pass per layer: ELC = the bits the local shapes have to cover.
  int kx = 0;
  if (h->world > 1 && h->cross_spare && h->dtype == QAOA_DTYPE_C64 && h->schedule != 1 && h->low_bits == 0) {
    auto counts = [&](int el, int* m4, int* m5) {
      if (pl.z2) { *m4 = (el - 13 + 9) / 10; *m5 = (el - 13 + 8) / 9; }
      else { *m4 = std::max(1, (el - 4 + 9) / 10); *m5 = std::max(1, (el - 5 + 8) / 9); }
    };
    int m4, m5;
    counts(EL, &m4, &m5);
    if (h->cross_spare >= 10) {   // test hook: force k = cross_spare - 10 spare bits
      kx = std::min(h->cross_spare - 10, 4 - h->gbits);
      kx = std::max(0, std::min(kx, EL - TILE_BITS - 1));
    } else if (m5 != m4)
      for (int k = 1; k <= 4 - h->gbits; k++) {
        int a4, a5;
        counts(EL - k, &a4, &a5);
        if (a5 == a4 && a5 <= m4 && EL - k >= TILE_BITS + 1) { kx = k; break; }
      }
  }
  const int ELC = EL - kx;
  pl.kx = kx;
  if (pl.z2) {
    // Symmetric register: EL = E - 1 real element bits.  Shape 0 = 13 real low bits + the virtual top qubit
    // (element bit E - 1); its tiles enumerate the free bits 13 .. EL-2 (the top real bit is 0 in the plain
    // half of a tile and 1 in its mirror image).  The high shapes tile the real bits 13 .. EL-1.
    const int nHz = ELC - 13;
    const int m4 = (nHz + 9) / 10, m5 = (nHz + 8) / 9;
    int clow = (m5 == m4) ? 5 : 4;
    if (h->low_bits == 4 || h->low_bits == 5) clow = h->low_bits;
    const int hper = TILE_BITS - clow;
    Shape s0;
    for (int j = 0; j < 13; j++) s0.pos[j] = j;
    s0.pos[13] = E - 1;
    for (int b = 13; b < EL - 1; b++) s0.freepos.push_back(b);
    s0.zshape = true;
    pl.shapes.push_back(s0);
    const int mh = (nHz + hper - 1) / hper;
    for (int i = 0; i < mh; i++) {
      std::vector<int> bits;
      for (int b = 0; b < clow; b++) bits.push_back(b);
      int lo = 13 + hper * i, hi = std::min(ELC, lo + hper);
      std::vector<int> own;
      for (int b = lo; b < hi; b++) own.push_back(b);
      int padb = lo - 1;
      while ((int)own.size() < hper) { own.push_back(padb); padb--; }
      for (int b : own) bits.push_back(b);
      std::sort(bits.begin(), bits.end());
      Shape s;
      for (int j = 0; j < TILE_BITS; j++) s.pos[j] = bits[j];
      for (int b = 0; b < EL; b++)
        if (!std::binary_search(bits.begin(), bits.end(), b)) s.freepos.push_back(b);
      pl.shapes.push_back(s);
    }
  }
  const int m4 = std::max(1, (ELC - 4 + 9) / 10), m5 = std::max(1, (ELC - 5 + 8) / 9);
  int clow = (m5 == m4 && ELC >= 15) ? 5 : 4;
  if (h->low_bits == 4 || h->low_bits == 5) clow = h->low_bits;
  if (pl.z2) clow = pl.shapes[1].pos[4] == 4 ? 5 : 4;
  const int hper = TILE_BITS - clow;
  const int nH = ELC - clow;
  const int m_local = pl.z2 ? 0 : std::max(1, (nH + hper - 1) / hper);
  for (int i = 0; i < m_local; i++) {
    std::vector<int> bits;
    for (int b = 0; b < clow; b++) bits.push_back(b);
    int lo = clow + hper * i, hi = std::min(ELC, lo + hper);
    std::vector<int> own;
    for (int b = lo; b < hi; b++) own.push_back(b);
    int padb = lo - 1;
    while ((int)own.size() < hper) { own.push_back(padb); padb--; }
    for (int b : own) bits.push_back(b);
    std::sort(bits.begin(), bits.end());
    Shape s;
    for (int j = 0; j < TILE_BITS; j++) s.pos[j] = bits[j];
    for (int b = 0; b < E; b++)
      if (!std::binary_search(bits.begin(), bits.end(), b)) s.freepos.push_back(b);
    s.cross = false;
    pl.shapes.push_back(s);
  }
  if (h->world > 1) {
    // CROSS shape: the 14 - gbits lowest element bits + the rank bits.  On every rank a tile is then
    // ONE contiguous run of 2^(14-gbits) elements (64 KB at 2 ranks, 16 KB at 8): peer traffic is
    // sequential, which is what NVLink and the peer TLBs want (tiles made of 256-byte rows 128 MB
    // apart ran at < 200 GB/s per direction at n = 33).
    // (+ the top kx local bits, see above: 2^kx runs per rank)
    Shape s;
    const int nlow = TILE_BITS - h->gbits - kx;
    for (int j = 0; j < nlow; j++) s.pos[j] = j;
    for (int j = 0; j < kx; j++) s.pos[nlow + j] = ELC + j;
    for (int j = 0; j < h->gbits; j++) s.pos[TILE_BITS - h->gbits + j] = EL + j;
    for (int b = nlow; b < ELC; b++) s.freepos.push_back(b);
    s.cross = true;
    pl.shapes.push_back(s);
  }
  const int m = (int)pl.shapes.size();
  // balanced schedule: complex64 with 4 or 5 low bits, complex128 with 5 (no A4 / C4 frames there)
  if (h->schedule != 1 && (h->dtype == QAOA_DTYPE_C64 || clow == 5) && m >= 3 && !h->keep_state) {
    build_balanced(h, p, pl);
    return pl;
  }
  std::vector<std::vector<char>> done(p, std::vector<char>(E, 0));
  auto layer_complete = [&](int d) {
    for (int b = qs; b < E; b++) if (!done[d][b]) return false;
    return true;
  };
  int cur = 0;
  bool phase_pending = true;
  bool first = true;
  bool finished = false;
  const bool load_first = (h->init.kind == INIT_STATEVECTOR);
  int guard = 0;
  while (!finished) {
    if (++guard > 100000) fail("planner did not terminate");
    // choose the shape with most unmixed qubits in the current layer
    int best = 0, bestc = -1;
    for (int s = 0; s < m; s++) {
      int c = 0;
      for (int j = 0; j < TILE_BITS; j++) {
        int b = pl.shapes[s].pos[j];
        if (b >= qs && !done[cur][b]) c++;
      }
      if (c > bestc) { bestc = c; best = s; }
    }
    const Shape& sh = pl.shapes[best];
    PlanPass pp;
    pp.cross = sh.cross;
    pp.shape_index = best;
    memset(&pp.tp, 0, sizeof(TilePass));
    TilePass& tp = pp.tp;
    for (int j = 0; j < TILE_BITS; j++) tp.pos[j] = sh.pos[j];
    tp.nfree = (int)sh.freepos.size();
    for (int j = 0; j < tp.nfree; j++) tp.freepos[j] = sh.freepos[j];
    if (!tilepass_set_free_runs(tp)) fail("tile shape with more than two runs of free bits");
    tp.qshift = qs;
    tp.init_kind = h->init.kind;
    tp.dicke_k = h->init.k;
    tp.init_amp = h->init.amp * (pl.z2 ? sqrt(2.0) : 1.0);
    pp.zshape = sh.zshape;
    if (sh.zshape) set_zshape(h, pl, tp);
    int frame = 0;
    auto push = [&](int type, int fr, int arg, int mask) {
      if (tp.nops >= TILE_MAX_OPS) fail("tile pass program overflow");
      tp.ops[tp.nops].type = (int16_t)type;
      tp.ops[tp.nops].frame = (int16_t)fr;
      tp.ops[tp.nops].arg = (int16_t)arg;
      tp.ops[tp.nops].mask = (int16_t)mask;
      tp.nops++;
    };
    if (first && !load_first) push(TOP_INIT, frame, 0, 0); else push(TOP_LOAD, frame, 0, 0);
    first = false;
    bool reduced = false;
    while (true) {
      // room check: one more layer segment needs <= 1 phase + 5 cover steps, plus the tail
      const bool room = tp.nops + 9 <= TILE_MAX_OPS && (int)pp.phases.size() + 2 <= TILE_MAX_PHASES;
      if (!room) break;
      // qubits of this tile still unmixed in layer `cur` (with a pending phase: all of them)
      std::vector<char> need(TILE_BITS, 0);
      int nneed = 0;
      bool need_low = false;
      for (int j = 0; j < TILE_BITS; j++) {
        int b = sh.pos[j];
        if (b >= qs && !done[cur][b]) { need[j] = 1; nneed++; if (j < 4) need_low = true; }
        // (with 5 low bits, tile bit 4 is reached by the high cover's lane round as well)
      }
      if (nneed == 0) break;
      if (phase_pending) {
        PhaseSpec ps{cur, best, frame};
        push(TOP_PHASE, frame, (int)pp.phases.size(), 0);
        pp.phases.push_back(ps);
        phase_pending = false;
      }
      // canonical cover: R WT R BT R, full (all 14 bits) or high (upper 10 bits)
      const int masks_full[3] = {M_ALL, M_ALL, M_HI4}, masks_high[3] = {M_HI4, M_TOP2, M_HI4};
      const int* masks = need_low ? masks_full : masks_high;
      for (int step = 0; step < 3; step++) {
        int rb[5];
        reg_bits(frame, step == 1 ? 1 : 0, rb);
        RoundSpec rs;
        rs.layer = cur;
        for (int s = 0; s < 5; s++) {
          rs.ebit[s] = -1;
          if (((masks[step] >> s) & 1) && need[rb[s]]) {
            rs.ebit[s] = sh.pos[rb[s]];
            need[rb[s]] = 0;
            nneed--;
            done[cur][sh.pos[rb[s]]] = 1;
          }
        }
        push(TOP_ROUND, (sh.zshape && step != 1 && frame == 0 && ((masks[step] >> 4) & 1)) ? 1 : 0, (int)pp.rounds.size(), masks[step]);
        pp.rounds.push_back(rs);
        if (step == 0) push(TOP_WT, 0, 0, 0);
        if (step == 1) { push(TOP_BT, 0, 0, 0); frame ^= 1; }
      }
      if (nneed != 0) fail("planner: cover left qubits unmixed");
      if (layer_complete(cur)) {
        if (cur == p - 1) {
          pp.has_reduce = true;
          pp.reduce_shape = best;
          pp.reduce_frame = frame;
          push(TOP_REDUCE, frame, (int)pp.phases.size(), 0);
          reduced = true;
          finished = true;
          break;
        }
        cur++;
        phase_pending = true;
        continue;
      }
      break;
    }
    if (tp.nops == 1 && !reduced) fail("planner: empty pass");
    if (!reduced || h->keep_state) push(TOP_STORE, frame, 0, 0);
    tp.nrounds = (int)pp.rounds.size();
    tp.nphases = (int)pp.phases.size();
    tp.tau_stride = tp.nrounds * TILE_ROUND_REALS;
    tp.phs_stride = 2 * tp.nphases + 1;
    for (int j = 0; j < tp.nphases; j++) tp.table_id[j] = pp.phases[j].layer;
    pp.tau_offset = pl.tau_per_point;
    pp.phs_offset = pl.phs_per_point;
    pl.tau_per_point += tp.tau_stride;
    pl.phs_per_point += tp.phs_stride;
    pp.prog = PROG_GENERIC;
    if (prog_matches<ProgFirstInit>(tp)) pp.prog = PROG_FIRST_INIT;
    else if (prog_matches<ProgFirstLoad>(tp)) pp.prog = PROG_FIRST_LOAD;
    else if (prog_matches<ProgInterior>(tp)) pp.prog = PROG_INTERIOR;
    else if (prog_matches<ProgBoundary>(tp)) pp.prog = PROG_BOUNDARY;
    else if (prog_matches<ProgFinal>(tp)) pp.prog = PROG_FINAL;
    else if (prog_matches<ProgMid>(tp)) pp.prog = PROG_MID;
    else if (prog_matches<ProgBound2>(tp)) pp.prog = PROG_BOUND2;
    else if (prog_matches<ProgFinal2>(tp)) pp.prog = PROG_FINAL2;
    pl.passes.push_back(pp);
  }
  // complex64: a last pass of the classic FINAL form (high cover through WT + BT, fused reduction)
  // becomes the one-transpose FINAL2 / FINAL2L4 form (p = 1 landscapes at 2-shape sizes: FIRST + this)
  // (the mirrored low tile of a symmetric register has no A4 / C4 addressing: with 4 low bits its last pass keeps
  // the classic FINAL form, whose cover reaches all ten bits t4..t13)
  if (h->dtype == QAOA_DTYPE_C64 && !h->keep_state && !pl.passes.empty() && pl.passes.back().prog == PROG_FINAL &&
      !pl.passes.back().cross && !(pl.passes.back().zshape && clow == 4)) {
    PlanPass& pp = pl.passes.back();
    TilePass& tp = pp.tp;
    std::vector<char> want(E, 0);
    int layer = p - 1;
    for (const auto& rs : pp.rounds) { layer = rs.layer; for (int k = 0; k < 5; k++) if (rs.ebit[k] >= 0) want[rs.ebit[k]] = 1; }
    const bool l4 = clow == 4;
    const int fio = l4 ? 3 : 0, fmid = l4 ? 4 : 2;
    pl.tau_per_point -= tp.tau_stride;
    pp.rounds.clear();
    tp.nops = 0;
    auto push = [&](int type, int fr, int arg, int mask) {
      tp.ops[tp.nops].type = (int16_t)type; tp.ops[tp.nops].frame = (int16_t)fr;
      tp.ops[tp.nops].arg = (int16_t)arg; tp.ops[tp.nops].mask = (int16_t)mask; tp.nops++;
    };
    auto round = [&](int frame, int mask) {
      RoundSpec rs;
      rs.layer = layer;
      for (int k = 0; k < 5; k++) {
        rs.ebit[k] = -1;
        const int eb = tp.pos[frame_reg_bit(frame, k)];
        if (((mask >> k) & 1) && want[eb]) { rs.ebit[k] = eb; want[eb] = 0; }
      }
      push(TOP_ROUND, (pp.zshape && frame == 0 && ((mask >> 4) & 1)) ? 1 : 0, (int)pp.rounds.size(), mask);
      pp.rounds.push_back(rs);
    };
    push(TOP_LOAD, fio, 0, 0);
    round(fio, l4 ? M_ALL : M_HI4);
    push(TOP_XT, 0, 0, 0);
    round(fmid, M_ALL);
    for (int b = 0; b < E; b++) if (want[b]) fail("planner: FINAL2 rewrite left qubits unmixed");
    pp.reduce_frame = fmid;
    push(TOP_REDUCE, fmid, 0, 0);
    tp.nrounds = (int)pp.rounds.size();
    tp.tau_stride = tp.nrounds * TILE_ROUND_REALS;
    pl.tau_per_point += tp.tau_stride;
    pp.prog = l4 ? PROG_FINAL2L4 : PROG_FINAL2;
    if (!(l4 ? prog_matches<ProgFinal2L4>(tp) : prog_matches<ProgFinal2>(tp))) fail("planner: FINAL2 rewrite mismatch");
  }
  return pl;
}

void* get_stream(Engine* h, const Plan& pl, int shape, int frame) {
  auto key = std::make_pair(shape + (pl.z2 ? 100 : 0), frame);
  auto it = h->streams.find(key);
  if (it != h->streams.end()) return it->second;
  const int vb = diag_value_bytes(h->dinfo.kind);
  void* s = nullptr;
  // complex128 in frame C: one value per REGISTER (the re / im halves of an amplitude sit in neighbouring lanes)
  const bool per_reg = h->dtype == QAOA_DTYPE_C128 && frame == 2;
  CUDA_OK(cudaMalloc(&s, plan_amps(h, pl) * vb * (per_reg ? 2 : 1)));
  TilePass tp;
  memset(&tp, 0, sizeof(tp));
  const Shape& sh = pl.shapes[shape];
  for (int j = 0; j < TILE_BITS; j++) tp.pos[j] = sh.pos[j];
  tp.nfree = (int)sh.freepos.size();
  for (int j = 0; j < tp.nfree; j++) tp.freepos[j] = sh.freepos[j];
  if (!tilepass_set_free_runs(tp)) fail("tile shape with more than two runs of free bits");
  tp.qshift = h->qshift;
  if (sh.zshape) set_zshape(h, pl, tp);
  const int cross = sh.cross ? 1 : 0;
  if (cross && !h->peers_diag_set) fail("sharded engine: peer cost-diagonal pointers have not been set");
  shard_fill(h, tp, sh.cross);
  const unsigned ntiles = plan_ntiles(h, pl);
  const int namp = (h->dtype == QAOA_DTYPE_C64 || per_reg) ? 32 : 16;
  if (frame >= 3 && namp != 32) fail("frame A4 / C4 streams exist for complex64 only");
  if (vb == 1) make_stream_kernel<1><<<ntiles, TILE_THREADS, 0, h->stream>>>(tp, frame, namp, cross, (const unsigned char*)h->diag, (unsigned char*)s);
  else if (vb == 4) make_stream_kernel<4><<<ntiles, TILE_THREADS, 0, h->stream>>>(tp, frame, namp, cross, (const unsigned char*)h->diag, (unsigned char*)s);
  else make_stream_kernel<8><<<ntiles, TILE_THREADS, 0, h->stream>>>(tp, frame, namp, cross, (const unsigned char*)h->diag, (unsigned char*)s);
  CUDA_OK(cudaGetLastError());
  h->launches += 1;
  h->streams[key] = s;
  return s;
}

void plan_chunks(Engine* h, Plan& pl);
void set_chunk_fields(const Plan& pl, PlanPass& pp, int c);

Plan& get_plan(Engine* h, int p) {
  auto it = h->plans.find(p);
  if (it != h->plans.end()) return it->second;
  Plan pl = build_plan(h, p);
  plan_chunks(h, pl);
  // attach the diagonal streams
  for (auto& pp : pl.passes) {
    for (size_t j = 0; j < pp.phases.size(); j++)
      pp.tp.streams[j] = get_stream(h, pl, pp.phases[j].shape, pp.phases[j].frame);
    if (pp.has_reduce) pp.tp.streams[pp.phases.size()] = get_stream(h, pl, pp.reduce_shape, pp.reduce_frame);
  }
  h->plans[p] = pl;
  return h->plans[p];
}

// beta of qubit q in layer d for one point (reference flat layout, qaoa/qaoa.py:937-990)
inline double beta_of(const Engine* h, const double* ang, int d, int q) {
  const int nb = h->mixer == MIXER_XMULTI ? h->n : 1;
  const int per = 1 + nb;
  return ang[d * per + 1 + (nb == 1 ? 0 : q)];
}
inline double gamma_of(const Engine* h, const double* ang, int d) {
  const int nb = h->mixer == MIXER_XMULTI ? h->n : 1;
  return ang[d * (1 + nb)];
}

// largest |tan beta| handled in the normal (lifted) form: bounds the per-layer growth so that the
// pre-shrunk amplitudes stay inside the normal exponent range of the state's real type
inline double tmax_normal_form(const Engine* h) {
  const double budget = h->dtype == QAOA_DTYPE_C64 ? 340.0 : 1800.0;  // 2 * log2 range allowed per layer
  return sqrt(pow(2.0, std::min(budget / h->n, 1000.0)) - 1.0);
}

// Lowers one point's angles into round coefficients (tau: doubles here, narrowed by the caller) and
// phase parameters.  Returns false in *all_normal if any slot needs the cot form.
void fill_point_params(Engine* h, const Plan& pl, const double* ang, double* tau /*tau_per_point*/,
                       double* phs /*phs_per_point*/, double* gamsig /*2*p or null*/,
                       std::vector<char>* pass_all_normal, double* final_scale_out) {
  const int p = pl.p;
  std::vector<double> logS(p, 0.0);   // log |prod scale| per layer
  std::vector<int> sgnS(p, 1);
  const double tmax = tmax_normal_form(h);
  if (pass_all_normal) pass_all_normal->assign(pl.passes.size(), 1);
  // The fixed bound tmax assumes EVERY qubit of the layer rotates by the same angle (plain X mixer).  With per-qubit
  // angles (XMultiAngle / XOrbit) a few angles near pi/2 are common, and what has to stay inside the exponent range is
  // the layer's TOTAL growth sum_q -log2|cos beta_q|: when that fits the budget every qubit of the layer keeps the
  // normal (lifted) form -- no cot form, no interpreter passes -- however large a single |tan beta_q| is.
  std::vector<char> layer_fits(p, 0);
  if (h->mixer == MIXER_XMULTI && h->layer_budget) {
    const double budget_bits = 0.5 * (h->dtype == QAOA_DTYPE_C64 ? 340.0 : 1800.0);
    for (int d = 0; d < p; d++) {
      double bits = 0.0;
      for (int q = 0; q < h->n; q++) {
        const double c = fabs(cos(beta_of(h, ang, d, q)));
        bits += c > 1e-280 ? -log2(c) : 1e9;
      }
      layer_fits[d] = bits <= budget_bits;
    }
  }
  for (size_t ip = 0; ip < pl.passes.size(); ip++) {
    const auto& pp = pl.passes[ip];
    double* base = tau + pp.tau_offset;
    for (size_t i = 0; i < pp.rounds.size(); i++) {
      const RoundSpec& rs = pp.rounds[i];
      double* rp = base + i * TILE_ROUND_REALS;
      for (int k = 0; k < TILE_ROUND_REALS; k++) rp[k] = 0.0;
      int forms = 0;
      for (int s = 0; s < 5; s++) {
        if (rs.ebit[s] < 0) continue;
        const int q = rs.ebit[s] - h->qshift;
        double beta = beta_of(h, ang, rs.layer, q);
        // Symmetric register + plain X mixer: U_M(beta + pi/2) = i^n X^(xn) U_M(beta), and the global flip X^(xn) acts
        // trivially on a state with psi(x) = psi(~x): energies have period pi/2 in beta.  Reduce beta to [-pi/4, pi/4]
        // (|tan| <= 1: no cot form -- the reference's default grids hit beta = pi/2 exactly --, least scaling).
        if (pl.z2 && h->mixer == MIXER_X) beta -= 1.5707963267948966 * nearbyint(beta / 1.5707963267948966);
        const double c = cos(beta), sn = sin(beta);
        if (layer_fits[rs.layer] || fabs(sn) <= tmax * fabs(c)) {   // normal form: c * [[1, t], [-t, 1]]
          const double t = sn / c;
          rp[s] = t;
          rp[5 + s] = -t;
          logS[rs.layer] += log(fabs(c));
          if (c < 0) sgnS[rs.layer] = -sgnS[rs.layer];
        } else {                            // cot form: s * [[u, 1], [-1, u]]
          const double u = c / sn;
          rp[s] = u;
          rp[5 + s] = u;
          forms |= 1 << s;
          logS[rs.layer] += log(fabs(sn));
          if (sn < 0) sgnS[rs.layer] = -sgnS[rs.layer];
          if (pass_all_normal) (*pass_all_normal)[ip] = 0;
        }
      }
      rp[10] = (double)forms;
    }
  }
  // split every layer's scale S_d into pre_d (applied with the layer's own phase) and post_d
  std::vector<double> pre(p), post(p);
  for (int d = 0; d < p; d++) {
    pre[d] = exp(0.5 * logS[d]);
    post[d] = sgnS[d] * exp(0.5 * logS[d]);
  }
  auto mult_of = [&](int d) { return (d == 0 ? 1.0 : post[d - 1]) * pre[d]; };
  for (const auto& pp : pl.passes) {
    double* ph = phs + pp.phs_offset;
    for (size_t j = 0; j < pp.phases.size(); j++) {
      const int d = pp.phases[j].layer;
      ph[2 * j] = gamma_of(h, ang, d);
      ph[2 * j + 1] = mult_of(d);
    }
    ph[2 * pp.phases.size()] = post[p - 1];
  }
  if (gamsig)
    for (int d = 0; d < p; d++) {
      gamsig[2 * d] = gamma_of(h, ang, d);
      gamsig[2 * d + 1] = mult_of(d);
    }
  if (final_scale_out) *final_scale_out = post[p - 1];
}

// ---- column kernels (column_kernel.cuh): tensor map of the state for a high tile shape --------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled tensor_map_encoder() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || !p)
      fail("cuTensorMapEncodeTiled is not available");
    fn = (PFN_encodeTiled)p;
  }
  return fn;
}

// high tile shape = element bits 0..4 + nine CONSECUTIVE bits hp..hp+8 (complex64: element = 8 bytes).  The register of
// `amps` elements is a 5-D tensor [16 (bits 0..3)][2 (bit 4)][2^(hp-5)][512 rows][amps >> (hp+9)]; a half tile is two
// boxes of 256 rows x 128 bytes, SWIZZLE_128B.
const CUtensorMap& column_tensor_map(Engine* h, void* state, int hp, size_t amps) {
  auto key = std::make_pair((const void*)state, hp * 64 + (int)(63 - __builtin_clzll((unsigned long long)amps)));
  auto it = h->tmaps.find(key);
  if (it != h->tmaps.end()) return it->second;
  CUtensorMap m;
  const cuuint64_t dims[5] = {16, 2, (cuuint64_t)1 << (hp - 5), 512, (cuuint64_t)(amps >> (hp + 9))};
  const cuuint64_t strides[4] = {128, 256, (cuuint64_t)8 << hp, (cuuint64_t)8 << (hp + 9)};   // bytes, dims 1..4
  const cuuint32_t box[5] = {16, 1, 1, 256, 1};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult rc = tensor_map_encoder()(&m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 5, state, dims, strides, box, estr,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) fail("cuTensorMapEncodeTiled failed (" + std::to_string((int)rc) + ")");
  h->tmaps[key] = m;
  return h->tmaps[key];
}

// can this pass run on the column kernels?
bool column_pass_ok(const Engine* h, const PlanPass& pp, int prog, int B) {
  if (!h->column_kernels || h->dtype != QAOA_DTYPE_C64 || h->dinfo.kind != DIAG_U8 || B != 1) return false;
  if (prog != PROG_BOUND2 && prog != PROG_FINAL2) return false;
  if (pp.cross || pp.zshape) return false;
  const TilePass& tp = pp.tp;
  for (int j = 0; j < 5; j++) if (tp.pos[j] != j) return false;
  for (int j = 6; j < TILE_BITS; j++) if (tp.pos[j] != tp.pos[5] + (j - 5)) return false;
  return tp.pos[5] >= 5;
}

// returns the number of partial-result slots per point a REDUCE pass writes
template <typename Elem>
int launch_pass(Engine* h, const PlanPass& pp, bool use_prog, unsigned ntiles, int B, size_t point_stride_elems,
                const void* d_tau, const double* d_phs, int tables_per_point) {
  typedef typename ElemTraits<Elem>::real real;
  TilePass tp = pp.tp;
  tp.ntiles = ntiles;
  if (ntiles & (ntiles - 1)) fail("the number of tiles must be a power of two");
  tilepass_set_offsets(tp);
  tp.prefetch_dist = h->prefetch_dist;
  shard_fill(h, tp, pp.cross);
  cudaStream_t lstream = h->cur_stream ? h->cur_stream : h->stream;
  const int lctas = h->cur_ctas > 0 ? h->cur_ctas : h->persistent_ctas;
  int prog = use_prog ? pp.prog : PROG_GENERIC;
  const bool cross = pp.cross;
  // cross passes have compiled programs for the high-tile passes of the balanced schedule only
  if (cross && prog != PROG_BOUNDX && prog != PROG_FINALX) prog = PROG_GENERIC;
  if (!cross && (prog == PROG_BOUNDX || prog == PROG_FINALX)) prog = PROG_GENERIC;
  const bool u8 = h->dinfo.kind == DIAG_U8;
  const bool zs = pp.zshape;   // low tile of a symmetric register: mirrored addressing (kernel mode 2)
  if (zs)
    for (int i = 0; i < tp.nops; i++)
      if (tp.ops[i].type != TOP_ROUND && tp.ops[i].frame >= 3) fail("symmetric register: the low tile has no A4 / C4 frames");
  int nparts = (int)ntiles;
  // interpreter: one CTA per (tile, point); programs: persistent CTAs walking the linear index
  auto go_generic = [&]() {
    auto kernel = zs ? tile_pass_kernel<Elem, 2> : (cross ? tile_pass_kernel<Elem, 1> : tile_pass_kernel<Elem, 0>);
    CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_SMEM_BYTES));
    tp.prefetch_dist = 0;
    // interpreter walks tiles of ONE point per launch row; use a flat grid over (point, tile)
    dim3 grid((unsigned)std::min<unsigned long long>((unsigned long long)ntiles * B, 0x7fffffffull));
    kernel<<<grid, TILE_THREADS, TILE_SMEM_BYTES, lstream>>>(tp, (Elem*)h->state, point_stride_elems, (const real*)d_tau,
                                                             d_phs, h->d_tables, tables_per_point, h->dinfo, h->d_partials);
  };
  auto go = [&](auto kernel) {
    CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TileSmem<Elem>::TOTAL));
    const unsigned long long total = (unsigned long long)ntiles * B;
    dim3 grid((unsigned)std::min<unsigned long long>(total, (unsigned long long)lctas));
    nparts = (int)grid.x;
    if (tp.ch_m > 0) {   // chunked launch: every chunk owns `grid` consecutive partial-result slots
      tp.ch_part0 = tp.ch_c * (int)grid.x;
      nparts = (int)grid.x << tp.ch_m;
    }
    kernel<<<grid, TILE_THREADS, TileSmem<Elem>::TOTAL, lstream>>>(tp, (Elem*)h->state, point_stride_elems,
                                                                      (const real*)d_tau, d_phs, h->d_tables,
                                                                      tables_per_point, h->dinfo, h->d_partials, (unsigned)B);
  };
  if constexpr (ElemTraits<Elem>::IS_C64) {
    if (column_pass_ok(h, pp, prog, B)) {
      const int hp = tp.pos[5];
      const CUtensorMap& tm = column_tensor_map(h, h->state, hp, point_stride_elems);
      ColGeom G;
      G.mid_bits = hp - 5;
      G.tiles = ntiles;
      G.tile0 = 0;
      dim3 grid((unsigned)std::min<unsigned long long>(ntiles, (unsigned long long)lctas));
      nparts = (int)grid.x;
      if (tp.ch_m > 0) { tp.ch_part0 = tp.ch_c * (int)grid.x; nparts = (int)grid.x << tp.ch_m; }
      if (h->column_sync & 1) CUDA_OK(cudaStreamSynchronize(lstream));
      if (prog == PROG_BOUND2) {
        CUDA_OK(cudaFuncSetAttribute(column_prog_kernel<ProgBound2>, cudaFuncAttributeMaxDynamicSharedMemorySize, COL_SMEM_BYTES));
        column_prog_kernel<ProgBound2><<<grid, TILE_THREADS, COL_SMEM_BYTES, lstream>>>(tm, tp, G, (const float*)d_tau, d_phs, h->d_tables,
                                                                                       tables_per_point, h->d_partials);
      } else {
        CUDA_OK(cudaFuncSetAttribute(column_prog_kernel<ProgFinal2>, cudaFuncAttributeMaxDynamicSharedMemorySize, COL_SMEM_BYTES));
        column_prog_kernel<ProgFinal2><<<grid, TILE_THREADS, COL_SMEM_BYTES, lstream>>>(tm, tp, G, (const float*)d_tau, d_phs, h->d_tables,
                                                                                       tables_per_point, h->d_partials);
      }
      CUDA_OK(cudaGetLastError());
      if (h->column_sync & 2) CUDA_OK(cudaStreamSynchronize(lstream));
      return nparts;
    }
  }
  if (zs) {
    switch (prog) {
      case PROG_FIRST_INIT: if (u8) go(tile_prog_kernel<Elem, ProgFirstInit, DIAG_U8, 2>); else go(tile_prog_kernel<Elem, ProgFirstInit, 0, 2>); break;
      case PROG_INTERIOR: go(tile_prog_kernel<Elem, ProgInterior, DIAG_U8, 2>); break;
      case PROG_BOUNDARY: if (u8) go(tile_prog_kernel<Elem, ProgBoundary, DIAG_U8, 2>); else go(tile_prog_kernel<Elem, ProgBoundary, 0, 2>); break;
      case PROG_FINAL: if (u8) go(tile_prog_kernel<Elem, ProgFinal, DIAG_U8, 2>); else go(tile_prog_kernel<Elem, ProgFinal, 0, 2>); break;
      case PROG_MID: go(tile_prog_kernel<Elem, ProgMid, DIAG_U8, 2>); break;
      case PROG_FINAL2: if (u8) go(tile_prog_kernel<Elem, ProgFinal2, DIAG_U8, 2>); else go(tile_prog_kernel<Elem, ProgFinal2, 0, 2>); break;
      default: go_generic(); break;
    }
    CUDA_OK(cudaGetLastError());
    return nparts;
  }
  switch (prog) {
    case PROG_FIRST_INIT: if (u8) go(tile_prog_kernel<Elem, ProgFirstInit, DIAG_U8, false>); else go(tile_prog_kernel<Elem, ProgFirstInit, 0, false>); break;
    case PROG_FIRST_LOAD: if (u8) go(tile_prog_kernel<Elem, ProgFirstLoad, DIAG_U8, false>); else go(tile_prog_kernel<Elem, ProgFirstLoad, 0, false>); break;
    case PROG_INTERIOR: go(tile_prog_kernel<Elem, ProgInterior, DIAG_U8, false>); break;
    case PROG_BOUNDARY: if (u8) go(tile_prog_kernel<Elem, ProgBoundary, DIAG_U8, false>); else go(tile_prog_kernel<Elem, ProgBoundary, 0, false>); break;
    case PROG_FINAL: if (u8) go(tile_prog_kernel<Elem, ProgFinal, DIAG_U8, false>); else go(tile_prog_kernel<Elem, ProgFinal, 0, false>); break;
    case PROG_MID: go(tile_prog_kernel<Elem, ProgMid, DIAG_U8, false>); break;
    case PROG_BOUND2: if (u8) go(tile_prog_kernel<Elem, ProgBound2, DIAG_U8, false>); else go(tile_prog_kernel<Elem, ProgBound2, 0, false>); break;
    case PROG_FINAL2: if (u8) go(tile_prog_kernel<Elem, ProgFinal2, DIAG_U8, false>); else go(tile_prog_kernel<Elem, ProgFinal2, 0, false>); break;
    case PROG_BOUND2L4: if (u8) go(tile_prog_kernel<Elem, ProgBound2L4, DIAG_U8, false>); else go(tile_prog_kernel<Elem, ProgBound2L4, 0, false>); break;
    case PROG_FINAL2L4: if (u8) go(tile_prog_kernel<Elem, ProgFinal2L4, DIAG_U8, false>); else go(tile_prog_kernel<Elem, ProgFinal2L4, 0, false>); break;
    case PROG_BOUNDX: if (u8) go(tile_prog_kernel<Elem, ProgBoundX, DIAG_U8, true>); else go(tile_prog_kernel<Elem, ProgBoundX, 0, true>); break;
    case PROG_FINALX: if (u8) go(tile_prog_kernel<Elem, ProgFinalX, DIAG_U8, true>); else go(tile_prog_kernel<Elem, ProgFinalX, 0, true>); break;
    default: go_generic(); break;
  }
  CUDA_OK(cudaGetLastError());
  return nparts;
}

// tile path for a chunk of points [b0, b0+Bc)
// ---- tile path, executed pass by pass -----------------------------------------------------
// tile_prepare lowers the angles of Bc points and uploads them (+ phase tables); tile_launch_pass
// launches pass ip; tile_finalize turns the partial sums into {E, E2, min, max, norm} per point.
void tile_prepare(Engine* h, Plan& pl, const double* angles, int n_angles, int Bc, bool record_last) {
  const int p = pl.p;
  const unsigned ntiles = plan_ntiles(h, pl);
  const bool c64 = h->dtype == QAOA_DTYPE_C64;
  const size_t rsz = c64 ? 4 : 8;
  const size_t tpp = pl.tau_per_point, ppp = pl.phs_per_point;
  // host block: [tau (real) pass-major][pad][phs (double) pass-major][gamsig]
  const size_t tau_bytes = (tpp * Bc * rsz + 15) / 16 * 16;
  const size_t phs_bytes = ppp * Bc * 8;
  const size_t gs_bytes = (size_t)2 * p * Bc * 8;
  // two pinned blocks used in turn; a block is rewritten only after the upload that last read it has completed
  h->param_slot ^= 1;
  const int slot = h->param_slot;
  if (!h->ev_params[slot]) CUDA_OK(cudaEventCreateWithFlags(&h->ev_params[slot], cudaEventDisableTiming));
  else CUDA_OK(cudaEventSynchronize(h->ev_params[slot]));
  double*& hpar = slot ? h->h_params2 : h->h_params;
  size_t& hcap = slot ? h->hparams2_cap : h->hparams_cap;
  ensure(hpar, hcap, tau_bytes + phs_bytes + gs_bytes, true);
  ensure(h->d_params, h->params_cap, tau_bytes + phs_bytes);
  ensure(h->d_gamsig, h->gamsig_cap, gs_bytes);
  ensure(h->d_partials, h->partials_cap, ((size_t)ntiles + h->persistent_ctas) * Bc * 5 * sizeof(double));
  unsigned char* hb = (unsigned char*)hpar;
  double* h_phs = (double*)(hb + tau_bytes);
  double* hgs = (double*)(hb + tau_bytes + phs_bytes);
  std::vector<double> tau_tmp(tpp), phs_tmp(ppp);
  std::vector<char> an;
  Engine::EvalCtx& ev = h->ev;
  ev.pl = &pl;
  ev.Bc = Bc;
  ev.tau_bytes = tau_bytes;
  ev.all_normal.assign(pl.passes.size(), 1);
  ev.tbase.assign(pl.passes.size(), 0);
  ev.pbase.assign(pl.passes.size(), 0);
  ev.nparts = (int)ntiles;
  for (int b = 0; b < Bc; b++) {
    double fsc = 1.0;
    fill_point_params(h, pl, angles + (size_t)b * n_angles, tau_tmp.data(), phs_tmp.data(), hgs + (size_t)2 * p * b,
                      &an, &fsc);
    size_t tbase = 0, pbase = 0;
    for (size_t ip = 0; ip < pl.passes.size(); ip++) {
      const auto& pp = pl.passes[ip];
      if (!an[ip]) ev.all_normal[ip] = 0;
      const size_t ts = pp.tp.tau_stride, ps = pp.tp.phs_stride;
      if (c64) {
        float* dst = (float*)hb + tbase + (size_t)b * ts;
        for (size_t k = 0; k < ts; k++) dst[k] = (float)tau_tmp[pp.tau_offset + k];
      } else {
        memcpy((double*)hb + tbase + (size_t)b * ts, tau_tmp.data() + pp.tau_offset, ts * 8);
      }
      memcpy(h_phs + pbase + (size_t)b * ps, phs_tmp.data() + pp.phs_offset, ps * 8);
      ev.tbase[ip] = tbase;
      ev.pbase[ip] = pbase;
      tbase += ts * Bc;
      pbase += ps * Bc;
    }
    if (record_last && b == Bc - 1) {
      h->last_z.assign(h->n, 0); h->last_gsign = 1; h->last_scale = fsc; h->last_valid = true; h->last_path = 1;
    }
  }
  unsigned char* db = (unsigned char*)h->d_params;
  CUDA_OK(cudaMemcpyAsync(db, hb, tau_bytes + phs_bytes, cudaMemcpyHostToDevice, h->stream));
  if (h->dinfo.kind == DIAG_U8) {
    CUDA_OK(cudaMemcpyAsync(h->d_gamsig, hgs, gs_bytes, cudaMemcpyHostToDevice, h->stream));
    const size_t tb = (size_t)p * Bc * 256 * (c64 ? 8 : 16);
    ensure(h->d_tables, h->tables_cap, tb);
    if (c64) build_tables_kernel<float, float2><<<p * Bc, 256, 0, h->stream>>>(h->d_gamsig, h->dinfo.c0, h->dinfo.delta, (float2*)h->d_tables);
    else build_tables_kernel<double, double2><<<p * Bc, 256, 0, h->stream>>>(h->d_gamsig, h->dinfo.c0, h->dinfo.delta, (double2*)h->d_tables);
    CUDA_OK(cudaGetLastError());
    h->launches += 1;
  }
  CUDA_OK(cudaEventRecord(h->ev_params[slot], h->stream));     // the uploads out of this pinned block are behind us
  if (h->init.kind == INIT_STATEVECTOR) {
    for (int b = 0; b < Bc; b++)
      CUDA_OK(cudaMemcpyAsync((char*)h->state + (size_t)b * h->N * h->amp_bytes, h->init_vec, h->N * h->amp_bytes,
                              cudaMemcpyDeviceToDevice, h->stream));
  }
}

// position of element bit e in the tile index of a shape (tiles enumerate the shape's free bits in ascending order)
static int tile_index_pos(const Shape& sh, int e) {
  int k = 0;
  for (int f : sh.freepos) { if (f == e) return k; if (f < e) k++; }
  return -1;
}

// fills the chunk fields of a pass' TilePass for chunk c (plan_chunks chose the bits)
void set_chunk_fields(const Plan& pl, PlanPass& pp, int c) {
  const Shape& sh = pl.shapes[pp.shape_index];
  TilePass& tp = pp.tp;
  tp.ch_m = pl.ch_m; tp.ch_c = c; tp.ch_part0 = 0;
  tp.ch_pu = tile_index_pos(sh, pl.ch_u);
  for (int i = 0; i < pl.ch_m; i++) tp.ch_pv[i] = tile_index_pos(sh, pl.ch_v[i]);
  if (tp.ch_pu < 0) fail("chunk pipeline: chunk bit is not a free bit of the shape");
  for (int i = 0; i < pl.ch_m; i++) if (tp.ch_pv[i] < 0) fail("chunk pipeline: chunk bit is not a free bit of the shape");
}

// Chunk pipeline of sharded registers.  The cross passes are NVLink-bound and need few SMs (24 CTAs keep the links
// busy, scripts/cross_ctas.py), the local passes are HBM / issue bound; they can run side by side if they touch
// disjoint data.  Pick three local element bits u, v0, v1 that exactly ONE local shape mixes (the "open" shape), that
// are free bits of every other shape and below the bits that deal the cross tiles out over the ranks: the four sets
// {z : z[v0]^z[u] = c0, z[v1]^z[u] = c1} are closed under every other pass -- the low tile's mirror pairing included,
// which complements all three bits -- so chunk c's cross pass may run while chunk c+1 is still in its local passes.
void plan_chunks(Engine* h, Plan& pl) {
  pl.ch_m = 0;
  if (h->world < 2 || !h->chunking) return;
  const int EL = plan_el(h, pl);
  const int m = (int)pl.shapes.size();
  int cross_shape = -1;
  for (int i = 0; i < m; i++) if (pl.shapes[i].cross) cross_shape = i;
  if (cross_shape < 0) return;
  // bits that deal the cross tiles out over the ranks: the top gbits free bits of the cross shape
  const auto& xf = pl.shapes[cross_shape].freepos;
  std::vector<int> owner(EL, -1), count(EL, 0);
  for (int i = 0; i < m; i++)
    for (int j = 0; j < TILE_BITS; j++) {
      const int e = pl.shapes[i].pos[j];
      if (e >= 0 && e < EL) { count[e]++; owner[e] = i; }
    }
  std::vector<std::vector<int>> excl(m);
  for (int e = h->qshift; e < EL; e++) {
    if (count[e] != 1) continue;
    const int o = owner[e];
    if (pl.shapes[o].cross || pl.shapes[o].zshape) continue;
    const int px = tile_index_pos(pl.shapes[cross_shape], e);
    if (px < 0 || px >= (int)xf.size() - h->gbits) continue;
    bool free_everywhere = true;
    for (int i = 0; i < m; i++) if (i != o && tile_index_pos(pl.shapes[i], e) < 0) free_everywhere = false;
    if (free_everywhere) excl[o].push_back(e);
  }
  // the open shape: preferably the local BOUNDARY shape (its passes carry a phase step and cannot move inside their
  // layer anyway; the phase-free passes of the other shapes can be placed next to the cross passes)
  int best = -1;
  for (const auto& pp : pl.passes)
    if (!pp.cross && !pp.zshape && !pp.phases.empty() && pp.tp.ops[0].type == TOP_LOAD && excl[pp.shape_index].size() >= 2) {
      best = pp.shape_index;
      break;
    }
  if (best < 0)
    for (int i = 0; i < m; i++) if (excl[i].size() >= 2 && (best < 0 || excl[i].size() > excl[best].size())) best = i;
  if (best < 0) return;
  // every pass over another shape must be a compiled-program pass candidate; passes over the open shape stay whole
  const auto& b = excl[best];
  pl.ch_open_shape = best;
  pl.ch_m = std::min({(int)b.size() - 1, 3, std::max(1, h->chunk_bits)});
  // (every chunk launch must still give each persistent CTA a few tiles)
  while (pl.ch_m > 1 && (plan_ntiles(h, pl) >> pl.ch_m) < 4u * (unsigned)h->persistent_ctas) pl.ch_m--;
  pl.ch_u = b[b.size() - 1];
  for (int i = 0; i < 3; i++) pl.ch_v[i] = -1;
  for (int i = 0; i < pl.ch_m; i++) pl.ch_v[i] = b[b.size() - 1 - pl.ch_m + i];   // ascending tile-index positions
  // chunked launches exist for the compiled-program kernels only (launch_pass sends a cross pass without a BOUNDX /
  // FINALX program, or any pass without a program, to the interpreter kernel, which walks whole passes)
  for (auto& pp : pl.passes) {
    const bool xprog = pp.prog == PROG_BOUNDX || pp.prog == PROG_FINALX;
    pp.chunk_closed = pp.shape_index != best && pp.prog != PROG_GENERIC && (pp.cross ? xprog : !xprog);
  }
  // layer 0: the open shape's pass goes BEFORE the closed pass that precedes the first cross pass (same layer, no
  // phase in between: the two commute), so that a closed pass sits right before the cross pass and can be pipelined
  for (size_t i = 0; i + 2 < pl.passes.size(); i++) {
    auto &a = pl.passes[i], &o = pl.passes[i + 1], &x = pl.passes[i + 2];
    if (!x.cross || a.cross || o.cross || !a.chunk_closed || o.chunk_closed) continue;
    if (!a.phases.empty() || !o.phases.empty() || a.has_reduce || o.has_reduce) continue;
    if (a.tp.ops[0].type != TOP_LOAD || o.tp.ops[0].type != TOP_LOAD) continue;
    bool same = true;
    const int L = a.rounds.empty() ? -1 : a.rounds[0].layer;
    for (auto& r : a.rounds) if (r.layer != L) same = false;
    for (auto& r : o.rounds) if (r.layer != L) same = false;
    if (same && L >= 0) std::swap(pl.passes[i], pl.passes[i + 1]);
    break;
  }
}

void tile_launch_pass(Engine* h, size_t ip) {
  Engine::EvalCtx& ev = h->ev;
  if (!ev.pl || ip >= ev.pl->passes.size()) fail("no evaluation in flight / pass index out of range");
  Plan& pl = *ev.pl;
  const int p = pl.p, Bc = ev.Bc;
  const unsigned ntiles = plan_ntiles(h, pl);
  const bool c64 = h->dtype == QAOA_DTYPE_C64;
  const size_t rsz = c64 ? 4 : 8;
  unsigned char* db = (unsigned char*)h->d_params;
  const size_t amps = plan_amps(h, pl);
  const size_t point_stride_elems = amps << h->qshift;
  auto& pp = pl.passes[ip];
  const bool use_prog = ev.all_normal[ip] && !h->force_generic;
  const void* d_tau = db + ev.tbase[ip] * rsz;
  const double* d_phs = (const double*)(db + ev.tau_bytes) + ev.pbase[ip];
  bool reads = pp.tp.ops[0].type == TOP_LOAD, writes = pp.tp.ops[pp.tp.nops - 1].type == TOP_STORE;
  const int chm = h->cur_ch_c >= 0 ? pl.ch_m : 0;
  if (chm > 0) {
    if (!pp.chunk_closed || !use_prog) fail("chunked launch of a pass that couples chunks / needs the interpreter kernel");
    set_chunk_fields(pl, pp, h->cur_ch_c);
  } else pp.tp.ch_m = 0;
  cudaStream_t lstream = h->cur_stream ? h->cur_stream : h->stream;
  const double pass_bytes = ((double)Bc * (double)amps * h->amp_bytes * ((reads ? 1 : 0) + (writes ? 1 : 0)) +
      (double)Bc * (double)amps * diag_value_bytes(h->dinfo.kind) * (pp.phases.size() + (pp.has_reduce ? 1 : 0))) / (double)(1 << chm);
  size_t evi = 0;
  if (h->time_passes) {
    evi = h->ev_pending.size();
    while (h->ev_pool.size() <= evi) {
      cudaEvent_t a, b;
      CUDA_OK(cudaEventCreate(&a));
      CUDA_OK(cudaEventCreate(&b));
      h->ev_pool.push_back({a, b});
    }
    CUDA_OK(cudaEventRecord(h->ev_pool[evi].first, lstream));
  }
  const int np = c64 ? launch_pass<float2>(h, pp, use_prog, ntiles >> chm, Bc, point_stride_elems, d_tau, d_phs, p)
                     : launch_pass<double>(h, pp, use_prog, ntiles >> chm, Bc, point_stride_elems, d_tau, d_phs, p);
  if (pp.has_reduce) ev.nparts = np;
  if (h->time_passes) {
    CUDA_OK(cudaEventRecord(h->ev_pool[evi].second, lstream));
    h->ev_pending.push_back({use_prog ? pp.prog : (int)PROG_GENERIC, evi, pass_bytes, h->cur_ch_c,
                             (h->cur_stream && h->cur_stream == h->stream2) ? 1 : 0});
  }
  h->launches += 1;
  if (use_prog && pp.prog != PROG_GENERIC) h->prog_launches += 1;
  h->bytes_alg += pass_bytes;
}

void tile_finalize(Engine* h, double* d_out_chunk, bool raw) {
  Engine::EvalCtx& ev = h->ev;
  if (raw) finalize_sums_kernel<<<ev.Bc, 256, 0, h->stream>>>(h->d_partials, ev.nparts, h->dinfo, d_out_chunk);
  else finalize_kernel<<<ev.Bc, 256, 0, h->stream>>>(h->d_partials, ev.nparts, h->dinfo, d_out_chunk);
  CUDA_OK(cudaGetLastError());
  h->launches += 1;
}

// tile path for a chunk of points [b0, b0+Bc)
void run_tile_chunk(Engine* h, Plan& pl, const double* angles, int n_angles, int Bc, double* d_out_chunk,
                    bool record_last) {
  tile_prepare(h, pl, angles, n_angles, Bc, record_last);
  for (size_t ip = 0; ip < pl.passes.size(); ip++) tile_launch_pass(h, ip);
  tile_finalize(h, d_out_chunk, false);
}

// small fused path for a chunk
void run_small_chunk(Engine* h, int p, const double* angles, int n_angles, int Bc, double* d_out_chunk) {
  ensure(h->d_angles, h->angles_cap, (size_t)Bc * n_angles * sizeof(double));
  CUDA_OK(cudaMemcpyAsync(h->d_angles, angles, (size_t)Bc * n_angles * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  SmallDesc D;
  D.n = h->n; D.p = p; D.mixer = h->mixer;
  D.n_beta = h->mixer == MIXER_XMULTI ? h->n : 1;
  D.n_gamma = h->n_gamma;
  D.n_init = h->n_init;
  D.terms = h->d_terms;
  D.n_pairs = (int)h->pairs.size() / 2;
  D.trotter = h->trotter;
  D.write_state = h->keep_state;
  D.init = h->init; D.sub = h->sub;
  D.node = h->node;
  if (h->keep_state) ensure_state(h, Bc);
  const int smem = (int)(h->N * h->amp_bytes);
  if (h->dtype == QAOA_DTYPE_C64) {
    CUDA_OK(cudaFuncSetAttribute(small_fused_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, std::max(smem, 1024)));
    small_fused_kernel<float><<<Bc, 512, smem, h->stream>>>(D, h->d_pairs, h->d_angles, n_angles, h->diag, h->dinfo,
                                                            (float*)h->state, d_out_chunk);
  } else {
    CUDA_OK(cudaFuncSetAttribute(small_fused_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, std::max(smem, 1024)));
    small_fused_kernel<double><<<Bc, 512, smem, h->stream>>>(D, h->d_pairs, h->d_angles, n_angles, h->diag, h->dinfo,
                                                             (double*)h->state, d_out_chunk);
  }
  CUDA_OK(cudaGetLastError());
  h->launches += 1;
  h->last_valid = true; h->last_path = 0; h->last_scale = 1.0; h->last_gsign = 1;
  h->last_z.assign(h->n, 0);
}

// launches grover_pass_c64_kernel<flags, diagonal kind, generator kind>; false if the combination is not instantiated
template <int FLAGS>
bool launch_grover_fast_f(Engine* h, int grid, float2* st, size_t Nv, const StateGen& g, double gamma, const void* diagv,
                          double* part_dot, double* part_red) {
#define QAOA_GF(DK, GK)                                                                                              \
  if (h->dinfo.kind == DK && g.kind == GK) {                                                                         \
    grover_pass_c64_kernel<FLAGS, DK, GK><<<grid, 256, 0, h->stream>>>(st, Nv, g, h->d_coefdot, gamma, diagv, h->dinfo, \
                                                                       part_dot, part_red);                          \
    return true;                                                                                                     \
  }
  QAOA_GF(DIAG_U8, INIT_PLUS) QAOA_GF(DIAG_U8, INIT_DICKE) QAOA_GF(DIAG_U8, INIT_UNIFORM)
  QAOA_GF(DIAG_U32FX, INIT_PLUS) QAOA_GF(DIAG_U32FX, INIT_DICKE) QAOA_GF(DIAG_U32FX, INIT_UNIFORM)
#undef QAOA_GF
  return false;
}
bool launch_grover_fast(Engine* h, int grid, float2* st, size_t Nv, int flags, const StateGen& g, double gamma,
                        const void* diagv, double* part_dot, double* part_red) {
  switch (flags) {
    case EW_INIT | EW_PHASE | EW_DOT | EW_STORE:
      return launch_grover_fast_f<EW_INIT | EW_PHASE | EW_DOT | EW_STORE>(h, grid, st, Nv, g, gamma, diagv, part_dot, part_red);
    case EW_AXPY | EW_PHASE | EW_DOT | EW_STORE:
      return launch_grover_fast_f<EW_AXPY | EW_PHASE | EW_DOT | EW_STORE>(h, grid, st, Nv, g, gamma, diagv, part_dot, part_red);
    case EW_AXPY | EW_REDUCE:
      return launch_grover_fast_f<EW_AXPY | EW_REDUCE>(h, grid, st, Nv, g, gamma, diagv, part_dot, part_red);
    case EW_AXPY | EW_REDUCE | EW_STORE:
      return launch_grover_fast_f<EW_AXPY | EW_REDUCE | EW_STORE>(h, grid, st, Nv, g, gamma, diagv, part_dot, part_red);
    default:
      return false;
  }
}

template <typename real>
void run_elementwise_point(Engine* h, int p, const double* ang, double* d_out_point) {
  // feasible-subspace registers (Dicke + Grover over the same Dicke state, Dicke + XY) need C(n,k) amplitudes only
  const bool xy_tiled_size = h->n >= (h->dtype == QAOA_DTYPE_C64 ? 13 : 12);
  const bool want_compact = h->init.kind == INIT_DICKE &&
      ((h->mixer == MIXER_GROVER && h->sub.kind == INIT_DICKE && h->init.k == h->sub.k) || (h->mixer == MIXER_XY && xy_tiled_size));
  const bool compact = want_compact && ensure_compact(h, h->init.k) && (h->mixer != MIXER_XY || ensure_compact_strings(h));
  ensure_state(h, 1, compact ? h->n_compact : h->N);
  const int threads = 256;
  const int grid = grid_for(h->N, threads);
  ensure(h->d_partials, h->partials_cap, (size_t)grid * 5 * sizeof(double) * 2);
  if (!h->d_coefdot) CUDA_OK(cudaMalloc((void**)&h->d_coefdot, 2 * sizeof(double)));
  double* part_dot = h->d_partials;
  double* part_red = h->d_partials + (size_t)grid * 5;
  int grid_red = grid;
  real* st = (real*)h->state;
  // Dicke initial state + Grover mixer over the same Dicke state: the register lives in the C(n,k)-dimensional
  // feasible subspace (compact index = rank of the weight-k string); every entry of |s> is the same amplitude
  size_t Nv = h->N;
  const void* diagv = h->diag;
  StateGen initv = h->init, subv = h->sub;
  if (h->mixer == MIXER_GROVER && compact) {
    Nv = h->n_compact;
    diagv = h->diag_compact;
    initv.kind = INIT_UNIFORM;
    subv.kind = INIT_UNIFORM;
  }
  const double sweep = (double)Nv * h->amp_bytes;
  const double dsz = (double)Nv * diag_value_bytes(h->dinfo.kind);
  // complex64 Grover over its own initial state (|+>, |D^n_k>, compact): the specialised passes (kernels_misc.cuh)
  const bool fast_grover = sizeof(real) == 4 && h->mixer == MIXER_GROVER && !h->force_generic && initv.kind == subv.kind &&
      initv.k == subv.k && initv.amp == subv.amp &&
      (initv.kind == INIT_PLUS || initv.kind == INIT_DICKE || initv.kind == INIT_UNIFORM) &&
      (h->dinfo.kind == DIAG_U8 || h->dinfo.kind == DIAG_U32FX) && (Nv >= 2);
  auto ew = [&](int flags, double gamma) {
    if (!(fast_grover && launch_grover_fast(h, grid, (float2*)st, Nv, flags, initv, gamma, diagv, part_dot, part_red)))
      elementwise_pass_kernel<real><<<grid, threads, 0, h->stream>>>(st, Nv, flags, initv, subv, h->d_coefdot, gamma,
                                                                    diagv, h->dinfo, part_dot, part_red);
    CUDA_OK(cudaGetLastError());
    h->launches += 1;
    h->bytes_alg += ((flags & EW_INIT) ? 0 : sweep) + ((flags & EW_STORE) ? sweep : 0) +
                    ((flags & (EW_PHASE | EW_REDUCE)) ? dsz : 0);
  };
  if (h->mixer == MIXER_GROVER) {
    for (int d = 0; d < p; d++) {
      int fl = (d == 0 ? EW_INIT : EW_AXPY) | EW_PHASE | EW_DOT | EW_STORE;
      ew(fl, ang[2 * d]);
      grover_dot_finalize_kernel<<<1, 1024, 0, h->stream>>>(part_dot, grid, ang[2 * d + 1], h->d_coefdot);
      CUDA_OK(cudaGetLastError());
      h->launches += 1;
    }
    ew(EW_AXPY | EW_REDUCE | (h->keep_state ? EW_STORE : 0), 0.0);
  } else if (h->mixer == MIXER_NODE_GROVER) {   // tensorised Grover mixer: one sweep per node register
    const int b = h->node.bits;
    for (int d = 0; d < p; d++) {
      ew((d == 0 ? EW_INIT : 0) | EW_PHASE | EW_STORE, ang[2 * d]);
      const double beta = ang[2 * d + 1];
      const int gg = grid_for(h->N >> b, threads);
      for (int v = 0; v < h->n / b; v++) {
        node_grover_kernel<real><<<gg, threads, 0, h->stream>>>(st, h->N, v * b, h->node, cos(beta) - 1.0, -sin(beta));
        CUDA_OK(cudaGetLastError());
        h->launches += 1;
        h->bytes_alg += 2 * sweep;
      }
    }
    ew(EW_REDUCE, 0.0);
  } else if (h->n < (h->dtype == QAOA_DTYPE_C64 ? 13 : 12)) {  // XY on a register smaller than a tile: one sweep per gate
    const int np = (int)h->pairs.size() / 2;
    for (int d = 0; d < p; d++) {
      ew((d == 0 ? EW_INIT : 0) | EW_PHASE | EW_STORE, ang[2 * d]);
      const double a = 0.25 * ang[2 * d + 1] / h->trotter;
      const int g4 = grid_for(h->N / 4, threads);
      for (int rep = 0; rep < h->trotter; rep++)
        for (int e = 0; e < np; e++) {
          xy_gate_kernel<real><<<g4, threads, 0, h->stream>>>(st, h->N, h->pairs[2 * e], h->pairs[2 * e + 1], cos(a), sin(a));
          CUDA_OK(cudaGetLastError());
          h->launches += 1;
          h->bytes_alg += sweep;
        }
    }
    ew(EW_REDUCE, 0.0);
  } else if (h->mixer == MIXER_XY && compact) {
    // XY + Dicke in the feasible subspace: C(n,k) entries instead of 2^n, one sweep per gate over the compact register
    const size_t M = h->n_compact;
    const int np = (int)h->pairs.size() / 2;
    StateGen uni = h->init;
    uni.kind = INIT_UNIFORM;
    const int cgrid = grid_for(M, threads);
    const double csweep = (double)M * h->amp_bytes, cdsz = (double)M * diag_value_bytes(h->dinfo.kind);
    const double sbytes = h->n <= 32 ? 4.0 : 8.0;
    auto cew = [&](int flags, double gamma) {
      elementwise_pass_kernel<real><<<cgrid, threads, 0, h->stream>>>(st, M, flags, uni, uni, h->d_coefdot, gamma,
                                                                     h->diag_compact, h->dinfo, part_dot, part_red);
      CUDA_OK(cudaGetLastError());
      h->launches += 1;
      h->bytes_alg += ((flags & EW_INIT) ? 0 : csweep) + ((flags & EW_STORE) ? csweep : 0) + cdsz;
    };
    // fraction of the entries a gate touches: strings with x[q0] != x[q1]
    const double touched = h->n > 1 ? 2.0 * h->init.k * (h->n - h->init.k) / ((double)h->n * (h->n - 1)) : 0.0;
    const int smem = (h->n + 1) * (h->init.k + 1) * 8;
    typedef typename Real2<real>::type real2;
    for (int d = 0; d < p; d++) {
      cew((d == 0 ? EW_INIT : 0) | EW_PHASE | EW_STORE, ang[2 * d]);
      const double a = 0.25 * ang[2 * d + 1] / h->trotter;
      for (int rep = 0; rep < h->trotter; rep++)
        for (int e = 0; e < np; e++) {
          const int q0 = h->pairs[2 * e], q1 = h->pairs[2 * e + 1];
          if (h->n <= 32)
            xy_compact_gate_kernel<real, unsigned><<<cgrid, threads, smem, h->stream>>>(
                (real2*)st, (const unsigned*)h->strings_compact, M, q0, q1, (real)cos(a), (real)sin(a), h->binom_compact, h->n, h->init.k);
          else
            xy_compact_gate_kernel<real, unsigned long long><<<cgrid, threads, smem, h->stream>>>(
                (real2*)st, (const unsigned long long*)h->strings_compact, M, q0, q1, (real)cos(a), (real)sin(a), h->binom_compact,
                h->n, h->init.k);
          CUDA_OK(cudaGetLastError());
          h->launches += 1;
          h->bytes_alg += (double)M * sbytes + 2.0 * touched * csweep;
        }
    }
    cew(EW_REDUCE, 0.0);
    grid_red = cgrid;
  } else {  // XY: ordered gate list, tiled (see xy_tile_kernel)
    const int np = (int)h->pairs.size() / 2;
    const bool c64 = h->dtype == QAOA_DTYPE_C64;
    const int tb = c64 ? 13 : 12, clow = c64 ? 5 : 4;
    std::vector<std::pair<int, int>> gates;
    for (int rep = 0; rep < h->trotter; rep++)
      for (int e = 0; e < np; e++) gates.push_back({h->pairs[2 * e], h->pairs[2 * e + 1]});
    // greedy cover: the longest run of consecutive gates whose qubits fit into one tile
    std::vector<XYTilePass> layer;
    size_t idx = 0;
    while (idx < gates.size()) {
      std::vector<int> extra;
      auto in_tile = [&](int q) { return q < clow || std::find(extra.begin(), extra.end(), q) != extra.end(); };
      std::vector<std::pair<int, int>> mine;
      while (idx < gates.size() && (int)mine.size() < XY_MAX_GATES) {
        std::vector<int> need;
        if (!in_tile(gates[idx].first)) need.push_back(gates[idx].first);
        if (!in_tile(gates[idx].second) && gates[idx].second != gates[idx].first) need.push_back(gates[idx].second);
        if ((int)(extra.size() + need.size()) > tb - clow) break;
        for (int q : need) extra.push_back(q);
        mine.push_back(gates[idx]);
        idx++;
      }
      if (mine.empty()) fail("XY planner: a gate does not fit into a tile");
      for (int q = clow; (int)extra.size() < tb - clow && q < h->n; q++) if (!in_tile(q)) extra.push_back(q);
      XYTilePass P;
      memset(&P, 0, sizeof(P));
      P.tb = tb;
      std::vector<int> bits;
      for (int q = 0; q < clow; q++) bits.push_back(q);
      for (int q : extra) bits.push_back(q);
      std::sort(bits.begin(), bits.end());
      for (int k = 0; k < tb; k++) P.pos[k] = bits[k];
      for (int q = 0; q < h->n; q++) if (!std::binary_search(bits.begin(), bits.end(), q)) P.freepos[P.nfree++] = q;
      P.ngates = (int)mine.size();
      for (int k = 0; k < P.ngates; k++) {
        P.gi[k] = (unsigned char)(std::lower_bound(bits.begin(), bits.end(), mine[k].first) - bits.begin());
        P.gj[k] = (unsigned char)(std::lower_bound(bits.begin(), bits.end(), mine[k].second) - bits.begin());
      }
      layer.push_back(P);
    }
    const size_t ntiles = h->N >> tb;
    const int smem = (int)(((size_t)1 << tb) * h->amp_bytes);
    auto kernel = xy_tile_kernel<real, sizeof(real) == 4 ? 13 : 12>;
    CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int xgrid = (int)std::min<size_t>(ntiles, (size_t)h->persistent_ctas * 2);   // 64 KB + 64 registers: two CTAs per SM
    ensure(h->d_partials, h->partials_cap, (size_t)std::max(grid, xgrid) * 5 * sizeof(double) * 2);
    part_red = h->d_partials;
    for (int d = 0; d < p; d++) {
      const double a = 0.25 * ang[2 * d + 1] / h->trotter;
      for (size_t k = 0; k < layer.size(); k++) {
        XYTilePass P = layer[k];
        const bool last = d == p - 1 && k + 1 == layer.size();
        P.flags = (k == 0 ? EW_PHASE : 0) | (d == 0 && k == 0 ? EW_INIT : 0) | (last ? EW_REDUCE : EW_STORE) |
                  (last && h->keep_state ? EW_STORE : 0);
        P.gamma = ang[2 * d];
        P.cd = cos(a);
        P.sd = sin(a);
        kernel<<<xgrid, 512, smem, h->stream>>>(P, st, ntiles, h->init, h->diag, h->dinfo, part_red);
        CUDA_OK(cudaGetLastError());
        h->launches += 1;
        h->bytes_alg += ((P.flags & EW_INIT) ? 0 : sweep) + ((P.flags & EW_STORE) ? sweep : 0) +
                        ((P.flags & (EW_PHASE | EW_REDUCE)) ? dsz : 0);
      }
    }
    grid_red = xgrid;
  }
  finalize_raw_kernel<<<1, 1024, 0, h->stream>>>(part_red, grid_red, d_out_point);
  CUDA_OK(cudaGetLastError());
  h->launches += 1;
  h->last_valid = true; h->last_path = 0; h->last_scale = 1.0; h->last_gsign = 1;
  h->last_z.assign(h->n, 0);
}

// Grover mixer on a sharded register (no state exchange at all -- one complex scalar per layer crosses the
// ranks).  The caller steps through the circuit: step d < p runs layer d's sweep over the local shard
//   v = (d == 0 ? |s> : v + (e^{-i beta_{d-1}} - 1) <s|v>_global s) ; v *= e^{i gamma_d c} ; partial <s|v>
// and returns the LOCAL partial dot product in out[0..1]; the caller all-reduces it and hands the GLOBAL value back as
// `dot_global` of the next step.  Step p applies the last reflection and returns the raw local statistics
// {sum p, sum p c, sum p c^2, min, max} in out[0..4] (combined like the tile path's, qaoa_shard_finish).
template <typename real>
void shard_grover_step(Engine* h, int step, int p, const double* ang, const double* dot_global, double* out) {
  ensure_state(h, 1);
  const int threads = 256;
  const int grid = grid_for(h->N, threads);
  ensure(h->d_partials, h->partials_cap, (size_t)grid * 5 * sizeof(double) * 2);
  if (!h->d_coefdot) CUDA_OK(cudaMalloc((void**)&h->d_coefdot, 2 * sizeof(double)));
  double* part_dot = h->d_partials;
  double* part_red = h->d_partials + (size_t)grid * 5;
  real* st = (real*)h->state;
  if (step > 0) {
    const double beta = ang[2 * (step - 1) + 1];
    const double kr = cos(beta) - 1.0, ki = -sin(beta);
    const double cd[2] = {kr * dot_global[0] - ki * dot_global[1], kr * dot_global[1] + ki * dot_global[0]};
    h2d(h, h->d_coefdot, cd, sizeof(cd));
  }
  const bool last = step == p;
  const int flags = last ? (EW_AXPY | EW_REDUCE | (h->keep_state ? EW_STORE : 0))
                         : ((step == 0 ? EW_INIT : EW_AXPY) | EW_PHASE | EW_DOT | EW_STORE);
  if (last) { h->last_valid = true; h->last_path = 0; h->last_scale = 1.0; h->last_gsign = 1; h->last_z.assign(h->n, 0); }
  const double gamma = last ? 0.0 : ang[2 * step];
  const bool fast = sizeof(real) == 4 && !h->force_generic && h->init.kind == h->sub.kind && h->init.k == h->sub.k &&
      h->init.amp == h->sub.amp && (h->init.kind == INIT_PLUS || h->init.kind == INIT_DICKE) &&
      (h->dinfo.kind == DIAG_U8 || h->dinfo.kind == DIAG_U32FX);
  if (!(fast && launch_grover_fast(h, grid, (float2*)st, h->N, flags, h->init, gamma, h->diag, part_dot, part_red)))
    elementwise_pass_kernel<real><<<grid, threads, 0, h->stream>>>(st, h->N, flags, h->init, h->sub, h->d_coefdot, gamma,
                                                                  h->diag, h->dinfo, part_dot, part_red);
  CUDA_OK(cudaGetLastError());
  h->launches += 1;
  const double sweep = (double)h->N * h->amp_bytes;
  h->bytes_alg += (step == 0 ? 0.0 : sweep) + (last ? 0.0 : sweep) + (double)h->N * diag_value_bytes(h->dinfo.kind);
  if (!last) {
    DevTmp t_sum(2 * 8);
    sum_partials_kernel<<<1, 32, 0, h->stream>>>(part_dot, grid, 2, t_sum.as<double>());
    CUDA_OK(cudaGetLastError());
    d2h(h, out, t_sum.as<double>(), 2 * 8);
  } else {
    std::vector<double> part((size_t)grid * 5);
    d2h(h, part.data(), part_red, part.size() * 8);
    out[0] = out[1] = out[2] = 0.0; out[3] = 1e300; out[4] = -1e300;
    for (int b = 0; b < grid; b++) {
      out[0] += part[(size_t)b * 5]; out[1] += part[(size_t)b * 5 + 1]; out[2] += part[(size_t)b * 5 + 2];
      out[3] = std::min(out[3], part[(size_t)b * 5 + 3]); out[4] = std::max(out[4], part[(size_t)b * 5 + 4]);
    }
  }
}

void check_ready(Engine* h, int p, int n_angles) {
  if (h->dinfo.kind == DIAG_NONE) fail("no cost function set");
  if (p < 1) fail("depth p must be >= 1");
  const int nb = h->mixer == MIXER_XMULTI ? h->n : 1;
  if (n_angles != h->n_init + p * (h->n_gamma + nb)) {
    char b[200];
    snprintf(b, sizeof(b), "expected %d angles for p=%d (n_init=%d, n_gamma=%d, n_beta=%d), got %d",
             h->n_init + p * (h->n_gamma + nb), p, h->n_init, h->n_gamma, nb, n_angles);
    fail(b);
  }
  if ((h->n_init > 0 || h->n_gamma > 1) && h->n > h->small_max()) {
    if (h->mixer != MIXER_X && h->mixer != MIXER_XMULTI)
      fail("per-term gammas / PlusParameterized on registers beyond one CTA: X and multi-angle X mixers only");
    if (h->world > 1) fail("sharded engine: per-term gammas / PlusParameterized are not supported");
  }
  if ((h->mixer == MIXER_XY) && h->pairs.empty()) fail("XY mixer without pairs");
}

void energy_batch_body(Engine* h, int p, const double* angles, int n_angles, int B);

// collects the per-pass event timings once the stream has drained
void harvest_events(Engine* h) {
  if (!h->ev_pending.empty()) h->timeline.clear();
  for (auto& pe : h->ev_pending) {
    float ms = 0;
    float a = 0, b = 0;
    const cudaEvent_t origin = h->ev_pool[h->ev_pending[0].idx].first;
    if (cudaEventElapsedTime(&a, origin, h->ev_pool[pe.idx].first) == cudaSuccess &&
        cudaEventElapsedTime(&b, origin, h->ev_pool[pe.idx].second) == cudaSuccess)
      h->timeline.push_back({pe.prog, pe.chunk, pe.side, a, b});
    if (cudaEventElapsedTime(&ms, h->ev_pool[pe.idx].first, h->ev_pool[pe.idx].second) == cudaSuccess) {
      h->pass_ms[pe.prog] += ms;
      h->pass_count[pe.prog] += 1;
      h->pass_bytes[pe.prog] += pe.bytes;
    }
  }
  h->ev_pending.clear();
  float ms = 0;
  if (h->ev_call_recorded) {
    if (cudaEventElapsedTime(&ms, h->ev_call0, h->ev_call1) == cudaSuccess) h->last_device_ms = ms;
    else cudaGetLastError();
    h->ev_call_recorded = false;
  }
}

void energy_batch_impl(Engine* h, int p, const double* angles, int n_angles, int B) {
  if (h->world > 1) fail("sharded engine: use qaoa_shard_begin / qaoa_shard_run_pass / qaoa_shard_finish");
  check_ready(h, p, n_angles);
  CUDA_OK(cudaSetDevice(h->device));
  if (!h->ev_call0) { CUDA_OK(cudaEventCreate(&h->ev_call0)); CUDA_OK(cudaEventCreate(&h->ev_call1)); }
  // make sure plans / streams exist before the timed region starts
  if (h->n > h->small_max() && (h->mixer == MIXER_X || h->mixer == MIXER_XMULTI) && h->n_gamma == 1 && h->n_init == 0)
    get_plan(h, p);
  CUDA_OK(cudaEventRecord(h->ev_call0, h->stream));
  energy_batch_body(h, p, angles, n_angles, B);
  CUDA_OK(cudaEventRecord(h->ev_call1, h->stream));
  h->ev_call_recorded = true;
}

// Per-term gammas (MaxCutFree / MaxCutOrbit) and PlusParameterized on registers beyond one CTA, X-type mixers: the
// circuit runs layer by layer.  layer_phase_kernel writes the segment's initial state (kept state of the previous
// segment x the layer's edge phases; first layer: generated state x RZ phases x edge phases), the tile plan of depth 1
// applies the mixer with gamma = 0 and keeps its final state; the last segment's statistics are the circuit's.  With
// one gamma per layer (PlusParameterized alone) the prepared state is followed by the ordinary depth-p plan.
void run_layered_point(Engine* h, int p, const double* ang, double* d_out_point) {
  const int ng = h->n_gamma, ni = h->n_init;
  const int nb = h->mixer == MIXER_XMULTI ? h->n : 1;
  const int per = ng + nb;
  const int ne = (int)h->term_u.size();
  if (ng > 1 && ne == 0) fail("per-term gammas without an edge list");
  if (!h->layer_vec) CUDA_OK(cudaMalloc(&h->layer_vec, h->N * h->amp_bytes));
  if (ni > 0) {
    if (!h->d_init_phi) CUDA_OK(cudaMalloc((void**)&h->d_init_phi, QAOA_MAX_QUBITS * sizeof(double)));
    h2d(h, h->d_init_phi, ang, ni * sizeof(double));
  }
  if (ng > 1 && !h->d_layer_w) {
    CUDA_OK(cudaMalloc((void**)&h->d_layer_w, ne * sizeof(double)));
    CUDA_OK(cudaMalloc((void**)&h->d_term_u, ne * sizeof(int)));
    CUDA_OK(cudaMalloc((void**)&h->d_term_v, ne * sizeof(int)));
    h2d(h, h->d_term_u, h->term_u.data(), ne * sizeof(int));
    h2d(h, h->d_term_v, h->term_v.data(), ne * sizeof(int));
  }
  MaxKCutDev D;
  memset(&D, 0, sizeof(D));
  D.bits = h->term_bits; D.n_free = h->term_n_free; D.fixed_colour = h->term_fixed_colour;
  for (int i = 0; i < 8; i++) D.lut[i] = h->term_lut[i];
  D.eu = h->d_term_u; D.ev = h->d_term_v; D.ew = h->d_layer_w; D.ewi = nullptr;
  const StateGen gen0 = h->init;
  void* const vec0 = h->init_vec;
  const int keep0 = h->keep_state;
  const int grid = grid_for(h->N, 256);
  const double sweep = (double)h->N * h->amp_bytes;
  auto prepare = [&](int d) {       // d < 0: initial state + RZ phases only
    D.n_edges = 0;
    if (d >= 0 && ng > 1) {
      std::vector<double> g(ne);
      for (int e = 0; e < ne; e++) g[e] = ang[ni + d * per + h->term_id[e]] * h->term_w[e];
      h2d(h, h->d_layer_w, g.data(), ne * sizeof(double));
      D.n_edges = ne;
    }
    const bool first = d <= 0;
    if (h->dtype == QAOA_DTYPE_C64)
      layer_phase_kernel<float><<<grid, 256, 0, h->stream>>>(first ? nullptr : (const float*)h->state, (float*)h->layer_vec, h->N, gen0,
                                                           (float)(h->last_scale * h->last_gsign), D, first ? ni : 0, h->d_init_phi);
    else
      layer_phase_kernel<double><<<grid, 256, 0, h->stream>>>(first ? nullptr : (const double*)h->state, (double*)h->layer_vec, h->N, gen0,
                                                            h->last_scale * h->last_gsign, D, first ? ni : 0, h->d_init_phi);
    CUDA_OK(cudaGetLastError());
    h->launches += 1;
    h->bytes_alg += (first ? 1.0 : 2.0) * sweep;
  };
  struct Restore {
    Engine* h; StateGen g; void* v; int keep, ng, ni;
    ~Restore() { h->init = g; h->init_vec = v; h->keep_state = keep; h->n_gamma = ng; h->n_init = ni; h->plans.clear(); }
  } restore{h, gen0, vec0, keep0, ng, ni};
  h->plans.clear();
  h->init.kind = INIT_STATEVECTOR; h->init.k = 0; h->init.amp = 1.0; h->init.vec = h->layer_vec;
  h->init_vec = h->layer_vec;
  h->n_gamma = 1; h->n_init = 0;
  if (ng == 1) {   // PlusParameterized with an ordinary cost function: prepared state, then the depth-p plan
    prepare(-1);
    Plan& pl = get_plan(h, p);
    ensure_state(h, 1, plan_amps(h, pl));
    run_tile_chunk(h, pl, ang + ni, p * per, 1, d_out_point, true);
    CUDA_OK(cudaStreamSynchronize(h->stream));
    return;
  }
  std::vector<double> a1(1 + nb);
  for (int d = 0; d < p; d++) {
    h->keep_state = (d + 1 < p) ? 1 : keep0;
    h->plans.clear();
    prepare(d);
    a1[0] = 0.0;
    for (int q = 0; q < nb; q++) a1[1 + q] = ang[ni + d * per + ng + q];
    Plan& pl = get_plan(h, 1);
    ensure_state(h, 1, plan_amps(h, pl));
    run_tile_chunk(h, pl, a1.data(), 1 + nb, 1, d_out_point, true);
    CUDA_OK(cudaStreamSynchronize(h->stream));   // pinned parameter buffer reuse
  }
}

void energy_batch_body(Engine* h, int p, const double* angles, int n_angles, int B) {
  ensure(h->d_out, h->out_cap, (size_t)B * 5 * sizeof(double));
  const bool small = h->n <= h->small_max();
  if (small) {
    int done = 0;
    while (done < B) {
      int Bc = std::min(B - done, h->keep_state ? 1 : 65535);
      run_small_chunk(h, p, angles + (size_t)done * n_angles, n_angles, Bc, h->d_out + (size_t)done * 5);
      if (Bc < B - done) CUDA_OK(cudaStreamSynchronize(h->stream));  // d_angles reuse
      done += Bc;
    }
    return;
  }
  if ((h->mixer == MIXER_X || h->mixer == MIXER_XMULTI) && (h->n_gamma > 1 || h->n_init > 0)) {
    for (int b = 0; b < B; b++) run_layered_point(h, p, angles + (size_t)b * n_angles, h->d_out + (size_t)b * 5);
    return;
  }
  if (h->mixer == MIXER_X || h->mixer == MIXER_XMULTI) {
    Plan& pl = get_plan(h, p);
    size_t per_point = plan_amps(h, pl) * h->amp_bytes;
    int Bmax = h->batch_points > 0 ? h->batch_points : (int)std::max<size_t>(1, std::min<size_t>(256, h->batch_bytes / per_point));
    Bmax = std::min(Bmax, 65535);
    ensure_state(h, std::min(B, Bmax), plan_amps(h, pl));
    int done = 0;
    while (done < B) {
      int Bc = std::min(B - done, Bmax);
      run_tile_chunk(h, pl, angles + (size_t)done * n_angles, n_angles, Bc, h->d_out + (size_t)done * 5, done + Bc == B);
      done += Bc;     // (no synchronise: tile_prepare alternates two pinned parameter blocks guarded by events)
    }
    return;
  }
  for (int b = 0; b < B; b++) {
    if (h->dtype == QAOA_DTYPE_C64) run_elementwise_point<float>(h, p, angles + (size_t)b * n_angles, h->d_out + (size_t)b * 5);
    else run_elementwise_point<double>(h, p, angles + (size_t)b * n_angles, h->d_out + (size_t)b * 5);
  }
}

void drop_term_edges(Engine* h) {
  h->term_u.clear(); h->term_v.clear(); h->term_id.clear(); h->term_w.clear();
  if (h->d_term_u) { cudaFree(h->d_term_u); h->d_term_u = nullptr; }
  if (h->d_term_v) { cudaFree(h->d_term_v); h->d_term_v = nullptr; }
  if (h->d_layer_w) { cudaFree(h->d_layer_w); h->d_layer_w = nullptr; }
}

void set_diag_common(Engine* h, int kind, double c0, double delta) {
  close_peers(h, 1);
  drop_compact(h);
  if (h->d_terms) { cudaFree(h->d_terms); h->d_terms = nullptr; }
  drop_term_edges(h);
  h->n_gamma = 1;
  if (h->diag) { cudaFree(h->diag); h->diag = nullptr; }
  free_streams(h);
  invalidate(h);
  h->cost_symmetric = -1;
  h->dinfo.kind = kind;
  h->dinfo.c0 = kind == DIAG_F64 ? 0.0 : c0;
  h->dinfo.delta = kind == DIAG_F64 ? 1.0 : delta;
  CUDA_OK(cudaMalloc(&h->diag, h->N * diag_value_bytes(kind)));
}

// choose the diagonal representation from value bounds
void choose_kind(Engine* h, double lo, double hi, bool integer_valued, int* kind, double* c0, double* delta) {
  if (integer_valued && hi - lo <= 255.0) { *kind = DIAG_U8; *c0 = lo; *delta = 1.0; return; }
  if (h->dtype == QAOA_DTYPE_C128) { *kind = DIAG_F64; *c0 = 0; *delta = 1; return; }
  *kind = DIAG_U32FX;
  *c0 = lo;
  if (integer_valued && hi - lo < 4294967295.0) *delta = 1.0;
  else *delta = (hi > lo) ? (hi - lo) / 4294967295.0 : 1.0;
}

StateGen make_gen(Engine* h, int kind, int k, const double* vec, void** devvec) {
  StateGen g;
  g.kind = kind; g.k = k; g.vec = nullptr; g.amp = 1.0;
  g.pc0 = h->world > 1 ? __builtin_popcountll((unsigned long long)h->x0) : 0;
  if (*devvec) { cudaFree(*devvec); *devvec = nullptr; }
  if (kind == INIT_PLUS) g.amp = pow(2.0, -0.5 * h->n);
  else if (kind == INIT_DICKE) {
    if (k < 0 || k > h->n) fail("Dicke weight out of range");
    double lc = lgamma(h->n + 1.0) - lgamma(k + 1.0) - lgamma(h->n - k + 1.0);
    g.amp = exp(-0.5 * lc);
  } else if (kind == INIT_STATEVECTOR) {
    if (!vec) fail("statevector pointer is null");
    CUDA_OK(cudaMalloc(devvec, h->N * h->amp_bytes));
    if (h->dtype == QAOA_DTYPE_C128) {
      h2d(h, *devvec, vec, h->N * 16);
      to_sframe_kernel<double><<<grid_for(h->N, 256), 256, 0, h->stream>>>((double*)*devvec, h->N);
    } else {
      std::vector<float> tmp(h->N * 2);
      for (size_t i = 0; i < h->N * 2; i++) tmp[i] = (float)vec[i];
      h2d(h, *devvec, tmp.data(), h->N * 8);
      to_sframe_kernel<float><<<grid_for(h->N, 256), 256, 0, h->stream>>>((float*)*devvec, h->N);
    }
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(h->stream));
    g.vec = *devvec;
  } else fail("unknown state kind");
  return g;
}

}  // namespace

struct qaoa_engine : Engine {};

#define API_BEGIN(h)            \
  if (!(h)) return 1;           \
  try {                         \
    cudaSetDevice((h)->device);
#define API_END(h)                  \
  }                                 \
  catch (const std::string& e) {    \
    (h)->err = e;                   \
    return 2;                       \
  }                                 \
  catch (...) {                     \
    (h)->err = "unknown error";     \
    return 3;                       \
  }                                 \
  return 0;

extern "C" {

const char* qaoa_global_error(void) { return g_global_error.c_str(); }

int qaoa_create(int n_qubits, int dtype, int device, qaoa_engine_t** out) {
  return qaoa_create_sharded(n_qubits, dtype, device, 0, 1, out);
}

int qaoa_create_sharded(int n_qubits, int dtype, int device, int rank, int world, qaoa_engine_t** out) {
  if (!out) return 1;
  *out = nullptr;
  try {
    if (n_qubits < 1 || n_qubits > QAOA_MAX_QUBITS - 1) fail("n_qubits out of range [1, 39]");
    int gbits = 0;
    while ((1 << gbits) < world) gbits++;
    if (world < 1 || world > QAOA_MAX_RANKS || (1 << gbits) != world) fail("world must be 1, 2, 4 or 8");
    if (rank < 0 || rank >= world) fail("rank out of range");
    if (world > 1 && n_qubits - gbits + (dtype == QAOA_DTYPE_C128 ? 1 : 0) < TILE_BITS + 1)
      fail("sharded engine: every shard must hold more than 14 element bits");
    if (dtype != QAOA_DTYPE_C64 && dtype != QAOA_DTYPE_C128) fail("dtype must be 0 (complex64) or 1 (complex128)");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) fail(std::string("no CUDA device available: ") + cudaGetErrorString(e));
    if (device < 0 || device >= ndev) fail("device index out of range");
    CUDA_OK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) fail("this engine requires an sm_100a (B200) device");
    qaoa_engine* h = new qaoa_engine();
    h->persistent_ctas = prop.multiProcessorCount;
    h->n = n_qubits; h->dtype = dtype; h->device = device;
    h->qshift = dtype == QAOA_DTYPE_C128 ? 1 : 0;
    h->E = n_qubits + h->qshift;
    h->rank = rank; h->world = world; h->gbits = gbits;
    h->El = h->E - gbits;
    h->N = (size_t)1 << (n_qubits - gbits);
    h->x0 = (size_t)rank << (n_qubits - gbits);
    h->amp_bytes = dtype == QAOA_DTYPE_C128 ? 16 : 8;
    CUDA_OK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->init = make_gen(h, INIT_PLUS, 0, nullptr, &h->init_vec);
    h->sub = h->init;
    *out = h;
  } catch (const std::string& e) {
    g_global_error = e;
    return 2;
  }
  return 0;
}

int qaoa_destroy(qaoa_engine_t* h) {
  if (!h) return 1;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  close_peers(h, 0);
  close_peers(h, 1);
  free_streams(h);
  void* bufs[] = {h->state, h->diag, h->diag_compact, h->strings_compact, h->binom_compact, h->d_term_u, h->d_term_v, h->d_layer_w,
                  h->d_init_phi, h->layer_vec, h->d_terms, h->d_pairs, h->init_vec, h->sub_vec, h->d_params, h->d_partials, h->d_out,
                  h->d_tables, h->d_gamsig, h->d_coefdot, h->d_angles};
  for (void* b : bufs) if (b) cudaFree(b);
  if (h->h_params) cudaFreeHost(h->h_params);
  if (h->h_params2) cudaFreeHost(h->h_params2);
  for (auto& e : h->ev_params) if (e) cudaEventDestroy(e);
  if (h->h_out) cudaFreeHost(h->h_out);
  cudaStreamDestroy(h->stream);
  if (h->stream2) cudaStreamDestroy(h->stream2);
  delete h;
  return 0;
}

const char* qaoa_last_error(qaoa_engine_t* h) { return h ? h->err.c_str() : "null handle"; }

int qaoa_set_option(qaoa_engine_t* h, const char* key, double value) {
  API_BEGIN(h)
  std::string k(key ? key : "");
  if (k == "small_max_qubits") {
    int lim = h->dtype == QAOA_DTYPE_C64 ? 14 : 13;
    h->small_max_qubits = std::min((int)value, lim);
  } else if (k == "keep_state") { h->keep_state = value != 0; invalidate(h); }
  else if (k == "batch_points") h->batch_points = (int)value;
  else if (k == "batch_bytes") h->batch_bytes = (size_t)value;
  else if (k == "force_generic") h->force_generic = value != 0;
  else if (k == "layer_budget") h->layer_budget = value != 0;
  else if (k == "column_kernels") h->column_kernels = value != 0;
  else if (k == "column_sync") h->column_sync = (int)value;
  else if (k == "chunk_pipeline") { h->chunking = value != 0; invalidate(h); }
  else if (k == "chunk_bits") { h->chunk_bits = std::max(1, std::min(3, (int)value)); invalidate(h); }
  else if (k == "prefetch_dist") h->prefetch_dist = (int)value;
  else if (k == "persistent_ctas") h->persistent_ctas = std::max(1, (int)value);
  else if (k == "time_passes") h->time_passes = value != 0;
  else if (k == "low_bits") { h->low_bits = (int)value; free_streams(h); invalidate(h); }
  else if (k == "schedule") { h->schedule = (int)value; free_streams(h); invalidate(h); }
  else if (k == "feasible_subspace") h->compact_enabled = value != 0;
  else if (k == "symmetric") { h->z2_enabled = value != 0; invalidate(h); }
  else if (k == "cross_spare") { h->cross_spare = (int)value; free_streams(h); invalidate(h); }   // 0 off, 1 auto, 10 + k: force k
  else if (k == "shard_symmetric") {
    // Sharded symmetric register: to be set right after qaoa_create_sharded (before the state buffer is exported
    // and before the cost function is set).  The engine then holds 2^(n-1-gbits) stored amplitudes.
    if ((value != 0) != (h->ztw_mode != 0)) {
      if (value == 0) fail("shard_symmetric cannot be switched off again");
      if (h->world < 2) fail("shard_symmetric applies to sharded engines (unsharded engines choose the symmetric register themselves)");
      if (h->dtype != QAOA_DTYPE_C64) fail("shard_symmetric: complex64 only");
      if (h->n & 1) fail("shard_symmetric: the number of qubits must be even");
      if (h->El - 1 < TILE_BITS + 1) fail("shard_symmetric: every shard of the stored half must hold more than 14 qubits");
      if (h->peers_state_set || h->peers_diag_set || h->dinfo.kind != DIAG_NONE) fail("shard_symmetric must be set before peers / cost function");
      if (h->state) { cudaFree(h->state); h->state = nullptr; h->state_bytes = 0; }
      h->ztw_mode = 1;
      h->N >>= 1;
      h->El -= 1;
      h->x0 = (size_t)h->rank << h->El;
      ensure_state(h, 1);
      invalidate(h);
    }
  }
  else fail("unknown option: " + k);
  API_END(h)
}

double qaoa_get_counter(qaoa_engine_t* h, const char* key) {
  if (!h || !key) return -1.0;
  std::string k(key);
  if (k == "launches") return h->launches;
  if (k == "prog_launches") return h->prog_launches;
  if (k == "bytes_alg") return h->bytes_alg;
  if (k == "diag_kind") return h->dinfo.kind;
  if (k == "diag_c0") return h->dinfo.c0;
  if (k == "diag_delta") return h->dinfo.delta;
  if (k == "device") return h->device;
  if (k == "compact_size") return (double)h->n_compact;
  if (k == "symmetric_plans") { double s = 0; for (auto& kv : h->plans) s += kv.second.z2 ? 1 : 0; return s; }
  if (k == "state_bytes") return (double)h->state_bytes;
  if (k == "cross_spare_bits") { double s = 0; for (auto& kv : h->plans) s = std::max(s, (double)kv.second.kx); return s; }
  if (k == "reset") {
    h->launches = 0; h->prog_launches = 0; h->bytes_alg = 0;
    for (int i = 0; i < PROG_COUNT; i++) { h->pass_ms[i] = 0; h->pass_count[i] = 0; h->pass_bytes[i] = 0; }
    return 0;
  }
  if (k == "last_device_ms") return h->last_device_ms;
  if (k.rfind("pass_ms_", 0) == 0) { int i = atoi(k.c_str() + 8); return (i >= 0 && i < PROG_COUNT) ? h->pass_ms[i] : -1.0; }
  if (k.rfind("pass_count_", 0) == 0) { int i = atoi(k.c_str() + 11); return (i >= 0 && i < PROG_COUNT) ? h->pass_count[i] : -1.0; }
  if (k.rfind("pass_bytes_", 0) == 0) { int i = atoi(k.c_str() + 11); return (i >= 0 && i < PROG_COUNT) ? h->pass_bytes[i] : -1.0; }
  if (k == "n_passes") { double s = 0; for (auto& kv : h->plans) s = (double)kv.second.passes.size(); return s; }
  return -1.0;
}

int qaoa_set_cost_maxkcut(qaoa_engine_t* h, int n_nodes, int n_edges, const int32_t* u, const int32_t* v,
                          const double* w, int bits_per_node, const uint8_t* lut, int fixed_node, int fixed_colour) {
  API_BEGIN(h)
  if (bits_per_node < 1 || bits_per_node > 3) fail("bits_per_node must be 1..3");
  const int n_free = n_nodes - (fixed_node ? 1 : 0);
  if (n_free * bits_per_node != h->n) fail("n_nodes * bits_per_node does not match the engine's qubit count");
  if (n_edges < 0 || (n_edges > 0 && (!u || !v || !w || !lut))) fail("null edge arrays");
  double pos = 0, neg = 0;
  bool integer = true;
  for (int e = 0; e < n_edges; e++) {
    if (u[e] < 0 || u[e] >= n_nodes || v[e] < 0 || v[e] >= n_nodes) fail("edge endpoint out of range");
    if (w[e] >= 0) pos += w[e]; else neg += w[e];
    if (w[e] != floor(w[e]) || fabs(w[e]) > 1e9) integer = false;
  }
  int kind; double c0, delta;
  choose_kind(h, neg, pos, integer, &kind, &c0, &delta);
  set_diag_common(h, kind, c0, delta);
  MaxKCutDev D;
  memset(&D, 0, sizeof(D));
  D.n_edges = n_edges; D.bits = bits_per_node; D.n_free = n_free; D.fixed_colour = fixed_colour;
  for (int i = 0; i < (1 << bits_per_node); i++) D.lut[i] = lut[i];
  int *du, *dv, *dwi; double* dw;
  std::vector<int> wi(std::max(n_edges, 1));
  // the U8 path stores cost - c0 with c0 = sum of the negative weights
  for (int e = 0; e < n_edges; e++) wi[e] = integer ? (int)w[e] : 0;
  DevTmp t_u(std::max(n_edges, 1) * 4), t_v(std::max(n_edges, 1) * 4), t_wi(std::max(n_edges, 1) * 4), t_w(std::max(n_edges, 1) * 8);
  du = t_u.as<int>(); dv = t_v.as<int>(); dwi = t_wi.as<int>(); dw = t_w.as<double>();
  h2d(h, du, u, n_edges * 4);
  h2d(h, dv, v, n_edges * 4);
  h2d(h, dwi, wi.data(), n_edges * 4);
  h2d(h, dw, w, n_edges * 8);
  D.eu = du; D.ev = dv; D.ewi = dwi; D.ew = dw;
  DiagInfo gi = h->dinfo;
  if (h->ztw_mode) {
    // the stored half (top qubit = 0) in twisted coordinates; MaxCut proper is invariant under the global flip
    if (bits_per_node != 1 || fixed_node || lut[0] == lut[1])
      fail("sharded symmetric register: MaxCut (one qubit per node, no fixed node) only");
    gen_maxkcut_kernel<<<grid_for(h->N, 256), 256, 0, h->stream>>>(D, h->N, h->x0, gi, h->diag, h->El, (1u << h->gbits) - 1u);
  } else {
    gen_maxkcut_kernel<<<grid_for(h->N, 256), 256, 0, h->stream>>>(D, h->N, h->x0, gi, h->diag);
  }
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaStreamSynchronize(h->stream));
  h->launches += 1;
  if (h->ztw_mode) h->cost_symmetric = 1;
  API_END(h)
}

int qaoa_set_cost_maxkcut_terms(qaoa_engine_t* h, int n_nodes, int n_edges, const int32_t* u, const int32_t* v,
                                const double* w, const int32_t* term_of_edge, int n_terms, int bits_per_node,
                                const uint8_t* lut, int fixed_node, int fixed_colour) {
  if (!h) return 1;
  // the total cost first (diagonal, bounds, representation), then one double-precision diagonal per term
  int rc = qaoa_set_cost_maxkcut(h, n_nodes, n_edges, u, v, w, bits_per_node, lut, fixed_node, fixed_colour);
  if (rc) return rc;
  API_BEGIN(h)
  if (h->world > 1) fail("sharded engine: per-term gammas are not supported");
  if (n_terms < 1 || !term_of_edge) fail("term_of_edge is null / n_terms < 1");
  for (int e = 0; e < n_edges; e++) if (term_of_edge[e] < 0 || term_of_edge[e] >= n_terms) fail("term index out of range");
  if (n_terms > 1) {   // the edge list itself: registers beyond one CTA evaluate the term phases on the fly (run_layered_point)
    h->term_u.assign(u, u + n_edges); h->term_v.assign(v, v + n_edges); h->term_w.assign(w, w + n_edges);
    h->term_id.assign(term_of_edge, term_of_edge + n_edges);
    h->term_bits = bits_per_node; h->term_n_free = n_nodes - (fixed_node ? 1 : 0); h->term_fixed_colour = fixed_colour;
    for (int i = 0; i < 8; i++) h->term_lut[i] = i < (1 << bits_per_node) ? lut[i] : 0;
    h->n_gamma = n_terms;
    invalidate(h);
  }
  if (n_terms > 1 && h->n <= (h->dtype == QAOA_DTYPE_C64 ? 14 : 13)) {
  CUDA_OK(cudaMalloc((void**)&h->d_terms, (size_t)n_terms * h->N * sizeof(double)));
  MaxKCutDev D;
  memset(&D, 0, sizeof(D));
  D.bits = bits_per_node; D.n_free = n_nodes - (fixed_node ? 1 : 0); D.fixed_colour = fixed_colour;
  for (int i = 0; i < (1 << bits_per_node); i++) D.lut[i] = lut[i];
  int *du, *dv, *dwi; double* dw;
  DevTmp t_u(std::max(n_edges, 1) * 4), t_v(std::max(n_edges, 1) * 4), t_wi(std::max(n_edges, 1) * 4), t_w(std::max(n_edges, 1) * 8);
  du = t_u.as<int>(); dv = t_v.as<int>(); dwi = t_wi.as<int>(); dw = t_w.as<double>();
  DiagInfo gi{DIAG_F64, 0.0, 1.0};
  for (int k = 0; k < n_terms; k++) {
    std::vector<int> su, sv, swi;
    std::vector<double> sw;
    for (int e = 0; e < n_edges; e++) if (term_of_edge[e] == k) { su.push_back(u[e]); sv.push_back(v[e]); sw.push_back(w[e]); swi.push_back(0); }
    const int ne = (int)su.size();
    if (ne) {
      h2d(h, du, su.data(), ne * 4);
      h2d(h, dv, sv.data(), ne * 4);
      h2d(h, dwi, swi.data(), ne * 4);
      h2d(h, dw, sw.data(), ne * 8);
    }
    D.n_edges = ne; D.eu = du; D.ev = dv; D.ewi = dwi; D.ew = dw;
    gen_maxkcut_kernel<<<grid_for(h->N, 256), 256, 0, h->stream>>>(D, h->N, h->x0, gi, h->d_terms + (size_t)k * h->N);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(h->stream));
    h->launches += 1;
  }
  h->n_gamma = n_terms;
  invalidate(h);
  }
  API_END(h)
}

int qaoa_set_cost_qubo(qaoa_engine_t* h, int n, const double* Q, const double* c, double b) {
  API_BEGIN(h)
  if (n != h->n) fail("QUBO size does not match the engine's qubit count");
  if (h->ztw_mode) fail("sharded symmetric register: MaxCut cost functions only");
  if (!Q) fail("Q is null");
  std::vector<double> d(n), M((size_t)n * n, 0.0);
  double lo = b, hi = b;  // bounds of f = x^T Q x + c^T x + b
  bool integer = (b == floor(b));
  for (int i = 0; i < n; i++) {
    d[i] = Q[(size_t)i * n + i] + (c ? c[i] : 0.0);
    if (d[i] > 0) hi += d[i]; else lo += d[i];
    if (d[i] != floor(d[i])) integer = false;
    for (int j = 0; j < i; j++) {
      double m = Q[(size_t)i * n + j] + Q[(size_t)j * n + i];
      M[(size_t)i * n + j] = m;
      if (m > 0) hi += m; else lo += m;
      if (m != floor(m)) integer = false;
    }
  }
  if (fabs(lo) > 1e9 || fabs(hi) > 1e9) integer = false;
  int kind; double c0, delta;
  choose_kind(h, -hi, -lo, integer, &kind, &c0, &delta);
  set_diag_common(h, kind, c0, delta);
  double *dd, *dM;
  DevTmp t_d(n * 8), t_M((size_t)n * n * 8);
  dd = t_d.as<double>(); dM = t_M.as<double>();
  h2d(h, dd, d.data(), n * 8);
  h2d(h, dM, M.data(), (size_t)n * n * 8);
  const int smem = (n + n * n) * 8;
  gen_qubo_kernel<<<grid_for(h->N, 256), 256, smem, h->stream>>>(n, dd, dM, b, h->N, h->x0, h->dinfo, h->diag);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaStreamSynchronize(h->stream));
  h->launches += 1;
  API_END(h)
}

int qaoa_set_cost_diagonal(qaoa_engine_t* h, const double* diag) {
  API_BEGIN(h)
  if (!diag) fail("diag is null");
  if (h->world > 1) fail("sharded engine: host-provided cost diagonals are not supported");
  double lo = 1e300, hi = -1e300;
  bool integer = true;
  for (size_t i = 0; i < h->N; i++) {
    lo = std::min(lo, diag[i]); hi = std::max(hi, diag[i]);
    if (diag[i] != floor(diag[i])) integer = false;
  }
  int kind; double c0, delta;
  choose_kind(h, lo, hi, integer, &kind, &c0, &delta);
  set_diag_common(h, kind, c0, delta);
  const size_t chunk = (size_t)1 << 24;
  DevTmp t_tmp(std::min(chunk, h->N) * 8);
  double* tmp = t_tmp.as<double>();
  for (size_t off = 0; off < h->N; off += chunk) {
    size_t cnt = std::min(chunk, h->N - off);
    h2d(h, tmp, diag + off, cnt * 8);
    quantise_diag_kernel<<<grid_for(cnt, 256), 256, 0, h->stream>>>(tmp, cnt, off, h->dinfo, h->diag);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(h->stream));
    h->launches += 1;
  }
  API_END(h)
}

int qaoa_read_cost_diagonal(qaoa_engine_t* h, size_t offset, size_t count, double* out) {
  API_BEGIN(h)
  if (h->dinfo.kind == DIAG_NONE) fail("no cost function set");
  if (offset + count > h->N) fail("range outside the diagonal");
  const size_t chunk = (size_t)1 << 24;
  DevTmp t_tmp(std::min(chunk, std::max<size_t>(count, 1)) * 8);
  double* tmp = t_tmp.as<double>();
  for (size_t o = 0; o < count; o += chunk) {
    size_t cnt = std::min(chunk, count - o);
    read_diag_kernel<<<grid_for(cnt, 256), 256, 0, h->stream>>>(h->diag, h->dinfo, offset + o, cnt, tmp);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpyAsync(out + o, tmp, cnt * 8, cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    h->launches += 1;
  }
  API_END(h)
}

int qaoa_cost_minmax(qaoa_engine_t* h, double* out_min, double* out_max) {
  API_BEGIN(h)
  if (h->dinfo.kind == DIAG_NONE) fail("no cost function set");
  const int grid = grid_for(h->N, 256);
  ensure(h->d_partials, h->partials_cap, (size_t)grid * 2 * sizeof(double));
  diag_minmax_kernel<<<grid, 256, 0, h->stream>>>(h->diag, h->dinfo, h->N, h->d_partials);
  CUDA_OK(cudaGetLastError());
  std::vector<double> hp((size_t)grid * 2);
  CUDA_OK(cudaMemcpyAsync(hp.data(), h->d_partials, hp.size() * 8, cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  h->launches += 1;
  double lo = 1e300, hi = -1e300;
  for (int i = 0; i < grid; i++) { lo = std::min(lo, hp[2 * i]); hi = std::max(hi, hp[2 * i + 1]); }
  if (out_min) *out_min = lo;
  if (out_max) *out_max = hi;
  API_END(h)
}

int qaoa_set_mixer(qaoa_engine_t* h, int kind, const int32_t* pairs, int n_pairs, int trotter_steps,
                   int sub_kind, int sub_k, const double* sub_vec) {
  API_BEGIN(h)
  if (kind < MIXER_X || kind > MIXER_NODE_GROVER) fail("unknown mixer kind");
  if (h->world > 1 && kind != MIXER_X && kind != MIXER_XMULTI &&
      !(kind == MIXER_GROVER && !h->ztw_mode && (sub_kind == INIT_PLUS || sub_kind == INIT_DICKE)))
    fail("sharded engine: the X and multi-angle X mixers, and (full register) the Grover mixer over |+> / Dicke states, are supported");
  if (trotter_steps < 1) fail("trotter_steps must be >= 1");
  invalidate(h);
  h->mixer = kind;
  h->trotter = trotter_steps;
  h->pairs.clear();
  if (h->d_pairs) { cudaFree(h->d_pairs); h->d_pairs = nullptr; }
  if (kind == MIXER_XY) {
    if (n_pairs < 1 || !pairs) fail("XY mixer needs at least one pair");
    for (int i = 0; i < 2 * n_pairs; i++) {
      if (pairs[i] < 0 || pairs[i] >= h->n) fail("XY pair index out of range");
      h->pairs.push_back(pairs[i]);
    }
    for (int i = 0; i < n_pairs; i++) if (pairs[2 * i] == pairs[2 * i + 1]) fail("XY pair with identical qubits");
    CUDA_OK(cudaMalloc((void**)&h->d_pairs, h->pairs.size() * 4));
    h2d(h, h->d_pairs, h->pairs.data(), h->pairs.size() * 4);
  }
  if (kind == MIXER_GROVER) h->sub = make_gen(h, sub_kind, sub_k, sub_vec, &h->sub_vec);
  if (kind == MIXER_NODE_GROVER) {   // sub_k = qubits per node, sub_vec = 2^sub_k complex amplitudes of |s_node>
    if (sub_k < 1 || sub_k > 3 || h->n % sub_k) fail("node Grover mixer: 1..3 qubits per node, dividing the qubit count");
    if (!sub_vec) fail("node Grover mixer: node state is null");
    double nrm = 0.0;
    for (int j = 0; j < (2 << sub_k); j++) nrm += sub_vec[j] * sub_vec[j];
    if (fabs(nrm - 1.0) > 1e-8) fail("node Grover mixer: node state is not normalised");
    h->node.bits = sub_k;
    for (int j = 0; j < (1 << sub_k); j++) {   // S frame: s~_j = i^popc(j) s_j
      double re = sub_vec[2 * j], im = sub_vec[2 * j + 1];
      const int k4 = __builtin_popcount((unsigned)j) & 3;
      const double r = re, m = im;
      if (k4 == 1) { re = -m; im = r; } else if (k4 == 2) { re = -r; im = -m; } else if (k4 == 3) { re = m; im = -r; }
      h->node.re[j] = re; h->node.im[j] = im;
    }
  }
  API_END(h)
}

int qaoa_set_initial(qaoa_engine_t* h, int kind, int k, const double* vec) {
  API_BEGIN(h)
  if (h->world > 1 && (kind == INIT_STATEVECTOR || kind == 3)) fail("sharded engine: this initial state is not supported");
  invalidate(h);
  h->n_init = 0;
  if (kind == 3) {   // QAOA_INIT_PLUS_PHASES: |+>^n followed by RZ(phi_q) on the first k qubits, phi = leading angles
    if (k < 1 || k > h->n) fail("number of initial-state phases out of range");
    h->init = make_gen(h, INIT_PLUS, 0, nullptr, &h->init_vec);
    h->n_init = k;
  } else
  h->init = make_gen(h, kind, k, vec, &h->init_vec);
  API_END(h)
}

int qaoa_energy_batch_resident(qaoa_engine_t* h, int p, const double* angles, int n_angles, int B) {
  API_BEGIN(h)
  if (B < 1 || !angles) fail("empty batch");
  energy_batch_impl(h, p, angles, n_angles, B);
  CUDA_OK(cudaStreamSynchronize(h->stream));
  harvest_events(h);
  API_END(h)
}

int qaoa_fetch_results(qaoa_engine_t* h, int B, double* out) {
  API_BEGIN(h)
  if (!h->d_out || (size_t)B * 5 * sizeof(double) > h->out_cap) fail("no results to fetch");
  CUDA_OK(cudaMemcpyAsync(out, h->d_out, (size_t)B * 5 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  API_END(h)
}

int qaoa_energy_batch(qaoa_engine_t* h, int p, const double* angles, int n_angles, int B, double* out) {
  API_BEGIN(h)
  if (B < 1 || !angles || !out) fail("empty batch");
  energy_batch_impl(h, p, angles, n_angles, B);
  CUDA_OK(cudaMemcpyAsync(out, h->d_out, (size_t)B * 5 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  harvest_events(h);
  API_END(h)
}

int qaoa_energy(qaoa_engine_t* h, int p, const double* angles, int n_angles, double* out) {
  return qaoa_energy_batch(h, p, angles, n_angles, 1, out);
}

int qaoa_read_statevector(qaoa_engine_t* h, double* out) {
  API_BEGIN(h)
  if (!h->keep_state || !h->last_valid || !h->state) fail("no kept state: set option keep_state=1 and evaluate first");
  const size_t N = h->N;
  std::vector<double> re(N), im(N);
  if (h->dtype == QAOA_DTYPE_C128) {
    std::vector<double> tmp(2 * N);
    d2h(h, tmp.data(), h->state, N * 16);
    for (size_t x = 0; x < N; x++) { re[x] = tmp[2 * x]; im[x] = tmp[2 * x + 1]; }
  } else {
    std::vector<float> tmp(2 * N);
    d2h(h, tmp.data(), h->state, N * 8);
    for (size_t x = 0; x < N; x++) { re[x] = tmp[2 * x]; im[x] = tmp[2 * x + 1]; }
  }
  unsigned long long zm = 0;
  for (int q = 0; q < h->n; q++) if (h->last_z[q]) zm |= 1ull << q;
  for (size_t x = 0; x < N; x++) {
    double r = re[x] * h->last_scale * h->last_gsign, m = im[x] * h->last_scale * h->last_gsign;
    if (__builtin_popcountll(x & zm) & 1) { r = -r; m = -m; }
    int k = (4 - (__builtin_popcountll(x) & 3)) & 3;  // multiply by (-i)^popc = i^(-popc)
    double rr = r, mm = m;
    if (k == 1) { rr = -m; mm = r; } else if (k == 2) { rr = -r; mm = -m; } else if (k == 3) { rr = m; mm = -r; }
    out[2 * x] = rr;
    out[2 * x + 1] = mm;
  }
  API_END(h)
}

// X gates between layers (the reference's flip=True, qaoa/qaoa.py:505-509): the state KEPT by the last evaluation, with X
// applied on the qubits of flip_mask, becomes the initial state -- device to device, as if qaoa_set_initial had been given
// that statevector -- so that the next evaluation continues the circuit with its remaining layers.
int qaoa_continue_from_kept_state(qaoa_engine_t* h, uint64_t flip_mask) {
  API_BEGIN(h)
  if (h->world > 1) fail("sharded engine: X gates between layers are not supported");
  if (!h->keep_state || !h->last_valid || !h->state) fail("no kept state: set option keep_state=1 and evaluate first");
  if (h->n < 64 && (flip_mask >> h->n)) fail("flip mask names qubits beyond the register");
  unsigned long long zm = 0;
  for (int q = 0; q < h->n; q++) if (h->last_z[q]) zm |= 1ull << q;
  const double scale = h->last_scale * h->last_gsign;
  const int rot = (4 - (__builtin_popcountll(flip_mask) & 3)) & 3;    // i^(-popc(m))
  void* dst = nullptr;
  CUDA_OK(cudaMalloc(&dst, h->N * h->amp_bytes));
  const int grid = grid_for(h->N, 256);
  if (h->dtype == QAOA_DTYPE_C64)
    flip_continue_kernel<float><<<grid, 256, 0, h->stream>>>((const float*)h->state, (float*)dst, h->N, flip_mask, zm, (float)scale, rot);
  else
    flip_continue_kernel<double><<<grid, 256, 0, h->stream>>>((const double*)h->state, (double*)dst, h->N, flip_mask, zm, scale, rot);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) { cudaFree(dst); CUDA_OK(e); }
  h->launches += 1;
  if (h->init_vec) cudaFree(h->init_vec);
  h->init_vec = dst;
  h->init.kind = INIT_STATEVECTOR; h->init.k = 0; h->init.amp = 1.0; h->init.vec = dst;
  h->n_init = 0;
  invalidate(h);
  API_END(h)
}

// ---- sampling / cost distribution of the final state -------------------------------------------
// evaluates `angles` with the final state kept in HBM (natural index order, S frame; |psi|^2 is what matters)
static void eval_keeping_state(Engine* h, int p, const double* angles, int n_angles) {
  if (h->world > 1) fail("sharded engine: sampling and cost distributions are not supported");
  if (!angles) fail("null angles");
  const int old = h->keep_state;
  if (!old) { h->keep_state = 1; invalidate(h); }
  try {
    energy_batch_impl(h, p, angles, n_angles, 1);
    CUDA_OK(cudaStreamSynchronize(h->stream));
    harvest_events(h);
  } catch (...) {
    if (!old) { h->keep_state = 0; invalidate(h); }
    throw;
  }
  if (!old) { h->keep_state = 0; invalidate(h); }
  if (!h->state) fail("final state was not kept");
}

int qaoa_sample(qaoa_engine_t* h, int p, const double* angles, int n_angles, int shots, const double* u,
                uint64_t* out_idx) {
  API_BEGIN(h)
  if (shots < 1 || !u || !out_idx) fail("empty sample request");
  eval_keeping_state(h, p, angles, n_angles);
  const int chunk_log2 = std::min(h->n, 16);
  const size_t nchunks = (h->N + ((size_t)1 << chunk_log2) - 1) >> chunk_log2;
  DevTmp t_sums(nchunks * 8);
  double* d_sums = t_sums.as<double>();
  if (h->dtype == QAOA_DTYPE_C64) chunk_prob_kernel<float><<<(unsigned)nchunks, 256, 0, h->stream>>>((const float*)h->state, h->N, chunk_log2, d_sums);
  else chunk_prob_kernel<double><<<(unsigned)nchunks, 256, 0, h->stream>>>((const double*)h->state, h->N, chunk_log2, d_sums);
  CUDA_OK(cudaGetLastError());
  std::vector<double> cum(nchunks);
  CUDA_OK(cudaMemcpyAsync(cum.data(), d_sums, nchunks * 8, cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  for (size_t i = 1; i < nchunks; i++) cum[i] += cum[i - 1];
  const double total = cum[nchunks - 1];
  if (!(total > 0.0)) fail("state has zero norm");
  std::vector<unsigned> chunk_of(shots);
  std::vector<double> resid(shots);
  for (int i = 0; i < shots; i++) {
    double t = std::min(std::max(u[i], 0.0), 1.0) * total;
    size_t c = std::upper_bound(cum.begin(), cum.end(), t) - cum.begin();
    if (c >= nchunks) c = nchunks - 1;
    while (c > 0 && cum[c] - cum[c - 1] <= 0.0) c--;   // never land in an empty chunk
    chunk_of[i] = (unsigned)c;
    resid[i] = t - (c ? cum[c - 1] : 0.0);
  }
  DevTmp t_chunk((size_t)shots * 4), t_resid((size_t)shots * 8), t_out((size_t)shots * 8);
  unsigned* d_chunk = t_chunk.as<unsigned>();
  double* d_resid = t_resid.as<double>();
  unsigned long long* d_out = t_out.as<unsigned long long>();
  CUDA_OK(cudaMemcpyAsync(d_chunk, chunk_of.data(), (size_t)shots * 4, cudaMemcpyHostToDevice, h->stream));
  CUDA_OK(cudaMemcpyAsync(d_resid, resid.data(), (size_t)shots * 8, cudaMemcpyHostToDevice, h->stream));
  if (h->dtype == QAOA_DTYPE_C64) sample_locate_kernel<float><<<shots, 256, 0, h->stream>>>((const float*)h->state, h->N, chunk_log2, d_chunk, d_resid, d_out);
  else sample_locate_kernel<double><<<shots, 256, 0, h->stream>>>((const double*)h->state, h->N, chunk_log2, d_chunk, d_resid, d_out);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaMemcpyAsync(out_idx, d_out, (size_t)shots * 8, cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  h->launches += 2;
  API_END(h)
}

int qaoa_extreme_solutions(qaoa_engine_t* h, int p, const double* angles, int n_angles, int want_max, int cap,
                           uint64_t* out_idx, int* out_count, double* out_cost) {
  API_BEGIN(h)
  if (cap < 1 || !out_idx || !out_count) fail("empty request");
  eval_keeping_state(h, p, angles, n_angles);
  double res[5];
  d2h(h, res, h->d_out, sizeof(res));
  const double target = want_max ? res[3] : res[2];
  DevTmp t_idx((size_t)cap * 8), t_cnt(4);
  unsigned long long* d_idx = t_idx.as<unsigned long long>();
  unsigned* d_cnt = t_cnt.as<unsigned>();
  CUDA_OK(cudaMemsetAsync(d_cnt, 0, 4, h->stream));
  const int grid = grid_for(h->N, 256);
  if (h->dtype == QAOA_DTYPE_C64) collect_cost_equal_kernel<float><<<grid, 256, 0, h->stream>>>((const float*)h->state, h->diag, h->dinfo, h->N, target, (unsigned)cap, d_idx, d_cnt);
  else collect_cost_equal_kernel<double><<<grid, 256, 0, h->stream>>>((const double*)h->state, h->diag, h->dinfo, h->N, target, (unsigned)cap, d_idx, d_cnt);
  CUDA_OK(cudaGetLastError());
  unsigned cnt = 0;
  CUDA_OK(cudaMemcpyAsync(&cnt, d_cnt, 4, cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  const unsigned got = std::min<unsigned>(cnt, (unsigned)cap);
  d2h(h, out_idx, d_idx, (size_t)got * 8);
  std::sort(out_idx, out_idx + got);
  *out_count = (int)cnt;
  if (out_cost) *out_cost = target;
  h->launches += 1;
  API_END(h)
}

int qaoa_cost_distribution(qaoa_engine_t* h, int p, const double* angles, int n_angles, double* probs256,
                           double* c0, double* delta) {
  API_BEGIN(h)
  if (!probs256) fail("null output");
  if (h->dinfo.kind != DIAG_U8) fail("cost distribution needs integer-valued costs with a range <= 255");
  eval_keeping_state(h, p, angles, n_angles);
  const int grid = grid_for(h->N, 256);
  DevTmp t_part((size_t)grid * 256 * 8);
  double* d_part = t_part.as<double>();
  if (h->dtype == QAOA_DTYPE_C64) cost_hist_kernel<float><<<grid, 256, 0, h->stream>>>((const float*)h->state, (const uint8_t*)h->diag, h->N, d_part);
  else cost_hist_kernel<double><<<grid, 256, 0, h->stream>>>((const double*)h->state, (const uint8_t*)h->diag, h->N, d_part);
  CUDA_OK(cudaGetLastError());
  std::vector<double> part((size_t)grid * 256);
  CUDA_OK(cudaMemcpyAsync(part.data(), d_part, part.size() * 8, cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  double tot = 0.0;
  for (int k = 0; k < 256; k++) {
    double s = 0.0;
    for (int b = 0; b < grid; b++) s += part[(size_t)b * 256 + k];
    probs256[k] = s;
    tot += s;
  }
  if (!(tot > 0.0)) fail("state has zero norm");
  for (int k = 0; k < 256; k++) probs256[k] /= tot;
  if (c0) *c0 = h->dinfo.c0;
  if (delta) *delta = h->dinfo.delta;
  h->launches += 1;
  API_END(h)
}

// exact CVaR_alpha of the final state's cost distribution for every cost representation (cvar_digit_kernel)
int qaoa_cvar(qaoa_engine_t* h, int p, const double* angles, int n_angles, double alpha, double* out2) {
  API_BEGIN(h)
  if (!out2) fail("null output");
  if (!(alpha > 0.0) || alpha > 1.0) fail("cvar fraction must be in (0, 1]");
  eval_keeping_state(h, p, angles, n_angles);
  const int grid = grid_for(h->N, 256);
  DevTmp t_part((size_t)grid * 512 * 8), t_sum(512 * 8);
  double* d_part = t_part.as<double>();
  double* d_sum = t_sum.as<double>();
  const int keybits = h->dinfo.kind == DIAG_U8 ? 8 : (h->dinfo.kind == DIAG_U32FX ? 32 : 64);
  unsigned long long prefix = 0;
  double above_mass = 0.0, above_pc = 0.0, total = 0.0, thr_mass = 0.0;
  std::vector<double> hist(512);
  for (int shift = keybits - 8; shift >= 0; shift -= 8) {
    const int check = shift != keybits - 8;
    if (h->dtype == QAOA_DTYPE_C64)
      cvar_digit_kernel<float><<<grid, 256, 0, h->stream>>>((const float*)h->state, h->diag, h->dinfo, h->N, prefix, shift, check, d_part);
    else
      cvar_digit_kernel<double><<<grid, 256, 0, h->stream>>>((const double*)h->state, h->diag, h->dinfo, h->N, prefix, shift, check, d_part);
    CUDA_OK(cudaGetLastError());
    sum_partials_kernel<<<2, 256, 0, h->stream>>>(d_part, grid, 512, d_sum);
    CUDA_OK(cudaGetLastError());
    h->launches += 2;
    d2h(h, hist.data(), d_sum, 512 * 8);
    if (!check) {
      for (int d = 0; d < 256; d++) total += hist[d];
      if (!(total > 0.0)) fail("state has zero norm");
    }
    const double want = alpha * total;
    // the largest digit d* with  above + mass(digits > d*) < want <= above + mass(digits >= d*)
    int dstar = 0;
    double am = above_mass, ap = above_pc;
    for (int d = 255; d >= 0; d--) {
      if (am + hist[d] >= want || d == 0) { dstar = d; break; }
      am += hist[d];
      ap += hist[256 + d];
    }
    // rounding can leave the target unreached at digit 0 with an empty bin: fall to the lowest populated digit
    if (hist[dstar] <= 0.0) for (int d = 0; d < 256; d++) if (hist[d] > 0.0) { dstar = d; break; }
    above_mass = am; above_pc = ap;
    thr_mass = hist[dstar];
    prefix = (prefix << 8) | (unsigned long long)dstar;
  }
  double cthr;
  if (h->dinfo.kind == DIAG_F64) {
    const unsigned long long b = (prefix >> 63) ? (prefix & 0x7fffffffffffffffull) : ~prefix;
    long long ll = (long long)b;
    memcpy(&cthr, &ll, 8);
  } else cthr = h->dinfo.c0 + h->dinfo.delta * (double)prefix;
  const double want = alpha * total;
  const double take = std::min(std::max(want - above_mass, 0.0), thr_mass);
  out2[0] = (above_pc + take * cthr) / (above_mass + take);
  out2[1] = cthr;
  API_END(h)
}

// the same call, returning the five statistics of the SAME evaluation as well: out7 = {CVaR, threshold cost, E[cost],
// E[cost^2], min, max, norm} -- a cvar < 1 run needs one circuit evaluation per angle vector, not two
int qaoa_cvar_stats(qaoa_engine_t* h, int p, const double* angles, int n_angles, double alpha, double* out7) {
  if (!h) return 1;
  if (!out7) { h->err = "null output"; return 2; }
  const int rc = qaoa_cvar(h, p, angles, n_angles, alpha, out7);
  if (rc) return rc;
  API_BEGIN(h)
  d2h(h, out7 + 2, h->d_out, 5 * sizeof(double));     // eval_keeping_state left the evaluation's statistics here
  API_END(h)
}

// ---- sharded registers ----------------------------------------------------------------------
int qaoa_shard_info(qaoa_engine_t* h, int* rank, int* world, int* local_qubits) {
  API_BEGIN(h)
  if (rank) *rank = h->rank;
  if (world) *world = h->world;
  if (local_qubits) *local_qubits = h->El - h->qshift;
  API_END(h)
}

static void* shard_local_buffer(Engine* h, int which) {
  if (which == 0) { ensure_state(h, 1); return h->state; }
  if (which == 1) { if (!h->diag) fail("no cost function set"); return h->diag; }
  fail("which must be 0 (state) or 1 (cost diagonal)");
}

int qaoa_shard_local_ptr(qaoa_engine_t* h, int which, void** out) {
  API_BEGIN(h)
  if (!out) fail("null output");
  *out = shard_local_buffer(h, which);
  API_END(h)
}

int qaoa_shard_set_peers(qaoa_engine_t* h, int which, void* const* ptrs) {
  API_BEGIN(h)
  if (h->world < 2) fail("not a sharded engine");
  void* mine = shard_local_buffer(h, which);
  close_peers(h, which);
  void** dst = which == 0 ? h->peer_state : h->peer_diag;
  for (int r = 0; r < h->world; r++) {
    if (!ptrs || !ptrs[r]) fail("null peer pointer");
    dst[r] = r == h->rank ? mine : ptrs[r];
    if (r != h->rank) {   // same-process peers on another GPU: map their memory into this device
      cudaPointerAttributes at;
      CUDA_OK(cudaPointerGetAttributes(&at, ptrs[r]));
      if (at.type == cudaMemoryTypeDevice && at.device != h->device) {
        int can = 0;
        CUDA_OK(cudaDeviceCanAccessPeer(&can, h->device, at.device));
        if (!can) fail("peer GPU memory is not accessible from this device (no NVLink / P2P)");
        cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else CUDA_OK(e);
      }
    }
  }
  if (which == 0) h->peers_state_set = true; else { h->peers_diag_set = true; free_streams(h); invalidate(h); }
  API_END(h)
}

int qaoa_shard_export(qaoa_engine_t* h, int which, unsigned char* handle64) {
  API_BEGIN(h)
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is expected to be 64 bytes");
  if (!handle64) fail("null handle buffer");
  cudaIpcMemHandle_t hd;
  CUDA_OK(cudaIpcGetMemHandle(&hd, shard_local_buffer(h, which)));
  memcpy(handle64, &hd, 64);
  API_END(h)
}

int qaoa_shard_import(qaoa_engine_t* h, int which, const unsigned char* handles) {
  API_BEGIN(h)
  if (h->world < 2) fail("not a sharded engine");
  if (!handles) fail("null handle buffer");
  void* mine = shard_local_buffer(h, which);
  close_peers(h, which);
  void** dst = which == 0 ? h->peer_state : h->peer_diag;
  bool* ipc = which == 0 ? h->peer_state_ipc : h->peer_diag_ipc;
  for (int r = 0; r < h->world; r++) {
    if (r == h->rank) { dst[r] = mine; continue; }
    cudaIpcMemHandle_t hd;
    memcpy(&hd, handles + (size_t)r * 64, 64);
    void* ptr = nullptr;
    CUDA_OK(cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess));
    dst[r] = ptr;
    ipc[r] = true;
  }
  if (which == 0) h->peers_state_set = true; else { h->peers_diag_set = true; free_streams(h); invalidate(h); }
  API_END(h)
}

int qaoa_shard_close_peers(qaoa_engine_t* h) {
  API_BEGIN(h)
  CUDA_OK(cudaStreamSynchronize(h->stream));
  close_peers(h, 0);
  close_peers(h, 1);
  free_streams(h);
  invalidate(h);
  API_END(h)
}

int qaoa_shard_begin(qaoa_engine_t* h, int p, const double* angles, int n_angles, int* n_passes) {
  API_BEGIN(h)
  if (h->world < 2) fail("not a sharded engine");
  if (!angles) fail("null angles");
  check_ready(h, p, n_angles);
  if (h->keep_state && h->ztw_mode) fail("sharded symmetric register: the final state cannot be kept (use the full sharded register)");
  if (!h->peers_state_set || !h->peers_diag_set) fail("sharded engine: peer pointers have not been set");
  Plan& pl = get_plan(h, p);
  ensure_state(h, 1);
  tile_prepare(h, pl, angles, n_angles, 1, h->keep_state != 0);
  if (n_passes) *n_passes = (int)pl.passes.size();
  API_END(h)
}

int qaoa_shard_pass_is_cross(qaoa_engine_t* h, int pass) {
  if (!h || !h->ev.pl || pass < 0 || pass >= (int)h->ev.pl->passes.size()) return -1;
  return h->ev.pl->passes[pass].cross ? 1 : 0;
}

int qaoa_shard_run_pass(qaoa_engine_t* h, int pass) {
  API_BEGIN(h)
  if (h->world < 2) fail("not a sharded engine");
  tile_launch_pass(h, (size_t)pass);
  API_END(h)
}

// the CUDA stream every launch of this engine goes to (cudaStream_t as an integer): lets the caller order its own
// stream work -- e.g. an NCCL collective standing in for a cross-rank barrier -- between two passes without a host sync
int qaoa_get_stream(qaoa_engine_t* h, unsigned long long* out) {
  API_BEGIN(h)
  if (!out) fail("null output");
  *out = (unsigned long long)(uintptr_t)h->stream;
  API_END(h)
}

// qaoa_shard_finish without the device -> host copy: the five raw partial sums stay on the device (d_out5 is a device
// pointer owned by the caller, e.g. a torch tensor that goes straight into an all-gather)
int qaoa_shard_finish_device(qaoa_engine_t* h, double* d_out5) {
  API_BEGIN(h)
  if (h->world < 2) fail("not a sharded engine");
  if (!d_out5) fail("null output");
  tile_finalize(h, d_out5, true);
  API_END(h)
}

// ---- sampling and exact CVaR on a sharded register: the local halves of qaoa_sample / qaoa_cvar --------------------
// After an evaluation with option keep_state = 1 every rank holds its shard of the final state.  The caller (sharded.py)
// gathers the per-chunk probability sums of all ranks, deals the shots out to (rank, chunk, residual), has every rank
// locate its own shots, and all-reduces the per-digit histograms of the weighted quantile select.
int qaoa_shard_chunk_probs(qaoa_engine_t* h, double* out, unsigned long long cap, unsigned long long* n_out) {
  API_BEGIN(h)
  if (!h->keep_state || !h->last_valid || !h->state) fail("no kept state: set option keep_state=1 and evaluate first");
  const int chunk_log2 = std::min(h->El - h->qshift, 16);
  const size_t nchunks = (h->N + ((size_t)1 << chunk_log2) - 1) >> chunk_log2;
  if (n_out) *n_out = nchunks;
  if (out) {
    if (cap < nchunks) fail("chunk buffer too small");
    DevTmp t_sums(nchunks * 8);
    double* d_sums = t_sums.as<double>();
    if (h->dtype == QAOA_DTYPE_C64) chunk_prob_kernel<float><<<(unsigned)nchunks, 256, 0, h->stream>>>((const float*)h->state, h->N, chunk_log2, d_sums);
    else chunk_prob_kernel<double><<<(unsigned)nchunks, 256, 0, h->stream>>>((const double*)h->state, h->N, chunk_log2, d_sums);
    CUDA_OK(cudaGetLastError());
    h->launches += 1;
    d2h(h, out, d_sums, nchunks * 8);
  }
  API_END(h)
}

// out_idx: LOCAL basis indices (the caller adds the shard's offset)
int qaoa_shard_locate(qaoa_engine_t* h, int shots, const uint32_t* chunk_of, const double* resid, uint64_t* out_idx) {
  API_BEGIN(h)
  if (!h->keep_state || !h->last_valid || !h->state) fail("no kept state: set option keep_state=1 and evaluate first");
  if (shots < 1) return 0;
  if (!chunk_of || !resid || !out_idx) fail("null pointer");
  const int chunk_log2 = std::min(h->El - h->qshift, 16);
  DevTmp t_chunk((size_t)shots * 4), t_resid((size_t)shots * 8), t_out((size_t)shots * 8);
  h2d(h, t_chunk.as<unsigned>(), chunk_of, (size_t)shots * 4);
  h2d(h, t_resid.as<double>(), resid, (size_t)shots * 8);
  if (h->dtype == QAOA_DTYPE_C64) sample_locate_kernel<float><<<shots, 256, 0, h->stream>>>((const float*)h->state, h->N, chunk_log2, t_chunk.as<unsigned>(), t_resid.as<double>(), t_out.as<unsigned long long>());
  else sample_locate_kernel<double><<<shots, 256, 0, h->stream>>>((const double*)h->state, h->N, chunk_log2, t_chunk.as<unsigned>(), t_resid.as<double>(), t_out.as<unsigned long long>());
  CUDA_OK(cudaGetLastError());
  h->launches += 1;
  d2h(h, out_idx, t_out.as<unsigned long long>(), (size_t)shots * 8);
  API_END(h)
}

// one digit of the weighted quantile select over the LOCAL shard: out512 = {mass[256], mass x cost[256]}
int qaoa_shard_cvar_digit(qaoa_engine_t* h, unsigned long long prefix, int shift, int check_prefix, double* out512,
                          int* key_bits) {
  API_BEGIN(h)
  if (!h->keep_state || !h->last_valid || !h->state) fail("no kept state: set option keep_state=1 and evaluate first");
  if (key_bits) *key_bits = h->dinfo.kind == DIAG_U8 ? 8 : (h->dinfo.kind == DIAG_U32FX ? 32 : 64);
  if (out512) {
    const int grid = grid_for(h->N, 256);
    DevTmp t_part((size_t)grid * 512 * 8), t_sum(512 * 8);
    if (h->dtype == QAOA_DTYPE_C64)
      cvar_digit_kernel<float><<<grid, 256, 0, h->stream>>>((const float*)h->state, h->diag, h->dinfo, h->N, prefix, shift, check_prefix, t_part.as<double>());
    else
      cvar_digit_kernel<double><<<grid, 256, 0, h->stream>>>((const double*)h->state, h->diag, h->dinfo, h->N, prefix, shift, check_prefix, t_part.as<double>());
    CUDA_OK(cudaGetLastError());
    sum_partials_kernel<<<2, 256, 0, h->stream>>>(t_part.as<double>(), grid, 512, t_sum.as<double>());
    CUDA_OK(cudaGetLastError());
    h->launches += 2;
    d2h(h, out512, t_sum.as<double>(), 512 * 8);
  }
  API_END(h)
}

// ---- chunk pipeline (plan_chunks): the caller overlaps the cross pass of one chunk with the local passes of another
// n_chunks: 1 = this plan has no chunks; closed[i] = 1 if pass i may be launched chunk by chunk
int qaoa_shard_chunk_info(qaoa_engine_t* h, int* n_chunks, int* closed, int cap) {
  API_BEGIN(h)
  if (!h->ev.pl) fail("no evaluation in flight");
  const Plan& pl = *h->ev.pl;
  bool ok = pl.ch_m > 0 && !h->force_generic;
  for (size_t i = 0; i < pl.passes.size(); i++)
    if (pl.passes[i].chunk_closed && !h->ev.all_normal[i]) ok = false;   // a cot-form pass needs the interpreter: no chunks
  if (n_chunks) *n_chunks = ok ? (1 << pl.ch_m) : 1;
  if (closed)
    for (int i = 0; i < cap && i < (int)pl.passes.size(); i++) closed[i] = ok && pl.passes[i].chunk_closed ? 1 : 0;
  API_END(h)
}

// which_stream: 0 = the engine's main stream, 1 = its side stream; ctas: persistent CTAs of this launch (0 = default)
int qaoa_shard_run_pass_chunk(qaoa_engine_t* h, int pass, int chunk, int which_stream, int ctas) {
  API_BEGIN(h)
  if (h->world < 2) fail("not a sharded engine");
  if (!h->ev.pl) fail("no evaluation in flight");
  if (chunk >= (1 << h->ev.pl->ch_m)) fail("chunk index out of range");
  if (which_stream == 1 && !h->stream2) CUDA_OK(cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking));
  h->cur_stream = which_stream == 1 ? h->stream2 : h->stream;
  h->cur_ctas = ctas;
  h->cur_ch_c = chunk;     // < 0: the whole pass
  try { tile_launch_pass(h, (size_t)pass); } catch (...) { h->cur_stream = nullptr; h->cur_ctas = 0; h->cur_ch_c = -1; throw; }
  h->cur_stream = nullptr; h->cur_ctas = 0; h->cur_ch_c = -1;
  API_END(h)
}

// launches of the last timed evaluation (option time_passes = 1): 5 doubles per launch {program id, chunk, stream,
// start ms, end ms}; returns the number of launches in *n_out (out may be null)
int qaoa_get_timeline(qaoa_engine_t* h, double* out, int cap, int* n_out) {
  API_BEGIN(h)
  if (n_out) *n_out = (int)h->timeline.size();
  if (out)
    for (int i = 0; i < cap && i < (int)h->timeline.size(); i++) {
      const auto& e = h->timeline[i];
      out[5 * i] = e.prog; out[5 * i + 1] = e.chunk; out[5 * i + 2] = e.side; out[5 * i + 3] = e.t0; out[5 * i + 4] = e.t1;
    }
  API_END(h)
}

int qaoa_get_stream2(qaoa_engine_t* h, unsigned long long* out) {
  API_BEGIN(h)
  if (!out) fail("null output");
  if (!h->stream2) CUDA_OK(cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking));
  *out = (unsigned long long)(uintptr_t)h->stream2;
  API_END(h)
}

int qaoa_shard_sync(qaoa_engine_t* h) {
  API_BEGIN(h)
  CUDA_OK(cudaStreamSynchronize(h->stream));
  if (h->stream2) CUDA_OK(cudaStreamSynchronize(h->stream2));
  harvest_events(h);
  API_END(h)
}

int qaoa_shard_finish(qaoa_engine_t* h, double* out5) {
  API_BEGIN(h)
  if (h->world < 2) fail("not a sharded engine");
  if (!out5) fail("null output");
  ensure(h->d_out, h->out_cap, 5 * sizeof(double));
  tile_finalize(h, h->d_out, true);
  CUDA_OK(cudaMemcpyAsync(out5, h->d_out, 5 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  harvest_events(h);
  API_END(h)
}

int qaoa_shard_grover_step(qaoa_engine_t* h, int step, int p, const double* angles, int n_angles,
                           const double* dot_global, double* out5) {
  API_BEGIN(h)
  if (h->world < 2) fail("not a sharded engine");
  if (h->mixer != MIXER_GROVER) fail("qaoa_shard_grover_step: the mixer is not the Grover mixer");
  if (!angles || !out5) fail("null pointer");
  check_ready(h, p, n_angles);
  if (step < 0 || step > p) fail("step out of range");
  if (step > 0 && !dot_global) fail("the global dot product of the previous step is required");
  if (h->init.kind == INIT_STATEVECTOR || h->sub.kind == INIT_STATEVECTOR) fail("sharded engine: statevector states are not supported");
  if (h->dtype == QAOA_DTYPE_C64) shard_grover_step<float>(h, step, p, angles, dot_global, out5);
  else shard_grover_step<double>(h, step, p, angles, dot_global, out5);
  API_END(h)
}

// Planner dump: builds the tile plan of a hypothetical engine WITHOUT touching a device (the planner is host code)
// and writes it as JSON -- lets CPU tests check the planner's invariants at sizes no test can execute.
int qaoa_debug_plan(int n_qubits, int dtype, int rank, int world, int symmetric, int mixer, int p, int options,
                    char* out, size_t cap) {
  if (!out || cap < 2) return 1;
  out[0] = 0;
  try {
    if (n_qubits < 1 || n_qubits > QAOA_MAX_QUBITS - 1) fail("n_qubits out of range");
    int gbits = 0;
    while ((1 << gbits) < world) gbits++;
    if (world < 1 || world > QAOA_MAX_RANKS || (1 << gbits) != world || rank < 0 || rank >= world) fail("bad rank / world");
    if (mixer != MIXER_X && mixer != MIXER_XMULTI) fail("the tile planner serves the X and multi-angle X mixers");
    if (world > 1 && n_qubits - gbits + (dtype == QAOA_DTYPE_C128 ? 1 : 0) < TILE_BITS + 1)
      fail("sharded engine: every shard must hold more than 14 element bits");   // (as qaoa_create_sharded)
    Engine e;
    Engine* h = &e;
    h->n = n_qubits; h->dtype = dtype;
    h->qshift = dtype == QAOA_DTYPE_C128 ? 1 : 0;
    h->E = n_qubits + h->qshift;
    h->rank = rank; h->world = world; h->gbits = gbits;
    h->El = h->E - gbits;
    h->N = (size_t)1 << (n_qubits - gbits);
    h->mixer = mixer;
    h->init.kind = INIT_PLUS;
    h->dinfo.kind = DIAG_U8;
    h->cost_symmetric = symmetric ? 1 : 0;
    h->z2_enabled = symmetric ? 1 : 0;
    h->schedule = options & 3;                 // 0 automatic, 1 classic, 2 balanced
    h->low_bits = (options >> 2) & 7;          // 0 automatic, 4, 5
    h->cross_spare = (options & 32) ? 0 : 1;
    if (symmetric && world > 1) {
      if (dtype != QAOA_DTYPE_C64 || (n_qubits & 1) || h->El - 1 < TILE_BITS + 1) fail("no sharded symmetric register for this configuration");
      h->ztw_mode = 1; h->N >>= 1; h->El -= 1;
    }
    if (world > 1) h->peers_state_set = true;
    Plan pl = build_plan(h, p);
    h->chunking = (options & 64) ? 0 : 1;
    plan_chunks(h, pl);
    static const char* names[] = {"Generic", "FirstInit", "FirstLoad", "Interior", "Boundary", "Final", "Mid", "Bound2",
                                  "Final2", "BoundX", "FinalX", "Bound2L4", "Final2L4"};
    std::string js = "{";
    js += "\"z2\": " + std::string(pl.z2 ? "true" : "false") + ", \"kx\": " + std::to_string(pl.kx) +
          ", \"el\": " + std::to_string(plan_el(h, pl)) + ", \"qshift\": " + std::to_string(h->qshift) +
          ", \"E\": " + std::to_string(h->E) + ", \"gbits\": " + std::to_string(h->gbits) +
          ", \"chunks\": {\"m\": " + std::to_string(pl.ch_m) + ", \"u\": " + std::to_string(pl.ch_u) + ", \"v\": [" +
          std::to_string(pl.ch_v[0]) + "," + std::to_string(pl.ch_v[1]) + "," + std::to_string(pl.ch_v[2]) +
          "], \"open_shape\": " + std::to_string(pl.ch_open_shape) +
          "}, \"shapes\": [";
    for (size_t i = 0; i < pl.shapes.size(); i++) {
      const Shape& sh = pl.shapes[i];
      js += std::string(i ? ", " : "") + "{\"pos\": [";
      for (int j = 0; j < TILE_BITS; j++) js += std::string(j ? "," : "") + std::to_string(sh.pos[j]);
      js += "], \"free\": [";
      for (size_t j = 0; j < sh.freepos.size(); j++) js += std::string(j ? "," : "") + std::to_string(sh.freepos[j]);
      js += "], \"cross\": " + std::string(sh.cross ? "true" : "false") + ", \"zshape\": " + (sh.zshape ? "true" : "false") + "}";
    }
    js += "], \"passes\": [";
    for (size_t i = 0; i < pl.passes.size(); i++) {
      const PlanPass& pp = pl.passes[i];
      js += std::string(i ? ", " : "") + "{\"prog\": \"" + names[pp.prog] + "\", \"shape\": " + std::to_string(pp.shape_index) +
            ", \"cross\": " + (pp.cross ? "true" : "false") + ", \"closed\": " + (pp.chunk_closed ? "true" : "false") +
            ", \"ops\": [";
      for (int k = 0; k < pp.tp.nops; k++) {
        const TileOp& op = pp.tp.ops[k];
        js += std::string(k ? ", " : "") + "[" + std::to_string(op.type) + "," + std::to_string(op.frame) + "," +
              std::to_string(op.arg) + "," + std::to_string(op.mask) + "]";
      }
      js += "], \"rounds\": [";
      for (size_t k = 0; k < pp.rounds.size(); k++) {
        js += std::string(k ? ", " : "") + "{\"layer\": " + std::to_string(pp.rounds[k].layer) + ", \"ebit\": [";
        for (int t = 0; t < 5; t++) js += std::string(t ? "," : "") + std::to_string(pp.rounds[k].ebit[t]);
        js += "]}";
      }
      js += "], \"phases\": [";
      for (size_t k = 0; k < pp.phases.size(); k++) js += std::string(k ? "," : "") + std::to_string(pp.phases[k].layer);
      js += "], \"reduce\": " + std::string(pp.has_reduce ? "true" : "false") + "}";
    }
    js += "]}";
    if (js.size() + 1 > cap) fail("output buffer too small: need " + std::to_string(js.size() + 1) + " bytes");
    memcpy(out, js.c_str(), js.size() + 1);
  } catch (const std::string& e) {
    g_global_error = e;
    return 2;
  }
  return 0;
}

// Test / profiling hook: runs an arbitrary step program on tiles over {0..3} U [hb, hb+10) with the
// interpreter kernel and returns the mean CUDA-event time per launch.  ops = (type, frame, mask)
// triples; PHASE / REDUCE are not allowed here (they need diagonal streams).
int qaoa_debug_tile_program(qaoa_engine_t* h, int hb, const int32_t* ops, int nops, int reps, double tau_val,
                            double* ms_out) {
  API_BEGIN(h)
  if (h->E < TILE_BITS) fail("register too small for the tile path");
  // hb >= 100 encodes c = hb / 100 low bits and a window start hb % 100 for the other 14 - c bits
  int clow = 4;
  if (hb >= 100) { clow = hb / 100; hb = hb % 100; }
  if (hb < clow || hb + (14 - clow) > h->E) fail("high-bit window out of range");
  if (nops < 1 || nops > TILE_MAX_OPS) fail("bad program length");
  ensure_state(h, 1);
  TilePass tp;
  memset(&tp, 0, sizeof(tp));
  for (int j = 0; j < clow; j++) tp.pos[j] = j;
  for (int j = 0; j < 14 - clow; j++) tp.pos[clow + j] = hb + j;
  for (int b = clow; b < h->E; b++) if (b < hb || b >= hb + (14 - clow)) tp.freepos[tp.nfree++] = b;
  if (!tilepass_set_free_runs(tp)) fail("tile shape with more than two runs of free bits");
  tp.qshift = h->qshift;
  tp.init_kind = INIT_PLUS;
  tp.init_amp = pow(2.0, -0.5 * h->n);
  int nr = 0;
  for (int i = 0; i < nops; i++) {
    const int ty = ops[3 * i];
    if (ty == TOP_PHASE || ty == TOP_REDUCE) fail("PHASE/REDUCE not supported by the debug hook");
    tp.ops[i].type = (int16_t)ty;
    tp.ops[i].frame = (int16_t)ops[3 * i + 1];
    tp.ops[i].mask = (int16_t)ops[3 * i + 2];
    tp.ops[i].arg = (int16_t)(ty == TOP_ROUND ? nr++ : 0);
  }
  tp.nops = nops;
  tp.nrounds = nr;
  tp.tau_stride = nr * TILE_ROUND_REALS;
  tp.phs_stride = 1;
  tp.prefetch_dist = h->prefetch_dist;
  const bool c64 = h->dtype == QAOA_DTYPE_C64;
  const size_t rsz = c64 ? 4 : 8;
  std::vector<unsigned char> hb_(std::max<size_t>(16, (size_t)tp.tau_stride * rsz + 16), 0);
  for (int r = 0; r < nr; r++)
    for (int s2 = 0; s2 < 5; s2++) {
      // only slots in the step's mask are active
      int mask = 0, seen = 0;
      for (int i = 0; i < nops; i++) if (ops[3 * i] == TOP_ROUND) { if (seen == r) mask = ops[3 * i + 2]; seen++; }
      const double k1 = ((mask >> s2) & 1) ? tau_val : 0.0;
      if (c64) { ((float*)hb_.data())[r * TILE_ROUND_REALS + s2] = (float)k1; ((float*)hb_.data())[r * TILE_ROUND_REALS + 5 + s2] = (float)-k1; }
      else { ((double*)hb_.data())[r * TILE_ROUND_REALS + s2] = k1; ((double*)hb_.data())[r * TILE_ROUND_REALS + 5 + s2] = -k1; }
    }
  ensure(h->d_params, h->params_cap, hb_.size() + 64);
  h2d(h, h->d_params, hb_.data(), hb_.size());
  double one = 1.0;
  h2d(h, (char*)h->d_params + hb_.size(), &one, 8);
  ensure(h->d_partials, h->partials_cap, 4096);
  const unsigned ntiles = 1u << (h->El - TILE_BITS);
  PlanPass pp;
  pp.tp = tp;
  cudaEvent_t e0, e1;
  CUDA_OK(cudaEventCreate(&e0));
  CUDA_OK(cudaEventCreate(&e1));
  auto once = [&]() {
    if (c64) launch_pass<float2>(h, pp, false, ntiles, 1, h->N << h->qshift, h->d_params, (const double*)((char*)h->d_params + hb_.size()), 1);
    else launch_pass<double>(h, pp, false, ntiles, 1, h->N << h->qshift, h->d_params, (const double*)((char*)h->d_params + hb_.size()), 1);
  };
  once();
  CUDA_OK(cudaStreamSynchronize(h->stream));
  CUDA_OK(cudaEventRecord(e0, h->stream));
  for (int r = 0; r < reps; r++) once();
  CUDA_OK(cudaEventRecord(e1, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  float ms = 0;
  CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (ms_out) *ms_out = ms / reps;
  h->launches += reps + 1;
  API_END(h)
}

}  // extern "C"
