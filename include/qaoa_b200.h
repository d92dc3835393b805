/* qaoa_b200.h -- C ABI of the B200-native QAOA statevector engine (libqaoa_b200.so).
 *
 * This is the drop-in boundary for the ONE hot path of OpenQuantumComputing/QAOA that the engine
 * replaces: the energy evaluation that the reference performs as
 *     transpile -> assign_parameters -> backend.run -> get_counts -> problem.cost loop -> Statistic
 * (reference qaoa/qaoa.py:516-519, 592-614, 653-661, 691-740, 912-922, 1053-1066).
 * All entry points are extern "C", take plain pointers / sizes, return 0 on success and a non-zero
 * code on failure (message via qaoa_last_error).  No C++ exception crosses the boundary.
 *
 * Ownership: the caller owns every host buffer for the duration of the call only (the engine
 * copies); the engine owns all device memory and streams.  One handle = one host thread at a time;
 * calls are blocking (stream-synchronised on return).
 *
 * Bit convention: bit i of a basis index x is qubit i and equals string[i] of the string the
 * reference hands to Problem.cost (qaoa/qaoa.py:605-609).
 */
#ifndef QAOA_B200_H
#define QAOA_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qaoa_engine qaoa_engine_t;

#define QAOA_DTYPE_C64 0
#define QAOA_DTYPE_C128 1

#define QAOA_MIXER_X 0       /* qaoa/mixers/x_mixer.py:28-35            */
#define QAOA_MIXER_XMULTI 1  /* qaoa/mixers/x_multiangle_mixer.py:24-50 */
#define QAOA_MIXER_XY 2      /* qaoa/mixers/xy_mixer.py:42-79           */
#define QAOA_MIXER_GROVER 3  /* qaoa/mixers/grover_mixer.py:43-82       */
#define QAOA_MIXER_NODE_GROVER 4 /* qaoa/mixers/maxkcut_grover_mixer.py:116-124 (tensorized=True): the Grover mixer of a
                                    b-qubit node state on every node register, one shared beta */

#define QAOA_INIT_PLUS 0         /* qaoa/initialstates/plus_initialstate.py:19-25        */
#define QAOA_INIT_DICKE 1        /* qaoa/initialstates/dicke_initialstate.py:23-56       */
#define QAOA_INIT_STATEVECTOR 2  /* qaoa/initialstates/statevector_initialstate.py:19-33 */
#define QAOA_INIT_PLUS_PHASES 3  /* qaoa/initialstates/plus_parameterized_initialstate.py:25-64: k phases lead the
                                    angle vector (n_init = k); beyond one CTA: X-type mixers, prepared state + tile plan */

/* Creates an engine for n_qubits on CUDA device `device`.  Replaces the construction of the
 * transpiled parameterised circuit + AerSimulator instance (qaoa/qaoa.py:204, 418-520). */
int qaoa_create(int n_qubits, int dtype, int device, qaoa_engine_t** out);
int qaoa_destroy(qaoa_engine_t* h);
const char* qaoa_last_error(qaoa_engine_t* h);
/* error text when no handle exists (failed qaoa_create) */
const char* qaoa_global_error(void);

/* Tunables / test hooks: "small_max_qubits", "keep_state", "batch_points", "batch_bytes",
 * "force_generic", "prefetch_dist", "persistent_ctas", "low_bits", "schedule", "feasible_subspace", "time_passes",
 * "cross_spare" (sharded registers: 1 = the cross tile may carry spare local qubits so that the local tiles keep 256-byte
 * rows, 0 = never, 10 + k = exactly k -- a test hook),
 * "symmetric" (default 1: when cost(x) == cost(~x) -- checked on the device --, the mixer is X / multi-angle X, the
 * initial state |+> and n even, the tile path stores only the 2^(n-1) amplitudes with the top qubit = 0; results are
 * those of the full register; 0 = always store the full register),
 * "shard_symmetric" (sharded engines, to be set to 1 right after qaoa_create_sharded, before any buffer is exported and
 * before the cost function is set: the engine holds 2^(n-1-log2 world) STORED amplitudes of the symmetric half register
 * in twisted shard coordinates -- see "Sharded registers" below; complex64, n even, MaxCut cost, X / multi-angle X mixer
 * and |+> only, anything else fails),
 * "layer_budget" (default 1: multi-angle mixers keep the lifted butterfly form for every qubit of a layer whose TOTAL
 * growth fits the exponent range, however close a single angle is to pi/2; 0 = the fixed per-angle bound),
 * "chunk_pipeline" (default 1) / "chunk_bits" (default 3): sharded registers, the plan marks passes that may be launched
 * chunk by chunk (up to 2^chunk_bits chunks) so that cross passes overlap local passes (qaoa_shard_chunk_info),
 * "column_kernels" (default 0, EXPERIMENTAL: BOUND2 / FINAL2 passes on the TMA / mbarrier kernels of
 * csrc/column_kernel.cuh, not validated) and "column_sync" (their debug hook). */
int qaoa_set_option(qaoa_engine_t* h, const char* key, double value);
double qaoa_get_counter(qaoa_engine_t* h, const char* key); /* "launches", "bytes_alg", ... */

/* Cost diagonal = GraphProblem.cost for every basis index (qaoa/problems/graph_problem.py:152-176).
 * u,v: relabelled edge endpoints in G.edges() order; w: weights; bits_per_node = ceil(log2 k);
 * lut[2^bits]: colour id of the b-bit integer whose bit j is qubit v*b+j; n_nodes includes a fixed
 * last node when fixed_node != 0 (fix_one_node, graph_problem.py:148-149), whose colour id is
 * fixed_colour. */
int qaoa_set_cost_maxkcut(qaoa_engine_t* h, int n_nodes, int n_edges, const int32_t* u,
                          const int32_t* v, const double* w, int bits_per_node, const uint8_t* lut,
                          int fixed_node, int fixed_colour);
/* Same cost, but the phase separator takes ONE GAMMA PER TERM: edge e belongs to term term_of_edge[e]
 * (MaxCutFree: one term per edge, qaoa/problems/maxcut_free_problem.py:31-99; MaxCutOrbit: one per edge orbit,
 * qaoa/problems/maxcut_orbit_problem.py:51-181) and layer d multiplies by exp(i sum_k gamma_{d,k} term_k(x)).
 * Angle layout per layer: n_terms gammas, then the betas.  Registers that fit one CTA (n <= 14 / 13) use term diagonals
 * in the single-CTA kernel; larger ones run layer by layer (edge phases on the fly, X-type mixers). */
int qaoa_set_cost_maxkcut_terms(qaoa_engine_t* h, int n_nodes, int n_edges, const int32_t* u, const int32_t* v,
                                const double* w, const int32_t* term_of_edge, int n_terms, int bits_per_node,
                                const uint8_t* lut, int fixed_node, int fixed_colour);
/* cost(x) = -(x^T Q x + c^T x + b)  (qaoa/problems/qubo_problem.py:101-108); Q row-major n*n. */
int qaoa_set_cost_qubo(qaoa_engine_t* h, int n, const double* Q, const double* c, double b);
/* Escape hatch for an arbitrary Problem.cost: host array of 2^n doubles. */
int qaoa_set_cost_diagonal(qaoa_engine_t* h, const double* diag);
/* Reads cost(x) for x in [offset, offset+count) back to the host. */
int qaoa_read_cost_diagonal(qaoa_engine_t* h, size_t offset, size_t count, double* out);
/* min / max of cost over all x (device reduction; Problem.computeMinMaxCosts, base_problem.py:121-134). */
int qaoa_cost_minmax(qaoa_engine_t* h, double* out_min, double* out_max);

/* pairs: XY topology as (q0,q1) pairs in application order; trotter_steps: qaoa/qaoa.py:494-503.
 * For GROVER, sub_kind/sub_k/sub_vec describe |s> = U_S|0> like qaoa_set_initial.
 * For NODE_GROVER, sub_k = qubits per node (1..3, dividing n) and sub_vec = the 2^sub_k complex amplitudes of the node
 * state (re, im interleaved); node v occupies qubits [v * sub_k, (v + 1) * sub_k). */
int qaoa_set_mixer(qaoa_engine_t* h, int kind, const int32_t* pairs, int n_pairs, int trotter_steps,
                   int sub_kind, int sub_k, const double* sub_vec);
/* vec: 2^n complex128 amplitudes (re,im interleaved), little-endian index; only for STATEVECTOR. */
int qaoa_set_initial(qaoa_engine_t* h, int kind, int k, const double* vec);

/* One energy evaluation (replaces QAOA.loss / _eval_cost, qaoa/qaoa.py:885-935, 1029-1066).
 * angles: the reference's flat layout [init_0..init_{n_init-1}, (gamma_d[0..n_gamma-1], beta_d[0..n_beta-1]) for d < p]
 * (qaoa/qaoa.py:937-990; n_init = 0 except for QAOA_INIT_PLUS_PHASES, n_gamma = 1 except after
 * qaoa_set_cost_maxkcut_terms).
 * out[5] = { E[cost], E[cost^2], min cost over support, max cost over support, norm }. */
int qaoa_energy(qaoa_engine_t* h, int p, const double* angles, int n_angles, double* out);
/* B evaluations in one call (replaces the batch branch of sample_cost_landscape,
 * qaoa/qaoa.py:630-662).  angles: B * n_angles doubles; out: B * 5 doubles. */
int qaoa_energy_batch(qaoa_engine_t* h, int p, const double* angles, int n_angles, int B, double* out);

/* Same evaluation as qaoa_energy_batch, but the B x 5 results stay in device memory until
 * qaoa_fetch_results is called (bench.py uses it to time the HBM-resident part on its own). */
int qaoa_energy_batch_resident(qaoa_engine_t* h, int p, const double* angles, int n_angles, int B);
int qaoa_fetch_results(qaoa_engine_t* h, int B, double* out);

/* Final state of the last qaoa_energy call made with option keep_state=1 (true amplitudes,
 * complex128 interleaved).  Test / debugging aid; 2^n * 16 bytes on the host. */
int qaoa_read_statevector(qaoa_engine_t* h, double* out);

/* X gates between layers -- the reference's flip=True (qaoa/qaoa.py:505-509, qaoa/utils/flip.py:76-88): the state kept by
 * the last evaluation (option keep_state=1), with X applied on every qubit of flip_mask (bit q = qubit q), is installed
 * as the initial state, device to device, exactly as qaoa_set_initial(QAOA_INIT_STATEVECTOR) would install it.  The
 * caller evaluates the circuit segment by segment: layers up to a flip, this call, the next layers, ...; the last
 * segment's result is the result of the whole circuit.  Single-GPU engines only. */
int qaoa_continue_from_kept_state(qaoa_engine_t* h, uint64_t flip_mask);

/* Draws `shots` basis indices from |psi(angles)|^2: replaces backend.run(shots=...) + get_counts of
 * QAOA.hist (qaoa/qaoa.py:1134-1172).  u[shots]: uniform numbers in [0,1) from the caller's seeded
 * generator (inverse-CDF sampling, deterministic for given u); out_idx[i] = basis index of shot i. */
int qaoa_sample(qaoa_engine_t* h, int p, const double* angles, int n_angles, int shots, const double* u,
                uint64_t* out_idx);
/* Basis indices of the support of |psi(angles)> that attain the maximum (want_max != 0) or minimum cost over that
 * support: the shots -> infinity limit of Statistic.get_max_sols / get_min_sols (statistic.py:69-82), which feed
 * OptResult.BestSols and QAOA.get_optimal_solutions (qaoa/qaoa.py:89-91, 1193-1212).  At most `cap` indices are
 * returned (ascending); *out_count is the true number. */
int qaoa_extreme_solutions(qaoa_engine_t* h, int p, const double* angles, int n_angles, int want_max, int cap,
                           uint64_t* out_idx, int* out_count, double* out_cost);
/* P(cost = *c0 + *delta * k), k < 256, over the final state (integer-valued costs with range <= 255).
 * The shots -> infinity limit of the sorted sample list behind Statistic.get_CVaR (statistic.py:132-142). */
int qaoa_cost_distribution(qaoa_engine_t* h, int p, const double* angles, int n_angles, double* probs256,
                           double* c0, double* delta);
/* Exact CVaR_alpha of the final state's cost distribution, 0 < alpha <= 1: the mean of the cost over the best alpha of
 * the probability mass -- the shots -> infinity limit of Statistic.get_CVaR (qaoa/utils/statistic.py:132-142), for every
 * cost representation (integer, 32-bit fixed point, double): a weighted quantile select on the device, one sweep of the
 * kept final state per 8-bit digit of the cost key.  out2 = {CVaR, threshold cost}. */
int qaoa_cvar(qaoa_engine_t* h, int p, const double* angles, int n_angles, double alpha, double* out2);
/* Same, plus the statistics of the same evaluation: out7 = {CVaR, threshold cost, E[cost], E[cost^2], min, max, norm}
 * (QAOA.loss with cvar < 1 needs both, qaoa/qaoa.py:885-935: one circuit evaluation instead of two). */
int qaoa_cvar_stats(qaoa_engine_t* h, int p, const double* angles, int n_angles, double alpha, double* out7);

/* ---- Sharded registers: one engine per GPU (one process per GPU), the register split on its top
 * log2(world) qubits; rank r holds the basis indices [r * 2^(n - log2 world), (r+1) * 2^(n - log2 world)).
 * There is nothing to replace in the reference (it has no multi-device path);
 * the boundary mirrors qaoa_energy: same angle layout, same result semantics.
 *
 * Mixer gates on the top ("global") qubits are executed by CROSS passes: ordinary tile passes whose
 * tile contains the rank bits, so that their loads and stores go to the peers' HBM over NVLink
 * (peer memory) -- the exchange is fused into the pass, there is no separate swap or staging buffer.
 * The caller provides the peers' buffers, either as CUDA IPC handles (one process per GPU) or as raw
 * device pointers (several engines inside one process), and separates passes with a barrier across
 * ranks whenever qaoa_shard_pass_is_cross says so for the pass before or after it:
 *
 *   qaoa_create_sharded on every rank; qaoa_set_cost_* / qaoa_set_mixer / qaoa_set_initial
 *   exchange qaoa_shard_export(which=0 state, 1 cost diagonal) handles -> qaoa_shard_import; barrier
 *   qaoa_shard_begin(p, angles) -> n_passes
 *   for i < n_passes: [barrier if pass i or pass i-1 is cross]  qaoa_shard_run_pass(i); qaoa_shard_sync
 *   qaoa_shard_finish(out5) -> this rank's {sum p, sum p c, sum p c^2, min c, max c};
 *   combine over ranks with (sum, sum, sum, min, max): E[c] = S1/S0, E[c^2] = S2/S0, norm = S0.
 * Supported: X / multi-angle X mixers, Plus / Dicke initial states, every cost generator.
 *
 * Symmetric variant (option "shard_symmetric"): for MaxCut (cost(x) = cost(~x)), X-type mixers and |+> the state obeys
 * psi(x) = psi(~x); only the half with the top qubit = 0 is stored, split over the ranks in TWISTED coordinates: the
 * stored index x' lives on rank r at local index z, where z = the low n-1-g bits of x' and r_j = x'[n-1-g+j] XOR x'[0].
 * The mirror image of a shard is then the shard itself (no NVLink traffic for the top qubit's mixer), and qubit 0 is
 * mixed inside the cross passes.  Same call sequence and results; qaoa_read_cost_diagonal and qaoa_shard_local_ptr
 * expose the shard in (z) order; qaoa_shard_info reports n-1-g local qubits. */
int qaoa_create_sharded(int n_qubits, int dtype, int device, int rank, int world, qaoa_engine_t** out);
int qaoa_shard_info(qaoa_engine_t* h, int* rank, int* world, int* local_qubits);
int qaoa_shard_local_ptr(qaoa_engine_t* h, int which, void** out);
int qaoa_shard_set_peers(qaoa_engine_t* h, int which, void* const* ptrs);
int qaoa_shard_export(qaoa_engine_t* h, int which, unsigned char* handle64);
int qaoa_shard_import(qaoa_engine_t* h, int which, const unsigned char* handles);
/* Unmaps every peer buffer (call on all ranks, with a barrier before and after, before any rank destroys its
 * engine: an exporter must not free memory that a peer still has mapped). */
int qaoa_shard_close_peers(qaoa_engine_t* h);
int qaoa_shard_begin(qaoa_engine_t* h, int p, const double* angles, int n_angles, int* n_passes);
int qaoa_shard_pass_is_cross(qaoa_engine_t* h, int pass); /* 1 / 0, -1 on error */
int qaoa_shard_run_pass(qaoa_engine_t* h, int pass);
int qaoa_shard_sync(qaoa_engine_t* h);
int qaoa_shard_finish(qaoa_engine_t* h, double* out5);
/* Stream-ordered variants (no host synchronisation inside an evaluation): qaoa_get_stream returns the engine's
 * cudaStream_t so that the caller can enqueue its cross-rank barrier (an NCCL collective) between two passes on that
 * stream; qaoa_shard_finish_device leaves the five raw partial sums in device memory (d_out5: device pointer). */
int qaoa_get_stream(qaoa_engine_t* h, unsigned long long* stream_out);
int qaoa_shard_finish_device(qaoa_engine_t* h, double* d_out5);
/* Chunk pipeline of a sharded evaluation: between qaoa_shard_begin and qaoa_shard_finish a pass flagged
 * closed[i] may be launched chunk by chunk (chunk = 0 .. n_chunks-1; chunk < 0: the whole pass) on the engine's main
 * (which_stream 0) or side stream (1) with `ctas` persistent CTAs (0 = one per SM).  Chunks are sets of local indices no
 * closed pass couples, so the cross pass of chunk c (few SMs, NVLink-bound) can run beside the local passes of chunk
 * c+1; the caller orders the two streams and the cross-rank barriers (qaoa_get_stream / qaoa_get_stream2). */
int qaoa_shard_chunk_info(qaoa_engine_t* h, int* n_chunks, int* closed, int cap);
int qaoa_shard_run_pass_chunk(qaoa_engine_t* h, int pass, int chunk, int which_stream, int ctas);
int qaoa_get_stream2(qaoa_engine_t* h, unsigned long long* stream_out);
/* Timeline of the last evaluation timed with option time_passes = 1: per launch {program id, chunk (-1: whole pass),
 * stream (0 main, 1 side), start ms, end ms} relative to the first launch; *n_out = number of launches. */
int qaoa_get_timeline(qaoa_engine_t* h, double* out, int cap, int* n_out);
/* Grover mixer on a sharded register (reference qaoa/mixers/grover_mixer.py:43-82 over |+> / |D^n_k>,
 * dicke_initialstate.py:23-56): no state exchange, one complex scalar per layer.  step = 0 .. p-1 runs layer `step`'s
 * sweep over the local shard and returns the LOCAL partial <s|v> in out5[0..1]; the caller sums it over the ranks
 * (all-reduce) and passes the global value as `dot_global` of the next step; step = p applies the last reflection and
 * returns the raw local statistics {sum p, sum p c, sum p c^2, min c, max c} (combine as for qaoa_shard_finish). */
int qaoa_shard_grover_step(qaoa_engine_t* h, int step, int p, const double* angles, int n_angles,
                           const double* dot_global, double* out5);
/* Sampling (QAOA.hist, qaoa/qaoa.py:1134-1172) and exact CVaR (statistic.py:132-142) on a sharded register: the local
 * halves of qaoa_sample / qaoa_cvar.  With option keep_state = 1 an evaluation (qaoa_shard_begin ... / the Grover steps)
 * leaves every rank's shard of the final state in HBM (full sharded register only).  qaoa_shard_chunk_probs: probability
 * mass of every 2^16-amplitude chunk of the local shard (out = NULL: only their number); qaoa_shard_locate: the LOCAL basis
 * index of each shot given its chunk and the residual mass inside it; qaoa_shard_cvar_digit: one 8-bit digit of the
 * weighted quantile select over the local shard (out512 = mass[256], mass x cost[256]; key_bits = 8 / 32 / 64).  The
 * driver gathers / all-reduces these small arrays (qaoa_b200/sharded.py). */
int qaoa_shard_chunk_probs(qaoa_engine_t* h, double* out, unsigned long long cap, unsigned long long* n_out);
int qaoa_shard_locate(qaoa_engine_t* h, int shots, const uint32_t* chunk_of, const double* resid, uint64_t* out_idx);
int qaoa_shard_cvar_digit(qaoa_engine_t* h, unsigned long long prefix, int shift, int check_prefix, double* out512,
                          int* key_bits);

/* Planner dump (no device needed): the tile plan the engine would build for this configuration, as JSON -- tile shapes,
 * passes, their step programs, which qubit of which layer every butterfly round mixes, which layer's phase separator a
 * pass applies.  options: bits 0-1 schedule (0 automatic, 1 classic, 2 balanced), bits 2-4 low_bits (0, 4, 5), bit 5 =
 * cross_spare off.  Error text via qaoa_global_error.  Used by tests/test_planner.py (CPU) to check the planner's
 * invariants at full BASELINE sizes. */
int qaoa_debug_plan(int n_qubits, int dtype, int rank, int world, int symmetric, int mixer, int p, int options,
                    char* out, size_t cap);

/* Profiling hook: runs an arbitrary (type, frame, mask) step program of the tiled pass kernel and
 * returns the mean CUDA-event time per launch (scripts/tile_floor.py). */
int qaoa_debug_tile_program(qaoa_engine_t* h, int hb, const int32_t* ops, int nops, int reps, double tau_val,
                            double* ms_out);

#ifdef __cplusplus
}
#endif
#endif
