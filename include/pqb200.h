/* pqb200.h — C ABI of the B200-native state-vector engine that replaces ProjectQ's C++ simulator.
 *
 * Every entry point below is what a binding of the reference's native seam would call.  The seam is the
 * pybind11 class `_cppsim.Simulator` (reference: projectq/backends/_sim/_cppsim.cpp:43-67) wrapping
 * `class Simulator` (reference: projectq/backends/_sim/_cppkernels/simulator.hpp:37-578).  Each function
 * cites the reference member it replaces.  Plain pointers and sizes only; no C++/torch types.
 *
 * Conventions
 *   - All functions return a status: PQB_OK, or an error class that the Python shim maps to the same
 *     exception class the reference raises (std::runtime_error -> RuntimeError, std::length_error /
 *     std::invalid_argument -> ValueError).  pqb_last_error() gives the message.
 *   - Arrays are caller-owned, contiguous, and only read during the call.
 *   - Complex numbers are interleaved (re, im) doubles.
 *   - Every observable call applies the pending fused-gate queue first (reference: run() at the top of every
 *     query, simulator.hpp:77,94,111,146,195,227,293,325,354,372,389,441,457,530).
 *   - There is no CPU fallback: without a CUDA device pqb_create fails with PQB_ERR_CUDA.
 */
#ifndef PQB200_H_
#define PQB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PQB_API __attribute__((visibility("default")))
#else
#define PQB_API
#endif

typedef struct pqb_sim pqb_sim;

enum pqb_status {
    PQB_OK = 0,
    PQB_ERR_RUNTIME = 1, /* reference throws std::runtime_error  -> RuntimeError */
    PQB_ERR_VALUE = 2,   /* std::length_error / std::invalid_argument -> ValueError */
    PQB_ERR_CUDA = 3,    /* CUDA / NCCL failure or no device       -> RuntimeError */
    PQB_ERR_MEMORY = 4   /* device memory exhausted                -> MemoryError */
};

/* Options beyond the reference constructor (simulator.hpp:48 takes only the seed).  Zero-initialise for defaults. */
typedef struct pqb_opts {
    int32_t device;            /* CUDA device ordinal (default 0) */
    int32_t fusion_max_qubits; /* widest fused dense gate, 1..5; 0 -> pick 4 or 5 per flush by cost (reference window: 4..5, simulator.hpp:48-49) */
    int32_t rank;              /* this process' rank in the sharded state (default 0) */
    int32_t world_size;        /* number of ranks = GPUs holding shards; power of two (0/1 -> single GPU) */
    const void* nccl_unique_id; /* 128-byte ncclUniqueId shared by all ranks (required when world_size > 1) */
    int32_t reserve_qubits;    /* optional: pre-size the state buffer for this many qubits (0 -> grow on demand) */
    int32_t reserved_[7];
} pqb_opts;

/* Pauli-string operators (reference: Term / TermsDict / ComplexTermsDict, simulator.hpp:44-46) are passed flattened:
 * term t covers entries [term_offsets[t], term_offsets[t+1]) of (qubit_index[], pauli[]); qubit_index is the position
 * in the `ids` array of the call, pauli is 'X', 'Y' or 'Z'. */
typedef struct pqb_terms {
    size_t n_terms;
    const size_t* term_offsets;   /* n_terms + 1 entries */
    const uint32_t* qubit_index;  /* term_offsets[n_terms] entries */
    const char* pauli;            /* term_offsets[n_terms] entries */
    const double* coefficients;   /* n_terms reals, or 2*n_terms (re,im) where the call says complex */
} pqb_terms;

/* Execution counters (what bench.py reports as gpu_launches, passes and remap traffic). */
typedef struct pqb_stats {
    uint64_t kernel_launches;     /* kernels of this library launched so far */
    uint64_t dense_passes[6];     /* [k] = fused dense passes applied with k target qubits (k = 1..5) */
    uint64_t gates_ingested;      /* apply_controlled_gate calls */
    uint64_t remaps;              /* global<->local qubit remaps (sharded state); one remap may move several qubits */
    uint64_t remap_bytes_sent;    /* bytes this rank sent over NVLink in remaps */
    double remap_ms;              /* device time the main stream spent waiting for remaps (the exposed part) */
    uint64_t diag_passes;         /* passes applied by the diagonal kernel */
    uint64_t p2p_remaps;          /* remaps done by the peer-memory exchange kernel rather than NCCL send/recv */
    uint64_t pipelined_remaps;    /* of those: remaps executed slice by slice, overlapped with the passes around them */
    uint64_t remap_qubits;        /* qubits moved between rank bits and local bits */
    double remap_comm_ms;         /* device time of the exchange kernels on the communication stream */
    double pass_ms[6];            /* profiling on: summed device time of the dense launches with k targets */
    double diag_ms;               /* profiling on: summed device time of the diagonal launches */
    uint64_t reserved_[4];
} pqb_stats;

/* ---- lifetime ------------------------------------------------------------------------------------------------ */
/* simulator.hpp:48-53  Simulator(unsigned seed): one amplitude = 1, mt19937(seed). */
PQB_API int pqb_create(uint32_t seed, const pqb_opts* opts, pqb_sim** out);
PQB_API void pqb_destroy(pqb_sim* sim);
/* message of the last failing call on `sim` (or of the last failing pqb_create when sim == NULL) */
PQB_API const char* pqb_last_error(const pqb_sim* sim);

/* ---- the 19 methods of _cppsim.Simulator (_cppsim.cpp:45-65) --------------------------------------------------- */
/* simulator.hpp:55-74   allocate_qubit(id): new most-significant bit, upper half zero; duplicate id -> RUNTIME */
PQB_API int pqb_allocate_qubit(pqb_sim* sim, uint32_t id);
/* simulator.hpp:194-202 deallocate_qubit(id): must be classical (else RUNTIME); compacts, shifts higher positions down */
PQB_API int pqb_deallocate_qubit(pqb_sim* sim, uint32_t id);
/* simulator.hpp:76-91   get_classical_value(id, tol) */
PQB_API int pqb_get_classical_value(pqb_sim* sim, uint32_t id, double tol, int* out);
/* simulator.hpp:93-108  is_classical(id, tol) */
PQB_API int pqb_is_classical(pqb_sim* sim, uint32_t id, double tol, int* out);
/* simulator.hpp:145-192 measure_qubits(ids) -> bits; exactly one draw of the host mt19937 stream per call */
PQB_API int pqb_measure_qubits(pqb_sim* sim, const uint32_t* ids, size_t n_ids, uint8_t* out_bits);
/* simulator.hpp:204-222 apply_controlled_gate(m, ids, ctrl): m is 2^k x 2^k row-major, matrix bit l <-> ids[l], k <= 5
 * (k > 5 -> VALUE, like run()'s std::invalid_argument at :523, but without poisoning the queue) */
PQB_API int pqb_apply_controlled_gate(pqb_sim* sim, const double* matrix_re_im, const uint32_t* ids, size_t k,
                                      const uint32_t* ctrl, size_t n_ctrl);
/* simulator.hpp:224-269 emulate_math(f, quregs, ctrl) for an arbitrary pure f, given as a table built by the caller:
 * the concatenated register value v (register 0 in the low bits, register r occupying reg_sizes[r] bits) maps to
 * table[v], the concatenated low-bits result.  table has 2^(sum reg_sizes) entries. */
PQB_API int pqb_emulate_math_table(pqb_sim* sim, const uint64_t* table, size_t table_len, const uint32_t* reg_ids_flat,
                                   const uint32_t* reg_sizes, size_t n_regs, const uint32_t* ctrl, size_t n_ctrl);
/* simulator.hpp:271-276 emulate_math_addConstant(a, quregs, ctrl): x -> x + a (low bits), every register */
PQB_API int pqb_emulate_math_add_constant(pqb_sim* sim, int64_t a, const uint32_t* reg_ids_flat,
                                          const uint32_t* reg_sizes, size_t n_regs, const uint32_t* ctrl, size_t n_ctrl);
/* simulator.hpp:278-283 emulate_math_addConstantModN(a, N, quregs, ctrl): x -> (x + a) % N (C remainder) */
PQB_API int pqb_emulate_math_add_constant_mod_n(pqb_sim* sim, int64_t a, int64_t N, const uint32_t* reg_ids_flat,
                                                const uint32_t* reg_sizes, size_t n_regs, const uint32_t* ctrl,
                                                size_t n_ctrl);
/* simulator.hpp:285-290 emulate_math_multiplyByConstantModN(a, N, quregs, ctrl): x -> (x * a) % N */
PQB_API int pqb_emulate_math_multiply_by_constant_mod_n(pqb_sim* sim, int64_t a, int64_t N,
                                                        const uint32_t* reg_ids_flat, const uint32_t* reg_sizes,
                                                        size_t n_regs, const uint32_t* ctrl, size_t n_ctrl);
/* simulator.hpp:292-322 get_expectation_value(td, ids): sum_t c_t Re<psi|P_t|psi>, real coefficients */
PQB_API int pqb_get_expectation_value(pqb_sim* sim, const pqb_terms* terms, const uint32_t* ids, size_t n_ids,
                                      double* out);
/* simulator.hpp:324-350 apply_qubit_operator(td, ids): psi <- sum_t c_t P_t psi, complex coefficients, no renormalisation */
PQB_API int pqb_apply_qubit_operator(pqb_sim* sim, const pqb_terms* terms_complex, const uint32_t* ids, size_t n_ids);
/* simulator.hpp:386-438 emulate_time_evolution(td, t, ids, ctrl): psi <- exp(-i H t) psi on the control subspace */
PQB_API int pqb_emulate_time_evolution(pqb_sim* sim, const pqb_terms* terms, double time, const uint32_t* ids,
                                       size_t n_ids, const uint32_t* ctrl, size_t n_ctrl);
/* simulator.hpp:352-368 get_probability(bits, ids); unknown id -> RUNTIME */
PQB_API int pqb_get_probability(pqb_sim* sim, const uint8_t* bits, const uint32_t* ids, size_t n_ids, double* out);
/* simulator.hpp:370-384 get_amplitude(bits, ids); ids must be a permutation of all qubits (else RUNTIME); out = (re, im) */
PQB_API int pqb_get_amplitude(pqb_sim* sim, const uint8_t* bits, const uint32_t* ids, size_t n_ids, double* out_re_im);
/* simulator.hpp:440-454 set_wavefunction(wf, ordering): position i <- ordering[i]; wf has 2^n (re,im) pairs */
PQB_API int pqb_set_wavefunction(pqb_sim* sim, const double* wf_re_im, size_t n_amplitudes, const uint32_t* ordering,
                                 size_t n_ordering);
/* simulator.hpp:456-485 collapse_wavefunction(ids, values); n_ids != n_values -> VALUE; P < 1e-12 -> RUNTIME */
PQB_API int pqb_collapse_wavefunction(pqb_sim* sim, const uint32_t* ids, size_t n_ids, const uint8_t* values,
                                      size_t n_values);
/* simulator.hpp:487-527 run(): fuse and apply everything queued */
PQB_API int pqb_run(pqb_sim* sim);
/* simulator.hpp:529-532 cheat(): the id -> bit-position map and a copy of the state.
 * pqb_num_qubits / pqb_cheat_map / pqb_cheat_state split it so the caller can size its buffers.
 * In a sharded run pqb_cheat_state returns the full 2^n logical state on every rank (small states only). */
PQB_API int pqb_num_qubits(pqb_sim* sim, size_t* out_n);
PQB_API int pqb_cheat_map(pqb_sim* sim, uint32_t* ids_out, uint32_t* positions_out, size_t capacity, size_t* out_n);
PQB_API int pqb_cheat_state(pqb_sim* sim, double* wf_re_im_out, size_t capacity_amplitudes);

/* ---- additions around the same path (no reference counterpart; used by bench.py / large-n parity) ----------------- */
/* amplitudes at `n` logical basis indices (O(n) device gathers; the large-n substitute for cheat()) */
PQB_API int pqb_get_amplitudes(pqb_sim* sim, const uint64_t* logical_indices, size_t n, double* out_re_im);
/* batched ingestion of a gate list in the layout of oracle/ref_harness.cpp's circuit file body:
 * per gate u32 k, u32 nc, u32 targets[k], u32 ctrls[nc], f64 matrix[2*4^k]; equivalent to n_gates
 * pqb_apply_controlled_gate calls (fuse = 0 additionally calls pqb_run after every gate, like gate_fusion=False) */
PQB_API int pqb_apply_gate_stream(pqb_sim* sim, const void* packed, size_t n_bytes, size_t n_gates, int fuse);
/* ---- f3: the state as data, around cheat()/set_wavefunction (simulator.hpp:440-454,529-532) ---------
 * Checkpoint of the (possibly sharded) state.  Every rank writes / reads its own file "<prefix>.rank<r>of<w>.pqbs": a
 * header (qubit ids and positions, physical layout, RNG stream position) followed by the shard's amplitudes exactly as they
 * lie in HBM — nothing is gathered or re-laid-out, so a 36-qubit state (1.1 TB over 8 ranks) is saved and restored shard by
 * shard.  Loading needs an engine with the same world size and rank; it replaces the engine's qubits, state and RNG (if it
 * fails half-way the amplitudes are undefined, the bookkeeping is unchanged). */
PQB_API int pqb_save_state(pqb_sim* sim, const char* path_prefix);
PQB_API int pqb_load_state(pqb_sim* sim, const char* path_prefix);
/* Zero-copy view for consumers that can wrap device memory (CUDA array interface / DLPack producers): the device pointer
 * of this rank's shard, its length in amplitudes, and for every logical bit position the physical bit it lives on (local
 * bit < 64, or 64 + rank bit).  The pointer is valid until the next call that allocates, deallocates or remaps. */
PQB_API int pqb_state_view(pqb_sim* sim, void** out_device_ptr, uint64_t* out_local_amplitudes, uint8_t* out_layout,
                           size_t layout_capacity, size_t* out_n_qubits);
/* set |psi> to a seeded pseudo-random normalised state directly on the device (benchmark input), n qubits ids 0..n-1 */
PQB_API int pqb_init_random_state(pqb_sim* sim, uint32_t n_qubits, uint64_t seed);
/* sum |psi_i|^2 over the whole state */
PQB_API int pqb_norm_squared(pqb_sim* sim, double* out);
PQB_API int pqb_synchronize(pqb_sim* sim);
/* CUDA-event stopwatch on the engine's stream: start, ... enqueue work ..., stop -> elapsed device ms */
PQB_API int pqb_timer_start(pqb_sim* sim);
PQB_API int pqb_timer_stop(pqb_sim* sim, double* out_ms);
/* synchronises the engine's streams and returns the counters */
PQB_API int pqb_get_stats(pqb_sim* sim, pqb_stats* out);
/* on: every dense / diagonal launch is bracketed by CUDA events on the engine's stream and its duration is added to
 * pqb_stats.pass_ms[k] / diag_ms (what bench.py reports as the dominant kernel's own launch time) */
PQB_API int pqb_set_profiling(pqb_sim* sim, int on);
PQB_API int pqb_reset_stats(pqb_sim* sim);
/* overwrite >= `bytes` of a scratch device buffer (L2 flush between timed iterations) */
PQB_API int pqb_flush_l2(pqb_sim* sim, size_t bytes);
/* micro-benchmark hook: apply one dense k-qubit gate pass directly at the given *bit positions* (bypasses the fuser) */
PQB_API int pqb_bench_dense_pass(pqb_sim* sim, const double* matrix_re_im, const uint32_t* positions, size_t k,
                                 uint64_t ctrl_mask, int repeats, double* out_ms_per_pass);
/* self-test hook: apply one dense pass slice by slice — the launch is restricted in turn to every value of the local bits in
 * slice_mask (none of which may be a target or control) — which must equal one unrestricted pass; this is how the sharded
 * engine runs the passes around a remap while other slices are on the wire */
PQB_API int pqb_selftest_sliced_pass(pqb_sim* sim, const double* matrix_re_im, const uint32_t* positions, size_t k,
                                     uint64_t ctrl_mask, uint64_t slice_mask);
/* self-test hook of the peer-memory exchange kernel on ONE device: `world` shards of 2^n_local_bits amplitudes live in one
 * process, every "rank" runs its exchange kernel for the (rank bit, local bit) pairs (one kernel per slice of slice_mask),
 * and the result is compared with the permutation a global<->local remap must perform.  out_mismatches = amplitudes in the
 * wrong place. */
PQB_API int pqb_selftest_exchange(int device, int world, int n_local_bits, const int32_t* pairs, size_t n_pairs,
                                  uint64_t slice_mask, uint64_t* out_mismatches);
/* measured FP64 FMA peak of the device in TFLOP/s (register-resident DFMA loop; roofline denominator for k = 5) */
PQB_API int pqb_measure_fp64_peak(pqb_sim* sim, double* out_tflops);
/* measured device-to-device copy bandwidth in GB/s (read + write bytes; cross-check of MEASURED_PEAKS.json) */
PQB_API int pqb_measure_copy_bandwidth(pqb_sim* sim, size_t bytes, double* out_gbs);

/* ---- host-only pieces of the path, exposed for CPU tests (no device needed) ---------------------------------------- */
/* fusion.hpp:65-109 perform_fusion restated by the new fuser: fuse the packed gate stream (same layout as
 * pqb_apply_gate_stream) into passes of at most max_qubits.  Writes per pass: u32 k, u32 nc, u32 targets[k] (ascending
 * qubit id), u32 ctrls[nc], f64 matrix[2*4^k] into out (capacity out_cap bytes); returns passes and bytes used. */
PQB_API int pqb_host_fuse_stream(const void* packed, size_t n_bytes, size_t n_gates, int max_qubits, void* out,
                                 size_t out_cap, size_t* out_bytes, size_t* out_passes);
/* simulator.hpp:43,48-53,153: the measurement RNG stream: n draws of uniform_real_distribution(0,1) on mt19937(seed) */
PQB_API int pqb_host_rng_stream(uint32_t seed, size_t n, double* out);
/* remap planner of the sharded state (projectq_b200/csrc/dist.h plan_remap): loc[p] = local bit (< 64) or 64 + rank bit of
 * logical position p (updated in place); need[] = logical positions that must become local; out_pairs receives
 * (rank bit, local bit) pairs to exchange */
PQB_API int pqb_host_plan_remap(uint8_t* loc, size_t n_logical, int n_local_bits, const uint32_t* need, size_t n_need,
                                int32_t* out_pairs, size_t cap_pairs, size_t* out_n_pairs);
/* dry run of the sharded scheduler (Engine::run_sharded) without a device: which fused passes run in which order and
 * when which qubits are exchanged, for `flushes` repetitions of a gate stream on n_qubits with the top rank_bits qubits
 * initially on rank bits.  Record layout in projectq_b200/csrc/capi.cpp. */
PQB_API int pqb_host_shard_schedule(const void* packed, size_t n_bytes, size_t n_gates, uint32_t n_qubits, uint32_t rank_bits,
                                    int max_qubits, uint32_t flushes, void* out, size_t out_cap, size_t* out_bytes);
/* self-test of the descriptor channel between rank processes (projectq_b200/csrc/fdpass.h): call from `world` processes
 * with the same run_tag; every pair exchanges a pipe descriptor and checks its content */
PQB_API int pqb_host_fdpass_selftest(uint64_t run_tag, int rank, int world);
/* exchange plan of a (multi-bit) remap for one rank (dist.h plan_exchange): pairs = (rank bit, local bit) x n_pairs;
 * for each partner rank the pattern of exchanged local bits of the sub-block swapped with it */
PQB_API int pqb_host_plan_exchange(int rank, const int32_t* pairs, size_t n_pairs, int32_t* out_peers, uint64_t* out_patterns,
                                   size_t cap, size_t* out_n);
/* planner of the tiled Pauli-operator kernels (engine.h plan_pauli_tiles): for terms given by their x/z masks over
 * n_local_bits index bits, which launch applies each term (-1: no tile can hold its X/Y support, it goes through global
 * gathers) and the tile bits of every launch */
PQB_API int pqb_host_plan_pauli_tiles(const uint64_t* xmasks, const uint64_t* zmasks, size_t n_terms, int n_local_bits,
                                      int32_t* out_launch_of_term, uint64_t* out_tile_masks, size_t cap_launches,
                                      size_t* out_n_launches);
/* a fresh 128-byte ncclUniqueId (rank 0 creates it, the launcher's plumbing broadcasts it to the other ranks) */
PQB_API int pqb_nccl_unique_id(void* out128);
/* library version string */
PQB_API const char* pqb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PQB200_H_ */
