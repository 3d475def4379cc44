// Host controller of the B200 state-vector engine: qubit bookkeeping, RNG replay, fuser, kernel orchestration.
//
// Replaces `class Simulator` (reference: projectq/backends/_sim/_cppkernels/simulator.hpp:37-578).  The logical
// bookkeeping (id -> bit position, new qubit = new most-significant bit, positions shift down on deallocation) is the
// reference's, bit for bit, because cheat() exposes it.  Underneath, every logical position is placed on a physical bit:
// either a bit of the local amplitude index or, in a sharded run, a bit of the rank number.
#pragma once
#include <cuda_runtime.h>

#include <complex>
#include <cstdint>
#include <map>
#include <memory>
#include <random>
#include <string>
#include <vector>

#include "../../include/pqb200.h"
#include "devmem.h"
#include "fuser.h"
#include "kernels.cuh"

namespace pqb {

// error classes mirrored onto pqb_status by capi.cpp
struct RuntimeErr : std::runtime_error {
    using std::runtime_error::runtime_error;
};
struct ValueErr : std::runtime_error {
    using std::runtime_error::runtime_error;
};
struct CudaErr : std::runtime_error {
    using std::runtime_error::runtime_error;
};

class Dist;  // sharded-state communicator (dist.h)
class ShardPlan;

// RAII: make `device` the calling thread's current CUDA device, restore the previous one on scope exit
class DeviceGuard {
public:
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev_) != cudaSuccess) {
            cudaGetLastError();
            prev_ = -1;
        }
        if (prev_ != device) cudaSetDevice(device);
        else prev_ = -1;  // nothing to restore
    }
    ~DeviceGuard() {
        if (prev_ >= 0) cudaSetDevice(prev_);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;

private:
    int prev_ = -1;
};


struct TermsView {
    size_t n_terms;
    const size_t* offsets;
    const uint32_t* qubit_index;
    const char* pauli;
    const double* coeff;  // real: n_terms, complex: 2*n_terms
    bool complex_coeff;
};

// How a Pauli-string operator is executed (engine.cpp plan_pauli_tiles): a few launches of the tiled kernel, each with its
// own set of tile bits, plus the terms whose X/Y support no tile can hold.  Pure host logic (CPU-tested through
// pqb_host_plan_pauli_tiles).
struct PauliPlan {
    std::vector<k::PauliTileArgs> launches;  // in execution order; first/final/scale are filled in by the caller
    std::vector<long> table_at;              // per launch: offset of its diagonal table in `tables`, or -1
    std::vector<uint64_t> tile_mask;         // per launch: its tile bits
    std::vector<double2> tables;             // 2^T entries per table
    std::vector<k::PauliTerm> wide;
    std::vector<int> launch_of_term;         // per input term: the launch that applies it, -1 = wide
    const k::PauliTerm* d_wide = nullptr;
    int block_bits = 0;                      // launches whose tile bits all lie below this bit may run fused (0: none)
    bool fusable(size_t i, int n_local_bits) const {
        return block_bits > 0 && (n_local_bits <= block_bits || (tile_mask[i] >> block_bits) == 0);
    }
};
// block size (log2 amplitudes) of the fused, L2-resident execution of tile-bit sets; PQB_PAULI_BLOCK_BITS overrides (0 = off)
int pauli_block_bits();
PauliPlan plan_pauli_tiles(const std::vector<k::PauliTerm>& terms, int n_local_bits, int block_bits);

class Engine {
public:
    Engine(uint32_t seed, const pqb_opts& opts);
    ~Engine();

    void allocate_qubit(uint32_t id);
    void deallocate_qubit(uint32_t id);
    bool get_classical_value(uint32_t id, double tol);
    bool is_classical(uint32_t id, double tol);
    void measure_qubits(const uint32_t* ids, size_t n, uint8_t* out);
    void apply_controlled_gate(const double* m, const uint32_t* ids, size_t k, const uint32_t* ctrl, size_t nc);
    void emulate_math(int mode, int64_t a, int64_t N, const uint64_t* table, size_t table_len, const uint32_t* reg_ids,
                      const uint32_t* reg_sizes, size_t n_regs, const uint32_t* ctrl, size_t nc);
    double get_expectation_value(const TermsView& t, const uint32_t* ids, size_t n_ids);
    void apply_qubit_operator(const TermsView& t, const uint32_t* ids, size_t n_ids);
    void emulate_time_evolution(const TermsView& t, double time, const uint32_t* ids, size_t n_ids, const uint32_t* ctrl,
                                size_t nc);
    double get_probability(const uint8_t* bits, const uint32_t* ids, size_t n);
    std::complex<double> get_amplitude(const uint8_t* bits, const uint32_t* ids, size_t n);
    void set_wavefunction(const double* wf, size_t n_amps, const uint32_t* ordering, size_t n);
    void collapse_wavefunction(const uint32_t* ids, size_t n_ids, const uint8_t* values, size_t n_values);
    void run();
    size_t num_qubits() const { return size_t(n_); }
    int device() const { return device_; }
    size_t cheat_map(uint32_t* ids, uint32_t* pos, size_t cap);
    void cheat_state(double* out, size_t cap_amps);

    // additions
    void get_amplitudes(const uint64_t* logical_idx, size_t n, double* out);
    void apply_gate_stream(const void* packed, size_t n_bytes, size_t n_gates, bool fuse);
    void init_random_state(uint32_t n_qubits, uint64_t seed);
    void save_state(const std::string& prefix);
    void load_state(const std::string& prefix);
    void state_view(void** ptr, uint64_t* n_amps, uint8_t* layout, size_t cap, size_t* n_qubits);
    double norm_squared();
    void synchronize();
    void timer_start();
    double timer_stop();
    void get_stats(pqb_stats* out);
    void reset_stats();
    void set_profiling(bool on) { profiling_ = on; }
    void flush_l2(size_t bytes);
    double bench_dense_pass(const double* m, const uint32_t* positions, size_t k, uint64_t ctrl_mask, int repeats);
    void selftest_sliced_pass(const double* m, const uint32_t* positions, size_t k, uint64_t ctrl_mask, uint64_t slice_mask);
    double measure_fp64_peak();
    double measure_copy_bandwidth(size_t bytes);

private:
    k::Ctx ctx() { return k::Ctx{stream_, &stats_.kernel_launches}; }
    uint64_t local_amps() const { return uint64_t(1) << L_; }
    double2* psi() { return state_->amps(); }
    uint32_t pos_of(uint32_t id, const char* what) const;
    bool known(uint32_t id) const { return map_.count(id) != 0; }
    // physical placement
    bool is_local(uint32_t logical_pos) const { return loc_[logical_pos] < 64; }
    uint64_t logical_to_local_index(uint64_t logical_index, bool* mine) const;
    bool layout_is_identity() const;
    // split a logical (mask, val) into the local part; returns false if this rank's bits contradict val
    bool split_mask(uint64_t lmask, uint64_t lval, uint64_t* local_mask, uint64_t* local_val) const;
    double read_scalar(const double* d_ptr);
    double allreduce_sum(double v);
    void ensure_scratch(GrowBuffer& b, size_t bytes);
    std::vector<k::PauliTerm> build_terms(const TermsView& t, const uint32_t* ids, size_t n_ids, bool skip_identity,
                                          double* identity_sum_re, double* identity_sum_im);
    void make_local(const std::vector<uint32_t>& logical_positions, const std::vector<uint32_t>* victims = nullptr);
    // A fused pass resolved to physical bit positions under the qubit layout that was current when it was resolved; it can
    // be launched later (and slice by slice) whatever has happened to the layout bookkeeping in between.
    struct Launch {
        enum Kind { NONE, DENSE, DIAG } kind = NONE;  // NONE: this rank's control bits switch the pass off
        int k = 0, n_ctrl = 0;
        uint8_t tpos[8] = {}, cpos[64] = {};
        std::vector<double> m;  // DENSE: 2^k x 2^k (re,im) row-major; DIAG: 2^k (re,im)
        uint64_t touched() const;  // mask of the local bits the pass uses as target or control (rank-independent)
    };
    Launch resolve_pass(const FusedPass& p);
    void launch(const Launch& l, const k::Slice& slice = k::Slice());
    void apply_pass(const FusedPass& p) { launch(resolve_pass(p)); }
    void run_sharded();
    // schedule everything that can run under the current layout, fuse it and resolve it into `out` (appended); with
    // `eager` all but the last `hold` launches are issued right away (they cannot be part of a remap pipeline's tail)
    void resolve_phase(ShardPlan& plan, std::vector<Launch>& out, bool eager, size_t hold);
    void serial_exchange(const std::vector<std::pair<int, int>>& swaps);
    void pipelined_exchange(const std::vector<Launch>& tail, const std::vector<Launch>& head,
                            const std::vector<uint8_t>& slice_bits);
    // An operator as the engine executes it: its terms grouped by the rank whose shard holds their partner amplitudes.
    // (P psi)[j] needs psi[j ^ xmask]; the local part of xmask stays inside this rank's shard, an X/Y on a qubit that sits on a
    // rank bit means the partner amplitude lives at the same local index on rank ^ xr.  Group 0 (xr = 0) reads this rank's own
    // vector; the others read the partner's vector through peer-mapped memory (NVLink), so an operator that flips more qubits
    // than one shard holds — a transverse field on every qubit of a sharded state — needs no remap at all.
    struct PauliProgram {
        struct Group {
            int xr = 0;
            PauliPlan plan;
        };
        std::vector<Group> groups;
        bool reads_peers() const { return groups.size() > 1 || (!groups.empty() && groups[0].xr != 0); }
    };
    PauliProgram build_pauli_program(const std::vector<k::PauliTerm>& logical_terms);
    std::vector<const double2*> pauli_sources(const PauliProgram& prog, const GrowBuffer& buf);
    // the tile launches [first, last) of one plan, consecutive fusable ones in one launch
    void run_tile_launches(PauliPlan& plan, size_t first, size_t last, const double2* in, double2* u, double2* acc, double* d_sum);
    void run_pauli_program(PauliProgram& prog, const std::vector<const double2*>& src, double2* u, double sre, double sim,
                           double2* acc, uint64_t cmask, double* d_norm);
    void check_exchange_error();
    bool leaves_from_low_bit(int local_bit, size_t n_swaps) const;
    // CUDA events: a pool, and (start, stop) pairs whose elapsed time is added to a counter once they have completed
    enum TimerKind { T_PASS0 = 0, T_DIAG = 6, T_STALL = 7, T_COMM = 8 };
    struct Timer {
        cudaEvent_t e0, e1;
        int kind;
    };
    cudaEvent_t get_event();
    void wait_on_main(cudaEvent_t ev);  // main stream waits for ev; the wait is timed as exposed remap time
    void harvest_timers(bool synchronize);
    std::vector<cudaEvent_t> free_events_, used_events_;
    std::vector<Timer> timers_;
    bool profiling_ = false;
    unsigned long long last_probe_[2] = {~0ULL, ~0ULL};  // result of the last classical probe (bit 0, bit 1)
    double draw_uniform();

    // device
    int device_ = 0;
    int sm_count_ = 148;
    cudaStream_t stream_ = nullptr;
    cudaEvent_t ev0_ = nullptr, ev1_ = nullptr;
    cudaEvent_t remap_e0_ = nullptr, remap_e1_ = nullptr;
    GrowBuffer buf_[3];
    GrowBuffer* state_ = &buf_[0];
    GrowBuffer* scratch1_ = &buf_[1];
    GrowBuffer* scratch2_ = &buf_[2];
    double* d_partials_ = nullptr;         // kReducePartials doubles
    unsigned* d_pauli_sync_ = nullptr;     // kPauliSyncWords words: ticket and block counters of the fused Pauli launches
    double* d_scalars_ = nullptr;          // small device scratch (bins, accumulators)
    double* h_pinned_ = nullptr;           // pinned host mirror of d_scalars_
    void* d_small_ = nullptr;              // terms / indices / tables upload area
    size_t d_small_cap_ = 0;
    double* d_flush_ = nullptr;
    size_t d_flush_cap_ = 0;
    void* small_upload(const void* src, size_t bytes);

    // logical bookkeeping (reference: map_, N_)
    std::map<uint32_t, uint32_t> map_;  // qubit id -> logical bit position
    int n_ = 0;
    // physical placement
    int L_ = 0;                  // local index bits
    std::vector<uint8_t> loc_;   // logical position -> local bit (< 64) or 64 + rank bit

    std::mt19937 rng_;
    Fuser fuser_;
    int fusion_max_ = 5;
    pqb_stats stats_{};
    std::unique_ptr<Dist> dist_;
    int rank_ = 0, world_ = 1;
};

}  // namespace pqb
