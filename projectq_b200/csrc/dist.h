// Sharded state across the GPUs of one NVSwitch box: one process per GPU, rank r holds the amplitudes whose rank bits
// spell r.  The reference has nothing like this (single address space, simulator.hpp:41); this is new work.
//
// Only dense gates with a *target* on a rank bit move data: the rank bit is exchanged with a local bit (a global<->local
// qubit remap), which is a pairwise half-shard exchange over NVLink (ncclSend/ncclRecv).  Controls and diagonal gates on
// rank bits never communicate.  Scalar results (probabilities, norms, expectation values, measurement bins) are summed
// with ncclAllReduce.  NCCL is loaded with dlopen so that the library has no link-time dependency on it.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <map>
#include <memory>
#include <set>
#include <vector>

#include "devmem.h"
#include "fdpass.h"
#include "fuser.h"
#include "kernels.cuh"

namespace pqb {

// Pure host logic of a remap, separated so it can be unit-tested without GPUs (pqb_host_plan_remap).
// loc[p] = placement of logical position p: local bit (< 64) or 64 + rank bit.  `need` lists logical positions that must
// become local.  Returns the (rank bit, local bit) pairs to exchange, choosing the highest local bits not in `need`
// (their halves are the largest contiguous blocks), and updates loc.  Throws std::runtime_error if it cannot be done.
// `victims` (optional) lists logical positions in eviction-preference order (the engine passes "needed last" first).
std::vector<std::pair<int, int>> plan_remap(std::vector<uint8_t>& loc, int n_local_bits, const std::vector<uint32_t>& need,
                                            const std::vector<uint32_t>* victims = nullptr);

// ---- when to remap and whom to evict (pure host logic, driven by Engine::run_sharded and, for CPU tests, by
// pqb_host_shard_schedule) ----------------------------------------------------------------------------------------------
using InteractionGraph = std::map<uint32_t, std::set<uint32_t>>;  // qubit id -> ids it shares a gate with
InteractionGraph interaction_graph(const Fuser& fuser);
struct RemapChoice {
    std::vector<uint32_t> need;     // logical positions that must be on-device next
    std::vector<uint32_t> victims;  // local logical positions in eviction-preference order
};
// Called when every pending gate waits for a qubit on a rank bit.  need = the qubits of the oldest waiting gate, plus the
// other rank-bit qubits that are needed sooner than the local qubits they would displace.  Eviction order: needed last
// first (Belady); ties (not needed again in this flush) go to a qubit that shares gates with a qubit already off-device,
// then to the one with the fewest interaction partners, then to the highest bit.
RemapChoice choose_remap(const Fuser& fuser, const std::map<uint32_t, uint32_t>& id_to_logical,
                         const std::vector<uint8_t>& loc, const InteractionGraph& adj, bool controls_needed = true);

// One flush of a sharded run, scheduled at PASS level.  The whole pending stream is scheduled once, exactly as on one GPU,
// so the number of fused passes does not depend on the sharding (scheduling gate by gate around the off-device qubits cost
// 2-4 extra passes per flush on the brickwork benchmark: the passes next to a remap came out half empty).  The passes are
// then executed in an order that respects their dependencies: everything whose target qubits are on-device runs before a
// remap is paid for, and the remap is chosen like before (choose_remap) with passes in the role of gates.
class ShardPlan {
public:
    ShardPlan(const Fuser& fuser, int max_qubits);
    bool finished() const { return n_left_ == 0; }
    size_t size() const { return clusters_.size(); }
    const Cluster& cluster(size_t i) const { return clusters_[i]; }
    // the passes that can run now (targets on local bits, all predecessors executed), in execution order; marks them executed
    std::vector<size_t> take_runnable(const std::map<uint32_t, uint32_t>& id_to_logical, const std::vector<uint8_t>& loc);
    // nothing is runnable: which qubits to bring on-device and whom to evict
    RemapChoice choose(const std::map<uint32_t, uint32_t>& id_to_logical, const std::vector<uint8_t>& loc,
                       const InteractionGraph& adj) const;

private:
    std::vector<Cluster> clusters_;
    std::vector<char> executed_;
    size_t n_left_ = 0;
};

// Who exchanges what in a (multi-bit) remap.  Exchanging rank bits r_i with local bits b_i sends, from every rank, the
// sub-block whose local bits (b_i) spell beta to the rank whose bits (r_i) spell beta, and the sub-block that arrives from
// that rank lands at the same addresses — so the remap is a set of pairwise in-place sub-block exchanges, one per peer
// (an all-to-all inside the group of 2^g ranks that differ only in the exchanged rank bits).
struct ExchangePeer {
    int peer;          // partner rank
    uint64_t pattern;  // values of the exchanged local bits (in place) of the sub-block swapped with that partner
};
std::vector<ExchangePeer> plan_exchange(int rank, const std::vector<std::pair<int, int>>& swaps);

class Dist {
public:
    Dist(int rank, int world, const void* nccl_unique_id, cudaStream_t stream, int device);
    ~Dist();

    int rank_bits() const { return g_; }
    // rank bits that carry no qubit yet: amplitude is non-zero only on ranks where all of them are 0
    uint64_t free_rank_bits_mask() const { return free_mask_; }
    bool has_free_rank_bit() const { return free_mask_ != 0; }
    int take_free_rank_bit();
    void reset_rank_bits(int n_used);  // rank bits [0, n_used) carry qubits, the rest are free
    void set_free_rank_bits_mask(uint64_t m) { free_mask_ = m & ((uint64_t(1) << g_) - 1); }  // restoring a checkpoint
    // the qubit on rank bit r was found classical with `value`: move the surviving shards onto bit r = 0, free the bit
    void release_rank_bit(int r, bool value, double2* shard, uint64_t n_amps);
    // exchange rank bit r with local bit b (n_local_bits = log2 n_amps); staging: >= staging_amps device amplitudes
    void swap_bits(int r, int b, double2* shard, int n_local_bits, double2* staging, uint64_t staging_amps,
                   uint64_t* bytes_sent);

    // exchange several (rank bit, local bit) pairs at once: every rank sends 1 - 2^-g of its shard instead of g/2
    void swap_bits_multi(const std::vector<std::pair<int, int>>& swaps, double2* shard, int n_local_bits, double2* staging,
                         uint64_t staging_amps, uint64_t* bytes_sent);

    // ---- peer-memory exchange (default; PQB_REMAP_P2P=0 switches it off) ------------------------------------------
    // Every rank maps the shards of the ranks it exchanges with (VMM handles exported as file descriptors and passed over
    // a Unix socket) and the whole remap — any number of (rank bit, local bit) pairs at once — is one kernel per rank
    // (k::peer_exchange) on a dedicated high-priority stream, ordered across GPUs by flags in peer-mapped sync pages.
    bool p2p_enabled() const { return p2p_ok_; }
    // Host-side preparation of an exchange of `swaps` on `state`: every member of the exchange group makes sure it has
    // every other member's current shard mapped, and the group agrees on the outcome.  false -> every member uses the NCCL
    // path for this remap.  Collective over the group (SPMD: all ranks call it with the same swaps).
    bool prepare_exchange(const std::vector<std::pair<int, int>>& swaps, const GrowBuffer& state, int device);
    // Enqueue the exchange of one slice of the shard on comm_stream() (the whole shard when slice.n == 0).  The caller
    // orders it after the passes that precede it (cudaStreamWaitEvent on comm_stream()) and waits for an event recorded
    // behind it before touching the slice again.
    void exchange_slice(const k::Slice& slice, int n_local_bits, int sm_count, uint64_t* bytes_sent);
    // `partner`'s buffer of the same role as `mine` (the ranks run the same program, so their buffers rotate alike), mapped
    // into this process; nullptr when peer mapping is not available.  Pairwise collective: the partner calls it with this
    // rank as its partner.
    const double2* peer_buffer(int partner, const GrowBuffer& mine);
    // rendezvous of all ranks on the engine's stream without blocking the host (a one-element all-reduce)
    void barrier_on_stream();
    cudaStream_t comm_stream() const { return comm_stream_; }
    // non-zero once a cross-GPU wait inside an exchange kernel timed out (the state is then undefined)
    int exchange_error() const { return h_error_ ? *h_error_ : 0; }

    double allreduce_sum(double v);
    void allreduce_sum_vec(double* v, size_t n);  // host vector, in place
    unsigned long long allreduce_min_u64(unsigned long long v);
    void barrier();

    static void get_unique_id(void* out128);

private:
    int rank_, world_, g_;
    uint64_t free_mask_;
    cudaStream_t stream_;
    cudaStream_t copy_stream_ = nullptr;  // staging -> shard copies, overlapped with the next NVLink transfer
    cudaEvent_t received_[2] = {nullptr, nullptr}, copied_[2] = {nullptr, nullptr}, packed_[2] = {nullptr, nullptr};
    cudaEvent_t ready_ = nullptr;
    void* comm_ = nullptr;
    struct PeerLink {
        // the partner's state buffers mapped into this process, by the partner's layout key (its three buffers rotate)
        std::vector<std::pair<uint64_t, std::unique_ptr<PeerMapping>>> maps;
        std::unique_ptr<PeerMapping> sync_map;  // the partner's sync page
        unsigned long long epoch = 0;           // exchanges done with this partner (the same number on both sides)
        PeerMapping* find(uint64_t key);
    };
    std::unique_ptr<FdChannel> fdchan_;
    std::map<int, PeerLink> links_;
    bool p2p_ok_ = false;
    int device_ = 0;
    cudaStream_t comm_stream_ = nullptr;       // exchange kernels (highest priority)
    GrowBuffer sync_page_;                      // arrive/done flags, mapped by every peer
    int* h_error_ = nullptr;                    // mapped pinned host word written by a kernel whose spin timed out
    int* d_error_ = nullptr;
    unsigned int* d_block_counter_ = nullptr;
    void setup_p2p();
    bool sync_mapping(int partner, const GrowBuffer& state, PeerMapping** out);
    void send_chunks(int partner, const GrowBuffer& buf);
    bool recv_chunks(int partner, PeerMapping& into);
    // the exchange prepared by prepare_exchange
    struct Prepared {
        std::vector<ExchangePeer> peers;
        std::vector<PeerMapping*> maps;
        std::vector<int> local_bits;  // ascending exchanged local bits
        uint64_t in_pattern = 0;
        double2* mine = nullptr;
    } cur_;
    void handshake(int peer);  // stream-ordered rendezvous with one partner (NCCL path)
    double* d_buf_ = nullptr;  // device staging for scalar collectives
    size_t d_buf_doubles_ = 0;
    void ensure_buf(size_t n);
};

}  // namespace pqb
