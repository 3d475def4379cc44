// Host-side gate fuser: merges queued (controlled) gates into dense k<=5-qubit passes.
//
// Replaces the reference's `class Fusion` (reference: _cppkernels/fusion.hpp:38-165) and the flush policy in
// `Simulator::apply_controlled_gate` (reference: _cppkernels/simulator.hpp:204-222).  Same algebra — a pass is the
// ordered product of its gates, controls common to every gate of a pass stay a control mask, any other control is
// folded into the matrix as an extra target (fusion.hpp:112-160) — but a different policy: the reference fires as soon
// as the queue spans 4 qubits, so it never builds 5-qubit passes and never looks past the next gate.  This fuser keeps
// the whole pending stream, treats it as a dependency graph (gates on disjoint qubits commute) and grows each pass
// greedily from the ready frontier up to `max_qubits`, which puts ~15 brickwork gates into one pass instead of ~4.
#pragma once
#include <complex>
#include <cstdint>
#include <functional>
#include <vector>

namespace pqb {

using cplx = std::complex<double>;

struct Gate {
    std::vector<uint32_t> targets;  // matrix bit l <-> targets[l]
    std::vector<uint32_t> ctrls;
    std::vector<cplx> m;            // 2^k x 2^k row-major
};

struct FusedPass {
    std::vector<uint32_t> targets;  // qubit ids, ascending sort key; matrix bit l <-> targets[l]
    std::vector<uint32_t> ctrls;    // controls common to every gate of the pass (kernel control mask)
    std::vector<cplx> m;            // 2^k x 2^k row-major
    size_t n_gates = 0;
    bool diagonal = false;          // every off-diagonal entry is exactly zero
};

// one scheduled pass before its matrix is built: which pending gates it contains, its width and common controls
struct Cluster {
    std::vector<uint32_t> gates;  // indices into the pending queue, in application order
    int width = 0;
    int n_ctrl = 0;
    std::vector<uint32_t> targets;  // qubit ids the fused matrix acts on (unordered)
    std::vector<uint32_t> ctrls;    // controls common to every gate of the pass
};

class Fuser {
public:
    void push(Gate g) { pending_.push_back(std::move(g)); }
    size_t pending() const { return pending_.size(); }
    void clear() { pending_.clear(); }

    // Schedule every pending gate into passes of at most max_qubits targets (gates wider than that keep their own
    // width).  sort_key(id) orders the target qubits of a pass (the engine passes the bit position).
    std::vector<FusedPass> drain(int max_qubits, const std::function<uint64_t(uint32_t)>& sort_key);
    // Same schedule, but the pending gates stay queued (used to compare fusion widths before committing to one).
    std::vector<FusedPass> plan(int max_qubits, const std::function<uint64_t(uint32_t)>& sort_key) const;

    // Two-step form used by the engine: schedule first (cheap: no matrix products), then build and launch one pass at a
    // time, so that the GPU works on pass i while the host multiplies the matrices of pass i+1.
    std::vector<Cluster> schedule(int max_qubits) const;
    std::vector<Cluster> schedule_unblocked(int max_qubits, const std::function<bool(uint32_t)>& blocked,
                                            std::vector<char>& done) const;
    FusedPass fuse_cluster(const Cluster& cl, const std::function<uint64_t(uint32_t)>& sort_key) const;
    void remove_done(const std::vector<char>& done);  // drop the gates a schedule_unblocked call marked

    // Sharded runs: schedule (and remove from the queue) only what can run without touching a blocked qubit; gates that
    // touch one, and everything that depends on them, stay queued in program order.
    std::vector<FusedPass> drain_unblocked(int max_qubits, const std::function<uint64_t(uint32_t)>& sort_key,
                                           const std::function<bool(uint32_t)>& blocked);
    // index of the first pending gate that touches qubit `id` (size_t(-1) if none): the remap planner evicts the local
    // qubit that is needed last
    size_t next_use(uint32_t id) const;
    const Gate& pending_gate(size_t i) const { return pending_[i]; }

    // Fuse an explicit list of gates (already chosen to fit) into one pass.
    static FusedPass fuse(const std::vector<const Gate*>& gates, const std::function<uint64_t(uint32_t)>& sort_key);

    // re-sort the targets of a pass by a (changed) sort key, permuting the matrix bits accordingly
    static void reorder(FusedPass& p, const std::function<uint64_t(uint32_t)>& sort_key);

private:
    std::vector<Cluster> schedule_impl(int max_qubits, const std::function<bool(uint32_t)>* blocked,
                                       std::vector<char>& done) const;
    std::vector<Gate> pending_;
};

}  // namespace pqb
