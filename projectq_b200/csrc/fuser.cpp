// Host-side gate fuser — see fuser.h.  Pure C++ (no CUDA), unit-tested on the CPU through pqb_host_fuse_stream.
#include "fuser.h"

#include <algorithm>
#include <stdexcept>
#include <unordered_map>

namespace pqb {

namespace {

constexpr size_t kLookahead = 768;  // how far past the oldest pending gate the scheduler searches for absorbable gates

inline bool contains(const std::vector<uint32_t>& v, uint32_t x) { return std::find(v.begin(), v.end(), x) != v.end(); }

// width of the pass if gate g joined a non-empty cluster with targets S and common controls G
int width_after(const std::vector<uint32_t>& S, const std::vector<uint32_t>& G, const Gate& g) {
    int w = int(S.size());
    auto in_s = [&](uint32_t q) { return contains(S, q); };
    uint32_t added[16];
    int n_added = 0;
    auto add = [&](uint32_t q) {
        if (in_s(q)) return;
        for (int i = 0; i < n_added; ++i)
            if (added[i] == q) return;
        if (n_added < 16) added[n_added] = q;
        ++n_added;
    };
    for (auto t : g.targets) add(t);
    for (auto c : g.ctrls)
        if (!contains(G, c)) add(c);  // control not common to the cluster -> becomes a target
    for (auto c : G)
        if (!contains(g.ctrls, c)) add(c);  // common control the new gate lacks -> demoted to a target
    return w + n_added;
}

void absorb(std::vector<uint32_t>& S, std::vector<uint32_t>& G, const Gate& g, bool first) {
    if (first) {
        S = g.targets;
        G = g.ctrls;
        return;
    }
    std::vector<uint32_t> keep;
    for (auto c : G) {
        if (contains(g.ctrls, c))
            keep.push_back(c);
        else if (!contains(S, c))
            S.push_back(c);
    }
    for (auto c : g.ctrls)
        if (!contains(keep, c) && !contains(S, c)) S.push_back(c);
    for (auto t : g.targets)
        if (!contains(S, t)) S.push_back(t);
    G.swap(keep);
}

}  // namespace

FusedPass Fuser::fuse(const std::vector<const Gate*>& gates, const std::function<uint64_t(uint32_t)>& sort_key) {
    if (gates.empty()) throw std::invalid_argument("fuse(): empty gate list");
    // common controls = intersection of all control lists (fusion.hpp:130-160 keeps exactly these in ctrl_set_)
    std::vector<uint32_t> G = gates[0]->ctrls;
    for (size_t i = 1; i < gates.size(); ++i) {
        std::vector<uint32_t> keep;
        for (auto c : G)
            if (contains(gates[i]->ctrls, c)) keep.push_back(c);
        G.swap(keep);
    }
    std::vector<uint32_t> S;
    for (auto* g : gates) {
        for (auto t : g->targets)
            if (!contains(S, t)) S.push_back(t);
        for (auto c : g->ctrls)
            if (!contains(G, c) && !contains(S, c)) S.push_back(c);
    }
    for (auto c : G)
        if (contains(S, c)) throw std::invalid_argument("fuse(): a qubit is both a common control and a target");
    std::sort(S.begin(), S.end(), [&](uint32_t a, uint32_t b) {
        const uint64_t ka = sort_key(a), kb = sort_key(b);
        return ka != kb ? ka < kb : a < b;
    });
    const int w = int(S.size());
    if (w > 12) throw std::invalid_argument("fuse(): pass too wide");
    const size_t D = size_t(1) << w;
    FusedPass out;
    out.targets = S;
    out.ctrls = G;
    std::sort(out.ctrls.begin(), out.ctrls.end());
    out.n_gates = gates.size();
    out.m.assign(D * D, cplx(0.0, 0.0));
    for (size_t i = 0; i < D; ++i) out.m[i * D + i] = 1.0;

    std::vector<cplx> tmp;
    for (auto* g : gates) {
        const int k = int(g->targets.size());
        const size_t d = size_t(1) << k;
        if (g->m.size() != d * d) throw std::invalid_argument("fuse(): matrix size does not match the target count");
        unsigned tb[16];
        size_t tmask = 0, cm = 0;
        for (int l = 0; l < k; ++l) {
            tb[l] = unsigned(std::find(S.begin(), S.end(), g->targets[l]) - S.begin());
            tmask |= size_t(1) << tb[l];
        }
        for (auto c : g->ctrls)
            if (!contains(G, c)) cm |= size_t(1) << unsigned(std::find(S.begin(), S.end(), c) - S.begin());
        // offsets of the 2^k group members inside the w-bit row index
        std::vector<size_t> off(d, 0);
        for (size_t j = 0; j < d; ++j)
            for (int l = 0; l < k; ++l)
                if ((j >> l) & 1) off[j] |= size_t(1) << tb[l];
        tmp.resize(d);
        // left-multiply the running product by the (expanded) gate: transform every column like a w-qubit state
        for (size_t col = 0; col < D; ++col) {
            for (size_t r = 0; r < D; ++r) {
                if ((r & tmask) != 0 || (r & cm) != cm) continue;
                for (size_t j = 0; j < d; ++j) tmp[j] = out.m[(r | off[j]) * D + col];
                for (size_t i = 0; i < d; ++i) {
                    cplx acc(0.0, 0.0);
                    for (size_t j = 0; j < d; ++j) acc += g->m[i * d + j] * tmp[j];
                    out.m[(r | off[i]) * D + col] = acc;
                }
            }
        }
    }
    out.diagonal = true;
    for (size_t r = 0; r < D && out.diagonal; ++r)
        for (size_t c = 0; c < D; ++c)
            if (r != c && (out.m[r * D + c].real() != 0.0 || out.m[r * D + c].imag() != 0.0)) {
                out.diagonal = false;
                break;
            }
    return out;
}

void Fuser::reorder(FusedPass& p, const std::function<uint64_t(uint32_t)>& sort_key) {
    const int w = int(p.targets.size());
    std::vector<int> order(w);
    for (int i = 0; i < w; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int a, int b) {
        const uint64_t ka = sort_key(p.targets[a]), kb = sort_key(p.targets[b]);
        return ka != kb ? ka < kb : p.targets[a] < p.targets[b];
    });
    bool same = true;
    for (int i = 0; i < w; ++i) same = same && order[i] == i;
    if (same) return;
    // new matrix bit i is old matrix bit order[i]
    const size_t D = size_t(1) << w;
    auto old_index = [&](size_t x) {
        size_t r = 0;
        for (int i = 0; i < w; ++i) r |= ((x >> i) & 1) << order[i];
        return r;
    };
    std::vector<cplx> m(D * D);
    for (size_t r = 0; r < D; ++r)
        for (size_t c = 0; c < D; ++c) m[r * D + c] = p.m[old_index(r) * D + old_index(c)];
    std::vector<uint32_t> t(w);
    for (int i = 0; i < w; ++i) t[i] = p.targets[order[i]];
    p.m.swap(m);
    p.targets.swap(t);
}

std::vector<FusedPass> Fuser::drain(int max_qubits, const std::function<uint64_t(uint32_t)>& sort_key) {
    std::vector<FusedPass> passes = plan(max_qubits, sort_key);
    pending_.clear();
    return passes;
}

std::vector<FusedPass> Fuser::plan(int max_qubits, const std::function<uint64_t(uint32_t)>& sort_key) const {
    std::vector<char> done;
    std::vector<FusedPass> passes;
    for (auto& cl : schedule_impl(max_qubits, nullptr, done)) passes.push_back(fuse_cluster(cl, sort_key));
    return passes;
}

std::vector<Cluster> Fuser::schedule(int max_qubits) const {
    std::vector<char> done;
    return schedule_impl(max_qubits, nullptr, done);
}

std::vector<Cluster> Fuser::schedule_unblocked(int max_qubits, const std::function<bool(uint32_t)>& blocked,
                                               std::vector<char>& done) const {
    return schedule_impl(max_qubits, &blocked, done);
}

FusedPass Fuser::fuse_cluster(const Cluster& cl, const std::function<uint64_t(uint32_t)>& sort_key) const {
    std::vector<const Gate*> members;
    members.reserve(cl.gates.size());
    for (auto g : cl.gates) members.push_back(&pending_[g]);
    return fuse(members, sort_key);
}

void Fuser::remove_done(const std::vector<char>& done) {
    std::vector<Gate> rest;
    for (size_t g = 0; g < pending_.size(); ++g)
        if (g >= done.size() || !done[g]) rest.push_back(std::move(pending_[g]));
    pending_.swap(rest);
}

std::vector<FusedPass> Fuser::drain_unblocked(int max_qubits, const std::function<uint64_t(uint32_t)>& sort_key,
                                              const std::function<bool(uint32_t)>& blocked) {
    std::vector<char> done;
    std::vector<FusedPass> passes;
    for (auto& cl : schedule_impl(max_qubits, &blocked, done)) passes.push_back(fuse_cluster(cl, sort_key));
    remove_done(done);
    return passes;
}

size_t Fuser::next_use(uint32_t id) const {
    for (size_t g = 0; g < pending_.size(); ++g) {
        if (contains(pending_[g].targets, id) || contains(pending_[g].ctrls, id)) return g;
    }
    return size_t(-1);
}

std::vector<Cluster> Fuser::schedule_impl(int max_qubits, const std::function<bool(uint32_t)>* blocked,
                                          std::vector<char>& done) const {
    std::vector<Cluster> passes;
    const size_t m = pending_.size();
    done.assign(m, 0);
    if (m == 0) return passes;
    // gates that touch a blocked qubit (one that sits on a rank bit of the sharded state) are left for later
    std::vector<char> is_blocked(m, 0);
    if (blocked)
        for (size_t g = 0; g < m; ++g) {
            for (auto t : pending_[g].targets)
                if ((*blocked)(t)) is_blocked[g] = 1;
            for (auto c : pending_[g].ctrls)
                if ((*blocked)(c)) is_blocked[g] = 1;
        }

    // per-qubit ordered lists of the gates touching it (targets and controls both order gates)
    std::unordered_map<uint32_t, uint32_t> dense;
    std::vector<std::vector<uint32_t>> touch;  // touch[qubit] = gate indices in program order
    std::vector<std::vector<uint32_t>> gq(m);  // gq[gate] = dense qubit numbers
    for (size_t g = 0; g < m; ++g) {
        auto reg = [&](uint32_t id) {
            auto it = dense.find(id);
            uint32_t d;
            if (it == dense.end()) {
                d = uint32_t(touch.size());
                dense.emplace(id, d);
                touch.emplace_back();
            } else
                d = it->second;
            touch[d].push_back(uint32_t(g));
            gq[g].push_back(d);
        };
        for (auto t : pending_[g].targets) reg(t);
        for (auto c : pending_[g].ctrls) reg(c);
    }
    std::vector<size_t> head(touch.size(), 0);
    auto ready = [&](size_t g) {
        for (auto d : gq[g]) {
            size_t& h = head[d];
            while (h < touch[d].size() && done[touch[d][h]]) ++h;
            if (h >= touch[d].size() || touch[d][h] != g) return false;
        }
        return true;
    };

    size_t first = 0;
    std::vector<uint32_t> S, G;
    std::vector<uint32_t> members;
    while (true) {
        while (first < m && done[first]) ++first;
        if (first >= m) break;
        // seed = oldest gate that can run now (with blocked qubits around, not necessarily the oldest pending one)
        size_t seed = first;
        if (blocked) {
            seed = m;
            for (size_t g = first; g < m; ++g)
                if (!done[g] && !is_blocked[g] && ready(g)) {
                    seed = g;
                    break;
                }
            if (seed == m) break;  // everything left waits for a blocked qubit
        }
        members.clear();
        S.clear();
        G.clear();
        absorb(S, G, pending_[seed], true);
        members.push_back(uint32_t(seed));
        done[seed] = 1;
        // grow the pass: prefer gates that fit without widening it, then the smallest widening, then program order
        while (true) {
            const size_t limit = std::min(m, seed + kLookahead);
            size_t best = m;
            int best_w = max_qubits + 1;
            const int cur_w = int(S.size());
            for (size_t g = first + 1; g < limit; ++g) {
                if (done[g] || is_blocked[g] || !ready(g)) continue;
                const int w = width_after(S, G, pending_[g]);
                if (w > max_qubits) continue;
                if (w < best_w) {
                    best_w = w;
                    best = g;
                    if (w == cur_w) break;
                }
            }
            if (best == m) break;
            absorb(S, G, pending_[best], false);
            members.push_back(uint32_t(best));
            done[best] = 1;
        }
        passes.push_back(Cluster{members, int(S.size()), int(G.size()), S, G});
    }
    return passes;
}

}  // namespace pqb
