// Host controller — see engine.h.
#include "engine.h"

#include <algorithm>
#include <cmath>
#include <set>
#include <sstream>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "bits.h"
#include "dist.h"

namespace pqb {

#define PQB_CHECK(expr)                                                                                       \
    do {                                                                                                      \
        cudaError_t err__ = (expr);                                                                           \
        if (err__ != cudaSuccess)                                                                             \
            throw CudaErr(std::string("CUDA error: ") + cudaGetErrorString(err__) + " (" #expr ") at " +      \
                          __FILE__ + ":" + std::to_string(__LINE__));                                         \
    } while (0)

namespace {
constexpr size_t kScalarDoubles = 8192;  // device/pinned scalar scratch (bins of the measurement search live here)
constexpr size_t kMaxPending = 4096;     // gates buffered before the fuser is drained on its own
constexpr int kBinBits = 10;             // measurement search: logical bits resolved per level
constexpr int kMinLocalBits = 8;         // sharded runs: qubits go to local bits until a shard holds this many, then to
                                         // free rank bits (so the widest gate plus its eviction room always fits on-device)
}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// construction
// ---------------------------------------------------------------------------------------------------------------
namespace {
// The per-engine stream, events and small device / pinned scratch are recycled between engines of one process (a ProjectQ
// program creates a Simulator per MainEngine; creating these cost 1-2 ms per engine, cudaMallocHost alone a good part of it).
struct EngineKit {
    int device;
    cudaStream_t stream;
    cudaEvent_t ev0, ev1;
    double* d_partials;
    unsigned* d_pauli_sync;
    double* d_scalars;
    double* h_pinned;
};
constexpr size_t kMaxKits = 4;
std::mutex g_kit_mutex;
std::vector<EngineKit> g_kits;
}  // namespace

Engine::Engine(uint32_t seed, const pqb_opts& o) : rng_(seed) {
    device_ = o.device;
    fusion_max_ = o.fusion_max_qubits <= 0 ? 0 : std::min(o.fusion_max_qubits, 5);  // 0 = choose per flush
    rank_ = o.rank;
    world_ = o.world_size <= 1 ? 1 : o.world_size;
    if (world_ & (world_ - 1)) throw ValueErr("pqb_create: world_size must be a power of two");
    if (rank_ < 0 || rank_ >= world_) throw ValueErr("pqb_create: rank out of range");

    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        throw CudaErr("pqb_create: no CUDA device available (this engine has no CPU fallback)");
    }
    if (device_ < 0 || device_ >= count) throw CudaErr("pqb_create: CUDA device ordinal out of range");
    PQB_CHECK(cudaSetDevice(device_));
    // (cudaGetDeviceProperties takes milliseconds; small circuits create engines often)
    PQB_CHECK(cudaDeviceGetAttribute(&sm_count_, cudaDevAttrMultiProcessorCount, device_));
    bool recycled = false;
    {
        std::lock_guard<std::mutex> lock(g_kit_mutex);
        for (size_t i = 0; i < g_kits.size(); ++i) {
            if (g_kits[i].device != device_) continue;
            const EngineKit kit = g_kits[i];
            g_kits.erase(g_kits.begin() + long(i));
            stream_ = kit.stream;
            ev0_ = kit.ev0;
            ev1_ = kit.ev1;
            d_partials_ = kit.d_partials;
            d_pauli_sync_ = kit.d_pauli_sync;
            d_scalars_ = kit.d_scalars;
            h_pinned_ = kit.h_pinned;
            recycled = true;
            break;
        }
    }
    if (!recycled) {
        PQB_CHECK(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
        PQB_CHECK(cudaEventCreate(&ev0_));
        PQB_CHECK(cudaEventCreate(&ev1_));
        PQB_CHECK(cudaMalloc(&d_partials_, sizeof(double) * k::kReducePartials));
        PQB_CHECK(cudaMalloc(&d_pauli_sync_, sizeof(unsigned) * k::kPauliSyncWords));
        PQB_CHECK(cudaMalloc(&d_scalars_, sizeof(double) * kScalarDoubles));
        PQB_CHECK(cudaMallocHost(&h_pinned_, sizeof(double) * kScalarDoubles));
    }
    // only the state buffer reserves its address range now; the two scratch buffers do so on first use (ensure_scratch)
    state_->init(device_, /*exportable=*/world_ > 1);

    if (world_ > 1) {
        if (!o.nccl_unique_id) throw ValueErr("pqb_create: nccl_unique_id is required when world_size > 1");
        dist_.reset(new Dist(rank_, world_, o.nccl_unique_id, stream_, device_));
    }

    // |psi> = 1 on one amplitude (simulator.hpp:48-50).  In a sharded run every rank bit starts free: the single
    // amplitude lives on rank 0 and the other ranks hold an (identically sized) all-zero shard.
    try {
        state_->ensure(std::max<size_t>(sizeof(double2), o.reserve_qubits > 0 && o.reserve_qubits < 40
                                                             ? (sizeof(double2) << o.reserve_qubits) / size_t(world_)
                                                             : 0));
    } catch (const std::bad_alloc&) {
        throw CudaErr("pqb_create: out of device memory");
    }
    const double2 one = make_double2(rank_ == 0 ? 1.0 : 0.0, 0.0);
    PQB_CHECK(cudaMemcpyAsync(psi(), &one, sizeof(one), cudaMemcpyHostToDevice, stream_));
    PQB_CHECK(cudaStreamSynchronize(stream_));
}

Engine::~Engine() {
    if (stream_) cudaStreamSynchronize(stream_);
    if (dist_ && dist_->comm_stream()) cudaStreamSynchronize(dist_->comm_stream());
    for (auto e : used_events_) cudaEventDestroy(e);
    for (auto e : free_events_) cudaEventDestroy(e);
    dist_.reset();
    if (d_small_) cudaFree(d_small_);
    if (d_flush_) cudaFree(d_flush_);
    if (remap_e0_) cudaEventDestroy(remap_e0_);
    if (remap_e1_) cudaEventDestroy(remap_e1_);
    // (the buffers park or release their memory in their own destructors, after this body)
    bool kept = false;
    if (stream_ && ev0_ && ev1_ && d_partials_ && d_pauli_sync_ && d_scalars_ && h_pinned_) {
        std::lock_guard<std::mutex> lock(g_kit_mutex);
        if (g_kits.size() < kMaxKits) {
            g_kits.push_back({device_, stream_, ev0_, ev1_, d_partials_, d_pauli_sync_, d_scalars_, h_pinned_});
            kept = true;
        }
    }
    if (!kept) {
        if (d_partials_) cudaFree(d_partials_);
        if (d_pauli_sync_) cudaFree(d_pauli_sync_);
        if (d_scalars_) cudaFree(d_scalars_);
        if (h_pinned_) cudaFreeHost(h_pinned_);
        if (ev0_) cudaEventDestroy(ev0_);
        if (ev1_) cudaEventDestroy(ev1_);
        if (stream_) cudaStreamDestroy(stream_);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------------------------
uint32_t Engine::pos_of(uint32_t id, const char* what) const {
    auto it = map_.find(id);
    if (it == map_.end()) throw RuntimeErr(std::string(what));
    return it->second;
}

bool Engine::layout_is_identity() const {
    for (int p = 0; p < n_; ++p)
        if (loc_[p] != p) return false;
    return true;
}

uint64_t Engine::logical_to_local_index(uint64_t logical_index, bool* mine) const {
    uint64_t local = 0;
    bool ok = true;
    for (int p = 0; p < n_; ++p) {
        const uint64_t bit = (logical_index >> p) & 1;
        if (loc_[p] < 64)
            local |= bit << loc_[p];
        else if (uint64_t((rank_ >> (loc_[p] - 64)) & 1) != bit)
            ok = false;
    }
    // free rank bits hold amplitude only where they are 0
    if (dist_ && (uint64_t(rank_) & dist_->free_rank_bits_mask()) != 0) ok = false;
    if (mine) *mine = ok;
    return local;
}

bool Engine::split_mask(uint64_t lmask, uint64_t lval, uint64_t* local_mask, uint64_t* local_val) const {
    uint64_t m = 0, v = 0;
    bool ok = true;
    for (int p = 0; p < n_; ++p) {
        if (!((lmask >> p) & 1)) continue;
        const uint64_t bit = (lval >> p) & 1;
        if (loc_[p] < 64) {
            m |= uint64_t(1) << loc_[p];
            v |= bit << loc_[p];
        } else if (uint64_t((rank_ >> (loc_[p] - 64)) & 1) != bit)
            ok = false;
    }
    *local_mask = m;
    *local_val = v;
    return ok;
}

double Engine::read_scalar(const double* d_ptr) {
    PQB_CHECK(cudaMemcpyAsync(h_pinned_, d_ptr, sizeof(double), cudaMemcpyDeviceToHost, stream_));
    PQB_CHECK(cudaStreamSynchronize(stream_));
    check_exchange_error();
    return h_pinned_[0];
}

void Engine::check_exchange_error() {
    if (dist_ && dist_->exchange_error())
        throw CudaErr("global<->local remap: a peer GPU did not answer within the time limit (exchange kernel wait code " +
                      std::to_string(dist_->exchange_error()) + "); the state is undefined");
}

// ---------------------------------------------------------------------------------------------------------------
// events and deferred timers
// ---------------------------------------------------------------------------------------------------------------
cudaEvent_t Engine::get_event() {
    if (used_events_.size() >= 16384) harvest_timers(true);
    cudaEvent_t e;
    if (!free_events_.empty()) {
        e = free_events_.back();
        free_events_.pop_back();
    } else
        PQB_CHECK(cudaEventCreate(&e));
    used_events_.push_back(e);
    return e;
}

void Engine::wait_on_main(cudaEvent_t ev) {
    cudaEvent_t s0 = get_event(), s1 = get_event();
    PQB_CHECK(cudaEventRecord(s0, stream_));
    PQB_CHECK(cudaStreamWaitEvent(stream_, ev, 0));
    PQB_CHECK(cudaEventRecord(s1, stream_));
    timers_.push_back(Timer{s0, s1, T_STALL});
}

void Engine::harvest_timers(bool synchronize) {
    if (used_events_.empty()) return;
    if (synchronize) {
        PQB_CHECK(cudaStreamSynchronize(stream_));
        if (dist_ && dist_->comm_stream()) PQB_CHECK(cudaStreamSynchronize(dist_->comm_stream()));
    }
    for (auto& t : timers_) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, t.e0, t.e1) != cudaSuccess) {
            cudaGetLastError();
            continue;
        }
        if (t.kind == T_STALL)
            stats_.remap_ms += ms;
        else if (t.kind == T_COMM)
            stats_.remap_comm_ms += ms;
        else if (t.kind == T_DIAG)
            stats_.diag_ms += ms;
        else if (t.kind >= 0 && t.kind < 6)
            stats_.pass_ms[t.kind] += ms;
    }
    timers_.clear();
    free_events_.insert(free_events_.end(), used_events_.begin(), used_events_.end());
    used_events_.clear();
}

void Engine::get_stats(pqb_stats* out) {
    harvest_timers(true);
    check_exchange_error();
    *out = stats_;
}

void Engine::reset_stats() {
    harvest_timers(true);
    stats_ = pqb_stats{};
}

double Engine::allreduce_sum(double v) { return dist_ ? dist_->allreduce_sum(v) : v; }

void Engine::ensure_scratch(GrowBuffer& b, size_t bytes) {
    b.init(device_, /*exportable=*/world_ > 1);  // no-op after the first time
    try {
        b.ensure(bytes, stream_);
    } catch (const std::bad_alloc&) {
        throw CudaErr("out of device memory for a scratch copy of the state (" + std::to_string(bytes >> 20) + " MiB)");
    }
}

void* Engine::small_upload(const void* src, size_t bytes) {
    if (bytes > d_small_cap_) {
        PQB_CHECK(cudaStreamSynchronize(stream_));
        if (d_small_) cudaFree(d_small_);
        d_small_cap_ = std::max<size_t>(bytes * 2, 1 << 16);
        PQB_CHECK(cudaMalloc(&d_small_, d_small_cap_));
    }
    // pageable source: the copy is staged by the runtime before the call returns
    PQB_CHECK(cudaMemcpyAsync(d_small_, src, bytes, cudaMemcpyHostToDevice, stream_));
    return d_small_;
}

double Engine::draw_uniform() {
    // std::uniform_real_distribution<double>(0,1) on std::mt19937 (reference: simulator.hpp:51-52,153) is
    // generate_canonical<double,53>: two 32-bit draws, low word first (libstdc++ bits/random.tcc).
    const double u0 = double(rng_());
    const double u1 = double(rng_());
    double r = (u0 + u1 * 4294967296.0) / 18446744073709551616.0;
    if (r >= 1.0) r = std::nextafter(1.0, 0.0);
    return r;
}

// ---------------------------------------------------------------------------------------------------------------
// allocation
// ---------------------------------------------------------------------------------------------------------------
void Engine::allocate_qubit(uint32_t id) {
    if (known(id)) throw RuntimeErr("AllocateQubit: ID already exists. Qubit IDs should be unique.");
    if (n_ >= 62) throw RuntimeErr("AllocateQubit: too many qubits");
    // pending gates keep their meaning: the new qubit is a new most-significant logical bit (simulator.hpp:57)
    if (dist_ && dist_->has_free_rank_bit() && L_ >= kMinLocalBits) {
        // a free rank bit is all-zero outside value 0, which is exactly a fresh |0> qubit: no data moves.  The first
        // kMinLocalBits qubits take local bits instead: a lazily allocating program (allocate q0; H | q0; flush) must be
        // able to run its first gates without a remap that has nothing to evict.
        const int r = dist_->take_free_rank_bit();
        loc_.push_back(uint8_t(64 + r));
    } else {
        const size_t old_bytes = sizeof(double2) << L_;
        try {
            state_->ensure(old_bytes * 2, stream_);
        } catch (const std::bad_alloc&) {
            // give the scratch copies back and retry once
            scratch1_->release();
            scratch2_->release();
            try {
                state_->ensure(old_bytes * 2, stream_);
            } catch (const std::bad_alloc&) {
                throw CudaErr("AllocateQubit: out of device memory at " + std::to_string(n_ + 1) + " qubits");
            }
        }
        PQB_CHECK(cudaMemsetAsync(reinterpret_cast<char*>(state_->ptr()) + old_bytes, 0, old_bytes, stream_));
        loc_.push_back(uint8_t(L_));
        ++L_;
    }
    map_[id] = uint32_t(n_);
    ++n_;
}

bool Engine::is_classical(uint32_t id, double tol) {
    run();
    const uint32_t lp = pos_of(id, "is_classical(): Unknown qubit id.");
    unsigned long long* d_out = reinterpret_cast<unsigned long long*>(d_scalars_);
    std::vector<uint8_t> phys2log;
    const bool ident = layout_is_identity() && !dist_;
    uint64_t rank_bits = 0;
    if (!ident) {
        phys2log.assign(64, 0);
        for (int p = 0; p < n_; ++p) phys2log[loc_[p] < 64 ? loc_[p] : L_ + (loc_[p] - 64)] = uint8_t(p);
        rank_bits = uint64_t(rank_) << L_;
    }
    // In a sharded run physical bit L_+r is rank bit r; free rank bits carry no logical qubit, and ranks with such a
    // bit set hold only zeros, so whatever phys2log says for them is never used.
    k::classical_probe(ctx(), psi(), local_amps(), loc_[lp] < 64 ? loc_[lp] : L_ + (loc_[lp] - 64), int(lp), tol,
                       ident ? nullptr : phys2log.data(), dist_ ? L_ + dist_->rank_bits() : n_, rank_bits, d_out);
    PQB_CHECK(cudaMemcpyAsync(h_pinned_, d_out, 16, cudaMemcpyDeviceToHost, stream_));
    PQB_CHECK(cudaStreamSynchronize(stream_));
    unsigned long long m[2];
    std::memcpy(m, h_pinned_, 16);
    if (dist_) {
        m[0] = dist_->allreduce_min_u64(m[0]);
        m[1] = dist_->allreduce_min_u64(m[1]);
    }
    last_probe_[0] = m[0];
    last_probe_[1] = m[1];
    const bool down = m[0] != ~0ULL, up = m[1] != ~0ULL;
    return down != up;
}

bool Engine::get_classical_value(uint32_t id, double tol) {
    // first amplitude above tol in the reference's scan order decides (simulator.hpp:81-88): the bit-0 member of a
    // pair is looked at before its bit-1 partner, so bit 1 wins only with a strictly smaller pair number
    is_classical(id, tol);
    return last_probe_[1] < last_probe_[0];
}

void Engine::deallocate_qubit(uint32_t id) {
    run();
    if (!known(id)) throw RuntimeErr("DeallocateQubit: Unknown qubit id.");
    if (!is_classical(id, 1e-12))
        throw RuntimeErr(
            "Error: Qubit has not been measured / uncomputed! There is most likely a bug in your code.");
    const bool value = last_probe_[1] < last_probe_[0];
    const uint32_t lp = map_[id];
    if (!is_local(lp)) {
        // the qubit sits on a rank bit: hand the bit back, moving the surviving half onto the ranks where it is 0
        dist_->release_rank_bit(loc_[lp] - 64, value, psi(), local_amps());
    } else {
        const int pb = loc_[lp];
        const uint64_t half = local_amps() >> 1;
        if (pb == L_ - 1) {
            // top local bit: the surviving half is contiguous
            if (value)
                PQB_CHECK(cudaMemcpyAsync(psi(), psi() + half, half * sizeof(double2), cudaMemcpyDeviceToDevice, stream_));
        } else {
            ensure_scratch(*scratch1_, half * sizeof(double2));
            k::compact_bit(ctx(), psi(), scratch1_->amps(), half, pb, value ? 1 : 0);
            std::swap(state_, scratch1_);
        }
        for (auto& l : loc_)
            if (l < 64 && l > pb) --l;
        --L_;
    }
    // logical bookkeeping (simulator.hpp:136-141)
    loc_.erase(loc_.begin() + lp);
    for (auto& kv : map_)
        if (kv.second > lp) --kv.second;
    map_.erase(id);
    --n_;
}

// ---------------------------------------------------------------------------------------------------------------
// gates
// ---------------------------------------------------------------------------------------------------------------
void Engine::apply_controlled_gate(const double* m, const uint32_t* ids, size_t k, const uint32_t* ctrl, size_t nc) {
    if (k > size_t(k::kMaxDense)) throw ValueErr("Gates with more than 5 qubits are not supported!");
    if (k == 0) throw ValueErr("apply_controlled_gate(): no target qubits");
    Gate g;
    g.targets.assign(ids, ids + k);
    g.ctrls.assign(ctrl, ctrl + nc);
    // a control listed twice is one control (the reference ORs the positions into a mask, simulator.hpp:551-556)
    std::sort(g.ctrls.begin(), g.ctrls.end());
    g.ctrls.erase(std::unique(g.ctrls.begin(), g.ctrls.end()), g.ctrls.end());
    for (size_t i = 0; i < k; ++i) {
        for (size_t j = i + 1; j < k; ++j)
            if (ids[i] == ids[j]) throw ValueErr("apply_controlled_gate(): duplicate target qubit");
        for (size_t j = 0; j < nc; ++j)
            if (ids[i] == ctrl[j]) throw ValueErr("apply_controlled_gate(): a qubit is both target and control");
    }
    const size_t d = size_t(1) << k;
    g.m.resize(d * d);
    for (size_t i = 0; i < d * d; ++i) g.m[i] = cplx(m[2 * i], m[2 * i + 1]);
    fuser_.push(std::move(g));
    ++stats_.gates_ingested;
    if (fuser_.pending() >= kMaxPending) run();
}

uint64_t Engine::Launch::touched() const {
    // the same on every rank, also where the pass is switched off: the ranks of an exchange group must agree on the slice
    // bits of a pipelined remap, which are chosen among the bits the neighbouring passes do not touch
    uint64_t m = 0;
    for (int l = 0; l < k; ++l) m |= uint64_t(1) << tpos[l];
    for (int l = 0; l < n_ctrl; ++l) m |= uint64_t(1) << cpos[l];
    return m;
}

Engine::Launch Engine::resolve_pass(const FusedPass& p) {
    Launch out;
    const int kq = int(p.targets.size());
    if (kq > k::kMaxDense) throw ValueErr("Gates with more than 5 qubits are not supported!");
    // controls on rank bits switch whole ranks on or off; controls on local bits become the kernel's control mask
    int ncl = 0;
    bool active = true;
    for (auto c : p.ctrls) {
        const uint32_t lp = map_.at(c);
        if (is_local(lp))
            out.cpos[ncl++] = loc_[lp];
        else if (!((rank_ >> (loc_[lp] - 64)) & 1))
            active = false;
    }
    if (dist_ && (uint64_t(rank_) & dist_->free_rank_bits_mask()) != 0) active = false;  // all-zero shard
    std::sort(out.cpos, out.cpos + ncl);
    out.n_ctrl = ncl;
    const size_t D = size_t(1) << kq;
    if (p.diagonal) {
        // a diagonal pass needs no remap: targets on rank bits just select a slice of the diagonal for this rank
        ++stats_.diag_passes;
        int which[8], kl = 0;
        size_t fixed = 0;
        for (int l = 0; l < kq; ++l) {
            const uint32_t lp = map_.at(p.targets[l]);
            if (is_local(lp)) {
                out.tpos[kl] = loc_[lp];
                which[kl++] = l;
            } else if ((rank_ >> (loc_[lp] - 64)) & 1)
                fixed |= size_t(1) << l;
        }
        out.m.resize(size_t(2) << kl);
        for (size_t v = 0; v < (size_t(1) << kl); ++v) {
            size_t idx = fixed;
            for (int j = 0; j < kl; ++j)
                if ((v >> j) & 1) idx |= size_t(1) << which[j];
            out.m[2 * v] = p.m[idx * D + idx].real();
            out.m[2 * v + 1] = p.m[idx * D + idx].imag();
        }
        out.k = kl;
        out.kind = active ? Launch::DIAG : Launch::NONE;
        return out;
    }
    ++stats_.dense_passes[kq];
    for (int l = 0; l < kq; ++l) {
        const uint32_t lp = map_.at(p.targets[l]);
        if (!is_local(lp)) throw RuntimeErr("internal: dense target on a rank bit was not remapped");
        out.tpos[l] = loc_[lp];
        if (l > 0 && out.tpos[l] <= out.tpos[l - 1]) throw RuntimeErr("internal: pass targets are not in ascending bit order");
    }
    out.k = kq;
    out.kind = active ? Launch::DENSE : Launch::NONE;
    if (active) {
        const double* m = reinterpret_cast<const double*>(p.m.data());
        out.m.assign(m, m + 2 * D * D);
    }
    return out;
}

void Engine::launch(const Launch& l, const k::Slice& slice) {
    if (l.kind == Launch::NONE) return;
    cudaEvent_t e0 = nullptr;
    if (profiling_) {
        e0 = get_event();
        PQB_CHECK(cudaEventRecord(e0, stream_));
    }
    if (l.kind == Launch::DIAG)
        k::apply_diagonal(ctx(), psi(), L_, l.k, l.tpos, l.n_ctrl, l.cpos, l.m.data(), slice);
    else
        k::apply_dense(ctx(), psi(), L_, l.k, l.tpos, l.n_ctrl, l.cpos, l.m.data(), slice);
    if (profiling_) {
        cudaEvent_t e1 = get_event();
        PQB_CHECK(cudaEventRecord(e1, stream_));
        timers_.push_back(Timer{e0, e1, l.kind == Launch::DIAG ? int(T_DIAG) : l.k});
    }
}

// bring the given logical positions onto local bits (global<->local qubit remap over NVLink)
void Engine::make_local(const std::vector<uint32_t>& need, const std::vector<uint32_t>* victims) {
    if (!dist_) return;
    bool any = false;
    for (auto lp : need)
        if (!is_local(lp)) any = true;
    if (!any) return;
    std::vector<std::pair<int, int>> swaps;
    try {
        swaps = plan_remap(loc_, L_, need, victims);
    } catch (const std::runtime_error& e) {
        throw RuntimeErr(e.what());
    }
    // A qubit that leaves from a low local bit leaves in runs of a few bytes.  The NCCL path then has to gather/scatter
    // through staging slots (measured 240-340 GB/s per direction; bit 0: every 32-byte sector is half used) against
    // 515-570 GB/s for the contiguous halves of a high bit; the peer-memory kernel only needs runs of a full 128-byte line
    // (bit 3 and up).  So first move the leaving qubit onto the highest free local bit with one local pass (16 B/amplitude
    // at HBM speed, ~25 ms on a 128 GiB shard), then exchange that bit.  The incoming qubit lands on that high bit and is
    // usually the next to leave, so this is paid once, not per remap.
    // plan_remap has already updated loc_ as if the exchange happened at the low bit; the bookkeeping below redirects it.
    if (L_ >= 2) {
        std::vector<int> taken;  // local bits that take part in this remap
        for (auto& sw : swaps) taken.push_back(sw.second);
        int t = L_ - 1;
        for (auto& sw : swaps) {
            const int b = sw.second;
            if (!leaves_from_low_bit(b, swaps.size())) continue;
            while (t >= 0 && std::find(taken.begin(), taken.end(), t) != taken.end()) --t;
            if (t <= b) break;
            int incoming = -1, at_t = -1;  // logical positions: the qubit plan_remap placed on b, the qubit sitting on t
            for (int p = 0; p < n_; ++p) {
                if (loc_[p] == b) incoming = p;
                if (loc_[p] == t) at_t = p;
            }
            k::swap_local_bits(ctx(), psi(), L_, b, t);
            if (incoming >= 0) loc_[incoming] = uint8_t(t);
            if (at_t >= 0) loc_[at_t] = uint8_t(b);
            sw.second = t;
            taken.push_back(t);
        }
    }
    serial_exchange(swaps);
}

// one remap, not overlapped with anything: the peer-memory exchange kernel (all bits at once) when every member of the
// exchange group can map the others' shards, NCCL send/recv through a staging slice otherwise
void Engine::serial_exchange(const std::vector<std::pair<int, int>>& swaps) {
    if (swaps.empty()) return;
    ++stats_.remaps;
    stats_.remap_qubits += swaps.size();
    try {
        if (dist_->prepare_exchange(swaps, *state_, device_)) {
            cudaEvent_t before = get_event(), x0 = get_event(), x1 = get_event();
            PQB_CHECK(cudaEventRecord(before, stream_));
            PQB_CHECK(cudaStreamWaitEvent(dist_->comm_stream(), before, 0));
            PQB_CHECK(cudaEventRecord(x0, dist_->comm_stream()));
            dist_->exchange_slice(k::Slice(), L_, sm_count_, &stats_.remap_bytes_sent);
            ++stats_.kernel_launches;
            PQB_CHECK(cudaEventRecord(x1, dist_->comm_stream()));
            timers_.push_back(Timer{x0, x1, T_COMM});
            wait_on_main(x1);
            ++stats_.p2p_remaps;
            return;
        }
        // NCCL path.  Staging: a bounded slice of the second scratch buffer (the state itself may fill most of HBM)
        const uint64_t want = std::min<uint64_t>(local_amps() >> 1, uint64_t(1) << 26);  // <= 1 GiB
        ensure_scratch(*scratch2_, std::max<uint64_t>(want, 1) * sizeof(double2));
        cudaEvent_t e0 = get_event(), e1 = get_event();
        PQB_CHECK(cudaEventRecord(e0, stream_));
        // Several bits at once can also go as pairwise rounds over sub-blocks (PQB_REMAP_MULTI=1); measured slower than
        // bit-by-bit exchanges with NCCL on 4 GPUs, so the NCCL path exchanges bit by bit.
        static const bool multi = [] {
            const char* e = getenv("PQB_REMAP_MULTI");
            return e && e[0] == '1';
        }();
        if (multi && swaps.size() >= 2 && swaps.size() <= 8 && want >= 4)
            dist_->swap_bits_multi(swaps, psi(), L_, scratch2_->amps(), want, &stats_.remap_bytes_sent);
        else
            for (auto& sw : swaps)
                dist_->swap_bits(sw.first, sw.second, psi(), L_, scratch2_->amps(), std::max<uint64_t>(want, 1),
                                 &stats_.remap_bytes_sent);
        PQB_CHECK(cudaEventRecord(e1, stream_));
        timers_.push_back(Timer{e0, e1, T_STALL});
    } catch (const CudaErr&) {
        throw;
    } catch (const std::runtime_error& e) {
        throw CudaErr(e.what());
    }
}

// One remap executed slice by slice.  The slice bits are local bits that neither the exchange nor the passes around it
// touch, so slice i can receive the last passes before the remap (`tail`), go over NVLink, and receive the first passes
// after it (`head`) independently of the other slices.  The main stream runs tail(0), tail(1), head(0), tail(2), head(1),
// ... while the communication stream exchanges slice i as soon as tail(i) is done: the exchange of a slice overlaps the
// passes on its neighbours, and only the first and last exchange can leave the main stream waiting.
void Engine::pipelined_exchange(const std::vector<Launch>& tail, const std::vector<Launch>& head,
                                const std::vector<uint8_t>& slice_bits) {
    const int s = int(slice_bits.size());
    const int n_slices = 1 << s;
    cudaStream_t comm = dist_->comm_stream();
    auto slice_of = [&](int i) {
        k::Slice sl;
        sl.n = s;
        for (int b = 0; b < s; ++b) sl.pos[b] = slice_bits[b];
        sl.val = deposit_bits(uint64_t(i), slice_bits.data(), s);
        return sl;
    };
    std::vector<cudaEvent_t> exchanged(n_slices);
    for (int i = 0; i < n_slices; ++i) {
        const k::Slice sl = slice_of(i);
        for (auto& l : tail) launch(l, sl);
        cudaEvent_t ready = get_event(), x0 = get_event();
        exchanged[i] = get_event();
        PQB_CHECK(cudaEventRecord(ready, stream_));
        PQB_CHECK(cudaStreamWaitEvent(comm, ready, 0));
        PQB_CHECK(cudaEventRecord(x0, comm));
        dist_->exchange_slice(sl, L_, sm_count_, &stats_.remap_bytes_sent);
        ++stats_.kernel_launches;
        PQB_CHECK(cudaEventRecord(exchanged[i], comm));
        timers_.push_back(Timer{x0, exchanged[i], T_COMM});
        if (i >= 1) {
            wait_on_main(exchanged[i - 1]);
            const k::Slice prev = slice_of(i - 1);
            for (auto& l : head) launch(l, prev);
        }
    }
    wait_on_main(exchanged[n_slices - 1]);
    const k::Slice last = slice_of(n_slices - 1);
    for (auto& l : head) launch(l, last);
}

void Engine::run() {
    if (fuser_.pending() == 0) return;
    if (dist_) {
        run_sharded();
        return;
    }
    // sort key of a qubit inside a pass = its physical place; rank bits sort above all local bits
    auto key = [this](uint32_t id) -> uint64_t {
        auto it = map_.find(id);
        if (it == map_.end()) throw RuntimeErr("apply_controlled_gate(): Unknown qubit id. Please allocate the qubit first.");
        return loc_[it->second];
    };
    // Relative cost of one pass by width, measured on B200 (profiles/): k <= 4 runs at the HBM roofline (32 B/amplitude);
    // k = 5 is bound by FP64 throughput (256 flop/amplitude): ~1.55x on the tensor pipe (DMMA kernel; sustained 9.2 ms against
    // 5.85 ms at 30 qubits), ~2.4x with the DFMA kernel, so a 5-wide pass only pays off when it swallows that many more gates.
    // A control bit halves the amplitudes a pass touches.
    auto cost = [&](const std::vector<Cluster>& cs) {
        double c = 0.0;
        for (auto& cl : cs) {
            double w = 1.0;
            if (cl.width >= 5) {
                uint64_t lowest = 64;
                for (auto t : cl.targets) lowest = std::min(lowest, key(t));
                for (auto q : cl.ctrls) lowest = std::min(lowest, key(q));
                w = k::dense_k5_dmma_applies(L_, int(lowest), cl.width + cl.n_ctrl) ? 1.55 : 2.4;
            }
            for (int i = 0; i < cl.n_ctrl && i < 6; ++i) w *= 0.5;
            c += w + 0.002;  // + launch overhead so tiny states prefer fewer passes
        }
        return c;
    };
    try {
        // every id must be known before anything is applied (the reference would silently insert into map_)
        for (size_t gi = 0; gi < fuser_.pending(); ++gi) {
            const Gate& gt = fuser_.pending_gate(gi);
            for (auto t : gt.targets) key(t);
            for (auto c : gt.ctrls) key(c);
        }
        // schedule first (no matrix products), then fuse and launch pass by pass: the GPU runs pass i while the host
        // builds the matrix of pass i+1
        std::vector<Cluster> clusters;
        if (fusion_max_ > 0) {
            clusters = fuser_.schedule(fusion_max_);
        } else {
            clusters = fuser_.schedule(4);
            if (clusters.size() > 1) {
                std::vector<Cluster> wide = fuser_.schedule(5);
                if (cost(wide) < cost(clusters)) clusters.swap(wide);
            }
        }
        for (auto& cl : clusters) apply_pass(fuser_.fuse_cluster(cl, key));
        fuser_.clear();
    } catch (...) {
        fuser_.clear();  // never leave a poisoned queue behind (the reference does, simulator.hpp:522-526)
        throw;
    }
}

bool Engine::leaves_from_low_bit(int b, size_t n_swaps) const {
    return dist_->p2p_enabled() ? b < 3 : b < L_ - int(n_swaps);
}

void Engine::resolve_phase(ShardPlan& plan, std::vector<Launch>& out, bool eager, size_t hold) {
    auto key = [this](uint32_t id) -> uint64_t { return loc_[map_.at(id)]; };
    const std::vector<size_t> runnable = plan.take_runnable(map_, loc_);
    for (size_t c = 0; c < runnable.size(); ++c) {
        Launch l = resolve_pass(fuser_.fuse_cluster(plan.cluster(runnable[c]), key));
        if (eager && out.empty() && c + hold < runnable.size())
            launch(l);  // the GPU works on this pass while the host multiplies the matrices of the next one
        else
            out.push_back(std::move(l));
    }
}

// Sharded run().  The flush is scheduled into passes once, exactly as on one GPU (ShardPlan), and every pass whose targets
// are on-device runs before a remap is paid for.  The remap brings in the rank-bit qubits of the oldest waiting pass and
// evicts the local qubits that are needed last (Belady), so a brickwork circuit needs one remap per flush instead of one
// per layer.  The remap itself moves all its qubits in one peer-memory exchange and is pipelined slice by slice against the
// last passes before it and the first passes after it (pipelined_exchange), so that NVLink traffic and HBM-bound passes
// run at the same time.
void Engine::run_sharded() {
    static const int max_slice_bits = [] {
        const char* e = getenv("PQB_REMAP_SLICE_BITS");  // 0 switches the pipeline off
        return e ? std::max(0, std::min(4, atoi(e))) : 3;
    }();
    constexpr size_t kTail = 3, kHead = 4;  // passes before / after the remap that may run slice by slice
    const int width = fusion_max_ > 0 ? fusion_max_ : 4;
    try {
        // every id must be known before anything is applied (the reference would silently insert into map_)
        for (size_t gi = 0; gi < fuser_.pending(); ++gi) {
            const Gate& gt = fuser_.pending_gate(gi);
            for (auto t : gt.targets) pos_of(t, "apply_controlled_gate(): Unknown qubit id. Please allocate the qubit first.");
            for (auto c : gt.ctrls) pos_of(c, "apply_controlled_gate(): Unknown qubit id. Please allocate the qubit first.");
        }
        const InteractionGraph adj = interaction_graph(fuser_);  // of this flush: breaks ties between eviction candidates
        ShardPlan plan(fuser_, width);
        std::vector<Launch> held;  // resolved but not yet launched (the candidates for a pipeline's tail)
        resolve_phase(plan, held, true, kTail);
        while (!plan.finished()) {
            const RemapChoice choice = plan.choose(map_, loc_, adj);
            // plan on a copy of the layout first: a remap that needs a local pre-pass is not pipelined
            std::vector<uint8_t> planned = loc_;
            std::vector<std::pair<int, int>> swaps;
            try {
                swaps = plan_remap(planned, L_, choice.need, &choice.victims);
            } catch (const std::runtime_error& e) {
                throw RuntimeErr(e.what());
            }
            bool pipelined = dist_->p2p_enabled() && max_slice_bits > 0 && !swaps.empty() && swaps.size() <= 3 &&
                             L_ - int(swaps.size()) >= 2;
            uint64_t exchanged_bits = 0;
            for (auto& sw : swaps) {
                if (leaves_from_low_bit(sw.second, swaps.size())) pipelined = false;
                exchanged_bits |= uint64_t(1) << sw.second;
            }
            if (pipelined) {
                try {
                    pipelined = dist_->prepare_exchange(swaps, *state_, device_);  // collective: same decision on every rank
                } catch (const std::runtime_error& e) {
                    throw CudaErr(e.what());
                }
            }
            if (!pipelined) {
                for (auto& l : held) launch(l);
                held.clear();
                make_local(choice.need, &choice.victims);
                resolve_phase(plan, held, true, kTail);
                continue;
            }
            loc_ = planned;
            ++stats_.remaps;
            ++stats_.p2p_remaps;
            ++stats_.pipelined_remaps;
            stats_.remap_qubits += swaps.size();
            const uint64_t all_local = (uint64_t(1) << L_) - 1;
            const int want = std::min(max_slice_bits, L_ - int(swaps.size()) - 1);
            auto free_bits = [&](uint64_t used) { return __builtin_popcountll(all_local & ~used); };
            // tail: the longest suffix of the held passes that leaves enough untouched bits to slice along
            size_t n_tail = 0;
            uint64_t used = exchanged_bits;
            while (n_tail < held.size() && n_tail < kTail) {
                const uint64_t u = used | held[held.size() - 1 - n_tail].touched();
                if (free_bits(u) < want) break;
                used = u;
                ++n_tail;
            }
            for (size_t i = 0; i + n_tail < held.size(); ++i) launch(held[i]);
            std::vector<Launch> tail(held.end() - n_tail, held.end());
            held.clear();
            // what can run once the exchanged qubits are on-device; resolved while the GPU works through the launches above
            std::vector<Launch> next;
            resolve_phase(plan, next, false, 0);
            size_t n_head = 0;
            while (n_head < next.size() && n_head < kHead) {
                const uint64_t u = used | next[n_head].touched();
                if (free_bits(u) < want) break;
                used = u;
                ++n_head;
            }
            std::vector<uint8_t> slice_bits;
            if (n_tail + n_head > 0)
                for (int b = L_ - 1; b >= 0 && int(slice_bits.size()) < want; --b)
                    if (!((used >> b) & 1)) slice_bits.push_back(uint8_t(b));
            std::sort(slice_bits.begin(), slice_bits.end());
            std::vector<Launch> head(next.begin(), next.begin() + n_head);
            pipelined_exchange(tail, head, slice_bits);
            // the rest of the phase: everything but the last kTail launches goes out now
            const size_t n_rest = next.size() - n_head;
            for (size_t i = 0; i < n_rest; ++i) {
                if (i + kTail < n_rest)
                    launch(next[n_head + i]);
                else
                    held.push_back(std::move(next[n_head + i]));
            }
        }
        for (auto& l : held) launch(l);
        fuser_.clear();
    } catch (...) {
        fuser_.clear();
        throw;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// measurement
// ---------------------------------------------------------------------------------------------------------------
void Engine::measure_qubits(const uint32_t* ids, size_t n, uint8_t* out) {
    run();
    std::vector<uint32_t> lpos(n);
    for (size_t i = 0; i < n; ++i) lpos[i] = pos_of(ids[i], "measure_qubits(): Unknown qubit id.");
    const double rnd = draw_uniform();

    // Inverse-CDF search in *logical* index order (simulator.hpp:156-160), kBinBits logical bits per level from the top:
    // each level sums |psi|^2 over the 2^m bins of the next m logical bits inside the prefix chosen so far, the host
    // walks the bins sequentially exactly like the reference walks amplitudes, and the search descends into the bin.
    uint64_t pick = 0;       // logical index bits decided so far (in place)
    uint64_t decided = 0;    // mask of decided logical bits
    double remaining = rnd;  // rnd minus the mass of everything before the chosen prefix
    bool overflowed = false; // the running sum never reached rnd: the reference ends on the last index
    int top = n_;
    if (n_ == 0) top = 0;
    // The first level sweeps the whole state anyway, so it also bins over the measured positions that lie below its own
    // bits (when they fit): the norm of the surviving amplitudes is then a sum of first-level bins and the separate norm
    // sweep before the collapse is not needed.
    std::vector<double> level1_bins;
    std::vector<int> level1_bits;  // logical position of every bit of a first-level bin index
    while (top > 0) {
        const int m = std::min(kBinBits, top);
        const int lo = top - m;  // this level decides logical bits [lo, top)
        std::vector<int> bin_lp;  // logical positions of the bin-index bits, least significant first
        if (top == n_) {
            std::vector<int> extra;
            for (auto lp : lpos)
                if (int(lp) < lo && std::find(extra.begin(), extra.end(), int(lp)) == extra.end()) extra.push_back(int(lp));
            std::sort(extra.begin(), extra.end());
            if (m + int(extra.size()) <= 12) bin_lp = extra;
        }
        const int n_extra = int(bin_lp.size());
        for (int lp = lo; lp < top; ++lp) bin_lp.push_back(lp);
        const int n_bin_bits = int(bin_lp.size());
        const int n_bins = 1 << m;
        // local view: decided bits and bin bits that live on local physical bits
        std::vector<uint8_t> ins;
        uint8_t bin_pos[16];
        int m_local = 0;
        uint64_t fixed_val = 0;
        bool rank_matches_prefix = true;
        int bin_is_local[16];
        for (int b = 0; b < n_bin_bits; ++b) {
            const int lp = bin_lp[b];
            bin_is_local[b] = is_local(lp);
            if (bin_is_local[b]) {
                bin_pos[m_local++] = loc_[lp];
                ins.push_back(loc_[lp]);
            }
        }
        for (int lp = top; lp < n_; ++lp) {
            const uint64_t bit = (pick >> lp) & 1;
            if (is_local(lp)) {
                ins.push_back(loc_[lp]);
                fixed_val |= bit << loc_[lp];
            } else if (uint64_t((rank_ >> (loc_[lp] - 64)) & 1) != bit)
                rank_matches_prefix = false;
        }
        if (dist_ && (uint64_t(rank_) & dist_->free_rank_bits_mask()) != 0) rank_matches_prefix = false;
        std::sort(ins.begin(), ins.end());
        std::vector<double> fine(size_t(1) << n_bin_bits, 0.0);
        if (rank_matches_prefix) {
            k::bin_sums(ctx(), psi(), L_, int(ins.size()), ins.data(), fixed_val, m_local, bin_pos, d_partials_, d_scalars_);
            PQB_CHECK(cudaMemcpyAsync(h_pinned_, d_scalars_, sizeof(double) << m_local, cudaMemcpyDeviceToHost, stream_));
            PQB_CHECK(cudaStreamSynchronize(stream_));
            // scatter the local bins into the full bin array (rank bits of this rank fill the non-local bin bits)
            for (int lb = 0; lb < (1 << m_local); ++lb) {
                int full = 0, j = 0;
                for (int b = 0; b < n_bin_bits; ++b) {
                    int bit;
                    if (bin_is_local[b])
                        bit = (lb >> j++) & 1;
                    else
                        bit = (rank_ >> (loc_[bin_lp[b]] - 64)) & 1;
                    full |= bit << b;
                }
                fine[full] = h_pinned_[lb];
            }
        }
        if (dist_) dist_->allreduce_sum_vec(fine.data(), fine.size());
        // the walk is over the bits this level decides: sum the extra (measured, lower) bits out
        std::vector<double> bins(n_bins, 0.0);
        for (size_t f = 0; f < fine.size(); ++f) bins[f >> n_extra] += fine[f];
        if (top == n_) {
            level1_bins.swap(fine);
            level1_bits = bin_lp;
        }
        // sequential walk (while (P < rnd && pick < size) P += ...; pick--)
        int chosen = -1;
        double before = 0.0;
        if (!overflowed) {
            double P = 0.0;
            for (int b = 0; b < n_bins; ++b) {
                if (P + bins[b] >= remaining && (bins[b] > 0.0 || remaining <= P)) {
                    chosen = b;
                    before = P;
                    break;
                }
                P += bins[b];
            }
        }
        if (chosen < 0) {
            // not reached: at the top level this is the reference's "last index" rule; below it, it can only be a
            // rounding difference between two summation orders -> stay on the last amplitude that carries weight
            chosen = n_bins - 1;
            if (top != n_ || overflowed) {
                for (int b = n_bins - 1; b >= 0; --b)
                    if (bins[b] > 0.0) {
                        chosen = b;
                        break;
                    }
            }
            overflowed = true;
        }
        remaining -= before;
        pick |= uint64_t(chosen) << lo;
        decided |= ((uint64_t(1) << m) - 1) << lo;
        top = lo;
        // only the measured bits of the picked index are needed: once every measured position is decided the lower levels
        // (one kernel and one host round trip each) cannot change the outcome
        bool all_decided = true;
        for (auto lp : lpos)
            if (int(lp) < lo) all_decided = false;
        if (all_decided) break;
    }
    (void)decided;

    uint64_t mask = 0, val = 0;
    for (size_t i = 0; i < n; ++i) {
        const uint64_t bit = (pick >> lpos[i]) & 1;
        out[i] = uint8_t(bit);
        mask |= uint64_t(1) << lpos[i];
        val |= bit << lpos[i];
    }
    // zero the amplitudes that disagree, renormalise the rest (simulator.hpp:172-185)
    uint64_t lmask, lval;
    const bool mine = split_mask(mask, val, &lmask, &lval);
    double kept = 0.0;
    bool from_bins = !level1_bins.empty();
    for (size_t i = 0; i < n && from_bins; ++i)
        if (std::find(level1_bits.begin(), level1_bits.end(), int(lpos[i])) == level1_bits.end()) from_bins = false;
    if (from_bins) {
        // every measured position is a bit of the first-level bin index: the surviving norm is the sum of the matching bins
        uint64_t bmask = 0, bval = 0;
        for (size_t b = 0; b < level1_bits.size(); ++b)
            if ((mask >> level1_bits[b]) & 1) {
                bmask |= uint64_t(1) << b;
                bval |= ((val >> level1_bits[b]) & 1) << b;
            }
        for (size_t f = 0; f < level1_bins.size(); ++f)
            if ((f & bmask) == bval) kept += level1_bins[f];
    } else {
        if (mine) {
            k::norm_masked(ctx(), psi(), local_amps(), lmask, lval, d_partials_, d_scalars_);
            kept = read_scalar(d_scalars_);
        }
        kept = allreduce_sum(kept);
    }
    const double scale = 1.0 / std::sqrt(kept);
    if (mine)
        k::collapse_scale(ctx(), psi(), local_amps(), lmask, lval, scale);
    else
        PQB_CHECK(cudaMemsetAsync(psi(), 0, local_amps() * sizeof(double2), stream_));
}

void Engine::collapse_wavefunction(const uint32_t* ids, size_t n_ids, const uint8_t* values, size_t n_values) {
    run();
    if (n_ids != n_values) throw ValueErr("collapse_wavefunction(): ids and values size mismatch");
    uint64_t mask = 0, val = 0;
    for (size_t i = 0; i < n_ids; ++i) {
        const uint32_t lp = pos_of(ids[i],
                                   "collapse_wavefunction(): Unknown qubit id(s) provided. Try calling eng.flush() before "
                                   "invoking this function.");
        mask |= uint64_t(1) << lp;
        val |= uint64_t(values[i] ? 1 : 0) << lp;
    }
    uint64_t lmask, lval;
    const bool mine = split_mask(mask, val, &lmask, &lval);
    double prob = 0.0;
    if (mine) {
        k::norm_masked(ctx(), psi(), local_amps(), lmask, lval, d_partials_, d_scalars_);
        prob = read_scalar(d_scalars_);
    }
    prob = allreduce_sum(prob);
    if (prob < 1e-12) throw RuntimeErr("collapse_wavefunction(): Invalid collapse! Probability is ~0.");
    const double scale = 1.0 / std::sqrt(prob);
    if (mine)
        k::collapse_scale(ctx(), psi(), local_amps(), lmask, lval, scale);
    else
        PQB_CHECK(cudaMemsetAsync(psi(), 0, local_amps() * sizeof(double2), stream_));
}

// ---------------------------------------------------------------------------------------------------------------
// queries
// ---------------------------------------------------------------------------------------------------------------
double Engine::get_probability(const uint8_t* bits, const uint32_t* ids, size_t n) {
    run();
    uint64_t mask = 0, val = 0;
    for (size_t i = 0; i < n; ++i) {
        const uint32_t lp =
            pos_of(ids[i], "get_probability(): Unknown qubit id. Please make sure you have called eng.flush().");
        mask |= uint64_t(1) << lp;
        val |= uint64_t(bits[i] ? 1 : 0) << lp;
    }
    uint64_t lmask, lval;
    double p = 0.0;
    if (split_mask(mask, val, &lmask, &lval)) {
        k::norm_masked(ctx(), psi(), local_amps(), lmask, lval, d_partials_, d_scalars_);
        p = read_scalar(d_scalars_);
    }
    return allreduce_sum(p);
}

double Engine::norm_squared() {
    run();
    k::norm_masked(ctx(), psi(), local_amps(), 0, 0, d_partials_, d_scalars_);
    return allreduce_sum(read_scalar(d_scalars_));
}

std::complex<double> Engine::get_amplitude(const uint8_t* bits, const uint32_t* ids, size_t n) {
    run();
    uint64_t chk = 0, index = 0;
    for (size_t i = 0; i < n; ++i) {
        auto it = map_.find(ids[i]);
        if (it == map_.end()) break;
        chk |= uint64_t(1) << it->second;
        index |= uint64_t(bits[i] ? 1 : 0) << it->second;
    }
    if (chk + 1 != (uint64_t(1) << n_))
        throw RuntimeErr(
            "The second argument to get_amplitude() must be a permutation of all allocated qubits. Please make sure you "
            "have called eng.flush().");
    double out[2];
    get_amplitudes(&index, 1, out);
    return {out[0], out[1]};
}

void Engine::get_amplitudes(const uint64_t* logical_idx, size_t n, double* out) {
    run();
    if (n == 0) return;
    std::vector<uint64_t> local(n);
    std::vector<char> mine(n);
    for (size_t i = 0; i < n; ++i) {
        if (n_ < 64 && (logical_idx[i] >> n_) != 0) throw ValueErr("get_amplitudes(): index out of range");
        bool ok;
        local[i] = logical_to_local_index(logical_idx[i], &ok);
        mine[i] = ok;
    }
    // device staging: room for n amplitudes (16-byte aligned) followed by the n indices
    std::vector<uint64_t> staged(3 * n, 0);
    std::copy(local.begin(), local.end(), staged.begin() + 2 * n);
    double2* d_out = static_cast<double2*>(small_upload(staged.data(), staged.size() * sizeof(uint64_t)));
    const uint64_t* d_idx = reinterpret_cast<const uint64_t*>(d_out + n);
    k::gather_indices(ctx(), psi(), d_idx, n, d_out);
    PQB_CHECK(cudaMemcpyAsync(out, d_out, n * sizeof(double2), cudaMemcpyDeviceToHost, stream_));
    PQB_CHECK(cudaStreamSynchronize(stream_));
    if (dist_) {
        for (size_t i = 0; i < n; ++i)
            if (!mine[i]) out[2 * i] = out[2 * i + 1] = 0.0;
        dist_->allreduce_sum_vec(out, 2 * n);
    }
}

void Engine::set_wavefunction(const double* wf, size_t n_amps, const uint32_t* ordering, size_t n) {
    run();
    bool ok = map_.size() == n;
    for (size_t i = 0; ok && i < n; ++i) ok = known(ordering[i]);
    if (!ok)
        throw RuntimeErr(
            "set_wavefunction(): Invalid mapping provided. Please make sure all qubits have been allocated previously "
            "(call eng.flush()).");
    if (n_amps != (size_t(1) << n)) throw ValueErr("set_wavefunction(): the wavefunction must have 2^n amplitudes");
    for (size_t i = 0; i < n; ++i) map_[ordering[i]] = uint32_t(i);
    if (map_.size() != n) throw RuntimeErr("set_wavefunction(): Invalid mapping provided (duplicate qubit ids).");
    if (!dist_) {
        for (int p = 0; p < n_; ++p) loc_[p] = uint8_t(p);
        PQB_CHECK(cudaMemcpyAsync(psi(), wf, n_amps * sizeof(double2), cudaMemcpyHostToDevice, stream_));
        PQB_CHECK(cudaStreamSynchronize(stream_));
        return;
    }
    // sharded: logical bits [0, L_) local, the rest on rank bits in order; every rank copies its slice
    dist_->reset_rank_bits(n_ - L_);
    for (int p = 0; p < n_; ++p) loc_[p] = uint8_t(p < L_ ? p : 64 + (p - L_));
    const uint64_t active_ranks = uint64_t(1) << (n_ - L_);
    if (uint64_t(rank_) < active_ranks)
        PQB_CHECK(cudaMemcpyAsync(psi(), wf + 2 * (uint64_t(rank_) << L_), local_amps() * sizeof(double2),
                                  cudaMemcpyHostToDevice, stream_));
    else
        PQB_CHECK(cudaMemsetAsync(psi(), 0, local_amps() * sizeof(double2), stream_));
    PQB_CHECK(cudaStreamSynchronize(stream_));
}

size_t Engine::cheat_map(uint32_t* ids, uint32_t* pos, size_t cap) {
    run();
    size_t i = 0;
    for (auto& kv : map_) {
        if (i < cap) {
            ids[i] = kv.first;
            pos[i] = kv.second;
        }
        ++i;
    }
    return i;
}

void Engine::cheat_state(double* out, size_t cap_amps) {
    run();
    const uint64_t total = uint64_t(1) << n_;
    if (cap_amps < total) throw ValueErr("cheat(): output buffer too small");
    if (!dist_) {
        const double2* src = psi();
        if (!layout_is_identity()) {
            ensure_scratch(*scratch1_, local_amps() * sizeof(double2));
            uint8_t perm[64];
            for (int p = 0; p < n_; ++p) perm[p] = loc_[p];  // logical out bit p <- physical in bit loc_[p]
            k::permute_gather(ctx(), psi(), scratch1_->amps(), local_amps(), n_, perm);
            src = scratch1_->amps();
        }
        PQB_CHECK(cudaMemcpyAsync(out, src, total * sizeof(double2), cudaMemcpyDeviceToHost, stream_));
        PQB_CHECK(cudaStreamSynchronize(stream_));
        return;
    }
    // sharded: every rank contributes its amplitudes at their logical indices; host-side sum (small states only)
    if (n_ > 26) throw RuntimeErr("cheat(): the sharded state is too large to gather on every rank (use get_amplitudes)");
    std::vector<double> mine(local_amps() * 2);
    PQB_CHECK(cudaMemcpyAsync(mine.data(), psi(), local_amps() * sizeof(double2), cudaMemcpyDeviceToHost, stream_));
    PQB_CHECK(cudaStreamSynchronize(stream_));
    std::memset(out, 0, total * sizeof(double2));
    if ((uint64_t(rank_) & dist_->free_rank_bits_mask()) == 0) {
        for (uint64_t i = 0; i < local_amps(); ++i) {
            uint64_t logical = 0;
            for (int p = 0; p < n_; ++p) {
                const uint64_t bit = loc_[p] < 64 ? (i >> loc_[p]) & 1 : uint64_t((rank_ >> (loc_[p] - 64)) & 1);
                logical |= bit << p;
            }
            out[2 * logical] = mine[2 * i];
            out[2 * logical + 1] = mine[2 * i + 1];
        }
    }
    dist_->allreduce_sum_vec(out, 2 * total);
}

// ---------------------------------------------------------------------------------------------------------------
// emulate_math
// ---------------------------------------------------------------------------------------------------------------
void Engine::emulate_math(int mode, int64_t a, int64_t N, const uint64_t* table, size_t table_len, const uint32_t* reg_ids,
                          const uint32_t* reg_sizes, size_t n_regs, const uint32_t* ctrl, size_t nc) {
    run();
    if (n_regs > 16) throw ValueErr("emulate_math(): more than 16 registers");
    if ((mode == k::MATH_ADD_MOD || mode == k::MATH_MUL_MOD) && N == 0) throw ValueErr("emulate_math(): N must not be 0");
    k::MathDesc d{};
    d.mode = mode;
    d.a = a;
    d.N = N;
    d.n_regs = int(n_regs);
    std::vector<uint32_t> need;
    size_t flat = 0;
    for (size_t r = 0; r < n_regs; ++r) {
        d.reg_off[r] = int(flat);
        for (uint32_t b = 0; b < reg_sizes[r]; ++b, ++flat) {
            if (flat >= 64) throw ValueErr("emulate_math(): registers too large");
            need.push_back(pos_of(reg_ids[flat], "emulate_math(): Unknown qubit id."));
        }
    }
    d.reg_off[n_regs] = int(flat);
    if (mode == k::MATH_TABLE && table_len != (size_t(1) << flat)) throw ValueErr("emulate_math(): table size mismatch");
    std::vector<uint32_t> cl;
    for (size_t i = 0; i < nc; ++i) cl.push_back(pos_of(ctrl[i], "emulate_math(): Unknown control qubit id."));
    if (dist_) make_local(need);
    for (size_t i = 0; i < flat; ++i) d.reg_pos[i] = loc_[need[i]];
    bool active = true;
    for (auto lp : cl) {
        if (is_local(lp))
            d.ctrl_mask |= uint64_t(1) << loc_[lp];
        else if (!((rank_ >> (loc_[lp] - 64)) & 1))
            active = false;
    }
    if (!active) return;  // this rank's control bits are not all set: identity on the whole shard
    const size_t bytes = local_amps() * sizeof(double2);
    if (flat <= size_t(k::kMathGatherBits)) {
        // Gather through the inverse map (kernels.cuh): tabulate v -> f(v) on the concatenated register value with the
        // reference's arithmetic (per register: x + a, (x + a) % N, (x * a) % N in C semantics, low bits of the result,
        // simulator.hpp:255-290 — in 64 bits), then list for every destination value its sources in
        // ascending order.  A bijection gives one source per destination and the kernel is a pure permutation (bit-exact);
        // inputs outside a gate's domain (x >= N) collide and are summed, as in the reference.
        const size_t V = size_t(1) << flat;
        std::vector<uint32_t> fwd(V);
        for (size_t v = 0; v < V; ++v) {
            if (mode == k::MATH_TABLE) {
                fwd[v] = uint32_t(table[v] & (V - 1));
                continue;
            }
            uint64_t y_all = 0;
            for (size_t r = 0; r < n_regs; ++r) {
                const int off = d.reg_off[r], nb = d.reg_off[r + 1] - d.reg_off[r];
                const long long x = (long long)((v >> off) & ((uint64_t(1) << nb) - 1));
                long long y;
                if (mode == k::MATH_ADD)
                    y = x + a;
                else if (mode == k::MATH_ADD_MOD)
                    y = (x + a) % N;
                else
                    y = (x * a) % N;
                y_all |= (uint64_t(y) & ((uint64_t(1) << nb) - 1)) << off;
            }
            fwd[v] = uint32_t(y_all);
        }
        std::vector<uint32_t> csr(2 * V + 1, 0);  // offsets (V + 1), then sources (V)
        uint32_t* off = csr.data();
        uint32_t* src = csr.data() + V + 1;
        for (size_t v = 0; v < V; ++v) ++off[fwd[v] + 1];
        for (size_t y = 0; y < V; ++y) off[y + 1] += off[y];
        {
            std::vector<uint32_t> fill(off, off + V);
            for (size_t v = 0; v < V; ++v) src[fill[fwd[v]]++] = uint32_t(v);
        }
        k::MathGatherDesc g{};
        const uint32_t* d_csr = static_cast<const uint32_t*>(small_upload(csr.data(), csr.size() * sizeof(uint32_t)));
        g.d_inv_off = d_csr;
        g.d_inv_src = d_csr + V + 1;
        g.ctrl_mask = d.ctrl_mask;
        // runs of consecutive positions
        for (size_t i = 0; i < flat;) {
            size_t j = i + 1;
            while (j < flat && d.reg_pos[j] == d.reg_pos[j - 1] + 1) ++j;
            g.seg[g.n_segs++] = {d.reg_pos[i], uint8_t(j - i), uint8_t(i)};
            i = j;
        }
        for (size_t i = 0; i < flat; ++i) g.reg_mask |= uint64_t(1) << d.reg_pos[i];
        ensure_scratch(*scratch1_, bytes);
        k::emulate_math_gather(ctx(), psi(), scratch1_->amps(), local_amps(), g);
        std::swap(state_, scratch1_);
        return;
    }
    // Registers too wide to tabulate: the closed forms are inverted in closed form (kernels.cuh MathInverseDesc) whenever the
    // gate's own preconditions hold — 0 <= a < N <= 2^bits for every register, gcd(a, N) = 1 for the multiplication.
    if (mode != k::MATH_TABLE) {
        bool ok = true;
        unsigned long long a_sub = 0, barrett = 0;
        if (mode == k::MATH_ADD) {
            a_sub = (unsigned long long)a;
        } else {
            ok = a >= 0 && N > 0 && a < N;
            for (size_t r = 0; r < n_regs && ok; ++r) {
                const int nb = d.reg_off[r + 1] - d.reg_off[r];
                if (nb < 63 && (unsigned long long)N > (1ull << nb)) ok = false;
                if (nb >= 63) ok = false;
            }
            a_sub = (unsigned long long)a;
            if (ok && mode == k::MATH_MUL_MOD) {
                // a^-1 mod N by the extended Euclidean algorithm
                long long r0 = N, r1 = a, t0 = 0, t1 = 1;
                while (r1 != 0) {
                    const long long q = r0 / r1;
                    const long long r2 = r0 - q * r1, t2 = t0 - q * t1;
                    r0 = r1, r1 = r2, t0 = t1, t1 = t2;
                }
                if (r0 != 1) {
                    ok = false;  // not invertible: the map is not a permutation of [0, N)
                } else {
                    a_sub = (unsigned long long)(t0 < 0 ? t0 + N : t0);
                    if ((unsigned long long)N < (1ull << 32)) barrett = ~0ull / (unsigned long long)N;  // floor((2^64 - 1) / N)
                }
            }
        }
        if (ok) {
            k::MathInverseDesc g{};
            g.mode = mode;
            g.a_sub = a_sub;
            g.N = (unsigned long long)N;
            g.barrett = barrett;
            g.ctrl_mask = d.ctrl_mask;
            g.n_regs = int(n_regs);
            int n_seg = 0;
            for (size_t r = 0; r < n_regs; ++r) {
                g.seg_off[r] = n_seg;
                const size_t lo = size_t(d.reg_off[r]), hi = size_t(d.reg_off[r + 1]);
                g.nb[r] = uint8_t(hi - lo);
                for (size_t i = lo; i < hi;) {
                    size_t j = i + 1;
                    while (j < hi && d.reg_pos[j] == d.reg_pos[j - 1] + 1) ++j;
                    g.seg[n_seg++] = {d.reg_pos[i], uint8_t(j - i), uint8_t(i - lo)};
                    i = j;
                }
            }
            g.seg_off[n_regs] = n_seg;
            for (size_t i = 0; i < flat; ++i) g.reg_mask |= uint64_t(1) << d.reg_pos[i];
            ensure_scratch(*scratch1_, bytes);
            k::emulate_math_inverse(ctx(), psi(), scratch1_->amps(), local_amps(), g);
            std::swap(state_, scratch1_);
            return;
        }
    }
    if (mode == k::MATH_TABLE)
        d.d_table = static_cast<const unsigned long long*>(small_upload(table, table_len * sizeof(uint64_t)));
    ensure_scratch(*scratch1_, bytes);
    PQB_CHECK(cudaMemsetAsync(scratch1_->ptr(), 0, bytes, stream_));
    k::emulate_math(ctx(), psi(), scratch1_->amps(), local_amps(), d);
    std::swap(state_, scratch1_);
}

// ---------------------------------------------------------------------------------------------------------------
// Pauli-string operators
// ---------------------------------------------------------------------------------------------------------------
std::vector<k::PauliTerm> Engine::build_terms(const TermsView& t, const uint32_t* ids, size_t n_ids, bool skip_identity,
                                              double* id_re, double* id_im) {
    std::vector<k::PauliTerm> out;
    out.reserve(t.n_terms);
    if (id_re) *id_re = 0.0;
    if (id_im) *id_im = 0.0;
    for (size_t ti = 0; ti < t.n_terms; ++ti) {
        double cre = t.complex_coeff ? t.coeff[2 * ti] : t.coeff[ti];
        double cim = t.complex_coeff ? t.coeff[2 * ti + 1] : 0.0;
        const size_t b = t.offsets[ti], e = t.offsets[ti + 1];
        if (b == e && skip_identity) {
            if (id_re) *id_re += cre;
            if (id_im) *id_im += cim;
            continue;
        }
        // compose the one-qubit Paulis left to right (apply_term queues them in this order, simulator.hpp:545-548) into
        // phase * X^x Z^z, in *logical* bits first
        uint64_t x = 0, z = 0;
        int quarter = 0;  // phase = i^quarter
        for (size_t j = b; j < e; ++j) {
            if (t.qubit_index[j] >= n_ids) throw ValueErr("qubit operator acts on a qubit outside the given register");
            const uint32_t lp = pos_of(ids[t.qubit_index[j]], "Pauli operator: Unknown qubit id.");
            const uint64_t bit = uint64_t(1) << lp;
            switch (t.pauli[j]) {
                case 'X': x ^= bit; break;
                case 'Z':
                    if (x & bit) quarter += 2;
                    z ^= bit;
                    break;
                case 'Y':
                    quarter += 1;
                    if (x & bit) quarter += 2;
                    x ^= bit;
                    z ^= bit;
                    break;
                default: throw ValueErr("Pauli operator must be 'X', 'Y' or 'Z'");
            }
        }
        switch (quarter & 3) {
            case 1: { const double r = -cim; cim = cre; cre = r; break; }
            case 2: cre = -cre; cim = -cim; break;
            case 3: { const double r = cim; cim = -cre; cre = r; break; }
            default: break;
        }
        out.push_back(k::PauliTerm{x, z, cre, cim});
    }
    return out;
}

// ---- tiled execution of a Pauli-string operator (kernels.cuh pauli_tile_pass) -------------------------------------------
// Cover the X/Y supports of the terms with sets of tile bits.  Every set holds index bits 0 and 1 (so that a tile is read
// in runs of at least 64 bytes) plus the bits that the most still-uncovered terms need; a term is applied by the first
// launch whose tile bits contain its whole xmask (z factors are signs and never need a partner amplitude).  Terms whose
// support is too large for any tile are returned in `wide` and go through per-term global gathers.
int pauli_block_bits() {
    static const int bits = [] {
        const char* e = std::getenv("PQB_PAULI_BLOCK_BITS");
        // Off by default: measured on TFIM-28 (profiles/r2_pauli_tile_history.md) the fused launch does keep the second
        // read of the vector in L2 (DRAM reads 12.9 -> 10.3 GB per application) but the launches are bound by the shared-
        // memory pipe, not by HBM, and the chained sets wait on one another: 8.4 ms instead of 7.6 ms per application.
        const int b = e ? std::atoi(e) : 0;
        return b < 0 ? 0 : (b > 40 ? 40 : b);
    }();
    return bits;
}

PauliPlan plan_pauli_tiles(const std::vector<k::PauliTerm>& terms, int L, int block_bits) {
    PauliPlan plan;
    plan.block_bits = block_bits;
    plan.launch_of_term.assign(terms.size(), -1);
    const int T = std::min(k::kTileBits, L);
    const uint64_t low2 = L >= 2 ? 3 : (L == 1 ? 1 : 0);
    std::vector<int> todo;
    for (size_t i = 0; i < terms.size(); ++i) {
        if (__builtin_popcountll(terms[i].xmask | low2) > T)
            plan.wide.push_back(terms[i]);
        else
            todo.push_back(int(i));
    }
    auto emit = [&](uint64_t S, const std::vector<int>& members) {
        // fill up with the lowest unused bits: a full-size tile keeps the loads long and the grid the same for every launch
        for (int b = 0; b < L && __builtin_popcountll(S) < T; ++b) S |= uint64_t(1) << b;
        k::PauliTileArgs base{};
        base.T = T;
        int n = 0;
        for (int b = 0; b < L; ++b)
            if ((S >> b) & 1) base.tile_pos[n++] = uint8_t(b);
        base.T_lo = 0;
        while (base.T_lo < T && base.tile_pos[base.T_lo] == base.T_lo) ++base.T_lo;
        base.n_tiles = uint64_t(1) << (L - T);
        // diagonal terms with z inside the tile -> one table for the whole pass
        std::vector<double2> table;
        std::vector<int> generic, outside;
        for (int i : members) {
            const k::PauliTerm& tm = terms[i];
            if (tm.xmask != 0) {
                generic.push_back(i);
            } else if ((tm.zmask & ~S) == 0) {
                plan.launch_of_term[i] = int(plan.launches.size());  // the table travels with the first launch of this set
                if (table.empty()) table.assign(size_t(1) << T, make_double2(0.0, 0.0));
                const uint32_t zl = uint32_t(extract_bits(tm.zmask, base.tile_pos, T));
                for (uint32_t t = 0; t < (1u << T); ++t) {
                    const bool neg = __builtin_popcount(t & zl) & 1;
                    table[t].x += neg ? -tm.cre : tm.cre;
                    table[t].y += neg ? -tm.cim : tm.cim;
                }
            } else if ((tm.zmask & S) == 0) {
                outside.push_back(i);
            } else {
                generic.push_back(i);
            }
        }
        // order the terms by where the kernel finds the partner amplitude: in the thread's registers (the x bits are all
        // among the top tile-coordinate bits, which number a thread's own elements) or in the shared tile
        constexpr uint32_t kThreadBits = 8;  // pauli_tile_kernel: 256 threads, element e of thread t = tile coordinate e * 256 + t
        auto klass = [&](int i) {
            const uint32_t xl = uint32_t(extract_bits(terms[i].xmask, base.tile_pos, T));
            return xl != 0 && T == k::kTileBits && (xl & ((1u << kThreadBits) - 1)) == 0 ? 0 : 1;
        };
        std::stable_sort(generic.begin(), generic.end(), [&](int x, int y) { return klass(x) < klass(y); });
        size_t at_g = 0, at_o = 0;
        bool first_chunk = true;
        do {
            k::PauliTileArgs a = base;
            for (; at_g < generic.size() && a.n_terms < k::kTileTerms; ++at_g) {
                const k::PauliTerm& tm = terms[generic[at_g]];
                plan.launch_of_term[generic[at_g]] = int(plan.launches.size());
                const uint32_t xl = uint32_t(extract_bits(tm.xmask, base.tile_pos, T));
                const uint32_t zl = uint32_t(extract_bits(tm.zmask, base.tile_pos, T));
                // sign on the source coordinate t ^ xl: parity((t ^ xl) & zl) = parity(t & zl) ^ parity(xl & zl)
                const double sgn = (__builtin_popcount(xl & zl) & 1) ? -1.0 : 1.0;
                a.coef[a.n_terms] = make_double2(sgn * tm.cre, sgn * tm.cim);
                a.xl[a.n_terms] = xl;
                a.zl[a.n_terms] = zl;
                a.z_out[a.n_terms] = tm.zmask & ~S;
                a.general[a.n_terms] = (zl != 0 || tm.cim != 0.0) ? 1 : 0;
                const int c = klass(generic[at_g]);
                ++a.n_terms;
                if (c == 0) a.n_reg = a.n_terms;
            }
            for (; at_o < outside.size() && a.n_outside < k::kTileTerms; ++at_o) {
                const k::PauliTerm& tm = terms[outside[at_o]];
                plan.launch_of_term[outside[at_o]] = int(plan.launches.size());
                a.coef_outside[a.n_outside] = make_double2(tm.cre, tm.cim);
                a.z_outside[a.n_outside] = tm.zmask;
                ++a.n_outside;
            }
            long at = -1;
            if (first_chunk && !table.empty()) {
                a.w_real = 1;
                for (const double2& w : table)
                    if (w.y != 0.0) a.w_real = 0;
                at = long(plan.tables.size());
                plan.tables.insert(plan.tables.end(), table.begin(), table.end());
            }
            plan.launches.push_back(a);
            plan.table_at.push_back(at);
            plan.tile_mask.push_back(S);
            first_chunk = false;
        } while (at_g < generic.size() || at_o < outside.size());
    };
    // Sets whose tile bits all lie below block_bits come first: the engine runs them fused, block by block, and what one
    // hands to the next stays in L2 (kernels.cuh PauliFusedArgs).  The sets that need higher bits follow.
    const uint64_t below_block = block_bits > T && block_bits < L ? (uint64_t(1) << block_bits) - 1 : ~uint64_t(0);
    auto cover = [&](std::vector<int>& todo, uint64_t allowed) {
        while (!todo.empty()) {
            uint64_t S = low2 | terms[todo[0]].xmask;  // the oldest uncovered term always fits: progress is guaranteed
            while (__builtin_popcountll(S) < T) {
                int count[64] = {0};
                for (int i : todo) {
                    const uint64_t extra = terms[i].xmask & ~S;
                    if (extra == 0 || __builtin_popcountll(S | terms[i].xmask) > T) continue;
                    for (uint64_t m = extra; m; m &= m - 1) ++count[__builtin_ctzll(m)];
                }
                int best = -1;
                for (int b = 0; b < L; ++b)
                    if (((allowed >> b) & 1) && count[b] > 0 && (best < 0 || count[b] > count[best])) best = b;
                if (best < 0) break;
                S |= uint64_t(1) << best;
            }
            std::vector<int> members, rest;
            for (int i : todo) ((terms[i].xmask & ~S) == 0 ? members : rest).push_back(i);
            emit(S, members);
            todo.swap(rest);
        }
    };
    std::vector<int> inside, beyond;
    for (int i : todo) ((terms[i].xmask & ~below_block) == 0 ? inside : beyond).push_back(i);
    cover(inside, below_block);
    cover(beyond, ~uint64_t(0));
    if (plan.launches.empty()) emit(low2, {});  // no tile-able term: an empty launch still finalises (scale / accumulate)
    return plan;
}

double Engine::get_expectation_value(const TermsView& t, const uint32_t* ids, size_t n_ids) {
    run();
    auto all_terms = build_terms(t, ids, n_ids, false, nullptr, nullptr);  // logical masks
    double* d_acc = d_scalars_;
    PQB_CHECK(cudaMemsetAsync(d_acc, 0, sizeof(double), stream_));
    // In a sharded run the X-support of a term must sit on local bits.  Terms are taken in batches whose combined
    // X-support fits on the device; each batch costs at most one remap of the state (nothing else has to move because
    // the expectation value is read-only and additive).  On one GPU this is a single batch.
    std::vector<char> done(all_terms.size(), 0);
    size_t left = all_terms.size();
    while (left > 0) {
        std::vector<k::PauliTerm> batch;
        uint64_t need_mask = 0;
        for (size_t i = 0; i < all_terms.size(); ++i) {
            if (done[i]) continue;
            const uint64_t merged = need_mask | all_terms[i].xmask;
            if (dist_ && __builtin_popcountll(merged) > L_) continue;
            need_mask = merged;
            batch.push_back(all_terms[i]);
            done[i] = 1;
            --left;
        }
        if (batch.empty()) {
            // what is left flips more qubits than one shard holds: no remap brings those partners on-device.  Apply the
            // rest of the operator with the partner amplitudes read from the peers' shards, then <psi|u>.
            std::vector<k::PauliTerm> rest;
            for (size_t i = 0; i < all_terms.size(); ++i)
                if (!done[i]) rest.push_back(all_terms[i]);
            PauliProgram prog = build_pauli_program(rest);
            ensure_scratch(*scratch1_, local_amps() * sizeof(double2));
            const std::vector<const double2*> src = pauli_sources(prog, *state_);
            dist_->barrier_on_stream();  // every rank's state is final before anybody reads it
            run_pauli_program(prog, src, scratch1_->amps(), 1.0, 0.0, nullptr, 0, nullptr);
            dist_->barrier_on_stream();  // the state may change again only after every peer has read it
            k::dot_real(ctx(), psi(), scratch1_->amps(), local_amps(), d_partials_, d_acc, true);
            break;
        }
        if (dist_) {
            std::vector<uint32_t> need;
            for (int p = 0; p < n_; ++p)
                if ((need_mask >> p) & 1) need.push_back(uint32_t(p));
            make_local(need);
        }
        const bool active = !dist_ || (uint64_t(rank_) & dist_->free_rank_bits_mask()) == 0;
        if (!active) continue;
        // tile-able terms: one read of the state per tile-bit set; the rest: one pair sweep per distinct xmask
        PauliProgram prog = build_pauli_program(batch);  // the batch's X/Y support is on-device: one group, this rank's shard
        if (prog.reads_peers()) throw RuntimeErr("internal: X on a rank bit was not remapped");
        PauliPlan& plan = prog.groups[0].plan;
        // consecutive fusable launches go out as one; an empty launch (nothing but a finalisation) is skipped
        for (size_t i = 0; i < plan.launches.size();) {
            k::PauliTileArgs& a = plan.launches[i];
            if (a.n_terms == 0 && a.n_outside == 0 && a.w_in == nullptr) {
                ++i;
                continue;
            }
            size_t j = i;
            for (; j < plan.launches.size(); ++j) {
                k::PauliTileArgs& b = plan.launches[j];
                if (b.n_terms == 0 && b.n_outside == 0 && b.w_in == nullptr) break;
                b.expectation = 1;
            }
            run_tile_launches(plan, i, j, psi(), nullptr, nullptr, d_acc);
            i = j;
        }
        size_t i = 0;
        while (i < plan.wide.size()) {
            size_t j = i;
            while (j < plan.wide.size() && plan.wide[j].xmask == plan.wide[i].xmask && j - i < 64) ++j;
            k::pauli_expectation_group(ctx(), psi(), L_, plan.wide[i].xmask, &plan.wide[i], int(j - i), d_partials_, d_acc);
            i = j;
        }
    }
    return allreduce_sum(read_scalar(d_acc));
}

// X/Y support of a term list as logical positions
static std::vector<uint32_t> x_support(const std::vector<k::PauliTerm>& terms, int n) {
    uint64_t m = 0;
    for (auto& tm : terms) m |= tm.xmask;
    std::vector<uint32_t> need;
    for (int p = 0; p < n; ++p)
        if ((m >> p) & 1) need.push_back(uint32_t(p));
    return need;
}

Engine::PauliProgram Engine::build_pauli_program(const std::vector<k::PauliTerm>& logical) {
    std::map<int, std::vector<k::PauliTerm>> by_rank;
    by_rank[0];  // the own-shard group always exists (it finalises even when it has no term)
    for (auto t : logical) {
        uint64_t x = 0, z = 0;
        int xr = 0, zr = 0;
        for (int p = 0; p < n_; ++p) {
            const uint64_t bit = uint64_t(1) << p;
            if (loc_[p] < 64) {
                if (t.xmask & bit) x |= uint64_t(1) << loc_[p];
                if (t.zmask & bit) z |= uint64_t(1) << loc_[p];
            } else {
                if (t.xmask & bit) xr |= 1 << (loc_[p] - 64);
                if (t.zmask & bit) zr |= 1 << (loc_[p] - 64);
            }
        }
        // the sign (-1)^{popcount(s & zmask)} is taken on the SOURCE index s = j ^ xmask: its rank bits are rank ^ xr
        if (__builtin_popcount(unsigned((rank_ ^ xr) & zr)) & 1) {
            t.cre = -t.cre;
            t.cim = -t.cim;
        }
        t.xmask = x;
        t.zmask = z;
        by_rank[xr].push_back(t);
    }
    PauliProgram prog;
    // peer groups first, the own-shard group last: its last launch scales / accumulates
    for (auto it = by_rank.rbegin(); it != by_rank.rend(); ++it) {
        PauliProgram::Group g;
        g.xr = it->first;
        std::stable_sort(it->second.begin(), it->second.end(),
                         [](const k::PauliTerm& a, const k::PauliTerm& b) { return a.xmask < b.xmask; });
        g.plan = plan_pauli_tiles(it->second, L_, pauli_block_bits());
        prog.groups.push_back(std::move(g));
    }
    // tables and wide terms of every group go to the device in one upload
    std::vector<char> blob;
    std::vector<size_t> table_off, wide_off;
    for (auto& g : prog.groups) {
        table_off.push_back(blob.size());
        const char* tb = reinterpret_cast<const char*>(g.plan.tables.data());
        blob.insert(blob.end(), tb, tb + g.plan.tables.size() * sizeof(double2));
    }
    for (auto& g : prog.groups) {
        wide_off.push_back(blob.size());
        const char* wb = reinterpret_cast<const char*>(g.plan.wide.data());
        blob.insert(blob.end(), wb, wb + g.plan.wide.size() * sizeof(k::PauliTerm));
    }
    if (!blob.empty()) {
        const char* d = static_cast<const char*>(small_upload(blob.data(), blob.size()));
        for (size_t gi = 0; gi < prog.groups.size(); ++gi) {
            PauliPlan& plan = prog.groups[gi].plan;
            for (size_t i = 0; i < plan.launches.size(); ++i)
                plan.launches[i].w_in =
                    plan.table_at[i] >= 0 ? reinterpret_cast<const double2*>(d + table_off[gi]) + plan.table_at[i] : nullptr;
            plan.d_wide = plan.wide.empty() ? nullptr : reinterpret_cast<const k::PauliTerm*>(d + wide_off[gi]);
        }
    }
    return prog;
}

// where every group of the program reads from when the operator is applied to `buf` (this rank's vector of that role)
std::vector<const double2*> Engine::pauli_sources(const PauliProgram& prog, const GrowBuffer& buf) {
    std::vector<const double2*> src;
    for (auto& g : prog.groups) {
        if (g.xr == 0) {
            src.push_back(buf.amps());
            continue;
        }
        const double2* p = nullptr;
        try {
            p = dist_ ? dist_->peer_buffer(rank_ ^ g.xr, buf) : nullptr;
        } catch (const std::runtime_error& e) {
            throw CudaErr(e.what());
        }
        if (!p)
            throw RuntimeErr("the operator flips more qubits than one shard holds, which needs peer-mapped shards "
                             "(not available here: PQB_REMAP_P2P=0 or no peer access between the GPUs)");
        src.push_back(p);
    }
    return src;
}

// u <- scale * sum_t c_t P_t in   (and optionally acc += u on the control subspace with |u|^2 summed into d_norm)
void Engine::run_pauli_program(PauliProgram& prog, const std::vector<const double2*>& src, double2* u, double sre, double sim,
                               double2* acc, uint64_t cmask, double* d_norm) {
    bool first = true;
    for (size_t gi = 0; gi < prog.groups.size(); ++gi) {
        PauliPlan& plan = prog.groups[gi].plan;
        const bool last_group = gi + 1 == prog.groups.size();
        if (!plan.wide.empty()) {
            k::pauli_gather_accumulate(ctx(), src[gi], u, local_amps(), plan.d_wide, int(plan.wide.size()), first);
            first = false;
        }
        // flags first, then the launches (consecutive fusable ones as one launch)
        std::vector<size_t> live;
        for (size_t i = 0; i < plan.launches.size(); ++i) {
            k::PauliTileArgs& a = plan.launches[i];
            if (!last_group && a.n_terms == 0 && a.n_outside == 0 && a.w_in == nullptr) continue;  // nothing to add
            a.first = first ? 1 : 0;
            a.final = last_group && i + 1 == plan.launches.size() ? 1 : 0;
            a.expectation = 0;
            a.sre = sre;
            a.sim = sim;
            a.cmask = cmask;
            first = false;
            live.push_back(i);
        }
        for (size_t p = 0; p < live.size();) {
            size_t q = p + 1;
            while (q < live.size() && live[q] == live[q - 1] + 1) ++q;  // a run of consecutive launches
            const bool finalises = plan.launches[live[q - 1]].final != 0;
            run_tile_launches(plan, live[p], live[q - 1] + 1, src[gi], u, finalises ? acc : nullptr, d_norm);
            p = q;
        }
    }
}

// d_sum: expectation launches add their <psi|.|psi> contributions to it; a finalising launch with an accumulator overwrites
// it with the squared norm of what it added
void Engine::run_tile_launches(PauliPlan& plan, size_t first, size_t last, const double2* in, double2* u, double2* acc,
                               double* d_sum) {
    for (size_t i = first; i < last;) {
        size_t j = i + 1;
        if (plan.fusable(i, L_))
            while (j < last && j - i < size_t(k::kTileSets) && plan.fusable(j, L_)) ++j;
        const k::PauliTileArgs& tail = plan.launches[j - 1];
        double2* acc_here = tail.final ? acc : nullptr;
        const int grid = k::pauli_tile_pass(ctx(), in, u, acc_here, &plan.launches[i], int(j - i), plan.block_bits, d_partials_,
                                            d_pauli_sync_);
        if (d_sum != nullptr) {
            if (tail.expectation)
                k::reduce_partials(ctx(), d_partials_, grid, d_sum, true);
            else if (tail.final && acc_here != nullptr)
                k::reduce_partials(ctx(), d_partials_, grid, d_sum, false);
        }
        i = j;
    }
}

void Engine::apply_qubit_operator(const TermsView& t, const uint32_t* ids, size_t n_ids) {
    run();
    auto terms = build_terms(t, ids, n_ids, false, nullptr, nullptr);
    if (dist_) {
        // bring the X/Y support on-device when it fits; otherwise the partner amplitudes are read from the peers' shards
        const std::vector<uint32_t> need = x_support(terms, n_);
        if (int(need.size()) <= L_) make_local(need);
    }
    PauliProgram prog = build_pauli_program(terms);
    const size_t bytes = local_amps() * sizeof(double2);
    ensure_scratch(*scratch1_, bytes);
    const std::vector<const double2*> src = pauli_sources(prog, *state_);
    const bool peers = prog.reads_peers();
    if (peers) dist_->barrier_on_stream();  // every rank's state is final before anybody reads it
    run_pauli_program(prog, src, scratch1_->amps(), 1.0, 0.0, nullptr, 0, nullptr);
    if (peers) dist_->barrier_on_stream();  // nobody reuses its old state buffer while a peer still reads it
    std::swap(state_, scratch1_);
}

void Engine::emulate_time_evolution(const TermsView& t, double time, const uint32_t* ids, size_t n_ids,
                                    const uint32_t* ctrl, size_t nc) {
    run();
    double tr = 0.0;
    auto terms = build_terms(t, ids, n_ids, true, &tr, nullptr);
    double op_nrm = 0.0;
    // |c| of the caller's coefficients (phases i^k folded in by build_terms do not change the modulus)
    for (auto& tm : terms) op_nrm += std::hypot(tm.cre, tm.cim);
    const unsigned s = unsigned(std::fabs(time) * op_nrm + 1.);
    const std::complex<double> correction = std::exp(std::complex<double>(0.0, -time * tr / double(s)));
    std::vector<uint32_t> cl;
    for (size_t i = 0; i < nc; ++i) cl.push_back(pos_of(ctrl[i], "emulate_time_evolution(): Unknown control qubit id."));
    if (dist_) {
        const std::vector<uint32_t> need = x_support(terms, n_);
        if (int(need.size()) <= L_) make_local(need);
    }
    PauliProgram prog = build_pauli_program(terms);
    uint64_t cmask = 0;
    bool active = true;  // accumulation is masked by the controls; H itself acts everywhere (simulator.hpp:411-425)
    for (auto lp : cl) {
        if (is_local(lp))
            cmask |= uint64_t(1) << loc_[lp];
        else if (!((rank_ >> (loc_[lp] - 64)) & 1))
            active = false;
    }
    const size_t bytes = local_amps() * sizeof(double2);
    ensure_scratch(*scratch1_, bytes);
    ensure_scratch(*scratch2_, bytes);
    // the Taylor vectors alternate between the two scratch buffers; with peer groups both are read by the partners
    const bool peers = prog.reads_peers();
    std::vector<const double2*> src_v = pauli_sources(prog, *scratch1_), src_u = pauli_sources(prog, *scratch2_);
    double* d_norm = d_scalars_;
    for (unsigned i = 0; i < s; ++i) {
        double2* v = scratch1_->amps();
        double2* u = scratch2_->amps();
        std::vector<const double2*>* sv = &src_v;
        std::vector<const double2*>* su = &src_u;
        PQB_CHECK(cudaMemcpyAsync(v, psi(), bytes, cudaMemcpyDeviceToDevice, stream_));
        if (peers) dist_->barrier_on_stream();  // every rank's v is in place before a partner reads it
        double nrm_change = 1.0;
        for (unsigned kk = 0; nrm_change > 1.e-12; ++kk) {
            // coeff = (-time * I) / (s * (k + 1))
            const double cim = -time / double(s * (kk + 1));
            if (active) {
                run_pauli_program(prog, *sv, u, 0.0, cim, psi(), cmask, d_norm);
                nrm_change = read_scalar(d_norm);
            } else {
                run_pauli_program(prog, *sv, u, 0.0, cim, nullptr, 0, nullptr);
                nrm_change = 0.0;
            }
            // (the all-reduce is also the rendezvous that keeps a rank from overwriting a vector a partner still reads)
            nrm_change = std::sqrt(allreduce_sum(nrm_change));
            std::swap(v, u);
            std::swap(sv, su);
        }
        if (active) k::scale_masked(ctx(), psi(), local_amps(), cmask, correction.real(), correction.imag());
    }
}

// ---------------------------------------------------------------------------------------------------------------
// additions
// ---------------------------------------------------------------------------------------------------------------
void Engine::apply_gate_stream(const void* packed, size_t n_bytes, size_t n_gates, bool fuse) {
    const uint8_t* p = static_cast<const uint8_t*>(packed);
    const uint8_t* end = p + n_bytes;
    std::vector<double> m;
    for (size_t g = 0; g < n_gates; ++g) {
        if (p + 8 > end) throw ValueErr("apply_gate_stream(): truncated stream");
        uint32_t k, nc;
        std::memcpy(&k, p, 4);
        std::memcpy(&nc, p + 4, 4);
        p += 8;
        if (k == 0 || k > 5 || nc > 64) throw ValueErr("apply_gate_stream(): bad gate header");
        const size_t d = size_t(1) << k;
        const size_t need = 4 * (k + nc) + 16 * d * d;
        if (p + need > end) throw ValueErr("apply_gate_stream(): truncated stream");
        uint32_t ids[8], ctrl[64];
        std::memcpy(ids, p, 4 * k);
        std::memcpy(ctrl, p + 4 * k, 4 * nc);
        m.resize(2 * d * d);
        std::memcpy(m.data(), p + 4 * (k + nc), 16 * d * d);
        p += need;
        apply_controlled_gate(m.data(), ids, k, ctrl, nc);
        if (!fuse) run();
    }
}

void Engine::init_random_state(uint32_t n_qubits, uint64_t seed) {
    fuser_.clear();
    if (n_qubits > 40) throw ValueErr("init_random_state(): too many qubits");
    map_.clear();
    loc_.clear();
    n_ = int(n_qubits);
    int g = 0;
    if (dist_) {
        g = std::min<int>(dist_->rank_bits(), std::max(0, n_ - kMinLocalBits));
        dist_->reset_rank_bits(g);
    }
    L_ = n_ - g;
    for (int p = 0; p < n_; ++p) {
        map_[uint32_t(p)] = uint32_t(p);
        loc_.push_back(uint8_t(p < L_ ? p : 64 + (p - L_)));
    }
    try {
        state_->ensure(local_amps() * sizeof(double2), stream_);
    } catch (const std::bad_alloc&) {
        scratch1_->release();
        scratch2_->release();
        try {
            state_->ensure(local_amps() * sizeof(double2), stream_);
        } catch (const std::bad_alloc&) {
            throw CudaErr("init_random_state(): out of device memory");
        }
    }
    const bool active = !dist_ || uint64_t(rank_) < (uint64_t(1) << g);
    if (active)
        k::init_random(ctx(), psi(), local_amps(), seed, uint64_t(rank_) << L_);
    else
        PQB_CHECK(cudaMemsetAsync(psi(), 0, local_amps() * sizeof(double2), stream_));
    const double nrm = norm_squared();
    k::scale_all(ctx(), psi(), local_amps(), 1.0 / std::sqrt(nrm));
}

// ---------------------------------------------------------------------------------------------------------------
// checkpoint / view (f3)
// ---------------------------------------------------------------------------------------------------------------
namespace {
constexpr size_t kIoChunk = size_t(64) << 20;  // pinned staging buffer between HBM and the file

std::string shard_file(const std::string& prefix, int rank, int world) {
    return prefix + ".rank" + std::to_string(rank) + "of" + std::to_string(world) + ".pqbs";
}

struct FileCloser {
    FILE* f;
    ~FileCloser() {
        if (f) fclose(f);
    }
};
}  // namespace

void Engine::save_state(const std::string& prefix) {
    run();
    const std::string path = shard_file(prefix, rank_, world_);
    FileCloser fc{fopen(path.c_str(), "wb")};
    if (!fc.f) throw RuntimeErr("save_state(): cannot open " + path);
    auto put = [&](const void* p, size_t n) {
        if (n && fwrite(p, 1, n, fc.f) != n) throw RuntimeErr("save_state(): write failed on " + path);
    };
    std::ostringstream rng_text;
    rng_text << rng_;
    const std::string rng = rng_text.str();
    const uint32_t head[8] = {1 /*version*/, uint32_t(n_), uint32_t(L_), uint32_t(rank_), uint32_t(world_),
                              uint32_t(rng.size()), 0, 0};
    const uint64_t free_mask = dist_ ? dist_->free_rank_bits_mask() : 0;
    put("PQBS", 4);
    put(head, sizeof(head));
    put(&free_mask, 8);
    for (auto& kv : map_) {
        const uint32_t e[2] = {kv.first, kv.second};
        put(e, 8);
    }
    put(loc_.data(), size_t(n_));
    put(rng.data(), rng.size());
    // the shard, as it lies in HBM
    void* pinned = nullptr;
    PQB_CHECK(cudaMallocHost(&pinned, kIoChunk));
    try {
        const size_t total = local_amps() * sizeof(double2);
        for (size_t off = 0; off < total; off += kIoChunk) {
            const size_t cnt = std::min(kIoChunk, total - off);
            PQB_CHECK(cudaMemcpyAsync(pinned, reinterpret_cast<const char*>(psi()) + off, cnt, cudaMemcpyDeviceToHost, stream_));
            PQB_CHECK(cudaStreamSynchronize(stream_));
            put(pinned, cnt);
        }
    } catch (...) {
        cudaFreeHost(pinned);
        throw;
    }
    cudaFreeHost(pinned);
    if (fflush(fc.f) != 0) throw RuntimeErr("save_state(): write failed on " + path);
}

void Engine::load_state(const std::string& prefix) {
    run();
    const std::string path = shard_file(prefix, rank_, world_);
    FileCloser fc{fopen(path.c_str(), "rb")};
    if (!fc.f) throw RuntimeErr("load_state(): cannot open " + path);
    auto get = [&](void* p, size_t n) {
        if (n && fread(p, 1, n, fc.f) != n) throw RuntimeErr("load_state(): " + path + " is truncated");
    };
    char magic[4];
    uint32_t head[8];
    uint64_t free_mask = 0;
    get(magic, 4);
    get(head, sizeof(head));
    get(&free_mask, 8);
    if (std::memcmp(magic, "PQBS", 4) != 0 || head[0] != 1) throw ValueErr("load_state(): " + path + " is not a state checkpoint");
    const uint32_t n = head[1], L = head[2];
    if (int(head[3]) != rank_ || int(head[4]) != world_)
        throw ValueErr("load_state(): the checkpoint was written by rank " + std::to_string(head[3]) + " of " +
                       std::to_string(head[4]) + ", this engine is rank " + std::to_string(rank_) + " of " +
                       std::to_string(world_));
    if (n > 62 || L > n || head[5] > (1u << 20)) throw ValueErr("load_state(): corrupt header in " + path);
    std::map<uint32_t, uint32_t> map;
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t e[2];
        get(e, 8);
        if (e[1] >= n) throw ValueErr("load_state(): corrupt qubit map in " + path);
        map[e[0]] = e[1];
    }
    std::vector<uint8_t> loc(n);
    get(loc.data(), n);
    std::string rng(head[5], '\0');
    get(&rng[0], rng.size());
    if (map.size() != n) throw ValueErr("load_state(): corrupt qubit map in " + path);
    const size_t total = (sizeof(double2) << L);
    try {
        state_->ensure(total, stream_);
    } catch (const std::bad_alloc&) {
        scratch1_->release();
        scratch2_->release();
        try {
            state_->ensure(total, stream_);
        } catch (const std::bad_alloc&) {
            throw CudaErr("load_state(): out of device memory");
        }
    }
    void* pinned = nullptr;
    PQB_CHECK(cudaMallocHost(&pinned, kIoChunk));
    try {
        for (size_t off = 0; off < total; off += kIoChunk) {
            const size_t cnt = std::min(kIoChunk, total - off);
            get(pinned, cnt);
            PQB_CHECK(cudaMemcpyAsync(reinterpret_cast<char*>(state_->ptr()) + off, pinned, cnt, cudaMemcpyHostToDevice, stream_));
            PQB_CHECK(cudaStreamSynchronize(stream_));
        }
    } catch (...) {
        cudaFreeHost(pinned);
        throw;
    }
    cudaFreeHost(pinned);
    // commit the bookkeeping only after the data is in place
    map_.swap(map);
    loc_.swap(loc);
    n_ = int(n);
    L_ = int(L);
    if (dist_) dist_->set_free_rank_bits_mask(free_mask);
    std::istringstream rng_text(rng);
    rng_text >> rng_;
}

void Engine::state_view(void** ptr, uint64_t* n_amps, uint8_t* layout, size_t cap, size_t* n_qubits) {
    run();
    PQB_CHECK(cudaStreamSynchronize(stream_));
    if (ptr) *ptr = state_->ptr();
    if (n_amps) *n_amps = local_amps();
    if (n_qubits) *n_qubits = size_t(n_);
    if (layout)
        for (int p = 0; p < n_ && size_t(p) < cap; ++p) layout[p] = loc_[p];
}

void Engine::synchronize() {
    PQB_CHECK(cudaStreamSynchronize(stream_));
    check_exchange_error();
}

void Engine::timer_start() { PQB_CHECK(cudaEventRecord(ev0_, stream_)); }

double Engine::timer_stop() {
    PQB_CHECK(cudaEventRecord(ev1_, stream_));
    PQB_CHECK(cudaEventSynchronize(ev1_));
    float ms = 0.f;
    PQB_CHECK(cudaEventElapsedTime(&ms, ev0_, ev1_));
    return double(ms);
}

void Engine::flush_l2(size_t bytes) {
    if (bytes > d_flush_cap_) {
        PQB_CHECK(cudaStreamSynchronize(stream_));
        if (d_flush_) cudaFree(d_flush_);
        PQB_CHECK(cudaMalloc(&d_flush_, bytes));
        d_flush_cap_ = bytes;
    }
    k::flush_l2(ctx(), d_flush_, bytes / sizeof(double));
}

double Engine::bench_dense_pass(const double* m, const uint32_t* positions, size_t kq, uint64_t ctrl_mask, int repeats) {
    run();
    if (kq == 0 || kq > 5) throw ValueErr("bench_dense_pass(): k must be 1..5");
    uint8_t tpos[8], cpos[64];
    int nc = 0;
    for (size_t i = 0; i < kq; ++i) {
        if (int(positions[i]) >= L_) throw ValueErr("bench_dense_pass(): position outside the local state");
        tpos[i] = uint8_t(positions[i]);
    }
    std::sort(tpos, tpos + kq);
    for (int b = 0; b < L_; ++b)
        if ((ctrl_mask >> b) & 1) cpos[nc++] = uint8_t(b);
    if (repeats < 1) repeats = 1;
    k::apply_dense(ctx(), psi(), L_, int(kq), tpos, nc, cpos, m);  // warm-up
    timer_start();
    for (int r = 0; r < repeats; ++r) k::apply_dense(ctx(), psi(), L_, int(kq), tpos, nc, cpos, m);
    return timer_stop() / repeats;
}

void Engine::selftest_sliced_pass(const double* m, const uint32_t* positions, size_t kq, uint64_t ctrl_mask,
                                  uint64_t slice_mask) {
    run();
    if (kq == 0 || kq > 5) throw ValueErr("selftest_sliced_pass(): k must be 1..5");
    Launch l;
    l.kind = Launch::DENSE;
    l.k = int(kq);
    uint64_t used = ctrl_mask;
    for (size_t i = 0; i < kq; ++i) {
        if (int(positions[i]) >= L_) throw ValueErr("selftest_sliced_pass(): position outside the local state");
        l.tpos[i] = uint8_t(positions[i]);
        used |= uint64_t(1) << positions[i];
    }
    if (!std::is_sorted(l.tpos, l.tpos + kq)) throw ValueErr("selftest_sliced_pass(): positions must ascend");
    if ((used & slice_mask) != 0 || (L_ < 64 && (slice_mask >> L_) != 0))
        throw ValueErr("selftest_sliced_pass(): slice bits must be free local bits");
    for (int b = 0; b < L_; ++b)
        if ((ctrl_mask >> b) & 1) l.cpos[l.n_ctrl++] = uint8_t(b);
    l.m.assign(m, m + (size_t(2) << (2 * kq)));
    k::Slice sl;
    for (int b = 0; b < L_; ++b)
        if ((slice_mask >> b) & 1) {
            if (sl.n >= 16) throw ValueErr("selftest_sliced_pass(): more than 16 slice bits");
            sl.pos[sl.n++] = uint8_t(b);
        }
    for (uint64_t v = 0; v < (uint64_t(1) << sl.n); ++v) {
        sl.val = deposit_bits(v, sl.pos, sl.n);
        launch(l, sl);
    }
}

double Engine::measure_fp64_peak() { return k::measure_fp64_tflops(ctx(), sm_count_); }

double Engine::measure_copy_bandwidth(size_t bytes) {
    ensure_scratch(*scratch1_, bytes);
    ensure_scratch(*scratch2_, bytes);
    PQB_CHECK(cudaMemsetAsync(scratch1_->ptr(), 1, bytes, stream_));
    double best = 0.0;
    for (int r = 0; r < 6; ++r) {
        timer_start();
        PQB_CHECK(cudaMemcpyAsync(scratch2_->ptr(), scratch1_->ptr(), bytes, cudaMemcpyDeviceToDevice, stream_));
        const double ms = timer_stop();
        const double gbs = 2.0 * double(bytes) / (ms * 1e-3) / 1e9;
        if (r > 0 && gbs > best) best = gbs;
    }
    return best;
}

}  // namespace pqb
