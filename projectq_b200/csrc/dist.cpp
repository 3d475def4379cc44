// Sharded state communicator — see dist.h.
#include "dist.h"
#include "kernels.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <unistd.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

namespace pqb {

namespace {

struct Nccl {
    void* lib = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
};

const Nccl& nccl() {
    static Nccl n = [] {
        Nccl x;
        x.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!x.lib) x.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!x.lib) throw std::runtime_error(std::string("cannot load libnccl.so.2: ") + dlerror());
        auto sym = [&](const char* name) {
            void* p = dlsym(x.lib, name);
            if (!p) throw std::runtime_error(std::string("libnccl lacks ") + name);
            return p;
        };
        x.GetUniqueId = reinterpret_cast<decltype(x.GetUniqueId)>(sym("ncclGetUniqueId"));
        x.CommInitRank = reinterpret_cast<decltype(x.CommInitRank)>(sym("ncclCommInitRank"));
        x.CommDestroy = reinterpret_cast<decltype(x.CommDestroy)>(sym("ncclCommDestroy"));
        x.AllReduce = reinterpret_cast<decltype(x.AllReduce)>(sym("ncclAllReduce"));
        x.Send = reinterpret_cast<decltype(x.Send)>(sym("ncclSend"));
        x.Recv = reinterpret_cast<decltype(x.Recv)>(sym("ncclRecv"));
        x.GroupStart = reinterpret_cast<decltype(x.GroupStart)>(sym("ncclGroupStart"));
        x.GroupEnd = reinterpret_cast<decltype(x.GroupEnd)>(sym("ncclGroupEnd"));
        x.GetErrorString = reinterpret_cast<decltype(x.GetErrorString)>(sym("ncclGetErrorString"));
        return x;
    }();
    return n;
}

void nccl_check(ncclResult_t r, const char* what) {
    if (r != ncclSuccess) throw std::runtime_error(std::string("NCCL error in ") + what + ": " + nccl().GetErrorString(r));
}

void cuda_check(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw std::runtime_error(std::string("CUDA error in ") + what + ": " + cudaGetErrorString(e));
}

}  // namespace

std::vector<std::pair<int, int>> plan_remap(std::vector<uint8_t>& loc, int n_local_bits,
                                            const std::vector<uint32_t>& need, const std::vector<uint32_t>* victims) {
    std::vector<std::pair<int, int>> swaps;
    std::vector<uint32_t> global_needed;
    for (auto lp : need) {
        if (lp >= loc.size()) throw std::runtime_error("plan_remap: logical position out of range");
        if (loc[lp] >= 64 && std::find(global_needed.begin(), global_needed.end(), lp) == global_needed.end())
            global_needed.push_back(lp);
    }
    if (global_needed.empty()) return swaps;
    auto needed = [&](uint32_t lp) { return std::find(need.begin(), need.end(), lp) != need.end(); };
    // eviction candidates: the caller's preference order, else the highest local bits (largest contiguous blocks)
    std::vector<uint32_t> order;
    if (victims) {
        for (auto lp : *victims)
            if (lp < loc.size() && loc[lp] < 64 && !needed(lp)) order.push_back(lp);
    } else {
        std::vector<int> owner(n_local_bits, -1);
        for (size_t p = 0; p < loc.size(); ++p)
            if (loc[p] < 64) owner[loc[p]] = int(p);
        for (int b = n_local_bits - 1; b >= 0; --b)
            if (owner[b] >= 0 && !needed(uint32_t(owner[b]))) order.push_back(uint32_t(owner[b]));
    }
    if (order.size() < global_needed.size())
        throw std::runtime_error("remap: not enough local qubits to bring every target of the gate on-device");
    for (size_t i = 0; i < global_needed.size(); ++i) {
        const uint32_t in = global_needed[i], out = order[i];
        const int r = loc[in] - 64, b = loc[out];
        swaps.emplace_back(r, b);
        loc[out] = uint8_t(64 + r);
        loc[in] = uint8_t(b);
    }
    return swaps;
}

InteractionGraph interaction_graph(const Fuser& fuser) {
    InteractionGraph adj;
    for (size_t gi = 0; gi < fuser.pending(); ++gi) {
        const Gate& gt = fuser.pending_gate(gi);
        std::vector<uint32_t> qs(gt.targets);
        qs.insert(qs.end(), gt.ctrls.begin(), gt.ctrls.end());
        for (auto a : qs)
            for (auto b : qs)
                if (a != b) adj[a].insert(b);
    }
    return adj;
}

RemapChoice choose_remap(const Fuser& fuser, const std::map<uint32_t, uint32_t>& map, const std::vector<uint8_t>& loc,
                         const InteractionGraph& adj) {
    RemapChoice out;
    std::vector<uint32_t>& need = out.need;
    auto is_local = [&](uint32_t lp) { return loc[lp] < 64; };
    auto in_need = [&](uint32_t lp) { return std::find(need.begin(), need.end(), lp) != need.end(); };
    // the oldest waiting gate has no unfinished predecessor, so it waits for one of its own qubits
    const Gate& g = fuser.pending_gate(0);
    for (auto t : g.targets) need.push_back(map.at(t));
    for (auto c : g.ctrls) need.push_back(map.at(c));
    // While an exchange is being paid for, bring in the other rank-bit qubits too if they are needed sooner than the local
    // qubits they would replace (plain Belady order for this decision).
    {
        std::vector<std::pair<size_t, uint32_t>> use0;  // (next use, logical position) of the local qubits
        std::vector<std::pair<size_t, uint32_t>> incoming;
        for (auto& kv : map) {
            if (is_local(kv.second))
                use0.emplace_back(fuser.next_use(kv.first), kv.second);
            else if (!in_need(kv.second))
                incoming.emplace_back(fuser.next_use(kv.first), kv.second);
        }
        std::sort(use0.begin(), use0.end(),
                  [](const std::pair<size_t, uint32_t>& a, const std::pair<size_t, uint32_t>& b) { return a.first > b.first; });
        std::sort(incoming.begin(), incoming.end());
        size_t n_global_needed = 0;
        for (auto lp : need)
            if (!is_local(lp)) ++n_global_needed;
        for (auto& in : incoming) {
            if (in.first == size_t(-1)) break;  // never used again
            size_t seen = 0;
            const std::pair<size_t, uint32_t>* victim = nullptr;
            for (auto& u : use0) {
                if (in_need(u.second)) continue;
                if (seen++ == n_global_needed) {
                    victim = &u;
                    break;
                }
            }
            if (!victim || victim->first <= in.first) break;  // the local qubit is needed sooner: keep it
            need.push_back(in.second);
            ++n_global_needed;
        }
    }
    struct Cand {
        size_t next_use;
        uint32_t id, pos;
    };
    std::vector<Cand> cands;
    std::set<uint32_t> off_device;  // ids on rank bits that stay there, plus victims chosen so far
    for (auto& kv : map) {
        if (in_need(kv.second)) continue;
        if (is_local(kv.second))
            cands.push_back({fuser.next_use(kv.first), kv.first, kv.second});
        else
            off_device.insert(kv.first);
    }
    std::vector<char> picked(cands.size(), 0);
    for (size_t round = 0; round < cands.size(); ++round) {
        int best = -1;
        bool best_touch = false;
        size_t best_deg = 0;
        for (size_t ci = 0; ci < cands.size(); ++ci) {
            if (picked[ci]) continue;
            const Cand& cd = cands[ci];
            bool touch = false;
            size_t deg = 0;
            auto it = adj.find(cd.id);
            if (it != adj.end()) {
                deg = it->second.size();
                for (auto o : it->second)
                    if (off_device.count(o)) {
                        touch = true;
                        break;
                    }
            }
            bool better;
            if (best < 0)
                better = true;
            else if (cd.next_use != cands[best].next_use)
                better = cd.next_use > cands[best].next_use;
            else if (touch != best_touch)
                better = touch;
            else if (deg != best_deg)
                better = deg < best_deg;
            else
                better = loc[cd.pos] > loc[cands[best].pos];
            if (better) {
                best = int(ci);
                best_touch = touch;
                best_deg = deg;
            }
        }
        picked[best] = 1;
        out.victims.push_back(cands[best].pos);
        off_device.insert(cands[best].id);
    }
    return out;
}

std::vector<ExchangePeer> plan_exchange(int rank, const std::vector<std::pair<int, int>>& swaps) {
    std::vector<ExchangePeer> out;
    const int g = int(swaps.size());
    uint64_t mine = 0;  // this rank's value of the exchanged rank bits
    for (int i = 0; i < g; ++i) mine |= uint64_t((rank >> swaps[i].first) & 1) << i;
    // round k pairs beta with beta ^ k: every round is a perfect matching, so the pairwise rounds never wait on a third rank
    for (uint64_t k = 1; k < (uint64_t(1) << g); ++k) {
        const uint64_t beta = mine ^ k;
        int peer = rank;
        uint64_t pattern = 0;
        for (int i = 0; i < g; ++i) {
            const int bit = int((beta >> i) & 1);
            peer = (peer & ~(1 << swaps[i].first)) | (bit << swaps[i].first);
            pattern |= uint64_t(bit) << swaps[i].second;
        }
        out.push_back({peer, pattern});
    }
    return out;
}

void Dist::get_unique_id(void* out128) {
    ncclUniqueId id;
    nccl_check(nccl().GetUniqueId(&id), "ncclGetUniqueId");
    static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
    std::memcpy(out128, &id, sizeof(id));
}

Dist::Dist(int rank, int world, const void* uid, cudaStream_t stream) : rank_(rank), world_(world), stream_(stream) {
    g_ = 0;
    while ((1 << g_) < world_) ++g_;
    free_mask_ = (uint64_t(1) << g_) - 1;
    ncclUniqueId id;
    std::memcpy(&id, uid, sizeof(id));
    {
        // the descriptor channel must listen before the (collective) communicator set-up returns on any rank
        const char* e = getenv("PQB_REMAP_P2P");
        if (e && e[0] == '1') {
            uint64_t tag = 1469598103934665603ULL;  // FNV-1a of the NCCL id: the same on every rank of this run
            for (size_t i = 0; i < sizeof(id); ++i) tag = (tag ^ reinterpret_cast<const unsigned char*>(&id)[i]) * 1099511628211ULL;
            fdchan_.reset(new FdChannel(tag, rank_));
        }
    }
    ncclComm_t comm;
    nccl_check(nccl().CommInitRank(&comm, world_, id, rank_), "ncclCommInitRank");
    comm_ = comm;
    ensure_buf(4096);
    cuda_check(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking), "cudaStreamCreate(copy)");
    for (int i = 0; i < 2; ++i) {
        cuda_check(cudaEventCreateWithFlags(&received_[i], cudaEventDisableTiming), "cudaEventCreate");
        cuda_check(cudaEventCreateWithFlags(&copied_[i], cudaEventDisableTiming), "cudaEventCreate");
        cuda_check(cudaEventCreateWithFlags(&packed_[i], cudaEventDisableTiming), "cudaEventCreate");
    }
    cuda_check(cudaEventCreateWithFlags(&ready_, cudaEventDisableTiming), "cudaEventCreate");
}

Dist::~Dist() {
    if (copy_stream_) cudaStreamSynchronize(copy_stream_);
    if (comm_) nccl().CommDestroy(static_cast<ncclComm_t>(comm_));
    if (d_buf_) cudaFree(d_buf_);
    for (int i = 0; i < 2; ++i) {
        if (received_[i]) cudaEventDestroy(received_[i]);
        if (copied_[i]) cudaEventDestroy(copied_[i]);
        if (packed_[i]) cudaEventDestroy(packed_[i]);
    }
    if (ready_) cudaEventDestroy(ready_);
    if (copy_stream_) cudaStreamDestroy(copy_stream_);
}

void Dist::ensure_buf(size_t n) {
    if (n <= d_buf_doubles_) return;
    if (d_buf_) {
        cudaStreamSynchronize(stream_);
        cudaFree(d_buf_);
    }
    cuda_check(cudaMalloc(&d_buf_, n * sizeof(double)), "cudaMalloc(collective staging)");
    d_buf_doubles_ = n;
}

int Dist::take_free_rank_bit() {
    for (int r = 0; r < g_; ++r)
        if ((free_mask_ >> r) & 1) {
            free_mask_ &= ~(uint64_t(1) << r);
            return r;
        }
    throw std::logic_error("take_free_rank_bit: none free");
}

void Dist::reset_rank_bits(int n_used) {
    free_mask_ = ((uint64_t(1) << g_) - 1) & ~((uint64_t(1) << n_used) - 1);
}

void Dist::allreduce_sum_vec(double* v, size_t n) {
    if (n == 0) return;
    ensure_buf(n);
    cuda_check(cudaMemcpyAsync(d_buf_, v, n * sizeof(double), cudaMemcpyHostToDevice, stream_), "H2D");
    nccl_check(nccl().AllReduce(d_buf_, d_buf_, n, ncclDouble, ncclSum, static_cast<ncclComm_t>(comm_), stream_),
               "ncclAllReduce");
    cuda_check(cudaMemcpyAsync(v, d_buf_, n * sizeof(double), cudaMemcpyDeviceToHost, stream_), "D2H");
    cuda_check(cudaStreamSynchronize(stream_), "sync");
}

double Dist::allreduce_sum(double v) {
    allreduce_sum_vec(&v, 1);
    return v;
}

unsigned long long Dist::allreduce_min_u64(unsigned long long v) {
    ensure_buf(1);
    cuda_check(cudaMemcpyAsync(d_buf_, &v, 8, cudaMemcpyHostToDevice, stream_), "H2D");
    nccl_check(nccl().AllReduce(d_buf_, d_buf_, 1, ncclUint64, ncclMin, static_cast<ncclComm_t>(comm_), stream_),
               "ncclAllReduce(min)");
    cuda_check(cudaMemcpyAsync(&v, d_buf_, 8, cudaMemcpyDeviceToHost, stream_), "D2H");
    cuda_check(cudaStreamSynchronize(stream_), "sync");
    return v;
}

void Dist::barrier() { allreduce_sum(0.0); }

void Dist::release_rank_bit(int r, bool value, double2* shard, uint64_t n_amps) {
    if (value) {
        const int partner = rank_ ^ (1 << r);
        const bool sender = (rank_ >> r) & 1;
        ncclComm_t comm = static_cast<ncclComm_t>(comm_);
        // ranks with the bit set hold the survivors: ship the whole shard to the partner with the bit clear
        const uint64_t chunk = uint64_t(1) << 26;  // doubles per call stay far below INT_MAX-sized counts
        for (uint64_t off = 0; off < 2 * n_amps; off += chunk) {
            const uint64_t cnt = std::min(chunk, 2 * n_amps - off);
            double* p = reinterpret_cast<double*>(shard) + off;
            if (sender)
                nccl_check(nccl().Send(p, cnt, ncclDouble, partner, comm, stream_), "ncclSend");
            else
                nccl_check(nccl().Recv(p, cnt, ncclDouble, partner, comm, stream_), "ncclRecv");
        }
        if (sender) cuda_check(cudaMemsetAsync(shard, 0, n_amps * sizeof(double2), stream_), "memset");
    } else if ((rank_ >> r) & 1) {
        // value 0: the ranks with the bit set may still hold residue below the is_classical tolerance (|psi|^2 <= 1e-12);
        // a free rank bit must be *exactly* zero outside value 0, because the bit is handed out again as a fresh |0> qubit
        // (the reference drops that half, simulator.hpp:123-133)
        cuda_check(cudaMemsetAsync(shard, 0, n_amps * sizeof(double2), stream_), "memset");
    }
    free_mask_ |= uint64_t(1) << r;
}

void Dist::swap_bits(int r, int b, double2* shard, int n_local_bits, double2* staging, uint64_t staging_amps,
                     uint64_t* bytes_sent) {
    const int partner = rank_ ^ (1 << r);
    const uint64_t my_bit = (rank_ >> r) & 1;
    const uint64_t send_bit = 1 - my_bit;  // rank bit 0 ships its local-bit-1 half, rank bit 1 its local-bit-0 half
    const uint64_t block = uint64_t(1) << b;  // contiguous run inside the half
    const uint64_t n_blocks = uint64_t(1) << (n_local_bits - 1 - b);
    ncclComm_t comm = static_cast<ncclComm_t>(comm_);
    if (staging_amps == 0) throw std::runtime_error("swap_bits: no staging memory");
    cuda_check(cudaEventRecord(ready_, stream_), "record(ready)");  // everything queued so far precedes the side stream
    // The half-shard that leaves and the half-shard that arrives occupy the same addresses, so arrivals land in a
    // staging slot first.  Two slots alternate: while NCCL moves piece i over NVLink on the main stream, the copy of
    // piece i-1 from its slot into place runs on a side stream, which hides the copies behind the transfers.
    if (block < (uint64_t(1) << 20) && staging_amps >= 4) {
        // Low local bit: the half is scattered in runs shorter than 16 MiB, far too many messages for NCCL (measured
        // 38 GB/s at b ~ 0 on a 128 GiB shard).  Gather each piece into a contiguous slot, exchange the slots, scatter the
        // arrival back.  Four slots (out/in x 2) so that packing piece i+1 and unpacking piece i-1 overlap the transfer.
        const uint64_t half_amps = uint64_t(1) << (n_local_bits - 1);
        const uint64_t slot = staging_amps / 4;
        uint64_t i = 0;
        for (uint64_t first = 0; first < half_amps; first += slot, ++i) {
            const uint64_t cnt = std::min(slot, half_amps - first);
            const int sl = int(i % 2);
            double2* out = staging + uint64_t(sl) * slot;
            double2* in = staging + uint64_t(2 + sl) * slot;
            // the side stream packs piece i as soon as slot `sl` has been drained by the send of piece i-2
            if (i >= 2) cuda_check(cudaStreamWaitEvent(copy_stream_, received_[sl], 0), "wait(sent)");
            else cuda_check(cudaStreamWaitEvent(copy_stream_, ready_, 0), "wait(ready)");
            k::pack_half(copy_stream_, shard, out, first, cnt, b, int(send_bit));
            cuda_check(cudaEventRecord(packed_[sl], copy_stream_), "record(packed)");
            cuda_check(cudaStreamWaitEvent(stream_, packed_[sl], 0), "wait(packed)");
            if (i >= 2) cuda_check(cudaStreamWaitEvent(stream_, copied_[sl], 0), "wait(unpacked)");
            nccl_check(nccl().GroupStart(), "ncclGroupStart");
            nccl_check(nccl().Send(out, 2 * cnt, ncclDouble, partner, comm, stream_), "ncclSend");
            nccl_check(nccl().Recv(in, 2 * cnt, ncclDouble, partner, comm, stream_), "ncclRecv");
            nccl_check(nccl().GroupEnd(), "ncclGroupEnd");
            cuda_check(cudaEventRecord(received_[sl], stream_), "record(received)");
            cuda_check(cudaStreamWaitEvent(copy_stream_, received_[sl], 0), "wait(received)");
            k::unpack_half(copy_stream_, shard, in, first, cnt, b, int(send_bit));
            cuda_check(cudaEventRecord(copied_[sl], copy_stream_), "record(unpacked)");
            if (bytes_sent) *bytes_sent += cnt * sizeof(double2);
        }
        for (int sl = 0; sl < 2 && uint64_t(sl) < i; ++sl)
            cuda_check(cudaStreamWaitEvent(stream_, copied_[sl], 0), "wait(unpacked, final)");
        return;
    }
    const uint64_t slot_amps = staging_amps >= 2 ? staging_amps / 2 : staging_amps;
    const int n_slots = staging_amps >= 2 ? 2 : 1;
    const uint64_t piece = std::min<uint64_t>(block, slot_amps);
    uint64_t i = 0;
    for (uint64_t j = 0; j < n_blocks; ++j) {
        const uint64_t start = (j << (b + 1)) | (send_bit << b);
        for (uint64_t o = 0; o < block; o += piece, ++i) {
            const uint64_t cnt = std::min(piece, block - o);
            const int slot = int(i % n_slots);
            double2* landing = staging + uint64_t(slot) * slot_amps;
            if (i >= uint64_t(n_slots)) cuda_check(cudaStreamWaitEvent(stream_, copied_[slot], 0), "wait(copied)");
            nccl_check(nccl().GroupStart(), "ncclGroupStart");
            nccl_check(nccl().Send(shard + start + o, 2 * cnt, ncclDouble, partner, comm, stream_), "ncclSend");
            nccl_check(nccl().Recv(landing, 2 * cnt, ncclDouble, partner, comm, stream_), "ncclRecv");
            nccl_check(nccl().GroupEnd(), "ncclGroupEnd");
            cuda_check(cudaEventRecord(received_[slot], stream_), "record(received)");
            cuda_check(cudaStreamWaitEvent(copy_stream_, received_[slot], 0), "wait(received)");
            cuda_check(cudaMemcpyAsync(shard + start + o, landing, cnt * sizeof(double2), cudaMemcpyDeviceToDevice,
                                       copy_stream_),
                       "staging copy");
            cuda_check(cudaEventRecord(copied_[slot], copy_stream_), "record(copied)");
            if (bytes_sent) *bytes_sent += cnt * sizeof(double2);
        }
    }
    for (int slot = 0; slot < n_slots && uint64_t(slot) < i; ++slot)
        cuda_check(cudaStreamWaitEvent(stream_, copied_[slot], 0), "wait(copied, final)");
}

void Dist::handshake(int peer) {
    // a one-element exchange on the stream: it completes on either side only when both sides have reached it
    ncclComm_t comm = static_cast<ncclComm_t>(comm_);
    ensure_buf(2);
    nccl_check(nccl().GroupStart(), "ncclGroupStart");
    nccl_check(nccl().Send(d_buf_, 1, ncclDouble, peer, comm, stream_), "ncclSend(handshake)");
    nccl_check(nccl().Recv(d_buf_ + 1, 1, ncclDouble, peer, comm, stream_), "ncclRecv(handshake)");
    nccl_check(nccl().GroupEnd(), "ncclGroupEnd");
}

bool Dist::swap_bits_p2p(int r, int b, const GrowBuffer& state, int n_local_bits, int device, const k::Ctx& ctx,
                         uint64_t* bytes_sent) {
    if (!fdchan_ || !state.uses_vmm()) return false;
    const int partner = rank_ ^ (1 << r);
    PeerLink& link = links_[partner];
    // 1. make sure each side has the other's current shard mapped: header {layout key, chunk count}, then the sizes and
    //    descriptors if the partner does not have this layout yet
    const uint64_t my_key = state.layout_key();
    std::vector<int> fds;
    std::vector<size_t> sizes;
    uint64_t header[2] = {my_key, 0};
    const bool need_send = link.sent_key != my_key;
    if (need_send) {
        if (!state.export_chunks(fds, sizes)) {
            // tell the partner we cannot do it (chunk count ~0ULL), both fall back together
            header[1] = ~0ULL;
            fdchan_->send(partner, header, sizeof(header), {});
            uint64_t theirs[2];
            std::vector<int> none;
            fdchan_->recv(partner, theirs, sizeof(theirs), none, 0);
            if (theirs[1] != ~0ULL && theirs[1] != 0) {  // drain the partner's chunk message
                std::vector<size_t> ts(theirs[1]);
                std::vector<int> tf;
                fdchan_->recv(partner, ts.data(), ts.size() * sizeof(size_t), tf, ts.size());
                for (int f : tf) ::close(f);
            }
            return false;
        }
        header[1] = fds.size();
    }
    fdchan_->send(partner, header, sizeof(header), {});
    if (need_send) {
        fdchan_->send(partner, sizes.data(), sizes.size() * sizeof(size_t), fds);
        for (int f : fds) ::close(f);
        link.sent_key = my_key;
    }
    uint64_t theirs[2];
    std::vector<int> none;
    fdchan_->recv(partner, theirs, sizeof(theirs), none, 0);
    if (theirs[1] == ~0ULL) return false;  // partner cannot export: both use the NCCL path
    uint64_t ok = 1;
    if (theirs[1] != 0) {
        std::vector<size_t> ts(theirs[1]);
        std::vector<int> tf;
        fdchan_->recv(partner, ts.data(), ts.size() * sizeof(size_t), tf, ts.size());
        if (!link.map) link.map.reset(new PeerMapping());
        try {
            link.map->map(device, tf, ts);
            link.mapped_key = theirs[0];
        } catch (const std::exception&) {
            ok = 0;  // e.g. no peer access between the two devices
            link.mapped_key = 0;
        }
        for (int f : tf) ::close(f);
    }
    if (!link.map || link.mapped_key != theirs[0]) ok = 0;
    // agree on the outcome before anything is enqueued: one side must never wait in a handshake the other skipped
    uint64_t their_ok = 0;
    fdchan_->send(partner, &ok, sizeof(ok), {});
    fdchan_->recv(partner, &their_ok, sizeof(their_ok), none, 0);
    if (!their_ok) link.sent_key = 0;        // the partner could not map my shard: resend next time
    if (!ok || !their_ok) return false;
    // 2. both sides idle on this shard -> swap -> both sides done
    const uint64_t half = uint64_t(1) << (n_local_bits - 1);
    const uint64_t my_bit = 1 - uint64_t((rank_ >> r) & 1);  // rank bit 0 trades its local-bit-1 half, and vice versa
    const uint64_t lo = (rank_ < partner) ? 0 : half / 2, cnt = (rank_ < partner) ? half / 2 : half - half / 2;
    handshake(partner);
    k::peer_swap(ctx, state.amps(), link.map->amps(), lo, cnt, b, int(my_bit));
    handshake(partner);
    if (bytes_sent) *bytes_sent += half * sizeof(double2);
    return true;
}

void Dist::swap_bits_multi(const std::vector<std::pair<int, int>>& swaps, double2* shard, int n_local_bits,
                           double2* staging, uint64_t staging_amps, uint64_t* bytes_sent) {
    const std::vector<ExchangePeer> peers = plan_exchange(rank_, swaps);
    const uint64_t P = peers.size();
    const int g = int(swaps.size());
    uint8_t pos[8];
    if (g > 8) throw std::runtime_error("swap_bits_multi: more than 8 bits");
    for (int i = 0; i < g; ++i) pos[i] = uint8_t(swaps[i].second);
    std::sort(pos, pos + g);
    const uint64_t sub_amps = uint64_t(1) << (n_local_bits - g);  // amplitudes per sub-block
    const uint64_t slot = staging_amps / 4;                        // out/in x double buffer
    if (slot == 0) throw std::runtime_error("swap_bits_multi: staging area too small");
    ncclComm_t comm = static_cast<ncclComm_t>(comm_);
    cuda_check(cudaEventRecord(ready_, stream_), "record(ready)");
    // Pairwise rounds (peer after peer) rather than one NCCL group with all peers: measured on 4 GPUs, a grouped
    // exchange with 3 peers at once moves 245 GB/s per direction, pair exchanges 440-570 GB/s.  The pieces of all rounds
    // form one pipeline: the side stream packs piece i+1 and unpacks piece i-1 while piece i is on the wire.
    uint64_t i = 0;
    for (uint64_t p = 0; p < P; ++p) {
        for (uint64_t first = 0; first < sub_amps; first += slot, ++i) {
            const uint64_t cnt = std::min(slot, sub_amps - first);
            const int sl = int(i % 2);
            double2* out = staging + uint64_t(sl) * slot;
            double2* in = staging + uint64_t(2 + sl) * slot;
            cuda_check(cudaStreamWaitEvent(copy_stream_, i >= 2 ? received_[sl] : ready_, 0), "wait(slot free)");
            k::pack_sub(copy_stream_, shard, out, first, cnt, pos, g, peers[p].pattern);
            cuda_check(cudaEventRecord(packed_[sl], copy_stream_), "record(packed)");
            cuda_check(cudaStreamWaitEvent(stream_, packed_[sl], 0), "wait(packed)");
            if (i >= 2) cuda_check(cudaStreamWaitEvent(stream_, copied_[sl], 0), "wait(unpacked)");
            nccl_check(nccl().GroupStart(), "ncclGroupStart");
            nccl_check(nccl().Send(out, 2 * cnt, ncclDouble, peers[p].peer, comm, stream_), "ncclSend");
            nccl_check(nccl().Recv(in, 2 * cnt, ncclDouble, peers[p].peer, comm, stream_), "ncclRecv");
            nccl_check(nccl().GroupEnd(), "ncclGroupEnd");
            cuda_check(cudaEventRecord(received_[sl], stream_), "record(received)");
            cuda_check(cudaStreamWaitEvent(copy_stream_, received_[sl], 0), "wait(received)");
            k::unpack_sub(copy_stream_, shard, in, first, cnt, pos, g, peers[p].pattern);
            cuda_check(cudaEventRecord(copied_[sl], copy_stream_), "record(unpacked)");
            if (bytes_sent) *bytes_sent += cnt * sizeof(double2);
        }
    }
    for (int sl = 0; sl < 2 && uint64_t(sl) < i; ++sl)
        cuda_check(cudaStreamWaitEvent(stream_, copied_[sl], 0), "wait(unpacked, final)");
}

}  // namespace pqb
