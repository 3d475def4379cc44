// extern "C" surface of the engine (include/pqb200.h): exceptions -> status codes, nothing else.
#include <algorithm>
#include <cstring>
#include <mutex>
#include <new>
#include <string>

#include <unistd.h>

#include "dist.h"
#include "engine.h"
#include "fdpass.h"
#include "bits.h"
#include "kernels.cuh"

using pqb::Engine;

struct pqb_sim {
    Engine* eng;
    std::string err;
};

namespace {
thread_local std::string g_create_error;  // message of the last failing call that has no handle to hang it on

template <class F>
int guarded(pqb_sim* s, F&& f) {
    if (!s || !s->eng) return PQB_ERR_RUNTIME;
    // the current device is per-thread state: make the engine's device current for this call and restore the caller's
    // afterwards, so that engines on different devices, or calls from other threads, never launch on the wrong one
    pqb::DeviceGuard on_device(s->eng->device());
    try {
        f(*s->eng);
        return PQB_OK;
    } catch (const pqb::ValueErr& e) {
        s->err = e.what();
        return PQB_ERR_VALUE;
    } catch (const pqb::CudaErr& e) {
        s->err = e.what();
        return PQB_ERR_CUDA;
    } catch (const pqb::RuntimeErr& e) {
        s->err = e.what();
        return PQB_ERR_RUNTIME;
    } catch (const std::bad_alloc&) {
        s->err = "out of memory";
        return PQB_ERR_MEMORY;
    } catch (const std::invalid_argument& e) {
        s->err = e.what();
        return PQB_ERR_VALUE;
    } catch (const std::exception& e) {
        s->err = e.what();
        return PQB_ERR_RUNTIME;
    }
}

pqb::TermsView view(const pqb_terms* t, bool cplx) {
    return pqb::TermsView{t->n_terms, t->term_offsets, t->qubit_index, t->pauli, t->coefficients, cplx};
}
}  // namespace

extern "C" {

const char* pqb_version(void) { return "pqb200 0.1 (sm_100a)"; }

int pqb_create(uint32_t seed, const pqb_opts* opts, pqb_sim** out) {
    if (!out) return PQB_ERR_VALUE;
    *out = nullptr;
    pqb_opts o;
    std::memset(&o, 0, sizeof(o));
    if (opts) o = *opts;
    // the constructor makes the engine's device current; give the caller's thread its own device back afterwards
    struct RestoreDevice {
        int prev = -1;
        RestoreDevice() {
            if (cudaGetDevice(&prev) != cudaSuccess) {
                cudaGetLastError();
                prev = -1;
            }
        }
        ~RestoreDevice() {
            if (prev >= 0) cudaSetDevice(prev);
        }
    } restore_device;
    try {
        pqb_sim* s = new pqb_sim{nullptr, {}};
        try {
            s->eng = new Engine(seed, o);
        } catch (...) {
            delete s;
            throw;
        }
        *out = s;
        return PQB_OK;
    } catch (const pqb::ValueErr& e) {
        g_create_error = e.what();
        return PQB_ERR_VALUE;
    } catch (const pqb::CudaErr& e) {
        g_create_error = e.what();
        return PQB_ERR_CUDA;
    } catch (const std::bad_alloc&) {
        g_create_error = "out of memory";
        return PQB_ERR_MEMORY;
    } catch (const std::exception& e) {
        g_create_error = e.what();
        return PQB_ERR_CUDA;
    }
}

void pqb_destroy(pqb_sim* sim) {
    if (!sim) return;
    if (sim->eng) {
        pqb::DeviceGuard on_device(sim->eng->device());
        delete sim->eng;
        sim->eng = nullptr;
    }
    delete sim;
}

const char* pqb_last_error(const pqb_sim* sim) { return sim ? sim->err.c_str() : g_create_error.c_str(); }

int pqb_allocate_qubit(pqb_sim* s, uint32_t id) {
    return guarded(s, [&](Engine& e) { e.allocate_qubit(id); });
}
int pqb_deallocate_qubit(pqb_sim* s, uint32_t id) {
    return guarded(s, [&](Engine& e) { e.deallocate_qubit(id); });
}
int pqb_get_classical_value(pqb_sim* s, uint32_t id, double tol, int* out) {
    return guarded(s, [&](Engine& e) { *out = e.get_classical_value(id, tol) ? 1 : 0; });
}
int pqb_is_classical(pqb_sim* s, uint32_t id, double tol, int* out) {
    return guarded(s, [&](Engine& e) { *out = e.is_classical(id, tol) ? 1 : 0; });
}
int pqb_measure_qubits(pqb_sim* s, const uint32_t* ids, size_t n, uint8_t* out) {
    return guarded(s, [&](Engine& e) { e.measure_qubits(ids, n, out); });
}
int pqb_apply_controlled_gate(pqb_sim* s, const double* m, const uint32_t* ids, size_t k, const uint32_t* ctrl,
                              size_t nc) {
    return guarded(s, [&](Engine& e) { e.apply_controlled_gate(m, ids, k, ctrl, nc); });
}
int pqb_emulate_math_table(pqb_sim* s, const uint64_t* table, size_t table_len, const uint32_t* reg_ids,
                           const uint32_t* reg_sizes, size_t n_regs, const uint32_t* ctrl, size_t nc) {
    return guarded(s, [&](Engine& e) {
        e.emulate_math(pqb::k::MATH_TABLE, 0, 0, table, table_len, reg_ids, reg_sizes, n_regs, ctrl, nc);
    });
}
int pqb_emulate_math_add_constant(pqb_sim* s, int64_t a, const uint32_t* reg_ids, const uint32_t* reg_sizes,
                                  size_t n_regs, const uint32_t* ctrl, size_t nc) {
    return guarded(s, [&](Engine& e) {
        e.emulate_math(pqb::k::MATH_ADD, a, 0, nullptr, 0, reg_ids, reg_sizes, n_regs, ctrl, nc);
    });
}
int pqb_emulate_math_add_constant_mod_n(pqb_sim* s, int64_t a, int64_t N, const uint32_t* reg_ids,
                                        const uint32_t* reg_sizes, size_t n_regs, const uint32_t* ctrl, size_t nc) {
    return guarded(s, [&](Engine& e) {
        e.emulate_math(pqb::k::MATH_ADD_MOD, a, N, nullptr, 0, reg_ids, reg_sizes, n_regs, ctrl, nc);
    });
}
int pqb_emulate_math_multiply_by_constant_mod_n(pqb_sim* s, int64_t a, int64_t N, const uint32_t* reg_ids,
                                                const uint32_t* reg_sizes, size_t n_regs, const uint32_t* ctrl,
                                                size_t nc) {
    return guarded(s, [&](Engine& e) {
        e.emulate_math(pqb::k::MATH_MUL_MOD, a, N, nullptr, 0, reg_ids, reg_sizes, n_regs, ctrl, nc);
    });
}
int pqb_get_expectation_value(pqb_sim* s, const pqb_terms* t, const uint32_t* ids, size_t n_ids, double* out) {
    return guarded(s, [&](Engine& e) { *out = e.get_expectation_value(view(t, false), ids, n_ids); });
}
int pqb_apply_qubit_operator(pqb_sim* s, const pqb_terms* t, const uint32_t* ids, size_t n_ids) {
    return guarded(s, [&](Engine& e) { e.apply_qubit_operator(view(t, true), ids, n_ids); });
}
int pqb_emulate_time_evolution(pqb_sim* s, const pqb_terms* t, double time, const uint32_t* ids, size_t n_ids,
                               const uint32_t* ctrl, size_t nc) {
    return guarded(s, [&](Engine& e) { e.emulate_time_evolution(view(t, false), time, ids, n_ids, ctrl, nc); });
}
int pqb_get_probability(pqb_sim* s, const uint8_t* bits, const uint32_t* ids, size_t n, double* out) {
    return guarded(s, [&](Engine& e) { *out = e.get_probability(bits, ids, n); });
}
int pqb_get_amplitude(pqb_sim* s, const uint8_t* bits, const uint32_t* ids, size_t n, double* out) {
    return guarded(s, [&](Engine& e) {
        const auto a = e.get_amplitude(bits, ids, n);
        out[0] = a.real();
        out[1] = a.imag();
    });
}
int pqb_set_wavefunction(pqb_sim* s, const double* wf, size_t n_amps, const uint32_t* ordering, size_t n) {
    return guarded(s, [&](Engine& e) { e.set_wavefunction(wf, n_amps, ordering, n); });
}
int pqb_collapse_wavefunction(pqb_sim* s, const uint32_t* ids, size_t n_ids, const uint8_t* values, size_t n_values) {
    return guarded(s, [&](Engine& e) { e.collapse_wavefunction(ids, n_ids, values, n_values); });
}
int pqb_run(pqb_sim* s) {
    return guarded(s, [&](Engine& e) { e.run(); });
}
int pqb_num_qubits(pqb_sim* s, size_t* out) {
    return guarded(s, [&](Engine& e) { *out = e.num_qubits(); });
}
int pqb_cheat_map(pqb_sim* s, uint32_t* ids, uint32_t* pos, size_t cap, size_t* out_n) {
    return guarded(s, [&](Engine& e) { *out_n = e.cheat_map(ids, pos, cap); });
}
int pqb_cheat_state(pqb_sim* s, double* out, size_t cap) {
    return guarded(s, [&](Engine& e) { e.cheat_state(out, cap); });
}
int pqb_get_amplitudes(pqb_sim* s, const uint64_t* idx, size_t n, double* out) {
    return guarded(s, [&](Engine& e) { e.get_amplitudes(idx, n, out); });
}
int pqb_apply_gate_stream(pqb_sim* s, const void* packed, size_t n_bytes, size_t n_gates, int fuse) {
    return guarded(s, [&](Engine& e) { e.apply_gate_stream(packed, n_bytes, n_gates, fuse != 0); });
}
int pqb_save_state(pqb_sim* s, const char* prefix) {
    return guarded(s, [&](Engine& e) { e.save_state(prefix ? prefix : ""); });
}
int pqb_load_state(pqb_sim* s, const char* prefix) {
    return guarded(s, [&](Engine& e) { e.load_state(prefix ? prefix : ""); });
}
int pqb_state_view(pqb_sim* s, void** ptr, uint64_t* n_amps, uint8_t* layout, size_t cap, size_t* n_qubits) {
    return guarded(s, [&](Engine& e) { e.state_view(ptr, n_amps, layout, cap, n_qubits); });
}
int pqb_init_random_state(pqb_sim* s, uint32_t n, uint64_t seed) {
    return guarded(s, [&](Engine& e) { e.init_random_state(n, seed); });
}
int pqb_norm_squared(pqb_sim* s, double* out) {
    return guarded(s, [&](Engine& e) { *out = e.norm_squared(); });
}
int pqb_synchronize(pqb_sim* s) {
    return guarded(s, [&](Engine& e) { e.synchronize(); });
}
int pqb_timer_start(pqb_sim* s) {
    return guarded(s, [&](Engine& e) { e.timer_start(); });
}
int pqb_timer_stop(pqb_sim* s, double* out_ms) {
    return guarded(s, [&](Engine& e) { *out_ms = e.timer_stop(); });
}
int pqb_get_stats(pqb_sim* s, pqb_stats* out) {
    return guarded(s, [&](Engine& e) { e.get_stats(out); });
}
int pqb_set_profiling(pqb_sim* s, int on) {
    return guarded(s, [&](Engine& e) { e.set_profiling(on != 0); });
}
int pqb_reset_stats(pqb_sim* s) {
    return guarded(s, [&](Engine& e) { e.reset_stats(); });
}
int pqb_flush_l2(pqb_sim* s, size_t bytes) {
    return guarded(s, [&](Engine& e) { e.flush_l2(bytes); });
}
int pqb_bench_dense_pass(pqb_sim* s, const double* m, const uint32_t* positions, size_t k, uint64_t ctrl_mask,
                         int repeats, double* out_ms) {
    return guarded(s, [&](Engine& e) { *out_ms = e.bench_dense_pass(m, positions, k, ctrl_mask, repeats); });
}
int pqb_selftest_sliced_pass(pqb_sim* s, const double* m, const uint32_t* positions, size_t k, uint64_t ctrl_mask,
                             uint64_t slice_mask) {
    return guarded(s, [&](Engine& e) { e.selftest_sliced_pass(m, positions, k, ctrl_mask, slice_mask); });
}

int pqb_selftest_exchange(int device, int world, int n_local_bits, const int32_t* pairs, size_t n_pairs, uint64_t slice_mask,
                          uint64_t* out_mismatches) {
    using pqb::k::ExchangeArgs;
    if (!out_mismatches || !pairs || n_pairs == 0 || n_pairs > 3 || world < 2 || (world & (world - 1)) || n_local_bits < 1 ||
        n_local_bits > 24)
        return PQB_ERR_VALUE;
    pqb::DeviceGuard on_device(device);
    std::vector<double2*> shard(world, nullptr);
    cudaStream_t stream = nullptr;
    int status = PQB_OK;
    try {
        const uint64_t n = uint64_t(1) << n_local_bits;
        int sm_count = 148;
        cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device);
        if (cudaStreamCreate(&stream) != cudaSuccess) throw std::runtime_error("cudaStreamCreate failed");
        std::vector<double2> host(n);
        for (int r = 0; r < world; ++r) {
            if (cudaMalloc(&shard[r], n * sizeof(double2)) != cudaSuccess) throw std::runtime_error("cudaMalloc failed");
            for (uint64_t i = 0; i < n; ++i) host[i] = make_double2(double((uint64_t(r) << n_local_bits) | i), double(r));
            cudaMemcpy(shard[r], host.data(), n * sizeof(double2), cudaMemcpyHostToDevice);
        }
        std::vector<std::pair<int, int>> swaps;
        std::vector<uint8_t> ex_bits, slice_bits;
        for (size_t i = 0; i < n_pairs; ++i) {
            swaps.emplace_back(pairs[2 * i], pairs[2 * i + 1]);
            ex_bits.push_back(uint8_t(pairs[2 * i + 1]));
        }
        std::sort(ex_bits.begin(), ex_bits.end());
        for (int b = 0; b < n_local_bits; ++b)
            if ((slice_mask >> b) & 1) slice_bits.push_back(uint8_t(b));
        if (ex_bits.size() + slice_bits.size() > 16) throw std::runtime_error("too many exchanged + slice bits");
        const int n_slices = 1 << slice_bits.size();
        for (int sl = 0; sl < n_slices; ++sl) {
            const uint64_t sval = pqb::deposit_bits(uint64_t(sl), slice_bits.data(), int(slice_bits.size()));
            for (int r = 0; r < world; ++r) {
                const std::vector<pqb::ExchangePeer> peers = pqb::plan_exchange(r, swaps);
                ExchangeArgs a;
                std::memset(&a, 0, sizeof(a));
                a.mine = shard[r];
                a.n_peers = int(peers.size());
                size_t e = 0, f = 0;
                int np = 0;
                while ((e < ex_bits.size() || f < slice_bits.size()) && np < 16) {
                    if (f >= slice_bits.size() || (e < ex_bits.size() && ex_bits[e] < slice_bits[f]))
                        a.pos[np++] = ex_bits[e++];
                    else
                        a.pos[np++] = slice_bits[f++];
                }
                a.n_pos = np;
                a.count = uint64_t(1) << (n_local_bits - np);
                uint64_t in_pattern = 0;
                for (auto& sw : swaps) in_pattern |= uint64_t((r >> sw.first) & 1) << sw.second;
                a.in_pattern = in_pattern | sval;
                for (int p = 0; p < a.n_peers; ++p) {
                    a.peer[p] = shard[peers[p].peer];
                    a.out_pattern[p] = peers[p].pattern | sval;
                    a.lower[p] = r < peers[p].peer ? 1 : 0;
                }
                a.sync = 0;  // one process, one stream: the kernels of the "ranks" simply run one after the other
                pqb::k::peer_exchange(stream, a, sm_count);
            }
        }
        if (cudaStreamSynchronize(stream) != cudaSuccess) throw std::runtime_error("exchange kernel failed");
        // where must every amplitude be now?  old (rank, idx) -> rank bit r_i and local bit b_i trade values
        uint64_t bad = 0;
        for (int r = 0; r < world; ++r) {
            cudaMemcpy(host.data(), shard[r], n * sizeof(double2), cudaMemcpyDeviceToHost);
            for (uint64_t i = 0; i < n; ++i) {
                int old_r = r;
                uint64_t old_i = i;
                for (auto& sw : swaps) {
                    const int rb = (r >> sw.first) & 1;
                    const int lb = int((i >> sw.second) & 1);
                    old_r = (old_r & ~(1 << sw.first)) | (lb << sw.first);
                    old_i = (old_i & ~(uint64_t(1) << sw.second)) | (uint64_t(rb) << sw.second);
                }
                const double want = double((uint64_t(old_r) << n_local_bits) | old_i);
                if (host[i].x != want || host[i].y != double(old_r)) ++bad;
            }
        }
        *out_mismatches = bad;
    } catch (const std::exception& e) {
        g_create_error = e.what();
        status = PQB_ERR_CUDA;
    }
    for (auto p : shard)
        if (p) cudaFree(p);
    if (stream) cudaStreamDestroy(stream);
    return status;
}

int pqb_measure_fp64_peak(pqb_sim* s, double* out) {
    return guarded(s, [&](Engine& e) { *out = e.measure_fp64_peak(); });
}
int pqb_measure_copy_bandwidth(pqb_sim* s, size_t bytes, double* out) {
    return guarded(s, [&](Engine& e) { *out = e.measure_copy_bandwidth(bytes); });
}

// ---- host-only helpers ------------------------------------------------------------------------------------------
int pqb_host_fuse_stream(const void* packed, size_t n_bytes, size_t n_gates, int max_qubits, void* out, size_t out_cap,
                         size_t* out_bytes, size_t* out_passes) {
    try {
        pqb::Fuser fuser;
        const uint8_t* p = static_cast<const uint8_t*>(packed);
        const uint8_t* end = p + n_bytes;
        for (size_t g = 0; g < n_gates; ++g) {
            if (p + 8 > end) return PQB_ERR_VALUE;
            uint32_t k, nc;
            std::memcpy(&k, p, 4);
            std::memcpy(&nc, p + 4, 4);
            p += 8;
            if (k == 0 || k > 5 || nc > 64) return PQB_ERR_VALUE;
            const size_t d = size_t(1) << k;
            if (p + 4 * (k + nc) + 16 * d * d > end) return PQB_ERR_VALUE;
            pqb::Gate gate;
            gate.targets.resize(k);
            gate.ctrls.resize(nc);
            std::memcpy(gate.targets.data(), p, 4 * k);
            std::memcpy(gate.ctrls.data(), p + 4 * k, 4 * nc);
            gate.m.resize(d * d);
            std::memcpy(gate.m.data(), p + 4 * (k + nc), 16 * d * d);
            p += 4 * (k + nc) + 16 * d * d;
            fuser.push(std::move(gate));
        }
        auto passes = fuser.drain(max_qubits <= 0 ? 5 : max_qubits, [](uint32_t id) { return uint64_t(id); });
        uint8_t* o = static_cast<uint8_t*>(out);
        size_t used = 0;
        for (auto& ps : passes) {
            const uint32_t k = uint32_t(ps.targets.size()), nc = uint32_t(ps.ctrls.size());
            const size_t d = size_t(1) << k;
            const size_t need = 8 + 4 * (k + nc) + 16 * d * d;
            if (used + need > out_cap) return PQB_ERR_MEMORY;
            std::memcpy(o + used, &k, 4);
            std::memcpy(o + used + 4, &nc, 4);
            std::memcpy(o + used + 8, ps.targets.data(), 4 * k);
            std::memcpy(o + used + 8 + 4 * k, ps.ctrls.data(), 4 * nc);
            std::memcpy(o + used + 8 + 4 * (k + nc), ps.m.data(), 16 * d * d);
            used += need;
        }
        if (out_bytes) *out_bytes = used;
        if (out_passes) *out_passes = passes.size();
        return PQB_OK;
    } catch (const std::exception& e) {
        g_create_error = e.what();
        return PQB_ERR_VALUE;
    }
}

int pqb_host_rng_stream(uint32_t seed, size_t n, double* out) {
    std::mt19937 rng(seed);
    for (size_t i = 0; i < n; ++i) {
        const double u0 = double(rng());
        const double u1 = double(rng());
        double r = (u0 + u1 * 4294967296.0) / 18446744073709551616.0;
        if (r >= 1.0) r = std::nextafter(1.0, 0.0);
        out[i] = r;
    }
    return PQB_OK;
}

int pqb_host_plan_remap(uint8_t* loc, size_t n_logical, int n_local_bits, const uint32_t* need, size_t n_need,
                        int32_t* out_pairs, size_t cap_pairs, size_t* out_n_pairs) {
    try {
        std::vector<uint8_t> l(loc, loc + n_logical);
        std::vector<uint32_t> nd(need, need + n_need);
        auto swaps = pqb::plan_remap(l, n_local_bits, nd);
        if (swaps.size() > cap_pairs) return PQB_ERR_MEMORY;
        for (size_t i = 0; i < swaps.size(); ++i) {
            out_pairs[2 * i] = swaps[i].first;
            out_pairs[2 * i + 1] = swaps[i].second;
        }
        std::memcpy(loc, l.data(), n_logical);
        *out_n_pairs = swaps.size();
        return PQB_OK;
    } catch (const std::exception& e) {
        g_create_error = e.what();
        return PQB_ERR_RUNTIME;
    }
}

int pqb_host_shard_schedule(const void* packed, size_t n_bytes, size_t n_gates, uint32_t n_qubits, uint32_t rank_bits,
                            int max_qubits, uint32_t flushes, void* out, size_t out_cap, size_t* out_bytes) {
    // Replays Engine::run_sharded without a device: qubit ids 0..n-1 at logical positions 0..n-1, the top `rank_bits`
    // positions on rank bits; the same gate stream is flushed `flushes` times.  Records, in order, every fused pass
    // (u32 k, u32 nc, u32 targets[k], u32 ctrls[nc], f64 matrix[2*4^k]) and every remap (u32 0xFFFFFFFF, u32 n_pairs, then
    // per pair u32 incoming qubit, u32 evicted qubit); a flush boundary is the record u32 0xFFFFFFFE, u32 0.
    try {
        if (rank_bits >= n_qubits) return PQB_ERR_VALUE;
        const int L = int(n_qubits - rank_bits);
        std::map<uint32_t, uint32_t> map;
        std::vector<uint8_t> loc(n_qubits);
        for (uint32_t p = 0; p < n_qubits; ++p) {
            map[p] = p;
            loc[p] = uint8_t(p < uint32_t(L) ? p : 64 + (p - L));
        }
        std::vector<pqb::Gate> gates;
        const uint8_t* q = static_cast<const uint8_t*>(packed);
        const uint8_t* end = q + n_bytes;
        for (size_t g = 0; g < n_gates; ++g) {
            if (q + 8 > end) return PQB_ERR_VALUE;
            uint32_t k, nc;
            std::memcpy(&k, q, 4);
            std::memcpy(&nc, q + 4, 4);
            q += 8;
            if (k == 0 || k > 5 || nc > 64) return PQB_ERR_VALUE;
            const size_t d = size_t(1) << k;
            if (q + 4 * (k + nc) + 16 * d * d > end) return PQB_ERR_VALUE;
            pqb::Gate gate;
            gate.targets.resize(k);
            gate.ctrls.resize(nc);
            std::memcpy(gate.targets.data(), q, 4 * k);
            std::memcpy(gate.ctrls.data(), q + 4 * k, 4 * nc);
            gate.m.resize(d * d);
            std::memcpy(gate.m.data(), q + 4 * (k + nc), 16 * d * d);
            q += 4 * (k + nc) + 16 * d * d;
            gates.push_back(std::move(gate));
        }
        uint8_t* o = static_cast<uint8_t*>(out);
        size_t used = 0;
        auto put = [&](const void* src, size_t n) {
            if (used + n > out_cap) throw std::length_error("output buffer too small");
            std::memcpy(o + used, src, n);
            used += n;
        };
        auto put32 = [&](uint32_t v) { put(&v, 4); };
        auto key = [&](uint32_t id) -> uint64_t { return loc[map.at(id)]; };
        auto blocked = [&](uint32_t id) -> bool { return loc[map.at(id)] >= 64; };
        pqb::Fuser fuser;
        for (uint32_t f = 0; f < flushes; ++f) {
            for (auto& g : gates) fuser.push(g);
            const pqb::InteractionGraph adj = pqb::interaction_graph(fuser);
            pqb::ShardPlan plan(fuser, max_qubits <= 0 ? 4 : max_qubits);
            while (true) {
                for (size_t ci : plan.take_runnable(map, loc)) {
                    const pqb::FusedPass ps = fuser.fuse_cluster(plan.cluster(ci), key);
                    const uint32_t k = uint32_t(ps.targets.size()), nc = uint32_t(ps.ctrls.size());
                    for (auto t : ps.targets)
                        if (loc[map.at(t)] >= 64) throw std::logic_error("scheduled a pass with an off-device target");
                    put32(k);
                    put32(nc);
                    put(ps.targets.data(), 4 * k);
                    put(ps.ctrls.data(), 4 * nc);
                    put(ps.m.data(), 16 * (size_t(1) << k) * (size_t(1) << k));
                }
                if (plan.finished()) break;
                pqb::RemapChoice choice = plan.choose(map, loc, adj);
                std::vector<uint8_t> before = loc;
                auto swaps = pqb::plan_remap(loc, L, choice.need, &choice.victims);
                if (swaps.empty()) throw std::logic_error("stuck without a remap");
                put32(0xFFFFFFFFu);
                put32(uint32_t(swaps.size()));
                for (auto& sw : swaps) {
                    uint32_t in = 0, ev = 0;
                    for (uint32_t p = 0; p < n_qubits; ++p) {
                        if (before[p] == 64 + sw.first) in = p;
                        if (before[p] == sw.second) ev = p;
                    }
                    put32(in);
                    put32(ev);
                }
            }
            fuser.clear();
            put32(0xFFFFFFFEu);
            put32(0);
        }
        if (out_bytes) *out_bytes = used;
        return PQB_OK;
    } catch (const std::length_error& e) {
        g_create_error = e.what();
        return PQB_ERR_MEMORY;
    } catch (const std::exception& e) {
        g_create_error = e.what();
        return PQB_ERR_RUNTIME;
    }
}

int pqb_host_fdpass_selftest(uint64_t run_tag, int rank, int world) {
    // every rank hands every other rank the read end of a pipe that holds (rank * 1000 + peer) and checks what it gets
    try {
        pqb::FdChannel ch(run_tag, rank, world);
        for (int k = 1; k < world; ++k) {
            const int peer = rank ^ k;
            if (peer >= world) continue;
            int fds[2];
            if (::pipe(fds) != 0) return PQB_ERR_RUNTIME;
            const int32_t token = rank * 1000 + peer;
            if (::write(fds[1], &token, sizeof(token)) != ssize_t(sizeof(token))) return PQB_ERR_RUNTIME;
            ::close(fds[1]);
            const int32_t me = rank;
            ch.send(peer, &me, sizeof(me), {fds[0]});
            ::close(fds[0]);
            int32_t who = -1;
            std::vector<int> got;
            ch.recv(peer, &who, sizeof(who), got, 1);
            int32_t seen = -1;
            const bool ok = who == peer && ::read(got[0], &seen, sizeof(seen)) == ssize_t(sizeof(seen)) &&
                            seen == peer * 1000 + rank;
            ::close(got[0]);
            if (!ok) {
                g_create_error = "fdpass selftest: wrong token";
                return PQB_ERR_RUNTIME;
            }
        }
        return PQB_OK;
    } catch (const std::exception& e) {
        g_create_error = e.what();
        return PQB_ERR_RUNTIME;
    }
}

int pqb_host_plan_exchange(int rank, const int32_t* pairs, size_t n_pairs, int32_t* out_peers, uint64_t* out_patterns,
                           size_t cap, size_t* out_n) {
    try {
        std::vector<std::pair<int, int>> swaps;
        for (size_t i = 0; i < n_pairs; ++i) swaps.emplace_back(pairs[2 * i], pairs[2 * i + 1]);
        auto peers = pqb::plan_exchange(rank, swaps);
        if (peers.size() > cap) return PQB_ERR_MEMORY;
        for (size_t i = 0; i < peers.size(); ++i) {
            out_peers[i] = peers[i].peer;
            out_patterns[i] = peers[i].pattern;
        }
        *out_n = peers.size();
        return PQB_OK;
    } catch (const std::exception& e) {
        g_create_error = e.what();
        return PQB_ERR_RUNTIME;
    }
}

int pqb_host_plan_pauli_tiles(const uint64_t* xmasks, const uint64_t* zmasks, size_t n_terms, int n_local_bits,
                              int32_t* out_launch_of_term, uint64_t* out_tile_masks, size_t cap_launches,
                              size_t* out_n_launches) {
    try {
        if (n_local_bits < 0 || n_local_bits > 62) return PQB_ERR_VALUE;
        std::vector<pqb::k::PauliTerm> terms(n_terms);
        for (size_t i = 0; i < n_terms; ++i) terms[i] = pqb::k::PauliTerm{xmasks[i], zmasks[i], 1.0 + double(i), 0.0};
        const pqb::PauliPlan plan = pqb::plan_pauli_tiles(terms, n_local_bits, pqb::pauli_block_bits());
        for (size_t i = 0; i < n_terms; ++i) out_launch_of_term[i] = plan.launch_of_term[i];
        if (out_n_launches) *out_n_launches = plan.launches.size();
        for (size_t l = 0; l < plan.launches.size() && l < cap_launches; ++l) out_tile_masks[l] = plan.tile_mask[l];
        return PQB_OK;
    } catch (const std::exception& e) {
        g_create_error = e.what();
        return PQB_ERR_RUNTIME;
    }
}

int pqb_nccl_unique_id(void* out128) {
    try {
        pqb::Dist::get_unique_id(out128);
        return PQB_OK;
    } catch (const std::exception& e) {
        g_create_error = e.what();
        return PQB_ERR_CUDA;
    }
}

}  // extern "C"
