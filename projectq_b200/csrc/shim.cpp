// pybind11 shim: the `_cppsim.Simulator` method surface (reference: projectq/backends/_sim/_cppsim.cpp:43-67) over the
// C ABI of include/pqb200.h.  Marshalling only — every call goes through a pqb_* entry point.
#include <pybind11/complex.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <complex>
#include <cstring>
#include <string>
#include <vector>

#include "pqb200.h"

namespace py = pybind11;
using cplx = std::complex<double>;
using carray = py::array_t<cplx, py::array::c_style | py::array::forcecast>;

namespace {

[[noreturn]] void raise(int status, const char* msg) {
    switch (status) {
        case PQB_ERR_VALUE: throw py::value_error(msg);
        case PQB_ERR_MEMORY: PyErr_SetString(PyExc_MemoryError, msg); throw py::error_already_set();
        default: throw std::runtime_error(msg);  // RuntimeError, like the reference's std::runtime_error
    }
}

struct Terms {
    std::vector<size_t> offsets;
    std::vector<uint32_t> qidx;
    std::vector<char> pauli;
    std::vector<double> coeff;
    pqb_terms view;
};

// terms: list of (list of (index, 'X'|'Y'|'Z'), coefficient)
void pack_terms(const py::object& terms, bool complex_coeff, Terms& out) {
    out.offsets.push_back(0);
    for (auto item : terms) {
        py::sequence tup = py::reinterpret_borrow<py::sequence>(item);
        if (py::len(tup) != 2) throw py::type_error("each term must be a (term, coefficient) pair");
        for (auto op : tup[0]) {
            py::sequence o = py::reinterpret_borrow<py::sequence>(op);
            if (py::len(o) != 2) throw py::type_error("each factor must be an (index, 'X'|'Y'|'Z') pair");
            out.qidx.push_back(o[0].cast<uint32_t>());
            std::string s = o[1].cast<std::string>();
            if (s.size() != 1) throw py::type_error("Pauli action must be a single character");
            out.pauli.push_back(s[0]);
        }
        out.offsets.push_back(out.qidx.size());
        py::object c = tup[1];
        if (complex_coeff) {
            const cplx v = c.cast<cplx>();
            out.coeff.push_back(v.real());
            out.coeff.push_back(v.imag());
        } else {
            // the reference binds std::vector<std::pair<Term, double>>: a Python complex is a TypeError
            if (PyComplex_Check(c.ptr())) throw py::type_error("coefficients must be real (got a complex number)");
            out.coeff.push_back(c.cast<double>());
        }
    }
    out.view.n_terms = out.offsets.size() - 1;
    out.view.term_offsets = out.offsets.data();
    out.view.qubit_index = out.qidx.data();
    out.view.pauli = out.pauli.data();
    out.view.coefficients = out.coeff.data();
}

struct Regs {
    std::vector<uint32_t> flat, sizes;
};

Regs pack_regs(const std::vector<std::vector<uint32_t>>& quregs) {
    Regs r;
    for (auto& q : quregs) {
        r.sizes.push_back(uint32_t(q.size()));
        r.flat.insert(r.flat.end(), q.begin(), q.end());
    }
    return r;
}

class Simulator {
public:
    Simulator(uint32_t seed, int device, int fusion_max_qubits, int rank, int world_size, py::object nccl_unique_id,
              int reserve_qubits) {
        pqb_opts o;
        std::memset(&o, 0, sizeof(o));
        o.device = device;
        o.fusion_max_qubits = fusion_max_qubits;
        o.rank = rank;
        o.world_size = world_size;
        o.reserve_qubits = reserve_qubits;
        std::string uid;
        if (!nccl_unique_id.is_none()) {
            uid = nccl_unique_id.cast<py::bytes>();
            if (uid.size() != 128) throw py::value_error("nccl_unique_id must be 128 bytes");
            o.nccl_unique_id = uid.data();
        }
        const int st = pqb_create(seed, &o, &sim_);
        if (st != PQB_OK) raise(st, pqb_last_error(nullptr));
    }
    ~Simulator() { pqb_destroy(sim_); }
    Simulator(const Simulator&) = delete;

    void check(int st) const {
        if (st != PQB_OK) raise(st, pqb_last_error(sim_));
    }

    void allocate_qubit(uint32_t id) { check(pqb_allocate_qubit(sim_, id)); }
    void deallocate_qubit(uint32_t id) { check(pqb_deallocate_qubit(sim_, id)); }
    bool get_classical_value(uint32_t id, double tol) {
        int v = 0;
        check(pqb_get_classical_value(sim_, id, tol, &v));
        return v != 0;
    }
    bool is_classical(uint32_t id, double tol) {
        int v = 0;
        check(pqb_is_classical(sim_, id, tol, &v));
        return v != 0;
    }
    std::vector<bool> measure_qubits(const std::vector<uint32_t>& ids) {
        std::vector<uint8_t> bits(ids.size());
        check(pqb_measure_qubits(sim_, ids.data(), ids.size(), bits.data()));
        return std::vector<bool>(bits.begin(), bits.end());
    }
    void apply_controlled_gate(const carray& m, const std::vector<uint32_t>& ids, const std::vector<uint32_t>& ctrl) {
        const size_t d = size_t(1) << ids.size();
        if (ids.size() > 5) throw py::value_error("Gates with more than 5 qubits are not supported!");
        if (m.ndim() != 2 || size_t(m.shape(0)) != d || size_t(m.shape(1)) != d)
            throw py::value_error("apply_controlled_gate(): the matrix must be 2^k x 2^k for k target qubits");
        check(pqb_apply_controlled_gate(sim_, reinterpret_cast<const double*>(m.data()), ids.data(), ids.size(),
                                        ctrl.data(), ctrl.size()));
    }
    void emulate_math(const py::function& f, const std::vector<std::vector<uint32_t>>& quregs,
                      const std::vector<uint32_t>& ctrl) {
        // The reference calls f once per basis state (_cppsim.cpp:33-41).  f is a pure function of the register values,
        // so it is evaluated once per distinct register-value tuple into a table and the permutation runs on the GPU.
        Regs r = pack_regs(quregs);
        size_t bits = 0;
        for (auto s : r.sizes) bits += s;
        if (bits > 26) throw py::value_error("emulate_math(): registers wider than 26 bits in total are not supported");
        std::vector<uint64_t> table(size_t(1) << bits);
        for (size_t v = 0; v < table.size(); ++v) {
            py::list args;
            size_t sh = 0;
            for (auto s : r.sizes) {
                args.append(py::int_((v >> sh) & ((size_t(1) << s) - 1)));
                sh += s;
            }
            py::object res = f(args);
            uint64_t packed = 0;
            sh = 0;
            size_t i = 0;
            for (auto item : res) {
                if (i >= r.sizes.size()) break;
                // low bits of the (possibly negative) result, two's complement (simulator.hpp:255-259)
                const long long y = py::reinterpret_borrow<py::object>(item).attr("__and__")(py::int_((1LL << r.sizes[i]) - 1))
                                        .cast<long long>();
                packed |= uint64_t(y) << sh;
                sh += r.sizes[i];
                ++i;
            }
            if (i != r.sizes.size()) throw py::value_error("emulate_math(): the function must return one value per register");
            table[v] = packed;
        }
        check(pqb_emulate_math_table(sim_, table.data(), table.size(), r.flat.data(), r.sizes.data(), r.sizes.size(),
                                     ctrl.data(), ctrl.size()));
    }
    void emulate_math_addConstant(long long a, const std::vector<std::vector<uint32_t>>& quregs,
                                  const std::vector<uint32_t>& ctrl) {
        Regs r = pack_regs(quregs);
        check(pqb_emulate_math_add_constant(sim_, a, r.flat.data(), r.sizes.data(), r.sizes.size(), ctrl.data(), ctrl.size()));
    }
    void emulate_math_addConstantModN(long long a, long long N, const std::vector<std::vector<uint32_t>>& quregs,
                                      const std::vector<uint32_t>& ctrl) {
        Regs r = pack_regs(quregs);
        check(pqb_emulate_math_add_constant_mod_n(sim_, a, N, r.flat.data(), r.sizes.data(), r.sizes.size(), ctrl.data(),
                                                  ctrl.size()));
    }
    void emulate_math_multiplyByConstantModN(long long a, long long N, const std::vector<std::vector<uint32_t>>& quregs,
                                             const std::vector<uint32_t>& ctrl) {
        Regs r = pack_regs(quregs);
        check(pqb_emulate_math_multiply_by_constant_mod_n(sim_, a, N, r.flat.data(), r.sizes.data(), r.sizes.size(),
                                                          ctrl.data(), ctrl.size()));
    }
    double get_expectation_value(const py::object& terms, const std::vector<uint32_t>& ids) {
        Terms t;
        pack_terms(terms, false, t);
        double out = 0.0;
        check(pqb_get_expectation_value(sim_, &t.view, ids.data(), ids.size(), &out));
        return out;
    }
    void apply_qubit_operator(const py::object& terms, const std::vector<uint32_t>& ids) {
        Terms t;
        pack_terms(terms, true, t);
        check(pqb_apply_qubit_operator(sim_, &t.view, ids.data(), ids.size()));
    }
    void emulate_time_evolution(const py::object& terms, double time, const std::vector<uint32_t>& ids,
                                const std::vector<uint32_t>& ctrl) {
        Terms t;
        pack_terms(terms, false, t);
        check(pqb_emulate_time_evolution(sim_, &t.view, time, ids.data(), ids.size(), ctrl.data(), ctrl.size()));
    }
    double get_probability(const std::vector<bool>& bits, const std::vector<uint32_t>& ids) {
        std::vector<uint8_t> b(bits.begin(), bits.end());
        b.resize(ids.size(), 0);
        double out = 0.0;
        check(pqb_get_probability(sim_, b.data(), ids.data(), ids.size(), &out));
        return out;
    }
    cplx get_amplitude(const std::vector<bool>& bits, const std::vector<uint32_t>& ids) {
        std::vector<uint8_t> b(bits.begin(), bits.end());
        b.resize(ids.size(), 0);
        double out[2] = {0.0, 0.0};
        check(pqb_get_amplitude(sim_, b.data(), ids.data(), ids.size(), out));
        return {out[0], out[1]};
    }
    void set_wavefunction(const carray& wf, const std::vector<uint32_t>& ordering) {
        check(pqb_set_wavefunction(sim_, reinterpret_cast<const double*>(wf.data()), size_t(wf.size()), ordering.data(),
                                   ordering.size()));
    }
    void collapse_wavefunction(const std::vector<uint32_t>& ids, const std::vector<bool>& values) {
        std::vector<uint8_t> v(values.begin(), values.end());
        check(pqb_collapse_wavefunction(sim_, ids.data(), ids.size(), v.data(), v.size()));
    }
    void run() {
        int st;
        {
            py::gil_scoped_release nogil;
            st = pqb_run(sim_);
        }
        check(st);
    }
    py::tuple cheat() {
        size_t n = 0;
        check(pqb_num_qubits(sim_, &n));
        std::vector<uint32_t> ids(n), pos(n);
        size_t got = 0;
        check(pqb_cheat_map(sim_, ids.data(), pos.data(), n, &got));
        py::dict map;
        for (size_t i = 0; i < got && i < n; ++i) map[py::int_(ids[i])] = py::int_(pos[i]);
        py::array_t<cplx> state(py::ssize_t(size_t(1) << n));
        check(pqb_cheat_state(sim_, reinterpret_cast<double*>(state.mutable_data()), size_t(1) << n));
        return py::make_tuple(map, state);
    }

    // ---- additions ----
    py::array_t<cplx> get_amplitudes(const py::array_t<uint64_t, py::array::c_style | py::array::forcecast>& idx) {
        py::array_t<cplx> out(idx.size());
        check(pqb_get_amplitudes(sim_, idx.data(), size_t(idx.size()), reinterpret_cast<double*>(out.mutable_data())));
        return out;
    }
    void apply_gate_stream(const py::bytes& packed, size_t n_gates, bool fuse) {
        std::string s = packed;
        int st;
        {
            py::gil_scoped_release nogil;
            st = pqb_apply_gate_stream(sim_, s.data(), s.size(), n_gates, fuse ? 1 : 0);
        }
        check(st);
    }
    // f1: a whole command list's worth of matrix gates in ONE native call.  gates = list of
    // (matrix, target ids, control ids); matrices are anything NumPy can view as a complex 2^k x 2^k array, so the
    // engine hands over `gate.matrix` as it is (the reference converts every matrix with .tolist(), _simulator.py:412).
    void apply_gate_list(const py::list& gates, bool fuse) {
        std::string buf;
        buf.reserve(size_t(py::len(gates)) * 96);
        size_t n_gates = 0;
        auto put32 = [&buf](uint32_t v) { buf.append(reinterpret_cast<const char*>(&v), 4); };
        for (auto item : gates) {
            py::sequence g = py::reinterpret_borrow<py::sequence>(item);
            if (py::len(g) != 3) throw py::type_error("each gate must be a (matrix, ids, ctrl) triple");
            carray m = carray::ensure(g[0]);
            if (!m) throw py::type_error("gate matrix is not convertible to a complex array");
            const std::vector<uint32_t> ids = g[1].cast<std::vector<uint32_t>>();
            const std::vector<uint32_t> ctrl = g[2].cast<std::vector<uint32_t>>();
            if (ids.empty() || ids.size() > 5) throw py::value_error("Gates with more than 5 qubits are not supported!");
            const size_t d = size_t(1) << ids.size();
            if (m.ndim() != 2 || size_t(m.shape(0)) != d || size_t(m.shape(1)) != d)
                throw py::value_error("apply_gate_list(): the matrix must be 2^k x 2^k for k target qubits");
            if (ctrl.size() > 64) throw py::value_error("apply_gate_list(): more than 64 control qubits");
            put32(uint32_t(ids.size()));
            put32(uint32_t(ctrl.size()));
            buf.append(reinterpret_cast<const char*>(ids.data()), 4 * ids.size());
            buf.append(reinterpret_cast<const char*>(ctrl.data()), 4 * ctrl.size());
            buf.append(reinterpret_cast<const char*>(m.data()), 16 * d * d);
            ++n_gates;
        }
        if (n_gates == 0) return;
        int st;
        {
            py::gil_scoped_release nogil;
            st = pqb_apply_gate_stream(sim_, buf.data(), buf.size(), n_gates, fuse ? 1 : 0);
        }
        check(st);
    }
    void save_state(const std::string& prefix) {
        int st;
        {
            py::gil_scoped_release nogil;
            st = pqb_save_state(sim_, prefix.c_str());
        }
        check(st);
    }
    void load_state(const std::string& prefix) {
        int st;
        {
            py::gil_scoped_release nogil;
            st = pqb_load_state(sim_, prefix.c_str());
        }
        check(st);
    }
    // zero-copy: a dict in the CUDA array interface (v3) format plus the qubit layout; wrap it with any consumer of
    // __cuda_array_interface__ (torch.as_tensor, cupy.asarray, numba) through projectq_b200.backend.DeviceStateView
    py::dict state_view() {
        void* ptr = nullptr;
        uint64_t n_amps = 0;
        uint8_t layout[64] = {0};
        size_t n = 0;
        check(pqb_state_view(sim_, &ptr, &n_amps, layout, 64, &n));
        py::dict cai;
        cai["shape"] = py::make_tuple(n_amps);
        cai["typestr"] = "<c16";
        cai["data"] = py::make_tuple(reinterpret_cast<uintptr_t>(ptr), false);
        cai["version"] = 3;
        cai["strides"] = py::none();
        py::list lay;
        for (size_t p = 0; p < n; ++p) lay.append(int(layout[p]));
        py::dict out;
        out["cuda_array_interface"] = cai;
        out["layout"] = lay;
        return out;
    }
    void init_random_state(uint32_t n, uint64_t seed) { check(pqb_init_random_state(sim_, n, seed)); }
    double norm_squared() {
        double v = 0.0;
        check(pqb_norm_squared(sim_, &v));
        return v;
    }
    void synchronize() {
        int st;
        {
            py::gil_scoped_release nogil;
            st = pqb_synchronize(sim_);
        }
        check(st);
    }
    void timer_start() { check(pqb_timer_start(sim_)); }
    double timer_stop() {
        double ms = 0.0;
        int st;
        {
            py::gil_scoped_release nogil;
            st = pqb_timer_stop(sim_, &ms);
        }
        check(st);
        return ms;
    }
    py::dict stats() {
        pqb_stats s;
        {
            int st;
            {
                py::gil_scoped_release nogil;
                st = pqb_get_stats(sim_, &s);
            }
            check(st);
        }
        py::dict d;
        d["kernel_launches"] = s.kernel_launches;
        py::list passes;
        for (int k = 0; k < 6; ++k) passes.append(s.dense_passes[k]);
        d["dense_passes"] = passes;
        d["diag_passes"] = s.diag_passes;
        d["gates_ingested"] = s.gates_ingested;
        d["remaps"] = s.remaps;
        d["remap_bytes_sent"] = s.remap_bytes_sent;
        d["remap_ms"] = s.remap_ms;
        d["p2p_remaps"] = s.p2p_remaps;
        d["pipelined_remaps"] = s.pipelined_remaps;
        d["remap_qubits"] = s.remap_qubits;
        d["remap_comm_ms"] = s.remap_comm_ms;
        py::list pass_ms;
        for (int k = 0; k < 6; ++k) pass_ms.append(s.pass_ms[k]);
        d["pass_ms"] = pass_ms;
        d["diag_ms"] = s.diag_ms;
        return d;
    }
    void reset_stats() { check(pqb_reset_stats(sim_)); }
    void set_profiling(bool on) { check(pqb_set_profiling(sim_, on ? 1 : 0)); }
    void flush_l2(size_t bytes) { check(pqb_flush_l2(sim_, bytes)); }
    double bench_dense_pass(const carray& m, const std::vector<uint32_t>& positions, uint64_t ctrl_mask, int repeats) {
        double ms = 0.0;
        check(pqb_bench_dense_pass(sim_, reinterpret_cast<const double*>(m.data()), positions.data(), positions.size(),
                                   ctrl_mask, repeats, &ms));
        return ms;
    }
    void selftest_sliced_pass(const carray& m, const std::vector<uint32_t>& positions, uint64_t ctrl_mask,
                              uint64_t slice_mask) {
        check(pqb_selftest_sliced_pass(sim_, reinterpret_cast<const double*>(m.data()), positions.data(), positions.size(),
                                       ctrl_mask, slice_mask));
    }
    double measure_fp64_peak() {
        double v = 0.0;
        check(pqb_measure_fp64_peak(sim_, &v));
        return v;
    }
    double measure_copy_bandwidth(size_t bytes) {
        double v = 0.0;
        check(pqb_measure_copy_bandwidth(sim_, bytes, &v));
        return v;
    }
    size_t num_qubits() {
        size_t n = 0;
        check(pqb_num_qubits(sim_, &n));
        return n;
    }

private:
    pqb_sim* sim_ = nullptr;
};

}  // namespace

PYBIND11_MODULE(_pqb_shim, m) {
    m.doc() = "pybind11 shim over the pqb200 C ABI (B200 state-vector engine)";
    m.def("version", [] { return std::string(pqb_version()); });
    m.def("nccl_unique_id", [] {
        char id[128];
        const int st = pqb_nccl_unique_id(id);
        if (st != PQB_OK) raise(st, pqb_last_error(nullptr));
        return py::bytes(id, 128);
    });
    m.def("selftest_exchange", [](int device, int world, int n_local_bits, const std::vector<std::pair<int, int>>& pairs,
                                  uint64_t slice_mask) {
        std::vector<int32_t> flat;
        for (auto& p : pairs) {
            flat.push_back(p.first);
            flat.push_back(p.second);
        }
        uint64_t bad = 0;
        const int st = pqb_selftest_exchange(device, world, n_local_bits, flat.data(), pairs.size(), slice_mask, &bad);
        if (st != PQB_OK) raise(st, pqb_last_error(nullptr));
        return bad;
    });
    py::class_<Simulator>(m, "Simulator")
        .def(py::init<uint32_t, int, int, int, int, py::object, int>(), py::arg("seed") = 1, py::arg("device") = 0,
             py::arg("fusion_max_qubits") = 0, py::arg("rank") = 0, py::arg("world_size") = 1,
             py::arg("nccl_unique_id") = py::none(), py::arg("reserve_qubits") = 0)
        .def("allocate_qubit", &Simulator::allocate_qubit)
        .def("deallocate_qubit", &Simulator::deallocate_qubit)
        .def("get_classical_value", &Simulator::get_classical_value, py::arg("id"), py::arg("tol") = 1e-12)
        .def("is_classical", &Simulator::is_classical, py::arg("id"), py::arg("tol") = 1e-12)
        .def("measure_qubits", &Simulator::measure_qubits)
        .def("apply_controlled_gate", &Simulator::apply_controlled_gate)
        .def("emulate_math", &Simulator::emulate_math)
        .def("emulate_math_addConstant", &Simulator::emulate_math_addConstant)
        .def("emulate_math_addConstantModN", &Simulator::emulate_math_addConstantModN)
        .def("emulate_math_multiplyByConstantModN", &Simulator::emulate_math_multiplyByConstantModN)
        .def("get_expectation_value", &Simulator::get_expectation_value)
        .def("apply_qubit_operator", &Simulator::apply_qubit_operator)
        .def("emulate_time_evolution", &Simulator::emulate_time_evolution)
        .def("get_probability", &Simulator::get_probability)
        .def("get_amplitude", &Simulator::get_amplitude)
        .def("set_wavefunction", &Simulator::set_wavefunction)
        .def("collapse_wavefunction", &Simulator::collapse_wavefunction)
        .def("run", &Simulator::run)
        .def("cheat", &Simulator::cheat)
        .def("get_amplitudes", &Simulator::get_amplitudes)
        .def("apply_gate_stream", &Simulator::apply_gate_stream, py::arg("packed"), py::arg("n_gates"),
             py::arg("fuse") = true)
        .def("apply_gate_list", &Simulator::apply_gate_list, py::arg("gates"), py::arg("fuse") = true)
        .def("save_state", &Simulator::save_state)
        .def("load_state", &Simulator::load_state)
        .def("state_view", &Simulator::state_view)
        .def("init_random_state", &Simulator::init_random_state)
        .def("norm_squared", &Simulator::norm_squared)
        .def("synchronize", &Simulator::synchronize)
        .def("timer_start", &Simulator::timer_start)
        .def("timer_stop", &Simulator::timer_stop)
        .def("stats", &Simulator::stats)
        .def("reset_stats", &Simulator::reset_stats)
        .def("set_profiling", &Simulator::set_profiling)
        .def("flush_l2", &Simulator::flush_l2)
        .def("bench_dense_pass", &Simulator::bench_dense_pass)
        .def("selftest_sliced_pass", &Simulator::selftest_sliced_pass)
        .def("measure_fp64_peak", &Simulator::measure_fp64_peak)
        .def("measure_copy_bandwidth", &Simulator::measure_copy_bandwidth)
        .def("num_qubits", &Simulator::num_qubits);
}
