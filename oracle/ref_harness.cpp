// TEST INFRASTRUCTURE ONLY (oracle / CPU baseline) — never linked into the product.
//
// A small driver around the *unmodified* reference header
//   /root/reference/projectq/backends/_sim/_cppkernels/simulator.hpp
// (included at compile time from where it lies; see oracle/Makefile).  It exists because the
// reference's Python binding turns the state into a Python list in cheat() (_cppsim.cpp:65), which is
// unusable at >= 27 qubits; here the same C++ class is driven directly.
//
// usage: refsim <circuit.bin> <fusion 0|1> [samples.bin out.bin [ops.bin results.bin]]
//   circuit.bin : "PQBC" u32 version=1, u32 n_qubits, u32 n_gates, then per gate
//                 u32 k, u32 nc, u32 targets[k], u32 ctrls[nc], f64 matrix[2*4^k] (row-major re,im)
//   samples.bin : u64 count, u64 idx[count]   -> amplitudes written to out.bin as f64 re,im pairs
//   ops.bin     : "PQBO" u32 version=1, u32 n_ops, then per op u32 kind + payload; executed after the gates and before
//                 the amplitudes are sampled, every result appended to results.bin as f64:
//                   1 time evolution : f64 t, ids, ctrl, terms                 (no result)
//                   2 expectation    : ids, terms                              -> 1 value
//                   3 probability    : ids, u32 bits[n_ids]                    -> 1 value
//                   4 measure        : ids                                     -> n_ids values (0/1)
//                   5 (x*a) mod N    : i64 a, i64 N, ids (one register), ctrl  (no result)
//                 ids/ctrl = u32 n, u32[n]; terms = u32 n_terms, per term f64 coeff, u32 len, (u32 index, u32 'X'|'Y'|'Z')[len]
// Prints one JSON line with timings (alloc, gates, ops) and the thread count.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#if defined(_OPENMP)
#include <omp.h>
#endif
#include "_cppkernels/simulator.hpp"

struct GateRec {
    std::vector<unsigned> targets, ctrls;
    Fusion::Matrix m;
};

static bool rd(FILE* f, void* p, size_t n) { return fread(p, 1, n, f) == n; }

static std::vector<unsigned> rd_ids(FILE* f) {
    uint32_t n = 0;
    if (!rd(f, &n, 4)) { fprintf(stderr, "truncated ops file\n"); exit(2); }
    std::vector<unsigned> v(n);
    if (n && !rd(f, v.data(), 4 * size_t(n))) { fprintf(stderr, "truncated ops file\n"); exit(2); }
    return v;
}

static Simulator::TermsDict rd_terms(FILE* f) {
    uint32_t nt = 0;
    if (!rd(f, &nt, 4)) { fprintf(stderr, "truncated ops file\n"); exit(2); }
    Simulator::TermsDict td(nt);
    for (auto& t : td) {
        uint32_t len = 0;
        if (!rd(f, &t.second, 8) || !rd(f, &len, 4)) { fprintf(stderr, "truncated ops file\n"); exit(2); }
        for (uint32_t i = 0; i < len; ++i) {
            uint32_t w[2];
            if (!rd(f, w, 8)) { fprintf(stderr, "truncated ops file\n"); exit(2); }
            t.first.emplace_back(unsigned(w[0]), char(w[1]));
        }
    }
    return td;
}

// the post-circuit script (see the header comment); results are appended to `res`
static void run_ops(Simulator& sim, const char* path, std::vector<double>& res) {
    FILE* f = fopen(path, "rb");
    if (!f) { perror("ops"); exit(2); }
    char magic[4];
    uint32_t ver = 0, n_ops = 0;
    if (!rd(f, magic, 4) || memcmp(magic, "PQBO", 4) || !rd(f, &ver, 4) || !rd(f, &n_ops, 4)) {
        fprintf(stderr, "bad ops header\n");
        exit(2);
    }
    for (uint32_t o = 0; o < n_ops; ++o) {
        uint32_t kind = 0;
        if (!rd(f, &kind, 4)) { fprintf(stderr, "truncated ops file\n"); exit(2); }
        if (kind == 1) {
            double t = 0.0;
            if (!rd(f, &t, 8)) exit(2);
            auto ids = rd_ids(f);
            auto ctrl = rd_ids(f);
            auto td = rd_terms(f);
            sim.emulate_time_evolution(td, t, ids, ctrl);
        } else if (kind == 2) {
            auto ids = rd_ids(f);
            auto td = rd_terms(f);
            res.push_back(sim.get_expectation_value(td, ids));
        } else if (kind == 3) {
            auto ids = rd_ids(f);
            std::vector<uint32_t> b(ids.size());
            if (!ids.empty() && !rd(f, b.data(), 4 * ids.size())) exit(2);
            std::vector<bool> bits(b.begin(), b.end());
            res.push_back(sim.get_probability(bits, ids));
        } else if (kind == 4) {
            auto ids = rd_ids(f);
            for (bool b : sim.measure_qubits_return(ids)) res.push_back(b ? 1.0 : 0.0);
        } else if (kind == 5) {
            int64_t a = 0, N = 0;
            if (!rd(f, &a, 8) || !rd(f, &N, 8)) exit(2);
            auto ids = rd_ids(f);
            auto ctrl = rd_ids(f);
            std::vector<std::vector<unsigned>> quregs{ids};
            sim.emulate_math_multiplyByConstantModN(int(a), int(N), quregs, ctrl);
        } else {
            fprintf(stderr, "unknown op kind %u\n", kind);
            exit(2);
        }
    }
    fclose(f);
}

int main(int argc, char** argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: %s circuit.bin fusion(0|1) [samples.bin out.bin]\n", argv[0]);
        return 2;
    }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror("circuit"); return 2; }
    char magic[4];
    uint32_t ver, nq, ng;
    if (!rd(f, magic, 4) || memcmp(magic, "PQBC", 4) || !rd(f, &ver, 4) || !rd(f, &nq, 4) || !rd(f, &ng, 4)) {
        fprintf(stderr, "bad circuit header\n");
        return 2;
    }
    std::vector<GateRec> gates(ng);
    for (auto& g : gates) {
        uint32_t k, nc;
        if (!rd(f, &k, 4) || !rd(f, &nc, 4)) return 2;
        g.targets.resize(k);
        g.ctrls.resize(nc);
        if (k && !rd(f, g.targets.data(), 4 * k)) return 2;
        if (nc && !rd(f, g.ctrls.data(), 4 * nc)) return 2;
        size_t d = size_t(1) << k;
        std::vector<double> buf(2 * d * d);
        if (!rd(f, buf.data(), buf.size() * 8)) return 2;
        g.m.assign(d, Fusion::Matrix::value_type(d));
        for (size_t r = 0; r < d; ++r)
            for (size_t c = 0; c < d; ++c) g.m[r][c] = {buf[2 * (r * d + c)], buf[2 * (r * d + c) + 1]};
    }
    fclose(f);
    bool fusion = atoi(argv[2]) != 0;

    using clk = std::chrono::steady_clock;
    Simulator sim(1);
    auto t0 = clk::now();
    for (unsigned q = 0; q < nq; ++q) sim.allocate_qubit(q);
    auto t1 = clk::now();
    for (auto const& g : gates) {
        sim.apply_controlled_gate(g.m, g.targets, g.ctrls);
        if (!fusion) sim.run();
    }
    sim.run();
    auto t2 = clk::now();

    int threads = 1;
#if defined(_OPENMP)
    threads = omp_get_max_threads();
#endif
    std::vector<double> results;
    if (argc >= 7) {
        run_ops(sim, argv[5], results);
        sim.run();
        FILE* o = fopen(argv[6], "wb");
        if (!o) { perror("results"); return 2; }
        if (!results.empty()) fwrite(results.data(), 8, results.size(), o);
        fclose(o);
    }
    auto t3 = clk::now();
    double ta = std::chrono::duration<double>(t1 - t0).count();
    double tg = std::chrono::duration<double>(t2 - t1).count();
    double to = std::chrono::duration<double>(t3 - t2).count();
    printf("{\"n_qubits\": %u, \"n_gates\": %u, \"fusion\": %d, \"threads\": %d, \"alloc_s\": %.6f, \"gates_s\": %.6f, "
           "\"ops_s\": %.6f}\n",
           nq, ng, int(fusion), threads, ta, tg, to);

    if (argc >= 5) {
        FILE* s = fopen(argv[3], "rb");
        if (!s) { perror("samples"); return 2; }
        uint64_t cnt;
        if (!rd(s, &cnt, 8)) return 2;
        std::vector<uint64_t> idx(cnt);
        if (cnt && !rd(s, idx.data(), 8 * cnt)) return 2;
        fclose(s);
        auto state = sim.cheat();
        auto const& vec = std::get<1>(state);
        FILE* o = fopen(argv[4], "wb");
        if (!o) { perror("out"); return 2; }
        for (auto i : idx) {
            double a[2] = {vec[i].real(), vec[i].imag()};
            fwrite(a, 8, 2, o);
        }
        fclose(o);
    }
    return 0;
}
