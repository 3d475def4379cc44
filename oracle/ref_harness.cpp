// TEST INFRASTRUCTURE ONLY (oracle / CPU baseline) — never linked into the product.
//
// A small driver around the *unmodified* reference header
//   /root/reference/projectq/backends/_sim/_cppkernels/simulator.hpp
// (included at compile time from where it lies; see oracle/Makefile).  It exists because the
// reference's Python binding turns the state into a Python list in cheat() (_cppsim.cpp:65), which is
// unusable at >= 27 qubits; here the same C++ class is driven directly.
//
// usage: refsim <circuit.bin> <fusion 0|1> [samples.bin] [out.bin]
//   circuit.bin : "PQBC" u32 version=1, u32 n_qubits, u32 n_gates, then per gate
//                 u32 k, u32 nc, u32 targets[k], u32 ctrls[nc], f64 matrix[2*4^k] (row-major re,im)
//   samples.bin : u64 count, u64 idx[count]   -> amplitudes written to out.bin as f64 re,im pairs
// Prints one JSON line with timings (alloc, gates) and the thread count.
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
    double ta = std::chrono::duration<double>(t1 - t0).count();
    double tg = std::chrono::duration<double>(t2 - t1).count();
    printf("{\"n_qubits\": %u, \"n_gates\": %u, \"fusion\": %d, \"threads\": %d, \"alloc_s\": %.6f, \"gates_s\": %.6f}\n",
           nq, ng, int(fusion), threads, ta, tg);

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
