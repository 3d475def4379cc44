// Launchers of the sm_100a kernels (kernels.cu).  Host code only sees these plain functions.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace pqb {
namespace k {

// launch context: the engine's stream and its launch counter (pqb_stats.kernel_launches)
struct Ctx {
    cudaStream_t stream;
    uint64_t* launches;
};

// restriction of a launch to a slice of the state: the (ascending) local bits pos[0..n) are held at the values in val
struct Slice {
    int n = 0;
    uint8_t pos[16] = {};
    uint64_t val = 0;
};

constexpr int kMaxDense = 5;        // widest dense gate (reference: simulator.hpp:522-523 throws above 5)
constexpr int kReducePartials = 65536;  // capacity (in doubles) of the partial-sum scratch the reductions use

// One fused dense pass (reference: kernel<k>, intrin/kernel1..5.hpp).  n_bits = log2(#local amplitudes).
// tpos: ascending bit positions of the k targets (matrix bit l <-> tpos[l]); cpos: ascending control positions.
// m_host: 2^k x 2^k row-major (re,im).
void apply_dense(const Ctx& c, double2* psi, int n_bits, int k, const uint8_t* tpos, int n_ctrl, const uint8_t* cpos,
                 const double* m_host, const Slice& slice = Slice());
// Diagonal pass: psi[i] *= d[bits of i at tpos] on the control-satisfying subspace; d_host has 2^k (re,im) entries.
// whether a k = 5 pass whose lowest target / control / slice bit is `lowest_fixed_bit` runs on the FP64 tensor pipe
// (apply_dense_k5_dmma_kernel) rather than the DFMA kernel: the scheduler prices the two differently
bool dense_k5_dmma_applies(int n_bits, int lowest_fixed_bit, int n_fixed);
void apply_diagonal(const Ctx& c, double2* psi, int n_bits, int k, const uint8_t* tpos, int n_ctrl,
                    const uint8_t* cpos, const double* d_host, const Slice& slice = Slice());

// sum_{(i & mask) == val} |psi_i|^2 -> d_out[0]  (reference: get_probability, simulator.hpp:363-367)
void norm_masked(const Ctx& c, const double2* psi, uint64_t n_amps, uint64_t mask, uint64_t val, double* d_partials,
                 double* d_out);
// psi_i <- (i & mask) == val ? psi_i * scale : 0   (reference: simulator.hpp:174-185, 478-484)
// d_out (+)= sum_j Re(conj(a_j) b_j)
void dot_real(const Ctx& c, const double2* a, const double2* b, uint64_t n_amps, double* d_partials, double* d_out,
              bool accumulate);
void collapse_scale(const Ctx& c, double2* psi, uint64_t n_amps, uint64_t mask, uint64_t val, double scale);
// psi_i *= scale
void scale_all(const Ctx& c, double2* psi, uint64_t n_amps, double scale);
// d_out2[b] = min over amplitudes with |psi|^2 > tol and bit `pos` == b of the *logical* index with that bit removed
// (UINT64_MAX if none): decides is_classical / get_classical_value (reference: simulator.hpp:76-108).
// phys2log: logical bit of each physical bit, or nullptr for the identity layout.  rank_bits: this rank's physical
// high bits (already shifted), 0 on a single GPU.
void classical_probe(const Ctx& c, const double2* psi, uint64_t n_amps, int pos_phys, int pos_log, double tol,
                     const uint8_t* phys2log, int n_total_bits, uint64_t rank_bits, unsigned long long* d_out2);
// out[j] = in[insert bit `pos` = value into j]  (reference: collapse_vector shrink branch, simulator.hpp:123-133)
void compact_bit(const Ctx& c, const double2* in, double2* out, uint64_t n_out, int pos, int value);
// out[i] = in[permute(i)], in-index bit perm[b] <- out-index bit b
void permute_gather(const Ctx& c, const double2* in, double2* out, uint64_t n_amps, int n_bits, const uint8_t* perm);
// out[j] = psi[indices[j]] for a handful of indices
void gather_indices(const Ctx& c, const double2* psi, const uint64_t* d_indices, uint64_t n, double2* d_out);

// exchange two local index bits of the state in place (n_bits >= 2)
void swap_local_bits(const Ctx& c, double2* psi, int n_bits, int b0, int b1);
// Global<->local remap over peer-mapped memory, one kernel per rank (see kernels.cu).  Sub-block element j of a shard is
// shard[insert_zero_bits(j, pos) | pattern]; pos = ascending union of the exchanged local bits and the slice bits.
constexpr int kMaxExchangePeers = 7;   // 2^3 - 1: three rank bits of an 8-GPU box exchanged at once
constexpr int kFlagArrive = 0;         // sync page: [kFlagArrive + rank] and [kFlagDone + rank], written by that rank
constexpr int kFlagDone = 64;
struct ExchangeArgs {
    double2* mine;                                  // my shard
    double2* peer[kMaxExchangePeers];               // the peers' shards, mapped into this process
    uint64_t out_pattern[kMaxExchangePeers];        // my sub-block that trades places with peer p's
    uint64_t in_pattern;                            // the peers' sub-block that is mine (same bits for every peer)
    uint64_t count;                                 // amplitudes per sub-block
    uint8_t lower[kMaxExchangePeers];               // 1: this rank handles the first half of the j range of that pair
    uint8_t pos[16];
    int n_peers, n_pos;
    // cross-GPU ordering (sync = 0: the caller orders the kernels itself, used by single-process tests)
    int sync, my_rank;
    int peer_rank[kMaxExchangePeers];
    unsigned long long epoch[kMaxExchangePeers];    // per-pair exchange counter (same value on both ranks of the pair)
    unsigned long long* my_flags;                   // my sync page
    unsigned long long* peer_flags[kMaxExchangePeers];  // the peers' sync pages, mapped
    unsigned int* block_counter;                    // zero-initialised device word
    int* host_error;                                // mapped pinned host word: non-zero after a spin timed out
};
void peer_exchange(cudaStream_t stream, const ExchangeArgs& a, int sm_count);
// remap support: gather / scatter the half of the shard whose local bit `pos` equals `value` (piece [first, first+count))
void pack_half(cudaStream_t stream, const double2* shard, double2* packed, uint64_t first, uint64_t count, int pos, int value);
void unpack_half(cudaStream_t stream, double2* shard, const double2* packed, uint64_t first, uint64_t count, int pos, int value);
// same for the sub-block whose local bits at pos[0..n_pos) (ascending) spell `pattern` (multi-bit exchange)
void pack_sub(cudaStream_t stream, const double2* shard, double2* packed, uint64_t first, uint64_t count, const uint8_t* pos,
              int n_pos, uint64_t pattern);
void unpack_sub(cudaStream_t stream, double2* shard, const double2* packed, uint64_t first, uint64_t count,
                const uint8_t* pos, int n_pos, uint64_t pattern);

// Measurement search support (reference: the serial inverse-CDF scan, simulator.hpp:156-158): sums of |psi|^2 over
// bins.  The subspace is fixed_val on the positions in ins_pos that are neither bin bits; bin b covers the amplitudes
// whose bits at bin_pos spell b.  ins_pos = ascending union of the fixed positions and bin positions.
// d_bins receives 2^m doubles (each bin reduced in a fixed order -> run-to-run deterministic).
void bin_sums(const Ctx& c, const double2* psi, int n_bits, int n_ins, const uint8_t* ins_pos, uint64_t fixed_val, int m,
              const uint8_t* bin_pos, double* d_partials, double* d_bins);

// emulate_math (reference: simulator.hpp:224-290): out[pi(i)] += in[i]; `out` must be zeroed by the caller.
enum MathMode { MATH_ADD = 0, MATH_ADD_MOD = 1, MATH_MUL_MOD = 2, MATH_TABLE = 3 };
struct MathDesc {
    int mode;
    long long a, N;
    const unsigned long long* d_table;  // MATH_TABLE only
    int n_regs;
    int reg_off[17];      // register r uses reg_pos[reg_off[r] .. reg_off[r+1])
    uint8_t reg_pos[64];  // physical bit position of each register bit, least-significant first
    uint64_t ctrl_mask;
};
void emulate_math(const Ctx& c, const double2* in, double2* out, uint64_t n_amps, const MathDesc& d);

// The same operation as a gather through the inverse map, for registers of up to kMathGatherBits bits in total: the host
// tabulates the map v -> f(v) on the concatenated register value once and inverts it into a CSR list (sources of every
// destination value, ascending), and every output amplitude sums its sources with plain loads and one plain store — no
// memset of the destination, no atomics, 32 B/amplitude.  Register bits are described as runs of consecutive positions.
constexpr int kMathGatherBits = 20;
struct MathGatherDesc {
    const uint32_t* d_inv_off;  // 2^bits + 1 offsets into d_inv_src
    const uint32_t* d_inv_src;  // 2^bits source values, grouped by destination value
    uint64_t ctrl_mask, reg_mask;
    int n_segs;
    struct Seg {
        uint8_t pos, len, shift;  // bits [pos, pos+len) of the index <-> bits [shift, shift+len) of the register value
    } seg[64];
};
void emulate_math_gather(const Ctx& c, const double2* in, double2* out, uint64_t n_amps, const MathGatherDesc& d);

// Gather through the inverse map in CLOSED FORM, for registers too wide to tabulate.  Per register (nb bits, value y):
//   x + a            : the single source (y - a) mod 2^nb;
//   (x + a) % N      : y >= N has no source; y < N has x0 = (y - a) mod N and, from outside the gate's domain, x0 + k N < 2^nb;
//   (x * a) % N      : the same with x0 = (y * a^-1) mod N  (needs gcd(a, N) = 1; Barrett reduction with m = floor(2^64 / N)).
// The host checks 0 <= a < N <= 2^nb (and the gcd) and otherwise falls back to the atomic scatter.  Sources of several
// registers combine as a product set; an output amplitude is the sum over it (one term whenever the state stays inside the
// gate's domain), written once: 16 B read + 16 B write per amplitude, no memset, no atomics.
struct MathInverseDesc {
    int mode;
    unsigned long long a_sub;   // ADD: a mod 2^64 (subtracted mod 2^nb);  ADD_MOD: a;  MUL_MOD: a^-1 mod N
    unsigned long long N, barrett;  // barrett = floor(2^64 / N) (MUL_MOD, N < 2^32)
    uint64_t ctrl_mask, reg_mask;
    int n_regs;
    int seg_off[17];            // register r is made of seg[seg_off[r] .. seg_off[r+1])
    uint8_t nb[16];             // bits of register r
    struct Seg {
        uint8_t pos, len, shift;  // bits [pos, pos+len) of the index <-> bits [shift, shift+len) of the register's value
    } seg[64];
};
void emulate_math_inverse(const Ctx& c, const double2* in, double2* out, uint64_t n_amps, const MathInverseDesc& d);

// Pauli strings in physical-bit form: (P psi)[j] = phase * (-1)^{popcount(s & zmask)} psi[s], s = j ^ xmask, with
// phase = coefficient * i^{nY} (reference: apply_term, simulator.hpp:538-550).
struct PauliTerm {
    uint64_t xmask, zmask;
    double cre, cim;
};
// sum_t Re<psi| c_t P_t |psi> for terms sharing one xmask (n_terms <= 64 per call); adds the result to d_acc[0].
void pauli_expectation_group(const Ctx& c, const double2* psi, int n_bits, uint64_t xmask, const PauliTerm* terms,
                             int n_terms, double* d_partials, double* d_acc);
// out[j] = scale * sum_t c_t (P_t in)[j]; optionally acc[j] += out[j] where (j & ctrl_mask) == ctrl_mask and
// d_norm[0] = sum over those j of |out[j]|^2 (one Taylor order of emulate_time_evolution, simulator.hpp:408-428).
// d_terms must be sorted by xmask.  rank_bits/sign handling of global bits is the caller's business.
void pauli_apply(const Ctx& c, const double2* in, double2* out, uint64_t n_amps, const PauliTerm* d_terms, int n_terms,
                 double scale_re, double scale_im, double2* acc, uint64_t ctrl_mask, double* d_partials, double* d_norm);
// Tiled form of the same operators (reference: apply_qubit_operator / get_expectation_value / the Taylor loop of
// emulate_time_evolution, simulator.hpp:292-350,386-438).  A tile is the set of 2^T amplitudes that differ only in the T
// "tile bits"; a CTA stages one tile in shared memory (one coalesced read of `in`), and every term whose X/Y support lies
// inside the tile bits finds its partner amplitude psi[j ^ xmask] in shared memory instead of in HBM.  The host covers the
// X-supports of an operator with a few tile-bit sets (one launch each), so a Taylor order costs a few sweeps instead of one
// gather stream per distinct xmask.
constexpr int kTileBits = 11;       // 2^11 amplitudes = 32 KB of shared memory per CTA
constexpr int kTileTerms = 64;      // terms per launch
// Everything the kernel needs is expressed in tile coordinates (bit i of a tile coordinate = index bit tile_pos[i]), so the
// per-amplitude work is 32-bit: a term's sign (-1)^{popcount(s & zmask)} on the source index s = j ^ xmask factors into a
// per-tile factor from the z bits outside the tile (s and j agree there; applied to the coefficient once per tile, in
// shared memory), a constant (-1)^{popcount(xl & zl)} folded into the coefficient by the host, and a per-amplitude factor
// (-1)^{popcount(t & zl)} on the OUTPUT tile coordinate t.  Diagonal terms (xmask = 0) are split three ways: z entirely
// inside the tile -> one host-built table w_in[tile coordinate] shared by all tiles; z entirely outside -> one scalar per
// tile; the rest are treated like any term.
struct PauliTileArgs {
    double2 coef[kTileTerms];       // c_t * i^{nY}
    uint64_t z_out[kTileTerms];     // z bits outside the tile (index positions)
    uint32_t xl[kTileTerms];        // xmask in tile coordinates (0 for a diagonal term)
    uint32_t zl[kTileTerms];        // z bits inside the tile, in tile coordinates
    uint8_t general[kTileTerms];    // 0: real coefficient and no z bit inside the tile (2 DFMA per amplitude, no sign work)
    // The terms are ordered by where their partner amplitude is found: [0, n_reg) flip only the tile-coordinate bits that
    // number a thread's own elements (the partner is in the thread's registers), the rest read it from the shared tile
    // (diagonal terms with z bits on both sides of the tile boundary are the case xl = 0).  The constant factor
    // (-1)^{popcount(xl & zl)} of the source-index sign is folded into coef by the host.
    int n_terms, n_reg;
    double2 coef_outside[kTileTerms];  // diagonal terms with z entirely outside the tile
    uint64_t z_outside[kTileTerms];
    int n_outside;
    const double2* w_in;            // sum of the diagonal terms with z inside the tile, per tile coordinate (or nullptr)
    int w_real;                     // 1: every entry of w_in is real
    int T, T_lo;                    // tile bits; the lowest T_lo of them are the index bits 0..T_lo-1
    uint8_t tile_pos[16];           // ascending
    uint64_t n_tiles;
    // what to do with s_j = sum_t c_t (P_t in)_j:
    int first;                      // 1: no partial sum in `u` yet;  0: u_j already holds the sum of earlier launches
    int final;                      // 1: u_j <- scale * (partial + s_j), then the optional accumulation below;  0: u_j <- partial + s_j
    int expectation;                // 1: only sum_j Re(conj(in_j) s_j) is wanted (u, acc unused)
    double sre, sim;                // scale
    uint64_t cmask;                 // acc_j += u_j and norm += |u_j|^2 where (j & cmask) == cmask (final only, acc != nullptr)
};
// One launch runs up to kTileSets sets back to back ("fused"): when their tile bits all lie below block_bits, the state is
// walked block by block and the partial sums one set hands to the next stay in L2 (see pauli_tile_kernel).
constexpr int kTileSets = 3;
constexpr int kPauliSyncWords = (1 << 16) + 1;  // unsigned words of d_sync: a ticket counter + one counter per block
struct PauliFusedArgs {
    PauliTileArgs set[kTileSets];
    int n_sets;
    int lag;                   // blocks between a block's set s and its set s+1 in the ticket order
    uint64_t tiles_per_block;  // per set
    uint64_t n_blocks;
};
// d_partials receives one double per CTA when `expectation` or (final && acc); grid size is returned.  The sets must agree in
// T, n_tiles and `expectation`; first / final describe the whole chain (set 0 may continue earlier launches, the last set may
// finalise).  d_sync: kPauliSyncWords unsigned words of device scratch.
int pauli_tile_pass(const Ctx& c, const double2* in, double2* u, double2* acc, const PauliTileArgs* sets, int n_sets,
                    int block_bits, double* d_partials, unsigned* d_sync);
void reduce_partials(const Ctx& c, const double* d_partials, int n, double* d_out, bool accumulate);
// out[j] (+)= sum_t c_t (P_t in)[j] by per-term gathers from global memory (terms whose X support no tile covers)
void pauli_gather_accumulate(const Ctx& c, const double2* in, double2* out, uint64_t n_amps, const PauliTerm* d_terms,
                             int n_terms, bool first);
// psi_i *= (re,im) where (i & ctrl_mask) == ctrl_mask
void scale_masked(const Ctx& c, double2* psi, uint64_t n_amps, uint64_t ctrl_mask, double re, double im);

// benchmark helpers
void init_random(const Ctx& c, double2* psi, uint64_t n_amps, uint64_t seed, uint64_t index_offset);
void flush_l2(const Ctx& c, double* buf, uint64_t n_doubles);
double measure_fp64_tflops(const Ctx& c, int sm_count);

}  // namespace k
}  // namespace pqb
