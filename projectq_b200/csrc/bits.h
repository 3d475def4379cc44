// Index arithmetic shared by host code and kernels.
//
// The reference enumerates the 2^(n-k) amplitude groups of a k-qubit gate with a (k+1)-deep loop nest over
// descending strides (reference: _cppkernels/nointrin/kernel2.hpp:31-58).  Here a flat group number is expanded
// into its base index by inserting zero bits at the (ascending) target/control positions, which is what lets one
// thread grid cover any placement of targets and controls.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define PQB_HD __host__ __device__ __forceinline__
#else
#define PQB_HD inline
#endif

namespace pqb {

// insert a zero bit at position p: bits >= p move up by one
PQB_HD uint64_t insert_zero_bit(uint64_t x, unsigned p) {
    const uint64_t low = x & ((uint64_t(1) << p) - 1);
    return ((x >> p) << (p + 1)) | low;
}

// insert zero bits at every position in pos[0..m) (ascending, distinct): the inverse of "compress away these bits"
PQB_HD uint64_t insert_zero_bits(uint64_t x, const uint8_t* pos, int m) {
    for (int j = 0; j < m; ++j) x = insert_zero_bit(x, pos[j]);
    return x;
}

// remove bit p: bits above p move down by one
PQB_HD uint64_t remove_bit(uint64_t x, unsigned p) {
    const uint64_t low = x & ((uint64_t(1) << p) - 1);
    return ((x >> (p + 1)) << p) | low;
}

// out bit perm[b] <- in bit b, for b in [0, n)
PQB_HD uint64_t permute_bits(uint64_t x, const uint8_t* perm, int n) {
    uint64_t r = 0;
    for (int b = 0; b < n; ++b) r |= ((x >> b) & 1) << perm[b];
    return r;
}

// spread the low m bits of v onto positions pos[0..m)
PQB_HD uint64_t deposit_bits(uint64_t v, const uint8_t* pos, int m) {
    uint64_t r = 0;
    for (int j = 0; j < m; ++j) r |= ((v >> j) & 1) << pos[j];
    return r;
}

// gather bits at positions pos[0..m) into the low m bits
PQB_HD uint64_t extract_bits(uint64_t x, const uint8_t* pos, int m) {
    uint64_t r = 0;
    for (int j = 0; j < m; ++j) r |= ((x >> pos[j]) & 1) << j;
    return r;
}

}  // namespace pqb
