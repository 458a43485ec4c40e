// Sortable keys shared by the NMS and sort kernels.
#pragma once
#include <cstdint>

namespace spb200 {

constexpr int kNmsCounters = 8;   // ints per image: [0] survivors, [1] undecided after round 0, [2..4] round totals

// monotone map float -> uint32 (larger float <=> larger key), and back
__device__ __forceinline__ unsigned sortable_bits(float v) {
    const unsigned u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_sortable_bits(unsigned s) {
    return __uint_as_float((s & 0x80000000u) ? (s & 0x7fffffffu) : ~s);
}
// survivor key: confidence in the high word, inverted pixel index in the low word, so that a descending
// sort gives descending confidence with ties by ascending pixel index (the oracle's tie rule)
__device__ __forceinline__ unsigned long long survivor_key(unsigned conf_key, unsigned pix) {
    return ((unsigned long long)conf_key << 32) | (unsigned)(~pix);
}

}  // namespace spb200
