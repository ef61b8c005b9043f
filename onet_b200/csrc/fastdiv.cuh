// Integer division by a runtime constant without the ~25-instruction (32-bit) / ~100-instruction (64-bit) division sequence.
#pragma once
#include <stdint.h>

namespace onet {

// n / d for 0 <= n < 2^31 by multiply-high (the tile-index decompositions of the persistent kernels: a runtime integer division
// is ~25 dependent instructions, and the epilogue warps did four of them per tile)
struct FastDiv { uint32_t d, mul, shr; };
inline FastDiv make_fastdiv(int d) {
    FastDiv f;
    f.d = static_cast<uint32_t>(d > 0 ? d : 1);
    if (f.d == 1u) { f.mul = 0u; f.shr = 0u; return f; }
    uint32_t lg = 0;
    while ((1ull << lg) < f.d) ++lg;
    const uint32_t pw = 31u + lg;
    f.mul = static_cast<uint32_t>(((1ull << pw) + f.d - 1ull) / f.d);
    f.shr = pw - 32u;
    return f;
}
__device__ __forceinline__ int fd_div(int n, const FastDiv& f) {
    return f.d == 1u ? n : static_cast<int>(__umulhi(static_cast<uint32_t>(n), f.mul) >> f.shr);
}

}  // namespace onet
