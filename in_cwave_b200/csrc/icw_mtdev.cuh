// icw_mtdev.cuh -- device-side MT19937 block regeneration shared by mt_words_kernel (words to HBM)
// and chain_mt_kernel (words to shared memory, consumed in place).
// Reference: src/mersene_twister/mt_jrnd.c:99-134 (regeneration + tempering).
#pragma once
#include <cstdint>
#include "icw_internal.h"

namespace icw {

constexpr int MT_WORDS_THREADS = 256;

__device__ __forceinline__ uint32_t shr_mul(uint32_t y, int k) { return __umulhi(y, 1u << (32 - k)); }   // y >> k on the fma pipe
__device__ __forceinline__ uint32_t mt_twist(uint32_t a, uint32_t b)
{
    uint32_t mix, mag;
    // (a & 0x80000000) | (b & 0x7FFFFFFF) as one bit-select, (b & 1) * 0x9908B0DF as a real multiply
    asm("lop3.b32 %0, %1, %2, 0x7FFFFFFF, 0xD8;" : "=r"(mix) : "r"(a), "r"(b));     // c ? b : a  == (a & ~c) | (b & c)
    asm("mul.lo.u32 %0, %1, 0x9908B0DF;" : "=r"(mag) : "r"(b & 1u));
    return shr_mul(mix, 1) ^ mag;
}
__device__ __forceinline__ uint32_t mt_temper_mul(uint32_t y)      // mt_temper with the shifts as multiplies
{
    y ^= shr_mul(y, 11);
    y ^= (y * 128u) & 0x9D2C5680u;
    y ^= (y * 32768u) & 0xEFC60000u;
    y ^= shr_mul(y, 18);
    return y;
}

// one block: s0, s1, s2 = this thread's words t, t+227, t+454 of `old` on entry, of `nw` on exit
template <bool EDGE>
__device__ __forceinline__ void mt_words_block(const uint32_t *__restrict__ old, uint32_t *__restrict__ nw, int t,
                                               uint32_t &s0, uint32_t &s1, uint32_t &s2, uint32_t *__restrict__ o,
                                               int64_t w0, int64_t want_lo, int64_t want_hi)
{
    const uint32_t n0 = old[t + 397] ^ mt_twist(s0, old[t + 1]);
    const uint32_t n1 = n0 ^ mt_twist(s1, old[t + 228]);
    nw[t] = n0;
    nw[t + 227] = n1;
    if (!EDGE || (w0 + t >= want_lo && w0 + t < want_hi)) o[0] = mt_temper_mul(n0);
    if (!EDGE || (w0 + t + 227 >= want_lo && w0 + t + 227 < want_hi)) o[227] = mt_temper_mul(n1);
    s0 = n0; s1 = n1;
    if (t < 170) {
        // the block's last word pairs with the NEW word 0 (mt_jrnd.c:121)
        uint32_t nxt = old[t + 455];
        if (t == 169) nxt = old[397] ^ mt_twist(old[0], old[1]);
        const uint32_t n2 = n1 ^ mt_twist(s2, nxt);
        nw[t + 454] = n2;
        if (!EDGE || (w0 + t + 454 >= want_lo && w0 + t + 454 < want_hi)) o[454] = mt_temper_mul(n2);
        s2 = n2;
    }
}

// ---- the stream in windows -----------------------------------------------------------------------------------
// u[j+624] = u[j+397] ^ twist(u[j], u[j+1]) reaches back 227 words at the nearest, so ANY window of up to 227
// consecutive words can be made at once from what precedes it -- the windows need not be the reference's 624-word
// blocks.  With the stream laid out LINEARLY in shared memory (624 words of history, then the tile's new words) a
// thread makes word q from three loads at fixed offsets of one pointer: 3 LDS + 5 logic/multiply + 1 STS, every lane
// busy, against ~28 issue slots a word for the block-shaped version above (227 of 256 lanes, then 170).  The words
// stay UNTEMPERED in the buffer (the recurrence needs them so); whoever draws from them tempers at use.
__device__ __forceinline__ uint32_t mt_window_word(uint32_t *__restrict__ u /* -> position of the new word */)
{
    const uint32_t n = u[-227] ^ mt_twist(u[-624], u[-623]);
    u[0] = n;
    return n;
}

}  // namespace icw
