// icw_scan_dev.cuh -- device helpers shared by the scan-mode kernels (icw_scan.cu, icw_sfused.cu): complex one-pole
// steps, the sign-free fs/4 mixer form, where the modal constants live.
#pragma once
#include "icw_dev.cuh"
#include "icw_scan.h"

namespace icw {

struct Cx { double re, im; };

// The fs/4 mixer without sign logic.  A filter's inputs arrive every second frame with alternating
// sign (reference src/lpf_hilbert_quad.c:129-131: I gets +x,0,-x,0; Q gets 0,-x,0,+x), and the up-mix
// (:132-153) multiplies the two outputs that follow an input by that same sign and by 2.  Carrying
// S~ = sign * S instead of S turns   S <- p^2 S + sign*x ; out = sign * 2 * sum(c S)   into
//     out1 = sum(2cp * S~) ; S~ <- (-p^2) S~ - x ; out2 = sum(-2c * S~)
// with no sign anywhere.  Negation and doubling are exact in binary floating point and round-to-
// nearest is symmetric, so every stored value is bit-identical to the signed formulation.
//
// The constants reach the sample loops through shared memory on purpose: as plain kernel parameters
// ptxas treats them as warp-uniform, parks ~60 doubles in the 63 uniform registers, spills those
// into vector registers and pays an R2UR per use -- more instructions than the DFMAs they feed.
// Values read back from shared memory are ordinary per-thread registers.
enum { K_PR = 0, K_PI, K_CR, K_CI, K_CPR, K_CPI };
__device__ __forceinline__ void stage_constants(double (*k)[SCAN_NMAX], const ModalCoef &mc)
{
    for (int i = threadIdx.x; i < 6 * SCAN_NMAX; i += blockDim.x) k[i / SCAN_NMAX][i % SCAN_NMAX] = mc.k[i / SCAN_NMAX][i % SCAN_NMAX];
    __syncthreads();
}
// UMASK bit a set: constant group a is read as a kernel parameter (uniform register operand of the
// DFMA: two vector operands, two cycles), clear: from shared memory into a per-thread register
// (three vector operands, three cycles -- the register file delivers two 64-bit operands a cycle).
#ifndef ICW_APPLY_UMASK
#define ICW_APPLY_UMASK 0x03
#endif
#ifndef ICW_LOCAL_UMASK
#define ICW_LOCAL_UMASK 0x03
#endif
constexpr int APPLY_UMASK = ICW_APPLY_UMASK, LOCAL_UMASK = ICW_LOCAL_UMASK;
#ifndef ICW_APPLY_UNROLL
#define ICW_APPLY_UNROLL 1
#endif
constexpr int APPLY_UNROLL = ICW_APPLY_UNROLL;      // steps of the pass-3 sample loop per trip
#ifndef ICW_LOCAL_UNROLL
#define ICW_LOCAL_UNROLL 4
#endif
#ifndef ICW_SCAN_PF_BYTES
#define ICW_SCAN_PF_BYTES 256
#endif
constexpr int LOCAL_UNROLL = ICW_LOCAL_UNROLL;      // steps of the pass-1 sample loop per trip (one prefetch per trip)
// (re-reading the residue constants from shared memory at every use -- a broadcast LDS each, 80 registers
// freed, 512 threads per SM -- was measured slower: 13.7 ms against 10.3 per C2 step)
template <int UMASK>
__device__ __forceinline__ double kconst(const ModalCoef &mc, const double (*kshared)[SCAN_NMAX], int a, int m)
{
    return ((UMASK >> a) & 1) ? mc.k[a][m] : kshared[a][m];
}
// sign of a filter's input at mixer phase q (q has the filter's parity): -1 for I at 2 and Q at 1
__device__ __forceinline__ double mixer_sign(int filt, unsigned q)
{
    return (((q >> 1) ^ (unsigned)filt) & 1u) ? -1.0 : 1.0;
}
__device__ __forceinline__ Cx cx_mul(double ar, double ai, Cx b)
{
    Cx r;
    r.re = fma(ar, b.re, -(ai * b.im));
    r.im = fma(ar, b.im, ai * b.re);
    return r;
}
// s <- m*s + u (u real)
__device__ __forceinline__ void cx_step(Cx &s, double mr, double mi, double u)
{
    const double nr = fma(mr, s.re, fma(-mi, s.im, u));
    const double ni = fma(mr, s.im, mi * s.re);
    s.re = nr; s.im = ni;
}

// the same for the real pole of an odd-order design: its state has no imaginary part, ever
__device__ __forceinline__ void re_step(Cx &s, double mr, double u) { s.re = fma(mr, s.re, u); }

}  // namespace icw
