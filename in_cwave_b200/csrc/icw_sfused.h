// icw_sfused.h -- the one-kernel scan path (icw_sfused.cu): geometry, constant tables and launch wrapper.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>
#include "icw_internal.h"
#include "icw_mt.h"
#include "icw_scan.h"

namespace icw {

constexpr int SF_LC = 16;               // frames per chunk (a multiple of 4: every chunk of a unit starts on the same mixer phase)
constexpr int SF_NI = SF_LC / 2;        // inputs of one filter per chunk (the fs/4 mixer feeds a filter every second frame)
#ifndef ICW_SF_CH
#define ICW_SF_CH 32
#endif
#ifndef ICW_SF_CTAS
#define ICW_SF_CTAS 2
#endif
constexpr int SF_CH = ICW_SF_CH;        // chunks per range
constexpr int SF_R = SF_LC * SF_CH;     // frames per range
constexpr int SF_CTAS_PER_SM = ICW_SF_CTAS;

// A chunk in BLOCK form.  With T the state of one mode right after the filter's last input (sign-free form of
// icw_scan_dev.cuh: T <- q T - x, q = -p^2) and x_0..x_7 the chunk's inputs,
//     state after input i-1:   T_i = q^i T_0 - sum_{j<i} q^(i-1-j) x_j
//     out1 before input i:     A_i = La(T_i)          La(S) = k4 S.re + k5 S.im   (2 c p)
//     out2 after input k-1:    B_k = Lb(T_k)          Lb(S) = k2 S.re + k3 S.im   (-2 c)
// are LINEAR in (T_0, x): every output is a dot product with constants -- no recurrence inside a chunk, and the
// chunk's effect on the state (E = T_8 from T_0 = 0) is needed once, not once per pass.  All powers are formed in
// long double from the double-double poles and rounded once (icw_scan.cu: sfused_make_tables).
struct SfTab {
    double q[SF_NI][SCAN_NMAX][2];      // -(q^(7-j)):  E += q[j] * x_j
    double q8[SCAN_NMAX][2];            // q^8: carry from chunk to chunk
    double qt[SCAN_NMAX][2];            // q itself (ragged last chunk of a range: stepped input by input)
    double p[SCAN_NMAX][2], pinv[SCAN_NMAX][2];     // pole and 1/pole: the state between ranges is the plain modal state
    double ca[SF_NI][SCAN_NMAX][2];     // A_i = sum_m ca[i][m][0] T_0.re + ca[i][m][1] T_0.im + sum_{j<i} ha[i-1-j] x_j (+ d0x2 x_i)
    double cb[SF_NI + 1][SCAN_NMAX][2]; // B_k likewise with hb[k-1-j]
    double ha[SF_NI], hb[SF_NI];
    double d0x2;                        // baseline summation: the direct term 2 d0 x_i; Kahan: 0 (src/hblpf.c:1056)
    int    nm, real_last;
};

struct SfGeom {
    int64_t n_frames;
    int64_t warm;                       // frames of warm-up before a unit (a multiple of SF_R)
    int     n_units, blocks_per_unit;   // dither: one unit = blocks_per_unit 624-word blocks of both generators' streams
    int64_t first_word, want_lo, want_hi, tail_block;   // as in MtPlan
    int64_t frames_per_unit;            // no dither: an even split
    int     out_aligned, pad_;
};

// what the kernel takes: one real-input stream, straight-line DSP list, no dither / RPDF / TPDF, no fades, no shaper
bool sfused_supports(const DevChain &ch, int n_streams, bool taps_or_pre);
// frames a unit runs ahead of its first frame for its filter state: |p|^warm < 1e-19 for the design's slowest pole
int64_t sfused_warm_frames(const SfTab &tb);
void sfused_make_tables(int filter_no, bool baseline, double d0, SfTab &tb);     // icw_scan.cu (host, long double)
size_t sfused_smem_bytes(int wps);
// pl / pr: both generators' plans (NULL without dither); n_cta: units to aim for when there is no dither
cudaError_t launch_scan_fused(const SfTab &tb, const DevChain &ch, DevStream *streams, int64_t n_frames, const uint8_t *in,
                              uint8_t *out, const MtPlan *pl, const MtPlan *pr, int n_cta, int64_t warm, cudaStream_t s);

}  // namespace icw
