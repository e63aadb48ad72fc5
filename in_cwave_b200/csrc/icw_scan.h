// icw_scan.h -- scan-mode (time-parallel) Hilbert converter: constants and launch wrapper.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include "icw_internal.h"

namespace icw {

constexpr int SCAN_L = 256;                 // frames per chunk (one thread each, per channel); multiple of 4.  The base
                                            // length: calls are cut on multiples of SCAN_L * SCAN_CH.  Long streams use
                                            // longer chunks (scan_chunk_len): a mode of pole radius r forgets in
                                            // log(1e-18)/log(r) samples, so pass 1 can start the fast modes late
constexpr int SCAN_CH = 128;                // chunks per tile
constexpr int SCAN_NMAX = 10;               // modes per filter (conjugate pairs + the real pole)
constexpr int64_t SCAN_SEGMENT = 1 << 25;   // frames per launch group (bounds the analytic scratch)

// partial-fraction form of one half-band design, folded for the two-sample step (icw_scan.cu)
struct ModalCoef {
    int    L, pad0_;                                // frames per chunk these constants are for (a multiple of SCAN_L)
    int    join[SCAN_NMAX + 1];                     // pass 1: step at which the mode of rank r (0 = slowest) starts; [nm] = L/2
    int    pad1_;
    int    nm, baseline;
    int    real_last, pad_;                         // the last mode is the real pole of an odd-order design
    double d0;
    double p_re[SCAN_NMAX], p_im[SCAN_NMAX];        // pole
    double p2_re[SCAN_NMAX], p2_im[SCAN_NMAX];      // pole^2
    double pinv_re[SCAN_NMAX], pinv_im[SCAN_NMAX];  // 1 / pole
    double c_re[SCAN_NMAX], c_im[SCAN_NMAX];        // y += c_re*s_re + c_im*s_im  == 2 Re(r s)
    double cp_re[SCAN_NMAX], cp_im[SCAN_NMAX];      // the same for r*p
    double pl_re[SCAN_NMAX], pl_im[SCAN_NMAX];      // pole^L
    double pt_re[SCAN_NMAX], pt_im[SCAN_NMAX];      // pole^(L*CH)
    double plp_re[16][SCAN_NMAX], plp_im[16][SCAN_NMAX];   // pole^(L*(j+1)), j = 0..15: in-CTA carry scan
    double k[6][SCAN_NMAX];                         // sample-loop constants of the sign-free form: -p^2 (re, im), -2c, 2cp
};

void scan_make_coef(int filter_no, bool baseline, double d0, int L, ModalCoef &mc, std::vector<double> &pw_table);
// frames per chunk for a group of n_streams x n_frames: the longest of 256 / 1024 / 2048 that still leaves
// every SM several waves of pass-3 CTAs
int scan_chunk_len(int n_streams, int64_t n_frames, int sm_count);
size_t scan_scratch_doubles(int n_streams, int64_t n_frames, int L);
cudaError_t launch_hb_scan(const ModalCoef &mc, const DevChain &ch, DevStream *streams, int n_streams,
                           int64_t n_frames, const uint8_t *in, size_t in_stride, const double *pw,
                           double *scratch, double *analytic, cudaStream_t s, int *launches,
                           cudaEvent_t mid_end = nullptr, cudaEvent_t mid_start = nullptr);   // recorded between pass 2 and pass 3

}  // namespace icw
