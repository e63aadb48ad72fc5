// icw_sfused.h -- the one-kernel scan path (icw_sfused.cu): geometry and launch wrapper.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>
#include "icw_internal.h"
#include "icw_mt.h"
#include "icw_scan.h"

namespace icw {

constexpr int SF_LC = 36;               // frames per chunk (a multiple of 4: every chunk of a unit starts on the same mixer phase)
constexpr int SF_CH = 64;               // chunks per range
constexpr int SF_R = SF_LC * SF_CH;     // frames per range

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
int64_t sfused_warm_frames(const ModalCoef &mc);
size_t sfused_smem_bytes(int wps);
// mc must be made for chunk length SF_LC (scan_make_coef(..., SF_LC, ...)); pl / pr: both generators' plans (NULL without dither)
cudaError_t launch_scan_fused(const ModalCoef &mc, const DevChain &ch, DevStream *streams, int64_t n_frames, const uint8_t *in,
                              uint8_t *out, const MtPlan *pl, const MtPlan *pr, int n_cta, int64_t warm, cudaStream_t s);

}  // namespace icw
