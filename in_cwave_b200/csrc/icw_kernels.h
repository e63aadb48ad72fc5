// icw_kernels.h -- launch wrappers exported by icw_kernels.cu to the host side (icw_api.cu).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>
#include "icw_internal.h"
#include "icw_mt.h"

namespace icw {

// half-band design in the form the recurrence consumes; passed as a __grid_constant__ kernel
// parameter so every coefficient is a constant-bank operand of its DMUL (no registers spent)
struct HbCoef {
    double fb[ICW_MAX_ORD];
    double ff[ICW_MAX_ORD];
    double d0;
};

struct HbLeafState {
    double   z[2][ICW_MAX_ORD];     // [0 = I, 1 = Q][newest first]
    unsigned long long rejects[2];
    unsigned quad;
    unsigned pad;
};

cudaError_t launch_hb_exact(const HbCoef &coef, const DevChain &ch, DevStream *streams, int n_streams,
                            int64_t n_frames, const uint8_t *in, size_t in_stride, double *analytic,
                            cudaStream_t s);
// pre != NULL (noise shaping on): (value, dither) pairs [stream][frame][4] for ns_render_kernel instead of PCM
cudaError_t launch_hb_fused(const HbCoef &coef, const DevChain &ch, DevStream *streams, int n_streams,
                            int64_t n_frames, const uint8_t *in, size_t in_stride,
                            const uint32_t *mtw_l, const uint32_t *mtw_r, size_t mt_stream_stride,
                            uint8_t *out, size_t out_stride, double *tap_bus, double *tap_lr, double *pre, cudaStream_t s);
// Kahan summation only: the state sum and the output sum of the recurrence on different warps (icw_split.cu)
cudaError_t launch_hb_split(const HbCoef &coef, const DevChain &ch, DevStream *streams, int n_streams, int64_t n_frames,
                            const uint8_t *in, size_t in_stride, const uint32_t *mtw_l, const uint32_t *mtw_r,
                            size_t mt_stream_stride, uint8_t *out, size_t out_stride, double *tap_bus, double *tap_lr,
                            double *pre, int fast, cudaStream_t s);
cudaError_t launch_hb_leaf(const HbCoef &coef, int ord, bool kahan, int reject, int n_chan, int64_t n,
                           const double *x, double *out, HbLeafState *st, cudaStream_t s);
cudaError_t launch_mt_words(const uint32_t *ckpt, int n_cta, int blocks_per_cta, int64_t first_word,
                            int64_t want_lo, int64_t want_hi, uint32_t *out, int64_t tail_block, uint32_t *tail,
                            cudaStream_t s);
cudaError_t launch_chain(const DevChain &ch, DevStream *streams, int n_streams, int64_t n_frames,
                         const uint8_t *in, size_t in_stride, int from_analytic,
                         const uint32_t *mtw_l, const uint32_t *mtw_r, size_t mt_stream_stride,
                         uint8_t *out, size_t out_stride, double *tap_bus, double *tap_lr, double *pre,
                         int sm_count, cudaStream_t s);
// the pointwise chain with both dither generators regenerated inside it (icw_chainmt.cu): one stream,
// RPDF / TPDF, both generators at the same draw.  The kernel fills pl.tail / pr.tail.
bool chain_mt_supports(const DevChain &ch);
int chain_mt_max_units(int sm_count);
cudaError_t launch_chain_mt(const DevChain &ch, DevStream *streams, int64_t n_frames, const uint8_t *in, int from_analytic,
                            const MtPlan &pl, const MtPlan &pr, uint8_t *out, double *tap_bus, double *tap_lr, double *pre,
                            cudaStream_t s);
// noise shaping on: chain_kernel leaves (value, dither) pairs in `pre`; one thread per (stream, channel) quantises
cudaError_t launch_ns_render(const DevChain &ch, DevStream *streams, int n_streams, int64_t n_frames,
                             const double *pre, uint8_t *out, size_t out_stride, cudaStream_t s);
// first pair of a stretch of generator words that the reference's dsopen would reject (atomicMin into *d_first)
cudaError_t launch_mt_find_reject(const uint32_t *w, int64_t n_pairs, long long base_pair, long long *d_first, int sm_count, cudaStream_t s);
cudaError_t launch_advance(const DevChain &ch, DevStream *streams, int n_streams, int64_t n_frames,
                           int advance_quad, cudaStream_t s);
cudaError_t launch_sincos_leaf(int64_t n, const double *x, double *out, cudaStream_t s);
cudaError_t launch_phase_leaf(const DevChain &ch, uint64_t n0, int64_t n, double f, double *out, cudaStream_t s);

}  // namespace icw
