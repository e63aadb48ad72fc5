// icw_frame.cuh -- everything that happens to one frame after its analytic value is known:
// oscillator -> DSP list -> dither -> quantise/clip/peak -> PCM bytes.  Shared by chain_kernel
// (one thread per frame) and the helper warps of the fused exact kernel.
#pragma once
#include "icw_dev.cuh"

namespace icw {

__device__ __forceinline__ double dither_value(const DevRender &r, const uint32_t *w, double prev_tr,
                                               double &tr_out, unsigned &redraws)
{
    // reference src/sound_render.c:711-751; w = this sample's tempered words
    bool rd;
    double v;
    switch (r.render_type) {
    case ICW_RENDER_RPDF:
        v = div_const(mt_dsopen(w[0], w[1], rd), ICW_SQRT2, ICW_RSQRT2);
        redraws += rd;
        return v;
    case ICW_RENDER_TPDF:
        v = mt_dsopen(w[0], w[1], rd); redraws += rd;
        v += mt_dsopen(w[2], w[3], rd); redraws += rd;
        return v * 0.5;                                         // /2.0, exact
    case ICW_RENDER_STPDF:
        tr_out = mt_dsopen(w[0], w[1], rd); redraws += rd;
        return (tr_out - prev_tr) * 0.5;
    case ICW_RENDER_GAUSS: {
        v = mt_dsopen(w[0], w[1], rd); redraws += rd;
#pragma unroll
        for (int j = 1; j < 12; ++j) { v += mt_dsopen(w[2 * j], w[2 * j + 1], rd); redraws += rd; }
        const double d = 2.0 * ICW_SQRT6;
        return div_const(v, d, 1.0 / d);
    }
    default:
        return 0.0;
    }
}

__device__ __forceinline__ void store_pcm(uint8_t *p, int val, int bytes)
{
    p[0] = (uint8_t)val;
    p[1] = (uint8_t)(val >> 8);
    if (bytes == 3) p[2] = (uint8_t)(val >> 16);
}

struct FrameAcc {
    unsigned clips_l = 0, clips_r = 0, redraws = 0;
    double peak_l = 0.0, peak_r = 0.0;
};

struct FrameIO {
    const uint32_t *mtw_l, *mtw_r;  // tempered MT words of this stream's call range (or NULL)
    uint8_t *dst;                   // this stream's output row
    double *tap_bus, *tap_lr;       // optional test taps for this stream ([frame][27][4], [frame][2])
};

// frame i of the call; v = (L.re, L.im, R.re, R.im) on plug 0; bus = thread-private plug values
__device__ __forceinline__ void finish_frame(const DevChain &ch, DevStream &st, int64_t i, int64_t n_frames,
                                             const double v[4], double (*bus)[4], const FrameIO &io, FrameAcc &acc)
{
    const DevRender &rq = ch.render;
    const int wps = rq.words_per_sample;
    bus[0][0] = v[0]; bus[0][1] = v[1]; bus[0][2] = v[2]; bus[0][3] = v[3];
    double omega = norm_omega(ch, frame_counter(ch, st.n_frame, (uint64_t)i));
    double lo, ro;
    run_graph(ch, bus, omega, lo, ro);

    double dl = 0.0, dr = 0.0;
    if (wps) {
        uint32_t wl[24], wr[24];
        for (int j = 0; j < wps; ++j) {
            wl[j] = io.mtw_l[(size_t)i * wps + j];
            wr[j] = io.mtw_r[(size_t)i * wps + j];
        }
        double prev_l = 0.0, prev_r = 0.0, tr;
        if (rq.render_type == ICW_RENDER_STPDF) {
            // the previous frame's draw, recomputed from that frame's words (frame 0: carried state)
            bool rd;
            if (i == 0) { prev_l = st.prev_rnd[0]; prev_r = st.prev_rnd[1]; }
            else {
                prev_l = mt_dsopen(io.mtw_l[(size_t)(i - 1) * wps], io.mtw_l[(size_t)(i - 1) * wps + 1], rd);
                prev_r = mt_dsopen(io.mtw_r[(size_t)(i - 1) * wps], io.mtw_r[(size_t)(i - 1) * wps + 1], rd);
            }
        }
        dl = dither_value(rq, wl, prev_l, tr, acc.redraws);
        dr = dither_value(rq, wr, prev_r, tr, acc.redraws);
    }
    RenderOut a = render_one(rq, lo, dl);
    RenderOut b = render_one(rq, ro, dr);
    acc.clips_l += a.clipped; acc.clips_r += b.clipped;
    acc.peak_l = fmax(acc.peak_l, a.level); acc.peak_r = fmax(acc.peak_r, b.level);
    uint8_t *p = io.dst + i * ch.out_frame_bytes;
    store_pcm(p, a.val, rq.bytes);
    store_pcm(p + rq.bytes, b.val, rq.bytes);
    if (io.tap_bus) {
        double *t = io.tap_bus + (size_t)i * (ICW_N_PLUGS * 4);
        for (int k = 0; k < ICW_N_PLUGS; ++k) { t[k * 4] = bus[k][0]; t[k * 4 + 1] = bus[k][1]; t[k * 4 + 2] = bus[k][2]; t[k * 4 + 3] = bus[k][3]; }
    }
    if (io.tap_lr) { io.tap_lr[(size_t)i * 2] = lo; io.tap_lr[(size_t)i * 2 + 1] = ro; }
    if (i == n_frames - 1) {
        // the context's bus after the call == the last frame's values (adv_modulator.c:634-751)
        for (int k = 0; k < ICW_N_PLUGS; ++k) { st.bus[k][0] = bus[k][0]; st.bus[k][1] = bus[k][1]; st.bus[k][2] = bus[k][2]; st.bus[k][3] = bus[k][3]; }
        if (rq.render_type == ICW_RENDER_STPDF) {
            bool rd;
            // frame 0 of this call may still be reading prev_rnd in another CTA: write the shadow copy
            st.prev_rnd_next[0] = mt_dsopen(io.mtw_l[(size_t)i * wps], io.mtw_l[(size_t)i * wps + 1], rd);
            st.prev_rnd_next[1] = mt_dsopen(io.mtw_r[(size_t)i * wps], io.mtw_r[(size_t)i * wps + 1], rd);
        }
    }
}

// plugs nobody writes keep whatever the context held (normally 0.0)
__device__ __forceinline__ void load_bus(const DevStream &st, double (*bus)[4])
{
    for (int k = 1; k < ICW_N_PLUGS; ++k) {
        bus[k][0] = st.bus[k][0]; bus[k][1] = st.bus[k][1]; bus[k][2] = st.bus[k][2]; bus[k][3] = st.bus[k][3];
    }
}

// fold a thread's counters into its stream: reduce over the `width` adjacent lanes that share the
// stream, then one atomic per group (peak >= 0, so its bit pattern orders like an integer).
// Every lane of the warp must call this (full-mask shuffles); lanes without a stream pass NULL.
__device__ __forceinline__ void commit_acc(DevStream *stp, FrameAcc acc, int width)
{
    for (int o = width >> 1; o; o >>= 1) {
        acc.clips_l += __shfl_xor_sync(0xffffffffu, acc.clips_l, o);
        acc.clips_r += __shfl_xor_sync(0xffffffffu, acc.clips_r, o);
        acc.redraws += __shfl_xor_sync(0xffffffffu, acc.redraws, o);
        acc.peak_l = fmax(acc.peak_l, __shfl_xor_sync(0xffffffffu, acc.peak_l, o));
        acc.peak_r = fmax(acc.peak_r, __shfl_xor_sync(0xffffffffu, acc.peak_r, o));
    }
    if (stp && (threadIdx.x & (width - 1)) == 0) {
        DevStream &st = *stp;
        if (acc.clips_l) atomicAdd(&st.clips[0], acc.clips_l);
        if (acc.clips_r) atomicAdd(&st.clips[1], acc.clips_r);
        if (acc.redraws) atomicAdd(&st.mt_redraws, (unsigned long long)acc.redraws);
        atomicMax(reinterpret_cast<unsigned long long *>(&st.peak[0]), (unsigned long long)__double_as_longlong(acc.peak_l));
        atomicMax(reinterpret_cast<unsigned long long *>(&st.peak[1]), (unsigned long long)__double_as_longlong(acc.peak_r));
    }
}

}  // namespace icw
