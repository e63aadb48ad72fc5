// icw_frame.cuh -- everything that happens to one frame after its analytic value is known:
// oscillator -> DSP list -> dither -> quantise/clip/peak -> PCM bytes.  Shared by chain_kernel
// (one thread per frame) and the helper warps of the fused exact kernel.
#pragma once
#include "icw_dev.cuh"

namespace icw {

// tempered words of sample i of one channel are w[i*wps .. i*wps + wps): 8- or 16-byte aligned
__device__ __forceinline__ double dsopen2(uint2 w, unsigned &redraws)
{
    bool rd;
    double v = mt_dsopen(w.x, w.y, rd);
    redraws += rd;
    return v;
}

// The words of sample i for the 2- and 4-word dither types, fetched at the top of the frame so that
// the oscillator and the DSP list run under the load's latency (Gauss reads its 24 words late).
__device__ __forceinline__ uint4 dither_fetch(int wps, const uint32_t *w, int64_t i)
{
    uint4 a = make_uint4(0u, 0u, 0u, 0u);
    if (wps == 4) {
        a = *reinterpret_cast<const uint4 *>(w + (size_t)i * 4);
    } else if (wps == 2) {
        const uint2 b = *reinterpret_cast<const uint2 *>(w + (size_t)i * 2);
        a.x = b.x; a.y = b.y;
    }
    return a;
}

// the dither value of sample i (reference src/sound_render.c:711-751); a = dither_fetch() of the sample
__device__ __forceinline__ double dither_sample(const DevRender &r, uint4 a, const uint32_t *w, int64_t i, double prev_tr,
                                                unsigned &redraws)
{
    switch (r.render_type) {
    case ICW_RENDER_RPDF:
        return div_const_finite(dsopen2(make_uint2(a.x, a.y), redraws), ICW_KC[KC_SQRT2], ICW_KC[KC_RSQRT2]);
    case ICW_RENDER_TPDF: {
        double v = dsopen2(make_uint2(a.x, a.y), redraws);
        v += dsopen2(make_uint2(a.z, a.w), redraws);
        return v * 0.5;                                         // /2.0, exact
    }
    case ICW_RENDER_STPDF:
        return (dsopen2(make_uint2(a.x, a.y), redraws) - prev_tr) * 0.5;
    case ICW_RENDER_GAUSS: {
        const uint4 *q = reinterpret_cast<const uint4 *>(w + (size_t)i * 24);
        double v = 0.0;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const uint4 a = q[j];
            const double d0 = dsopen2(make_uint2(a.x, a.y), redraws);
            v = j == 0 ? d0 : v + d0;
            v += dsopen2(make_uint2(a.z, a.w), redraws);
        }
        const double d = 2.0 * ICW_SQRT6;
        return div_const_finite(v, d, 1.0 / d);
    }
    default:
        return 0.0;
    }
}

// first draw of sample i (sloped TPDF needs the previous sample's)
__device__ __forceinline__ double first_draw(const uint32_t *w, int64_t i)
{
    bool rd;
    const uint2 a = *reinterpret_cast<const uint2 *>(w + (size_t)i * 2);
    return mt_dsopen(a.x, a.y, rd);
}

// interleaved L,R little-endian: 2 x 2 bytes as one 32-bit store, 2 x 3 bytes as three 16-bit stores
// (frame i starts at byte 4*i or 6*i of a 16-byte aligned row)
__device__ __forceinline__ void store_frame_pcm(uint8_t *p, int l, int r, int bytes)
{
    if (bytes == 2) {
        *reinterpret_cast<uint32_t *>(p) = ((uint32_t)l & 0xFFFFu) | ((uint32_t)r << 16);
    } else {
        uint16_t *q = reinterpret_cast<uint16_t *>(p);
        q[0] = (uint16_t)l;
        q[1] = (uint16_t)(((uint32_t)l >> 16) & 0xFFu) | (uint16_t)(((uint32_t)r & 0xFFu) << 8);
        q[2] = (uint16_t)((uint32_t)r >> 8);
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
// The reference's own way to draw (mtrnd_gen_dsopen, src/mersene_twister/mt_jrnd.c:218-226,245-256): words taken one after
// the other, a pair that maps to -1 thrown away and the next pair taken.  Used for the one frame of a call that met such a
// pair (the closed form "draw j = words 2j, 2j+1" ends there); `cur` = next word of the channel's buffer.
__device__ __forceinline__ double serial_dsopen(const uint32_t *w, unsigned &cur)
{
    for (;;) {
        bool rd;
        const double v = mt_dsopen(w[cur], w[cur + 1], rd);
        cur += 2;
        if (!rd) return v;
    }
}
__device__ __forceinline__ double serial_dither(const DevRender &r, const uint32_t *w, unsigned &cur, double prev_tr, double &first)
{
    switch (r.render_type) {
    case ICW_RENDER_RPDF:
        first = serial_dsopen(w, cur);
        return div_const_finite(first, ICW_KC[KC_SQRT2], ICW_KC[KC_RSQRT2]);
    case ICW_RENDER_TPDF: {
        first = serial_dsopen(w, cur);
        double v = first;
        v += serial_dsopen(w, cur);
        return v * 0.5;
    }
    case ICW_RENDER_STPDF:
        first = serial_dsopen(w, cur);
        return (first - prev_tr) * 0.5;
    case ICW_RENDER_GAUSS: {
        first = serial_dsopen(w, cur);
        double v = first;
        for (int j = 1; j < 12; ++j) v += serial_dsopen(w, cur);
        const double d = 2.0 * ICW_SQRT6;
        return div_const_finite(v, d, 1.0 / d);
    }
    default:
        first = 0.0;
        return 0.0;
    }
}

struct FrameIO {
    const uint32_t *mtw_l, *mtw_r;  // tempered MT words of this stream's call range (or NULL)
    uint8_t *dst;                   // this stream's output row
    int dst_aligned;                // row start is 4-byte aligned: frames can be stored as 16/32-bit words
    double *tap_bus, *tap_lr;       // optional test taps for this stream ([frame][27][4], [frame][2])
    double *pre;                    // noise shaping on: [frame][4] = (L value, L dither, R value, R dither) for
                                    // ns_render_kernel instead of PCM (the error feedback is serial per channel)
};

// frame i of the call; v = (L.re, L.im, R.re, R.im) on plug 0; bus = thread-private plug values.
// DITHER_AT: where the dither words of the 2- and 4-word types come from -- fetched before the DSP
// list (8 more live registers), fetched at their use, or handed in by the caller (chain_mt_kernel
// makes them in shared memory; Gauss and sloped TPDF never take that route).
enum { DITHER_EARLY = 1, DITHER_LATE = 0, DITHER_GIVEN = 2,
       DITHER_SERIAL = 3 /* a one-frame call replayed with the reference's rejection loop (chain_serial_kernel only) */ };
template <int DITHER_AT = DITHER_EARLY>
__device__ __forceinline__ void finish_frame(const DevChain &ch, DevStream &st, int64_t i, int64_t n_frames,
                                             const double v[4], double (*bus)[4], const FrameIO &io, FrameAcc &acc,
                                             OscCounter &osc, uint4 wl = make_uint4(0u, 0u, 0u, 0u),
                                             uint4 wr = make_uint4(0u, 0u, 0u, 0u))
{
    const DevRender &rq = ch.render;
    const int wps = rq.words_per_sample;
    constexpr bool EARLY_DITHER = DITHER_AT == DITHER_EARLY;
    if (EARLY_DITHER) { wl = dither_fetch(wps, io.mtw_l, i); wr = dither_fetch(wps, io.mtw_r, i); }
    double omega = norm_omega(ch, osc.at(ch, i));
    double lo, ro;
    if (ch.shape != ICW_SHAPE_GENERIC) {
        // straight-line list: the bus only has to be materialised where somebody looks at it
        double o[4];
        run_shape(ch, v, omega, o, lo, ro);
        if (io.tap_bus || i == n_frames - 1) {
            bus[0][0] = v[0]; bus[0][1] = v[1]; bus[0][2] = v[2]; bus[0][3] = v[3];
            if (ch.shape == ICW_SHAPE_SHIFT_MASTER) {
                const int k = ch.nodes[0].n_out;
                bus[k][0] = o[0]; bus[k][1] = o[1]; bus[k][2] = o[2]; bus[k][3] = o[3];
            }
        }
    } else {
        bus[0][0] = v[0]; bus[0][1] = v[1]; bus[0][2] = v[2]; bus[0][3] = v[3];
        run_graph(ch, bus, omega, lo, ro);
    }

    double dl = 0.0, dr = 0.0;
    if (wps) {
        double prev_l = 0.0, prev_r = 0.0;
        if (DITHER_AT != DITHER_GIVEN && rq.render_type == ICW_RENDER_STPDF) {
            // the previous frame's draw, recomputed from that frame's words (frame 0: carried state)
            if (i == 0) { prev_l = st.prev_rnd[0]; prev_r = st.prev_rnd[1]; }
            else { prev_l = first_draw(io.mtw_l, i - 1); prev_r = first_draw(io.mtw_r, i - 1); }
        }
        if (DITHER_AT == DITHER_SERIAL) {
            // replay of a frame that met the rejection loop: a one-frame call (i == 0), the channel's words in order
            unsigned cl = 0, cr = 0;
            double fl, fr;
            dl = serial_dither(rq, io.mtw_l, cl, st.prev_rnd[0], fl);
            dr = serial_dither(rq, io.mtw_r, cr, st.prev_rnd[1], fr);
            st.serial_used[0] = cl; st.serial_used[1] = cr;
            if (rq.render_type == ICW_RENDER_STPDF) { st.prev_rnd_next[0] = fl; st.prev_rnd_next[1] = fr; }
        } else {
            if (DITHER_AT == DITHER_LATE) { wl = dither_fetch(wps, io.mtw_l, i); wr = dither_fetch(wps, io.mtw_r, i); }
            dl = dither_sample(rq, wl, io.mtw_l, i, prev_l, acc.redraws);
            dr = dither_sample(rq, wr, io.mtw_r, i, prev_r, acc.redraws);
        }
    }
    if (io.pre) {
        double4 *q = reinterpret_cast<double4 *>(io.pre) + i;
        *q = make_double4(lo, dl, ro, dr);
    } else {
        // FP_CHECK: the checked twin, counters fes_sr_left / _right (adv_modulator.c:757-758)
        RenderOut a = ch.fp_check ? render_one_checked(rq, lo, dl, st.fp_cnt[2]) : render_one(rq, lo, dl);
        RenderOut b = ch.fp_check ? render_one_checked(rq, ro, dr, st.fp_cnt[3]) : render_one(rq, ro, dr);
        acc.clips_l += a.clipped; acc.clips_r += b.clipped;
        acc.peak_l = fmax(acc.peak_l, a.level); acc.peak_r = fmax(acc.peak_r, b.level);
        uint8_t *p = io.dst + i * ch.out_frame_bytes;
        if (io.dst_aligned) store_frame_pcm(p, a.val, b.val, rq.bytes);
        else { store_pcm(p, a.val, rq.bytes); store_pcm(p + rq.bytes, b.val, rq.bytes); }
    }
    if (io.tap_bus) {
        double *t = io.tap_bus + (size_t)i * (ICW_N_PLUGS * 4);
        for (int k = 0; k < ICW_N_PLUGS; ++k) { t[k * 4] = bus[k][0]; t[k * 4 + 1] = bus[k][1]; t[k * 4 + 2] = bus[k][2]; t[k * 4 + 3] = bus[k][3]; }
    }
    if (io.tap_lr) { io.tap_lr[(size_t)i * 2] = lo; io.tap_lr[(size_t)i * 2 + 1] = ro; }
    if (i == n_frames - 1) {
        // the context's bus after the call == the last frame's values (adv_modulator.c:634-751)
        for (int k = 0; k < ICW_N_PLUGS; ++k) { st.bus[k][0] = bus[k][0]; st.bus[k][1] = bus[k][1]; st.bus[k][2] = bus[k][2]; st.bus[k][3] = bus[k][3]; }
        if (DITHER_AT != DITHER_GIVEN && DITHER_AT != DITHER_SERIAL && rq.render_type == ICW_RENDER_STPDF) {
            // frame 0 of this call may still be reading prev_rnd in another CTA: write the shadow copy
            st.prev_rnd_next[0] = first_draw(io.mtw_l, i);
            st.prev_rnd_next[1] = first_draw(io.mtw_r, i);
        }
    }
}

// ---- lean frame path: straight-line DSP lists, plain PCM output ---------------------------------
// List shape (ICW_SHAPE_MASTER / ICW_SHAPE_SHIFT_MASTER) and dither type (none, RPDF, TPDF) are
// template parameters, the dither words arrive as values, there is no thread-private bus.
// dither value from the words of one channel-sample (reference src/sound_render.c:711-733)
template <int RT>
__device__ __forceinline__ double lean_dither(uint4 a, unsigned &redraws)
{
    if (RT == ICW_RENDER_ROUND) return 0.0;
    if (RT == ICW_RENDER_TPDF) {
        double v = dsopen2(make_uint2(a.x, a.y), redraws);
        v += dsopen2(make_uint2(a.z, a.w), redraws);
        return v * 0.5;                                         // /2.0, exact
    }
    return div_const_finite(dsopen2(make_uint2(a.x, a.y), redraws), ICW_KC[KC_SQRT2], ICW_KC[KC_RSQRT2]);
}

// one frame of a straight-line list, plain PCM out.  Operation for operation finish_frame()'s
// ch.shape != GENERIC route with DITHER_GIVEN.
// LAST: FRAME_CHECK_LAST = render, and leave the bus state behind if i == i_last; FRAME_NO_LAST = render only (the
// caller revisits the call's last frame); FRAME_LAST_ONLY = that revisit: the bus state alone, nothing rendered.
// The eight predicated stores of the check are issue slots on every frame of an issue-bound loop.
enum { FRAME_CHECK_LAST = 0, FRAME_NO_LAST = 1, FRAME_LAST_ONLY = 2 };
template <int SHAPE, int RT, int LAST = FRAME_CHECK_LAST>
__device__ __forceinline__ void lean_frame(const DevChain &ch, DevStream &st, int64_t i, int64_t i_last, const double v[4],
                                           uint4 wl, uint4 wr, uint8_t *dst, int dst_aligned, FrameAcc &acc, OscCounter &osc)
{
    const DevRender &rq = ch.render;
    const double omega = norm_omega(ch, osc.at(ch, i));
    double o[4], lo, ro;
    run_shape<SHAPE>(ch, v, omega, o, lo, ro);
    if (LAST != FRAME_LAST_ONLY) {
        const double dl = lean_dither<RT>(wl, acc.redraws);
        const double dr = lean_dither<RT>(wr, acc.redraws);
        const RenderOut a = render_one(rq, lo, dl);
        const RenderOut b = render_one(rq, ro, dr);
        acc.clips_l += a.clipped; acc.clips_r += b.clipped;
        acc.peak_l = fmax(acc.peak_l, a.level); acc.peak_r = fmax(acc.peak_r, b.level);
        uint8_t *p = dst + i * ch.out_frame_bytes;
        if (dst_aligned) store_frame_pcm(p, a.val, b.val, rq.bytes);
        else { store_pcm(p, a.val, rq.bytes); store_pcm(p + rq.bytes, b.val, rq.bytes); }
    }
    if (LAST == FRAME_LAST_ONLY || (LAST == FRAME_CHECK_LAST && i == i_last)) {
        // the context's bus after the call == the last frame's values (adv_modulator.c:634-751);
        // plugs this list never writes keep what the context held
        st.bus[0][0] = v[0]; st.bus[0][1] = v[1]; st.bus[0][2] = v[2]; st.bus[0][3] = v[3];
        if (SHAPE == ICW_SHAPE_SHIFT_MASTER || (SHAPE == ICW_SHAPE_GENERIC && ch.shape == ICW_SHAPE_SHIFT_MASTER)) {
            const int k = ch.nodes[0].n_out;
            st.bus[k][0] = o[0]; st.bus[k][1] = o[1]; st.bus[k][2] = o[2]; st.bus[k][3] = o[3];
        }
    }
}

// The commonest chain of all -- Shift(In -> X) at one |frequency| on both channels, Master(X) with
// I+Q or I-Q outputs, scaled oscillator, 24-bit PCM into a 4-byte aligned row -- with every one of
// those facts known at compile time (the host checks them: lean_fast_ok).  Same operations in the
// same order as run_shape<ICW_SHAPE_SHIFT_MASTER> + master_out + render_one; x - y is computed as
// x + (-y), which is the same IEEE operation.
#define ICW_SHAPE_SHIFT_MASTER_FAST 3
__host__ __device__ inline bool lean_fast_ok(const DevChain &ch)
{
    if (ch.shape != ICW_SHAPE_SHIFT_MASTER || !ch.is_frmod_scaled || ch.render.bytes != 3) return false;
    const DevNode &sh = ch.nodes[0], &ms = ch.nodes[ch.n_nodes - 1];
    if (!sh.l_on || !sh.r_on || sh.l_f != sh.r_f || !(sh.l_f < 1.0e8)) return false;
    const bool lt = ms.l_tout == ICW_OUT_ADD_REIM || ms.l_tout == ICW_OUT_SUB_REIM;
    const bool rt = ms.r_tout == ICW_OUT_ADD_REIM || ms.r_tout == ICW_OUT_SUB_REIM;
    return lt && rt;
}

// n_value: the reference's scaled frame counter at frame i (adv_modulator.c:611-625), kept by the caller
template <int RT, int LAST = FRAME_CHECK_LAST>
__device__ __forceinline__ void lean_frame_fast_at(const DevChain &ch, DevStream &st, int64_t i, int64_t i_last, uint64_t n_value,
                                                   const double v[4], uint4 wl, uint4 wr, uint8_t *dst, FrameAcc &acc)
{
    const DevRender &rq = ch.render;
    const DevNode &sh = ch.nodes[0], &ms = ch.nodes[1];
    const double omega = norm_omega(ch, n_value);
    // shift node: mix from +0.0, gain, one phase for both channels (:519-550)
    double d0 = (0.0 + v[0]) * sh.l_gain, d1 = (0.0 + v[1]) * sh.l_gain;
    double d2 = (0.0 + v[2]) * sh.r_gain, d3 = (0.0 + v[3]) * sh.r_gain;
    const double x = omega * sh.l_f;                            // < 2*pi * 1e8: the quotient below is exact enough
    const double q = floor(x * ICW_KC[KC_INV_TWO_PI]);
    double ph = fma(-q, ICW_KC[KC_TWO_PI], x);
    if (ph < 0.0) ph += ICW_KC[KC_TWO_PI];
    else if (ph >= ICW_KC[KC_TWO_PI]) ph -= ICW_KC[KC_TWO_PI];
    double s, c;
    sincos_2pi(ph, s, c);
    const double sl = sh.l_neg ? -s : s, sr = sh.r_neg ? -s : s;
    double o[4];
    rotate(c, sl, d0, d1, o[0], o[1]);
    rotate(c, sr, d2, d3, o[2], o[3]);
    if (LAST != FRAME_LAST_ONLY) {
        // master (:485-507)
        d0 = (0.0 + o[0]) * ms.l_gain; d1 = (0.0 + o[1]) * ms.l_gain;
        d2 = (0.0 + o[2]) * ms.r_gain; d3 = (0.0 + o[3]) * ms.r_gain;
        const double lo = div_const(d0 + (ms.l_tout == ICW_OUT_SUB_REIM ? -d1 : d1), ICW_KC[KC_SQRT2], ICW_KC[KC_RSQRT2]);
        const double ro = div_const(d2 + (ms.r_tout == ICW_OUT_SUB_REIM ? -d3 : d3), ICW_KC[KC_SQRT2], ICW_KC[KC_RSQRT2]);
        const double dl = lean_dither<RT>(wl, acc.redraws);
        const double dr = lean_dither<RT>(wr, acc.redraws);
        const RenderOut a = render_one(rq, lo, dl);
        const RenderOut b = render_one(rq, ro, dr);
        acc.clips_l += a.clipped; acc.clips_r += b.clipped;
        acc.peak_l = fmax(acc.peak_l, a.level); acc.peak_r = fmax(acc.peak_r, b.level);
        store_frame_pcm(dst + i * 6, a.val, b.val, 3);
    }
    if (LAST == FRAME_LAST_ONLY || (LAST == FRAME_CHECK_LAST && i == i_last)) {
        st.bus[0][0] = v[0]; st.bus[0][1] = v[1]; st.bus[0][2] = v[2]; st.bus[0][3] = v[3];
        const int k = sh.n_out;
        st.bus[k][0] = o[0]; st.bus[k][1] = o[1]; st.bus[k][2] = o[2]; st.bus[k][3] = o[3];
    }
}

template <int RT, int LAST = FRAME_CHECK_LAST>
__device__ __forceinline__ void lean_frame_fast(const DevChain &ch, DevStream &st, int64_t i, int64_t i_last, const double v[4],
                                                uint4 wl, uint4 wr, uint8_t *dst, FrameAcc &acc, OscCounter &osc)
{
    // oscillator, scaled counter (adv_modulator.c:611-625)
    uint64_t d = (uint64_t)(i - osc.frame);
    osc.frame = i;
    if (d >= ch.scale_sr) d %= ch.scale_sr;
    osc.value += d;
    if (osc.value >= ch.scale_sr) osc.value -= ch.scale_sr;
    lean_frame_fast_at<RT, LAST>(ch, st, i, i_last, osc.value, v, wl, wr, dst, acc);
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
