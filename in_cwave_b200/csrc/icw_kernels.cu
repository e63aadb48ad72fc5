// icw_kernels.cu -- sm_100a kernels of the in_cwave chain (round-1 set).
//
//   hb_exact_kernel   real -> analytic, one thread per half-band recurrence, the reference's own
//                     operation order (bit-exact), time-serial per chain, parallel over chains
//   mt_words_kernel   MT19937 block regeneration from per-CTA checkpoints -> tempered words
//   chain_kernel      unpack -> fade -> oscillator -> DSP list -> dither + quantise -> PCM
//
// Built with -fmad=false: a*b+c is never contracted, fused operations are explicit fma().
#include <cstdio>
#include "icw_dev.cuh"
#include "icw_kernels.h"
#include "icw_hb.cuh"
#include "icw_frame.cuh"
#include "icw_mtdev.cuh"

namespace icw {

// =============================================================================================
// exact half-band recurrences (reference src/hblpf.c:894-926 baseline, :1008-1057 Kahan;
// mixer src/lpf_hilbert_quad.c:129-156)
// =============================================================================================

// chain id = ((stream * 2 + channel) * 2 + iq).  Output: analytic frames as 4 doubles
// (L.re, L.im, R.re, R.im) == the layout of ICW_FMT_CW_F64 stereo, so chain_kernel reads it back
// as complex input.
template <int ORD, bool KAHAN, bool CHECK>
__global__ void __launch_bounds__(128)
hb_exact_kernel(const __grid_constant__ HbCoef coef, const __grid_constant__ DevChain ch,
                DevStream *__restrict__ streams, int n_streams, int64_t n_frames,
                const uint8_t *__restrict__ in, size_t in_stride,
                double *__restrict__ analytic /* [stream][frame][4] */)
{
    int cid = blockIdx.x * blockDim.x + threadIdx.x;
    if (cid >= n_streams * 4) return;
    const int iq = cid & 1, chan = (cid >> 1) & 1, stream = cid >> 2;
    DevStream &st = streams[stream];

    double z[ORD];
#pragma unroll
    for (int i = 0; i < ORD; ++i) z[i] = st.hb[chan][iq][i];
    unsigned long long rejects = st.hb_rejects[chan][iq];
    const unsigned q0 = st.quad[chan];
    const int64_t pos0 = st.pos;
    // mono: the right channel filters the (already faded) left value (xwave_reader.c:988-998)
    const uint8_t *src = in + (size_t)stream * in_stride + (ch.n_channels > 1 ? chan : 0) * ch.chan_bytes;
    double *dst = analytic + (size_t)stream * (size_t)n_frames * 4 + chan * 2;
    const bool fading = (ch.n_fade_in | ch.n_fade_out) != 0;

    hb_run<ORD, KAHAN, CHECK>(z, rejects, coef, ch.reject_flag, n_frames,
        [&](int64_t i) {
            double x = unpack_real(ch.fmt, src + i * ch.frame_bytes);
            if (fading) {
                double g = fade_gain(ch, pos0 + i);
                if (g >= 0.0) x *= g;
            }
            return mix_down(iq, (q0 + (unsigned)i) & 3u, x);
        },
        [&](int64_t i, double y) {
            int slot;
            double v = mix_up(iq, (q0 + (unsigned)i) & 3u, y, slot);
            dst[i * 4 + slot] = v;
        }, st.fp_cnt[chan]);                                    // fes_hilb_left / _right (xwave_reader.c:980,998)
#pragma unroll
    for (int i = 0; i < ORD; ++i) st.hb[chan][iq][i] = z[i];
    st.hb_rejects[chan][iq] = rejects;
    // quad is advanced by advance_streams_kernel once per call (both iq threads share it)
}

template <int ORD>
static cudaError_t launch_hb_ord(bool kahan, const HbCoef &coef, const DevChain &ch, DevStream *streams,
                                 int n_streams, int64_t n_frames, const uint8_t *in, size_t in_stride,
                                 double *analytic, cudaStream_t s)
{
    int threads = 128;
    int blocks = (n_streams * 4 + threads - 1) / threads;
    if (ch.fp_check) {
        if (kahan) hb_exact_kernel<ORD, true, true><<<blocks, threads, 0, s>>>(coef, ch, streams, n_streams, n_frames, in, in_stride, analytic);
        else       hb_exact_kernel<ORD, false, true><<<blocks, threads, 0, s>>>(coef, ch, streams, n_streams, n_frames, in, in_stride, analytic);
    } else if (kahan)
        hb_exact_kernel<ORD, true, false><<<blocks, threads, 0, s>>>(coef, ch, streams, n_streams, n_frames, in, in_stride, analytic);
    else
        hb_exact_kernel<ORD, false, false><<<blocks, threads, 0, s>>>(coef, ch, streams, n_streams, n_frames, in, in_stride, analytic);
    return cudaGetLastError();
}

cudaError_t launch_hb_exact(const HbCoef &coef, const DevChain &ch, DevStream *streams, int n_streams,
                            int64_t n_frames, const uint8_t *in, size_t in_stride, double *analytic,
                            cudaStream_t s)
{
    switch (ch.hb_ord) {
    case 15: return launch_hb_ord<15>(ch.is_kahan, coef, ch, streams, n_streams, n_frames, in, in_stride, analytic, s);
    case 18: return launch_hb_ord<18>(ch.is_kahan, coef, ch, streams, n_streams, n_frames, in, in_stride, analytic, s);
    case 19: return launch_hb_ord<19>(ch.is_kahan, coef, ch, streams, n_streams, n_frames, in, in_stride, analytic, s);
    case 20: return launch_hb_ord<20>(ch.is_kahan, coef, ch, streams, n_streams, n_frames, in, in_stride, analytic, s);
    default: return cudaErrorInvalidValue;
    }
}

// leaf: hq_rp_process over independent channels given as doubles (stage-wise parity tap T2)
template <int ORD, bool KAHAN>
__global__ void __launch_bounds__(64)
hb_leaf_kernel(const __grid_constant__ HbCoef coef, int reject, int n_chan, int64_t n,
               const double *__restrict__ x, double *__restrict__ out_iq, HbLeafState *__restrict__ states)
{
    int cid = blockIdx.x * blockDim.x + threadIdx.x;
    if (cid >= n_chan * 2) return;
    const int iq = cid & 1, chan = cid >> 1;
    HbLeafState &st = states[chan];
    double z[ORD];
#pragma unroll
    for (int i = 0; i < ORD; ++i) z[i] = st.z[iq][i];
    unsigned long long rejects = st.rejects[iq];
    const unsigned q0 = st.quad;
    const double *src = x + (size_t)chan * n;
    double *dst = out_iq + (size_t)chan * n * 2;
    hb_run<ORD, KAHAN>(z, rejects, coef, reject, n,
        [&](int64_t i) { return mix_down(iq, (q0 + (unsigned)i) & 3u, src[i]); },
        [&](int64_t i, double y) {
            int slot;
            double v = mix_up(iq, (q0 + (unsigned)i) & 3u, y, slot);
            dst[i * 2 + slot] = v;
        });
#pragma unroll
    for (int i = 0; i < ORD; ++i) st.z[iq][i] = z[i];
    st.rejects[iq] = rejects;
    __syncwarp();
}

__global__ void hb_leaf_advance_kernel(HbLeafState *states, int n_chan, int64_t n)
{
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n_chan) states[c].quad = (unsigned)((states[c].quad + (uint64_t)n) & 3u);
}

template <int ORD>
static cudaError_t launch_leaf_ord(bool kahan, const HbCoef &coef, int reject, int n_chan, int64_t n,
                                   const double *x, double *out, HbLeafState *st, cudaStream_t s)
{
    int threads = 64, blocks = (n_chan * 2 + threads - 1) / threads;
    if (kahan) hb_leaf_kernel<ORD, true><<<blocks, threads, 0, s>>>(coef, reject, n_chan, n, x, out, st);
    else       hb_leaf_kernel<ORD, false><<<blocks, threads, 0, s>>>(coef, reject, n_chan, n, x, out, st);
    hb_leaf_advance_kernel<<<(n_chan + 127) / 128, 128, 0, s>>>(st, n_chan, n);
    return cudaGetLastError();
}

cudaError_t launch_hb_leaf(const HbCoef &coef, int ord, bool kahan, int reject, int n_chan, int64_t n,
                           const double *x, double *out, HbLeafState *st, cudaStream_t s)
{
    switch (ord) {
    case 15: return launch_leaf_ord<15>(kahan, coef, reject, n_chan, n, x, out, st, s);
    case 18: return launch_leaf_ord<18>(kahan, coef, reject, n_chan, n, x, out, st, s);
    case 19: return launch_leaf_ord<19>(kahan, coef, reject, n_chan, n, x, out, st, s);
    case 20: return launch_leaf_ord<20>(kahan, coef, reject, n_chan, n, x, out, st, s);
    default: return cudaErrorInvalidValue;
    }
}

// =============================================================================================
// MT19937 words (reference src/mersene_twister/mt_jrnd.c:99-134)
// =============================================================================================

// One CTA regenerates `blocks_per_unit` consecutive 624-word blocks starting from its checkpoint
// (the untempered state array that PRECEDES its first block) and writes the tempered words.
// new[k] = far ^ twist(old[k], old[k+1]) with far = old[k+397] for k < 227 and new[k-227] after
// that: thread t < 227 makes the words t, t+227, t+454 as one dependent chain -- each "far" after
// the first is the word it made the step before -- so a block needs no barrier inside it, only the
// one that hands the finished block (double-buffered in shared memory) to the next iteration.
//
// The thread keeps its own three words in registers from block to block (they are the old[k] of
// its next step) and reads only the neighbours' from shared memory.  Shifts go through the
// multiplier (IMAD / IMAD.HI) so that the logic pipe is left with ~7 LOP3 a word: the kernel is
// then bound by the HBM write of the words, not by instruction issue.
__global__ void __launch_bounds__(MT_WORDS_THREADS, 8)
mt_words_kernel(const uint32_t *__restrict__ ckpt /* [unit][624] */, int n_units, int blocks_per_unit,
                int64_t first_word /* stream word index of block 0's first word */,
                int64_t want_lo, int64_t want_hi, uint32_t *__restrict__ out /* out[w - want_lo] */,
                int64_t tail_block /* relative index of the block holding the last wanted word */,
                uint32_t *__restrict__ tail /* [2][624]: state before and after that block, for the next call */)
{
    __shared__ uint32_t buf[2][ICW_MT_N + 8];       // + slack: thread 169 reads old[624] before overriding it
    const int t = threadIdx.x;
    const int unit = blockIdx.x;
    for (int i = t; i < ICW_MT_N; i += MT_WORDS_THREADS) buf[0][i] = ckpt[(size_t)unit * ICW_MT_N + i];
    if (t < 8) buf[0][ICW_MT_N + t] = buf[1][ICW_MT_N + t] = 0u;
    __syncthreads();
    const bool active = t < 227;
    uint32_t s0 = 0u, s1 = 0u, s2 = 0u;
    if (active) { s0 = buf[0][t]; s1 = buf[0][t + 227]; if (t < 170) s2 = buf[0][t + 454]; }
    // block b of this unit holds words [u0 + 624 b, u0 + 624 (b + 1)): which of them are wanted
    // in full, in part (the two ends of the range, and the block whose states go to `tail`), not at all
    const int64_t u0 = first_word + (int64_t)unit * blocks_per_unit * ICW_MT_N;
    int64_t nb64 = (want_hi - u0 + ICW_MT_N - 1) / ICW_MT_N;                    // blocks that start below want_hi
    const int n_blk = (int)(nb64 < 0 ? 0 : nb64 > blocks_per_unit ? blocks_per_unit : nb64);
    int64_t lo64 = (want_lo - u0 + ICW_MT_N - 1) / ICW_MT_N;                    // first block that starts at or above want_lo
    const int e_lo = (int)(lo64 < 0 ? 0 : lo64 > n_blk ? n_blk : lo64);
    int64_t hi64 = (want_hi - u0) / ICW_MT_N;                                   // first block that ends above want_hi
    int e_hi = (int)(hi64 < 0 ? 0 : hi64 > n_blk ? n_blk : hi64);
    const int64_t tl = tail ? tail_block - (int64_t)unit * blocks_per_unit : -1;
    const int tail_blk = (tl >= 0 && tl < n_blk) ? (int)tl : -1;
    if (tail_blk >= 0 && tail_blk < e_hi) e_hi = tail_blk;
    uint32_t *o = out + (u0 - want_lo) + t;
    // one block with every special case: partial ranges, the tail states
    auto step = [&](int blk, int cur) {
        const uint32_t *old = buf[cur];
        uint32_t *nw = buf[cur ^ 1];
        const bool is_tail = blk == tail_blk;
        if (is_tail)
            for (int i = t; i < ICW_MT_N; i += MT_WORDS_THREADS) tail[i] = old[i];
        if (active) {
            if (blk < e_lo || blk >= e_hi) mt_words_block<true>(old, nw, t, s0, s1, s2, o, u0 + (int64_t)blk * ICW_MT_N, want_lo, want_hi);
            else                           mt_words_block<false>(old, nw, t, s0, s1, s2, o, 0, 0, 0);
        }
        __syncthreads();
        if (is_tail)
            for (int i = t; i < ICW_MT_N; i += MT_WORDS_THREADS) tail[ICW_MT_N + i] = nw[i];
        o += ICW_MT_N;
    };
    int blk = 0;
    for (; blk < n_blk && blk < e_lo; ++blk) step(blk, blk & 1);
    // interior blocks two at a time: fixed buffer roles, no per-block case analysis
    if (blk & 1) { if (blk < n_blk) { step(blk, 1); ++blk; } }
    for (; blk + 2 <= e_hi; blk += 2, o += 2 * ICW_MT_N) {
        if (active) mt_words_block<false>(buf[0], buf[1], t, s0, s1, s2, o, 0, 0, 0);
        __syncthreads();
        if (active) mt_words_block<false>(buf[1], buf[0], t, s0, s1, s2, o + ICW_MT_N, 0, 0, 0);
        __syncthreads();
    }
    for (; blk < n_blk; ++blk) step(blk, blk & 1);
}

cudaError_t launch_mt_words(const uint32_t *ckpt, int n_units, int blocks_per_unit, int64_t first_word,
                            int64_t want_lo, int64_t want_hi, uint32_t *out, int64_t tail_block, uint32_t *tail,
                            cudaStream_t s)
{
    mt_words_kernel<<<n_units, MT_WORDS_THREADS, 0, s>>>(ckpt, n_units, blocks_per_unit, first_word,
                                                         want_lo, want_hi, out, tail_block, tail);
    return cudaGetLastError();
}

// =============================================================================================
// the pointwise chain: unpack -> oscillator -> DSP list -> render
// =============================================================================================

// one thread per frame, grid-stride inside a stream (blockIdx.y = stream).
// src: raw file bytes (complex formats) or the analytic scratch written by hb_exact_kernel
// (from_analytic: 4 doubles per frame, no fade -- it was applied before the Hilbert converter).
#ifndef ICW_CHAIN_THREADS
#define ICW_CHAIN_THREADS 256
#define ICW_CHAIN_CTAS 4
#endif
// the interpreter's own launch shape (A/B: profiles/r2_ab_interpreter.txt)
#ifndef ICW_GCHAIN_THREADS
#define ICW_GCHAIN_THREADS 256
#define ICW_GCHAIN_CTAS 4
#endif
__global__ void __launch_bounds__(ICW_GCHAIN_THREADS, ICW_GCHAIN_CTAS)
chain_kernel(const __grid_constant__ DevChain ch, DevStream *__restrict__ streams, int64_t n_frames,
             const uint8_t *__restrict__ in, size_t in_stride, int from_analytic,
             const uint32_t *__restrict__ mtw_l, const uint32_t *__restrict__ mtw_r, size_t mt_stream_stride,
             uint8_t *__restrict__ out, size_t out_stride,
             double *__restrict__ tap_bus /* optional [stream][frame][ICW_N_PLUGS][4] */,
             double *__restrict__ tap_lr /* optional [stream][frame][2] */,
             double *__restrict__ pre /* noise shaping: [stream][frame][4] values + dither instead of PCM */)
{
    const int stream = blockIdx.y;
    DevStream &st = streams[stream];
    const uint8_t *src = in + (size_t)stream * in_stride;
    const size_t mt_off = (size_t)stream * mt_stream_stride;    // 0 when every stream shares one generator state
    FrameIO io;
    io.mtw_l = mtw_l ? mtw_l + mt_off : nullptr;
    io.mtw_r = mtw_r ? mtw_r + mt_off : nullptr;
    io.dst = out + (size_t)stream * out_stride;
    io.dst_aligned = ((size_t)(uintptr_t)io.dst & 3u) == 0;
    io.tap_bus = tap_bus ? tap_bus + (size_t)stream * n_frames * (ICW_N_PLUGS * 4) : nullptr;
    io.tap_lr = tap_lr ? tap_lr + (size_t)stream * n_frames * 2 : nullptr;
    io.pre = pre ? pre + (size_t)stream * n_frames * 4 : nullptr;

    FrameAcc acc;
    double bus[ICW_N_PLUGS][4];
    load_bus(st, bus);
    OscCounter osc;
    osc.init(ch, st.n_frame, (int64_t)blockIdx.x * blockDim.x + threadIdx.x);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_frames;
         i += (int64_t)gridDim.x * blockDim.x) {
        double v[4];
        {
            // the thread's next frame is a grid stride away: asked for now, it is in L1 when its turn comes (the load at the top
            // of a trip was 8 % of the kernel's stall samples on C3)
            const int64_t nx = i + (int64_t)gridDim.x * blockDim.x;
            if (nx < n_frames) {
                const uint8_t *pn = from_analytic ? src + nx * 32 : src + nx * ch.frame_bytes;
                asm volatile("prefetch.global.L1 [%0];" :: "l"(pn));
            }
        }
        if (from_analytic) {
            const double *a = reinterpret_cast<const double *>(src) + i * 4;
            v[0] = a[0]; v[1] = a[1]; v[2] = a[2]; v[3] = a[3];
        } else {
            unpack_frame(ch, src + i * ch.frame_bytes, st.pos + i, v);
        }
        finish_frame(ch, st, i, n_frames, v, bus, io, acc, osc);
    }
    commit_acc(&st, acc, 32);
}

// Feedback lists (a node reads a plug that a LATER node of the list writes, or its own output): the value read is the
// previous frame's (the bus persists in the context, reference src/adv_modulator.c:634-751), so the frames of a stream
// are serial.  One thread per stream walks them in order with its bus in local memory -- the same frame path, the bus
// simply never reloaded; parallel only across streams (like the noise shapers).
__global__ void __launch_bounds__(32)
chain_serial_kernel(const __grid_constant__ DevChain ch, DevStream *__restrict__ streams, int n_streams, int64_t n_frames,
                    const uint8_t *__restrict__ in, size_t in_stride, int from_analytic,
                    const uint32_t *__restrict__ mtw_l, const uint32_t *__restrict__ mtw_r, size_t mt_stream_stride,
                    uint8_t *__restrict__ out, size_t out_stride, double *__restrict__ tap_bus, double *__restrict__ tap_lr,
                    double *__restrict__ pre)
{
    const int stream = blockIdx.x * blockDim.x + threadIdx.x;
    if (stream >= n_streams) return;
    DevStream &st = streams[stream];
    const uint8_t *src = in + (size_t)stream * in_stride;
    const size_t mt_off = (size_t)stream * mt_stream_stride;
    FrameIO io;
    io.mtw_l = mtw_l ? mtw_l + mt_off : nullptr;
    io.mtw_r = mtw_r ? mtw_r + mt_off : nullptr;
    io.dst = out + (size_t)stream * out_stride;
    io.dst_aligned = ((size_t)(uintptr_t)io.dst & 3u) == 0;
    io.tap_bus = tap_bus ? tap_bus + (size_t)stream * n_frames * (ICW_N_PLUGS * 4) : nullptr;
    io.tap_lr = tap_lr ? tap_lr + (size_t)stream * n_frames * 2 : nullptr;
    io.pre = pre ? pre + (size_t)stream * n_frames * 4 : nullptr;
    FrameAcc acc;
    double bus[ICW_N_PLUGS][4];
    load_bus(st, bus);
    bus[0][0] = st.bus[0][0]; bus[0][1] = st.bus[0][1]; bus[0][2] = st.bus[0][2]; bus[0][3] = st.bus[0][3];
    OscCounter osc;
    osc.init(ch, st.n_frame, 0);
    for (int64_t i = 0; i < n_frames; ++i) {
        double v[4];
        if (from_analytic) {
            const double *a = reinterpret_cast<const double *>(src) + i * 4;
            v[0] = a[0]; v[1] = a[1]; v[2] = a[2]; v[3] = a[3];
        } else {
            unpack_frame(ch, src + i * ch.frame_bytes, st.pos + i, v);
        }
        if (ch.feedback & 2) finish_frame<DITHER_SERIAL>(ch, st, i, n_frames, v, bus, io, acc, osc);      // replay of a rejected draw's frame
        else finish_frame<DITHER_LATE>(ch, st, i, n_frames, v, bus, io, acc, osc);
    }
    if (acc.clips_l) atomicAdd(&st.clips[0], acc.clips_l);
    if (acc.clips_r) atomicAdd(&st.clips[1], acc.clips_r);
    if (acc.redraws) atomicAdd(&st.mt_redraws, (unsigned long long)acc.redraws);
    atomicMax(reinterpret_cast<unsigned long long *>(&st.peak[0]), (unsigned long long)__double_as_longlong(acc.peak_l));
    atomicMax(reinterpret_cast<unsigned long long *>(&st.peak[1]), (unsigned long long)__double_as_longlong(acc.peak_r));
}

// The same pass for the straight-line lists with plain PCM output and no dither or a 2-/4-word one:
// list shape and dither type fixed at compile time (lean_frame, icw_frame.cuh), no thread-private bus.
template <int SHAPE, int RT>
__global__ void __launch_bounds__(ICW_CHAIN_THREADS, ICW_CHAIN_CTAS)
chain_lean_kernel(const __grid_constant__ DevChain ch, DevStream *__restrict__ streams, int64_t n_frames,
                  const uint8_t *__restrict__ in, size_t in_stride, int from_analytic,
                  const uint32_t *__restrict__ mtw_l, const uint32_t *__restrict__ mtw_r, size_t mt_stream_stride,
                  uint8_t *__restrict__ out, size_t out_stride)
{
    constexpr int WPS = RT == ICW_RENDER_TPDF ? 4 : RT == ICW_RENDER_RPDF ? 2 : 0;
    const int stream = blockIdx.y;
    DevStream &st = streams[stream];
    const uint8_t *src = in + (size_t)stream * in_stride;
    const size_t mt_off = (size_t)stream * mt_stream_stride;
    const uint32_t *wl_p = WPS ? mtw_l + mt_off : nullptr, *wr_p = WPS ? mtw_r + mt_off : nullptr;
    uint8_t *dst = out + (size_t)stream * out_stride;
    const int dst_aligned = ((size_t)(uintptr_t)dst & 3u) == 0;
    const int64_t pos0 = st.pos;
    FrameAcc acc;
    OscCounter osc;
    osc.init(ch, st.n_frame, (int64_t)blockIdx.x * blockDim.x + threadIdx.x);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_frames;
         i += (int64_t)gridDim.x * blockDim.x) {
        const uint4 wl = dither_fetch(WPS, wl_p, i), wr = dither_fetch(WPS, wr_p, i);
        double v[4];        // (a prefetch of the thread's next frame, which pays in chain_kernel, costs 5 % here: measured)
        if (SHAPE == ICW_SHAPE_SHIFT_MASTER_FAST || from_analytic) {
            const double2 *a = reinterpret_cast<const double2 *>(src) + i * 2;
            const double2 a0 = a[0], a1 = a[1];
            v[0] = a0.x; v[1] = a0.y; v[2] = a1.x; v[3] = a1.y;
        } else {
            unpack_frame(ch, src + i * ch.frame_bytes, pos0 + i, v);
        }
        if (SHAPE == ICW_SHAPE_SHIFT_MASTER_FAST) lean_frame_fast<RT>(ch, st, i, n_frames - 1, v, wl, wr, dst, acc, osc);
        else lean_frame<SHAPE, RT>(ch, st, i, n_frames - 1, v, wl, wr, dst, dst_aligned, acc, osc);
    }
    commit_acc(&st, acc, 32);
}

cudaError_t launch_chain(const DevChain &ch, DevStream *streams, int n_streams, int64_t n_frames,
                         const uint8_t *in, size_t in_stride, int from_analytic,
                         const uint32_t *mtw_l, const uint32_t *mtw_r, size_t mt_stream_stride,
                         uint8_t *out, size_t out_stride, double *tap_bus, double *tap_lr, double *pre,
                         int sm_count, cudaStream_t s)
{
    int threads = ICW_CHAIN_THREADS;
    int64_t need = (n_frames + threads - 1) / threads;
    int per_stream = (int)(need < 1 ? 1 : need);
    // keep the grid near a few waves of the machine
    int cap = (sm_count * 2 * ICW_CHAIN_CTAS + n_streams - 1) / n_streams;
    if (cap < 1) cap = 1;
    if (per_stream > cap) per_stream = cap;
    dim3 grid(per_stream, n_streams);
    const int rt = ch.render.render_type;
    if (ch.feedback) {
        chain_serial_kernel<<<(n_streams + 31) / 32, 32, 0, s>>>(ch, streams, n_streams, n_frames, in, in_stride, from_analytic, mtw_l, mtw_r,
                                                                 mt_stream_stride, out, out_stride, tap_bus, tap_lr, pre);
        return cudaGetLastError();
    }
    if (!tap_bus && !tap_lr && !pre && !ch.bypass && !ch.fp_check && (ch.shape == ICW_SHAPE_MASTER || ch.shape == ICW_SHAPE_SHIFT_MASTER) &&
        (rt == ICW_RENDER_ROUND || rt == ICW_RENDER_RPDF || rt == ICW_RENDER_TPDF)) {
#define ICW_LEAN(SH, RT) chain_lean_kernel<SH, RT><<<grid, threads, 0, s>>>(ch, streams, n_frames, in, in_stride, from_analytic, \
                                                                            mtw_l, mtw_r, mt_stream_stride, out, out_stride)
        // every stream's row must be 4-byte aligned for the fast path's 16-bit PCM stores
        const bool fast = from_analytic && lean_fast_ok(ch) && ((size_t)(uintptr_t)out & 3u) == 0 && (n_streams == 1 || (out_stride & 3u) == 0);
        if (fast) {
            if (rt == ICW_RENDER_ROUND) ICW_LEAN(ICW_SHAPE_SHIFT_MASTER_FAST, ICW_RENDER_ROUND);
            else if (rt == ICW_RENDER_RPDF) ICW_LEAN(ICW_SHAPE_SHIFT_MASTER_FAST, ICW_RENDER_RPDF);
            else ICW_LEAN(ICW_SHAPE_SHIFT_MASTER_FAST, ICW_RENDER_TPDF);
        } else if (ch.shape == ICW_SHAPE_MASTER) {
            if (rt == ICW_RENDER_ROUND) ICW_LEAN(ICW_SHAPE_MASTER, ICW_RENDER_ROUND);
            else if (rt == ICW_RENDER_RPDF) ICW_LEAN(ICW_SHAPE_MASTER, ICW_RENDER_RPDF);
            else ICW_LEAN(ICW_SHAPE_MASTER, ICW_RENDER_TPDF);
        } else {
            if (rt == ICW_RENDER_ROUND) ICW_LEAN(ICW_SHAPE_SHIFT_MASTER, ICW_RENDER_ROUND);
            else if (rt == ICW_RENDER_RPDF) ICW_LEAN(ICW_SHAPE_SHIFT_MASTER, ICW_RENDER_RPDF);
            else ICW_LEAN(ICW_SHAPE_SHIFT_MASTER, ICW_RENDER_TPDF);
        }
#undef ICW_LEAN
        return cudaGetLastError();
    }
    {
        const int gthreads = ICW_GCHAIN_THREADS;
        const int64_t gneed = (n_frames + gthreads - 1) / gthreads;
        int gper = (int)(gneed < 1 ? 1 : gneed);
        int gcap = (sm_count * 2 * ICW_GCHAIN_CTAS + n_streams - 1) / n_streams;
        if (gcap < 1) gcap = 1;
        if (gper > gcap) gper = gcap;
        chain_kernel<<<dim3(gper, n_streams), gthreads, 0, s>>>(ch, streams, n_frames, in, in_stride, from_analytic, mtw_l, mtw_r,
                                                                mt_stream_stride, out, out_stride, tap_bus, tap_lr, pre);
    }
    return cudaGetLastError();
}

// =============================================================================================
// noise-shaped quantiser (reference src/sound_render.c:403-489 filters, :753-810 the feedback)
// =============================================================================================
// The shaper feeds the quantisation error of sample n into sample n + 1: one thread per
// (stream, channel) walks its channel in order -- parallel over streams only, which is why
// chain_kernel stops before the quantiser (FrameIO::pre) when a shaper is on.  The error memory is a
// register array ordered by age (static indices after unrolling), the reference's circular buffer
// read newest-first is the same sum.  ORD is the shaper's order; plain mul + add as in the reference.
template <int ORD, bool IIR, bool CHECK>
__global__ void __launch_bounds__(64)
ns_render_kernel(const __grid_constant__ DevChain ch, DevStream *__restrict__ streams, int n_streams, int64_t n_frames,
                 const double *__restrict__ pre, uint8_t *__restrict__ out, size_t out_stride)
{
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n_streams * 2) return;
    const int stream = id >> 1, c = id & 1;
    const DevRender &q = ch.render;
    DevStream &st = streams[stream];
    double e[ORD], o[IIR ? ORD : 1], ce[ORD], co[IIR ? ORD : 1];
#pragma unroll
    for (int i = 0; i < ORD; ++i) { e[i] = st.ns_e[c][i]; ce[i] = q.ns_coef[i]; }
    if (IIR) {
#pragma unroll
        for (int i = 0; i < ORD; ++i) { o[i] = st.ns_o[c][i]; co[i] = q.ns_coef[ICW_NS_MAX_TAPS + i]; }
    }
    double prev_err = st.ns_prev_err[c];
    unsigned clips = 0;
    double peak = 0.0;
    const double2 *src = reinterpret_cast<const double2 *>(pre + (size_t)stream * n_frames * 4) + c;
    uint8_t *dst = out + (size_t)stream * out_stride + c * q.bytes;
    // (value, dither) pairs are fetched a group of NS_AHEAD frames ahead: one thread per channel means one or two warps
    // per SM sub-partition, nothing else hides the DRAM latency of a load, and a sample is only ~250 cycles of work
    constexpr int NS_AHEAD = 8;
    double2 grp[NS_AHEAD], ngrp[NS_AHEAD];
#pragma unroll
    for (int j = 0; j < NS_AHEAD; ++j) grp[j] = n_frames > 0 ? src[2 * (j < n_frames ? j : n_frames - 1)] : make_double2(0.0, 0.0);
    for (int64_t i0 = 0; i0 < n_frames; i0 += NS_AHEAD) {
#pragma unroll
        for (int j = 0; j < NS_AHEAD; ++j) {
            const int64_t f = i0 + NS_AHEAD + j;
            ngrp[j] = src[2 * (f < n_frames ? f : n_frames - 1)];
        }
#pragma unroll
        for (int j = 0; j < NS_AHEAD; ++j) {
        const int64_t i = i0 + j;
        if (i >= n_frames) break;
        const double2 cur = grp[j];
        // CHECK: the FP-exception-checked twins (src/sound_render.c:846-897, :415-441, :458-489)
        uint32_t *cnt = st.fp_cnt[2 + c];
        auto F = [&](double x) { return CHECK ? fc(x, cnt) : x; };
        const double v = F(F(cur.x * q.norm_mul) - prev_err);
        double qv = F(v + F(cur.y * q.dth_mul));
        int delta;
        if (qv < 0.0) { qv = F(qv - q.round_off); delta = q.neg_delta; }
        else          { qv = F(qv + q.round_off); delta = 0; }
        peak = fmax(peak, fabs(qv) * q.inv_hi);
        if (qv >= q.hi) { qv = q.hi - 1.0; ++clips; }
        if (qv <= q.lo) { qv = q.lo + 1.0; ++clips; }
        int val = __double2int_rz(qv) + delta;
        // the error of this sample through the shaper
        const double err = F((double)val - v);
#pragma unroll
        for (int k = ORD - 1; k > 0; --k) e[k] = e[k - 1];
        e[0] = F(err);
        double res = 0.0;
        if (IIR) {
#pragma unroll
            for (int k = 0; k < ORD; ++k) res = F(res + F(F(ce[k] * e[k]) - F(co[k] * o[k])));
#pragma unroll
            for (int k = ORD - 1; k > 0; --k) o[k] = o[k - 1];
            o[0] = res;
        } else {
#pragma unroll
            for (int k = 0; k < ORD; ++k) res = F(res + F(ce[k] * e[k]));
        }
        prev_err = res;
        val = (int)((unsigned)val << q.shift);
        store_pcm(dst + i * ch.out_frame_bytes, val, q.bytes);
        }
#pragma unroll
        for (int j = 0; j < NS_AHEAD; ++j) grp[j] = ngrp[j];
    }
#pragma unroll
    for (int i = 0; i < ORD; ++i) st.ns_e[c][i] = e[i];
    if (IIR) {
#pragma unroll
        for (int i = 0; i < ORD; ++i) st.ns_o[c][i] = o[i];
    }
    st.ns_prev_err[c] = prev_err;
    st.clips[c] += clips;
    st.peak[c] = fmax(st.peak[c], peak);
}

cudaError_t launch_ns_render(const DevChain &ch, DevStream *streams, int n_streams, int64_t n_frames,
                             const double *pre, uint8_t *out, size_t out_stride, cudaStream_t s)
{
    const int threads = 64;
    const int grid = (n_streams * 2 + threads - 1) / threads;
#define ICW_NS_CASE(O, I) if (ch.fp_check) ns_render_kernel<O, I, true><<<grid, threads, 0, s>>>(ch, streams, n_streams, n_frames, pre, out, out_stride); \
                          else ns_render_kernel<O, I, false><<<grid, threads, 0, s>>>(ch, streams, n_streams, n_frames, pre, out, out_stride); break
    if (ch.render.ns_kind == 2) {
        switch (ch.render.ns_order) {
        case 4: ICW_NS_CASE(4, true);
        default: return cudaErrorInvalidValue;
        }
    } else {
        switch (ch.render.ns_order) {
        case 5:  ICW_NS_CASE(5, false);
        case 9:  ICW_NS_CASE(9, false);
        case 15: ICW_NS_CASE(15, false);
        case 16: ICW_NS_CASE(16, false);
        case 20: ICW_NS_CASE(20, false);
        default: return cudaErrorInvalidValue;
        }
    }
#undef ICW_NS_CASE
    return cudaGetLastError();
}

// Where in a stretch of a generator's words is the first pair that mtrnd_gen_dsopen would throw away (both halves zero after
// their shifts: u == 0, the draw maps to -1; src/mersene_twister/mt_jrnd.c:218-226,245-256)?  Only run after a kernel has
// counted such a draw: the hot kernels keep a count and nothing else (the event has probability 2^-53).
__global__ void mt_find_reject_kernel(const uint32_t *__restrict__ w, int64_t n_pairs, long long base_pair, long long *first)
{
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_pairs; p += (int64_t)gridDim.x * blockDim.x)
        if (((w[2 * p] >> 5) | (w[2 * p + 1] >> 6)) == 0u) atomicMin(first, base_pair + (long long)p);
}

cudaError_t launch_mt_find_reject(const uint32_t *w, int64_t n_pairs, long long base_pair, long long *d_first, int sm_count, cudaStream_t s)
{
    const int threads = 256;
    int64_t blocks = (n_pairs + threads - 1) / threads;
    if (blocks > (int64_t)sm_count * 16) blocks = (int64_t)sm_count * 16;
    if (blocks < 1) blocks = 1;
    mt_find_reject_kernel<<<(int)blocks, threads, 0, s>>>(w, n_pairs, base_pair, d_first);
    return cudaGetLastError();
}

// after a process call: advance the per-stream scalars that are closed forms of the frame count
__global__ void advance_streams_kernel(const __grid_constant__ DevChain ch, DevStream *streams, int n_streams,
                                       int64_t n_frames, int advance_quad)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_streams) return;
    DevStream &st = streams[s];
    st.n_frame = ch.is_frmod_scaled ? (st.n_frame + (uint64_t)n_frames % ch.scale_sr) % ch.scale_sr
                                    : st.n_frame + (uint64_t)n_frames;
    st.pos += n_frames;
    if (advance_quad) {
        st.quad[0] = (unsigned)((st.quad[0] + (uint64_t)n_frames) & 3u);
        st.quad[1] = (unsigned)((st.quad[1] + (uint64_t)n_frames) & 3u);
    }
    if (ch.render.render_type == ICW_RENDER_STPDF) { st.prev_rnd[0] = st.prev_rnd_next[0]; st.prev_rnd[1] = st.prev_rnd_next[1]; }
    uint64_t words = (uint64_t)n_frames * (uint64_t)ch.render.words_per_sample;
    const bool serial = (ch.feedback & 2) != 0 && ch.render.words_per_sample != 0;     // the draws were taken one by one
    st.mt_drawn[0] += serial ? (uint64_t)st.serial_used[0] : words;
    st.mt_drawn[1] += serial ? (uint64_t)st.serial_used[1] : words;
}

cudaError_t launch_advance(const DevChain &ch, DevStream *streams, int n_streams, int64_t n_frames,
                           int advance_quad, cudaStream_t s)
{
    advance_streams_kernel<<<(n_streams + 127) / 128, 128, 0, s>>>(ch, streams, n_streams, n_frames, advance_quad);
    return cudaGetLastError();
}

// debug / parity leaf: the oscillator phase of one shift frequency for frames [n0, n0+n)
__global__ void phase_leaf_kernel(const __grid_constant__ DevChain ch, uint64_t n0, int64_t n, double f,
                                  double *__restrict__ out /* [n][2] = omega, phase */)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double om = norm_omega(ch, frame_counter(ch, n0, (uint64_t)i));
    out[i * 2] = om;
    out[i * 2 + 1] = fmod_2pi(om * f);
}

// debug / parity leaf: the modulator's sincos on [0, 2*pi) beside libdevice's
__global__ void sincos_leaf_kernel(int64_t n, const double *__restrict__ x, double *__restrict__ out /* [n][4] */)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s, c, s2, c2;
    sincos_2pi(x[i], s, c);
    sincos(x[i], &s2, &c2);
    out[i * 4] = s; out[i * 4 + 1] = c; out[i * 4 + 2] = s2; out[i * 4 + 3] = c2;
}

cudaError_t launch_sincos_leaf(int64_t n, const double *x, double *out, cudaStream_t s)
{
    sincos_leaf_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, x, out);
    return cudaGetLastError();
}

cudaError_t launch_phase_leaf(const DevChain &ch, uint64_t n0, int64_t n, double f, double *out, cudaStream_t s)
{
    phase_leaf_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ch, n0, n, f, out);
    return cudaGetLastError();
}

}  // namespace icw
