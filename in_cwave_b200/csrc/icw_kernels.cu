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

namespace icw {

// =============================================================================================
// exact half-band recurrences (reference src/hblpf.c:894-926 baseline, :1008-1057 Kahan;
// mixer src/lpf_hilbert_quad.c:129-156)
// =============================================================================================

struct Comp { double s, c; };
__device__ __forceinline__ void comp_add(Comp &a, double x)
{
    // hblpf.c:991-997 -- y = x - c; t = s + y; c = (t - s) - y; s = t
    double y = __dsub_rn(x, a.c);
    double t = __dadd_rn(a.s, y);
    a.c = __dsub_rn(__dsub_rn(t, a.s), y);
    a.s = t;
}

// three independent compensated sums advanced in lock step, written stage by stage so that the
// instruction stream alternates between the chains (ptxas otherwise emits one chain after another
// and every DADD waits its full 8-cycle latency)
__device__ __forceinline__ void comp_add3(Comp &a, double xa, Comp &b, double xb, Comp &c, double xc)
{
    double ya = __dsub_rn(xa, a.c), yb = __dsub_rn(xb, b.c), yc = __dsub_rn(xc, c.c);
    double ta = __dadd_rn(a.s, ya), tb = __dadd_rn(b.s, yb), tc = __dadd_rn(c.s, yc);
    double da = __dsub_rn(ta, a.s), db = __dsub_rn(tb, b.s), dc = __dsub_rn(tc, c.s);
    a.c = __dsub_rn(da, ya); b.c = __dsub_rn(db, yb); c.c = __dsub_rn(dc, yc);
    a.s = ta; b.s = tb; c.s = tc;
}
__device__ __forceinline__ void comp_add2(Comp &a, double xa, Comp &b, double xb)
{
    double ya = __dsub_rn(xa, a.c), yb = __dsub_rn(xb, b.c);
    double ta = __dadd_rn(a.s, ya), tb = __dadd_rn(b.s, yb);
    double da = __dsub_rn(ta, a.s), db = __dsub_rn(tb, b.s);
    a.c = __dsub_rn(da, ya); b.c = __dsub_rn(db, yb);
    a.s = ta; b.s = tb;
}

// ---------------------------------------------------------------------------------------------
// One half-band recurrence over n samples.  z[j] = state j+1 samples ago (newest first: the order
// the reference's circular walk visits them).  `in(i)` gives the mixed-down input of sample i,
// `out(i, y)` receives the filter output of sample i.
//
// Kahan variant: per sample the reference forms two compensated sums -- the state sum
// (x, z0*fb0, z1*fb1, ...: 4 dependent DADDs per term, 76 in a row) and the output sum
// (z0*ff0, (z0*fb0)*d0, z1*ff1, ...: 148 in a row).  Only the first feeds the next sample, so
// the output sum of sample n-1 (first half) and of sample n-2 (second half) are evaluated while
// the state sum of sample n runs: three independent dependency chains of ~75 DADDs each per
// iteration instead of one of ~225.  Every operation and its order within each sum is the
// reference's; only the interleaving in time differs, which rounding cannot see.
// ---------------------------------------------------------------------------------------------
template <int ORD, int I0, int I1, int OFF>
__device__ __forceinline__ void out_sum_part(Comp &o, const double (&z)[ORD + 2], const HbCoef &k)
{
#pragma unroll
    for (int i = I0; i < I1; ++i) {
        const double zi = z[i + OFF];
        const double t = __dmul_rn(zi, k.fb[i]);
        if (i == 0) { o.s = __dmul_rn(zi, k.ff[0]); o.c = 0.0; }
        else comp_add(o, __dmul_rn(zi, k.ff[i]));
        comp_add(o, __dmul_rn(t, k.d0));
    }
}

template <int ORD, bool KAHAN, class In, class Out>
__device__ __forceinline__ void hb_run(double *zstate, unsigned long long &rejects, const HbCoef &k,
                                       int reject, int64_t n, In in, Out out)
{
    const double thr = (double)reject;              // compared as a number, bug-for-bug (hblpf.c:915,1046)
    if (n <= 0) return;
    if (KAHAN) {
        constexpr int H = (ORD + 1) / 2;
        double z[ORD + 2];
#pragma unroll
        for (int i = 0; i < ORD; ++i) z[i] = zstate[i];
        z[ORD] = z[ORD + 1] = 0.0;                  // only reach outputs of samples before this call
        Comp o_half; o_half.s = 0.0; o_half.c = 0.0;
        double xc = in(0);
        for (int64_t i = 0; i < n; ++i) {
            const double xn = in(i + 1 < n ? i + 1 : i);        // next input is fetched a sample ahead
            // ---- one basic block: three independent dependency chains, interleaved by hand ------
            // term lists in the reference's order:
            //   state sum  a : x, z0*fb0, z1*fb1, ...                      (ORD additions)
            //   output sum of sample i-1, first half  (o_new):  z1*ff0 | t0*d0, z2*ff1, t1*d0, ...
            //   output sum of sample i-2, second half (o_half): z[H+2]*ffH, tH*d0, ...
            Comp a; a.s = xc; a.c = 0.0;
            Comp o_new;
            {
                constexpr int NA = ORD;                 // additions into a
                constexpr int NB = 2 * H - 1;           // additions into o_new (its first term initialises)
                constexpr int NC = 2 * (ORD - H);       // additions into o_half
                const double tb0 = __dmul_rn(z[1], k.fb[0]);
                o_new.s = __dmul_rn(z[1], k.ff[0]); o_new.c = 0.0;
                constexpr int NSTEP = NA > NB ? (NA > NC ? NA : NC) : (NB > NC ? NB : NC);
#pragma unroll
                for (int st = 0; st < NSTEP; ++st) {
                    // the st-th addend of each chain (compile-time indices)
                    const double xa = st < NA ? __dmul_rn(z[st < NA ? st : 0], k.fb[st < NA ? st : 0]) : 0.0;
                    // o_new addends: st = 0 -> t0*d0; then pairs (z*ff[i], t_i*d0) for i = 1..H-1
                    const int ib = (st + 1) / 2;                                   // filter tap index
                    const int ibc = ib < H ? ib : 0;
                    const double xb = st < NB
                        ? ((st & 1) == 0 ? __dmul_rn(st == 0 ? tb0 : __dmul_rn(z[ibc + 1], k.fb[ibc]), k.d0)
                                         : __dmul_rn(z[ibc + 1], k.ff[ibc]))
                        : 0.0;
                    // o_half addends: pairs (z*ff[i], t_i*d0) for i = H..ORD-1
                    const int ic = H + st / 2;
                    const int icc = ic < ORD ? ic : H;
                    const double xc2 = st < NC
                        ? ((st & 1) == 0 ? __dmul_rn(z[icc + 2], k.ff[icc])
                                         : __dmul_rn(__dmul_rn(z[icc + 2], k.fb[icc]), k.d0))
                        : 0.0;
                    if (st < NA && st < NB && st < NC) comp_add3(a, xa, o_new, xb, o_half, xc2);
                    else if (st < NA && st < NB) comp_add2(a, xa, o_new, xb);
                    else if (st < NA && st < NC) comp_add2(a, xa, o_half, xc2);
                    else if (st < NB && st < NC) comp_add2(o_new, xb, o_half, xc2);
                    else if (st < NA) comp_add(a, xa);
                    else if (st < NB) comp_add(o_new, xb);
                    else if (st < NC) comp_add(o_half, xc2);
                }
            }
            double w = a.s;
            const bool rj = (reject != 0) & (fabs(w) < thr);
            w = rj ? 0.0 : w;
            rejects += rj ? 1ull : 0ull;
            const double y2 = o_half.s;                         // no d0*x term: bug-for-bug (hblpf.c:1056)
            // ------------------------------------------------------------------------------------
            if (i >= 2) out(i - 2, y2);
            o_half = o_new;
#pragma unroll
            for (int j = ORD + 1; j > 0; --j) z[j] = z[j - 1];
            z[0] = w;
            xc = xn;
        }
        // drain the two output sums still in flight (no shift in between: both read z[1..])
        {
            Comp o_new; o_new.s = 0.0; o_new.c = 0.0;
            out_sum_part<ORD, 0, H, 1>(o_new, z, k);
            out_sum_part<ORD, H, ORD, 2>(o_half, z, k);
            if (n >= 2) out(n - 2, o_half.s);
            out_sum_part<ORD, H, ORD, 1>(o_new, z, k);
            out(n - 1, o_new.s);
        }
#pragma unroll
        for (int i = 0; i < ORD; ++i) zstate[i] = z[i];
    } else {
        double z[ORD];
#pragma unroll
        for (int i = 0; i < ORD; ++i) z[i] = zstate[i];
        double xc = in(0);
        for (int64_t i = 0; i < n; ++i) {
            const double xn = in(i + 1 < n ? i + 1 : i);
            double acc_in = xc, acc_out = 0.0;
#pragma unroll
            for (int j = 0; j < ORD; ++j) {
                acc_in = __dadd_rn(acc_in, __dmul_rn(z[j], k.fb[j]));
                acc_out = __dadd_rn(acc_out, __dmul_rn(z[j], k.ff[j]));
            }
            double w = acc_in;
            const bool rj = (reject != 0) & (fabs(w) < thr);
            w = rj ? 0.0 : w;
            rejects += rj ? 1ull : 0ull;
            out(i, __dadd_rn(__dmul_rn(w, k.d0), acc_out));
#pragma unroll
            for (int j = ORD - 1; j > 0; --j) z[j] = z[j - 1];
            z[0] = w;
            xc = xn;
        }
#pragma unroll
        for (int i = 0; i < ORD; ++i) zstate[i] = z[i];
    }
}

// fs/4 mixer around one filter (reference src/lpf_hilbert_quad.c:132-153).  Down-mix: the I filter
// gets (+x, 0, -x, 0), the Q filter (0, -x, 0, +x).  Up-mix and *2: the I filter feeds
// (+re, +im, -re, -im), the Q filter (+im, -re, -im, +re); slot 0 = re, 1 = im.
__device__ __forceinline__ double mix_down(int iq, unsigned q, double x)
{
    if (iq == 0) return (q == 0) ? x : (q == 2) ? -x : 0.0;
    return (q == 1) ? -x : (q == 3) ? x : 0.0;
}
__device__ __forceinline__ double mix_up(int iq, unsigned q, double y, int &slot)
{
    double v = __dmul_rn(y, 2.0);
    if (iq == 0) { slot = q & 1; if (q >= 2) v = -v; }
    else         { slot = (q & 1) ^ 1; if (q == 1 || q == 2) v = -v; }
    return v;
}

// chain id = ((stream * 2 + channel) * 2 + iq).  Output: analytic frames as 4 doubles
// (L.re, L.im, R.re, R.im) == the layout of ICW_FMT_CW_F64 stereo, so chain_kernel reads it back
// as complex input.
template <int ORD, bool KAHAN>
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

    hb_run<ORD, KAHAN>(z, rejects, coef, ch.reject_flag, n_frames,
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
        });
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
    if (kahan)
        hb_exact_kernel<ORD, true><<<blocks, threads, 0, s>>>(coef, ch, streams, n_streams, n_frames, in, in_stride, analytic);
    else
        hb_exact_kernel<ORD, false><<<blocks, threads, 0, s>>>(coef, ch, streams, n_streams, n_frames, in, in_stride, analytic);
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

// One CTA regenerates `blocks_per_cta` consecutive 624-word blocks starting from its checkpoint
// (the untempered state array that PRECEDES its first block) and writes the tempered words.
// Regeneration is three dependent phases of <= 227 independent words each (new[i] needs
// new[i-227] from the phase before), double-buffered in shared memory.
__global__ void __launch_bounds__(256)
mt_words_kernel(const uint32_t *__restrict__ ckpt /* [cta][624] */, int blocks_per_cta,
                int64_t first_word /* stream word index of block 0's first word */,
                int64_t want_lo, int64_t want_hi, uint32_t *__restrict__ out /* out[w - want_lo] */)
{
    __shared__ uint32_t a[ICW_MT_N], b[ICW_MT_N];
    uint32_t *cur = a, *nxt = b;
    const int tid = threadIdx.x;
    for (int i = tid; i < ICW_MT_N; i += blockDim.x) cur[i] = ckpt[(size_t)blockIdx.x * ICW_MT_N + i];
    __syncthreads();
    for (int blk = 0; blk < blocks_per_cta; ++blk) {
        // phase A: i in [0,227): new[i] = old[i+397] ^ tw(old[i], old[i+1])
        // phase B: i in [227,454): new[i] = new[i-227] ^ tw(old[i], old[i+1])
        // phase C: i in [454,624): same, and i = 623 pairs old[623] with new[0]
        for (int ph = 0; ph < 3; ++ph) {
            int i = ph * 227 + tid;
            if (tid < 227 && i < ICW_MT_N) {
                uint32_t u = cur[i];
                uint32_t v = (i + 1 < ICW_MT_N) ? cur[i + 1] : nxt[0];
                uint32_t mix = (u & 0x80000000u) | (v & 0x7FFFFFFFu);
                uint32_t tw = (mix >> 1) ^ ((v & 1u) ? 0x9908B0DFu : 0u);
                uint32_t far = (ph == 0) ? cur[i + ICW_MT_M] : nxt[i - 227];
                nxt[i] = far ^ tw;
            }
            __syncthreads();
        }
        int64_t w0 = first_word + ((int64_t)blockIdx.x * blocks_per_cta + blk) * ICW_MT_N;
        for (int i = tid; i < ICW_MT_N; i += blockDim.x) {
            int64_t w = w0 + i;
            if (w >= want_lo && w < want_hi) out[w - want_lo] = mt_temper(nxt[i]);
        }
        uint32_t *t = cur; cur = nxt; nxt = t;
        __syncthreads();
    }
}

cudaError_t launch_mt_words(const uint32_t *ckpt, int n_cta, int blocks_per_cta, int64_t first_word,
                            int64_t want_lo, int64_t want_hi, uint32_t *out, cudaStream_t s)
{
    mt_words_kernel<<<n_cta, 256, 0, s>>>(ckpt, blocks_per_cta, first_word, want_lo, want_hi, out);
    return cudaGetLastError();
}

// =============================================================================================
// the pointwise chain: unpack -> oscillator -> DSP list -> render
// =============================================================================================

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

// one thread per frame, grid-stride inside a stream (blockIdx.y = stream).
// src: raw file bytes (complex formats) or the analytic scratch written by hb_exact_kernel
// (fmt_override = ICW_FMT_CW_F64, no fade: it was applied before the Hilbert converter).
__global__ void __launch_bounds__(256)
chain_kernel(const __grid_constant__ DevChain ch, DevStream *__restrict__ streams, int64_t n_frames,
             const uint8_t *__restrict__ in, size_t in_stride, int from_analytic,
             const uint32_t *__restrict__ mtw_l, const uint32_t *__restrict__ mtw_r, int mt_shared,
             uint8_t *__restrict__ out, size_t out_stride,
             double *__restrict__ tap_bus /* optional [stream][frame][ICW_N_PLUGS][4] */,
             double *__restrict__ tap_lr /* optional [stream][frame][2] */)
{
    const int stream = blockIdx.y;
    DevStream &st = streams[stream];
    const uint8_t *src = in + (size_t)stream * in_stride;
    uint8_t *dst = out + (size_t)stream * out_stride;
    const DevRender &rq = ch.render;
    const int wps = rq.words_per_sample;
    const size_t mt_off = mt_shared ? 0 : (size_t)stream * (size_t)n_frames * wps;

    unsigned clips_l = 0, clips_r = 0, redraws = 0;
    double peak_l = 0.0, peak_r = 0.0;
    double bus[ICW_N_PLUGS][4];
    // plugs nobody writes keep whatever the context held (normally 0.0)
    for (int k = 1; k < ICW_N_PLUGS; ++k) {
        bus[k][0] = st.bus[k][0]; bus[k][1] = st.bus[k][1]; bus[k][2] = st.bus[k][2]; bus[k][3] = st.bus[k][3];
    }

    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_frames;
         i += (int64_t)gridDim.x * blockDim.x) {
        double v[4];
        if (from_analytic) {
            const double *a = reinterpret_cast<const double *>(src) + i * 4;
            v[0] = a[0]; v[1] = a[1]; v[2] = a[2]; v[3] = a[3];
        } else {
            unpack_frame(ch, src + i * ch.frame_bytes, st.pos + i, v);
        }
        bus[0][0] = v[0]; bus[0][1] = v[1]; bus[0][2] = v[2]; bus[0][3] = v[3];
        double omega = norm_omega(ch, frame_counter(ch, st.n_frame, (uint64_t)i));
        double lo, ro;
        run_graph(ch, bus, omega, lo, ro);

        double dl = 0.0, dr = 0.0;
        if (wps) {
            uint32_t wl[24], wr[24];
            for (int j = 0; j < wps; ++j) {
                wl[j] = mtw_l[mt_off + (size_t)i * wps + j];
                wr[j] = mtw_r[mt_off + (size_t)i * wps + j];
            }
            double prev_l = 0.0, prev_r = 0.0, tr;
            if (rq.render_type == ICW_RENDER_STPDF) {
                // previous frame's draw: recompute it from that frame's words (frame 0: carried state)
                bool rd;
                if (i == 0) { prev_l = st.prev_rnd[0]; prev_r = st.prev_rnd[1]; }
                else {
                    prev_l = mt_dsopen(mtw_l[mt_off + (size_t)(i - 1) * wps], mtw_l[mt_off + (size_t)(i - 1) * wps + 1], rd);
                    prev_r = mt_dsopen(mtw_r[mt_off + (size_t)(i - 1) * wps], mtw_r[mt_off + (size_t)(i - 1) * wps + 1], rd);
                }
            }
            dl = dither_value(rq, wl, prev_l, tr, redraws);
            dr = dither_value(rq, wr, prev_r, tr, redraws);
        }
        RenderOut a = render_one(rq, lo, dl);
        RenderOut b = render_one(rq, ro, dr);
        clips_l += a.clipped; clips_r += b.clipped;
        peak_l = fmax(peak_l, a.level); peak_r = fmax(peak_r, b.level);
        uint8_t *p = dst + i * ch.out_frame_bytes;
        store_pcm(p, a.val, rq.bytes);
        store_pcm(p + rq.bytes, b.val, rq.bytes);
        if (tap_bus) {
            double *t = tap_bus + ((size_t)stream * n_frames + i) * (ICW_N_PLUGS * 4);
            for (int k = 0; k < ICW_N_PLUGS; ++k) { t[k * 4] = bus[k][0]; t[k * 4 + 1] = bus[k][1]; t[k * 4 + 2] = bus[k][2]; t[k * 4 + 3] = bus[k][3]; }
        }
        if (tap_lr) {
            double *t = tap_lr + ((size_t)stream * n_frames + i) * 2;
            t[0] = lo; t[1] = ro;
        }
        if (i == n_frames - 1) {
            // the context's bus after the call == the last frame's values (adv_modulator.c:634-751)
            for (int k = 0; k < ICW_N_PLUGS; ++k) { st.bus[k][0] = bus[k][0]; st.bus[k][1] = bus[k][1]; st.bus[k][2] = bus[k][2]; st.bus[k][3] = bus[k][3]; }
            if (rq.render_type == ICW_RENDER_STPDF) {
                bool rd;
                st.prev_rnd[0] = mt_dsopen(mtw_l[mt_off + (size_t)i * wps], mtw_l[mt_off + (size_t)i * wps + 1], rd);
                st.prev_rnd[1] = mt_dsopen(mtw_r[mt_off + (size_t)i * wps], mtw_r[mt_off + (size_t)i * wps + 1], rd);
            }
        }
    }

    // counters: warp-reduce, then one atomic per warp (peak >= 0, so its bit pattern orders like an integer)
    for (int o = 16; o; o >>= 1) {
        clips_l += __shfl_xor_sync(0xffffffffu, clips_l, o);
        clips_r += __shfl_xor_sync(0xffffffffu, clips_r, o);
        redraws += __shfl_xor_sync(0xffffffffu, redraws, o);
        peak_l = fmax(peak_l, __shfl_xor_sync(0xffffffffu, peak_l, o));
        peak_r = fmax(peak_r, __shfl_xor_sync(0xffffffffu, peak_r, o));
    }
    if ((threadIdx.x & 31) == 0) {
        if (clips_l) atomicAdd(&st.clips[0], clips_l);
        if (clips_r) atomicAdd(&st.clips[1], clips_r);
        if (redraws) atomicAdd(&st.mt_redraws, (unsigned long long)redraws);
        atomicMax(reinterpret_cast<unsigned long long *>(&st.peak[0]), (unsigned long long)__double_as_longlong(peak_l));
        atomicMax(reinterpret_cast<unsigned long long *>(&st.peak[1]), (unsigned long long)__double_as_longlong(peak_r));
    }
}

cudaError_t launch_chain(const DevChain &ch, DevStream *streams, int n_streams, int64_t n_frames,
                         const uint8_t *in, size_t in_stride, int from_analytic,
                         const uint32_t *mtw_l, const uint32_t *mtw_r, int mt_shared,
                         uint8_t *out, size_t out_stride, double *tap_bus, double *tap_lr,
                         int sm_count, cudaStream_t s)
{
    int threads = 256;
    int64_t need = (n_frames + threads - 1) / threads;
    int per_stream = (int)(need < 1 ? 1 : need);
    // keep the grid near a few waves of the machine
    int cap = (sm_count * 8 + n_streams - 1) / n_streams;
    if (cap < 1) cap = 1;
    if (per_stream > cap) per_stream = cap;
    dim3 grid(per_stream, n_streams);
    chain_kernel<<<grid, threads, 0, s>>>(ch, streams, n_frames, in, in_stride, from_analytic, mtw_l, mtw_r,
                                          mt_shared, out, out_stride, tap_bus, tap_lr);
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
    uint64_t words = (uint64_t)n_frames * (uint64_t)ch.render.words_per_sample;
    st.mt_drawn[0] += words;
    st.mt_drawn[1] += words;
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

cudaError_t launch_phase_leaf(const DevChain &ch, uint64_t n0, int64_t n, double f, double *out, cudaStream_t s)
{
    phase_leaf_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ch, n0, n, f, out);
    return cudaGetLastError();
}

}  // namespace icw
