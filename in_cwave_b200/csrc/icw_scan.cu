// icw_scan.cu -- time-parallel real -> analytic conversion ("scan mode").
//
// The reference's half-band low-pass H(z) (order 15..20 elliptic, reference src/hblpf.c:740-820)
// is evaluated in its partial-fraction form  H(z) = d0 + sum_k r_k z^-1 / (1 - p_k z^-1):
// one complex one-pole recurrence s_k[n] = p_k s_k[n-1] + u[n] per conjugate pair,
// y[n] = sum_k 2 Re(r_k s_k[n-1]) (+ d0 u[n] for the baseline summation; the reference's Kahan
// path drops that term, src/hblpf.c:1056, and so do we).  One-pole recurrences are a linear
// scan: a chunk's end state from zero (pass 1), a carry scan over chunks (pass 2), the real run
// from the carried state (pass 3).  The fs/4 mixer of the converter (reference
// src/lpf_hilbert_quad.c:129-156) feeds the I filter only on even phases and the Q filter only on
// odd phases, so each state advances two samples at a time with p^2 and its zero-input output
// uses the residue r*p: 16 FMA per mode per sample PAIR per channel instead of 24.
//
// Numerics: this is the mathematically exact filter evaluated in FP64 (within 3e-15 of a
// quad-precision evaluation of the reference's recurrence, all six designs), NOT the reference's
// rounding sequence; it differs from the reference by the reference's own rounding noise
// (1e-9 .. 1e-3 of RMS depending on the design) and ignores the |w|<1
// state zeroing.  Exact mode (icw_fused.cu) is the bit-exact path.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "icw_dev.cuh"
#include "icw_hb.cuh"
#include "icw_kernels.h"
#include "icw_scan.h"
#include "icw_scan_dev.cuh"
#include "icw_sfused.h"
#include "icw_hb_modal.inc"

// 384 threads at 168 registers (12 warps per SM, ~100 B of spill per thread outside the sample loop's
// steady state) beat 256 at 207-224 (8 warps): 10.3 against 11.0 ms per C2 step; 320 and 512 lose
#ifndef ICW_APPLY_THREADS
#define ICW_APPLY_THREADS 384
#endif

namespace icw {

// ---------------------------------------------------------------------------------------------
// host: modal constants in extended precision, rounded once
// ---------------------------------------------------------------------------------------------
typedef long double ld;
struct cld { ld re, im; };
static inline cld cmul(cld a, cld b) { return { a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re }; }
static inline cld cpowi(cld a, long long e)
{
    cld r = { 1.0L, 0.0L };
    while (e) { if (e & 1) r = cmul(r, a); a = cmul(a, a); e >>= 1; }
    return r;
}

int scan_chunk_len(int n_streams, int64_t n_frames, int sm_count)
{
    const char *f = getenv("ICW_SCAN_L");                       // tests force a length on short inputs
    if (f && *f) { const int v = atoi(f); if (v >= SCAN_L && v % SCAN_L == 0 && v <= 16384) return v; }
    const double chunks256 = (double)n_streams * (double)n_frames / SCAN_L;
    const double wave = (double)sm_count * (ICW_APPLY_THREADS / 4);    // pass-3 chunks resident at once
    // measured on C2 (i24, 691 M frames): 256 -> 5.0 + 10.3 ms (pass 1 + pass 3), 1024 -> 4.3 + 10.2,
    // 2048 -> 4.2 + 10.3, 4096 -> 4.2 + 10.5, 8192 -> 4.5 + 10.5; on the C5 shape (f32) 2048 is within 1 % of 4096
    if (chunks256 / 8 >= 8 * wave) return 8 * SCAN_L;
    if (chunks256 / 4 >= 8 * wave) return 4 * SCAN_L;
    return SCAN_L;
}

void scan_make_coef(int filter_no, bool baseline, double d0, int L, ModalCoef &mc, std::vector<double> &pw_table)
{
    memset(&mc, 0, sizeof mc);
    const int nm = ICW_HB_NMODES[filter_no];
    mc.nm = nm;
    mc.L = L;
    mc.real_last = ICW_HB_MODES[filter_no][nm - 1].is_real;
    mc.baseline = baseline;
    mc.d0 = d0;
    pw_table.assign((size_t)SCAN_CH * SCAN_NMAX * 2, 0.0);
    for (int m = 0; m < nm; ++m) {
        const icw_hb_mode &md = ICW_HB_MODES[filter_no][m];
        cld p = { (ld)md.p_re[0] + (ld)md.p_re[1], (ld)md.p_im[0] + (ld)md.p_im[1] };
        cld r = { (ld)md.r_re[0] + (ld)md.r_re[1], (ld)md.r_im[0] + (ld)md.r_im[1] };
        const ld w = md.is_real ? 1.0L : 2.0L;          // a pair contributes r s + conj(r s) = 2 Re(r s)
        cld p2 = cmul(p, p);
        ld n2 = p.re * p.re + p.im * p.im;
        cld pinv = { p.re / n2, -p.im / n2 };
        cld rp = cmul(r, p);
        cld pl = cpowi(p, L);
        cld pt = cpowi(p, (long long)L * SCAN_CH);
        mc.p_re[m] = (double)p.re;       mc.p_im[m] = (double)p.im;
        mc.p2_re[m] = (double)p2.re;     mc.p2_im[m] = (double)p2.im;
        mc.pinv_re[m] = (double)pinv.re; mc.pinv_im[m] = (double)pinv.im;
        mc.c_re[m] = (double)(w * r.re);   mc.c_im[m] = (double)(-w * r.im);
        mc.cp_re[m] = (double)(w * rp.re); mc.cp_im[m] = (double)(-w * rp.im);
        mc.k[0][m] = -mc.p2_re[m];       mc.k[1][m] = -mc.p2_im[m];         // exact: sign and power-of-two scaling
        mc.k[2][m] = -2.0 * mc.c_re[m];  mc.k[3][m] = -2.0 * mc.c_im[m];
        mc.k[4][m] = 2.0 * mc.cp_re[m];  mc.k[5][m] = 2.0 * mc.cp_im[m];
        mc.pl_re[m] = (double)pl.re;     mc.pl_im[m] = (double)pl.im;
        mc.pt_re[m] = (double)pt.re;     mc.pt_im[m] = (double)pt.im;
        cld pj = pl;
        for (int j = 0; j < 16; ++j) {                  // p^(L*(j+1))
            mc.plp_re[j][m] = (double)pj.re; mc.plp_im[j][m] = (double)pj.im;
            pj = cmul(pj, pl);
        }
        cld acc = { 1.0L, 0.0L };
        for (int j = 0; j < SCAN_CH; ++j) {             // p^(L*j)
            pw_table[((size_t)j * SCAN_NMAX + m) * 2] = (double)acc.re;
            pw_table[((size_t)j * SCAN_NMAX + m) * 2 + 1] = (double)acc.im;
            acc = cmul(acc, pl);
        }
    }
    // pass 1 wants a chunk's END state from zero: a mode whose pole has radius r has forgotten an input after
    // log(1e-18)/log(r) samples, so it may start that many samples before the chunk end.  Ranks: the slowest
    // mode first; the table lists conjugate pairs by ascending radius and the real pole (the fastest) last.
    {
        const bool rl = mc.real_last != 0;
        int prev = 0;
        for (int r = 0; r < nm; ++r) {
            const int m = rl ? (r == nm - 1 ? nm - 1 : nm - 2 - r) : nm - 1 - r;
            const double rad = std::sqrt(mc.p_re[m] * mc.p_re[m] + mc.p_im[m] * mc.p_im[m]);
            const double forget = rad > 0.0 && rad < 1.0 ? std::log(1e-18) / std::log(rad) : 1e30;   // samples
            int j = 0;
            if (forget < (double)L) j = (int)(((double)L - forget) / 2.0) - 2;       // steps of two samples
            j = (j < 0 || r == 0) ? 0 : (j / 4) * 4;                                // the slowest mode walks the whole chunk
            if (j < prev) j = prev;                                                 // ranks join in order
            mc.join[r] = j;
            prev = j;
        }
        for (int r = nm; r <= SCAN_NMAX; ++r) mc.join[r] = L / 2;
    }
}

// The one-kernel path's chunk tables (icw_sfused.h: SfTab): every power of q = -p^2 a 16-frame chunk needs, the
// functionals of the two outputs applied to them, and the chunk's own impulse-response taps -- long double from the
// double-double poles and residues, rounded once.
void sfused_make_tables(int filter_no, bool baseline, double d0, SfTab &tb)
{
    memset(&tb, 0, sizeof tb);
    const int nm = ICW_HB_NMODES[filter_no];
    tb.nm = nm;
    tb.real_last = ICW_HB_MODES[filter_no][nm - 1].is_real;
    tb.d0x2 = baseline ? 2.0 * d0 : 0.0;
    ld ha[SF_NI], hb[SF_NI];
    for (int k = 0; k < SF_NI; ++k) ha[k] = hb[k] = 0.0L;
    for (int m = 0; m < nm; ++m) {
        const icw_hb_mode &md = ICW_HB_MODES[filter_no][m];
        const cld p = { (ld)md.p_re[0] + (ld)md.p_re[1], (ld)md.p_im[0] + (ld)md.p_im[1] };
        const cld r = { (ld)md.r_re[0] + (ld)md.r_re[1], (ld)md.r_im[0] + (ld)md.r_im[1] };
        const ld w = md.is_real ? 1.0L : 2.0L;
        cld q = cmul(p, p);
        q.re = -q.re; q.im = -q.im;
        const cld rp = cmul(r, p);
        const ld n2 = p.re * p.re + p.im * p.im;
        // out1 = k4 S.re + k5 S.im (2 c p), out2 = k2 S.re + k3 S.im (-2 c): the constants of scan_make_coef, unrounded
        const ld k4 = 2.0L * w * rp.re, k5 = -2.0L * w * rp.im, k2 = -2.0L * w * r.re, k3 = 2.0L * w * r.im;
        cld qpow[SF_NI + 1];
        cld g = { 1.0L, 0.0L };
        for (int k = 0; k <= SF_NI; ++k) { qpow[k] = g; g = cmul(g, q); }
        for (int j = 0; j < SF_NI; ++j) {
            tb.q[j][m][0] = (double)-qpow[SF_NI - 1 - j].re;
            tb.q[j][m][1] = (double)-qpow[SF_NI - 1 - j].im;
        }
        tb.q8[m][0] = (double)qpow[SF_NI].re; tb.q8[m][1] = (double)qpow[SF_NI].im;
        tb.qt[m][0] = (double)q.re;           tb.qt[m][1] = (double)q.im;
        tb.p[m][0] = (double)p.re;            tb.p[m][1] = (double)p.im;
        tb.pinv[m][0] = (double)(p.re / n2);  tb.pinv[m][1] = (double)(-p.im / n2);
        for (int i = 0; i < SF_NI; ++i) {
            tb.ca[i][m][0] = (double)(k4 * qpow[i].re + k5 * qpow[i].im);
            tb.ca[i][m][1] = (double)(-k4 * qpow[i].im + k5 * qpow[i].re);
        }
        for (int k = 0; k <= SF_NI; ++k) {
            tb.cb[k][m][0] = (double)(k2 * qpow[k].re + k3 * qpow[k].im);
            tb.cb[k][m][1] = (double)(-k2 * qpow[k].im + k3 * qpow[k].re);
        }
        for (int k = 0; k < SF_NI; ++k) {
            ha[k] -= k4 * qpow[k].re + k5 * qpow[k].im;
            hb[k] -= k2 * qpow[k].re + k3 * qpow[k].im;
        }
    }
    for (int k = 0; k < SF_NI; ++k) { tb.ha[k] = (double)ha[k]; tb.hb[k] = (double)hb[k]; }
}

// ---------------------------------------------------------------------------------------------
// device
// ---------------------------------------------------------------------------------------------
// FMT is a compile-time constant: the format switch folds away inside the sample loops.
// A thread walks its filter's samples by POINTER (p advances two frames a step): forming the address
// from the frame index costs a 64-bit multiply-add chain per sample, a quarter of the loop's
// instructions before this form.
template <int FMT, bool FADE>
__device__ __forceinline__ double scan_sample_at(const DevChain &ch, const uint8_t *p, int64_t file_pos)
{
    double x = unpack_real(FMT, p, ch.aligned);
    if (FADE) {
        double g = fade_gain(ch, file_pos);
        if (g >= 0.0) x *= g;
    }
    return x;
}

// With only one or two warps per scheduler to hide a miss, ask for the line two ahead of the one
// being read (the clamp keeps the request inside the row); once every fourth step is enough, a
// 128-byte line holds five to sixteen steps.
__device__ __forceinline__ void scan_prefetch_at(const uint8_t *p, const uint8_t *row_last)
{
    const uint8_t *q = p + ICW_SCAN_PF_BYTES;
    if (q > row_last) q = row_last;
#ifdef ICW_SCAN_PF_L2
    asm volatile("prefetch.global.L2 [%0];" :: "l"(q));
#else
    asm volatile("prefetch.global.L1 [%0];" :: "l"(q));
#endif
}

// E layout: [stream][comp][chan][chunk], comp = (filter * SCAN_NMAX + mode) * 2 + {re, im}
__device__ __forceinline__ size_t e_index(int stream, int comp, int chan, int64_t chunk, int64_t n_chunks)
{
    return (((size_t)stream * (4 * SCAN_NMAX) + comp) * 2 + chan) * (size_t)n_chunks + (size_t)chunk;
}

// does any frame of [pos, pos + n) fall inside the file's fade-in or fade-out ramp?
__device__ __forceinline__ bool chunk_fades(const DevChain &ch, int64_t pos, int n)
{
    if ((ch.n_fade_in | ch.n_fade_out) == 0) return false;
    return pos < ch.n_fade_in || pos + n > ch.n_samples - ch.n_fade_out;
}

// pass-1 inner loop: L / 2 inputs of one filter from a zero state, sign-free form.  The slowest mode runs
// over the whole chunk; the mode of rank r joins at step mc.join[r] with a zero state (what it would hold by
// then has decayed below 1e-18 by the chunk end).  ACT = modes running in the current phase.
template <int NM, bool RL> __device__ __forceinline__ constexpr int mode_rank(int m)
{
    return RL ? (m == NM - 1 ? NM - 1 : NM - 2 - m) : NM - 1 - m;
}
template <int ACT, int NM, bool RL>
__device__ __forceinline__ void local_advance(Cx (&s)[NM], const double (&kpr)[NM], const double (&kpi)[NM], double xin)
{
#pragma unroll
    for (int m = 0; m < NM; ++m) {
        if (mode_rank<NM, RL>(m) >= ACT) continue;
        if (RL && m == NM - 1) re_step(s[m], kpr[m], -xin);
        else cx_step(s[m], kpr[m], kpi[m], -xin);
    }
}
template <int ACT, int NM, bool RL, int FMT, bool FADE>
__device__ __forceinline__ void local_phase(Cx (&s)[NM], const double (&kpr)[NM], const double (&kpi)[NM], const DevChain &ch,
                                            const uint8_t *&p, int step_bytes, int64_t &fpos, const uint8_t *row_last,
                                            double &x, int k0, int k1)
{
#pragma unroll LOCAL_UNROLL
    for (int k = k0; k < k1; ++k) {                             // the next input is fetched one step ahead
        p += step_bytes;
        if (FADE) fpos += 2;
        const double xn = scan_sample_at<FMT, FADE>(ch, p, fpos);
        if ((k & (LOCAL_UNROLL - 1)) == 0) scan_prefetch_at(p, row_last);
        local_advance<ACT, NM, RL>(s, kpr, kpi, x);
        x = xn;
    }
}
template <int ACT, int NM, bool RL, int FMT, bool FADE>
struct LocalPhases {
    static __device__ __forceinline__ void run(Cx (&s)[NM], const double (&kpr)[NM], const double (&kpi)[NM], const DevChain &ch,
                                               const ModalCoef &mc, const uint8_t *&p, int step_bytes, int64_t &fpos,
                                               const uint8_t *row_last, double &x, int last)
    {
        // phase ACT: steps [join[ACT-1], join[ACT]) -- never the chunk's last step, which has nothing to fetch
        const int k0 = mc.join[ACT - 1], k1 = mc.join[ACT] < last ? mc.join[ACT] : last;
        local_phase<ACT, NM, RL, FMT, FADE>(s, kpr, kpi, ch, p, step_bytes, fpos, row_last, x, k0, k1);
        if (ACT < NM) LocalPhases<(ACT < NM ? ACT + 1 : NM), NM, RL, FMT, FADE>::run(s, kpr, kpi, ch, mc, p, step_bytes, fpos, row_last, x, last);
    }
};
template <int NM, bool RL, int FMT, bool FADE>
__device__ __forceinline__ void local_run(Cx (&s)[NM], const double (&kpr)[NM], const double (&kpi)[NM], const DevChain &ch,
                                          const ModalCoef &mc, const uint8_t *row, int64_t f0, int chan_off, int64_t pos0,
                                          const uint8_t *row_last)
{
    const uint8_t *p = row + f0 * ch.frame_bytes + chan_off;
    const int step_bytes = 2 * ch.frame_bytes;
    int64_t fpos = pos0 + f0;
    double x = scan_sample_at<FMT, FADE>(ch, p, fpos);
    LocalPhases<1, NM, RL, FMT, FADE>::run(s, kpr, kpi, ch, mc, p, step_bytes, fpos, row_last, x, mc.L / 2 - 1);
    local_advance<NM, NM, RL>(s, kpr, kpi, x);                  // the chunk's last input: every mode is in by now
}

// pass-3 inner loop: one filter over one chunk from its true state S~, writing its half of every
// frame.  The filter's inputs sit at frames off, off + 2, ...; the frame of an input gets
// sum(2cp S~) (+ 2 d0 x) in its re slot, the frame after it sum(-2c S~') in its im slot.
template <int NM, bool RL, int FMT, bool FADE>
__device__ __forceinline__ int apply_run(Cx (&S)[NM], const ModalCoef &mc, const double (*kshared)[SCAN_NMAX], bool direct, double d0x2, const DevChain &ch,
                                         const uint8_t *row, int64_t f0, int off, int len, int chan_off, int64_t pos0,
                                         const uint8_t *row_last, double *__restrict__ dst)
{
    double kpr[NM], kpi[NM], kcr[NM], kci[NM], kcpr[NM], kcpi[NM];
#pragma unroll
    for (int m = 0; m < NM; ++m) {
        kpr[m] = kconst<APPLY_UMASK>(mc, kshared, K_PR, m);   kpi[m] = kconst<APPLY_UMASK>(mc, kshared, K_PI, m);
        kcr[m] = kconst<APPLY_UMASK>(mc, kshared, K_CR, m);   kci[m] = kconst<APPLY_UMASK>(mc, kshared, K_CI, m);
        kcpr[m] = kconst<APPLY_UMASK>(mc, kshared, K_CPR, m); kcpi[m] = kconst<APPLY_UMASK>(mc, kshared, K_CPI, m);
    }
    if (off) {                                                  // frame 0 follows an input of the previous chunk
        double y2 = 0.0;
#pragma unroll
        for (int m = 0; m < NM; ++m) y2 = fma(kcr[m], S[m].re, fma(kci[m], S[m].im, y2));
        dst[1] = y2;
    }
    // this filter's inputs in the chunk: n_s of them, the first n_full followed by a frame of the chunk
    const int n_s = len > off ? (len - off + 1) >> 1 : 0, n_full = len > off ? (len - off) >> 1 : 0;
    const uint8_t *p = row + (f0 + off) * ch.frame_bytes + chan_off;
    const int step_bytes = 2 * ch.frame_bytes;
    int64_t fpos = pos0 + f0 + off;
    double *dp = dst + (size_t)off * 4;
    double x = 0.0, xn = 0.0;
    if (n_s > 0) x = scan_sample_at<FMT, FADE>(ch, p, fpos);
    if (n_s > 1) { p += step_bytes; fpos += 2; xn = scan_sample_at<FMT, FADE>(ch, p, fpos); }
    auto one = [&](double xin) {                                // one input: both outputs it shapes
        double y1 = direct ? d0x2 * xin : 0.0, y2 = 0.0;
#pragma unroll
        for (int m = 0; m < NM; ++m) {
            if (RL && m == NM - 1) {                            // the real pole: im == 0 and its weights are 0
                y1 = fma(kcpr[m], S[m].re, y1);
                re_step(S[m], kpr[m], -xin);
                y2 = fma(kcr[m], S[m].re, y2);
            } else {
                y1 = fma(kcpr[m], S[m].re, fma(kcpi[m], S[m].im, y1));
                cx_step(S[m], kpr[m], kpi[m], -xin);
                y2 = fma(kcr[m], S[m].re, fma(kci[m], S[m].im, y2));
            }
        }
        dp[0] = y1;
        dp[5] = y2;
        dp += 8;
    };
    int i = 0;
#pragma unroll APPLY_UNROLL
    for (; i + 2 < n_s; ++i) {                                  // the input after next is fetched two steps ahead
        p += step_bytes;
        if (FADE) fpos += 2;
        const double xnn = scan_sample_at<FMT, FADE>(ch, p, fpos);
        if ((i & 1) == 0) scan_prefetch_at(p, row_last);
        one(x);
        x = xn; xn = xnn;
    }
    for (; i < n_full; ++i) { one(x); x = xn; }                 // the last one or two, nothing left to fetch
    if (n_s > n_full) {                                         // an input on the call's very last frame
        double y1 = direct ? d0x2 * x : 0.0;
#pragma unroll
        for (int m = 0; m < NM; ++m) {
            y1 = fma(kcpr[m], S[m].re, fma(kcpi[m], S[m].im, y1));
            cx_step(S[m], kpr[m], kpi[m], -x);
        }
        dp[0] = y1;
    }
    return n_s;
}

// Thread = one recurrence set: (chunk, channel, filter I/Q) with the filter's NM modal states in
// registers -- the same decomposition as the exact kernel's "one thread per half-band filter".
// The filter fed on the first sample of every pair (phase parity of the chunk start) is "X", the
// other "Y"; a thread only ever touches the samples that feed its own filter.
//
// pass 1 + 2a: end state of every full chunk started from zero, then -- inside the CTA, which is
// exactly one tile of 128 chunks -- the carry scan over chunks: warp shuffles (8 chunks per warp),
// one shared-memory hop across the 16 warps.  Writes, per chunk, its carry-in from the tile start
// (exclusive prefix) and, per tile, the tile's end state.
template <int NM, bool RL, int FMT>
__global__ void __launch_bounds__(512, 1)
scan_local_kernel(const __grid_constant__ ModalCoef mc, const __grid_constant__ DevChain ch,
                  const DevStream *__restrict__ streams, int64_t n_frames, int64_t n_chunks, int64_t n_tiles,
                  const uint8_t *__restrict__ in, size_t in_stride, double *__restrict__ E,
                  double *__restrict__ Tend /* [stream][tile][chan][filter][mode][2] */)
{
    __shared__ double wtot[16][2][2][SCAN_NMAX][2];     // inclusive total of each warp's 8 chunks
    __shared__ double wcar[16][2][2][SCAN_NMAX][2];     // carry into each warp from the tile start
    __shared__ double kshared[6][SCAN_NMAX];
    stage_constants(kshared, mc);
    const int stream = blockIdx.y;
    const int64_t tile = blockIdx.x;
    const int filt = threadIdx.x & 1, chan = (threadIdx.x >> 1) & 1;
    const int cl = threadIdx.x >> 2;                    // chunk inside the tile, 0..127
    const int64_t chunk = tile * SCAN_CH + cl;
    const int warp = threadIdx.x >> 5, cw = cl & 7;     // chunk inside the warp
    const bool have = chunk < n_chunks - 1;             // the call's last chunk has no end state to hand on
    const DevStream &st = streams[stream];
    const unsigned q0 = st.quad[chan];
    const bool is_x = filt == (int)(q0 & 1);            // fed on the first sample of a pair

    Cx s[NM];
#pragma unroll
    for (int m = 0; m < NM; ++m) s[m].re = s[m].im = 0.0;
    if (have) {
        const uint8_t *row = in + (size_t)stream * in_stride;
        const int chan_off = (ch.n_channels > 1 ? chan : 0) * ch.chan_bytes;
        const uint8_t *row_last = row + n_frames * ch.frame_bytes - 1;
        const int64_t f0 = chunk * mc.L + (is_x ? 0 : 1);       // this filter's first input sample
        const unsigned qf = (q0 + (is_x ? 0u : 1u)) & 3u;       // its mixer phase (chunks start on a multiple of 4)
        double kpr[NM], kpi[NM];
#pragma unroll
        for (int m = 0; m < NM; ++m) { kpr[m] = kconst<LOCAL_UMASK>(mc, kshared, K_PR, m); kpi[m] = kconst<LOCAL_UMASK>(mc, kshared, K_PI, m); }
        if (chunk_fades(ch, st.pos + chunk * mc.L, mc.L))
            local_run<NM, RL, FMT, true>(s, kpr, kpi, ch, mc, row, f0, chan_off, st.pos, row_last);
        else
            local_run<NM, RL, FMT, false>(s, kpr, kpi, ch, mc, row, f0, chan_off, st.pos, row_last);
        // S~ -> S: L / 2 inputs is an even count, the sign is the first input's
        const double sg = mixer_sign(filt, qf);
#pragma unroll
        for (int m = 0; m < NM; ++m) { s[m].re *= sg; s[m].im *= sg; }
        // state after the chunk's last sample (a Y-input sample): X has idled one sample since its input
        if (is_x) {
#pragma unroll
            for (int m = 0; m < NM; ++m) s[m] = cx_mul(mc.p_re[m], mc.p_im[m], s[m]);
        }
    }

    // inclusive scan over the warp's 8 chunks (lanes of one recurrence are 4 apart)
#pragma unroll
    for (int d = 1; d < 8; d <<= 1) {
#pragma unroll
        for (int m = 0; m < NM; ++m) {
            const double ore = __shfl_up_sync(0xffffffffu, s[m].re, 4 * d);
            const double oim = __shfl_up_sync(0xffffffffu, s[m].im, 4 * d);
            if (cw >= d) {
                Cx o; o.re = ore; o.im = oim;
                Cx w = cx_mul(mc.plp_re[d - 1][m], mc.plp_im[d - 1][m], o);         // p^(L*d) * earlier
                s[m].re += w.re; s[m].im += w.im;
            }
        }
    }
    if (cw == 7) {
#pragma unroll
        for (int m = 0; m < NM; ++m) { wtot[warp][chan][filt][m][0] = s[m].re; wtot[warp][chan][filt][m][1] = s[m].im; }
    }
    __syncthreads();
    // carry into each warp: sequential over the 16 warps, one thread per (chan, filter, mode)
    if (threadIdx.x < 2 * 2 * NM) {
        int r = threadIdx.x;
        const int m = r % NM; r /= NM;
        const int f = r & 1, c = r >> 1;
        Cx acc; acc.re = acc.im = 0.0;
        for (int w = 0; w < 16; ++w) {
            wcar[w][c][f][m][0] = acc.re; wcar[w][c][f][m][1] = acc.im;
            // acc <- p^(8 L) * acc + total(w)
            const double nr = fma(mc.plp_re[7][m], acc.re, fma(-mc.plp_im[7][m], acc.im, wtot[w][c][f][m][0]));
            const double ni = fma(mc.plp_re[7][m], acc.im, fma(mc.plp_im[7][m], acc.re, wtot[w][c][f][m][1]));
            acc.re = nr; acc.im = ni;
        }
    }
    __syncthreads();
    // inclusive prefix from the tile start, then shift by one chunk for the carry-IN of this chunk
#pragma unroll
    for (int m = 0; m < NM; ++m) {
        Cx c; c.re = wcar[warp][chan][filt][m][0]; c.im = wcar[warp][chan][filt][m][1];
        Cx w = cx_mul(mc.plp_re[cw][m], mc.plp_im[cw][m], c);                       // p^(L*(cw+1)) * warp carry
        const double inc_re = s[m].re + w.re, inc_im = s[m].im + w.im;
        double ex_re = __shfl_up_sync(0xffffffffu, inc_re, 4);
        double ex_im = __shfl_up_sync(0xffffffffu, inc_im, 4);
        if (cw == 0) { ex_re = c.re; ex_im = c.im; }
        if (chunk < n_chunks) {
            const int comp = (filt * SCAN_NMAX + m) * 2;
            E[e_index(stream, comp, chan, chunk, n_chunks)] = ex_re;
            E[e_index(stream, comp + 1, chan, chunk, n_chunks)] = ex_im;
        }
        if (cl == SCAN_CH - 1) {
            double *t = Tend + ((((size_t)stream * n_tiles + tile) * 2 + chan) * 2 + filt) * (SCAN_NMAX * 2) + m * 2;
            t[0] = inc_re; t[1] = inc_im;
        }
    }
}

// pass 2b: carry into every tile.  The filters forget: |p|^(tile) <= 1e-3, so eight tiles back is
// below 1e-24 and the sum over the previous SCAN_W tile end states (Horner in p^tile) is exact to
// rounding.  Index -1 stands for the stream's state before this call.
constexpr int SCAN_W = 8;
__global__ void __launch_bounds__(128)
scan_tile_carry_kernel(const __grid_constant__ ModalCoef mc, const DevStream *__restrict__ streams, int64_t n_tiles,
                       const double *__restrict__ Tend, double *__restrict__ Tin)
{
    const int stream = blockIdx.y;
    const int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int per_tile = 2 * 2 * mc.nm;
    if (id >= n_tiles * per_tile) return;
    const int64_t tile = id / per_tile;
    int r = (int)(id % per_tile);
    const int m = r % mc.nm; r /= mc.nm;
    const int f = r & 1, chan = r >> 1;
    const double tr = mc.pt_re[m], ti = mc.pt_im[m];
    Cx acc; acc.re = acc.im = 0.0;
    for (int i = SCAN_W; i >= 1; --i) {
        const int64_t src = tile - i;
        Cx e; e.re = e.im = 0.0;
        if (src >= 0) {
            const double *t = Tend + ((((size_t)stream * n_tiles + src) * 2 + chan) * 2 + f) * (SCAN_NMAX * 2) + m * 2;
            e.re = t[0]; e.im = t[1];
        } else if (src == -1) {
            e.re = streams[stream].hb[chan][f][2 * m];
            e.im = streams[stream].hb[chan][f][2 * m + 1];
        }
        const double nr = fma(tr, acc.re, fma(-ti, acc.im, e.re));
        const double ni = fma(tr, acc.im, fma(ti, acc.re, e.im));
        acc.re = nr; acc.im = ni;
    }
    double *o = Tin + ((((size_t)stream * n_tiles + tile) * 2 + chan) * 2 + f) * (SCAN_NMAX * 2) + m * 2;
    o[0] = acc.re; o[1] = acc.im;
}

// pass 3: every chunk again, from its true initial state, producing the analytic signal.
// Each thread writes its own filter's half of every frame: re or im, alternating with the phase.
template <int NM, bool RL, int FMT>
__global__ void __launch_bounds__(ICW_APPLY_THREADS, 1)
scan_apply_kernel(const __grid_constant__ ModalCoef mc, const __grid_constant__ DevChain ch,
                  DevStream *__restrict__ streams, int64_t n_frames, int64_t n_chunks, int64_t n_tiles,
                  const uint8_t *__restrict__ in, size_t in_stride,
                  const double *__restrict__ E, const double *__restrict__ Tin, const double *__restrict__ pw,
                  double *__restrict__ analytic /* [stream][frame][4] */)
{
    __shared__ double kshared[6][SCAN_NMAX];
    stage_constants(kshared, mc);
    const int stream = blockIdx.y;
    const int filt = threadIdx.x & 1, chan = (threadIdx.x >> 1) & 1;
    const int64_t chunk = (int64_t)blockIdx.x * (blockDim.x / 4) + (threadIdx.x >> 2);
    if (chunk >= n_chunks) return;
    DevStream &st = streams[stream];
    const unsigned q0 = st.quad[chan];
    const bool is_x = filt == (int)(q0 & 1);
    const uint8_t *row = in + (size_t)stream * in_stride;
    const int chan_off = (ch.n_channels > 1 ? chan : 0) * ch.chan_bytes;
    const int64_t f0 = chunk * mc.L;
    const int len = (int)((n_frames - f0 < mc.L) ? n_frames - f0 : mc.L);
    const int64_t tile = chunk / SCAN_CH;
    const int jl = (int)(chunk % SCAN_CH);
    double *dst = analytic + ((size_t)stream * (size_t)n_frames + (size_t)f0) * 4 + chan * 2;

    // carried state after the sample before this chunk: prefix inside the tile + p^(L*j) * tile carry
    Cx S[NM];
#pragma unroll
    for (int m = 0; m < NM; ++m) {
        const int comp = (filt * SCAN_NMAX + m) * 2;
        const double *t = Tin + ((((size_t)stream * n_tiles + tile) * 2 + chan) * 2 + filt) * (SCAN_NMAX * 2) + m * 2;
        Cx c; c.re = t[0]; c.im = t[1];
        const double wr = pw[((size_t)jl * SCAN_NMAX + m) * 2], wi = pw[((size_t)jl * SCAN_NMAX + m) * 2 + 1];
        Cx cw = cx_mul(wr, wi, c);
        Cx s0;
        s0.re = E[e_index(stream, comp, chan, chunk, n_chunks)] + cw.re;
        s0.im = E[e_index(stream, comp + 1, chan, chunk, n_chunks)] + cw.im;
        // X last saw input two samples ago: step it back one sample (exact algebra: S = s / p)
        S[m] = is_x ? cx_mul(mc.pinv_re[m], mc.pinv_im[m], s0) : s0;
    }

    const uint8_t *row_last = row + n_frames * ch.frame_bytes - 1;
    // S -> S~ with the sign of this filter's first input of the chunk (chunks start on a multiple of 4)
    const int off = is_x ? 0 : 1;
    const double sg = mixer_sign(filt, (q0 + (unsigned)off) & 3u);
#pragma unroll
    for (int m = 0; m < NM; ++m) { S[m].re *= sg; S[m].im *= sg; }
    const bool direct = mc.baseline != 0;
    const double d0x2 = 2.0 * mc.d0;
    int n_in;
    if (chunk_fades(ch, st.pos + f0, len))
        n_in = apply_run<NM, RL, FMT, true>(S, mc, kshared, direct, d0x2, ch, row, f0, off, len, chan_off, st.pos, row_last, dst);
    else
        n_in = apply_run<NM, RL, FMT, false>(S, mc, kshared, direct, d0x2, ch, row, f0, off, len, chan_off, st.pos, row_last, dst);
    if (chunk == n_chunks - 1) {
        // state after the call's last sample: undo the sign (it flips with every input), and a filter
        // whose last input was not the last frame has idled one sample since
        const double se = (n_in & 1) ? -sg : sg;
        const bool idle = ((len - 1 - off) & 1) != 0;
#pragma unroll
        for (int m = 0; m < NM; ++m) {
            Cx v; v.re = S[m].re * se; v.im = S[m].im * se;
            Cx e = idle ? cx_mul(mc.p_re[m], mc.p_im[m], v) : v;
            st.hb[chan][filt][2 * m] = e.re; st.hb[chan][filt][2 * m + 1] = e.im;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// launch sequence
// ---------------------------------------------------------------------------------------------
size_t scan_scratch_doubles(int n_streams, int64_t n_frames, int L)
{
    const int64_t n_chunks = (n_frames + L - 1) / L;
    const int64_t n_tiles = (n_chunks + SCAN_CH - 1) / SCAN_CH;
    const size_t e = (size_t)n_streams * (4 * SCAN_NMAX) * 2 * (size_t)n_chunks;
    const size_t t = (size_t)n_streams * (size_t)n_tiles * 2 * 2 * (SCAN_NMAX * 2);
    return e + 2 * t;
}

template <int NM, bool RL, int FMT>
static cudaError_t scan_launch_nf(const ModalCoef &mc, const DevChain &ch, DevStream *streams, int n_streams,
                                  int64_t n_frames, const uint8_t *in, size_t in_stride, const double *pw,
                                  double *scratch, double *analytic, cudaStream_t s, int *launches,
                                  cudaEvent_t mid_end, cudaEvent_t mid_start)
{
    const int64_t n_chunks = (n_frames + mc.L - 1) / mc.L;
    const int64_t n_tiles = (n_chunks + SCAN_CH - 1) / SCAN_CH;
    double *E = scratch;
    double *Tend = E + (size_t)n_streams * (4 * SCAN_NMAX) * 2 * (size_t)n_chunks;
    double *Tin = Tend + (size_t)n_streams * (size_t)n_tiles * 2 * 2 * (SCAN_NMAX * 2);
    const unsigned cgrid = (unsigned)n_tiles;                   // one CTA = one tile: 128 chunks x 2 channels x 2 filters
    scan_local_kernel<NM, RL, FMT><<<dim3(cgrid, n_streams), 512, 0, s>>>(mc, ch, streams, n_frames, n_chunks, n_tiles, in, in_stride, E, Tend);
    const int64_t items = n_tiles * 2 * 2 * mc.nm;
    const unsigned tgrid = (unsigned)((items + 127) / 128);
    scan_tile_carry_kernel<<<dim3(tgrid, n_streams), 128, 0, s>>>(mc, streams, n_tiles, Tend, Tin);
    if (mid_end) { cudaEventRecord(mid_end, s); cudaEventRecord(mid_start, s); }
    const unsigned agrid = (unsigned)((n_chunks + ICW_APPLY_THREADS / 4 - 1) / (ICW_APPLY_THREADS / 4));
    scan_apply_kernel<NM, RL, FMT><<<dim3(agrid, n_streams), ICW_APPLY_THREADS, 0, s>>>(mc, ch, streams, n_frames, n_chunks, n_tiles, in, in_stride,
                                                                    E, Tin, pw, analytic);
    *launches += 3;
    return cudaGetLastError();
}

template <int NM, bool RL>
static cudaError_t scan_launch_nm(const ModalCoef &mc, const DevChain &ch, DevStream *streams, int n_streams,
                                  int64_t n_frames, const uint8_t *in, size_t in_stride, const double *pw,
                                  double *scratch, double *analytic, cudaStream_t s, int *launches,
                                  cudaEvent_t mid_end, cudaEvent_t mid_start)
{
#define ICW_SCAN_FMT(F) case F: return scan_launch_nf<NM, RL, F>(mc, ch, streams, n_streams, n_frames, in, in_stride, pw, scratch, analytic, s, launches, mid_end, mid_start)
    switch (ch.fmt) {
        ICW_SCAN_FMT(ICW_FMT_WAV_U8);
        ICW_SCAN_FMT(ICW_FMT_WAV_I16);
        ICW_SCAN_FMT(ICW_FMT_WAV_I24);
        ICW_SCAN_FMT(ICW_FMT_WAV_I32);
        ICW_SCAN_FMT(ICW_FMT_WAV_F32);
        ICW_SCAN_FMT(ICW_FMT_INTERNAL_F64);
    default: return cudaErrorInvalidValue;
    }
#undef ICW_SCAN_FMT
}

cudaError_t launch_hb_scan(const ModalCoef &mc, const DevChain &ch, DevStream *streams, int n_streams,
                           int64_t n_frames, const uint8_t *in, size_t in_stride, const double *pw,
                           double *scratch, double *analytic, cudaStream_t s, int *launches, cudaEvent_t mid_end, cudaEvent_t mid_start)
{
    switch (mc.nm) {
    // odd orders end with a real pole (15 -> 7 pairs + 1, 19 -> 9 + 1): its mode runs in scalar form
    case 8:  return scan_launch_nm<8, true>(mc, ch, streams, n_streams, n_frames, in, in_stride, pw, scratch, analytic, s, launches, mid_end, mid_start);
    case 9:  return scan_launch_nm<9, false>(mc, ch, streams, n_streams, n_frames, in, in_stride, pw, scratch, analytic, s, launches, mid_end, mid_start);
    case 10: return mc.real_last ? scan_launch_nm<10, true>(mc, ch, streams, n_streams, n_frames, in, in_stride, pw, scratch, analytic, s, launches, mid_end, mid_start)
                                : scan_launch_nm<10, false>(mc, ch, streams, n_streams, n_frames, in, in_stride, pw, scratch, analytic, s, launches, mid_end, mid_start);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace icw

// host-only hook for the CPU tests: the chunk length the scan picks for a launch group
extern "C" int icw_host_scan_chunk_len(int n_streams, int64_t n_frames, int sm_count)
{
    return icw::scan_chunk_len(n_streams, n_frames, sm_count);
}
