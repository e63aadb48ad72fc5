// icw_sfused.cu -- scan mode in ONE kernel: file bytes -> Hilbert (modal scan) -> oscillator -> DSP list -> dither ->
// quantise -> PCM, with nothing but the input bytes and the PCM bytes crossing HBM.
//
// The three-kernel scan path (icw_scan.cu passes 1..3, then chain_mt_kernel) writes the analytic signal to HBM and
// reads it back: 64 of the 84 bytes per frame the C2 workload moved, and two half-idle kernels back to back -- pass 3
// is bound by the FP64 pipe (issue slots 46 % busy), the pointwise kernel by instruction issue (FP64 pipe 29 % busy).
// Here both run on the same SM at the same time, in different warps:
//
//   one CTA per SM = one UNIT: a contiguous run of the stream's frames (the generators' jump-ahead unit, icw_mt.cu,
//   when there is dither -- the draw index is a closed form of the frame index -- an even split otherwise), walked in
//   RANGES of 2304 frames = 64 chunks x 36 frames.
//
//   warps 0-7   SCAN      thread = (chunk, channel, I/Q filter).  Per range: input bytes arrive by one cp.async.bulk
//                         (TMA, issued a range ahead by one thread, mbarrier-tracked) and are unpacked once to doubles
//                         in shared memory; LOCAL pass: every chunk's end state from zero; a serial carry over the 64
//                         chunks from the CTA's own running state (40 threads, shared memory); APPLY pass: every chunk
//                         again from its true state, its half of every analytic frame written into a ring of SLICES
//                         (6 frames of each chunk) in shared memory.
//   warps 8-19  POINTWISE thread = one frame of a slice: oscillator, DSP list, dither, quantiser, PCM store -- the code
//                         of icw_frame.cuh.  While the scan warps are in their LOCAL pass these warps regenerate the
//                         range's MT19937 words for both channels' generators (icw_mtdev.cuh, window form).
//
// Slices are handed over with named barriers (bar.arrive / bar.sync, one FULL and one EMPTY barrier per ring slot);
// the register file is split with setmaxnreg (scan warps 144, pointwise warps 64: the pool is what the CTA's own
// warps give back, so the two must add up to the 640 x 96 registers of the launch).  A unit's filter state at its
// first frame comes from a WARM-UP: the LOCAL pass alone over the frames before it (the filters forget: |p|^warm is
// below 1e-19; the same fact the multi-GPU hand-off uses) -- so there is no pass 1 / pass 2 over the whole stream,
// no carry arrays in HBM and no second read of the input.
//
// Numerics: the same modal recurrences as icw_scan.cu (sign-free two-sample steps, the same constants); only the
// chunking differs, which moves results at the 1e-16 level (tests/test_gpu_sfused.py holds it to the binary128 truth
// at 1e-12 like the other scan tests, and to the three-kernel path within an LSB).
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "icw_dev.cuh"
#include "icw_frame.cuh"
#include "icw_kernels.h"
#include "icw_mtdev.cuh"
#include "icw_scan_dev.cuh"
#include "icw_sfused.h"

namespace icw {

constexpr int SF_SCAN_THREADS = SF_CH * 4;              // 256
constexpr int SF_PW_THREADS = 384;
constexpr int SF_THREADS = SF_SCAN_THREADS + SF_PW_THREADS;
constexpr int SF_SL = 6;                                // frames of a chunk per slice: 64 chunks x 6 = one frame per pointwise thread
constexpr int SF_NSLICE = SF_LC / SF_SL;                // 6 slices per range
constexpr int SF_NSLOT = 4;                             // ring depth
constexpr int SF_PLANE = SF_CH * SF_SL * 16 + 32;       // one channel's (re, im) pairs of a slice, padded
constexpr int SF_SLOT = 2 * SF_PLANE;
constexpr int SF_MT_PASS = SF_PW_THREADS / 2;           // words of each generator per pass (<= 227)
static_assert(SF_LC % 4 == 0 && SF_LC % SF_SL == 0 && SF_SL % 2 == 0, "chunk geometry");
static_assert(SF_CH * SF_SL == SF_PW_THREADS, "one frame per pointwise thread and slice");
static_assert(SF_NSLOT * SF_SLOT >= SF_CH * 4 * 2 * SCAN_NMAX * 8, "the chunk end states borrow the slice ring");
static_assert(SF_MT_PASS <= 227, "MT19937 reaches back 227 words");

// named barriers (id 0 is __syncthreads)
enum { NB_SCAN = 1, NB_PW = 2, NB_FULL = 3, NB_EMPTY = 3 + SF_NSLOT };

__device__ __forceinline__ void nb_sync(int id, int n) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void nb_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" :: "r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" :: "r"(bar), "r"(parity) : "memory");
}

// shared-memory map (bytes); the word buffers' size depends on the dither type
struct SfSmem {
    static constexpr int BAR = 0;                                   // one mbarrier (input staging)
    static constexpr int CARRY = 64;                                // [4][SCAN_NMAX][2] doubles: state after the last frame done
    static constexpr int KSH = CARRY + 4 * SCAN_NMAX * 2 * 8;       // [6][SCAN_NMAX] doubles: sample-loop constants
    static constexpr int STAGE = KSH + 6 * SCAN_NMAX * 8;           // the range's input bytes as they lie in the file (+ alignment slack)
    static constexpr int XD = STAGE + SF_R * 8 + 32;                // [SF_R][2] doubles: the same, unpacked (mono duplicated)
    static constexpr int SLOTS = XD + SF_R * 2 * 8;                 // slice ring; between LOCAL and APPLY: the chunk end states
    static constexpr int WORDS = SLOTS + SF_NSLOT * SF_SLOT;        // two generators' untempered words
};
__host__ __device__ constexpr int sf_gen_words(int wps) { return wps ? 3 * ICW_MT_N + SF_R * wps : 0; }   // history + skipped head + range + tail block
size_t sfused_smem_bytes(int wps) { return (size_t)SfSmem::WORDS + 2 * (size_t)sf_gen_words(wps) * 4; }

// ------------------------------------------------------------------------------------------------------------------
// SCAN side
// ------------------------------------------------------------------------------------------------------------------
// chunk end state from zero over the chunk's inputs of one filter (sign-free form, every mode from the first input)
template <int NM, bool RL>
__device__ __forceinline__ void sf_local(Cx (&s)[NM], const double (&kpr)[NM], const double (&kpi)[NM], const double *xp, int n_in)
{
#pragma unroll
    for (int m = 0; m < NM; ++m) s[m].re = s[m].im = 0.0;
    double x = n_in > 0 ? xp[0] : 0.0;
#pragma unroll 2
    for (int i = 0; i < n_in; ++i) {
        const double xn = xp[(i + 1 < n_in ? i + 1 : i) * 4];
#pragma unroll
        for (int m = 0; m < NM; ++m) {
            if (RL && m == NM - 1) re_step(s[m], kpr[m], -x);
            else cx_step(s[m], kpr[m], kpi[m], -x);
        }
        x = xn;
    }
}

template <int NM, bool RL, int RT, int FAST>
__global__ void __launch_bounds__(SF_THREADS, 1)
scan_fused_kernel(const __grid_constant__ ModalCoef mc, const __grid_constant__ DevChain ch, const __grid_constant__ SfGeom g,
                  DevStream *__restrict__ streams, const uint8_t *__restrict__ in, uint8_t *__restrict__ out,
                  const uint32_t *__restrict__ ckpt_l, const uint32_t *__restrict__ ckpt_r,
                  uint32_t *__restrict__ tail_l, uint32_t *__restrict__ tail_r)
{
    constexpr int WPS = RT == ICW_RENDER_TPDF ? 4 : RT == ICW_RENDER_RPDF ? 2 : 0;
    constexpr int WPSD = WPS ? WPS : 1;                     // divisor that exists for every instantiation
    constexpr int GEN_WORDS = sf_gen_words(WPS);
    extern __shared__ __align__(128) uint8_t sf_smem[];
    const int tid = threadIdx.x;
    const int unit = blockIdx.x;
    DevStream &st = streams[0];

    // ---- this unit's frames [U0, U1) and the stream words that belong to them ------------------------------------
    int64_t U0, U1, w_unit = 0;
    if (WPS) {
        w_unit = g.first_word + (int64_t)unit * g.blocks_per_unit * ICW_MT_N;
        const int64_t w_next = w_unit + (int64_t)g.blocks_per_unit * ICW_MT_N;
        U0 = (w_unit - g.want_lo) / WPSD;  U0 = U0 < 0 ? 0 : U0;
        U1 = (w_next - g.want_lo) / WPSD;  U1 = U1 > g.n_frames ? g.n_frames : U1;
    } else {
        U0 = (int64_t)unit * g.frames_per_unit;
        U1 = U0 + g.frames_per_unit;  U1 = U1 > g.n_frames ? g.n_frames : U1;
    }
    if (U0 >= U1) return;                                   // nothing of the call falls into this unit (whole CTA leaves)
    const int64_t Wb = U0 - g.warm > 0 ? U0 - g.warm : 0;   // warm-up starts here (from the stream's state when that is frame 0)
    const int n_warm = (int)((U0 - Wb + SF_R - 1) / SF_R);
    const int n_real = (int)((U1 - U0 + SF_R - 1) / SF_R);

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(sf_smem + SfSmem::BAR)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {   // constants and the running state
        double *ksh = reinterpret_cast<double *>(sf_smem + SfSmem::KSH);
        for (int i = tid; i < 6 * SCAN_NMAX; i += SF_THREADS) ksh[i] = mc.k[i / SCAN_NMAX][i % SCAN_NMAX];
        double *carry = reinterpret_cast<double *>(sf_smem + SfSmem::CARRY);
        for (int i = tid; i < 4 * SCAN_NMAX * 2; i += SF_THREADS) {
            const int cf = i / (SCAN_NMAX * 2), j = i % (SCAN_NMAX * 2);
            carry[i] = Wb == 0 ? st.hb[cf >> 1][cf & 1][j] : 0.0;
        }
    }
    __syncthreads();

    if (tid < SF_SCAN_THREADS) {
        // ==========================================================================================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 144;");     // 256 x 144 + 384 x 64 = 640 x 96: exactly what the CTA was given at launch
        const int filt = tid & 1, chan = (tid >> 1) & 1, cf = tid & 3, c = tid >> 2;
        const double (*ksh)[SCAN_NMAX] = reinterpret_cast<const double (*)[SCAN_NMAX]>(sf_smem + SfSmem::KSH);
        double *carry = reinterpret_cast<double *>(sf_smem + SfSmem::CARRY);
        double *E = reinterpret_cast<double *>(sf_smem + SfSmem::SLOTS);          // [chunk][cf][mode][2]
        const double *xd = reinterpret_cast<const double *>(sf_smem + SfSmem::XD);
        uint8_t *stage = sf_smem + SfSmem::STAGE;
        const uint32_t bar = smem_u32(sf_smem + SfSmem::BAR);
        const int fb = ch.frame_bytes;
        const unsigned q0 = st.quad[chan];
        const bool direct = mc.baseline != 0;
        const double d0x2 = 2.0 * mc.d0;
        uint32_t parity = 0;
        int gslice = 0;                                     // slices produced so far (ring position)

        // bytes of frames [F, F + len) -> stage, by one bulk copy (16-byte granules around the span) or, for the call's
        // last range (nothing may be read past the caller's buffer), by plain loads
        auto fetch = [&](int64_t F, int len) {
            const uint8_t *src = in + F * fb;
            const uint32_t a = (uint32_t)((uintptr_t)src & 15u);
            if (F + len < g.n_frames) {
                if (tid == 0) {
                    const uint32_t bytes = (a + (uint32_t)len * (uint32_t)fb + 15u) & ~15u;
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 :: "r"(smem_u32(stage)), "l"(src - a), "r"(bytes), "r"(bar) : "memory");
                }
            } else {
                const int nb = len * fb;
                for (int i = tid; i < nb; i += SF_SCAN_THREADS) stage[a + i] = src[i];
                nb_sync(NB_SCAN, SF_SCAN_THREADS);
                if (tid == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
            }
        };
        auto range_of = [&](int r, int64_t &F, int &len) {  // r < n_warm: warm-up ranges, then the unit's own
            if (r < n_warm) { F = Wb + (int64_t)r * SF_R; const int64_t l = U0 - F; len = (int)(l < SF_R ? l : SF_R); }
            else { F = U0 + (int64_t)(r - n_warm) * SF_R; const int64_t l = U1 - F; len = (int)(l < SF_R ? l : SF_R); }
        };

        double kpr[NM], kpi[NM];
#pragma unroll
        for (int m = 0; m < NM; ++m) { kpr[m] = kconst<APPLY_UMASK>(mc, ksh, K_PR, m); kpi[m] = kconst<APPLY_UMASK>(mc, ksh, K_PI, m); }

        const int n_ranges = n_warm + n_real;
        {
            int64_t F; int len;
            range_of(0, F, len);
            fetch(F, len);
        }
        for (int r = 0; r < n_ranges; ++r) {
            int64_t F; int len;
            range_of(r, F, len);
            const bool real = r >= n_warm;
            const int nch = (len + SF_LC - 1) / SF_LC;                      // chunks with frames in them
            const int lc = c < nch - 1 ? SF_LC : c == nch - 1 ? len - (nch - 1) * SF_LC : 0;   // this chunk's length
            // ---- input: wait for the bulk copy, unpack once -----------------------------------------------------
            mbar_wait(bar, parity);
            parity ^= 1u;
            {
                const uint32_t a = (uint32_t)((uintptr_t)(in + F * fb) & 15u);
                double *xw = reinterpret_cast<double *>(sf_smem + SfSmem::XD);
                const int ns = len * 2;
                const int cb = ch.chan_bytes, stereo = ch.n_channels > 1;
                for (int i = tid; i < ns; i += SF_SCAN_THREADS)
                    xw[i] = unpack_real(ch.fmt, stage + a + (i >> 1) * fb + (stereo ? (i & 1) * cb : 0), 0);
            }
            nb_sync(NB_SCAN, SF_SCAN_THREADS);
            if (r + 1 < n_ranges) {
                int64_t Fn; int ln;
                range_of(r + 1, Fn, ln);
                fetch(Fn, ln);                                              // lands while this range is computed
            }
            // ---- LOCAL: this chunk's end state from zero -----------------------------------------------------------
            const unsigned qr = (q0 + (unsigned)(F & 3)) & 3u;              // mixer phase of the range's (and every chunk's) first frame
            const bool is_x = filt == (int)(qr & 1u);                       // fed on the chunk's first frame
            const int off = is_x ? 0 : 1;
            const double sg = mixer_sign(filt, (qr + (unsigned)off) & 3u);
            const int n_in = lc > off ? (lc - off + 1) >> 1 : 0;            // this filter's inputs in the chunk
            const double *xp = xd + (size_t)(c * SF_LC + off) * 2 + chan;
            {
                Cx s[NM];
                sf_local<NM, RL>(s, kpr, kpi, xp, n_in);
                // S~ -> S (the sign flips with every input); a filter whose last input was not the chunk's last frame has idled one sample
                const double se = (n_in & 1) ? -sg : sg;
                const bool idle = lc > 0 && ((lc - 1 - off) & 1) != 0;
                if (real) {
                    // the previous range's last slices must have been read before their memory holds end states
                    for (int k = 0; k < SF_NSLOT; ++k)
                        if (gslice + k >= SF_NSLOT) nb_sync(NB_EMPTY + ((gslice + k) % SF_NSLOT), SF_THREADS);
                }
#pragma unroll
                for (int m = 0; m < NM; ++m) {
                    Cx v; v.re = s[m].re * se; v.im = s[m].im * se;
                    if (idle) v = cx_mul(mc.p_re[m], mc.p_im[m], v);
                    E[((c * 4 + cf) * SCAN_NMAX + m) * 2] = v.re;
                    E[((c * 4 + cf) * SCAN_NMAX + m) * 2 + 1] = v.im;
                }
            }
            nb_sync(NB_SCAN, SF_SCAN_THREADS);
            // ---- carry over the chunks, serial, one thread per (channel, filter, mode) -----------------------------------
            if (tid < 4 * NM) {
                const int m = tid % NM, f4 = tid / NM;
                Cx acc; acc.re = carry[(f4 * SCAN_NMAX + m) * 2]; acc.im = carry[(f4 * SCAN_NMAX + m) * 2 + 1];
                const double plr = mc.pl_re[m], pli = mc.pl_im[m];           // p^SF_LC
                const int nfull = len / SF_LC;
                double *e = E + (f4 * SCAN_NMAX + m) * 2;
                double tr = e[0], ti = e[1];
                for (int k = 0; k < nfull; ++k) {
                    double *en = e + 4 * SCAN_NMAX * 2;
                    const double nr = k + 1 < nch ? en[0] : 0.0, ni = k + 1 < nch ? en[1] : 0.0;    // next chunk's, a step ahead
                    e[0] = acc.re; e[1] = acc.im;                           // state before chunk k
                    const double ar = fma(plr, acc.re, fma(-pli, acc.im, tr));
                    const double ai = fma(plr, acc.im, fma(pli, acc.re, ti));
                    acc.re = ar; acc.im = ai;
                    tr = nr; ti = ni;
                    e = en;
                }
                if (nfull < nch) {                                          // a ragged last chunk: p^(its length), by repeated multiplication
                    const int rest = len - nfull * SF_LC;
                    e[0] = acc.re; e[1] = acc.im;
                    Cx pw; pw.re = 1.0; pw.im = 0.0;
                    for (int k = 0; k < rest; ++k) pw = cx_mul(mc.p_re[m], mc.p_im[m], pw);
                    const Cx w = cx_mul(pw.re, pw.im, acc);
                    acc.re = w.re + tr; acc.im = w.im + ti;
                }
                carry[(f4 * SCAN_NMAX + m) * 2] = acc.re; carry[(f4 * SCAN_NMAX + m) * 2 + 1] = acc.im;
            }
            nb_sync(NB_SCAN, SF_SCAN_THREADS);
            if (!real) continue;                                            // warm-up: only the state moves on
            // ---- APPLY: the chunk again from its true state, analytic halves into the slice ring ------------------------------
            Cx S[NM];
#pragma unroll
            for (int m = 0; m < NM; ++m) {
                Cx s0; s0.re = E[((c * 4 + cf) * SCAN_NMAX + m) * 2]; s0.im = E[((c * 4 + cf) * SCAN_NMAX + m) * 2 + 1];
                // X last saw input two samples ago: step it back one sample (exact algebra: S = s / p); then S -> S~
                if (is_x) s0 = cx_mul(mc.pinv_re[m], mc.pinv_im[m], s0);
                S[m].re = s0.re * sg; S[m].im = s0.im * sg;
            }
            nb_sync(NB_SCAN, SF_SCAN_THREADS);                              // every start state is in registers: the ring is a ring again
            double kcr[NM], kci[NM], kcpr[NM], kcpi[NM];
#pragma unroll
            for (int m = 0; m < NM; ++m) {
                kcr[m] = kconst<APPLY_UMASK>(mc, ksh, K_CR, m);   kci[m] = kconst<APPLY_UMASK>(mc, ksh, K_CI, m);
                kcpr[m] = kconst<APPLY_UMASK>(mc, ksh, K_CPR, m); kcpi[m] = kconst<APPLY_UMASK>(mc, ksh, K_CPI, m);
            }
            // where this thread's values go inside a slot: plane = channel, (chunk, frame of the slice) -> 16 bytes (re, im)
            const int tb = chan * SF_PLANE + c * (SF_SL * 16);
            auto slot_of = [&](int k) { return sf_smem + SfSmem::SLOTS + ((gslice + k) % SF_NSLOT) * SF_SLOT + tb; };
            if (off && lc > 0) {                                            // frame 0 follows an input of the previous chunk
                double y2 = 0.0;
#pragma unroll
                for (int m = 0; m < NM; ++m) {
                    if (RL && m == NM - 1) y2 = fma(kcr[m], S[m].re, y2);
                    else y2 = fma(kcr[m], S[m].re, fma(kci[m], S[m].im, y2));
                }
                *reinterpret_cast<double *>(slot_of(0) + 8) = y2;
            }
            for (int k = 0; k < SF_NSLICE; ++k) {
                uint8_t *sl = slot_of(k);
#pragma unroll
                for (int j = 0; j < SF_SL / 2; ++j) {
                    const int f = k * SF_SL + 2 * j + off;                  // frame of the chunk this filter is fed on
                    if (j == SF_SL / 2 - 1 && k + 1 < SF_NSLICE && k + 1 >= SF_NSLOT)
                        nb_sync(NB_EMPTY + ((gslice + k + 1) % SF_NSLOT), SF_THREADS);   // the odd filter's last value lands in the next slice
                    if (f < lc) {
                        const double xin = xp[(size_t)(k * (SF_SL / 2) + j) * 4];
                        // four interleaved partial sums per output: with two scan warps per scheduler a 20-term chain of
                        // dependent DFMAs (8 cycles each) is what the warp would otherwise spend its time waiting on
                        double a1[4] = { direct ? d0x2 * xin : 0.0, 0.0, 0.0, 0.0 }, a2[4] = { 0.0, 0.0, 0.0, 0.0 };
#pragma unroll
                        for (int m = 0; m < NM; ++m) {
                            if (RL && m == NM - 1) {                        // the real pole: im == 0 and its weights are 0
                                a1[m & 3] = fma(kcpr[m], S[m].re, a1[m & 3]);
                                re_step(S[m], kpr[m], -xin);
                                a2[m & 3] = fma(kcr[m], S[m].re, a2[m & 3]);
                            } else {
                                a1[m & 3] = fma(kcpr[m], S[m].re, fma(kcpi[m], S[m].im, a1[m & 3]));
                                cx_step(S[m], kpr[m], kpi[m], -xin);
                                a2[m & 3] = fma(kcr[m], S[m].re, fma(kci[m], S[m].im, a2[m & 3]));
                            }
                        }
                        const double y1 = (a1[0] + a1[1]) + (a1[2] + a1[3]), y2 = (a2[0] + a2[1]) + (a2[2] + a2[3]);
                        *reinterpret_cast<double *>(sl + (2 * j + off) * 16) = y1;                       // re of its own frame
                        if (f + 1 < lc) {                                                               // im of the frame after it
                            if (2 * j + off + 1 < SF_SL) *reinterpret_cast<double *>(sl + (2 * j + off + 1) * 16 + 8) = y2;
                            else *reinterpret_cast<double *>(slot_of(k + 1) + 8) = y2;
                        }
                    }
                }
                nb_arrive(NB_FULL + ((gslice + k) % SF_NSLOT), SF_THREADS);
            }
            gslice += SF_NSLICE;
        }
        // the stream's filter state after the call's last frame (modal basis, like scan_apply_kernel leaves it)
        if (U1 == g.n_frames && tid < 4 * NM) {
            const int m = tid % NM, f4 = tid / NM;
            st.hb[f4 >> 1][f4 & 1][2 * m] = carry[(f4 * SCAN_NMAX + m) * 2];
            st.hb[f4 >> 1][f4 & 1][2 * m + 1] = carry[(f4 * SCAN_NMAX + m) * 2 + 1];
        }
    } else {
        // ==============================================================================================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        const int t = tid - SF_SCAN_THREADS;
        uint32_t *ub_l = reinterpret_cast<uint32_t *>(sf_smem + SfSmem::WORDS);      // [624 history | new words], untempered
        uint32_t *ub_r = ub_l + GEN_WORDS;
        const int gen = t / SF_MT_PASS, lane_w = t % SF_MT_PASS;
        uint32_t *my_ub = gen ? ub_r : ub_l;
        int64_t gen_pos = w_unit;                           // next stream word to make (both generators stand at the same draw)
        int n_prev = 0;                                     // new words of the previous fill
        if (WPS) {
            for (int i = t; i < ICW_MT_N; i += SF_PW_THREADS) {     // history of the first fill = the unit's checkpoint
                ub_l[i] = ckpt_l[(size_t)unit * ICW_MT_N + i];
                ub_r[i] = ckpt_r[(size_t)unit * ICW_MT_N + i];
            }
        }
        // the states around the call's last block go back to the host's bookkeeping (MtPlan::tail)
        const int64_t tailw = g.first_word + g.tail_block * ICW_MT_N;       // first word of that block
        const bool has_tail = WPS && g.tail_block >= (int64_t)unit * g.blocks_per_unit && g.tail_block < (int64_t)(unit + 1) * g.blocks_per_unit;

        FrameAcc acc;
        OscCounter osc;
        const int pc = t / SF_SL, pf = t % SF_SL;           // this thread's (chunk, frame of the slice)
        osc.init(ch, st.n_frame, U0 + pc * SF_LC + pf < g.n_frames ? U0 + pc * SF_LC + pf : 0);
        int gslice = 0;
        for (int r = 0; r < n_real; ++r) {
            const int64_t F = U0 + (int64_t)r * SF_R;
            const int len = (int)(U1 - F < SF_R ? U1 - F : SF_R);
            int64_t base = 0;                               // stream word at buffer position 624
            if (WPS) {
                // ---- the range's words: both generators, SF_MT_PASS words each per pass ---------------------------------------
                if (r > 0) {                                // the last 624 words of the previous fill become the history
                    uint32_t keep[4];
                    int nk = 0;
                    for (int i = t; i < 2 * ICW_MT_N; i += SF_PW_THREADS) {
                        const uint32_t *b = i < ICW_MT_N ? ub_l : ub_r;
                        keep[nk++] = b[n_prev + (i < ICW_MT_N ? i : i - ICW_MT_N)];
                    }
                    nb_sync(NB_PW, SF_PW_THREADS);
                    nk = 0;
                    for (int i = t; i < 2 * ICW_MT_N; i += SF_PW_THREADS) {
                        uint32_t *b = i < ICW_MT_N ? ub_l : ub_r;
                        b[i < ICW_MT_N ? i : i - ICW_MT_N] = keep[nk++];
                    }
                }
                nb_sync(NB_PW, SF_PW_THREADS);
                base = gen_pos;
                int64_t end = g.want_lo + (F + len) * WPS;
                if (has_tail && F + len == g.n_frames) end = tailw + ICW_MT_N;      // through the end of the call's last block
                const int n_new = (int)(end - base);
                uint32_t *u = my_ub + ICW_MT_N + lane_w;
                const int full = n_new / SF_MT_PASS, rag = n_new % SF_MT_PASS;
                for (int k = 0; k < full; ++k, u += SF_MT_PASS) {
                    mt_window_word(u);
                    nb_sync(NB_PW, SF_PW_THREADS);
                }
                if (rag) {
                    if (lane_w < rag) mt_window_word(u);
                    nb_sync(NB_PW, SF_PW_THREADS);
                }
                if (has_tail) {
                    // buffer position p holds stream word base - 624 + p; a 624-word block that lies wholly inside goes out
                    for (int h = 0; h < 2; ++h) {
                        const int64_t b0 = tailw - ICW_MT_N + (int64_t)h * ICW_MT_N - (base - ICW_MT_N);
                        if (b0 >= 0 && b0 + ICW_MT_N <= ICW_MT_N + n_new)
                            for (int i = t; i < ICW_MT_N; i += SF_PW_THREADS) {
                                tail_l[h * ICW_MT_N + i] = ub_l[b0 + i];
                                tail_r[h * ICW_MT_N + i] = ub_r[b0 + i];
                            }
                    }
                }
                gen_pos = end;
                n_prev = n_new;
            }
            // ---- the frames, slice by slice as the scan warps finish them ---------------------------------------------------------
            for (int k = 0; k < SF_NSLICE; ++k) {
                const int slot = (gslice + k) % SF_NSLOT;
                nb_sync(NB_FULL + slot, SF_THREADS);
                const uint8_t *sl = sf_smem + SfSmem::SLOTS + slot * SF_SLOT + (pc * SF_SL + pf) * 16;
                const double2 a0 = *reinterpret_cast<const double2 *>(sl), a1 = *reinterpret_cast<const double2 *>(sl + SF_PLANE);
                nb_arrive(NB_EMPTY + slot, SF_THREADS);
                const int f = pc * SF_LC + k * SF_SL + pf;
                if (f < len) {
                    const int64_t i = F + f;
                    uint4 wl = make_uint4(0u, 0u, 0u, 0u), wr = wl;
                    if (WPS) {
                        const int o = ICW_MT_N + (int)(g.want_lo + i * WPS - base);
                        if (WPS == 4) {
                            wl = *reinterpret_cast<const uint4 *>(ub_l + o);
                            wr = *reinterpret_cast<const uint4 *>(ub_r + o);
                            wl.z = mt_temper_mul(wl.z); wl.w = mt_temper_mul(wl.w); wr.z = mt_temper_mul(wr.z); wr.w = mt_temper_mul(wr.w);
                        } else {
                            const uint2 a = *reinterpret_cast<const uint2 *>(ub_l + o), b = *reinterpret_cast<const uint2 *>(ub_r + o);
                            wl.x = a.x; wl.y = a.y; wr.x = b.x; wr.y = b.y;
                        }
                        wl.x = mt_temper_mul(wl.x); wl.y = mt_temper_mul(wl.y); wr.x = mt_temper_mul(wr.x); wr.y = mt_temper_mul(wr.y);
                    }
                    double v[4] = { a0.x, a0.y, a1.x, a1.y };
                    if (FAST) {
                        lean_frame_fast<RT, FRAME_NO_LAST>(ch, st, i, -1, v, wl, wr, out, acc, osc);
                        if (i == g.n_frames - 1) lean_frame_fast<RT, FRAME_LAST_ONLY>(ch, st, i, i, v, wl, wr, out, acc, osc);
                    } else {
                        lean_frame<ICW_SHAPE_GENERIC, RT, FRAME_NO_LAST>(ch, st, i, -1, v, wl, wr, out, g.out_aligned, acc, osc);
                        if (i == g.n_frames - 1) lean_frame<ICW_SHAPE_GENERIC, RT, FRAME_LAST_ONLY>(ch, st, i, i, v, wl, wr, out, g.out_aligned, acc, osc);
                    }
                }
            }
            gslice += SF_NSLICE;
        }
        commit_acc(&st, acc, 32);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------------------------
bool sfused_supports(const DevChain &ch, int n_streams, bool taps_or_pre)
{
    if (n_streams != 1 || ch.is_complex || taps_or_pre || ch.fp_check || ch.bypass) return false;
    if (ch.fmt < ICW_FMT_WAV_U8 || ch.fmt > ICW_FMT_WAV_F32) return false;
    if ((ch.n_fade_in | ch.n_fade_out) != 0) return false;              // fades: the three-kernel path
    if (ch.render.ns_kind != 0) return false;
    const int rt = ch.render.render_type;
    if (rt != ICW_RENDER_ROUND && rt != ICW_RENDER_RPDF && rt != ICW_RENDER_TPDF) return false;
    return ch.shape == ICW_SHAPE_MASTER || ch.shape == ICW_SHAPE_SHIFT_MASTER;
}

int64_t sfused_warm_frames(const ModalCoef &mc)
{
    double rmax = 0.0;
    for (int m = 0; m < mc.nm; ++m) {
        const double r = std::sqrt(mc.p_re[m] * mc.p_re[m] + mc.p_im[m] * mc.p_im[m]);
        if (r > rmax) rmax = r;
    }
    const double need = rmax > 0.0 && rmax < 1.0 ? std::log(1e-19) / std::log(rmax) : 1e9;
    const int64_t ranges = (int64_t)(need / SF_R) + 1;
    return ranges * SF_R;
}

template <int NM, bool RL, int RT, int FAST>
static cudaError_t sf_launch(const ModalCoef &mc, const DevChain &ch, const SfGeom &g, DevStream *streams, const uint8_t *in,
                             uint8_t *out, const MtPlan *pl, const MtPlan *pr, cudaStream_t s)
{
    const int wps = RT == ICW_RENDER_TPDF ? 4 : RT == ICW_RENDER_RPDF ? 2 : 0;
    const size_t smem = sfused_smem_bytes(wps);
    cudaError_t e = cudaFuncSetAttribute(scan_fused_kernel<NM, RL, RT, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    scan_fused_kernel<NM, RL, RT, FAST><<<g.n_units, SF_THREADS, smem, s>>>(mc, ch, g, streams, in, out,
                                                                            pl ? pl->ckpt : nullptr, pr ? pr->ckpt : nullptr,
                                                                            pl ? pl->tail : nullptr, pr ? pr->tail : nullptr);
    return cudaGetLastError();
}

template <int NM, bool RL>
static cudaError_t sf_launch_nm(const ModalCoef &mc, const DevChain &ch, const SfGeom &g, DevStream *streams, const uint8_t *in,
                                uint8_t *out, const MtPlan *pl, const MtPlan *pr, bool fast, cudaStream_t s)
{
#define ICW_SF(RT) (fast ? sf_launch<NM, RL, RT, 1>(mc, ch, g, streams, in, out, pl, pr, s) : sf_launch<NM, RL, RT, 0>(mc, ch, g, streams, in, out, pl, pr, s))
    switch (ch.render.render_type) {
    case ICW_RENDER_ROUND: return ICW_SF(ICW_RENDER_ROUND);
    case ICW_RENDER_RPDF:  return ICW_SF(ICW_RENDER_RPDF);
    case ICW_RENDER_TPDF:  return ICW_SF(ICW_RENDER_TPDF);
    default: return cudaErrorInvalidValue;
    }
#undef ICW_SF
}

cudaError_t launch_scan_fused(const ModalCoef &mc, const DevChain &ch, DevStream *streams, int64_t n_frames, const uint8_t *in,
                              uint8_t *out, const MtPlan *pl, const MtPlan *pr, int n_cta, int64_t warm, cudaStream_t s)
{
    if (mc.L != SF_LC) return cudaErrorInvalidValue;
    SfGeom g;
    memset(&g, 0, sizeof g);
    g.n_frames = n_frames;
    g.warm = warm;
    g.out_aligned = ((size_t)(uintptr_t)out & 3u) == 0;
    const int wps = ch.render.words_per_sample;
    if (wps) {
        if (!pl || !pr || pl->n_units != pr->n_units || pl->blocks_per_unit != pr->blocks_per_unit || pl->first_word != pr->first_word ||
            pl->want_lo != pr->want_lo || pl->want_hi != pr->want_hi)
            return cudaErrorInvalidValue;                   // the two generators must stand at the same draw
        g.n_units = pl->n_units; g.blocks_per_unit = pl->blocks_per_unit;
        g.first_word = pl->first_word; g.want_lo = pl->want_lo; g.want_hi = pl->want_hi; g.tail_block = pl->tail_block;
    } else {
        int64_t per = (n_frames + n_cta - 1) / n_cta;
        per = (per + SF_R - 1) / SF_R * SF_R;               // whole ranges: only the last unit ends ragged
        g.frames_per_unit = per;
        g.n_units = (int)((n_frames + per - 1) / per);
    }
    const bool fast = lean_fast_ok(ch) && g.out_aligned;
    switch (mc.nm) {
    case 8:  return sf_launch_nm<8, true>(mc, ch, g, streams, in, out, pl, pr, fast, s);
    case 9:  return sf_launch_nm<9, false>(mc, ch, g, streams, in, out, pl, pr, fast, s);
    case 10: return mc.real_last ? sf_launch_nm<10, true>(mc, ch, g, streams, in, out, pl, pr, fast, s)
                                : sf_launch_nm<10, false>(mc, ch, g, streams, in, out, pl, pr, fast, s);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace icw
