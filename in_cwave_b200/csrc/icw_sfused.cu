// icw_sfused.cu -- scan mode in ONE kernel: file bytes -> Hilbert (modal block scan) -> oscillator -> DSP list -> dither ->
// quantise -> PCM, with nothing but the input bytes and the PCM bytes crossing HBM.
//
// The three-kernel scan path (icw_scan.cu passes 1..3, then chain_mt_kernel) writes the analytic signal to HBM and
// reads it back: 64 of the 84 bytes per frame the C2 workload moved, and it runs every recurrence twice (chunk end
// states from zero, then the chunks again from their true states).  Here a chunk is 16 frames in BLOCK form
// (icw_sfused.h: SfTab): its effect on the filter state and each of its outputs are dot products with constants, so
//
//   level 0   E_c   = sum_j q^(7-j) (-x_j)                    thread = (chunk, channel), both filters     20 FMA / input
//   level 1   T_c+1 = q^8 T_c + E_c    serial over the chunks thread = (channel, filter, mode)             5 FMA / input
//   level 2   out   = C T_c + H x_c                           thread = (chunk, channel), both filters     48 FMA / input
//
// -- 73 FMA per filter input against 120 for the recurrences run twice, every one of them a DFMA whose constant is a
// uniform-register operand (the 2-cycle form), and a constant fetched once (LDCU.128) feeds four of
// them because a thread runs both filters of its channel.
//
// One CTA = one UNIT: a contiguous run of the stream's frames (the generators' jump-ahead unit, icw_mt.cu, when there
// is dither -- the draw index is a closed form of the frame index -- an even split otherwise), walked in RANGES of
// 512 frames by three groups of warps that only meet at two pairs of named barriers (FULL / EMPTY per buffer):
//
//   warps 0-1  SCAN   the range's bytes arrive by one cp.async.bulk (TMA), issued two ranges ahead by one thread and
//                     tracked by an mbarrier; a thread unpacks its own 16 samples, runs level 0, level 1 (40 threads),
//                     level 2, and writes its chunk's analytic frames into one of two buffers in shared memory
//   warps 2-3  GEN    one warp per channel's generator: the range's MT19937 words, 624 a pass (icw_mtdev.cuh, window
//                     form), into one of two slices of a linear buffer
//   warps 4-7  PW     four frames per thread: oscillator, DSP list, dither, quantiser, PCM store (icw_frame.cuh)
//
// so the FP64-bound scan of range r+1, the integer-only generators and the issue-bound pointwise code of range r run
// on the same SM at the same time.  Two CTAs per SM.  A unit's filter state at its first frame comes from a WARM-UP:
// levels 0 and 1 alone over the frames before it (the filters forget: |p|^warm is below 1e-19; the same fact the
// multi-GPU hand-off uses) -- no pass over the whole stream, no carry arrays in HBM, no second read of the input.
//
// Numerics: the same modal decomposition as icw_scan.cu with the same sign-free two-sample step; powers of the step
// are rounded once from long double instead of accumulating in the recurrence (tests/test_gpu_sfused.py holds the path
// to the three-kernel path within an LSB, tests/test_gpu_scan.py holds both to the binary128 truth at 1e-12).
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

constexpr int SF_HALF = 2 * SF_CH;                              // 64 threads = (chunk, channel): one half of the scan group
constexpr int SF_SCAN_THREADS = 2 * SF_HALF;                    // 128: thread = (half, chunk, channel); a half is two warps
constexpr int SF_GEN_THREADS = 64;                              // one warp per generator
constexpr int SF_PW_THREADS = SF_R / 4;                         // 128: four frames per thread and range
constexpr int SF_THREADS = SF_SCAN_THREADS + SF_GEN_THREADS + SF_PW_THREADS;
constexpr int SF_PW0 = SF_SCAN_THREADS + SF_GEN_THREADS;        // first pointwise thread
constexpr int SF_FPT = SF_R / SF_PW_THREADS;                    // frames per pointwise thread and range
#ifndef ICW_SF_PW_TRIP
#define ICW_SF_PW_TRIP 1
#endif
#ifndef ICW_SF_ORDER
#define ICW_SF_ORDER 0
#endif
constexpr int SF_PW_TRIP = ICW_SF_PW_TRIP;                      // of which side by side in one loop trip (1, 2 or 4)
constexpr int SF_APITCH = SF_LC * 16 + 16;                      // bytes per chunk of one channel's analytic plane (16 x (re, im), padded: banks)
constexpr int SF_PLANE = SF_CH * SF_APITCH + 64;                // the two planes sit 16 banks apart
constexpr int SF_ABUF = 2 * SF_PLANE;                           // one analytic buffer
constexpr int SF_STAGE = SF_R * 8 + 32;                         // a range's bytes as they lie in the file (<= 8 per frame) + alignment slack
constexpr int SF_ESLOTS = 2 * SF_HALF + 1;                      // (filter, chunk, channel) per mode; +1: the carry threads (one per mode) walk
                                                                // the chunks side by side, their 16-byte accesses must not share banks
static_assert(SF_LC % 4 == 0 && SF_NI == 8, "chunk geometry (the tables are written out for 8 inputs)");
static_assert(SF_HALF % 32 == 0 && SF_HALF >= 4 * SCAN_NMAX && SF_THREADS == 320, "warp roles");

// named barriers (id 0 is __syncthreads): scan-only, and FULL / EMPTY per buffer between the three groups
enum { NB_SCAN = 1, NB_FULL = 2 /* + buffer */, NB_EMPTY = 4 /* + buffer */ };

__device__ __forceinline__ void nb_sync(int id, int n) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void nb_arrive(int id, int n) { __threadfence_block(); asm volatile("bar.arrive %0, %1;" :: "r"(id), "r"(n) : "memory"); }
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
    static constexpr int BAR = 0;                                   // two mbarriers (input staging, one per stage buffer)
    static constexpr int CARRY = 64;                                // [4][SCAN_NMAX][2] doubles: modal state after the last frame done
    static constexpr int STAGE = 768;                               // two stage buffers
    static constexpr int ET = (STAGE + 2 * SF_STAGE + 127) / 128 * 128;     // [mode][SF_ESLOTS] (re, im): chunk effects E, then chunk start states T
    static constexpr int AN = (ET + SCAN_NMAX * SF_ESLOTS * 16 + 127) / 128 * 128;   // two analytic buffers of two planes
    static constexpr int WORDS = AN + 2 * SF_ABUF;                  // two generators' untempered words
};
static_assert(SfSmem::WORDS % 16 == 0, "16-byte loads of the dither words");
// per generator: history (624) | slice of an even range | slice of an odd range; the first fill of a unit also makes
// the words skipped in front of it (< 624), the last one runs through the end of the call's last block (< 624 more)
__host__ __device__ constexpr int sf_gen_words(int wps) { return wps ? 3 * ICW_MT_N + 2 * SF_R * wps : 0; }
size_t sfused_smem_bytes(int wps) { return (size_t)SfSmem::WORDS + 2 * (size_t)sf_gen_words(wps) * 4; }

// ------------------------------------------------------------------------------------------------------------------
// a scan thread's own 16 samples out of the staged bytes (mono: both channel threads read the same sample), zeros
// behind a short range.  A sample is at most 4 bytes at any byte alignment: two aligned words and a funnel shift;
// the values are unpack_real()'s (icw_dev.cuh), format known at compile time.
// ------------------------------------------------------------------------------------------------------------------
template <int FMT>
__device__ __forceinline__ void sf_load_x(uint32_t sb /* shared address of the thread's first sample */, int fb, int n_valid, double (&x)[SF_LC])
{
#pragma unroll
    for (int f = 0; f < SF_LC; ++f) {
        double v = 0.0;
        if (f < n_valid) {
            const uint32_t addr = sb + (uint32_t)(f * fb);
            uint32_t w0, w1;
            asm volatile("ld.shared.u32 %0, [%2];\n\tld.shared.u32 %1, [%2+4];" : "=r"(w0), "=r"(w1) : "r"(addr & ~3u));
            const uint32_t lo = __funnelshift_r(w0, w1, (addr & 3u) * 8u);
            switch (FMT) {
            case ICW_FMT_WAV_U8:  v = 256.0 * (double)(int)(int8_t)(uint8_t)((lo & 0xFFu) - 0x80u); break;
            case ICW_FMT_WAV_I16: v = (double)(int)(int16_t)(lo & 0xFFFFu); break;
            case ICW_FMT_WAV_I24: v = (double)((int32_t)(lo << 8) >> 8) * 0.00390625; break;         // /256.0, exact
            case ICW_FMT_WAV_I32: v = (double)(int32_t)lo * (1.0 / 65536.0); break;
            default:              v = 32768.0 * (double)__uint_as_float(lo); break;
            }
        }
        x[f] = v;
    }
}

// The same for a whole chunk with the frame size known at compile time: the thread's span of bytes comes in as aligned
// words, is shifted to the thread's own byte alignment once, and every sample is then a fixed word and a fixed shift.
template <int FMT, int FB>
__device__ __forceinline__ void sf_load_x_full(uint32_t sb, double (&x)[SF_LC])
{
    constexpr int SB = FMT == ICW_FMT_WAV_U8 ? 1 : FMT == ICW_FMT_WAV_I16 ? 2 : FMT == ICW_FMT_WAV_I24 ? 3 : 4;
    constexpr int NA = ((SF_LC - 1) * FB + SB + 3) / 4;     // words of the span once it starts on a word
    const uint32_t *wp = reinterpret_cast<const uint32_t *>(__cvta_shared_to_generic(sb & ~3u));
    const uint32_t sh = (sb & 3u) * 8u;
    uint32_t w[NA + 1], al[NA];
#pragma unroll
    for (int k = 0; k <= NA; ++k) w[k] = wp[k];
#pragma unroll
    for (int k = 0; k < NA; ++k) al[k] = __funnelshift_r(w[k], w[k + 1], sh);
#pragma unroll
    for (int f = 0; f < SF_LC; ++f) {
        constexpr int dummy = 0; (void)dummy;
        const int o = f * FB, k = o >> 2, s8 = (o & 3) * 8;
        const uint32_t lo = s8 == 0 ? al[k] : (k + 1 < NA ? __funnelshift_r(al[k], al[k + 1], s8) : al[k] >> s8);
        double v;
        switch (FMT) {
        case ICW_FMT_WAV_U8:  v = 256.0 * (double)(int)(int8_t)(uint8_t)((lo & 0xFFu) - 0x80u); break;
        case ICW_FMT_WAV_I16: v = (double)(int)(int16_t)(lo & 0xFFFFu); break;
        case ICW_FMT_WAV_I24: v = (double)((int32_t)(lo << 8) >> 8) * 0.00390625; break;
        case ICW_FMT_WAV_I32: v = (double)(int32_t)lo * (1.0 / 65536.0); break;
        default:              v = 32768.0 * (double)__uint_as_float(lo); break;
        }
        x[f] = v;
    }
}

template <int FMT>
__device__ __forceinline__ void sf_load_x_any(uint32_t sb, int fb, int n_valid, double (&x)[SF_LC])
{
    constexpr int SB = FMT == ICW_FMT_WAV_U8 ? 1 : FMT == ICW_FMT_WAV_I16 ? 2 : FMT == ICW_FMT_WAV_I24 ? 3 : 4;
    if (n_valid >= SF_LC && fb == SB) sf_load_x_full<FMT, SB>(sb, x);               // mono
    else if (n_valid >= SF_LC && fb == 2 * SB) sf_load_x_full<FMT, 2 * SB>(sb, x);  // stereo
    else sf_load_x<FMT>(sb, fb, n_valid, x);
}

// ------------------------------------------------------------------------------------------------------------------
// level 0: both filters' chunk effects E = sum_j q[j] x_j  (X is fed on the chunk's even frames, Y on the odd ones)
// ------------------------------------------------------------------------------------------------------------------
template <int NM, bool RL, int M0, int M1>
__device__ __forceinline__ void sf_level0_half(const SfTab &tb, const double (&x)[SF_LC], double2 *ex, double2 *ey)
{
    double xr[M1 - M0], xi[M1 - M0], yr[M1 - M0], yi[M1 - M0];
#pragma unroll
    for (int m = M0; m < M1; ++m) {
        xr[m - M0] = tb.q[0][m][0] * x[0];  yr[m - M0] = tb.q[0][m][0] * x[1];
        xi[m - M0] = tb.q[0][m][1] * x[0];  yi[m - M0] = tb.q[0][m][1] * x[1];
    }
#pragma unroll
    for (int j = 1; j < SF_NI; ++j) {
#pragma unroll
        for (int m = M0; m < M1; ++m) {
            xr[m - M0] = fma(tb.q[j][m][0], x[2 * j], xr[m - M0]);
            yr[m - M0] = fma(tb.q[j][m][0], x[2 * j + 1], yr[m - M0]);
            if (!(RL && m == NM - 1)) {                             // the real pole: its state has no imaginary part
                xi[m - M0] = fma(tb.q[j][m][1], x[2 * j], xi[m - M0]);
                yi[m - M0] = fma(tb.q[j][m][1], x[2 * j + 1], yi[m - M0]);
            }
        }
    }
#pragma unroll
    for (int m = M0; m < M1; ++m) {
        ex[m * SF_ESLOTS] = make_double2(xr[m - M0], (RL && m == NM - 1) ? 0.0 : xi[m - M0]);
        ey[m * SF_ESLOTS] = make_double2(yr[m - M0], (RL && m == NM - 1) ? 0.0 : yi[m - M0]);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// level 2: the chunk's 16 frames of one channel from the two filters' start states and its own inputs
//   X (fed on even frames): A_i -> re of frame 2i,     B_k (k = 1..8) -> im of frame 2k - 1
//   Y (fed on odd frames):  A_i -> re of frame 2i + 1, B_k (k = 0..7) -> im of frame 2k
// The A outputs (real parts) and the B outputs (imaginary parts) are made by different warps: the same inputs and
// start states, half of the sums each, and a constant pair still feeds four DFMAs (both filters, re and im).
// ------------------------------------------------------------------------------------------------------------------
template <int NM, bool RL>
__device__ __forceinline__ void sf_level2_a(const SfTab &tb, const double (&x)[SF_LC], const double2 *tx, const double2 *ty,
                                            double (&ax)[SF_NI], double (&ay)[SF_NI])
{
    // the chunk's own inputs (taps of the impulse response); the direct term only with the baseline summation
#pragma unroll
    for (int i = 0; i < SF_NI; ++i) {
        ax[i] = tb.d0x2 * x[2 * i];
        ay[i] = tb.d0x2 * x[2 * i + 1];
#pragma unroll
        for (int j = 0; j < i; ++j) {
            ax[i] = fma(tb.ha[i - 1 - j], x[2 * j], ax[i]);
            ay[i] = fma(tb.ha[i - 1 - j], x[2 * j + 1], ay[i]);
        }
    }
    // the start states, mode by mode: a LOOP whose counter only indexes the constants (so it lives in a uniform register and
    // the constants still arrive as LDCU / uniform operands) -- unrolled, the two halves of level 2 alone were 30 KB of
    // straight-line code, and the SM's instruction caches are 6 KB (L0) and 32 KB (L1.5) for everything the CTAs run
    double2 sx = *tx, sy = *ty;
#pragma unroll 1
    for (int m = 0; m < NM; ++m) {
        tx += SF_ESLOTS; ty += SF_ESLOTS;
        const double2 nx = *tx, ny = *ty;               // the next mode's states, a trip ahead (behind the last mode: shared memory that is there, unused)
#pragma unroll
        for (int i = 0; i < SF_NI; ++i) {
            const double2 k = *reinterpret_cast<const double2 *>(&tb.ca[i][m][0]);      // the real pole's imaginary weight is 0, and so is its state's
            ax[i] = fma(k.x, sx.x, fma(k.y, sx.y, ax[i]));
            ay[i] = fma(k.x, sy.x, fma(k.y, sy.y, ay[i]));
        }
        sx = nx; sy = ny;
    }
}

template <int NM, bool RL>
__device__ __forceinline__ void sf_level2_b(const SfTab &tb, const double (&x)[SF_LC], const double2 *tx, const double2 *ty,
                                            double (&bx)[SF_NI], double (&by)[SF_NI])
{
#pragma unroll
    for (int k = 0; k < SF_NI; ++k) {
        // bx[k] is B_(k+1) of X: inputs 0..k; by[k] is B_k of Y: inputs 0..k-1
        bx[k] = tb.hb[k] * x[0];
        by[k] = 0.0;
#pragma unroll
        for (int j = 1; j <= k; ++j) bx[k] = fma(tb.hb[k - j], x[2 * j], bx[k]);
#pragma unroll
        for (int j = 0; j < k; ++j) by[k] = fma(tb.hb[k - 1 - j], x[2 * j + 1], by[k]);
    }
    double2 sx = *tx, sy = *ty;
#pragma unroll 1
    for (int m = 0; m < NM; ++m) {
        tx += SF_ESLOTS; ty += SF_ESLOTS;
        const double2 nx = *tx, ny = *ty;
#pragma unroll
        for (int k = 0; k < SF_NI; ++k) {
            const double2 k1 = *reinterpret_cast<const double2 *>(&tb.cb[k + 1][m][0]), k0 = *reinterpret_cast<const double2 *>(&tb.cb[k][m][0]);
            bx[k] = fma(k1.x, sx.x, fma(k1.y, sx.y, bx[k]));
            by[k] = fma(k0.x, sy.x, fma(k0.y, sy.y, by[k]));
        }
        sx = nx; sy = ny;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// MT19937 in windows of up to 624 words: u -> the first new word, everything before it is there.  A lane owns the
// virtual lanes v = l, l + LANES, ... < 227; word v needs nothing newer than v - 227, word v + 227 needs word v (its
// own), word v + 454 needs word v + 227 (its own) and words v - 170, v - 169: older than the window for v < 169, and
// for v = 169 the window's word 0, made again here rather than waited for (mt_jrnd.c:121 pairs the block's last word
// with the NEW word 0 in the same way).
// ------------------------------------------------------------------------------------------------------------------
template <int LANES>
__device__ __forceinline__ void mt_window_pass(uint32_t *__restrict__ u, int n, int l)
{
#pragma unroll
    for (int v0 = 0; v0 < 227; v0 += LANES) {
        const int v = v0 + l;
        if (v < 227 && v < n) {
            const uint32_t a = u[v - 227] ^ mt_twist(u[v - 624], u[v - 623]);
            u[v] = a;
            if (v + 227 < n) {
                const uint32_t b = a ^ mt_twist(u[v - 397], u[v - 396]);
                u[v + 227] = b;
                if (v < 170 && v + 454 < n) {
                    const uint32_t nx = v == 169 ? (u[-227] ^ mt_twist(u[-624], u[-623])) : u[v - 169];
                    u[v + 454] = b ^ mt_twist(u[v - 170], nx);
                }
            }
        }
    }
}

// a frame's dither words out of the generators' buffers (untempered there: the recurrence needs them so), tempered here
template <int WPS>
__device__ __forceinline__ void sf_words(const uint32_t *ub_l, const uint32_t *ub_r, int o, uint4 &wl, uint4 &wr)
{
    if (WPS == 4) {
        wl = *reinterpret_cast<const uint4 *>(ub_l + o);
        wr = *reinterpret_cast<const uint4 *>(ub_r + o);
        wl.z = mt_temper_mul(wl.z); wl.w = mt_temper_mul(wl.w); wr.z = mt_temper_mul(wr.z); wr.w = mt_temper_mul(wr.w);
    } else {
        const uint2 wa = *reinterpret_cast<const uint2 *>(ub_l + o), wb = *reinterpret_cast<const uint2 *>(ub_r + o);
        wl.x = wa.x; wl.y = wa.y; wr.x = wb.x; wr.y = wb.y;
    }
    wl.x = mt_temper_mul(wl.x); wl.y = mt_temper_mul(wl.y); wr.x = mt_temper_mul(wr.x); wr.y = mt_temper_mul(wr.y);
}

// where a range's generator words lie: fills alternate between the two slices of the linear buffer (an even fill
// starts at 624 behind the history, an odd one follows the even one); pure arithmetic, kept by GEN and PW alike
struct SfWordGeom {
    int64_t gen_pos;        // next stream word to make (both generators stand at the same draw)
    int wp;                 // buffer position behind the last word made
    // this range's fill
    int64_t base;           // stream word at buffer position `pos`
    int pos, n_new;
    bool move;              // the last 624 words made go to the front first
    __device__ __forceinline__ void next(int rr, int64_t end)
    {
        base = gen_pos;
        n_new = (int)(end - base);
        move = (rr & 1) == 0 && rr > 0;
        pos = (rr & 1) == 0 ? ICW_MT_N : wp;
        gen_pos = end;
    }
    __device__ __forceinline__ void done() { wp = pos + n_new; }
};

template <int NM, bool RL, int RT, int FAST>
__global__ void __launch_bounds__(SF_THREADS, SF_CTAS_PER_SM)
scan_fused_kernel(const __grid_constant__ SfTab tb, const __grid_constant__ DevChain ch, const __grid_constant__ SfGeom g,
                  DevStream *__restrict__ streams, const uint8_t *__restrict__ in, uint8_t *__restrict__ out,
                  const uint32_t *__restrict__ ckpt_l, const uint32_t *__restrict__ ckpt_r,
                  uint32_t *__restrict__ tail_l, uint32_t *__restrict__ tail_r)
{
    constexpr int WPS = RT == ICW_RENDER_TPDF ? 4 : RT == ICW_RENDER_RPDF ? 2 : 0;
    constexpr int WPSD = WPS ? WPS : 1;                     // divisor that exists for every instantiation
    constexpr int GEN_WORDS = sf_gen_words(WPS);
    extern __shared__ __align__(128) uint8_t sf_smem[];
    // which hardware warps play which role: a sub-partition's arbiter serves the warp with the highest id first
    // (B300_MICROARCH.md "multi-warp arbiter"), so the order decides whose instructions wait when two roles want the same slot.
    // Measured (profiles/r2_ab_warp_roles.txt): the four orders are within 1.7 % of each other, order 0 is the fastest.
#if ICW_SF_ORDER == 0
    const int tid = threadIdx.x;                                                    // scan | gen | pw
#elif ICW_SF_ORDER == 1
    const int tid = (threadIdx.x + SF_SCAN_THREADS) % SF_THREADS;                   // gen | pw | scan
#elif ICW_SF_ORDER == 2
    const int tid = (threadIdx.x + SF_SCAN_THREADS + SF_GEN_THREADS) % SF_THREADS;  // pw | scan | gen
#else
    const int tid = threadIdx.x < SF_PW_THREADS ? SF_PW0 + threadIdx.x              // pw | gen | scan
                  : threadIdx.x < SF_PW_THREADS + SF_GEN_THREADS ? threadIdx.x : threadIdx.x - (SF_PW_THREADS + SF_GEN_THREADS);
#endif
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
    const int n_ranges = n_warm + n_real;

    double *carry = reinterpret_cast<double *>(sf_smem + SfSmem::CARRY);
    const uint32_t bar0 = smem_u32(sf_smem + SfSmem::BAR);
    const int fb = ch.frame_bytes;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar0));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar0 + 8));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 4 * SCAN_NMAX * 2; i += SF_THREADS) {         // the running state, plain modal basis
        const int cf = i / (SCAN_NMAX * 2), j = i % (SCAN_NMAX * 2);
        carry[i] = Wb == 0 ? st.hb[cf >> 1][cf & 1][j] : 0.0;
    }
    __syncthreads();                                        // the only time the whole CTA meets

    // the states around the call's last block go back to the host's bookkeeping (MtPlan::tail)
    const int64_t tailw = g.first_word + g.tail_block * ICW_MT_N;       // first word of that block
    const bool has_tail = WPS && g.tail_block >= (int64_t)unit * g.blocks_per_unit && g.tail_block < (int64_t)(unit + 1) * g.blocks_per_unit;
    auto words_end = [&](int64_t F, int len) {              // stream word behind the range's last one
        int64_t end = g.want_lo + (F + len) * WPS;
        if (has_tail && F + len == g.n_frames) end = tailw + ICW_MT_N;  // through the end of the call's last block
        return end;
    };

    if (tid < SF_SCAN_THREADS) {
        // ==========================================================================================================
        // SCAN
        // ==========================================================================================================
        double2 *et = reinterpret_cast<double2 *>(sf_smem + SfSmem::ET);
        const int half = tid / SF_HALF, hid = tid % SF_HALF;               // half: warp-uniform
        const int c = hid >> 1, chan = hid & 1;
        const int cb = ch.chan_bytes, stereo = ch.n_channels > 1;
        auto range_of = [&](int r, int64_t &F, int &len) {  // r < n_warm: warm-up ranges, then the unit's own
            if (r < n_warm) { F = Wb + (int64_t)r * SF_R; const int64_t l = U0 - F; len = (int)(l < SF_R ? l : SF_R); }
            else { F = U0 + (int64_t)(r - n_warm) * SF_R; const int64_t l = U1 - F; len = (int)(l < SF_R ? l : SF_R); }
        };
        // bytes of frames [F, F + len) -> stage buffer r & 1 by one bulk copy (16-byte granules around the span); the call's
        // last range -- nothing may be read past the caller's buffer -- is read with plain loads when its turn comes
        auto bulk_ok = [&](int64_t F, int len) { return F + len < g.n_frames; };
        auto issue_fetch = [&](int r) {
            int64_t F; int len;
            range_of(r, F, len);
            if (!bulk_ok(F, len)) return;
            const uint8_t *src = in + F * fb;
            const uint32_t a = (uint32_t)((uintptr_t)src & 15u);
            const uint32_t bytes = (a + (uint32_t)len * (uint32_t)fb + 15u) & ~15u;
            const uint32_t bar = bar0 + 8u * (uint32_t)(r & 1);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(smem_u32(sf_smem + SfSmem::STAGE + (r & 1) * SF_STAGE)), "l"(src - a), "r"(bytes), "r"(bar) : "memory");
        };
        if (tid == 0) {
            issue_fetch(0);
            if (n_ranges > 1) issue_fetch(1);
        }
        for (int r = 0; r < n_ranges; ++r) {
            int64_t F; int len;
            range_of(r, F, len);
            const bool real = r >= n_warm;
            uint8_t *stage = sf_smem + SfSmem::STAGE + (r & 1) * SF_STAGE;
            const uint32_t a = (uint32_t)((uintptr_t)(in + F * fb) & 15u);
            // ---- input ---------------------------------------------------------------------------------------------------------
            if (bulk_ok(F, len)) {
                mbar_wait(bar0 + 8u * (uint32_t)(r & 1), (uint32_t)(r >> 1) & 1u);
            } else {
                const uint8_t *src = in + F * fb;
                const int nb = len * fb;
                for (int i = tid; i < nb; i += SF_SCAN_THREADS) stage[a + i] = src[i];
                nb_sync(NB_SCAN, SF_SCAN_THREADS);
            }
            double x[SF_LC];
            {
                const uint32_t sb = smem_u32(stage) + a + (uint32_t)(c * SF_LC * fb + (stereo ? chan * cb : 0));
                const int nv = len - c * SF_LC;
                switch (ch.fmt) {
                case ICW_FMT_WAV_U8:  sf_load_x_any<ICW_FMT_WAV_U8>(sb, fb, nv, x); break;
                case ICW_FMT_WAV_I16: sf_load_x_any<ICW_FMT_WAV_I16>(sb, fb, nv, x); break;
                case ICW_FMT_WAV_I24: sf_load_x_any<ICW_FMT_WAV_I24>(sb, fb, nv, x); break;
                case ICW_FMT_WAV_I32: sf_load_x_any<ICW_FMT_WAV_I32>(sb, fb, nv, x); break;
                default:              sf_load_x_any<ICW_FMT_WAV_F32>(sb, fb, nv, x); break;
                }
            }
            // ---- level 0 -------------------------------------------------------------------------------------------------------
            const unsigned qr = (st.quad[chan] + (unsigned)(F & 3)) & 3u;   // mixer phase of the range's (and every chunk's) first frame
            const int xq = (int)(qr & 1u);                                  // the filter fed on even frames of a chunk: I on phases 0 / 2
            const int xslot = xq * SF_HALF + hid, yslot = (1 - xq) * SF_HALF + hid;
            if (half == 0) sf_level0_half<NM, RL, 0, NM / 2>(tb, x, et + xslot, et + yslot);     // the modes are split between the halves
            else           sf_level0_half<NM, RL, NM / 2, NM>(tb, x, et + xslot, et + yslot);
            nb_sync(NB_SCAN, SF_SCAN_THREADS);
            // ---- level 1: carry over the chunks, serial, one thread per (channel, filter, mode) ----------------------------------
            if (tid < 4 * NM) {
                const int m = tid % NM, f4 = tid / NM, s_chan = f4 >> 1, filt = f4 & 1;
                const unsigned q1 = (st.quad[s_chan] + (unsigned)(F & 3)) & 3u;
                const bool is_x = filt == (int)(q1 & 1u);                   // fed on the range's first frame
                const int off = is_x ? 0 : 1;
                const double sg = mixer_sign(filt, (q1 + (unsigned)off) & 3u);
                Cx T; T.re = carry[(f4 * SCAN_NMAX + m) * 2]; T.im = carry[(f4 * SCAN_NMAX + m) * 2 + 1];
                // the state between ranges is the plain modal state as of the frame before; T is the state right after the
                // filter's last input, which for X was two frames ago (exact algebra: S = s / p); then S -> S~
                if (is_x) T = cx_mul(tb.pinv[m][0], tb.pinv[m][1], T);
                T.re *= sg; T.im *= sg;
                const double q8r = tb.q8[m][0], q8i = tb.q8[m][1];
                const int nfull = len / SF_LC, nch = (len + SF_LC - 1) / SF_LC;
                double2 *e = et + m * SF_ESLOTS + filt * SF_HALF + s_chan;
                {
                    const double nq8i = -q8i;
                    double2 *ep = e;
                    int k = 0;
                    for (; k + 4 <= nfull; k += 4, ep += 8) {                   // four chunks a trip: their effects are fetched together
                        const double2 e0 = ep[0], e1 = ep[2], e2 = ep[4], e3 = ep[6];
                        double tr = T.re, ti = T.im, nr, ni;
                        ep[0] = make_double2(tr, ti);                           // state before the chunk
                        nr = fma(q8r, tr, fma(nq8i, ti, e0.x)); ni = fma(q8r, ti, fma(q8i, tr, e0.y)); tr = nr; ti = ni;
                        ep[2] = make_double2(tr, ti);
                        nr = fma(q8r, tr, fma(nq8i, ti, e1.x)); ni = fma(q8r, ti, fma(q8i, tr, e1.y)); tr = nr; ti = ni;
                        ep[4] = make_double2(tr, ti);
                        nr = fma(q8r, tr, fma(nq8i, ti, e2.x)); ni = fma(q8r, ti, fma(q8i, tr, e2.y)); tr = nr; ti = ni;
                        ep[6] = make_double2(tr, ti);
                        nr = fma(q8r, tr, fma(nq8i, ti, e3.x)); ni = fma(q8r, ti, fma(q8i, tr, e3.y));
                        T.re = nr; T.im = ni;
                    }
                    for (; k < nfull; ++k, ep += 2) {
                        const double2 e0 = ep[0];
                        ep[0] = make_double2(T.re, T.im);
                        const double nr = fma(q8r, T.re, fma(nq8i, T.im, e0.x));
                        const double ni = fma(q8r, T.im, fma(q8i, T.re, e0.y));
                        T.re = nr; T.im = ni;
                    }
                }
                int n_in = 0, lc = SF_LC;
                if (nfull < nch) {                                          // a ragged last chunk: input by input
                    lc = len - nfull * SF_LC;
                    e[2 * nfull] = make_double2(T.re, T.im);
                    n_in = lc > off ? (lc - off + 1) >> 1 : 0;
                    const uint8_t *sp = stage + a + (nfull * SF_LC + off) * fb + (stereo ? s_chan * cb : 0);
                    for (int j = 0; j < n_in; ++j) cx_step(T, tb.qt[m][0], tb.qt[m][1], -unpack_real(ch.fmt, sp + 2 * j * fb, 0));
                }
                // S~ -> S (the sign flips with every input); a filter whose last input was not the range's last frame has idled one sample
                const double se = (n_in & 1) ? -sg : sg;
                T.re *= se; T.im *= se;
                if (((lc - 1 - off) & 1) != 0) T = cx_mul(tb.p[m][0], tb.p[m][1], T);
                carry[(f4 * SCAN_NMAX + m) * 2] = T.re; carry[(f4 * SCAN_NMAX + m) * 2 + 1] = T.im;
            }
            nb_sync(NB_SCAN, SF_SCAN_THREADS);
            if (tid == 0 && r + 2 < n_ranges) issue_fetch(r + 2);           // this range's stage buffer is free again
            if (!real) continue;                                            // warm-up: only the state moves on
            // ---- level 2: the chunk's frames from its start states, into the analytic buffer ------------------------------------------
            const int rr = r - n_warm, b = rr & 1;
            double u[SF_NI], v[SF_NI];
            if (half == 0) sf_level2_a<NM, RL>(tb, x, et + xslot, et + yslot, u, v);       // u = A of X (re of even frames), v = A of Y (re of odd frames)
            else           sf_level2_b<NM, RL>(tb, x, et + xslot, et + yslot, v, u);       // u = B of Y (im of even frames), v = B of X (im of odd frames)
            if (rr >= 2) nb_sync(NB_EMPTY + b, SF_THREADS);                 // the pointwise warps are done with the buffer's last content
            uint8_t *pl = sf_smem + SfSmem::AN + b * SF_ABUF + chan * SF_PLANE + c * SF_APITCH + half * 8;
#pragma unroll
            for (int i = 0; i < SF_NI; ++i) {
                *reinterpret_cast<double *>(pl + (2 * i) * 16) = u[i];
                *reinterpret_cast<double *>(pl + (2 * i + 1) * 16) = v[i];
            }
            nb_arrive(NB_FULL + b, SF_THREADS);
        }
        // the stream's filter state after the call's last frame (modal basis, like scan_apply_kernel leaves it)
        if (U1 == g.n_frames && tid < 4 * NM) {
            const int m = tid % NM, f4 = tid / NM;
            st.hb[f4 >> 1][f4 & 1][2 * m] = carry[(f4 * SCAN_NMAX + m) * 2];
            st.hb[f4 >> 1][f4 & 1][2 * m + 1] = carry[(f4 * SCAN_NMAX + m) * 2 + 1];
        }
    } else if (tid < SF_PW0) {
        // ==========================================================================================================
        // GEN: one warp per generator
        // ==========================================================================================================
        const int gen = (tid - SF_SCAN_THREADS) >> 5, lane = tid & 31;
        uint32_t *my_ub = reinterpret_cast<uint32_t *>(sf_smem + SfSmem::WORDS) + gen * GEN_WORDS;   // [624 history | slices], untempered
        uint32_t *my_tail = gen ? tail_r : tail_l;
        SfWordGeom wg;
        wg.gen_pos = w_unit; wg.wp = ICW_MT_N;
        if (WPS) {
            const uint32_t *ck = (gen ? ckpt_r : ckpt_l) + (size_t)unit * ICW_MT_N;     // history of the first fill = the unit's checkpoint
            for (int i = lane; i < ICW_MT_N; i += 32) my_ub[i] = ck[i];
            __syncwarp();
        }
        for (int rr = 0; rr < n_real; ++rr) {
            const int b = rr & 1;
            const int64_t F = U0 + (int64_t)rr * SF_R;
            const int len = (int)(U1 - F < SF_R ? U1 - F : SF_R);
            if (rr >= 2) nb_sync(NB_EMPTY + b, SF_THREADS);
            if (WPS) {
                wg.next(rr, words_end(F, len));
                if (wg.move) {                              // the last 624 words made become the history in front of the even slice
                    uint32_t keep[(ICW_MT_N + 31) / 32];
#pragma unroll
                    for (int k = 0; k < (ICW_MT_N + 31) / 32; ++k) {
                        const int i = lane + k * 32;
                        keep[k] = i < ICW_MT_N ? my_ub[wg.wp - ICW_MT_N + i] : 0u;
                    }
                    __syncwarp();
#pragma unroll
                    for (int k = 0; k < (ICW_MT_N + 31) / 32; ++k) {
                        const int i = lane + k * 32;
                        if (i < ICW_MT_N) my_ub[i] = keep[k];
                    }
                    __syncwarp();
                }
                for (int done = 0; done < wg.n_new; done += ICW_MT_N) {
                    mt_window_pass<32>(my_ub + wg.pos + done, wg.n_new - done < ICW_MT_N ? wg.n_new - done : ICW_MT_N, lane);
                    __syncwarp();
                }
                if (has_tail) {
                    // buffer position p holds stream word base - pos + p; a 624-word block that lies wholly inside what is there goes out
                    for (int h = 0; h < 2; ++h) {
                        const int64_t p0 = tailw - ICW_MT_N + (int64_t)h * ICW_MT_N - wg.base + wg.pos;
                        if (p0 >= wg.pos - ICW_MT_N && p0 + ICW_MT_N <= wg.pos + wg.n_new)
                            for (int i = lane; i < ICW_MT_N; i += 32) my_tail[h * ICW_MT_N + i] = my_ub[p0 + i];
                    }
                }
                wg.done();
            }
            nb_arrive(NB_FULL + b, SF_THREADS);
        }
    } else {
        // ==========================================================================================================
        // PW
        // ==========================================================================================================
        const int t = tid - SF_PW0;
        const uint32_t *ub_l = reinterpret_cast<const uint32_t *>(sf_smem + SfSmem::WORDS), *ub_r = ub_l + GEN_WORDS;
        SfWordGeom wg;
        wg.gen_pos = w_unit; wg.wp = ICW_MT_N; wg.pos = ICW_MT_N; wg.base = 0;
        FrameAcc acc;
        OscCounter osc;
        osc.init(ch, st.n_frame, U0 + t < g.n_frames ? U0 + t : 0);
        for (int rr = 0; rr < n_real; ++rr) {
            const int b = rr & 1;
            const int64_t F = U0 + (int64_t)rr * SF_R;
            const int len = (int)(U1 - F < SF_R ? U1 - F : SF_R);
            if (WPS) wg.next(rr, words_end(F, len));
            const uint8_t *an = sf_smem + SfSmem::AN + b * SF_ABUF;
            nb_sync(NB_FULL + b, SF_THREADS);               // the range's analytic frames and generator words are there
            if (FAST && len == SF_R && F + len < g.n_frames) {
                // a whole range that does not end the call: the thread's four frames side by side, no guards -- a frame is a long
                // chain of dependent FP64 operations
                uint64_t nvc = osc.at(ch, F + t);
#pragma unroll 1
                for (int k2 = 0; k2 < SF_FPT; k2 += SF_PW_TRIP) {           // SF_PW_TRIP frames a trip: a loop the instruction cache can hold
                    uint64_t nv[SF_PW_TRIP];
#pragma unroll
                    for (int k = 0; k < SF_PW_TRIP; ++k) {
                        nv[k] = nvc;
                        nvc += SF_PW_THREADS;                               // scaled counter (lean_fast_ok): wraps at scale_sr > SF_PW_THREADS
                        if (nvc >= ch.scale_sr) nvc -= ch.scale_sr;
                    }
#pragma unroll
                    for (int k = 0; k < SF_PW_TRIP; ++k) {
                        const int f = t + (k2 + k) * SF_PW_THREADS;
                        const uint8_t *pa = an + (f >> 4) * SF_APITCH + (f & 15) * 16;
                        const double2 a0 = *reinterpret_cast<const double2 *>(pa), a1 = *reinterpret_cast<const double2 *>(pa + SF_PLANE);
                        const int64_t i = F + f;
                        uint4 wl = make_uint4(0u, 0u, 0u, 0u), wr = wl;
                        if (WPS) sf_words<WPS>(ub_l, ub_r, wg.pos + (int)(g.want_lo + F * WPS - wg.base) + f * WPS, wl, wr);
                        const double v[4] = { a0.x, a0.y, a1.x, a1.y };
                        lean_frame_fast_at<RT, FRAME_NO_LAST>(ch, st, i, -1, nv[k], v, wl, wr, out, acc);
                    }
                }
                osc.frame = F + t + (SF_R - SF_PW_THREADS);
                osc.value = nvc >= SF_PW_THREADS ? nvc - SF_PW_THREADS : nvc + ch.scale_sr - SF_PW_THREADS;      // the last frame done
            } else {
#pragma unroll 1
                for (int k = 0; k < SF_FPT; ++k) {
                    const int f = t + k * SF_PW_THREADS;
                    if (f < len) {
                        const uint8_t *pa = an + (f >> 4) * SF_APITCH + (f & 15) * 16;
                        const double2 a0 = *reinterpret_cast<const double2 *>(pa), a1 = *reinterpret_cast<const double2 *>(pa + SF_PLANE);
                        const int64_t i = F + f;
                        uint4 wl = make_uint4(0u, 0u, 0u, 0u), wr = wl;
                        if (WPS) sf_words<WPS>(ub_l, ub_r, wg.pos + (int)(g.want_lo + i * WPS - wg.base), wl, wr);
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
            }
            if (WPS) wg.done();
            if (rr + 2 < n_real) nb_arrive(NB_EMPTY + b, SF_THREADS);       // somebody will fill this buffer again
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

int64_t sfused_warm_frames(const SfTab &tb)
{
    double rmax = 0.0;
    for (int m = 0; m < tb.nm; ++m) {
        const double r = std::sqrt(tb.p[m][0] * tb.p[m][0] + tb.p[m][1] * tb.p[m][1]);
        if (r > rmax) rmax = r;
    }
    const double need = rmax > 0.0 && rmax < 1.0 ? std::log(1e-19) / std::log(rmax) : 1e9;
    const int64_t ranges = (int64_t)(need / SF_R) + 1;
    return ranges * SF_R;
}

template <int NM, bool RL, int RT, int FAST>
static cudaError_t sf_launch(const SfTab &tb, const DevChain &ch, const SfGeom &g, DevStream *streams, const uint8_t *in,
                             uint8_t *out, const MtPlan *pl, const MtPlan *pr, cudaStream_t s)
{
    const int wps = RT == ICW_RENDER_TPDF ? 4 : RT == ICW_RENDER_RPDF ? 2 : 0;
    const size_t smem = sfused_smem_bytes(wps);
    cudaError_t e = cudaFuncSetAttribute(scan_fused_kernel<NM, RL, RT, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    scan_fused_kernel<NM, RL, RT, FAST><<<g.n_units, SF_THREADS, smem, s>>>(tb, ch, g, streams, in, out,
                                                                            pl ? pl->ckpt : nullptr, pr ? pr->ckpt : nullptr,
                                                                            pl ? pl->tail : nullptr, pr ? pr->tail : nullptr);
    return cudaGetLastError();
}

template <int NM, bool RL>
static cudaError_t sf_launch_nm(const SfTab &tb, const DevChain &ch, const SfGeom &g, DevStream *streams, const uint8_t *in,
                                uint8_t *out, const MtPlan *pl, const MtPlan *pr, bool fast, cudaStream_t s)
{
#define ICW_SF(RT) (fast ? sf_launch<NM, RL, RT, 1>(tb, ch, g, streams, in, out, pl, pr, s) : sf_launch<NM, RL, RT, 0>(tb, ch, g, streams, in, out, pl, pr, s))
    switch (ch.render.render_type) {
    case ICW_RENDER_ROUND: return ICW_SF(ICW_RENDER_ROUND);
    case ICW_RENDER_RPDF:  return ICW_SF(ICW_RENDER_RPDF);
    case ICW_RENDER_TPDF:  return ICW_SF(ICW_RENDER_TPDF);
    default: return cudaErrorInvalidValue;
    }
#undef ICW_SF
}

cudaError_t launch_scan_fused(const SfTab &tb, const DevChain &ch, DevStream *streams, int64_t n_frames, const uint8_t *in,
                              uint8_t *out, const MtPlan *pl, const MtPlan *pr, int n_cta, int64_t warm, cudaStream_t s)
{
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
    switch (tb.nm) {
    case 8:  return sf_launch_nm<8, true>(tb, ch, g, streams, in, out, pl, pr, fast, s);
    case 9:  return sf_launch_nm<9, false>(tb, ch, g, streams, in, out, pl, pr, fast, s);
    case 10: return tb.real_last ? sf_launch_nm<10, true>(tb, ch, g, streams, in, out, pl, pr, fast, s)
                                : sf_launch_nm<10, false>(tb, ch, g, streams, in, out, pl, pr, fast, s);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace icw
