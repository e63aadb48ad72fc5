// icw_chainmt.cu -- the pointwise chain with the dither generator inside it.
//
//   chain_mt_kernel   MT19937 block regeneration (both channels' generators) into shared memory,
//                     then unpack -> oscillator -> DSP list -> dither + quantise -> PCM for the
//                     frames those words belong to; tile after tile along the CTA's own span.
//
// chain_kernel reads its dither from a buffer that mt_words_kernel wrote: 8 (RPDF) or 16 (TPDF)
// bytes per channel-sample go to HBM and come back, 64 of the ~140 bytes per frame the C2 workload
// moved.  Here a CTA owns one "unit" of the generator's stream -- a run of 2^k consecutive
// 624-word blocks that starts at a jump-ahead checkpoint (icw_mt.cu) -- and the frames whose draws
// fall into it (the draw index is a closed form of the frame index: reference
// src/sound_render.c:711-751 draws 2 or 4 words per channel-sample, in order).  The words never
// leave the SM.  Block regeneration is mt_words_block (icw_mtdev.cuh), the frame arithmetic is
// the same code chain_kernel runs (icw_frame.cuh): the bytes are identical by construction and
// tests/test_gpu_parity.py compares both routes with the oracle.
//
// Straight-line DSP lists (Master alone, Shift -> Master) with plain PCM output get a lean frame
// path: list shape and dither type are template parameters, no thread-private bus.
#include <type_traits>

#include "icw_dev.cuh"
#include "icw_frame.cuh"
#include "icw_kernels.h"
#include "icw_mtdev.cuh"

namespace icw {

constexpr int CMT_THREADS = 256;
constexpr int CMT_CTAS = 4;                         // resident per SM
constexpr int CMT_TB = 8;                           // blocks per tile: 1248 TPDF frames, 2496 RPDF frames
constexpr int CMT_GEN_WORDS = (CMT_TB + 1) * ICW_MT_N;     // per generator: 624 words of history + the tile's new words, untempered
constexpr size_t CMT_SMEM = (size_t)(2 * CMT_GEN_WORDS) * sizeof(uint32_t);   // 44 928 B

int chain_mt_max_units(int sm_count) { return sm_count * CMT_CTAS; }

struct CmtGeom {
    int n_units, blocks_per_unit;
    int64_t first_word, want_lo, want_hi, tail_block;
};

// SHAPE == ICW_SHAPE_GENERIC: any list, taps, noise-shaping hand-off -- finish_frame() itself.
template <int SHAPE, int RT>
__global__ void __launch_bounds__(CMT_THREADS, CMT_CTAS)
chain_mt_kernel(const __grid_constant__ DevChain ch, DevStream *__restrict__ streams, int64_t n_frames,
                const uint8_t *__restrict__ in, int from_analytic, const __grid_constant__ CmtGeom g,
                const uint32_t *__restrict__ ckpt_l, const uint32_t *__restrict__ ckpt_r,
                uint32_t *__restrict__ tail_l, uint32_t *__restrict__ tail_r,
                uint8_t *__restrict__ out, double *__restrict__ tap_bus, double *__restrict__ tap_lr, double *__restrict__ pre)
{
    constexpr int WPS = RT == ICW_RENDER_TPDF ? 4 : 2;
    constexpr bool LEAN = SHAPE != ICW_SHAPE_GENERIC;
    extern __shared__ __align__(16) uint32_t cmt_smem[];
    uint32_t *ub_l = cmt_smem;                                  // [624 history + CMT_TB * 624] the left generator's stream, untempered
    uint32_t *ub_r = cmt_smem + CMT_GEN_WORDS;                  // the right generator's
    const int t = threadIdx.x;
    const int unit = blockIdx.x;
    DevStream &st = streams[0];
    // threads 0..127 make the left generator's words, 128..255 the right one's: 128 words of each per pass (icw_mtdev.cuh)
    const int gen = t >> 7, lane_w = t & 127;
    uint32_t *my_ub = gen ? ub_r : ub_l;

    for (int i = t; i < ICW_MT_N; i += CMT_THREADS) {           // history of the first tile = the unit's checkpoint
        ub_l[i] = ckpt_l[(size_t)unit * ICW_MT_N + i];
        ub_r[i] = ckpt_r[(size_t)unit * ICW_MT_N + i];
    }
    // block b of this unit holds stream words [u0 + 624 b, u0 + 624 (b + 1))
    const int64_t u0 = g.first_word + (int64_t)unit * g.blocks_per_unit * ICW_MT_N;
    int64_t nb64 = (g.want_hi - u0 + ICW_MT_N - 1) / ICW_MT_N;                      // blocks that start below want_hi
    const int n_blk = (int)(nb64 < 0 ? 0 : nb64 > g.blocks_per_unit ? g.blocks_per_unit : nb64);
    const int64_t tl = g.tail_block - (int64_t)unit * g.blocks_per_unit;
    const int tail_blk = (tl >= 0 && tl < n_blk) ? (int)tl : -1;
    __syncthreads();

    FrameIO io;
    io.mtw_l = io.mtw_r = nullptr;
    io.dst = out;
    io.dst_aligned = ((size_t)(uintptr_t)out & 3u) == 0;
    io.tap_bus = tap_bus; io.tap_lr = tap_lr; io.pre = pre;
    FrameAcc acc;
    double bus[LEAN ? 1 : ICW_N_PLUGS][4];
    if (!LEAN) load_bus(st, bus);
    OscCounter osc;
    {
        int64_t f = (u0 - g.want_lo) / WPS;
        f = (f < 0 ? 0 : f) + t;
        osc.init(ch, st.n_frame, f < n_frames ? f : 0);
    }
    const int64_t pos0 = st.pos;

    for (int b0 = 0; b0 < n_blk; b0 += CMT_TB) {
        const int tb = n_blk - b0 < CMT_TB ? n_blk - b0 : CMT_TB;
        // ---- the tile's words: both generators, 128 words each per pass, one barrier per pass ----------------
        {
            uint32_t *u = my_ub + ICW_MT_N + lane_w;
            const int n_new = tb * ICW_MT_N;
            const int full = n_new >> 7, rag = n_new & 127;      // 624 tb is not a multiple of 128: one ragged pass
            for (int k = 0; k < full; ++k, u += 128) {
                mt_window_word(u);
                __syncthreads();
            }
            if (rag) {
                if (lane_w < rag) mt_window_word(u);
                __syncthreads();
            }
            // (a barrier per generator -- bar.sync over its 128 threads -- and 227 words between barriers were both
            // measured: 13.4 and 14.6 ms per C2 step against 13.3 for this form; the phase is not barrier-bound)
        }
        // the states around the call's last block go back to the host's bookkeeping (MtPlan::tail): the block before
        // it (or the history, when it is the tile's first) and the block itself
        if (tail_blk >= b0 && tail_blk < b0 + tb) {
            const uint32_t *sl = ub_l + (tail_blk - b0) * ICW_MT_N, *sr = ub_r + (tail_blk - b0) * ICW_MT_N;
            for (int i = t; i < 2 * ICW_MT_N; i += CMT_THREADS) { tail_l[i] = sl[i]; tail_r[i] = sr[i]; }
        }
        // ---- the frames whose draws are those words -----------------------------------------------------
        const int64_t w_lo = u0 + (int64_t)b0 * ICW_MT_N;       // stream word held by wd[0]
        int64_t i_lo = w_lo - g.want_lo;
        i_lo = i_lo < 0 ? 0 : i_lo / WPS;
        int64_t i_hi = (w_lo + (int64_t)tb * ICW_MT_N - g.want_lo) / WPS;
        if (i_hi > n_frames) i_hi = n_frames;
        // (requesting a thread's next analytic frame one iteration ahead was measured slower: the eight
        // extra live registers spill at 64 per thread)
        const double2 *ana = reinterpret_cast<const double2 *>(in);
        // one frame from the tile's words; LAST as in icw_frame.cuh (the call's last frame leaves the bus state behind)
        auto frame = [&](int64_t i, auto last_tag) {
            constexpr int LAST = decltype(last_tag)::value;
            const uint32_t off = (uint32_t)(g.want_lo + i * WPS - w_lo);
            uint4 wl = make_uint4(0u, 0u, 0u, 0u), wr = wl;
            if (WPS == 4) {
                wl = *reinterpret_cast<const uint4 *>(ub_l + ICW_MT_N + off);
                wr = *reinterpret_cast<const uint4 *>(ub_r + ICW_MT_N + off);
                wl.z = mt_temper_mul(wl.z); wl.w = mt_temper_mul(wl.w); wr.z = mt_temper_mul(wr.z); wr.w = mt_temper_mul(wr.w);
            } else {
                const uint2 a = *reinterpret_cast<const uint2 *>(ub_l + ICW_MT_N + off), b = *reinterpret_cast<const uint2 *>(ub_r + ICW_MT_N + off);
                wl.x = a.x; wl.y = a.y; wr.x = b.x; wr.y = b.y;
            }
            wl.x = mt_temper_mul(wl.x); wl.y = mt_temper_mul(wl.y); wr.x = mt_temper_mul(wr.x); wr.y = mt_temper_mul(wr.y);
            double v[4];
            if (SHAPE == ICW_SHAPE_SHIFT_MASTER_FAST || from_analytic) {
                const double2 a0 = ana[i * 2], a1 = ana[i * 2 + 1];
                v[0] = a0.x; v[1] = a0.y; v[2] = a1.x; v[3] = a1.y;
            } else {
                unpack_frame(ch, in + i * ch.frame_bytes, pos0 + i, v);
            }
            if (SHAPE == ICW_SHAPE_SHIFT_MASTER_FAST) lean_frame_fast<RT, LAST>(ch, st, i, n_frames - 1, v, wl, wr, io.dst, acc, osc);
            else if (LEAN) lean_frame<SHAPE, RT, LAST>(ch, st, i, n_frames - 1, v, wl, wr, io.dst, io.dst_aligned, acc, osc);
            else if (LAST != FRAME_LAST_ONLY) finish_frame<DITHER_GIVEN>(ch, st, i, n_frames, v, bus, io, acc, osc, wl, wr);
        };
        for (int64_t i = i_lo + t; i < i_hi; i += CMT_THREADS) frame(i, std::integral_constant<int, LEAN ? FRAME_NO_LAST : FRAME_CHECK_LAST>());
        // the call's last frame again, for the bus state it leaves in the context (the words are still in the tile)
        if (LEAN && i_hi == n_frames && i_lo < i_hi && t == (int)((n_frames - 1 - i_lo) % CMT_THREADS))
            frame(n_frames - 1, std::integral_constant<int, FRAME_LAST_ONLY>());
        __syncthreads();                                        // every frame of the tile has its words
        if (b0 + CMT_TB < n_blk) {                              // the tile's last block becomes the next tile's history
            for (int i = t; i < ICW_MT_N; i += CMT_THREADS) {
                ub_l[i] = ub_l[tb * ICW_MT_N + i];
                ub_r[i] = ub_r[tb * ICW_MT_N + i];
            }
            __syncthreads();
        }
    }
    commit_acc(&st, acc, 32);
}

template <int SHAPE, int RT>
static cudaError_t launch_cmt(const DevChain &ch, DevStream *streams, int64_t n_frames, const uint8_t *in, int from_analytic,
                              const CmtGeom &g, const MtPlan &pl, const MtPlan &pr, uint8_t *out, double *tap_bus, double *tap_lr,
                              double *pre, cudaStream_t s)
{
    // per device, not per process: set on every launch (a few microseconds) rather than cached in a static
    cudaError_t e = cudaFuncSetAttribute(chain_mt_kernel<SHAPE, RT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CMT_SMEM);
    if (e != cudaSuccess) return e;
    chain_mt_kernel<SHAPE, RT><<<g.n_units, CMT_THREADS, CMT_SMEM, s>>>(ch, streams, n_frames, in, from_analytic, g, pl.ckpt, pr.ckpt,
                                                                      pl.tail, pr.tail, out, tap_bus, tap_lr, pre);
    return cudaGetLastError();
}

bool chain_mt_supports(const DevChain &ch)
{
    return ch.render.render_type == ICW_RENDER_RPDF || ch.render.render_type == ICW_RENDER_TPDF;
}

cudaError_t launch_chain_mt(const DevChain &ch, DevStream *streams, int64_t n_frames, const uint8_t *in, int from_analytic,
                            const MtPlan &pl, const MtPlan &pr, uint8_t *out, double *tap_bus, double *tap_lr, double *pre,
                            cudaStream_t s)
{
    if (pl.n_units != pr.n_units || pl.blocks_per_unit != pr.blocks_per_unit || pl.first_word != pr.first_word ||
        pl.want_lo != pr.want_lo || pl.want_hi != pr.want_hi)
        return cudaErrorInvalidValue;                           // the two generators must stand at the same draw
    CmtGeom g;
    g.n_units = pl.n_units; g.blocks_per_unit = pl.blocks_per_unit;
    g.first_word = pl.first_word; g.want_lo = pl.want_lo; g.want_hi = pl.want_hi; g.tail_block = pl.tail_block;
    const bool lean = !tap_bus && !tap_lr && !pre && ch.shape != ICW_SHAPE_GENERIC && !ch.bypass && !ch.fp_check;
    const bool tpdf = ch.render.render_type == ICW_RENDER_TPDF;
#define ICW_CMT(SH, RT) return launch_cmt<SH, RT>(ch, streams, n_frames, in, from_analytic, g, pl, pr, out, tap_bus, tap_lr, pre, s)
    if (lean && from_analytic && lean_fast_ok(ch) && ((size_t)(uintptr_t)out & 3u) == 0) {
        if (tpdf) ICW_CMT(ICW_SHAPE_SHIFT_MASTER_FAST, ICW_RENDER_TPDF);
        ICW_CMT(ICW_SHAPE_SHIFT_MASTER_FAST, ICW_RENDER_RPDF);
    }
    if (lean && ch.shape == ICW_SHAPE_SHIFT_MASTER) { if (tpdf) ICW_CMT(ICW_SHAPE_SHIFT_MASTER, ICW_RENDER_TPDF); ICW_CMT(ICW_SHAPE_SHIFT_MASTER, ICW_RENDER_RPDF); }
    if (lean && ch.shape == ICW_SHAPE_MASTER) { if (tpdf) ICW_CMT(ICW_SHAPE_MASTER, ICW_RENDER_TPDF); ICW_CMT(ICW_SHAPE_MASTER, ICW_RENDER_RPDF); }
    if (tpdf) ICW_CMT(ICW_SHAPE_GENERIC, ICW_RENDER_TPDF);
    ICW_CMT(ICW_SHAPE_GENERIC, ICW_RENDER_RPDF);
#undef ICW_CMT
}

}  // namespace icw
