// icw_split.cu -- the reference-exact chain (Kahan summation, the reference's default: src/hblpf.c:974-1057) for
// batches of real-input streams, with the two sums of the recurrence on DIFFERENT warps:
//
//     helper warps : unpack + fade tile s                 -> XS          (src/xwave_reader.c:908-1009)
//     state warps  : state sum of tile s-1                -> W           (hblpf.c:1008-1046: w = x + sum z[i]*fb[i])
//     output warps : output sum of tile s-2 from W        -> YS          (hblpf.c:1013-1056: y = sum z[i]*ff[i], (z[i]*fb[i])*d0)
//     helper warps : YS of tile s-3 -> un-mix -> oscillator -> DSP list -> quantise -> PCM
//
// Only the state sum is serial in time: 19 compensated additions = 76 dependent DADDs = 608 cycles a sample (order 19),
// and nothing can shorten it without changing the reference's rounding.  The output sum (148 dependent DADDs) reads the
// same states but feeds nothing back, so it runs a tile behind on another warp, FOUR consecutive samples of a
// recurrence advanced in lock step by one thread (four independent dependency chains keep the FP64 pipe fed from one
// warp; the 22 states they read are loaded once).  The state warps' loop is then the critical sum and nothing else
// (95 FP64 instructions a sample, 190 of the 608 cycles of its sub-partition's FP64 pipe); the hardware scheduler fills
// the gaps with the output and helper warps (icw_fused.cu interleaves the three sums in ONE instruction stream at
// compile time and reaches 61 % of the pipe).  Every operation and its order inside each sum is the reference's; which
// warp performs it rounding cannot see.
//
// Who shares an FP64 pipe with the state warps decides the run time: a dependent DADD that finds the pipe taken by
// another warp's instruction waits, 76 times a sample.  Measured with one state, one output and one helper warp on every
// sub-partition (pipe 65 % busy): 11.6 cycles per dependent DADD instead of 8, 1020 cycles a sample.  So the four state
// warps of a CTA (32 streams, 128 recurrences) sit ALONE on sub-partition 0 (warp ids 0, 4, 8, 12: a warp runs on
// sub-partition id mod 4): four chains of the same 8-cycle period, 2 pipe cycles each, fall into step with one another,
// and the pipe of that sub-partition carries nothing else.  Sub-partitions 1-3 carry the throughput work: on each, four
// output warps (32 recurrences each, every third pair of samples of a tile, the pair advanced in lock step) and one
// helper warp (3 threads per stream).
//
// W holds the states w[n] as rows of 4 NS doubles, three tile buffers, each preceded by the last HIST rows of the tile
// before it (written twice by the state warp), so that both consumers read w[n-1-i] at compile-time offsets from one
// row pointer: no circular-buffer arithmetic, no register moves (the z[] shift costs icw_fused.cu 40 MOVs a sample).
#include <cstdlib>
#include "icw_dev.cuh"
#include "icw_kernels.h"
#include "icw_hb.cuh"
#include "icw_frame.cuh"

namespace icw {

#ifndef ICW_SPLIT_T
#define ICW_SPLIT_T 32
#endif
#ifndef ICW_SPLIT_K
#define ICW_SPLIT_K 1
#endif
#ifndef ICW_SPLIT_PH
#define ICW_SPLIT_PH 2
#endif
#ifndef ICW_SPLIT_HPS
#define ICW_SPLIT_HPS 8
#endif
// timing experiments only (wrong results): leave one role's work out
#ifndef ICW_SPLIT_SKIP
#define ICW_SPLIT_SKIP 0
#endif
constexpr bool SP_DEBUG_SKIP_STATE = (ICW_SPLIT_SKIP & 1) != 0, SP_DEBUG_SKIP_OUT = (ICW_SPLIT_SKIP & 2) != 0,
               SP_DEBUG_SKIP_UNPACK = (ICW_SPLIT_SKIP & 4) != 0, SP_DEBUG_SKIP_RENDER = (ICW_SPLIT_SKIP & 8) != 0;
constexpr int SP_T = ICW_SPLIT_T;               // frames per tile
constexpr int SP_K = ICW_SPLIT_K;               // samples of a recurrence whose output sums one thread advances in lock step (1: see DESIGN.md 5.1)
constexpr int SP_PH = ICW_SPLIT_PH;             // output lanes per recurrence: each takes every PH-th group of K samples
constexpr int SP_HPS = ICW_SPLIT_HPS;           // helper threads per stream: a frame is ~3000 cycles of dependent work (more with dither and a
                                                // noise shaper's hand-off: 4 a stream is enough for C4 and costs c4ns 25 %).  18 warps at 28
                                                // streams; the SM allots registers to warps in fours, so 96 a thread is the ceiling
constexpr int SP_HIST = 20;                     // rows of history in front of a tile (>= the highest order)
constexpr int SP_WROWS = SP_HIST + SP_T;
static_assert(SP_T >= SP_HIST && SP_HIST >= ICW_MAX_ORD && SP_T % (SP_K * SP_PH) == 0 && SP_T % SP_HPS == 0,
              "a tile must cover the history it hands on and divide among the output and helper warps");

// Streams per CTA are a launch-time choice: 28 fills 147 SMs with 4096 streams; a smaller batch takes 14 / 7 / 4 so that
// it still covers the machine -- the FP64 pipe of a sub-partition then carries half / a quarter of the throughput work
// next to its state warp, and the state sum (the run time) loses fewer arbitrations.
template <int NS_>
struct SplitGeom {
    static constexpr int NS = NS_;                      // streams per CTA (<= 32: one lane per recurrence in <= 4 state warps)
    static constexpr int NCH = NS * 4;                  // recurrences per CTA = doubles per row of W / YS
    static constexpr int STATE_WARPS = (NCH + 31) / 32; // warp ids 0 .. : one per sub-partition
    static constexpr int OUT_WARPS = (NCH * SP_PH + 31) / 32;   // output tasks (recurrence, phase) packed into full warps
    static constexpr int HELP_WARPS = (NS * SP_HPS + 31) / 32;
    static constexpr int THREADS = 32 * (STATE_WARPS + OUT_WARPS + HELP_WARPS);
    static constexpr size_t XS_BYTES = sizeof(double) * 2 * SP_T * NS * 2;
    static constexpr size_t W_BYTES = sizeof(double) * 3 * SP_WROWS * NCH;
    static constexpr size_t YS_BYTES = sizeof(double) * 2 * SP_T * NCH;
    static constexpr size_t SMEM = XS_BYTES + W_BYTES + YS_BYTES;
    static_assert(NS <= 32 && SMEM <= 227 * 1024, "tile buffers exceed the shared memory of an SM");
};

// kind 0 state, 1 output, 2 helper, 3 none; index = which warp of that kind (a warp runs on sub-partition id mod 4: the
// state warps, ids 0.., sit on different ones).  Written as a count over the warps below this one on purpose: the
// closed form (warp - STATE_WARPS ...) lets the compiler prove the role warp-uniform and move everything derived from it
// into uniform registers, which the three tile loops need for their coefficients -- measured 223 ms per C4 step
// against 217 with this form (tools/ab_variants.sh).
struct SplitRole { int kind, index; };
template <class G>
__device__ __forceinline__ SplitRole split_role(int warp)
{
    const int sub = warp & 3, k = warp >> 2;
    if (sub < G::STATE_WARPS && k < 1) return SplitRole{0, sub};
    int n = 0;
    for (int v = 0; v < warp; ++v) n += !((v & 3) < G::STATE_WARPS && (v >> 2) < 1);
    if (n < G::OUT_WARPS) return SplitRole{1, n};
    n -= G::OUT_WARPS;
    if (n < G::HELP_WARPS) return SplitRole{2, n};
    return SplitRole{3, 0};
}

// a helper thread's counters into its stream (commit_acc, icw_frame.cuh, without the shuffle tree: 3 threads a stream)
__device__ __forceinline__ void commit_acc_one(DevStream &st, const FrameAcc &acc)
{
    if (acc.clips_l) atomicAdd(&st.clips[0], acc.clips_l);
    if (acc.clips_r) atomicAdd(&st.clips[1], acc.clips_r);
    if (acc.redraws) atomicAdd(&st.mt_redraws, (unsigned long long)acc.redraws);
    atomicMax(reinterpret_cast<unsigned long long *>(&st.peak[0]), (unsigned long long)__double_as_longlong(acc.peak_l));
    atomicMax(reinterpret_cast<unsigned long long *>(&st.peak[1]), (unsigned long long)__double_as_longlong(acc.peak_r));
}

// K compensated sums advanced in lock step, stage by stage (icw_hb.cuh: comp_add3)
template <int K>
__device__ __forceinline__ void comp_add_k(Comp (&o)[K], const double (&x)[K])
{
    double y[K], t[K], d[K];
#pragma unroll
    for (int k = 0; k < K; ++k) y[k] = __dsub_rn(x[k], o[k].c);
#pragma unroll
    for (int k = 0; k < K; ++k) t[k] = __dadd_rn(o[k].s, y[k]);
#pragma unroll
    for (int k = 0; k < K; ++k) d[k] = __dsub_rn(t[k], o[k].s);
#pragma unroll
    for (int k = 0; k < K; ++k) { o[k].c = __dsub_rn(d[k], y[k]); o[k].s = t[k]; }
}

template <int ORD, int NS>
__global__ void __launch_bounds__(SplitGeom<NS>::THREADS, 1)
hb_split_kernel(const __grid_constant__ HbCoef coef, const __grid_constant__ DevChain ch,
                DevStream *__restrict__ streams, int n_streams, int64_t n_frames,
                const uint8_t *__restrict__ in, size_t in_stride,
                const uint32_t *__restrict__ mtw_l, const uint32_t *__restrict__ mtw_r, size_t mt_stream_stride,
                uint8_t *__restrict__ out, size_t out_stride,
                double *__restrict__ tap_bus, double *__restrict__ tap_lr, double *__restrict__ pre, int fast)
{
    using G = SplitGeom<NS>;
    extern __shared__ __align__(16) unsigned char split_smem[];
    double (*xs)[SP_T][G::NS * 2] = reinterpret_cast<double (*)[SP_T][G::NS * 2]>(split_smem);
    double (*wb)[SP_WROWS][G::NCH] = reinterpret_cast<double (*)[SP_WROWS][G::NCH]>(split_smem + G::XS_BYTES);
    double (*ys)[SP_T][G::NCH] = reinterpret_cast<double (*)[SP_T][G::NCH]>(split_smem + G::XS_BYTES + G::W_BYTES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const SplitRole role = split_role<G>(warp);
    const int stream0 = blockIdx.x * G::NS;
    const int64_t n_tiles = (n_frames + SP_T - 1) / SP_T;

    // ---- state-warp set-up: lane = one recurrence ---------------------------------------------------------------
    const int c_local = role.index * 32 + lane;                 // state warps: recurrence index inside the CTA
    const bool is_state = role.kind == 0 && c_local < G::NCH;
    const int c_iq = c_local & 1, c_chan = (c_local >> 1) & 1, c_sl = c_local >> 2;
    const bool c_live = is_state && stream0 + c_sl < n_streams;
    unsigned long long rejects = 0;
    unsigned q0c = 0;
    const double thr = (double)ch.reject_flag;
    if (is_state) {
        // history rows in front of tile 0 = the stream's delay line, newest last (hblpf.c: z[0] is the newest state)
#pragma unroll
        for (int i = 0; i < SP_HIST; ++i) {
            double z = 0.0;
            if (c_live && i < ORD) z = streams[stream0 + c_sl].hb[c_chan][c_iq][i];
            wb[0][SP_HIST - 1 - i][c_local] = z;
        }
        if (c_live) {
            rejects = streams[stream0 + c_sl].hb_rejects[c_chan][c_iq];
            q0c = streams[stream0 + c_sl].quad[c_chan];
        }
    }

    // ---- output-warp set-up: lane = (recurrence, phase): every PH-th group of K samples of that recurrence ----------
    const int o_task = role.index * 32 + lane;
    const bool is_out = role.kind == 1 && o_task < G::NCH * SP_PH;
    const int o_chain = o_task % G::NCH, o_phase = o_task / G::NCH;

    // ---- helper-warp set-up ------------------------------------------------------------------------------------------
    const int h = role.index * 32 + lane;
    const int h_sl = h / SP_HPS, h_part = h % SP_HPS;
    const bool h_live = role.kind == 2 && h_sl < G::NS && stream0 + h_sl < n_streams;
    FrameAcc acc;
    FrameIO io;
    io.mtw_l = io.mtw_r = nullptr; io.dst = nullptr; io.dst_aligned = 0; io.tap_bus = io.tap_lr = nullptr; io.pre = nullptr;
    OscCounter osc;
    osc.frame = 0; osc.value = 0;
    const uint8_t *h_src = nullptr;
    int64_t h_pos0 = 0;
    unsigned h_q0[2] = { 0, 0 };
    double bus[ICW_N_PLUGS][4];
    if (h_live) {
        const int stream = stream0 + h_sl;
        DevStream &st = streams[stream];
        const size_t mt_off = (size_t)stream * mt_stream_stride;
        io.mtw_l = mtw_l ? mtw_l + mt_off : nullptr;
        io.mtw_r = mtw_r ? mtw_r + mt_off : nullptr;
        io.dst = out + (size_t)stream * out_stride;
        io.dst_aligned = ((size_t)(uintptr_t)io.dst & 3u) == 0;
        io.tap_bus = tap_bus ? tap_bus + (size_t)stream * n_frames * (ICW_N_PLUGS * 4) : nullptr;
        io.tap_lr = tap_lr ? tap_lr + (size_t)stream * n_frames * 2 : nullptr;
        io.pre = pre ? pre + (size_t)stream * n_frames * 4 : nullptr;
        h_src = in + (size_t)stream * in_stride;
        h_pos0 = st.pos;
        h_q0[0] = st.quad[0]; h_q0[1] = st.quad[1];
        load_bus(st, bus);
        osc.init(ch, st.n_frame, 0);
    }
    __syncthreads();

    // step s: helpers unpack tile s and render tile s-3, state warps run tile s-1, output warps tile s-2; one CTA barrier
    // a step.  Every role has its own copy of the step loop: the constants each keeps in uniform registers (the state
    // warps the feedback coefficients, the output warps all three sets) then only compete with their own kind.
    const int64_t n_steps = n_tiles + 3;
    if (role.kind == 0) {
        for (int64_t s = 0; s < n_steps; ++s) {
            const int64_t ts = s - 1;
            if (is_state && ts >= 0 && ts < n_tiles && !SP_DEBUG_SKIP_STATE) {
                const int xb = (int)(ts & 1), wbi = (int)(ts % 3), wbn = (int)((ts + 1) % 3);
                const int64_t i0 = ts * SP_T;
                const int len = (int)((n_frames - i0 < SP_T) ? n_frames - i0 : SP_T);
                const int col = c_sl * 2 + c_chan;
                const double *xp = &xs[xb][0][col];
                double *wr = &wb[wbi][SP_HIST][c_local];                // row of the tile's first sample
                double *wn = &wb[wbn][0][c_local] + (SP_HIST - SP_T) * G::NCH;   // the same rows seen from the next buffer's history
                // products of the states the first sample reads, except the newest (its product heads the serial chain)
                double q[ORD];
                double w_prev = wr[-1 * G::NCH];
#pragma unroll
                for (int i = 1; i < ORD; ++i) q[i] = __dmul_rn(wr[(-1 - i) * G::NCH], coef.fb[i]);
                // fs/4 down-mix without branches (lpf_hilbert_quad.c:132-153): the I filter takes +x on phase 0 and -x on
                // phase 2, the Q filter -x on phase 1 and +x on phase 3, zero otherwise
                const unsigned ph_pos = c_iq ? 3u : 0u, ph_neg = c_iq ? 1u : 2u;
                unsigned qd = (q0c + (unsigned)i0) & 3u;
                double x = xp[0];
                x = qd == ph_pos ? x : (qd == ph_neg ? -x : 0.0);
                for (int t = 0; t < len; ++t) {
                    const int tn = t + 1 < len ? t + 1 : t;
                    const double xraw = xp[tn * (G::NS * 2)];           // the next input, a sample ahead
                    // first compensated addition: y = p - 0 is p itself for every p (also -0 and NaN)
                    Comp a;
                    {
                        const double y = __dmul_rn(w_prev, coef.fb[0]);
                        const double t0 = __dadd_rn(x, y);
                        a.c = __dsub_rn(__dsub_rn(t0, x), y);
                        a.s = t0;
                    }
#pragma unroll
                    for (int i = 1; i < ORD; ++i) {
                        comp_add(a, q[i]);
                        // the same term of the NEXT sample: its state is already there (row t - i)
                        q[i] = __dmul_rn(wr[(t - i) * G::NCH], coef.fb[i]);
                    }
                    double w = a.s;
                    const bool rj = (ch.reject_flag != 0) & (fabs(w) < thr);    // |w| < flag value: bug-for-bug (hblpf.c:1046)
                    w = rj ? 0.0 : w;
                    rejects += rj ? 1ull : 0ull;
                    wr[t * G::NCH] = w;
                    if (t >= SP_T - SP_HIST) wn[t * G::NCH] = w;
                    w_prev = w;
                    qd = (qd + 1u) & 3u;
                    x = qd == ph_pos ? xraw : (qd == ph_neg ? -xraw : 0.0);
                }
                if (ts == n_tiles - 1 && c_live) {
                    DevStream &st = streams[stream0 + c_sl];
#pragma unroll
                    for (int i = 0; i < ORD; ++i) st.hb[c_chan][c_iq][i] = wr[(len - 1 - i) * G::NCH];
                    st.hb_rejects[c_chan][c_iq] = rejects;
                }
            }
            __syncthreads();
        }
    } else if (role.kind == 1) {
        for (int64_t s = 0; s < n_steps; ++s) {
            const int64_t ts = s - 2;
            if (is_out && ts >= 0 && ts < n_tiles && !SP_DEBUG_SKIP_OUT) {
                const int wbi = (int)(ts % 3), yb = (int)(ts & 1);
                const int64_t i0 = ts * SP_T;
                const int len = (int)((n_frames - i0 < SP_T) ? n_frames - i0 : SP_T);
                const double *wr = &wb[wbi][SP_HIST][o_chain];
                double *yp = &ys[yb][0][o_chain];
                for (int t = o_phase * SP_K; t < len; t += SP_K * SP_PH) {
                    // samples t .. t+K-1: term i of sample t+k reads the state of row t+k-1-i = r[K - 1 - k + i]
                    constexpr int K = SP_K;
                    double r[ORD + K - 1];
#pragma unroll
                    for (int j = 0; j < ORD + K - 1; ++j) r[j] = wr[(t + K - 2 - j) * G::NCH];
                    Comp o[K];
                    double x[K];
#pragma unroll
                    for (int k = 0; k < K; ++k) { o[k].s = __dmul_rn(r[K - 1 - k], coef.ff[0]); o[k].c = 0.0; }
#pragma unroll
                    for (int k = 0; k < K; ++k) x[k] = __dmul_rn(__dmul_rn(r[K - 1 - k], coef.fb[0]), coef.d0);
                    comp_add_k<K>(o, x);
#pragma unroll
                    for (int i = 1; i < ORD; ++i) {
#pragma unroll
                        for (int k = 0; k < K; ++k) x[k] = __dmul_rn(r[K - 1 - k + i], coef.ff[i]);
                        comp_add_k<K>(o, x);
#pragma unroll
                        for (int k = 0; k < K; ++k) x[k] = __dmul_rn(__dmul_rn(r[K - 1 - k + i], coef.fb[i]), coef.d0);
                        comp_add_k<K>(o, x);
                    }
#pragma unroll
                    for (int k = 0; k < K; ++k)
                        if (t + k < len) yp[(t + k) * G::NCH] = o[k].s;     // no d0*x term: bug-for-bug (hblpf.c:1056)
                }
            }
            __syncthreads();
        }
    } else if (role.kind == 2) {
        for (int64_t s = 0; s < n_steps; ++s) {
            // ---- unpack tile s ---------------------------------------------------------------------------------------
            if (s < n_tiles && h_live && !SP_DEBUG_SKIP_UNPACK) {
                const int b = (int)(s & 1);
                const int64_t i0 = s * SP_T;
                const int len = (int)((n_frames - i0 < SP_T) ? n_frames - i0 : SP_T);
                for (int t = h_part; t < len; t += SP_HPS) {
                    double v[4];
                    unpack_frame(ch, h_src + (i0 + t) * ch.frame_bytes, h_pos0 + i0 + t, v);
                    xs[b][t][h_sl * 2] = v[0];
                    xs[b][t][h_sl * 2 + 1] = v[2];
                }
            }
            // ---- filter outputs of tile s-3 -> PCM -------------------------------------------------------------------
            const int64_t sp = s - 3;
            if (sp >= 0 && sp < n_tiles && h_live && !SP_DEBUG_SKIP_RENDER) {
                const int b = (int)(sp & 1);
                const int64_t i0 = sp * SP_T;
                const int len = (int)((n_frames - i0 < SP_T) ? n_frames - i0 : SP_T);
                for (int t = h_part; t < len; t += SP_HPS) {
                    const int64_t f = i0 + t;
                    const double *y = &ys[b][t][h_sl * 4];
                    double v[4];
                    int slot;
                    // up-mix: (yI, yQ) of each channel -> (re, im), reference lpf_hilbert_quad.c:132-153
                    {
                        const unsigned q = (h_q0[0] + (unsigned)f) & 3u;
                        double a = mix_up(0, q, y[0], slot); v[slot] = a;
                        double c = mix_up(1, q, y[1], slot); v[slot] = c;
                    }
                    {
                        const unsigned q = (h_q0[1] + (unsigned)f) & 3u;
                        double a = mix_up(0, q, y[2], slot); v[2 + slot] = a;
                        double c = mix_up(1, q, y[3], slot); v[2 + slot] = c;
                    }
                    if (fast) {
                        const uint4 nw = make_uint4(0u, 0u, 0u, 0u);
                        lean_frame_fast<ICW_RENDER_ROUND>(ch, streams[stream0 + h_sl], f, n_frames - 1, v, nw, nw, io.dst, acc, osc);
                    } else {
                        finish_frame<DITHER_LATE>(ch, streams[stream0 + h_sl], f, n_frames, v, bus, io, acc, osc);
                    }
                }
            }
            __syncthreads();
        }
        if (h_live) commit_acc_one(streams[stream0 + h_sl], acc);
    } else {
        for (int64_t s = 0; s < n_steps; ++s) __syncthreads();
    }
}

template <int ORD, int NS>
static cudaError_t launch_split_geom(const HbCoef &coef, const DevChain &ch, DevStream *streams, int n_streams, int64_t n_frames,
                                     const uint8_t *in, size_t in_stride, const uint32_t *mtw_l, const uint32_t *mtw_r,
                                     size_t mt_stream_stride, uint8_t *out, size_t out_stride, double *tap_bus, double *tap_lr,
                                     double *pre, int fast, cudaStream_t s)
{
    using G = SplitGeom<NS>;
    const int blocks = (n_streams + NS - 1) / NS;
    cudaError_t e1 = cudaFuncSetAttribute(hb_split_kernel<ORD, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM);
    if (e1 != cudaSuccess) return e1;
    hb_split_kernel<ORD, NS><<<blocks, G::THREADS, G::SMEM, s>>>(coef, ch, streams, n_streams, n_frames, in, in_stride, mtw_l, mtw_r,
                                                                 mt_stream_stride, out, out_stride, tap_bus, tap_lr, pre, fast);
    return cudaGetLastError();
}

// streams per CTA: the largest of 28 / 14 / 7 / 4 that still gives every SM a CTA (ICW_SPLIT_NS forces one)
static int split_streams_per_cta(int n_streams)
{
    static const int forced = [] { const char *v = getenv("ICW_SPLIT_NS"); return v ? atoi(v) : 0; }();
    if (forced == 28 || forced == 14 || forced == 7 || forced == 4) return forced;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    for (int ns : { 28, 14, 7 })
        if ((n_streams + ns - 1) / ns >= sms - 1) return ns;
    return 4;
}

template <int ORD>
static cudaError_t launch_split_ord(const HbCoef &coef, const DevChain &ch, DevStream *streams, int n_streams, int64_t n_frames,
                                    const uint8_t *in, size_t in_stride, const uint32_t *mtw_l, const uint32_t *mtw_r,
                                    size_t mt_stream_stride, uint8_t *out, size_t out_stride, double *tap_bus, double *tap_lr,
                                    double *pre, int fast, cudaStream_t s)
{
#define ICW_SPLIT_GEOM(N) \
    case N: return launch_split_geom<ORD, N>(coef, ch, streams, n_streams, n_frames, in, in_stride, mtw_l, mtw_r, mt_stream_stride, \
                                             out, out_stride, tap_bus, tap_lr, pre, fast, s)
    switch (split_streams_per_cta(n_streams)) {
        ICW_SPLIT_GEOM(28);
        ICW_SPLIT_GEOM(14);
        ICW_SPLIT_GEOM(7);
    default:
        ICW_SPLIT_GEOM(4);
    }
#undef ICW_SPLIT_GEOM
}

cudaError_t launch_hb_split(const HbCoef &coef, const DevChain &ch, DevStream *streams, int n_streams, int64_t n_frames,
                            const uint8_t *in, size_t in_stride, const uint32_t *mtw_l, const uint32_t *mtw_r,
                            size_t mt_stream_stride, uint8_t *out, size_t out_stride, double *tap_bus, double *tap_lr,
                            double *pre, int fast, cudaStream_t s)
{
#define ICW_SPLIT_CASE(O) \
    case O: return launch_split_ord<O>(coef, ch, streams, n_streams, n_frames, in, in_stride, mtw_l, mtw_r, mt_stream_stride, \
                                       out, out_stride, tap_bus, tap_lr, pre, fast, s)
    switch (ch.hb_ord) {
        ICW_SPLIT_CASE(15);
        ICW_SPLIT_CASE(18);
        ICW_SPLIT_CASE(19);
        ICW_SPLIT_CASE(20);
    default: return cudaErrorInvalidValue;
    }
#undef ICW_SPLIT_CASE
}

}  // namespace icw
