// icw_api.cu -- the C ABI declared in include/icw_b200.h: spec validation and folding, sessions,
// stream state, and the per-call launch sequence.  Host side only; kernels live in icw_kernels.cu.
#include <algorithm>
#include <cmath>
#include <climits>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "icw_internal.h"
#include "icw_kernels.h"
#include "icw_mt.h"
#include "icw_scan.h"
#include "icw_crc.h"
#include "icw_comm.h"
#include "icw_sfused.h"
#include "icw_hb_tables.inc"
#include "icw_ns_tables.inc"

using namespace icw;

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(ICW_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

extern "C" const char *icw_last_error(void) { return g_err.c_str(); }
extern "C" int icw_abi_version(void) { return ICW_ABI_VERSION; }

// ---------------------------------------------------------------------------------------------
// objects
// ---------------------------------------------------------------------------------------------
struct Scratch {
    void  *p = nullptr;
    size_t cap = 0;
    int reserve(size_t n)
    {
        if (n <= cap) return ICW_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 8 + 256;
        if (cudaMalloc(&p, want) != cudaSuccess) { cudaGetLastError(); return fail(ICW_E_NOMEM, "cudaMalloc(%zu) failed", want); }
        cap = want;
        return ICW_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct icw_engine {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t aux = nullptr;         // dither words are generated here, concurrently with the Hilbert kernels
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaStream_t h2d = nullptr, d2h = nullptr;      // copy streams of the host entry point's pipeline
    cudaEvent_t ev_in[2] = { nullptr, nullptr }, ev_comp[2] = { nullptr, nullptr }, ev_out[2] = { nullptr, nullptr };
    // engine-wide scratch (analytic, mtw, scan_scratch, ns_pre, checkpoints) is shared by all sessions: a call on a
    // stream other than the one that used it last first waits for that stream's work (event below)
    cudaStream_t last_stream = nullptr;
    bool last_stream_valid = false;
    cudaEvent_t ev_scratch = nullptr;
    int n_sessions = 0;
    bool dying = false;                 // icw_engine_destroy was called while sessions were alive
    Scratch analytic, mtw[2], ckpt, io_in, io_out, leaf;
    bool unfused = false;               // ICW_UNFUSED=1: keep the two-kernel exact path (A/B measurements)
    bool no_fuse_mt = false;            // ICW_NO_FUSE_MT=1: dither words through HBM even where chain_mt_kernel applies (A/B)
    Scratch scan_scratch;
    Scratch crc_partial;                // per-tile CRC registers + the result word
    Scratch ns_pre;                     // noise shaping: (value, dither) pairs between chain_kernel and ns_render_kernel
    struct ScanPlan { bool ready = false; ModalCoef mc; double *d_pw = nullptr; };
    struct SfPlan { bool ready = false; SfTab tb; int64_t warm = 0; };
    SfPlan sf[ICW_HB_NTYPES][2];        // [filter_no][baseline]: the one-kernel scan path's chunk tables (icw_sfused.h)
    ScanPlan scan[ICW_HB_NTYPES][2][4]; // [filter_no][baseline][chunk length: 256, 1024, 2048 (or the forced one); 3 = the one-kernel path's]
    int sfused = -1;                    // ICW_SFUSED: 1 = the one-kernel scan path wherever it applies, 0 = never, unset = where it pays
    MtJump mt;                          // MT19937 checkpoint service (icw_mt.cu)
};

struct icw_session {
    icw_engine *e = nullptr;
    icw_chain_spec spec;
    DevChain ch;
    HbCoef coef;
    int n_streams = 0;
    DevStream *d_streams = nullptr;
    // host mirror of the generator identity per stream (closed form: never needs a device read)
    std::vector<uint32_t> mt_seed[2];
    std::vector<uint64_t> mt_drawn[2];
    uint64_t launches = 0;
    unsigned long long redraws_seen = 0;    // sum of the streams' mt_redraws already reported (ICW_E_MT_REDRAW)
    uint64_t redraws_handled = 0;           // dither draws that met the rejection loop and were replayed the reference's way
    DevStream *d_snap = nullptr;            // the streams as they stood when the current host call began (replay)
    struct MtPatch { int chan; uint64_t idx; uint32_t value; };
    std::vector<MtPatch> mt_patches;        // test hook: icw_debug_patch_mt_word
    cudaStream_t last_stream = nullptr; // the stream of the most recent process call (NULL: none yet / engine stream)
    int hb_basis = 0;                   // basis of the Hilbert state on the device
    bool hb_live = false;               // that state is not all-zero / has been used
    double *d_tap_bus = nullptr;        // test taps, set through icw_session_set_taps
    double *d_tap_lr = nullptr;
    // measurement: event pairs around each kernel class (icw_session_profile)
    bool profiling = false;
    std::vector<cudaEvent_t> ev_pool;
    struct Span { int k; cudaEvent_t a, b; };
    std::vector<Span> spans;
    double prof_ms[ICW_K_COUNT] = {};
    uint64_t prof_n[ICW_K_COUNT] = {};
};

// RAII: records an event pair around a group of launches of one kernel class
struct ProfSpan {
    icw_session *s; cudaStream_t st; int k; cudaEvent_t a = nullptr, b = nullptr;
    static cudaEvent_t take(icw_session *s)
    {
        cudaEvent_t e = nullptr;
        if (!s->ev_pool.empty()) { e = s->ev_pool.back(); s->ev_pool.pop_back(); }
        else cudaEventCreate(&e);
        return e;
    }
    ProfSpan(icw_session *s_, cudaStream_t st_, int k_) : s(s_), st(st_), k(k_)
    {
        if (!s->profiling) return;
        a = take(s); b = take(s);
        cudaEventRecord(a, st);
    }
    ~ProfSpan()
    {
        if (!a) return;
        cudaEventRecord(b, st);
        s->spans.push_back({ k, a, b });
    }
};

// ---------------------------------------------------------------------------------------------
// defaults and small queries
// ---------------------------------------------------------------------------------------------
static int chan_bytes_of(int fmt)
{
    switch (fmt) {
    case ICW_FMT_WAV_U8: return 1;
    case ICW_FMT_WAV_I16: return 2;
    case ICW_FMT_WAV_I24: return 3;
    case ICW_FMT_WAV_I32: return 4;
    case ICW_FMT_WAV_F32: return 4;
    case ICW_FMT_CW_F64: return 16;
    case ICW_FMT_CW_I16: return 4;
    case ICW_FMT_CW_I16F32: return 6;
    case ICW_FMT_CW_F32: return 8;
    default: return -1;
    }
}

extern "C" void icw_default_spec(icw_chain_spec *sp)
{
    memset(sp, 0, sizeof *sp);
    sp->fmt = ICW_FMT_WAV_F32;
    sp->n_channels = 2;
    sp->sample_rate = 48000;
    sp->filter_no = 1;              // IIR_HBLPF_IX   (reference src/config.c:158)
    sp->is_kahan = 1;               // IIR_SUM_KAHAN  (:161)
    sp->is_subnorm_reject = 1;      // IIR_SUBN_ZERO  (:164)
    sp->hilbert_mode = ICW_HILBERT_EXACT;
    sp->is_frmod_scaled = 1;        // FRMOD_SCALED   (:154)
    sp->need24bits = 1;             // NEED24BITS     (:183)
    sp->dth_bits = 1.0;
    sp->quantz_type = 1;            // mid riser      (:190)
    sp->render_type = ICW_RENDER_ROUND;
    sp->nshape_type = 0;
    sp->sign_bits16 = 16;
    sp->sign_bits24 = 24;
    sp->n_nodes = 1;                // the lone master (reference src/adv_modulator.c:112-118)
    sp->nodes[0].mode = ICW_MODE_MASTER;
    sp->nodes[0].inputs_mask = 1u;
    sp->nodes[0].l_gain = sp->nodes[0].r_gain = 0.8;
    sp->nodes[0].l_tout = sp->nodes[0].r_tout = ICW_OUT_ADD_REIM;
}

extern "C" void icw_default_state(icw_stream_state *st)
{
    memset(st, 0, sizeof *st);
    st->mt_seed[0] = 0x13579BDFu;   // reference src/in_cwave.c:69
    st->mt_seed[1] = 0x479B22ABu;   // reference src/in_cwave.c:70
}

extern "C" int icw_frame_bytes(const icw_chain_spec *sp)
{
    int cb = chan_bytes_of(sp->fmt);
    return (cb < 0 || sp->n_channels < 1 || sp->n_channels > 2) ? -1 : cb * sp->n_channels;
}
extern "C" int icw_out_frame_bytes(const icw_chain_spec *sp) { return sp->need24bits ? 6 : 4; }
extern "C" double icw_peak_db(double lin) { return lin ? 20.0 * log10(lin) : ICW_SILENCE_DB; }

// ---------------------------------------------------------------------------------------------
// spec -> device form
// ---------------------------------------------------------------------------------------------
static double word_as_double(unsigned long long w) { double d; memcpy(&d, &w, 8); return d; }

static double scaled_freq(double f, int scaled)
{
    // reference src/adv_modulator.c:36-38
    return scaled ? (double)(unsigned)(f * (double)ICW_HZ_SCALE + 0.5) : f;
}

static int fold_spec(const icw_chain_spec &sp, DevChain &ch, HbCoef &coef)
{
    memset(&ch, 0, sizeof ch);
    memset(&coef, 0, sizeof coef);
    int cb = chan_bytes_of(sp.fmt);
    if (cb < 0) return fail(ICW_E_ARG, "unknown sample format %d", sp.fmt);
    if (sp.n_channels < 1 || sp.n_channels > 2) return fail(ICW_E_ARG, "n_channels must be 1 or 2");
    if (sp.sample_rate == 0 || sp.sample_rate > 2000000u) return fail(ICW_E_ARG, "sample_rate out of range");
    if (sp.filter_no < 0 || sp.filter_no >= ICW_HB_NTYPES) return fail(ICW_E_ARG, "filter_no must be 0..5");
    if (sp.n_nodes < 1 || sp.n_nodes > ICW_MAX_NODES) return fail(ICW_E_ARG, "n_nodes out of range");
    if (sp.nodes[sp.n_nodes - 1].mode != ICW_MODE_MASTER) return fail(ICW_E_ARG, "the last node must be the master");
    if (sp.render_type > ICW_RENDER_GAUSS) return fail(ICW_E_ARG, "render_type out of range");
    if (sp.quantz_type > 1) return fail(ICW_E_ARG, "quantz_type out of range");
    if (sp.nshape_type >= ICW_NS_COUNT) return fail(ICW_E_ARG, "nshape_type out of range");
    if (sp.sign_bits16 < 2 || sp.sign_bits16 > 16 || sp.sign_bits24 < 2 || sp.sign_bits24 > 24)
        return fail(ICW_E_ARG, "significant bits out of range");
    if (sp.n_fade_in < 0 || sp.n_fade_out < 0) return fail(ICW_E_ARG, "negative fade length");

    ch.fmt = sp.fmt;
    ch.n_channels = sp.n_channels;
    ch.chan_bytes = cb;
    ch.frame_bytes = cb * sp.n_channels;
    ch.out_frame_bytes = sp.need24bits ? 6 : 4;
    ch.is_complex = sp.fmt >= ICW_FMT_CW_F64;
    ch.is_frmod_scaled = sp.is_frmod_scaled != 0;
    ch.bypass = sp.bypass != 0;
    ch.n_nodes = sp.n_nodes;
    ch.n_samples = sp.n_samples;
    ch.n_fade_in = sp.n_fade_in;
    ch.n_fade_out = sp.n_fade_out;
    if (ch.is_frmod_scaled) {
        // reference src/adv_modulator.c:614: unsigned scale_sr = sample_rate * HZ_SCALE
        unsigned scale = sp.sample_rate * ICW_HZ_SCALE;
        ch.scale_sr = scale;
        ch.osc_div = (double)scale;
    } else {
        ch.scale_sr = 0;
        ch.osc_div = (double)sp.sample_rate;
    }
    ch.osc_rdiv = 1.0 / ch.osc_div;

    // half-band design (reference src/hblpf.c:849-856)
    ch.filter_no = sp.filter_no;
    ch.hb_ord = ICW_HB_ORDER[sp.filter_no];
    ch.is_kahan = sp.is_kahan != 0;
    ch.reject_flag = sp.is_subnorm_reject;
    ch.fp_check = sp.is_fp_check != 0;
    {
        double a0 = word_as_double(ICW_HB_A[sp.filter_no][0]);
        coef.d0 = word_as_double(ICW_HB_B[sp.filter_no][0]) / a0;
        for (int i = 0; i < ch.hb_ord; ++i) {
            coef.fb[i] = -word_as_double(ICW_HB_A[sp.filter_no][i + 1]) / a0;
            coef.ff[i] = word_as_double(ICW_HB_B[sp.filter_no][i + 1]) / a0;
        }
        memcpy(ch.hb_fb, coef.fb, sizeof coef.fb);
        memcpy(ch.hb_ff, coef.ff, sizeof coef.ff);
        ch.hb_d0 = coef.d0;
    }

    // quantiser (reference src/sound_render.c:499-551)
    DevRender &q = ch.render;
    q.dth_mul = pow(2.0, sp.dth_bits) - 1.0;
    if (sp.quantz_type == 0) { q.round_off = 0.5; q.neg_delta = 0; }
    else                     { q.round_off = 0.0; q.neg_delta = -1; }
    long long top;
    if (sp.need24bits) {
        q.bytes = 3;
        q.shift = 24 - (int)sp.sign_bits24;
        top = 0x800000LL >> q.shift;
        q.norm_mul = q.shift < 8 ? (double)(0x100 >> q.shift) : 1.0 / (double)(1ULL << (q.shift - 8));
    } else {
        q.bytes = 2;
        q.shift = 16 - (int)sp.sign_bits16;
        top = 0x8000LL >> q.shift;
        q.norm_mul = 1.0 / (double)(1ULL << q.shift);
    }
    q.hi = (double)top;
    q.lo = -(double)(top + 1 + q.neg_delta);
    q.lo -= (double)q.neg_delta;
    q.inv_hi = 1.0 / q.hi;
    q.render_type = (int)sp.render_type;
    static const int wps[5] = { 0, 2, 4, 2, 24 };       // reference src/sound_render.c:711-751
    q.words_per_sample = wps[sp.render_type];
    q.ns_kind = ICW_NS_KIND[sp.nshape_type];
    q.ns_order = ICW_NS_ORDER[sp.nshape_type];
    for (int i = 0; i < 2 * ICW_NS_MAX_TAPS; ++i)
        q.ns_coef[i] = i < ICW_NS_MAX_TAPS ? word_as_double(ICW_NS_COEF[sp.nshape_type][i]) : 0.0;
    if (q.ns_kind == 2)     // IIR: [order, 2*order) of the table weigh the filter's own outputs
        for (int i = 0; i < q.ns_order; ++i) { q.ns_coef[ICW_NS_MAX_TAPS + i] = q.ns_coef[q.ns_order + i]; }

    // DSP list.  First pass: every output plug in range (the feedback analysis below shifts by it)
    for (int i = 0; i < sp.n_nodes; ++i) {
        const icw_node &s = sp.nodes[i];
        if (s.mode < ICW_MODE_MASTER || s.mode > ICW_MODE_MIX) return fail(ICW_E_ARG, "node %d: bad mode %d", i, s.mode);
        if (s.mode != ICW_MODE_MASTER && (s.n_out < 1 || s.n_out >= ICW_N_PLUGS))
            return fail(ICW_E_ARG, "node %d: output plug %d out of range", i, s.n_out);
    }
    uint32_t written = 1u;                                // plug 0 is written by the unpacker
    for (int i = 0; i < sp.n_nodes; ++i) {
        const icw_node &s = sp.nodes[i];
        DevNode &d = ch.nodes[i];
        if (s.mode < ICW_MODE_MASTER || s.mode > ICW_MODE_MIX) return fail(ICW_E_ARG, "node %d: bad mode %d", i, s.mode);
        if (s.mode == ICW_MODE_MASTER && i != sp.n_nodes - 1) return fail(ICW_E_ARG, "node %d: a second master", i);
        if (s.inputs_mask >> ICW_N_PLUGS) return fail(ICW_E_ARG, "node %d: input plug out of range", i);
        if (s.mode != ICW_MODE_MASTER && (s.n_out < 1 || s.n_out >= ICW_N_PLUGS))
            return fail(ICW_E_ARG, "node %d: output plug %d out of range", i, s.n_out);
        if (!sp.bypass) {
            // a plug read before its writer in the same frame carries the PREVIOUS frame's value
            // (reference src/adv_modulator.c:634-751): a one-frame feedback loop, serial in time
            uint32_t later = 0;
            for (int j = i; j < sp.n_nodes; ++j)
                if (sp.nodes[j].mode != ICW_MODE_MASTER) later |= 1u << sp.nodes[j].n_out;
            if (s.inputs_mask & later & ~written) ch.feedback = 1;  // run frame by frame, one thread a stream (chain_serial_kernel)
        }
        d.mode = s.mode;
        d.inputs_mask = s.inputs_mask;
        d.xch_mode = s.xch_mode;
        d.l_iq_invert = s.l_iq_invert != 0;
        d.r_iq_invert = s.r_iq_invert != 0;
        d.n_out = s.n_out;
        d.l_tout = s.l_tout; d.r_tout = s.r_tout;
        d.l_on = s.l_on != 0; d.r_on = s.r_on != 0;
        d.l_gain = s.l_gain; d.r_gain = s.r_gain;
        if (s.mode == ICW_MODE_SHIFT) {
            d.l_neg = s.l_p[0] < 0.0; d.r_neg = s.r_p[0] < 0.0;
            d.l_f = scaled_freq(d.l_neg ? -s.l_p[0] : s.l_p[0], ch.is_frmod_scaled);
            d.r_f = scaled_freq(d.r_neg ? -s.r_p[0] : s.r_p[0], ch.is_frmod_scaled);
        } else if (s.mode == ICW_MODE_PM) {
            d.l_f = scaled_freq(s.l_p[0], ch.is_frmod_scaled);
            d.r_f = scaled_freq(s.r_p[0], ch.is_frmod_scaled);
            d.l_ph0 = s.l_p[1] * ICW_PI;    d.r_ph0 = s.r_p[1] * ICW_PI;
            d.l_lvlpi = s.l_p[2] * ICW_PI;  d.r_lvlpi = s.r_p[2] * ICW_PI;
            d.l_angle = s.l_p[3];           d.r_angle = s.r_p[3];
        }
        if (s.mode != ICW_MODE_MASTER) written |= 1u << s.n_out;
        d.n_in = 0;
        memset(d.pad_, 0, sizeof d.pad_);
        memset(d.in_off, 0, sizeof d.in_off);
        {
            int idx[3] = { 0, 0, 0 }, n_in = 0;
            for (int k = 0; k < ICW_N_PLUGS; ++k)
                if (s.inputs_mask >> k & 1u) { if (n_in < 3) idx[n_in] = k; ++n_in; }
            int m[4] = { 0, 1, 2, 3 };
            bool ok = n_in >= 1 && n_in <= 3;
            switch (s.xch_mode) {                               // reference src/adv_modulator.c:668-722: XCH, then I/Q inversion
            case ICW_XCH_NORMAL: break;
            case ICW_XCH_SWAP:      m[0] = 2; m[1] = 3; m[2] = 0; m[3] = 1; break;
            case ICW_XCH_LEFTONLY:  m[2] = 0; m[3] = 1; break;
            case ICW_XCH_RIGHTONLY: m[0] = 2; m[1] = 3; break;
            default: ok = false; break;                         // (L+R)/2 is arithmetic, not a move
            }
            if (d.l_iq_invert) { const int t = m[0]; m[0] = m[1]; m[1] = t; }
            if (d.r_iq_invert) { const int t = m[2]; m[2] = m[3]; m[3] = t; }
            if (ok) {
                d.n_in = n_in;
                for (int i = 0; i < n_in; ++i) {
                    auto off = [&](int j) { return (idx[i] * 4 + m[j]) * (int)sizeof(double); };
                    d.in_off[i] = make_int4(off(0), off(1), off(2), off(3));
                }
            }
        }
    }
    if (!ch.is_complex && sp.hilbert_mode != ICW_HILBERT_EXACT && sp.hilbert_mode != ICW_HILBERT_SCAN)
        return fail(ICW_E_ARG, "hilbert_mode out of range");
    // common list shapes get straight-line device code (same operations, no interpreter overhead)
    ch.shape = ICW_SHAPE_GENERIC;
    {
        auto plain = [](const DevNode &d) { return d.xch_mode == ICW_XCH_NORMAL && !d.l_iq_invert && !d.r_iq_invert; };
        const DevNode &last = ch.nodes[ch.n_nodes - 1];
        if (!ch.bypass && !ch.feedback && plain(last)) {
            if (ch.n_nodes == 1 && last.inputs_mask == 1u) ch.shape = ICW_SHAPE_MASTER;
            if (ch.n_nodes == 2 && ch.nodes[0].mode == ICW_MODE_SHIFT && plain(ch.nodes[0]) && ch.nodes[0].inputs_mask == 1u &&
                last.inputs_mask == (1u << ch.nodes[0].n_out))
                ch.shape = ICW_SHAPE_SHIFT_MASTER;
        }
    }
    return ICW_OK;
}

// ---------------------------------------------------------------------------------------------
// engine
// ---------------------------------------------------------------------------------------------
extern "C" int icw_engine_create(int device, icw_engine **out)
{
    if (!out) return fail(ICW_E_ARG, "out is NULL");
    *out = nullptr;
    int n = 0;
    CK(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) return fail(ICW_E_ARG, "device %d of %d", device, n);
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(ICW_E_CUDA, "this library is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
    icw_engine *e = new icw_engine;
    e->device = device;
    e->sm_count = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&e->aux, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&e->ev_scratch, cudaEventDisableTiming));
    { const char *u = getenv("ICW_UNFUSED"); e->unfused = u && *u == '1'; }
    { const char *u = getenv("ICW_NO_FUSE_MT"); e->no_fuse_mt = u && *u == '1'; }
    { const char *u = getenv("ICW_SFUSED"); e->sfused = (u && *u) ? atoi(u) : -1; }
    *out = e;
    return ICW_OK;
}

static void engine_free(icw_engine *e);

extern "C" void icw_engine_destroy(icw_engine *e)
{
    if (!e) return;
    if (e->n_sessions > 0) { e->dying = true; return; }     // the last session to go frees the engine
    engine_free(e);
}

static void engine_free(icw_engine *e)
{
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
    e->analytic.release(); e->mtw[0].release(); e->mtw[1].release(); e->ckpt.release();
    e->io_in.release(); e->io_out.release(); e->leaf.release(); e->scan_scratch.release(); e->ns_pre.release(); e->crc_partial.release();
    for (auto &row : e->scan) for (auto &col : row) for (auto &pl : col) if (pl.d_pw) cudaFree(pl.d_pw);
    e->mt.release();
    cudaStreamSynchronize(e->aux);
    if (e->h2d) {
        cudaStreamSynchronize(e->h2d); cudaStreamSynchronize(e->d2h);
        for (int i = 0; i < 2; ++i) { cudaEventDestroy(e->ev_in[i]); cudaEventDestroy(e->ev_comp[i]); cudaEventDestroy(e->ev_out[i]); }
        cudaStreamDestroy(e->h2d); cudaStreamDestroy(e->d2h);
    }
    cudaEventDestroy(e->ev_fork); cudaEventDestroy(e->ev_join); cudaEventDestroy(e->ev_scratch);
    cudaStreamDestroy(e->aux);
    cudaStreamDestroy(e->stream);
    delete e;
}

// ---------------------------------------------------------------------------------------------
// sessions and state
// ---------------------------------------------------------------------------------------------
static void to_dev(const icw_stream_state &s, DevStream &d)
{
    memset(&d, 0, sizeof d);
    d.n_frame = s.n_frame;
    d.pos = s.pos;
    memcpy(d.hb, s.hb, sizeof d.hb);
    for (int c = 0; c < 2; ++c) {
        for (int f = 0; f < 2; ++f) d.hb_rejects[c][f] = s.hb_rejects[c][f];
        d.quad[c] = s.quad[c] & 3u;
        d.mt_seed[c] = s.mt_seed[c];
        d.mt_drawn[c] = s.mt_drawn[c];
        d.prev_rnd[c] = s.prev_rnd[c];
        d.clips[c] = s.clips[c];
        d.peak[c] = s.peak[c];
    }
    memcpy(d.bus, s.bus, sizeof d.bus);
    d.hb_basis = s.hb_basis;
    memcpy(d.ns_e, s.ns_e, sizeof d.ns_e);
    memcpy(d.ns_o, s.ns_o, sizeof d.ns_o);
    memcpy(d.ns_prev_err, s.ns_prev_err, sizeof d.ns_prev_err);
}

static void from_dev(const DevStream &d, icw_stream_state &s)
{
    memset(&s, 0, sizeof s);
    s.n_frame = d.n_frame;
    s.pos = d.pos;
    memcpy(s.hb, d.hb, sizeof s.hb);
    for (int c = 0; c < 2; ++c) {
        for (int f = 0; f < 2; ++f) s.hb_rejects[c][f] = d.hb_rejects[c][f];
        s.quad[c] = d.quad[c];
        s.mt_seed[c] = d.mt_seed[c];
        s.mt_drawn[c] = d.mt_drawn[c];
        s.prev_rnd[c] = d.prev_rnd[c];
        s.clips[c] = d.clips[c];
        s.peak[c] = d.peak[c];
    }
    memcpy(s.bus, d.bus, sizeof s.bus);
    s.hb_basis = d.hb_basis;
    memcpy(s.ns_e, d.ns_e, sizeof s.ns_e);
    memcpy(s.ns_o, d.ns_o, sizeof s.ns_o);
    memcpy(s.ns_prev_err, d.ns_prev_err, sizeof s.ns_prev_err);
}

extern "C" int icw_session_create(icw_engine *e, const icw_chain_spec *spec, int n_streams, icw_session **out)
{
    if (!e || !spec || !out) return fail(ICW_E_ARG, "NULL argument");
    *out = nullptr;
    if (n_streams < 1) return fail(ICW_E_ARG, "n_streams must be >= 1");
    icw_session *s = new icw_session;
    s->e = e;
    s->spec = *spec;
    int rc = fold_spec(*spec, s->ch, s->coef);
    if (rc) { delete s; return rc; }
    s->n_streams = n_streams;
    CK(cudaSetDevice(e->device));
    if (cudaMalloc(&s->d_streams, sizeof(DevStream) * (size_t)n_streams) != cudaSuccess) {
        cudaGetLastError();
        delete s;
        return fail(ICW_E_NOMEM, "cudaMalloc of %d stream states failed", n_streams);
    }
    icw_stream_state fresh;
    icw_default_state(&fresh);
    DevStream d;
    to_dev(fresh, d);
    std::vector<DevStream> all((size_t)n_streams, d);
    CK(cudaMemcpyAsync(s->d_streams, all.data(), sizeof(DevStream) * all.size(), cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    for (int c = 0; c < 2; ++c) {
        s->mt_seed[c].assign((size_t)n_streams, fresh.mt_seed[c]);
        s->mt_drawn[c].assign((size_t)n_streams, 0);
    }
    e->n_sessions++;
    *out = s;
    return ICW_OK;
}

static int quiesce(icw_session *s);

extern "C" void icw_session_destroy(icw_session *s)
{
    if (!s) return;
    quiesce(s);
    if (s->d_streams) cudaFree(s->d_streams);
    if (s->d_snap) cudaFree(s->d_snap);
    for (auto &sp : s->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    for (auto ev : s->ev_pool) cudaEventDestroy(ev);
    icw_engine *e = s->e;
    delete s;
    if (--e->n_sessions == 0 && e->dying) engine_free(e);
}

// Everything that may still read or write this session's stream states: the engine's own stream and the caller's
// stream of the last icw_session_process_device (which may be a non-blocking stream that nothing else orders).
static int quiesce(icw_session *s)
{
    CK(cudaSetDevice(s->e->device));
    if (s->last_stream && s->last_stream != s->e->stream) CK(cudaStreamSynchronize(s->last_stream));
    CK(cudaStreamSynchronize(s->e->stream));
    return ICW_OK;
}

// The streams are idle (quiesce / a synchronised host call): did any dither draw run into the reference's rejection loop
// since the last look?  The draw index is a closed form of the frame index here, so the draws after such an event are two
// words early -- not the reference's output, and said so rather than passed on.
static int check_redraws(icw_session *s)
{
    if (!s->ch.render.words_per_sample) return ICW_OK;
    std::vector<unsigned long long> cnt((size_t)s->n_streams);
    CK(cudaMemcpy2D(cnt.data(), sizeof(unsigned long long), &s->d_streams[0].mt_redraws, sizeof(DevStream), sizeof(unsigned long long),
                    (size_t)s->n_streams, cudaMemcpyDeviceToHost));
    unsigned long long tot = 0;
    for (auto c : cnt) tot += c;
    if (tot < s->redraws_seen) s->redraws_seen = tot;           // the counters were reset
    if (tot != s->redraws_seen) {
        const unsigned long long d = tot - s->redraws_seen;
        s->redraws_seen = tot;
        return fail(ICW_E_MT_REDRAW, "%llu dither draw(s) hit the generator's rejection loop (mt_jrnd.c:249-253): the output after the first "
                                     "of them is not the reference's; replay the stream from a state before the event in smaller calls", d);
    }
    return ICW_OK;
}

extern "C" int icw_session_sync(icw_session *s)
{
    if (!s) return fail(ICW_E_ARG, "NULL session");
    int rc = quiesce(s);
    return rc ? rc : check_redraws(s);
}

extern "C" int icw_session_get_state(icw_session *s, int k, icw_stream_state *out)
{
    if (!s || !out || k < 0 || k >= s->n_streams) return fail(ICW_E_ARG, "bad stream index");
    DevStream d;
    int rc = quiesce(s);
    if (rc) return rc;
    CK(cudaMemcpy(&d, s->d_streams + k, sizeof d, cudaMemcpyDeviceToHost));
    from_dev(d, *out);
    out->hb_basis = (uint32_t)s->hb_basis;
    return ICW_OK;
}

extern "C" int icw_session_set_state(icw_session *s, int k, const icw_stream_state *in)
{
    if (!s || !in || k < 0 || k >= s->n_streams) return fail(ICW_E_ARG, "bad stream index");
    // validate before anything is touched: a refused call leaves the stream as it was
    bool any = false;
    for (int c = 0; c < 2 && !any; ++c) for (int f = 0; f < 2 && !any; ++f) for (int i = 0; i < ICW_MAX_ORD; ++i) if (in->hb[c][f][i] != 0.0) { any = true; break; }
    if (any) {
        if (in->hb_basis > 1u) return fail(ICW_E_ARG, "stream %d: hb_basis %u is neither 0 (delay line) nor 1 (modal)", k, in->hb_basis);
        if (s->hb_live && s->hb_basis != (int)in->hb_basis && s->n_streams > 1)
            return fail(ICW_E_ARG, "stream %d: hb_basis %u differs from the session's live Hilbert state basis %d", k, in->hb_basis, s->hb_basis);
    }
    int rc = quiesce(s);
    if (rc) return rc;
    // what only the kernels keep (re-draw flags, FP_CHECK counters) is not part of the ABI state and stays
    DevStream d, old;
    CK(cudaMemcpy(&old, s->d_streams + k, sizeof old, cudaMemcpyDeviceToHost));
    to_dev(*in, d);
    d.mt_redraws = old.mt_redraws;
    memcpy(d.fp_cnt, old.fp_cnt, sizeof d.fp_cnt);
    CK(cudaMemcpy(s->d_streams + k, &d, sizeof d, cudaMemcpyHostToDevice));
    for (int c = 0; c < 2; ++c) { s->mt_seed[c][k] = in->mt_seed[c]; s->mt_drawn[c][k] = in->mt_drawn[c]; }
    if (any) {
        s->hb_basis = (int)in->hb_basis;
        s->hb_live = true;
    }
    return ICW_OK;
}

static int rewrite_states(icw_session *s, void (*fn)(DevStream &, void *), void *arg)
{
    std::vector<DevStream> all((size_t)s->n_streams);
    int rc = quiesce(s);
    if (rc) return rc;
    CK(cudaMemcpy(all.data(), s->d_streams, sizeof(DevStream) * all.size(), cudaMemcpyDeviceToHost));
    for (auto &d : all) fn(d, arg);
    CK(cudaMemcpy(s->d_streams, all.data(), sizeof(DevStream) * all.size(), cudaMemcpyHostToDevice));
    return ICW_OK;
}

extern "C" int icw_session_reset(icw_session *s, unsigned what)
{
    if (!s) return fail(ICW_E_ARG, "NULL session");
    if (what & ICW_RESET_HILBERT) s->hb_live = false;
    if (what & ICW_RESET_RENDER) {
        for (int c = 0; c < 2; ++c) std::fill(s->mt_drawn[c].begin(), s->mt_drawn[c].end(), (uint64_t)0);
        s->redraws_seen = 0;
    }
    return rewrite_states(s, [](DevStream &d, void *a) {
        unsigned w = *(unsigned *)a;
        if (w & ICW_RESET_HILBERT) {            // hq_rp_reset, reference src/lpf_hilbert_quad.c:160-165
            memset(d.hb, 0, sizeof d.hb);
            memset(d.hb_rejects, 0, sizeof d.hb_rejects);
            d.quad[0] = d.quad[1] = 0;
        }
        if (w & ICW_RESET_FRAMECNT) d.n_frame = 0;
        if (w & ICW_RESET_COUNTERS) {
            d.clips[0] = d.clips[1] = 0; d.peak[0] = d.peak[1] = 0.0;
            memset(d.fp_cnt, 0, sizeof d.fp_cnt);               // except_stats_reset, reference src/fp_check.c:37-47
        }
        if (w & ICW_RESET_FILEPOS) d.pos = 0;
        if (w & ICW_RESET_RENDER_MEMORY) {      // sound_render_recalc (src/sound_render.c:509,556-580)
            d.prev_rnd[0] = d.prev_rnd[1] = 0.0; d.prev_rnd_next[0] = d.prev_rnd_next[1] = 0.0;
            memset(d.ns_e, 0, sizeof d.ns_e); memset(d.ns_o, 0, sizeof d.ns_o);
            d.ns_prev_err[0] = d.ns_prev_err[1] = 0.0;
        }
        if (w & ICW_RESET_RENDER) {             // what a fresh context holds (winampGetInModule2, src/in_cwave.c:46-80,551-572)
            d.mt_drawn[0] = d.mt_drawn[1] = 0;  // mtrnd_init_seed: the generators back at their seeds' first draw
            d.prev_rnd[0] = d.prev_rnd[1] = 0.0; d.prev_rnd_next[0] = d.prev_rnd_next[1] = 0.0;
            memset(d.ns_e, 0, sizeof d.ns_e); memset(d.ns_o, 0, sizeof d.ns_o);
            d.ns_prev_err[0] = d.ns_prev_err[1] = 0.0;
            memset(d.bus, 0, sizeof d.bus);
            d.mt_redraws = 0;
        }
    }, &what);
}

extern "C" int icw_session_set_spec(icw_session *s, const icw_chain_spec *spec)
{
    if (!s || !spec) return fail(ICW_E_ARG, "NULL argument");
    DevChain ch;
    HbCoef coef;
    int rc = fold_spec(*spec, ch, coef);
    if (rc) return rc;
    unsigned flags = 0;
    // a different half-band design replaces the converters (reference src/in_cwave.c:135-150,177-191)
    if (spec->filter_no != s->spec.filter_no) { flags |= 1; s->hb_live = false; }
    // any sound_render_setup / set_outbits clears prev_rnd (reference src/sound_render.c:509)
    if (memcmp(&ch.render, &s->ch.render, sizeof ch.render) != 0) flags |= 2;
    // iir_rp_setcfg clears the reject counters (reference src/hblpf.c:1114-1125)
    if (spec->is_kahan != s->spec.is_kahan || spec->is_subnorm_reject != s->spec.is_subnorm_reject) flags |= 4;
    // switching FP_CHECK on or off clears its counters (reference mod_context_fpcheck_endis, src/in_cwave.c:325-379)
    if ((spec->is_fp_check != 0) != (s->spec.is_fp_check != 0)) flags |= 8;
    if (flags) {
        rc = rewrite_states(s, [](DevStream &d, void *a) {
            unsigned f = *(unsigned *)a;
            if (f & 1) { memset(d.hb, 0, sizeof d.hb); d.quad[0] = d.quad[1] = 0; }
            if (f & (1 | 4)) memset(d.hb_rejects, 0, sizeof d.hb_rejects);
            if (f & 8) memset(d.fp_cnt, 0, sizeof d.fp_cnt);
            if (f & 2) {            // sound_render_recalc also clears the shaper memory (src/sound_render.c:546-571)
                d.prev_rnd[0] = d.prev_rnd[1] = 0.0;
                memset(d.ns_e, 0, sizeof d.ns_e); memset(d.ns_o, 0, sizeof d.ns_o);
                d.ns_prev_err[0] = d.ns_prev_err[1] = 0.0;
            }
        }, &flags);
        if (rc) return rc;
    }
    s->spec = *spec;
    s->ch = ch;
    s->coef = coef;
    return ICW_OK;
}

// page-locked host memory for the layers above this ABI (they are plain C and do not link the CUDA runtime)
extern "C" int icw_pinned_alloc(size_t bytes, void **out)
{
    if (!out) return fail(ICW_E_ARG, "out is NULL");
    *out = nullptr;
    if (!bytes) return ICW_OK;
    if (cudaHostAlloc(out, bytes, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        *out = nullptr;
        return fail(ICW_E_NOMEM, "cudaHostAlloc(%zu) failed", bytes);
    }
    return ICW_OK;
}
extern "C" void icw_pinned_free(void *p) { if (p) cudaFreeHost(p); }

// CRC-32 of a device buffer with the reference's conventions (src/crc32.c:55-108: crc32init / update / final)
extern "C" int icw_crc32_device(icw_engine *e, const void *d_data, size_t n_bytes, uint32_t *crc_out)
{
    if (!e || !crc_out || (!d_data && n_bytes)) return fail(ICW_E_ARG, "NULL argument");
    CK(cudaSetDevice(e->device));
    const size_t words = n_bytes / 32768 + 4;
    int rc = e->crc_partial.reserve(words * sizeof(uint32_t));
    if (rc) return rc;
    uint32_t *d_part = (uint32_t *)e->crc_partial.p, *d_res = d_part + (words - 1);
    int nl = 0;
    CK(launch_crc32_raw((const uint8_t *)d_data, n_bytes, d_part, d_res, e->stream, &nl));
    uint32_t raw = 0;
    CK(cudaMemcpyAsync(&raw, d_res, sizeof raw, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    // start value 0xFFFFFFFF adds its own image after n bytes; then the final inversion
    *crc_out = ~(raw ^ crc32_shift(0xFFFFFFFFu, n_bytes));
    return ICW_OK;
}

// the same over host memory: blocks go through the engine's input staging and are joined with crc32_combine
extern "C" int icw_crc32_host(icw_engine *e, const void *data, size_t n_bytes, uint32_t *crc_out)
{
    if (!e || !crc_out || (!data && n_bytes)) return fail(ICW_E_ARG, "NULL argument");
    CK(cudaSetDevice(e->device));
    const size_t block = (size_t)64 << 20;
    int rc = e->io_in.reserve(n_bytes < block ? (n_bytes ? n_bytes : 16) : block);
    if (rc) return rc;
    uint32_t crc = 0;                                   // CRC of the empty message
    for (size_t off = 0; off < n_bytes; off += block) {
        const size_t n = n_bytes - off < block ? n_bytes - off : block;
        uint32_t part;
        CK(cudaMemcpyAsync(e->io_in.p, (const uint8_t *)data + off, n, cudaMemcpyHostToDevice, e->stream));
        rc = icw_crc32_device(e, e->io_in.p, n, &part);
        if (rc) return rc;
        crc = off ? crc32_combine(crc, part, n) : part;
    }
    *crc_out = crc;
    return ICW_OK;
}

extern "C" uint32_t icw_crc32_combine(uint32_t crc_a, uint32_t crc_b, uint64_t len_b)
{
    return crc32_combine(crc_a, crc_b, len_b);
}

// test-only taps: device buffers that receive the whole bus and the master output per frame
extern "C" int icw_session_set_taps(icw_session *s, double *d_tap_bus, double *d_tap_lr)
{
    if (!s) return fail(ICW_E_ARG, "NULL session");
    s->d_tap_bus = d_tap_bus;
    s->d_tap_lr = d_tap_lr;
    return ICW_OK;
}

// ---------------------------------------------------------------------------------------------
// the hot call
// ---------------------------------------------------------------------------------------------
struct DitherWords { const uint32_t *l = nullptr, *r = nullptr; size_t stream_stride = 0; cudaEvent_t join = nullptr; };

// words [first, first + n) of channel c's generator (seeded with `seed`) into dst
static int generate_words(icw_session *s, int c, uint32_t seed, uint64_t first, int64_t n, uint32_t *dst, cudaStream_t st)
{
    icw_engine *e = s->e;
    int rc = e->mt.generate(seed, first, n, dst, e->sm_count, st, &s->launches);
    if (rc) return fail(rc, "%s", e->mt.error());
    for (const auto &pt : s->mt_patches)        // test hook: one word of the stream replaced (icw_debug_patch_mt_word)
        if (pt.chan == c && pt.idx >= first && pt.idx < first + (uint64_t)n)
            CK(cudaMemcpyAsync(dst + (pt.idx - first), &pt.value, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    return ICW_OK;
}

static int make_dither_words(icw_session *s, int64_t n_frames, cudaStream_t st, DitherWords &dw)
{
    int mt_shared;
    const uint32_t *wl, *wr;
    icw_engine *e = s->e;
    const int wps = s->ch.render.words_per_sample;
    wl = wr = nullptr;
    mt_shared = 1;
    if (!wps) return ICW_OK;
    const int K = s->n_streams;
    for (int c = 0; c < 2; ++c)
        for (int k = 1; k < K; ++k)
            if (s->mt_seed[c][k] != s->mt_seed[c][0] || s->mt_drawn[c][k] != s->mt_drawn[c][0]) mt_shared = 0;
    for (int c = 0; c < 2; ++c)
        for (int k = 0; k < K; ++k)
            if (s->mt_drawn[c][k] % 2u)
                return fail(ICW_E_UNSUPPORTED, "stream %d channel %d: generator offset %llu is odd (every draw takes two words)", k, c,
                            (unsigned long long)s->mt_drawn[c][k]);
    // a frame replayed with the reference's rejection loop may take more words than its share: room for 32 rejected pairs
    const int64_t words = n_frames * wps + ((s->ch.feedback & 2) ? 64 : 0);
    const int groups = mt_shared ? 1 : K;
    for (int c = 0; c < 2; ++c) {
        int rc = e->mtw[c].reserve((size_t)groups * (size_t)words * sizeof(uint32_t));
        if (rc) return rc;
        for (int g = 0; g < groups; ++g) {
            uint32_t *dst = (uint32_t *)e->mtw[c].p + (size_t)g * (size_t)words;
            rc = generate_words(s, c, s->mt_seed[c][g], s->mt_drawn[c][g], words, dst, st);
            if (rc) return rc;
        }
    }
    wl = (const uint32_t *)e->mtw[0].p;
    wr = (const uint32_t *)e->mtw[1].p;
    dw.l = wl; dw.r = wr;
    dw.stream_stride = mt_shared ? 0 : (size_t)words;
    return ICW_OK;
}

static int get_scan_plan(icw_engine *e, int filter_no, bool baseline, double d0, int L, icw_engine::ScanPlan **out)
{
    icw_engine::ScanPlan &pl = e->scan[filter_no][baseline ? 1 : 0][L == SF_LC ? 3 : L <= SCAN_L ? 0 : L <= 4 * SCAN_L ? 1 : 2];
    if (pl.ready && pl.mc.L != L) {             // only when ICW_SCAN_L forces odd lengths: rebuild in place
        CK(cudaStreamSynchronize(e->stream));
        cudaFree(pl.d_pw);
        pl.d_pw = nullptr;
        pl.ready = false;
    }
    if (!pl.ready) {
        std::vector<double> pw;
        scan_make_coef(filter_no, baseline, d0, L, pl.mc, pw);
        if (cudaMalloc(&pl.d_pw, pw.size() * sizeof(double)) != cudaSuccess) { cudaGetLastError(); return fail(ICW_E_NOMEM, "cudaMalloc(scan power table) failed"); }
        CK(cudaMemcpy(pl.d_pw, pw.data(), pw.size() * sizeof(double), cudaMemcpyHostToDevice));
        pl.ready = true;
    }
    *out = &pl;
    return ICW_OK;
}

// one launch group: n_frames of every stream, state read from and left in the DevStream array
// `pre`: dither words generated once for the whole call (pointing at this group's first frame), or NULL
// `fuse_mt`: no word buffers at all -- chain_mt_kernel regenerates the dither inside the pointwise pass
// scan mode, one long stream: the whole chain in one kernel (icw_sfused.cu)
static icw_engine::SfPlan &get_sf_plan(icw_engine *e, int filter_no, bool baseline, double d0)
{
    icw_engine::SfPlan &p = e->sf[filter_no][baseline ? 1 : 0];
    if (!p.ready) {
        sfused_make_tables(filter_no, baseline, d0, p.tb);
        p.warm = sfused_warm_frames(p.tb);
        p.ready = true;
    }
    return p;
}

static int process_sfused(icw_session *s, int64_t n_frames, const uint8_t *d_in, uint8_t *d_out, cudaStream_t st)
{
    icw_engine *e = s->e;
    const DevChain &ch = s->ch;
    const icw_engine::SfPlan &sp = get_sf_plan(e, s->spec.filter_no, !s->spec.is_kahan, s->coef.d0);
    const int wps = ch.render.words_per_sample;
    const int slots = e->sm_count * SF_CTAS_PER_SM;         // one wave of units
    MtPlan pl[2];
    int rc;
    if (wps) {
        ProfSpan ps(s, st, ICW_K_MT);
        const uint32_t seeds[2] = { s->mt_seed[0][0], s->mt_seed[1][0] };
        rc = e->mt.plan_pair(seeds, s->mt_drawn[0][0], n_frames * wps, slots, e->sm_count, st, &s->launches, pl);
        if (rc) return fail(rc, "%s", e->mt.error());
    }
    {
        ProfSpan ps(s, st, ICW_K_SCAN_FUSED);
        CK(launch_scan_fused(sp.tb, ch, s->d_streams, n_frames, d_in, d_out, wps ? &pl[0] : nullptr, wps ? &pl[1] : nullptr,
                             slots, sp.warm, st));
    }
    s->launches++;
    {
        ProfSpan ps(s, st, ICW_K_MISC);
        CK(launch_advance(ch, s->d_streams, 1, n_frames, 1, st));
    }
    s->launches++;
    const uint64_t words = (uint64_t)n_frames * (uint64_t)wps;
    for (int c = 0; c < 2; ++c) s->mt_drawn[c][0] += words;
    return ICW_OK;
}

static int process_group(icw_session *s, int64_t n_frames, const uint8_t *d_in, size_t in_stride,
                         uint8_t *d_out, size_t out_stride, cudaStream_t st, const DitherWords *pre, bool fuse_mt, bool sfused = false)
{
    icw_engine *e = s->e;
    const DevChain &ch = s->ch;
    const int K = s->n_streams;
    int rc;
    if (sfused) return process_sfused(s, n_frames, d_in, d_out, st);
    DitherWords dw;
    if (fuse_mt) { }
    else if (pre) dw = *pre;
    else {
        ProfSpan ps(s, st, ICW_K_MT);
        rc = make_dither_words(s, n_frames, st, dw);
        if (rc) return rc;
    }
    const uint32_t *wl = dw.l, *wr = dw.r;
    const size_t mt_shared = dw.stream_stride;      // words between consecutive streams' dither (0 = shared)
    const bool scan = !ch.is_complex && s->spec.hilbert_mode == ICW_HILBERT_SCAN;
    const bool shaped = ch.render.ns_kind != 0;     // error feedback through the quantiser: serial per channel
    if (!ch.is_complex && !scan && !e->unfused && !ch.fp_check && !ch.feedback) {
        // real input, reference-exact Hilbert: the whole chain in one kernel; with a noise shaper on it stops in front of
        // the quantiser ((value, dither) pairs in `pre`, 32 B/frame) and the serial quantiser follows
        if (dw.join) CK(cudaStreamWaitEvent(st, dw.join, 0));
        double *pre = nullptr;
        if (shaped) {
            rc = e->ns_pre.reserve((size_t)n_frames * 4 * sizeof(double) * (size_t)K);
            if (rc) return rc;
            pre = (double *)e->ns_pre.p;
        }
        {
            ProfSpan ps(s, st, ICW_K_HILBERT);
            CK(launch_hb_fused(s->coef, ch, s->d_streams, K, n_frames, d_in, in_stride, wl, wr, mt_shared,
                               d_out, out_stride, s->d_tap_bus, s->d_tap_lr, pre, st));
            s->launches++;
        }
        if (shaped) {
            ProfSpan ps(s, st, ICW_K_CHAIN);
            CK(launch_ns_render(ch, s->d_streams, K, n_frames, pre, d_out, out_stride, st));
            s->launches++;
        }
    } else {
        const uint8_t *src = d_in;
        size_t src_stride = in_stride;
        int from_analytic = 0;
        if (!ch.is_complex) {
            const size_t per = (size_t)n_frames * 4 * sizeof(double);
            rc = e->analytic.reserve(per * (size_t)K);
            if (rc) return rc;
            if (scan) {
                // time-parallel modal scan (icw_scan.cu): chunk end states -> carries -> apply
                icw_engine::ScanPlan *plp;
                const int L = scan_chunk_len(K, n_frames, e->sm_count);
                rc = get_scan_plan(e, s->spec.filter_no, !s->spec.is_kahan, s->coef.d0, L, &plp);
                if (rc) return rc;
                icw_engine::ScanPlan &pl = *plp;
                rc = e->scan_scratch.reserve(scan_scratch_doubles(K, n_frames, L) * sizeof(double));
                if (rc) return rc;
                int nl = 0;
                // profiling: one span for passes 1 + 2, one for pass 3 (two events recorded between them)
                cudaEvent_t a = nullptr, m0 = nullptr, m1 = nullptr, b = nullptr;
                if (s->profiling) {
                    a = ProfSpan::take(s); m0 = ProfSpan::take(s); m1 = ProfSpan::take(s); b = ProfSpan::take(s);
                    cudaEventRecord(a, st);
                }
                CK(launch_hb_scan(pl.mc, ch, s->d_streams, K, n_frames, src, in_stride, pl.d_pw,
                                  (double *)e->scan_scratch.p, (double *)e->analytic.p, st, &nl, m0, m1));
                if (a) {
                    cudaEventRecord(b, st);
                    s->spans.push_back({ ICW_K_SCAN_LOCAL, a, m0 });
                    s->spans.push_back({ ICW_K_SCAN_APPLY, m1, b });
                }
                s->launches += nl;
            } else {
                // unfused exact pair (ICW_UNFUSED=1): recurrences -> analytic scratch -> pointwise kernel
                ProfSpan ps(s, st, ICW_K_HILBERT);
                CK(launch_hb_exact(s->coef, ch, s->d_streams, K, n_frames, src, in_stride, (double *)e->analytic.p, st));
                s->launches++;
            }
            src = (const uint8_t *)e->analytic.p;
            src_stride = per;
            from_analytic = 1;
        }
        if (dw.join) CK(cudaStreamWaitEvent(st, dw.join, 0));
        double *pre = nullptr;
        if (shaped) {
            rc = e->ns_pre.reserve((size_t)n_frames * 4 * sizeof(double) * (size_t)K);
            if (rc) return rc;
            pre = (double *)e->ns_pre.p;
        }
        if (fuse_mt) {
            MtPlan pl[2];
            {
                ProfSpan ps(s, st, ICW_K_MT);
                const uint32_t seeds[2] = { s->mt_seed[0][0], s->mt_seed[1][0] };
                rc = e->mt.plan_pair(seeds, s->mt_drawn[0][0], n_frames * ch.render.words_per_sample,
                                     chain_mt_max_units(e->sm_count), e->sm_count, st, &s->launches, pl);
                if (rc) return fail(rc, "%s", e->mt.error());
            }
            ProfSpan ps(s, st, ICW_K_CHAIN);
            CK(launch_chain_mt(ch, s->d_streams, n_frames, src, from_analytic, pl[0], pl[1], d_out,
                               s->d_tap_bus, s->d_tap_lr, pre, st));
        } else {
            ProfSpan ps(s, st, ICW_K_CHAIN);
            CK(launch_chain(ch, s->d_streams, K, n_frames, src, src_stride, from_analytic, wl, wr, mt_shared,
                            d_out, out_stride, s->d_tap_bus, s->d_tap_lr, pre, e->sm_count, st));
        }
        s->launches++;
        if (shaped) {
            CK(launch_ns_render(ch, s->d_streams, K, n_frames, pre, d_out, out_stride, st));
            s->launches++;
        }
    }
    {
        ProfSpan ps(s, st, ICW_K_MISC);
        CK(launch_advance(ch, s->d_streams, K, n_frames, !ch.is_complex, st));
    }
    s->launches++;
    const uint64_t words = (uint64_t)n_frames * (uint64_t)ch.render.words_per_sample;
    for (int c = 0; c < 2; ++c)
        for (int k = 0; k < K; ++k) s->mt_drawn[c][k] += words;
    return ICW_OK;
}

// every stream's Hilbert memory from the session's live basis to `to_basis` (0 delay line, 1 modal)
static int convert_basis(icw_session *s, int to_basis)
{
    struct Arg { int ft, to, bad; } arg = { s->spec.filter_no, to_basis, 0 };
    int rc = rewrite_states(s, [](DevStream &d, void *a) {
        Arg &g = *(Arg *)a;
        for (int c = 0; c < 2; ++c)
            for (int f = 0; f < 2; ++f)
                if (icw_host_hb_convert(g.ft, g.to, d.hb[c][f], d.hb[c][f]) != 0) g.bad = 1;
        d.hb_basis = (uint32_t)g.to;
    }, &arg);
    if (rc) return rc;
    if (arg.bad) return fail(ICW_E_ARG, "filter state conversion failed for design %d", s->spec.filter_no);
    s->hb_basis = to_basis;
    return ICW_OK;
}

// ---- one API call = begin (checks, dither words for the whole call) + ranges + end ------------------
struct CallCtx {
    DitherWords all;
    bool pre = false;
    bool fuse_mt = false;   // dither regenerated inside the pointwise kernel: no word buffers
    bool sfused = false;    // scan mode in one kernel (icw_sfused.cu): no scratch at all, the call is one launch
    int64_t step = 0;       // frames per launch group
    int mode = 0;
    bool real_in = false;
};

// frames per launch group when the group's dither comes from chain_mt_kernel: its CTAs are the
// generator's jump-ahead units, so a group should be the whole call (analytic scratch: 32 B/frame)
constexpr int64_t FUSE_MT_GROUP = (int64_t)1 << 30;
constexpr int64_t BIG_GROUP = (int64_t)1 << 30;     // frames per launch group of one long stream (34 GB of analytic scratch)

// `one_range`: the caller hands the whole call over in a single call_range()
static int call_begin(icw_session *s, int64_t n_total, cudaStream_t st, CallCtx &cx, bool one_range)
{
    icw_engine *e = s->e;
    const DevChain &ch = s->ch;
    const int K = s->n_streams;
    cx.real_in = !ch.is_complex;
    cx.mode = s->spec.hilbert_mode;
    if (e->last_stream_valid && e->last_stream != st) CK(cudaStreamWaitEvent(st, e->ev_scratch, 0));
    if (s->last_stream && s->last_stream != st && s->last_stream != e->last_stream) CK(cudaStreamSynchronize(s->last_stream));
    if (cx.real_in) {
        if (cx.mode != ICW_HILBERT_EXACT && cx.mode != ICW_HILBERT_SCAN) return fail(ICW_E_ARG, "hilbert_mode out of range");
        if (s->hb_live && s->hb_basis != cx.mode) {
            // the filter memory goes over to the other basis (icw_hbconv.cpp: binary128 on the host); the reference
            // changes filter settings on a live stream as well (src/in_cwave.c:135-191)
            int rc = convert_basis(s, cx.mode);
            if (rc) return rc;
        }
    }
    if (cx.real_in && cx.mode == ICW_HILBERT_SCAN && ch.fp_check)
        return fail(ICW_E_UNSUPPORTED, "FP_CHECK counts exceptional intermediates of the reference's own recurrences: "
                                       "exact Hilbert mode only (the modal scan has different intermediates)");
    // scan mode and the unfused path go through per-frame scratch: bound it by walking the call in groups
    // per-frame scratch: the analytic signal (real input unless the fused exact kernel runs) and, with a noise
    // shaper, the (value, dither) pairs between the pointwise pass and the serial quantiser -- 32 B/frame each
    const bool shaped = ch.render.ns_kind != 0;
    const bool scratchy = (cx.real_in && (cx.mode == ICW_HILBERT_SCAN || e->unfused || ch.fp_check || ch.feedback || shaped)) || shaped;
    // one stream: groups as long as the scratch may grow -- long groups let the scan use long chunks
    // (scan_chunk_len); many streams: 2^25 frames over ALL of them, never less than one scan tile per stream
    const int64_t seg = !scratchy ? n_total
                      : K == 1 ? BIG_GROUP : SCAN_SEGMENT / K / (SCAN_L * SCAN_CH) * (SCAN_L * SCAN_CH);
    cx.step = seg < SCAN_L * SCAN_CH ? SCAN_L * SCAN_CH : seg;
    const int wps = ch.render.words_per_sample;
    // scan mode, one stream, straight-line list: everything in one kernel.  A unit (one CTA) first runs the filters over
    // `warm` frames before its own: that only pays on streams long enough to give every SM a unit many times that long
    if (cx.real_in && cx.mode == ICW_HILBERT_SCAN && one_range && e->sfused != 0 && !e->unfused && !ch.feedback &&
        sfused_supports(ch, K, s->d_tap_bus || s->d_tap_lr) &&
        (!wps || (s->mt_drawn[0][0] == s->mt_drawn[1][0] && s->mt_drawn[0][0] % (uint64_t)wps == 0))) {
        bool go = e->sfused == 1;
        // ... and on streams with dither: without it the generator warps idle and two scan warps a CTA are all the SM gets for
        // the filters -- the three-kernel path, analytic round trip and all, is then the faster one (C5 shape, 400 M frames of
        // f32 without dither: 11.4 ms against 16.2; C2, i24 with TPDF: 30.8 against 26.1)
        if (!go) go = wps != 0 && n_total / (e->sm_count * SF_CTAS_PER_SM) >= 2 * get_sf_plan(e, s->spec.filter_no, !s->spec.is_kahan, s->coef.d0).warm;
        if (go) { cx.sfused = true; cx.step = n_total; return ICW_OK; }
    }
    // the fused exact kernel reads word buffers; everything else that ends in chain_kernel can make its own
    const bool hb_fused = cx.real_in && cx.mode == ICW_HILBERT_EXACT && !e->unfused && !ch.fp_check && !ch.feedback;
    if (wps && one_range && !hb_fused && !e->no_fuse_mt && !ch.feedback && K == 1 && chain_mt_supports(ch) &&
        s->mt_drawn[0][0] == s->mt_drawn[1][0] && s->mt_drawn[0][0] % (uint64_t)wps == 0) {
        cx.fuse_mt = true;
        if (scratchy) cx.step = n_total < FUSE_MT_GROUP ? (n_total < cx.step ? cx.step : n_total) : FUSE_MT_GROUP;
        return ICW_OK;
    }
    // dither words for the whole call in one go when they fit (one jump tree instead of one per group)
    if (wps && (double)n_total * wps * 4.0 * 2.0 * (K > 1 ? K : 1) <= 48e9) {
        // integer work on the aux stream next to the FP64-bound Hilbert kernels on `st`; the word
        // buffers may still be read by the previous call, so the aux stream first joins `st`
        CK(cudaEventRecord(e->ev_fork, st));
        CK(cudaStreamWaitEvent(e->aux, e->ev_fork, 0));
        {
            ProfSpan ps(s, e->aux, ICW_K_MT);
            int rc = make_dither_words(s, n_total, e->aux, cx.all);
            if (rc) return rc;
        }
        CK(cudaEventRecord(e->ev_join, e->aux));
        cx.all.join = e->ev_join;
        cx.pre = true;
    } else if (wps) {
        cx.step = std::min<int64_t>(cx.step, SCAN_SEGMENT);     // per-group word buffers
    }
    return ICW_OK;
}

// frames [f0, f0 + n) of the call; d_in / d_out point at frame f0
static int call_range(icw_session *s, CallCtx &cx, int64_t f0, int64_t n, const uint8_t *d_in, size_t in_stride,
                      uint8_t *d_out, size_t out_stride, cudaStream_t st)
{
    const DevChain &ch = s->ch;
    const int wps = ch.render.words_per_sample;
    if ((s->d_tap_bus || s->d_tap_lr) && n > cx.step)
        return fail(ICW_E_ARG, "taps are a test aid for calls of at most %lld frames", (long long)cx.step);
    for (int64_t g0 = 0; g0 < n; g0 += cx.step) {
        const int64_t gn = n - g0 < cx.step ? n - g0 : cx.step;
        DitherWords here = cx.all;
        if (cx.pre) {
            here.l += (size_t)(f0 + g0) * wps; here.r += (size_t)(f0 + g0) * wps;
            here.join = cx.all.join;                            // waited for once, by the first group that runs
            cx.all.join = nullptr;
        }
        int rc = process_group(s, gn, d_in + (size_t)g0 * ch.frame_bytes, in_stride,
                               d_out + (size_t)g0 * ch.out_frame_bytes, out_stride, st, cx.pre ? &here : nullptr, cx.fuse_mt, cx.sfused);
        if (rc) return rc;
    }
    return ICW_OK;
}

static void call_end(icw_session *s, const CallCtx &cx, cudaStream_t st)
{
    if (cx.real_in) { s->hb_basis = cx.mode; s->hb_live = true; }
    icw_engine *e = s->e;
    cudaEventRecord(e->ev_scratch, st);
    e->last_stream = st; e->last_stream_valid = true;
    s->last_stream = st;
}

static void note_alignment(icw_session *s, const void *d_in, size_t in_stride)
{
    // typed loads are only legal when every sample sits on its natural alignment
    // (a complex format's unit is its (I, Q) pair: one 32- or 64-bit word, or two doubles; the 6-byte i16 + f32 pair never is)
    const int cb = s->ch.chan_bytes;
    size_t a = (cb == 2 || cb == 4) ? (size_t)cb : 0;
    if (s->ch.is_complex) a = s->ch.fmt == ICW_FMT_CW_I16 ? 4 : (s->ch.fmt == ICW_FMT_CW_F32 || s->ch.fmt == ICW_FMT_CW_F64) ? 8 : 0;
    s->ch.aligned = a && ((size_t)(uintptr_t)d_in % a == 0) && (s->n_streams == 1 || in_stride % a == 0);
}

extern "C" int icw_session_process_device(icw_session *s, int64_t n_frames, const void *d_in, size_t in_stride,
                                          void *d_out, size_t out_stride, void *cuda_stream)
{
    if (!s || !d_in || !d_out) return fail(ICW_E_ARG, "NULL argument");
    if (n_frames < 0) return fail(ICW_E_ARG, "negative frame count");
    if (n_frames == 0) return ICW_OK;
    icw_engine *e = s->e;
    note_alignment(s, d_in, in_stride);
    const DevChain &ch = s->ch;
    if (s->n_streams > 1 && (in_stride < (size_t)n_frames * ch.frame_bytes || out_stride < (size_t)n_frames * ch.out_frame_bytes))
        return fail(ICW_E_ARG, "stream strides are shorter than one stream's data");
    CK(cudaSetDevice(e->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : e->stream;
    CallCtx cx;
    int rc = call_begin(s, n_frames, st, cx, true);
    if (rc) return rc;
    rc = call_range(s, cx, 0, n_frames, (const uint8_t *)d_in, in_stride, (uint8_t *)d_out, out_stride, st);
    if (rc) return rc;
    call_end(s, cx, st);
    return ICW_OK;
}

// Host buffers: the call is cut into time segments that flow through a three-stage pipeline --
// H2D copy (copy stream), kernels (engine stream), D2H copy (second copy stream) -- over
// double-buffered device staging, so that on a long call the PCIe transfers of neighbouring
// segments hide behind the kernels.  Works from pageable memory too (the copies then stage
// through the driver and overlap less).
static int host_run(icw_session *s, int64_t n_frames, const void *in, size_t in_stride, void *out, size_t out_stride)
{
    icw_engine *e = s->e;
    const DevChain &ch = s->ch;
    const int K = s->n_streams;
    const size_t in_row = (size_t)n_frames * ch.frame_bytes, out_row = (size_t)n_frames * ch.out_frame_bytes;
    if (K > 1 && (in_stride < in_row || out_stride < out_row)) return fail(ICW_E_ARG, "stream strides too short");
    const size_t h_in_stride = K > 1 ? in_stride : in_row, h_out_stride = K > 1 ? out_stride : out_row;
    CK(cudaSetDevice(e->device));

    // segment length: a multiple of the scan tile, about 2^24 frames over all streams
    int64_t seg = ((int64_t)1 << 24) / K;
    seg = seg / (SCAN_L * SCAN_CH) * (SCAN_L * SCAN_CH);
    if (seg < 4096) seg = 4096;
    if (seg > n_frames) seg = n_frames;
    const int nbuf = seg < n_frames ? 2 : 1;
    // device rows are padded to 16 bytes so every stream starts on a vector boundary
    const size_t din_stride = ((size_t)seg * ch.frame_bytes + 15) & ~(size_t)15;
    const size_t dout_stride = ((size_t)seg * ch.out_frame_bytes + 15) & ~(size_t)15;
    int rc = e->io_in.reserve(din_stride * K * nbuf);
    if (rc) return rc;
    rc = e->io_out.reserve(dout_stride * K * nbuf);
    if (rc) return rc;
    if (!e->h2d) {
        CK(cudaStreamCreateWithFlags(&e->h2d, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&e->d2h, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            CK(cudaEventCreateWithFlags(&e->ev_in[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&e->ev_comp[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&e->ev_out[i], cudaEventDisableTiming));
        }
    }
    note_alignment(s, e->io_in.p, din_stride);
    cudaStream_t st = e->stream;
    CallCtx cx;
    rc = call_begin(s, n_frames, st, cx, nbuf == 1);
    if (rc) return rc;
    int64_t k = 0;
    for (int64_t f0 = 0; f0 < n_frames; f0 += seg, ++k) {
        const int64_t n = n_frames - f0 < seg ? n_frames - f0 : seg;
        const int b = (int)(k & 1) % nbuf;
        uint8_t *din = (uint8_t *)e->io_in.p + (size_t)b * din_stride * K;
        uint8_t *dout = (uint8_t *)e->io_out.p + (size_t)b * dout_stride * K;
        // the staging input of this slot is free once the kernels of segment k-2 are done
        if (k >= 2) CK(cudaStreamWaitEvent(e->h2d, e->ev_comp[b], 0));
        CK(cudaMemcpy2DAsync(din, din_stride, (const uint8_t *)in + (size_t)f0 * ch.frame_bytes, h_in_stride,
                             (size_t)n * ch.frame_bytes, K, cudaMemcpyHostToDevice, e->h2d));
        CK(cudaEventRecord(e->ev_in[b], e->h2d));
        CK(cudaStreamWaitEvent(st, e->ev_in[b], 0));
        if (k >= 2) CK(cudaStreamWaitEvent(st, e->ev_out[b], 0));   // and its staging output once segment k-2 left
        rc = call_range(s, cx, f0, n, din, din_stride, dout, dout_stride, st);
        if (rc) return rc;
        CK(cudaEventRecord(e->ev_comp[b], st));
        CK(cudaStreamWaitEvent(e->d2h, e->ev_comp[b], 0));
        CK(cudaMemcpy2DAsync((uint8_t *)out + (size_t)f0 * ch.out_frame_bytes, h_out_stride, dout, dout_stride,
                             (size_t)n * ch.out_frame_bytes, K, cudaMemcpyDeviceToHost, e->d2h));
        CK(cudaEventRecord(e->ev_out[b], e->d2h));
    }
    CK(cudaStreamSynchronize(e->d2h));
    CK(cudaStreamSynchronize(st));
    call_end(s, cx, st);
    return ICW_OK;
}

// The streams are idle and some kernel of the call that just ran has counted a dither draw that met the generator's
// rejection loop.  Which frame of the call was the first?  The kernels keep a count and nothing else, so the answer is
// looked up in the generators' own words: every stream whose count moved has the words its channels drew in this call
// made again (in pieces of 2^24) and searched for the first pair the reference would throw away.  Streams that stand at
// the same place of the same generator share the answer.  *frame = LLONG_MAX if nothing was counted.
static int first_redraw_frame(icw_session *s, const std::vector<unsigned long long> &cnt0, const std::vector<uint64_t> (&drawn0)[2],
                              int64_t n_frames, int64_t *frame)
{
    icw_engine *e = s->e;
    const int K = s->n_streams, wps = s->ch.render.words_per_sample;
    std::vector<unsigned long long> cnt((size_t)K);
    CK(cudaMemcpy2D(cnt.data(), sizeof(unsigned long long), &s->d_streams[0].mt_redraws, sizeof(DevStream), sizeof(unsigned long long),
                    (size_t)K, cudaMemcpyDeviceToHost));
    *frame = LLONG_MAX;
    std::map<std::pair<uint32_t, uint64_t>, long long> seen;       // (seed, first word) -> first rejected word or LLONG_MAX
    long long *d_first = nullptr;
    const int64_t piece = (int64_t)1 << 24;
    for (int k = 0; k < K; ++k) {
        if (cnt[(size_t)k] == cnt0[(size_t)k]) continue;
        for (int c = 0; c < 2; ++c) {
            const uint32_t seed = s->mt_seed[c][(size_t)k];
            const uint64_t w0 = drawn0[c][(size_t)k];
            auto key = std::make_pair(seed, w0);
            auto it = seen.find(key);
            long long first = LLONG_MAX;
            if (it != seen.end()) first = it->second;
            else {
                if (!d_first && cudaMalloc(&d_first, sizeof(long long)) != cudaSuccess) { cudaGetLastError(); return fail(ICW_E_NOMEM, "cudaMalloc failed"); }
                int rc = e->mtw[c].reserve((size_t)piece * sizeof(uint32_t));
                if (rc) { cudaFree(d_first); return rc; }
                const int64_t total = n_frames * wps;
                for (int64_t off = 0; off < total && first == LLONG_MAX; off += piece) {
                    const int64_t nw = total - off < piece ? total - off : piece;
                    const long long init = LLONG_MAX;
                    CK(cudaMemcpyAsync(d_first, &init, sizeof init, cudaMemcpyHostToDevice, e->stream));
                    rc = generate_words(s, c, seed, w0 + (uint64_t)off, nw, (uint32_t *)e->mtw[c].p, e->stream);
                    if (rc) { cudaFree(d_first); return rc; }
                    CK(launch_mt_find_reject((const uint32_t *)e->mtw[c].p, nw / 2, (long long)((w0 + (uint64_t)off) / 2), d_first, e->sm_count, e->stream));
                    CK(cudaMemcpyAsync(&first, d_first, sizeof first, cudaMemcpyDeviceToHost, e->stream));
                    CK(cudaStreamSynchronize(e->stream));
                    if (first != LLONG_MAX) first *= 2;             // pair number -> word number
                }
                seen[key] = first;
            }
            if (first != LLONG_MAX) *frame = std::min<int64_t>(*frame, (int64_t)(((uint64_t)first - w0) / (uint64_t)wps));
        }
    }
    if (d_first) cudaFree(d_first);
    return ICW_OK;
}

// Host entry point.  The kernels draw the dither by frame index (draw j of a channel = words 2j, 2j+1 of its generator);
// the reference draws word after word and throws a pair away when it maps to -1 (mtrnd_gen_dsopen,
// src/mersene_twister/mt_jrnd.c:245-256; probability 2^-53 a draw).  A call whose kernels met such a pair is REPLAYED: the
// streams go back to where they stood when the call began, the frames before the event run again as they did, the one frame
// with the event runs through the frame-serial route with the reference's own loop (its generator ends two words further
// on), and the rest of the call follows from there -- the output is the reference's for every frame.  The input is still in
// the caller's buffer, which is why this is the host entry point's privilege (the device entry point reports the event).
extern "C" int icw_session_process_host(icw_session *s, int64_t n_frames, const void *in, size_t in_stride,
                                        void *out, size_t out_stride)
{
    if (!s || !in || !out) return fail(ICW_E_ARG, "NULL argument");
    if (n_frames <= 0) return n_frames ? fail(ICW_E_ARG, "negative frame count") : ICW_OK;
    const int K = s->n_streams;
    const DevChain &ch = s->ch;
    if (!ch.render.words_per_sample) return host_run(s, n_frames, in, in_stride, out, out_stride);
    const size_t in_row = (size_t)n_frames * ch.frame_bytes, out_row = (size_t)n_frames * ch.out_frame_bytes;
    if (K > 1 && (in_stride < in_row || out_stride < out_row)) return fail(ICW_E_ARG, "stream strides too short");
    // a sub-range of the call keeps the caller's row strides (one stream: the rows are the whole buffers)
    const size_t hs_in = K > 1 ? in_stride : in_row, hs_out = K > 1 ? out_stride : out_row;
    CK(cudaSetDevice(s->e->device));
    int rc = quiesce(s);
    if (rc) return rc;
    if (!s->d_snap && cudaMalloc(&s->d_snap, sizeof(DevStream) * (size_t)K) != cudaSuccess) {
        cudaGetLastError();
        return fail(ICW_E_NOMEM, "cudaMalloc(state snapshot) failed");
    }
    auto sub = [&](int64_t f0, int64_t n) {
        // rows of K > 1 streams keep their stride; a single stream's sub-range is contiguous
        return host_run(s, n, (const uint8_t *)in + (size_t)f0 * ch.frame_bytes, K > 1 ? hs_in : 0,
                        (uint8_t *)out + (size_t)f0 * ch.out_frame_bytes, K > 1 ? hs_out : 0);
    };
    int64_t done = 0;
    for (int guard = 0; done < n_frames; ++guard) {
        // where the streams stand now
        CK(cudaMemcpy(s->d_snap, s->d_streams, sizeof(DevStream) * (size_t)K, cudaMemcpyDeviceToDevice));
        std::vector<unsigned long long> cnt0((size_t)K);
        CK(cudaMemcpy2D(cnt0.data(), sizeof(unsigned long long), &s->d_streams[0].mt_redraws, sizeof(DevStream), sizeof(unsigned long long),
                        (size_t)K, cudaMemcpyDeviceToHost));
        const std::vector<uint64_t> drawn0[2] = { s->mt_drawn[0], s->mt_drawn[1] };
        const int basis0 = s->hb_basis;
        const bool live0 = s->hb_live;
        rc = sub(done, n_frames - done);
        if (rc) return rc;
        int64_t f;
        rc = first_redraw_frame(s, cnt0, drawn0, n_frames - done, &f);
        if (rc) return rc;
        if (f == LLONG_MAX) return ICW_OK;
        if (guard >= 64) return fail(ICW_E_MT_REDRAW, "more than 64 rejected dither draws in one call");
        // back to the start of this stretch, then: the frames before the event as before, the event's frame the reference's way
        CK(cudaMemcpy(s->d_streams, s->d_snap, sizeof(DevStream) * (size_t)K, cudaMemcpyDeviceToDevice));
        s->mt_drawn[0] = drawn0[0]; s->mt_drawn[1] = drawn0[1];
        s->hb_basis = basis0; s->hb_live = live0;
        if (f > 0) {
            rc = sub(done, f);
            if (rc) return rc;
        }
        s->ch.feedback |= 2;
        rc = sub(done + f, 1);
        s->ch.feedback &= ~2;
        if (rc) return rc;
        // the generators stand where the serial draws left them
        for (int c = 0; c < 2; ++c)
            CK(cudaMemcpy2D(s->mt_drawn[c].data(), sizeof(uint64_t), &s->d_streams[0].mt_drawn[c], sizeof(DevStream), sizeof(uint64_t),
                            (size_t)K, cudaMemcpyDeviceToHost));
        ++s->redraws_handled;
        done += f + 1;
    }
    return ICW_OK;
}

extern "C" int icw_session_profile(icw_session *s, int on)
{
    if (!s) return fail(ICW_E_ARG, "NULL session");
    s->profiling = on != 0;
    return ICW_OK;
}

extern "C" int icw_session_profile_read(icw_session *s, icw_profile *out, int reset)
{
    if (!s || !out) return fail(ICW_E_ARG, "NULL argument");
    CK(cudaDeviceSynchronize());
    for (auto &sp : s->spans) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) { s->prof_ms[sp.k] += ms; s->prof_n[sp.k]++; }
        else cudaGetLastError();
        s->ev_pool.push_back(sp.a);
        s->ev_pool.push_back(sp.b);
    }
    s->spans.clear();
    for (int k = 0; k < ICW_K_COUNT; ++k) { out->ms[k] = s->prof_ms[k]; out->launches[k] = s->prof_n[k]; }
    if (reset) for (int k = 0; k < ICW_K_COUNT; ++k) { s->prof_ms[k] = 0; s->prof_n[k] = 0; }
    return ICW_OK;
}

extern "C" const char *icw_kernel_class_name(int k)
{
    static const char *names[ICW_K_COUNT] = { "hilbert", "chain", "mt", "misc", "scan_local", "scan_apply", "scan_fused" };
    return (k >= 0 && k < ICW_K_COUNT) ? names[k] : "?";
}

extern "C" int icw_session_stats(icw_session *s, icw_stats *out)
{
    if (!s || !out) return fail(ICW_E_ARG, "NULL argument");
    std::vector<DevStream> all((size_t)s->n_streams);
    int rc = quiesce(s);
    if (rc) return rc;
    CK(cudaMemcpy(all.data(), s->d_streams, sizeof(DevStream) * all.size(), cudaMemcpyDeviceToHost));
    memset(out, 0, sizeof *out);
    double pk[2] = { 0.0, 0.0 };
    for (const auto &d : all) {
        for (int c = 0; c < 2; ++c) {
            out->clips[c] += d.clips[c];
            pk[c] = std::max(pk[c], d.peak[c]);
            out->hb_rejects += d.hb_rejects[c][0] + d.hb_rejects[c][1];
        }
        out->mt_redraws += d.mt_redraws;
    }
    out->mt_redraws += s->redraws_handled;
    out->peak_db[0] = icw_peak_db(pk[0]);
    out->peak_db[1] = icw_peak_db(pk[1]);
    out->kernel_launches = s->launches;
    return ICW_OK;
}

extern "C" int icw_session_fp_stats(icw_session *s, int stream, uint32_t out[4][7])
{
    if (!s || !out) return fail(ICW_E_ARG, "NULL argument");
    if (stream < 0 || stream >= s->n_streams) return fail(ICW_E_ARG, "stream index out of range");
    { int rc = quiesce(s); if (rc) return rc; }
    CK(cudaMemcpy(out, (const uint8_t *)(s->d_streams + stream) + offsetof(DevStream, fp_cnt), sizeof(uint32_t) * 4 * 7,
                  cudaMemcpyDeviceToHost));
    return ICW_OK;
}

// ---------------------------------------------------------------------------------------------
// leaves
// ---------------------------------------------------------------------------------------------
extern "C" int icw_hilbert_device(icw_engine *e, int filter_no, int is_kahan, int is_reject, int mode,
                                  int n_chan, int64_t n, const double *d_x, double *d_out_iq,
                                  icw_stream_state *chan_state)
{
    if (!e || !d_x || !d_out_iq || !chan_state || n_chan < 1 || n < 0) return fail(ICW_E_ARG, "bad argument");
    if (filter_no < 0 || filter_no >= ICW_HB_NTYPES) return fail(ICW_E_ARG, "filter_no must be 0..5");
    if (mode != ICW_HILBERT_EXACT && mode != ICW_HILBERT_SCAN) return fail(ICW_E_ARG, "hilbert_mode out of range");
    CK(cudaSetDevice(e->device));
    if (mode == ICW_HILBERT_SCAN) {
        // every channel becomes a mono stream of doubles; the L half of the analytic frame is the answer
        if (n == 0) return ICW_OK;
        const double a0 = word_as_double(ICW_HB_A[filter_no][0]);
        const double d0 = word_as_double(ICW_HB_B[filter_no][0]) / a0;
        icw_engine::ScanPlan *pl;
        const int L = scan_chunk_len(n_chan, n, e->sm_count);
        int rc = get_scan_plan(e, filter_no, !is_kahan, d0, L, &pl);
        if (rc) return rc;
        DevChain ch;
        memset(&ch, 0, sizeof ch);
        ch.fmt = ICW_FMT_INTERNAL_F64; ch.n_channels = 1; ch.chan_bytes = 8; ch.frame_bytes = 8;
        std::vector<DevStream> ds((size_t)n_chan);
        for (int c = 0; c < n_chan; ++c) {
            memset(&ds[c], 0, sizeof(DevStream));
            for (int ch2 = 0; ch2 < 2; ++ch2) {
                memcpy(ds[c].hb[ch2], chan_state[c].hb[0], sizeof ds[c].hb[ch2]);
                ds[c].quad[ch2] = chan_state[c].quad[0] & 3u;
            }
        }
        rc = e->leaf.reserve(sizeof(DevStream) * ds.size());
        if (rc) return rc;
        rc = e->analytic.reserve((size_t)n_chan * (size_t)n * 4 * sizeof(double));
        if (rc) return rc;
        rc = e->scan_scratch.reserve(scan_scratch_doubles(n_chan, n, L) * sizeof(double));
        if (rc) return rc;
        CK(cudaMemcpyAsync(e->leaf.p, ds.data(), sizeof(DevStream) * ds.size(), cudaMemcpyHostToDevice, e->stream));
        int nl = 0;
        CK(launch_hb_scan(pl->mc, ch, (DevStream *)e->leaf.p, n_chan, n, (const uint8_t *)d_x, (size_t)n * 8, pl->d_pw,
                          (double *)e->scan_scratch.p, (double *)e->analytic.p, e->stream, &nl));
        CK(cudaMemcpy2DAsync(d_out_iq, 16, e->analytic.p, 32, 16, (size_t)n_chan * (size_t)n, cudaMemcpyDeviceToDevice, e->stream));
        CK(cudaMemcpyAsync(ds.data(), e->leaf.p, sizeof(DevStream) * ds.size(), cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        for (int c = 0; c < n_chan; ++c) {
            memcpy(chan_state[c].hb[0], ds[c].hb[0], sizeof ds[c].hb[0]);
            chan_state[c].quad[0] = (uint32_t)((chan_state[c].quad[0] + (uint64_t)n) & 3u);
            chan_state[c].hb_basis = 1;
        }
        return ICW_OK;
    }
    HbCoef coef;
    memset(&coef, 0, sizeof coef);
    const int ord = ICW_HB_ORDER[filter_no];
    double a0 = word_as_double(ICW_HB_A[filter_no][0]);
    coef.d0 = word_as_double(ICW_HB_B[filter_no][0]) / a0;
    for (int i = 0; i < ord; ++i) {
        coef.fb[i] = -word_as_double(ICW_HB_A[filter_no][i + 1]) / a0;
        coef.ff[i] = word_as_double(ICW_HB_B[filter_no][i + 1]) / a0;
    }
    std::vector<HbLeafState> hs((size_t)n_chan);
    for (int c = 0; c < n_chan; ++c) {
        memcpy(hs[c].z, chan_state[c].hb[0], sizeof hs[c].z);
        hs[c].rejects[0] = chan_state[c].hb_rejects[0][0];
        hs[c].rejects[1] = chan_state[c].hb_rejects[0][1];
        hs[c].quad = chan_state[c].quad[0] & 3u;
        hs[c].pad = 0;
    }
    int rc = e->leaf.reserve(sizeof(HbLeafState) * hs.size());
    if (rc) return rc;
    CK(cudaMemcpyAsync(e->leaf.p, hs.data(), sizeof(HbLeafState) * hs.size(), cudaMemcpyHostToDevice, e->stream));
    if (n) CK(launch_hb_leaf(coef, ord, is_kahan != 0, is_reject, n_chan, n, d_x, d_out_iq, (HbLeafState *)e->leaf.p, e->stream));
    CK(cudaMemcpyAsync(hs.data(), e->leaf.p, sizeof(HbLeafState) * hs.size(), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    for (int c = 0; c < n_chan; ++c) {
        memcpy(chan_state[c].hb[0], hs[c].z, sizeof hs[c].z);
        chan_state[c].hb_rejects[0][0] = hs[c].rejects[0];
        chan_state[c].hb_rejects[0][1] = hs[c].rejects[1];
        chan_state[c].quad[0] = hs[c].quad;
    }
    return ICW_OK;
}

extern "C" int icw_mt_words_device(icw_engine *e, uint32_t seed, uint64_t skip, int64_t n, uint32_t *d_out)
{
    if (!e || !d_out || n < 0) return fail(ICW_E_ARG, "bad argument");
    if (!n) return ICW_OK;
    CK(cudaSetDevice(e->device));
    uint64_t launches = 0;
    int rc = e->mt.generate(seed, skip, n, d_out, e->sm_count, e->stream, &launches);
    if (rc) return fail(rc, "%s", e->mt.error());
    CK(cudaStreamSynchronize(e->stream));
    return ICW_OK;
}

extern "C" int icw_debug_sincos_device(icw_engine *e, int64_t n, const double *d_x, double *d_out)
{
    if (!e || !d_x || !d_out || n < 0) return fail(ICW_E_ARG, "bad argument");
    if (!n) return ICW_OK;
    CK(cudaSetDevice(e->device));
    CK(launch_sincos_leaf(n, d_x, d_out, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return ICW_OK;
}

// test hook for the ICW_E_MT_REDRAW report: the event itself has probability 2^-53 a draw and no seed is known that
// produces it, so the test raises the stream's counter the way the kernels' commit would
extern "C" int icw_debug_note_redraw(icw_session *s, int k, uint64_t n)
{
    if (!s || k < 0 || k >= s->n_streams) return fail(ICW_E_ARG, "bad stream index");
    int rc = quiesce(s);
    if (rc) return rc;
    unsigned long long v = 0;
    CK(cudaMemcpy(&v, &s->d_streams[k].mt_redraws, sizeof v, cudaMemcpyDeviceToHost));
    v += n;
    CK(cudaMemcpy(&s->d_streams[k].mt_redraws, &v, sizeof v, cudaMemcpyHostToDevice));
    return ICW_OK;
}

// test hook for the replay of a rejected dither draw: word number `idx` (counted from the seeding) of channel `chan`'s
// generator comes out as `value` wherever the dither words go through a buffer (every route but the two that make them
// inside the pointwise kernel: exact mode, batches, interpreted lists, noise shaping, sloped TPDF, Gauss all do).
// n_patches < 0 clears the list.
extern "C" int icw_debug_patch_mt_word(icw_session *s, int chan, uint64_t idx, uint32_t value)
{
    if (!s) return fail(ICW_E_ARG, "NULL session");
    if (chan < 0) { s->mt_patches.clear(); return ICW_OK; }
    if (chan > 1) return fail(ICW_E_ARG, "channel 0 or 1");
    int rc = quiesce(s);
    if (rc) return rc;
    s->mt_patches.push_back({ chan, idx, value });
    return ICW_OK;
}

// parity leaf for the oscillator: omega and fmod(omega*f, 2pi) for frames n0 .. n0+n-1 of a spec
extern "C" int icw_debug_phase_device(icw_engine *e, const icw_chain_spec *spec, uint64_t n0, int64_t n,
                                      double freq_hz, double *d_out)
{
    if (!e || !spec || !d_out || n < 0) return fail(ICW_E_ARG, "bad argument");
    DevChain ch;
    HbCoef coef;
    int rc = fold_spec(*spec, ch, coef);
    if (rc) return rc;
    if (!n) return ICW_OK;
    CK(cudaSetDevice(e->device));
    CK(launch_phase_leaf(ch, n0, n, scaled_freq(fabs(freq_hz), ch.is_frmod_scaled), d_out, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return ICW_OK;
}

// ---------------------------------------------------------------------------------------------
// time-sharded streams: boundary state on the device, NCCL hand-off (icw_comm.cu)
// ---------------------------------------------------------------------------------------------
static cudaStream_t pick_stream(icw_session *s, void *cuda_stream) { return cuda_stream ? (cudaStream_t)cuda_stream : s->e->stream; }

extern "C" int icw_session_boundary_export(icw_session *s, int stream, double *d_state, int *basis_out, void *cuda_stream)
{
    if (!s || !d_state || stream < 0 || stream >= s->n_streams) return fail(ICW_E_ARG, "bad argument");
    CK(cudaSetDevice(s->e->device));
    cudaStream_t st = pick_stream(s, cuda_stream);
    if (s->last_stream && s->last_stream != st) CK(cudaStreamSynchronize(s->last_stream));
    CK(cudaMemcpyAsync(d_state, (const uint8_t *)(s->d_streams + stream) + offsetof(DevStream, hb), sizeof(double) * ICW_BOUNDARY_DOUBLES,
                       cudaMemcpyDeviceToDevice, st));
    if (basis_out) *basis_out = s->hb_basis;
    s->last_stream = st;
    return ICW_OK;
}

extern "C" int icw_session_boundary_import(icw_session *s, int stream, const double *d_state, int basis, void *cuda_stream)
{
    if (!s || !d_state || stream < 0 || stream >= s->n_streams || (basis != 0 && basis != 1)) return fail(ICW_E_ARG, "bad argument");
    if (s->hb_live && s->hb_basis != basis && s->n_streams > 1)
        return fail(ICW_E_ARG, "basis %d differs from the session's live Hilbert state basis %d", basis, s->hb_basis);
    CK(cudaSetDevice(s->e->device));
    cudaStream_t st = pick_stream(s, cuda_stream);
    if (s->last_stream && s->last_stream != st) CK(cudaStreamSynchronize(s->last_stream));
    CK(cudaMemcpyAsync((uint8_t *)(s->d_streams + stream) + offsetof(DevStream, hb), d_state, sizeof(double) * ICW_BOUNDARY_DOUBLES,
                       cudaMemcpyDeviceToDevice, st));
    s->hb_basis = basis;
    s->hb_live = true;
    s->last_stream = st;
    return ICW_OK;
}

extern "C" int icw_session_seek_closed_form(icw_session *s, int k, int64_t frames, const icw_stream_state *base)
{
    if (!s || k < 0 || k >= s->n_streams || frames < 0) return fail(ICW_E_ARG, "bad argument");
    icw_stream_state b;
    if (base) b = *base; else icw_default_state(&b);
    int rc = quiesce(s);
    if (rc) return rc;
    DevStream d;
    CK(cudaMemcpy(&d, s->d_streams + k, sizeof d, cudaMemcpyDeviceToHost));
    const uint64_t wps = (uint64_t)s->ch.render.words_per_sample;
    if (s->ch.is_frmod_scaled) d.n_frame = (b.n_frame + (uint64_t)frames % s->ch.scale_sr) % s->ch.scale_sr;
    else d.n_frame = b.n_frame + (uint64_t)frames;
    d.pos = b.pos + frames;
    for (int c = 0; c < 2; ++c) {
        d.quad[c] = (uint32_t)((b.quad[c] + (uint64_t)frames) & 3u);
        d.mt_seed[c] = b.mt_seed[c];
        d.mt_drawn[c] = b.mt_drawn[c] + (uint64_t)frames * wps;
        s->mt_seed[c][k] = d.mt_seed[c];
        s->mt_drawn[c][k] = d.mt_drawn[c];
    }
    if (s->ch.render.render_type == ICW_RENDER_STPDF && frames > 0) {
        // sloped TPDF carries the previous sample's first draw (src/sound_render.c:729): regenerate it from the
        // generator at draw (frames - 1) * wps
        for (int c = 0; c < 2; ++c) {
            rc = s->e->leaf.reserve(2 * sizeof(uint32_t));
            if (rc) return rc;
            uint64_t launches = 0;
            rc = s->e->mt.generate(d.mt_seed[c], d.mt_drawn[c] - wps, 2, (uint32_t *)s->e->leaf.p, s->e->sm_count, s->e->stream, &launches);
            if (rc) return fail(rc, "%s", s->e->mt.error());
            uint32_t w[2];
            CK(cudaMemcpyAsync(w, s->e->leaf.p, sizeof w, cudaMemcpyDeviceToHost, s->e->stream));
            CK(cudaStreamSynchronize(s->e->stream));
            const uint32_t a = w[0] >> 5, bb = w[1] >> 6;             // mtrnd_gen_dsemi / dsopen, src/mt_jrnd.c:218-256
            d.prev_rnd[c] = ((double)a * 67108864.0 + (double)bb) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
        }
    }
    CK(cudaMemcpy(s->d_streams + k, &d, sizeof d, cudaMemcpyHostToDevice));
    return ICW_OK;
}

extern "C" int icw_comm_unique_id(unsigned char id[ICW_COMM_ID_BYTES])
{
    std::string err;
    int rc = comm_unique_id(id, err);
    return rc ? fail(rc == -3 ? ICW_E_UNSUPPORTED : ICW_E_CUDA, "%s", err.c_str()) : ICW_OK;
}
extern "C" int icw_comm_init(const unsigned char id[ICW_COMM_ID_BYTES], int rank, int world, void **comm_out)
{
    if (!comm_out || world < 1 || rank < 0 || rank >= world) return fail(ICW_E_ARG, "bad argument");
    std::string err;
    int rc = comm_init(id, rank, world, comm_out, err);
    return rc ? fail(rc == -3 ? ICW_E_UNSUPPORTED : ICW_E_CUDA, "%s", err.c_str()) : ICW_OK;
}
extern "C" int icw_comm_destroy(void *comm)
{
    std::string err;
    int rc = comm_destroy(comm, err);
    return rc ? fail(rc == -3 ? ICW_E_UNSUPPORTED : ICW_E_CUDA, "%s", err.c_str()) : ICW_OK;
}
extern "C" int icw_nccl_version(void) { return nccl_version(); }

extern "C" int icw_session_handoff(icw_session *from, icw_session *to, void *comm, int rank, int world, void *cuda_stream)
{
    if (!comm || world < 1 || rank < 0 || rank >= world) return fail(ICW_E_ARG, "bad argument");
    if (rank + 1 < world && !from) return fail(ICW_E_ARG, "rank %d of %d has a right neighbour but no state to send", rank, world);
    if (rank > 0 && !to) return fail(ICW_E_ARG, "rank %d has a left neighbour but no session to receive into", rank);
    icw_session *any = from ? from : to;
    if (!any) return ICW_OK;                                // world == 1
    if (from && from->hb_live && from->hb_basis != ICW_HILBERT_SCAN) {
        int rc = convert_basis(from, ICW_HILBERT_SCAN);     // a shard may have been run in exact mode: send modal states
        if (rc) return rc;
    }
    CK(cudaSetDevice(any->e->device));
    cudaStream_t st = pick_stream(any, cuda_stream);
    for (icw_session *x : { from, to })
        if (x && x->last_stream && x->last_stream != st) CK(cudaStreamSynchronize(x->last_stream));
    const double *snd = (from && rank + 1 < world) ? (const double *)((const uint8_t *)from->d_streams + offsetof(DevStream, hb)) : nullptr;
    double *rcv = (to && rank > 0) ? (double *)((uint8_t *)to->d_streams + offsetof(DevStream, hb)) : nullptr;
    std::string err;
    int rc = comm_shift_right(comm, rank, world, snd, rcv, ICW_BOUNDARY_DOUBLES, st, err);
    if (rc) return fail(rc == -3 ? ICW_E_UNSUPPORTED : ICW_E_CUDA, "%s", err.c_str());
    if (rcv) { to->hb_basis = ICW_HILBERT_SCAN; to->hb_live = true; to->last_stream = st; }
    if (from) from->last_stream = st;
    return ICW_OK;
}

extern "C" int icw_session_reduce_counters(icw_session *s, void *comm, void *cuda_stream, uint64_t clips_out[2], double peak_db_out[2])
{
    if (!s || !comm || !clips_out || !peak_db_out) return fail(ICW_E_ARG, "bad argument");
    icw_stats mine;
    int rc = icw_session_stats(s, &mine);                   // sums / maxima over this rank's streams (synchronises)
    if (rc) return rc;
    std::vector<DevStream> all((size_t)s->n_streams);
    CK(cudaMemcpy(all.data(), s->d_streams, sizeof(DevStream) * all.size(), cudaMemcpyDeviceToHost));
    struct { unsigned long long clips[2]; double peak[2]; } h;
    h.clips[0] = mine.clips[0]; h.clips[1] = mine.clips[1];
    h.peak[0] = h.peak[1] = 0.0;
    for (const auto &d : all) { h.peak[0] = std::max(h.peak[0], d.peak[0]); h.peak[1] = std::max(h.peak[1], d.peak[1]); }
    rc = s->e->leaf.reserve(sizeof h);
    if (rc) return rc;
    cudaStream_t st = pick_stream(s, cuda_stream);
    CK(cudaMemcpyAsync(s->e->leaf.p, &h, sizeof h, cudaMemcpyHostToDevice, st));
    std::string err;
    rc = comm_reduce(comm, (unsigned long long *)s->e->leaf.p, 2, (double *)((uint8_t *)s->e->leaf.p + 16), 2, st, err);
    if (rc) return fail(rc == -3 ? ICW_E_UNSUPPORTED : ICW_E_CUDA, "%s", err.c_str());
    CK(cudaMemcpyAsync(&h, s->e->leaf.p, sizeof h, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    clips_out[0] = h.clips[0]; clips_out[1] = h.clips[1];
    peak_db_out[0] = icw_peak_db(h.peak[0]); peak_db_out[1] = icw_peak_db(h.peak[1]);
    return ICW_OK;
}
