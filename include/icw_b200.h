/* icw_b200.h -- C ABI of the B200 (sm_100a) implementation of the in_cwave signal chain
 *
 *      file bytes --unpack--> real/complex double --Hilbert (real input only)--> analytic
 *                 --modulator graph (DSP list)--> (L, R) double --render--> 16/24-bit LE PCM
 *
 * Plain C: opaque handles, POD structs, raw pointers and sizes.  No CUDA or torch types cross
 * this boundary (a CUDA stream is passed as void*).  Every entry point names the reference
 * interface it stands in for (paths relative to the reference tree, file:line).
 *
 * Vocabulary (the reference's): a FRAME is one L+R sample pair; sample values are doubles in
 * +-32768 units; a PLUG is one of the 27 cells of the in/out bus (0 = In, 1..26 = A..Z); the
 * DSP LIST is the modulator graph; a STREAM is one file played through one fresh MOD_CONTEXT.
 *
 * Threading: one icw_engine per process and GPU; calls on one session must not overlap.
 * Streams: icw_session_process_device is asynchronous on the caller's stream.  Every state accessor
 * (get/set_state, reset, set_spec, stats, fp_stats, sync, destroy) first waits for the engine's own stream
 * AND for the stream of the session's last process call; scratch shared between sessions of one engine is
 * ordered across streams by an event, so sessions may use different streams.
 * Errors: every int-returning call gives ICW_OK (0) or a negative ICW_E_* code and leaves a
 * message in icw_last_error().  There is NO CPU fallback anywhere behind this header.
 */
#ifndef ICW_B200_H
#define ICW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ICW_ABI_VERSION   1
#define ICW_N_PLUGS       27    /* reference src/in_cwave.h:198  N_INPUTS */
#define ICW_MAX_NODES     32
#define ICW_NS_MAX_TAPS   20      /* longest noise-shaping filter (src/sound_render.c:75-235) */
#define ICW_MAX_ORD       20    /* highest half-band filter order, reference src/hblpf.c:740-820 */

enum {
    ICW_OK = 0,
    ICW_E_ARG = -1,             /* malformed argument / spec */
    ICW_E_CUDA = -2,            /* CUDA runtime failure (message has the detail) */
    ICW_E_UNSUPPORTED = -3,     /* valid in the reference but not modelled on the GPU (says which) */
    ICW_E_NOMEM = -4,
    ICW_E_MT_REDRAW = -5        /* a dither draw hit the reference's rejection loop (src/mersene_twister/mt_jrnd.c:249-253:
                                 * probability 2^-53 a draw).  The HOST entry point replays such a call the reference's way
                                 * (icw_session_process_host) and never returns this; the DEVICE entry point cannot (the
                                 * input may be gone): the frames before the event are the reference's, the later ones drew
                                 * their words two places early, and the next icw_session_sync says so -- once;
                                 * icw_stats.mt_redraws counts both kinds */
};

/* input sample encodings.  0..4 = reference HRW_FMT_* (src/in_cwave.h:326-330, unpackers
 * src/xwave_reader.c:203-239); 16+n = reference HCW_FMT_* n (src/cwave.h:76-84, unpackers
 * src/xwave_reader.c:171-199). */
enum {
    ICW_FMT_WAV_U8 = 0, ICW_FMT_WAV_I16 = 1, ICW_FMT_WAV_I24 = 2, ICW_FMT_WAV_I32 = 3,
    ICW_FMT_WAV_F32 = 4,
    ICW_FMT_CW_F64 = 16, ICW_FMT_CW_I16 = 17, ICW_FMT_CW_I16F32 = 18, ICW_FMT_CW_F32 = 19
};
/* node kinds, reference src/in_cwave.h:289-292 */
enum { ICW_MODE_MASTER = 0, ICW_MODE_SHIFT = 1, ICW_MODE_PM = 2, ICW_MODE_MIX = 3 };
/* channel exchange, reference src/in_cwave.h:186-190 */
enum { ICW_XCH_NORMAL = 0, ICW_XCH_SWAP = 1, ICW_XCH_LEFTONLY = 2, ICW_XCH_RIGHTONLY = 3, ICW_XCH_MIXLR = 4 };
/* master output, reference src/in_cwave.h:154-157 */
enum { ICW_OUT_ADD_REIM = 0, ICW_OUT_SUB_REIM = 1, ICW_OUT_RE = 2, ICW_OUT_IM = 3 };
/* dither, reference src/sound_render.h:62-66 */
enum { ICW_RENDER_ROUND = 0, ICW_RENDER_RPDF = 1, ICW_RENDER_TPDF = 2, ICW_RENDER_STPDF = 3, ICW_RENDER_GAUSS = 4 };
/* how the half-band recurrences are evaluated */
enum {
    ICW_HILBERT_EXACT = 0,      /* the reference's own operation order; bit-exact; serial per stream */
    ICW_HILBERT_SCAN = 1        /* modal block scan; time-parallel; differs from the reference by
                                   the reference's own rounding noise (DESIGN.md "numerics").  A live stream may
                                   switch modes: its filter memory is converted between the bases (icw_host_hb_convert) */
};

/* One DSP-list node in EXECUTION order (the reference walks its list tail -> head, so the
 * master comes last).  Field meaning = NODE_DSP, reference src/in_cwave.h:207-287; the
 * per-frame arithmetic is reference src/adv_modulator.c:485-583 and :637-751. */
typedef struct icw_node {
    int32_t  mode;              /* ICW_MODE_* */
    uint32_t inputs_mask;       /* bit k set: plug k is summed into this node's input */
    int32_t  xch_mode;          /* ICW_XCH_* */
    int32_t  l_iq_invert, r_iq_invert;
    double   l_gain, r_gain;
    int32_t  n_out;             /* output plug 1..26 (ignored for the master) */
    int32_t  l_tout, r_tout;    /* master: ICW_OUT_* */
    int32_t  l_on, r_on;        /* shift: is_shift; pm: is_pm (0 = copy through) */
    double   l_p[4], r_p[4];    /* shift: {fr_shift Hz, signed}; pm: {freq, phase, level, angle} */
} icw_node;

/* Everything the reference keeps in its config (src/config.c:118-207), reader
 * (src/in_cwave.h:375-405) and DSP list that changes what one frame computes. */
typedef struct icw_chain_spec {
    int32_t  fmt;               /* ICW_FMT_* */
    int32_t  n_channels;        /* 1 or 2 (mono is duplicated to R, src/xwave_reader.c:949,988) */
    uint32_t sample_rate;       /* Hz, <= 2 000 000 (src/in_cwave.h:160) */
    int64_t  n_samples;         /* frames in the file: fade-out anchor (src/xwave_reader.c:930) */
    int64_t  n_fade_in, n_fade_out;     /* frames; 0 = off (src/xwave_reader.c:921-936) */
    int32_t  filter_no;         /* half-band design 0..5 (src/lpf_hilbert_quad.h:68-84) */
    int32_t  is_kahan;          /* src/hblpf.c:1008 vs :894 */
    int32_t  is_subnorm_reject; /* compared as a number, bug-for-bug (src/hblpf.c:915,1046) */
    int32_t  hilbert_mode;      /* ICW_HILBERT_* */
    int32_t  is_frmod_scaled;   /* mHz frequency grid + wrapping frame counter (src/adv_modulator.c:612-624) */
    int32_t  need24bits;
    double   dth_bits;
    uint32_t quantz_type;       /* 0 mid tread, 1 mid riser (src/sound_render.h:57-58) */
    uint32_t render_type;       /* ICW_RENDER_* */
    uint32_t nshape_type;       /* 0 = flat, 1..15 FIR, 16..17 IIR shapers (src/sound_render.h:73-92): the error
                                   feedback is serial per channel -> one thread per (stream, channel) */
    uint32_t sign_bits16, sign_bits24;
    int32_t  bypass;            /* src/adv_modulator.c:637,644 */
    int32_t  n_nodes;
    icw_node nodes[ICW_MAX_NODES];
    int32_t  is_fp_check;       /* FP_CHECK (src/config.c:180, off by default): the FP-exception-checked twins of the
                                   half-band filters, the renderer and the noise shapers (src/fp_check.c:52-99,
                                   src/hblpf.c:928-950,1059-1096, src/sound_render.c:811-913): NaN and denormal
                                   intermediates become 0, +-Inf becomes +-65535, every event counted.  Exact Hilbert
                                   mode only (scan mode refuses it) */
    int32_t  reserved_;
} icw_chain_spec;

/* What persists from frame to frame for one stream: the reference's MOD_CONTEXT
 * (src/in_cwave.h:410-424) reduced to its arithmetic content. */
typedef struct icw_stream_state {
    uint64_t n_frame;                       /* oscillator frame counter */
    int64_t  pos;                           /* frames already taken from the current file */
    double   hb[2][2][ICW_MAX_ORD];         /* [channel][0 = I, 1 = Q][j] = filter state j+1 frames ago */
    uint64_t hb_rejects[2][2];              /* "subnorm reject" hits (src/hblpf.c:918) */
    uint32_t quad[2];                       /* frame index mod 4 of the fs/4 mixer, per channel */
    uint32_t mt_seed[2];                    /* MT19937 identity: mtrnd_init_seed(seed) ... */
    uint64_t mt_drawn[2];                   /* ... advanced by this many 32-bit words */
    double   prev_rnd[2];                   /* sloped-TPDF memory (src/sound_render.c:729) */
    uint32_t clips[2];                      /* src/sound_render.c:782-797 */
    double   peak[2];                       /* max |q|/hi_bound, LINEAR; dB via icw_peak_db() */
    double   bus[ICW_N_PLUGS][4];           /* (L.re, L.im, R.re, R.im) per plug */
    uint32_t hb_basis;                      /* what hb[] holds: 0 = delay line (exact mode), 1 = modal states
                                               (scan mode: hb[c][f][2m], [2m+1] = Re, Im of mode m) */
    uint32_t reserved;
    /* noise-shaper memory per channel (reference NS_SHAPER, src/sound_render.h:113-139) ordered by AGE:
     * ns_e[c][0] = latest quantisation error, ns_o[c][0] = the IIR shapers' latest output; ns_prev_err[c] =
     * what the next sample subtracts (src/sound_render.c:756,800).  All zero for FLAT shaping. */
    double   ns_e[2][ICW_NS_MAX_TAPS], ns_o[2][ICW_NS_MAX_TAPS];
    double   ns_prev_err[2];
} icw_stream_state;

typedef struct icw_engine  icw_engine;      /* one GPU: streams, scratch, MT jump tables */
typedef struct icw_session icw_session;     /* a chain spec + K device-resident stream states */

const char *icw_last_error(void);
int  icw_abi_version(void);

/* spec with the reference's defaults: src/config.c:118-207, master node src/adv_modulator.c:112-118 */
void icw_default_spec(icw_chain_spec *spec);
/* fresh stream == state right after winampGetInModule2(): src/in_cwave.c:46-80 (seeds :69-70) */
void icw_default_state(icw_stream_state *st);
int  icw_frame_bytes(const icw_chain_spec *spec);       /* input bytes per frame */
int  icw_out_frame_bytes(const icw_chain_spec *spec);   /* 4 or 6; src/sound_render.c:682-685 */
/* 20*log10(peak) or -555 dB for silence; src/sound_render.c:773-775 */
double icw_peak_db(double peak_linear);

int  icw_engine_create(int device, icw_engine **out);
void icw_engine_destroy(icw_engine *e);
/* page-locked host memory (cudaHostAlloc) for callers that are plain C: buffers handed to
 * icw_session_process_host from it move by asynchronous DMA and overlap the kernels; stands where the
 * reference's reader mallocs its read buffer (src/xwave_reader.c:822-835) */
int  icw_pinned_alloc(size_t bytes, void **out);
void icw_pinned_free(void *p);

/* K independent streams sharing one spec; stands in for K fresh MOD_CONTEXTs
 * (src/in_cwave.c:46-80) plus amod_init (src/adv_modulator.c:216-331). */
int  icw_session_create(icw_engine *e, const icw_chain_spec *spec, int n_streams, icw_session **out);
void icw_session_destroy(icw_session *s);
/* parameter snapshot: what the GUI thread does through adbl_write / srenders_set_vcfg /
 * mod_context_change_all_hilberts_* (src/amod_gui_control.c:259-312,1555-1602); takes effect at
 * the next process call.  A new filter_no clears the Hilbert state like the reference does
 * (src/in_cwave.c:135-150); a render change clears prev_rnd (src/sound_render.c:499-581). */
int  icw_session_set_spec(icw_session *s, const icw_chain_spec *spec);
int  icw_session_get_state(icw_session *s, int stream, icw_stream_state *out);
int  icw_session_set_state(icw_session *s, int stream, const icw_stream_state *in);
/* mod_context_reset_hilbert / _reset_framecnt (src/in_cwave.c:161,287), counters
 * (amod_get_clips_peaks isReset, src/adv_modulator.c:445), new file position */
enum { ICW_RESET_HILBERT = 1, ICW_RESET_FRAMECNT = 2, ICW_RESET_COUNTERS = 4, ICW_RESET_FILEPOS = 8,
       ICW_RESET_RENDER = 16,   /* dither generators back at their seeds, prev_rnd, shaper memory and bus cleared: with the
                                 * other four, the state of a fresh context (winampGetInModule2, src/in_cwave.c:551-572) */
       ICW_RESET_RENDER_MEMORY = 32,    /* what sound_render_recalc clears (src/sound_render.c:509,556-580): prev_rnd, the shaper's
                                         * buffers and prev_ns_err -- run by EVERY mod_context_fopen through sound_render_set_outbits
                                         * (src/in_cwave.c:231-234); the generators keep their place */
       ICW_RESET_ALL = 255 };
int  icw_session_reset(icw_session *s, unsigned what);

/* The hot call: n frames of every stream.  Stands in for the frame loop of
 * amod_process_samples (src/adv_modulator.c:587-763) run n_streams times.
 * Input stream k starts at in + k*in_stride bytes, output at out + k*out_stride bytes.
 * _host: pageable or pinned host memory, copies included.  _device: device pointers,
 * asynchronous on cuda_stream (a cudaStream_t, NULL = engine's own stream). */
int  icw_session_process_host(icw_session *s, int64_t n_frames, const void *in, size_t in_stride,
                              void *out, size_t out_stride);
int  icw_session_process_device(icw_session *s, int64_t n_frames, const void *d_in, size_t in_stride,
                                void *d_out, size_t out_stride, void *cuda_stream);
int  icw_session_sync(icw_session *s);

/* sums / maxima over the session's streams: amod_get_clips_peaks (src/adv_modulator.c:445-465),
 * mod_context_get_desubnorm_counter (src/in_cwave.c:296) */
typedef struct icw_stats {
    uint64_t clips[2];
    double   peak_db[2];
    uint64_t hb_rejects;
    uint64_t mt_redraws;        /* dsopen re-draws seen (src/mt_jrnd.c:249-253); each reported once as ICW_E_MT_REDRAW */
    uint64_t kernel_launches;   /* our kernels launched by this session so far */
} icw_stats;
int  icw_session_stats(icw_session *s, icw_stats *out);
/* FP exception counters of one stream (reference fecs_getcnts, src/in_cwave.c:383-440; FP_EXCEPT_STATS,
 * src/fp_check.h:66-76): out[block][class], block = hilbert L, hilbert R, render L, render R; class = total,
 * snan, qnan, ninf, nden, pden, pinf.  Arithmetic never yields a signalling NaN: snan stays 0. */
int  icw_session_fp_stats(icw_session *s, int stream, uint32_t out[4][7]);

/* leaf entry points (device buffers in, device buffers out) for stage-wise parity:
 * hq_rp_process (src/lpf_hilbert_quad.c:129-156) over n samples of n_chan independent channels,
 * x[c*n + i] -> out_iq[(c*n + i)*2 + {0,1}]; state per channel in/out on the host. */
int  icw_hilbert_device(icw_engine *e, int filter_no, int is_kahan, int is_reject, int mode,
                        int n_chan, int64_t n, const double *d_x, double *d_out_iq,
                        icw_stream_state *chan_state /* uses hb[0], quad[0], hb_rejects[0] */);
/* mtrnd_gen_ui32 stream (src/mt_jrnd.c:99-134): words [skip, skip+n) after mtrnd_init_seed(seed),
 * produced on the GPU through the jump-ahead path */
int  icw_mt_words_device(icw_engine *e, uint32_t seed, uint64_t skip, int64_t n, uint32_t *d_out);

/* CRC-32 of a device buffer, the reference's CWAVE data check (src/crc32.c:55-108 as driven by
 * src/gui_cwave.c:82-130; == CRC-32/ISO-HDLC).  Synchronous.  icw_crc32_combine joins the CRCs of two
 * consecutive pieces (crc of A, crc of B, length of B) so a file can be checked block by block. */
int  icw_crc32_device(icw_engine *e, const void *d_data, size_t n_bytes, uint32_t *crc_out);
int  icw_crc32_host(icw_engine *e, const void *data, size_t n_bytes, uint32_t *crc_out);   /* host memory, staged in blocks */
uint32_t icw_crc32_combine(uint32_t crc_a, uint32_t crc_b, uint64_t len_b);

/* ---- time-sharded streams: the block-boundary hand-off ---------------------------------------
 * One long stream cut in time over several GPUs: everything a frame needs except the Hilbert filter memory is a
 * closed form of its index, so rank r only has to receive the four half-band filters' state at its first frame.
 * That state is the reference's IIR_RAT_POLY delay line x 4 plus the mixer phase (src/hblpf.h:115-127,
 * src/lpf_hilbert_quad.h:52-57); here it is hb[2][2][ICW_MAX_ORD] of the stream, in the session's current basis,
 * and it never leaves device memory: */
#define ICW_BOUNDARY_DOUBLES (2 * 2 * ICW_MAX_ORD)
/* stream's hb[] -> d_state[ICW_BOUNDARY_DOUBLES] (device memory), asynchronous on cuda_stream; *basis_out (optional)
 * receives the basis it is in (0 delay line, 1 modal) */
int  icw_session_boundary_export(icw_session *s, int stream, double *d_state, int *basis_out, void *cuda_stream);
/* d_state (device memory, `basis` as exported) -> stream's hb[]; asynchronous on cuda_stream */
int  icw_session_boundary_import(icw_session *s, int stream, const double *d_state, int basis, void *cuda_stream);
/* closed-form part of a segment start: what frame counter, file position, mixer phase and dither offset a stream
 * that began fresh (or at `base`, if not NULL) holds `frames` frames later (src/adv_modulator.c:611-625,
 * src/lpf_hilbert_quad.c:155, src/sound_render.c:711-751); the Hilbert memory, counters and bus are left alone */
int  icw_session_seek_closed_form(icw_session *s, int stream, int64_t frames, const icw_stream_state *base);
/* a communicator of our own: rank 0 makes an id, the caller distributes its ICW_COMM_ID_BYTES bytes, every rank inits.
 * NCCL is loaded at run time (the copy already in the process, e.g. torch's); ICW_E_UNSUPPORTED if there is none. */
#define ICW_COMM_ID_BYTES 128
int  icw_comm_unique_id(unsigned char id[ICW_COMM_ID_BYTES]);
int  icw_comm_init(const unsigned char id[ICW_COMM_ID_BYTES], int rank, int world, void **comm_out);
int  icw_comm_destroy(void *comm);
int  icw_nccl_version(void);                    /* 0 when NCCL cannot be loaded */
/* THE collective of the path: `from` (stream 0's filter state; NULL on the last rank) goes to rank + 1 and rank - 1's
 * arrives in `to` (NULL on rank 0), device to device over NVLink by ncclSend / ncclRecv on cuda_stream; `to` is then
 * in the modal basis.  comm: an ncclComm_t (icw_comm_init or the caller's own).  Both sessions must be scan-mode. */
int  icw_session_handoff(icw_session *from, icw_session *to, void *comm, int rank, int world, void *cuda_stream);
/* clip counters summed and peaks maximised over the shards (the reference keeps ONE pair of accumulators,
 * src/adv_modulator.c:54-55): ncclAllReduce on cuda_stream, results to the host (synchronises) */
int  icw_session_reduce_counters(icw_session *s, void *comm, void *cuda_stream, uint64_t clips_out[2], double peak_db_out[2]);
/* one filter's state between the two bases, on the host in binary128 (icw_hbconv.cpp): to_basis 1 = delay line ->
 * modal, 0 = modal -> delay line.  A live stream whose hilbert_mode changes is converted with this automatically. */
int  icw_host_hb_convert(int filter_no, int to_basis, const double *in, double *out);

/* ---- measurement: per-kernel device time from CUDA events on the launching stream ---------- */
/* HILBERT: the exact recurrences (fused with the chain unless ICW_UNFUSED); SCAN_LOCAL / SCAN_APPLY: passes 1+2 and
 * pass 3 of the time-parallel converter; CHAIN: the pointwise kernel; MT: dither word generation incl. jump-ahead */
enum { ICW_K_HILBERT = 0, ICW_K_CHAIN = 1, ICW_K_MT = 2, ICW_K_MISC = 3, ICW_K_SCAN_LOCAL = 4, ICW_K_SCAN_APPLY = 5,
       ICW_K_SCAN_FUSED = 6 /* scan mode, whole chain in one kernel (icw_sfused.cu) */, ICW_K_COUNT = 7 };
typedef struct icw_profile {
    double   ms[ICW_K_COUNT];       /* summed device time per kernel class since the last reset */
    uint64_t launches[ICW_K_COUNT];
} icw_profile;
/* on != 0: bracket every kernel class of this session with events (a few us each); reading syncs */
int  icw_session_profile(icw_session *s, int on);
int  icw_session_profile_read(icw_session *s, icw_profile *out, int reset);
const char *icw_kernel_class_name(int k);

/* ---- diagnostics used by the parity tests (not part of the reference's surface) -------------- */
/* device buffers that receive, per stream and frame, the whole bus [ICW_N_PLUGS][4] and the
 * master output [2] of the next process calls (NULL = off) */
int  icw_session_set_taps(icw_session *s, double *d_tap_bus, double *d_tap_lr);
/* oscillator leaf: out[i] = { norm_omega, fmod(norm_omega * f, 2*pi) } for frame counter n0+i,
 * f = |freq_hz| on the spec's frequency grid (reference src/adv_modulator.c:36-38,537,611-625) */
int  icw_debug_phase_device(icw_engine *e, const icw_chain_spec *spec, uint64_t n0, int64_t n,
                            double freq_hz, double *d_out);
/* trig leaf: out[i] = { s, c, sin(x[i]), cos(x[i]) } with (s, c) from the modulator's own sincos for
 * phases in [0, 2*pi) (icw_dev.cuh sincos_2pi) and the other two from CUDA's libdevice */
int  icw_debug_sincos_device(icw_engine *e, int64_t n, const double *d_x, double *d_out);
/* test hook: count n more dsopen re-draws on stream k, as a kernel that met one would (see ICW_E_MT_REDRAW) */
int  icw_debug_note_redraw(icw_session *s, int k, uint64_t n);
/* test hook: word `idx` (counted from the seeding) of channel chan's generator is handed out as `value` wherever the dither
 * words pass through a buffer; chan < 0 clears the patches.  Two patched words (0, 0) make a pair the reference rejects. */
int  icw_debug_patch_mt_word(icw_session *s, int chan, uint64_t idx, uint32_t value);
/* host-only checks of the MT19937 jump-ahead mathematics (no GPU is touched):
 * characteristic polynomial found by Berlekamp-Massey; the state array after `blocks` block
 * regenerations computed sequentially, through x^J mod phi, and through the x^(624*2^k) family */
int  icw_mt_host_charpoly(int *n_terms, int *degree);
void icw_mt_host_seq_state(uint32_t seed, uint64_t blocks, uint32_t *out624);
int  icw_mt_host_jump_state(uint32_t seed, uint64_t blocks, uint32_t *out624);
int  icw_mt_host_jump_state_family(uint32_t seed, uint64_t blocks, uint32_t *out624);
/* the same distance as ONE polynomial, the product of the x^(624*2^k) family over the set bits of the
 * distance -- how the checkpoint tree reaches units of any length (MtJump::poly_for) */
int  icw_mt_host_jump_state_product(uint32_t seed, uint64_t blocks, uint32_t *out624);
/* planning arithmetic, host only: blocks per jump-ahead unit for nb blocks over at most max_units CTAs (m * 2^k,
 * m < 16); frames per scan chunk for a launch group (256, 1024 or 2048; ICW_SCAN_L overrides) */
uint64_t icw_mt_host_unit_blocks(uint64_t nb, int max_units);
int  icw_host_scan_chunk_len(int n_streams, int64_t n_frames, int sm_count);

#ifdef __cplusplus
}
#endif
#endif /* ICW_B200_H */
