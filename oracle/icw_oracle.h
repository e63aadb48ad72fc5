/* icw_oracle.h -- CPU restatement of the in_cwave signal chain (the "port" oracle).
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.  The product
 * (in_cwave_b200/) never links, imports or calls anything in oracle/.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py checks every function here against
 * the reference's own sources compiled in place (oracle/_ref/libicw_ref.so, see Makefile)
 * and against tests/golden/ (vectors produced by that same compiled reference, plus the
 * reference's one known-answer test, mt19937ar_out.c).
 *
 * Units follow the reference: one "frame" = one L+R sample pair; sample values are doubles
 * in +-32768 units (reference src/sound_render.c:36-39).
 */
#ifndef ICW_ORACLE_H
#define ICW_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ICWO_N_PLUGS    27      /* In + A..Z          (reference src/in_cwave.h:198) */
#define ICWO_MAX_NODES  32
#define ICWO_MAX_ORD    20
#define ICWO_MT_N       624

/* input encodings: WAV (real) 0..4 = reference HRW_FMT_*, CWAVE (complex I/Q) 16+HCW_FMT_* */
enum {
    ICWO_FMT_WAV_U8 = 0, ICWO_FMT_WAV_I16 = 1, ICWO_FMT_WAV_I24 = 2, ICWO_FMT_WAV_I32 = 3,
    ICWO_FMT_WAV_F32 = 4,
    ICWO_FMT_CW_F64 = 16, ICWO_FMT_CW_I16 = 17, ICWO_FMT_CW_I16F32 = 18, ICWO_FMT_CW_F32 = 19
};
enum { ICWO_MODE_MASTER = 0, ICWO_MODE_SHIFT = 1, ICWO_MODE_PM = 2, ICWO_MODE_MIX = 3 };
enum { ICWO_XCH_NORMAL = 0, ICWO_XCH_SWAP = 1, ICWO_XCH_LEFT = 2, ICWO_XCH_RIGHT = 3, ICWO_XCH_MIXLR = 4 };
enum { ICWO_OUT_ADD = 0, ICWO_OUT_SUB = 1, ICWO_OUT_RE = 2, ICWO_OUT_IM = 3 };
enum { ICWO_DITHER_NONE = 0, ICWO_DITHER_RPDF = 1, ICWO_DITHER_TPDF = 2, ICWO_DITHER_STPDF = 3,
       ICWO_DITHER_GAUSS = 4 };

/* one DSP-list node, EXECUTION order (the reference walks its list tail -> head, so the
 * master is last).  Same field meaning as NODE_DSP, reference src/in_cwave.h:207-287. */
typedef struct icwo_node {
    int      mode;
    unsigned inputs_mask;       /* bit k: plug k is summed into the node input */
    int      xch_mode;
    int      l_iq_invert, r_iq_invert;
    double   l_gain, r_gain;
    int      n_out;             /* output plug 1..26 (not for master) */
    int      l_tout, r_tout;    /* master: ICWO_OUT_* */
    int      l_on, r_on;        /* shift: is_shift, pm: is_pm */
    double   l_p[4], r_p[4];    /* shift {fr_shift}; pm {freq, phase, level, angle} */
} icwo_node;

typedef struct icwo_spec {
    int      fmt;
    int      n_channels;        /* 1 or 2 */
    unsigned sample_rate;
    int64_t  n_samples;         /* frames in the file (fade-out anchor) */
    int64_t  n_fade_in, n_fade_out;   /* frames, 0 = off */
    int      filter_no;         /* 0..5 */
    int      is_kahan;
    int      is_subnorm_reject;
    int      is_frmod_scaled;
    int      need24bits;
    double   dth_bits;
    unsigned quantz_type;       /* 0 mid tread, 1 mid riser */
    unsigned render_type;       /* ICWO_DITHER_* */
    unsigned nshape_type;       /* 0 flat, 1..15 FIR shapers, 16..17 IIR shapers (src/sound_render.h:73-92) */
    unsigned sign_bits16, sign_bits24;
    int      bypass;
    int      n_nodes;
    icwo_node nodes[ICWO_MAX_NODES];
    int      is_fp_check;       /* FP_CHECK: the FP-exception-checked twins (src/fp_check.c:52-99) */
} icwo_spec;

typedef struct icwo_iir {
    double   z[ICWO_MAX_ORD];   /* circular delay line */
    int      ix;
    uint64_t rejects;
} icwo_iir;

typedef struct icwo_mt {
    uint32_t w[ICWO_MT_N];
    int      pos;               /* next word to hand out; ICWO_MT_N = regenerate first */
    uint64_t drawn;             /* words consumed so far */
} icwo_mt;
/* test hook (tests/ only): replace word `idx` of a generator's stream, if it comes out as `match`, by `value` */
void icwo_debug_patch_word(uint64_t idx, uint32_t match, uint32_t value);
void icwo_debug_clear_patches(void);

/* everything that persists from frame to frame (reference MOD_CONTEXT, src/in_cwave.h:410-424) */
/* noise-shaper memory of one channel (reference NS_SHAPER, src/sound_render.h:113-139), ordered by AGE:
 * e[0] = the latest quantisation error, o[0] = the filter's latest output (IIR shapers).  The
 * reference keeps circular buffers + an index; the sums run newest-first either way. */
#define ICWO_NS_MAX_TAPS 20
typedef struct icwo_ns {
    double e[ICWO_NS_MAX_TAPS], o[ICWO_NS_MAX_TAPS];
    double prev_err;                    /* what the next sample subtracts (src/sound_render.c:756) */
} icwo_ns;

typedef struct icwo_state {
    uint64_t n_frame;                   /* oscillator frame counter */
    int64_t  pos;                       /* frames already taken from the current file */
    double   bus[ICWO_N_PLUGS][4];      /* (L.re, L.im, R.re, R.im) */
    icwo_iir lpf[2][2];                 /* [channel][0 = I, 1 = Q] */
    unsigned quad[2];                   /* n mod 4 per channel */
    icwo_mt  mt[2];
    double   prev_rnd[2];
    unsigned clips[2];
    double   peak_db[2];
    icwo_ns  ns[2];
    /* FP_EXCEPT_STATS of the context (src/in_cwave.h:418-421): [hilbert L, hilbert R, render L, render R]
     * x [total, snan, qnan, ninf, nden, pden, pinf] (src/fp_check.h:66-76) */
    unsigned fp_cnt[4][7];
} icwo_state;

void    icwo_default_spec(icwo_spec *sp);
/* fresh stream: zero filters/bus/counters, MT seeded like the plugin (src/in_cwave.c:69-70) */
void    icwo_state_init(icwo_state *st);
int     icwo_frame_bytes(const icwo_spec *sp);      /* input bytes per frame, <0 if bad fmt */
int     icwo_out_frame_bytes(const icwo_spec *sp);  /* 4 or 6 */

/* Run n frames.  Optional taps (NULL to skip): analytic [n][4] = plug 0 after unpack/Hilbert;
 * bus_tap [n][n_tap][4] for the listed plugs after the graph ran; lr_tap [n][2] = master out.
 * Returns 0, or <0 for a spec the oracle does not model. */
int     icwo_process(const icwo_spec *sp, icwo_state *st, const uint8_t *in, int64_t n,
                     uint8_t *pcm, double *analytic, double *bus_tap, const int *tap_plugs,
                     int n_tap, double *lr_tap);

/* leaf entry points for stage-wise parity */
void    icwo_mt_seed(icwo_mt *mt, uint32_t seed);
void    icwo_mt_seed_key(icwo_mt *mt, const uint32_t *key, uint32_t key_len);
uint32_t icwo_mt_u32(icwo_mt *mt);
double  icwo_mt_dsopen(icwo_mt *mt);
void    icwo_unpack(int fmt, int n_channels, const uint8_t *in, int64_t n, double *out4);
void    icwo_hilbert(int filter_no, int is_kahan, int is_reject, icwo_iir lpf[2], unsigned *quad,
                     const double *x, int64_t n, double *out_i, double *out_q);
void    icwo_iir_run(int filter_no, int is_kahan, int is_reject, icwo_iir *f,
                     const double *x, int64_t n, double *y);
/* exact-arithmetic (binary128) value of the same converter from zero state; see icw_oracle.c */
void    icwo_hilbert_truth(int filter_no, int drop_direct, unsigned quad0, const double *x, int64_t n,
                           double *out_i, double *out_q);
/* ns may be NULL when sp->nshape_type == 0 */
int64_t icwo_render(const icwo_spec *sp, icwo_mt *mt, double *prev_rnd, icwo_ns *ns, const double *in, int64_t n,
                    uint8_t *out, unsigned *clips, double *peak_db);

#ifdef __cplusplus
}
#endif
#endif
