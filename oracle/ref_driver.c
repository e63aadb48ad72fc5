/* ref_driver.c -- flat C entry points over the UNMODIFIED reference DSP sources.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is compiled together with the reference's own
 * C files (read in place from /root/reference/src, never copied) into
 * oracle/_ref/libicw_ref.so by oracle/Makefile.  It supplies the four symbols that live
 * in reference files we do not compile as they stand (playback.c: GUI code; config.c's default-
 * path lookup -- config.c itself IS compiled, with its two *_default functions renamed away, so the
 * reference's own load_config()/save_config() text format is available to the tests) and
 * a ctypes-friendly facade so tests/ and bench.py's cpu_baseline leg can run the reference
 * chain on a file and read back PCM, per-frame bus taps and the counters.
 *
 * Reference entry points driven here:
 *   winampGetInModule2            src/in_cwave.c:551
 *   amod_add_lastdsp/get_headdsp  src/adv_modulator.c:394,429
 *   mod_context_fopen/fclose      src/in_cwave.c:207,240
 *   amod_process_samples          src/adv_modulator.c:587
 *   winampGetExtendedRead_*       src/transcode.c:40-118
 *   hq_rp_*                       src/lpf_hilbert_quad.c:70-188
 *   sound_render_*                src/sound_render.c:588-915
 *   mtrnd_*                       src/mersene_twister/mt_jrnd.c:28-256
 */
#include "in_cwave.h"

#include <stdio.h>

/* ---- knobs that replace the config file (reference defaults: src/config.c:118-207) ---- */
typedef struct icwref_cfg {
    int      filter_no;          /* IIR_HBLPF_IX   default 1 */
    int      is_kahan;           /* IIR_SUM_KAHAN  default 1 */
    int      is_subnorm_reject;  /* IIR_SUBN_ZERO  default 1 */
    double   subnorm_thr;        /* IIR_SUBN_THR   default 1e-150 */
    int      is_frmod_scaled;    /* FRMOD_SCALED   default 1 */
    int      need24bits;         /* NEED24BITS     default 1 */
    int      is_fp_check;        /* FP_CHECK       default 0 */
    double   dth_bits;           /* DITHER_BITS    default 1.0 */
    unsigned quantz_type;        /* QUANTIZE_TYPE  default 1 (mid riser) */
    unsigned render_type;        /* RENDER_TYPE    default 0 (round) */
    unsigned nshape_type;        /* NOISE_SHAPING  default 0 (flat) */
    unsigned sign_bits16;        /* SIGNBITS16     default 16 */
    unsigned sign_bits24;        /* SIGNBITS24     default 24 */
    unsigned sec_align;          /* SEC_ALIGN      default 0 */
    unsigned fade_in;            /* FADE_IN  ms    default 0 */
    unsigned fade_out;           /* FADE_OUT ms    default 0 */
    int      clr_nframe_trk;     /* CLR_NFRAME_PT  default 0 */
    int      clr_hilb_trk;       /* CLR_HILB_PT    default 0 */
} icwref_cfg;

static icwref_cfg g_cfg;
static const char *g_cfg_path;      /* non-NULL: the next plugin init reads this reference-format config file */
static int g_cfg_loaded;            /* what load_config() returned for it */

/* ---- the four symbols from files we do not build -------------------------------------- */
BOOL load_config_default(void)
{
    if (g_cfg_path) {
        g_cfg_loaded = load_config(g_cfg_path);     /* src/config.c:815: defaults + NULL list on any error */
        return g_cfg_loaded;
    }
    the.cfg.ver_config            = 10;
    the.cfg.is_wav_support        = TRUE;
    the.cfg.is_rwave_support      = FALSE;
    the.cfg.infobox_parenting     = INFOBOX_LISTPARENT;
    the.cfg.enable_unload_cleanup = FALSE;
    the.cfg.play_sleep            = DEF_PLAY_SLEEP;
    the.cfg.disable_play_sleep    = FALSE;
    the.cfg.sec_align             = g_cfg.sec_align;
    the.cfg.fade_in               = g_cfg.fade_in;
    the.cfg.fade_out              = g_cfg.fade_out;
    the.cfg.is_frmod_scaled       = g_cfg.is_frmod_scaled;
    the.cfg.iir_filter_no         = (unsigned)g_cfg.filter_no;
    the.cfg.iir_comp_config.is_kahan          = g_cfg.is_kahan;
    the.cfg.iir_comp_config.is_subnorm_reject = g_cfg.is_subnorm_reject;
    the.cfg.iir_comp_config.subnorm_thr       = g_cfg.subnorm_thr;
    the.cfg.is_clr_nframe_trk     = g_cfg.clr_nframe_trk;
    the.cfg.is_clr_hilb_trk       = g_cfg.clr_hilb_trk;
    the.cfg.show_long_numbers     = FALSE;
    the.cfg.is_fp_check           = g_cfg.is_fp_check;
    the.cfg.need24bits            = g_cfg.need24bits;
    the.cfg.sr_config.dth_bits    = g_cfg.dth_bits;
    the.cfg.sr_config.quantz_type = g_cfg.quantz_type;
    the.cfg.sr_config.render_type = g_cfg.render_type;
    the.cfg.sr_config.nshape_type = g_cfg.nshape_type;
    the.cfg.sr_config.sign_bits16 = g_cfg.sign_bits16;
    the.cfg.sr_config.sign_bits24 = g_cfg.sign_bits24;
    the.cfg.dsp_list              = NULL;
    return TRUE;
}
BOOL save_config_default(void) { return TRUE; }
void make_exts_list(BOOL is_wav, BOOL is_rwave) { (void)is_wav; (void)is_rwave; }
In_Module *get_playback_iface(void) { return NULL; }

/* ---- facade ---------------------------------------------------------------------------- */

void icwref_default_cfg(icwref_cfg *c)
{
    memset(c, 0, sizeof(*c));
    c->filter_no = IX_LPF_HILB_DEF;
    c->is_kahan = 1;
    c->is_subnorm_reject = 1;
    c->subnorm_thr = SBN_THR_DEF;
    c->is_frmod_scaled = 1;
    c->need24bits = 1;
    c->dth_bits = DEF_DITHER_BITS;
    c->quantz_type = SND_QUANTZ_MID_RISER;
    c->render_type = SND_RENDER_ROUND;
    c->nshape_type = SND_NSHAPE_FLAT;
    c->sign_bits16 = DEF_SIGN_BITS16;
    c->sign_bits24 = DEF_SIGN_BITS24;
}

/* fresh plugin state == one independent stream ("checkpoint/resume") */
/* present only in the build whose frame loop runs on the GPU (in_cwave_b200/host/adv_modulator_gpu.c): a fresh plugin
   instance re-seeds the dither generators, so the sessions that hold their positions go too */
void amod_gpu_reset(void) __attribute__((weak));

void icwref_reset(const icwref_cfg *c)
{
    g_cfg = *c;
    (void)winampGetInModule2();
    if (amod_gpu_reset) amod_gpu_reset();
}

/* fresh plugin state configured by the reference's own parser from a config file; returns load_config()'s verdict */
int icwref_reset_from_file(const char *path)
{
    g_cfg_path = path;
    g_cfg_loaded = 0;
    (void)winampGetInModule2();
    if (amod_gpu_reset) amod_gpu_reset();
    g_cfg_path = NULL;
    return g_cfg_loaded;
}

/* the live configuration, flattened */
void icwref_get_cfg(icwref_cfg *c)
{
    memset(c, 0, sizeof(*c));
    c->filter_no = (int)the.cfg.iir_filter_no;
    c->is_kahan = the.cfg.iir_comp_config.is_kahan;
    c->is_subnorm_reject = the.cfg.iir_comp_config.is_subnorm_reject;
    c->subnorm_thr = the.cfg.iir_comp_config.subnorm_thr;
    c->is_frmod_scaled = the.cfg.is_frmod_scaled;
    c->need24bits = the.cfg.need24bits;
    c->is_fp_check = the.cfg.is_fp_check;
    c->dth_bits = the.cfg.sr_config.dth_bits;
    c->quantz_type = the.cfg.sr_config.quantz_type;
    c->render_type = the.cfg.sr_config.render_type;
    c->nshape_type = the.cfg.sr_config.nshape_type;
    c->sign_bits16 = the.cfg.sr_config.sign_bits16;
    c->sign_bits24 = the.cfg.sr_config.sign_bits24;
    c->sec_align = the.cfg.sec_align;
    c->fade_in = the.cfg.fade_in;
    c->fade_out = the.cfg.fade_out;
    c->clr_nframe_trk = the.cfg.is_clr_nframe_trk;
    c->clr_hilb_trk = the.cfg.is_clr_hilb_trk;
}

/* Write the live configuration and DSP list with the reference's save_config() (src/config.c:924).
 * As in the reference's module_cleanup (src/in_cwave.c:575-595) the list is handed to the config and
 * freed while it is written: the plugin must be re-initialised (icwref_reset*) afterwards. */
int icwref_save_config(const char *path)
{
    the.cfg.dsp_list = amod_cleanup(TRUE);
    return save_config(path);
}

/* a renderer-parameter change on the LIVE plugin, as the GUI thread does it (src/amod_gui_control.c:1555-1602):
   no re-initialisation, both contexts keep their state; takes effect at the next block */
void icwref_set_render_live(const icwref_cfg *c)
{
    SR_VCONFIG v;
    v.dth_bits = c->dth_bits;
    v.quantz_type = c->quantz_type;
    v.render_type = c->render_type;
    v.nshape_type = c->nshape_type;
    v.sign_bits16 = c->sign_bits16;
    v.sign_bits24 = c->sign_bits24;
    srenders_set_vcfg(&v);
}

/* mod_context_reset_hilbert / _reset_framecnt on the transcode context (src/in_cwave.c:162,287) */
void icwref_reset_live(int hilbert, int framecnt)
{
    if (hilbert) mod_context_reset_hilbert(&the.mc_transcode);
    if (framecnt) mod_context_reset_framecnt(&the.mc_transcode);
}

/* flat description of one DSP-list node, in EXECUTION order (master last) */
typedef struct icwref_node {
    int      mode;               /* MODE_MASTER/SHIFT/PM/MIX */
    unsigned inputs_mask;        /* bit k = plug k enabled (0 = In, 1..26 = A..Z) */
    int      xch_mode;
    int      l_iq_invert, r_iq_invert;
    double   l_gain, r_gain;
    int      n_out;              /* ignored for master */
    int      l_tout, r_tout;     /* master */
    int      l_on, r_on;         /* shift: is_shift; pm: is_pm */
    double   l_p[4], r_p[4];     /* shift: {fr_shift}; pm: {freq, phase, level, angle} */
} icwref_node;

static void fill_common(NODE_DSP *nd, const icwref_node *s)
{
    int k;
    for (k = 0; k < N_INPUTS; ++k)
        nd->inputs[k] = (char)((s->inputs_mask >> k) & 1u);
    nd->xch_mode = s->xch_mode;
    nd->l_iq_invert = s->l_iq_invert;
    nd->r_iq_invert = s->r_iq_invert;
    nd->l_gain = s->l_gain;
    nd->r_gain = s->r_gain;
    nd->lock_gain = 0;
}

/* replaces the whole list; nodes[n-1] must be the master. returns 0 on success */
int icwref_set_graph(const icwref_node *nodes, int n, int bypass)
{
    int i;
    NODE_DSP *nd;

    if (n < 1 || nodes[n - 1].mode != MODE_MASTER)
        return -1;
    amod_del_dsplist();
    nd = amod_get_headdsp();
    fill_common(nd, &nodes[n - 1]);
    nd->dsp.mk_master.le.tout = nodes[n - 1].l_tout;
    nd->dsp.mk_master.ri.tout = nodes[n - 1].r_tout;

    /* tail executes first: append the node that runs just before the master first */
    for (i = n - 2; i >= 0; --i) {
        const icwref_node *s = &nodes[i];
        nd = amod_add_lastdsp("n", s->mode);
        if (!nd)
            return -2;
        fill_common(nd, s);
        switch (s->mode) {
        case MODE_SHIFT:
            nd->dsp.mk_shift.le.fr_shift = s->l_p[0];
            nd->dsp.mk_shift.ri.fr_shift = s->r_p[0];
            nd->dsp.mk_shift.le.is_shift = s->l_on;
            nd->dsp.mk_shift.ri.is_shift = s->r_on;
            nd->dsp.mk_shift.n_out = s->n_out;
            nd->dsp.mk_shift.lock_shift = 0;
            break;
        case MODE_PM:
            nd->dsp.mk_pm.le.freq = s->l_p[0];  nd->dsp.mk_pm.ri.freq = s->r_p[0];
            nd->dsp.mk_pm.le.phase = s->l_p[1]; nd->dsp.mk_pm.ri.phase = s->r_p[1];
            nd->dsp.mk_pm.le.level = s->l_p[2]; nd->dsp.mk_pm.ri.level = s->r_p[2];
            nd->dsp.mk_pm.le.angle = s->l_p[3]; nd->dsp.mk_pm.ri.angle = s->r_p[3];
            nd->dsp.mk_pm.le.is_pm = s->l_on;
            nd->dsp.mk_pm.ri.is_pm = s->r_on;
            nd->dsp.mk_pm.n_out = s->n_out;
            nd->dsp.mk_pm.lock_freq = nd->dsp.mk_pm.lock_phase = 0;
            nd->dsp.mk_pm.lock_level = nd->dsp.mk_pm.lock_angle = 0;
            break;
        case MODE_MIX:
            nd->dsp.mk_mix.n_out = s->n_out;
            break;
        default:
            return -3;
        }
    }
    amod_set_bypass_list_flag(bypass ? TRUE : FALSE);
    return 0;
}

/* the live DSP list in execution order (tail first, master last); returns the node count or -1 if cap is short */
int icwref_get_graph(icwref_node *nodes, int cap)
{
    NODE_DSP *nd = amod_get_headdsp();
    int n = 0, k;
    while (nd->next) nd = nd->next;
    for (; nd; nd = nd->prev, ++n) {
        icwref_node *o;
        if (n >= cap) return -1;
        o = &nodes[n];
        memset(o, 0, sizeof(*o));
        o->mode = nd->mode;
        for (k = 0; k < N_INPUTS; ++k)
            if (nd->inputs[k]) o->inputs_mask |= 1u << k;
        o->xch_mode = nd->xch_mode;
        o->l_iq_invert = nd->l_iq_invert; o->r_iq_invert = nd->r_iq_invert;
        o->l_gain = nd->l_gain; o->r_gain = nd->r_gain;
        switch (nd->mode) {
        case MODE_MASTER:
            o->l_tout = nd->dsp.mk_master.le.tout; o->r_tout = nd->dsp.mk_master.ri.tout;
            break;
        case MODE_SHIFT:
            o->n_out = nd->dsp.mk_shift.n_out;
            o->l_on = nd->dsp.mk_shift.le.is_shift; o->r_on = nd->dsp.mk_shift.ri.is_shift;
            o->l_p[0] = nd->dsp.mk_shift.le.fr_shift; o->r_p[0] = nd->dsp.mk_shift.ri.fr_shift;
            break;
        case MODE_PM:
            o->n_out = nd->dsp.mk_pm.n_out;
            o->l_on = nd->dsp.mk_pm.le.is_pm; o->r_on = nd->dsp.mk_pm.ri.is_pm;
            o->l_p[0] = nd->dsp.mk_pm.le.freq;  o->l_p[1] = nd->dsp.mk_pm.le.phase;
            o->l_p[2] = nd->dsp.mk_pm.le.level; o->l_p[3] = nd->dsp.mk_pm.le.angle;
            o->r_p[0] = nd->dsp.mk_pm.ri.freq;  o->r_p[1] = nd->dsp.mk_pm.ri.phase;
            o->r_p[2] = nd->dsp.mk_pm.ri.level; o->r_p[3] = nd->dsp.mk_pm.ri.angle;
            break;
        case MODE_MIX:
            o->n_out = nd->dsp.mk_mix.n_out;
            break;
        }
    }
    return n;
}

/* Run one file through mod_context_fopen / amod_process_samples / mod_context_fclose on the
 * transcode context.  read_quant = frames per amod_process_samples call.  When bus_tap is
 * non-NULL the file is stepped ONE frame per call and after each frame the plugs listed in
 * tap_plugs are copied out as (L.re, L.im, R.re, R.im).  Returns frames rendered, <0 on error. */
int64_t icwref_process_file(const char *path, unsigned read_quant,
                            char *pcm, int64_t pcm_cap,
                            double *bus_tap, const int *tap_plugs, int n_tap)
{
    MOD_CONTEXT *mc = &the.mc_transcode;
    unsigned fb;
    int64_t frames = 0;
    int got;

    if (bus_tap)
        read_quant = 1;
    if (!read_quant)
        return -1;
    if (!mod_context_fopen(path, read_quant, mc))
        return -2;
    fb = sound_render_size(&mc->sr_left) + sound_render_size(&mc->sr_right);
    for (;;) {
        if ((frames + read_quant) * (int64_t)fb > pcm_cap) {
            /* shrink the last request instead of overrunning the caller */
            int64_t room = pcm_cap / fb - frames;
            if (room <= 0)
                break;
            xwave_change_read_quant((unsigned)room, mc->xr);
            read_quant = (unsigned)room;
        }
        got = amod_process_samples(pcm + frames * fb, mc);
        if (got <= 0)
            break;
        if (bus_tap) {
            int j;
            for (j = 0; j < n_tap; ++j) {
                const LRCOMPLEX *c = &mc->inout[tap_plugs[j]];
                double *d = bus_tap + (frames * n_tap + j) * 4;
                d[0] = c->le.re; d[1] = c->le.im; d[2] = c->ri.re; d[3] = c->ri.im;
            }
        }
        frames += got;
    }
    mod_context_fclose(mc);
    return frames;
}

/* transcode.c exports these without a header declaration (Winamp resolves them by name) */
intptr_t winampGetExtendedRead_open(const TCHAR *filename, int *size, int *bps, int *nch, int *srate);
intptr_t winampGetExtendedRead_getData(intptr_t handle, char *dest, int len, int *killswitch);
int winampGetExtendedRead_setTime(intptr_t handle, int decode_pos_ms);
void winampGetExtendedRead_close(intptr_t handle);

/* The exported transcode wrapper itself (src/transcode.c:40-118), chunk = bytes per getData */
int64_t icwref_transcode_file(const char *path, int chunk, char *pcm, int64_t pcm_cap, int info[4])
{
    int kill = 0;
    int64_t total = 0;
    intptr_t h = winampGetExtendedRead_open(path, &info[0], &info[1], &info[2], &info[3]);

    if (!h)
        return -1;
    for (;;) {
        int want = chunk;
        intptr_t got;
        if (total + want > pcm_cap)
            want = (int)(pcm_cap - total);
        if (want <= 0)
            break;
        got = winampGetExtendedRead_getData(h, pcm + total, want, &kill);
        if (got <= 0)
            break;
        total += got;
    }
    winampGetExtendedRead_close(h);
    return total;
}

typedef struct icwref_stats {
    unsigned l_clips, r_clips;
    double   l_peak, r_peak;
    uint64_t subnorm_cnt;
    uint64_t n_frame;
} icwref_stats;

void icwref_get_stats(icwref_stats *st, int reset)
{
    amod_get_clips_peaks(&st->l_clips, &st->r_clips, &st->l_peak, &st->r_peak, FALSE);
    st->subnorm_cnt = mod_context_get_desubnorm_counter(&the.mc_transcode);
    st->n_frame = mod_context_get_framecnt(&the.mc_transcode);
    if (reset) {
        unsigned a, b; double c, d;
        amod_get_clips_peaks(&a, &b, &c, &d, TRUE);
    }
}

/* FP exception counters of the transcode context: [hilbert L, hilbert R, render L, render R] x
 * [total, snan, qnan, ninf, nden, pden, pinf] (the reference's own getter, src/in_cwave.c:383-440) */
void icwref_fp_stats(unsigned out[4][7])
{
    FP_EXCEPT_STATS s[4];
    memset(s, 0, sizeof s);
    fecs_getcnts(&s[0], &s[1], &s[2], &s[3]);
    for (int i = 0; i < 4; ++i) {
        out[i][0] = s[i].cnt_total; out[i][1] = s[i].cnt_snan; out[i][2] = s[i].cnt_qnan; out[i][3] = s[i].cnt_ninf;
        out[i][4] = s[i].cnt_nden;  out[i][5] = s[i].cnt_pden; out[i][6] = s[i].cnt_pinf;
    }
}

/* ---- leaf taps ------------------------------------------------------------------------- */

/* real -> analytic for one channel from a fresh converter; returns the subnorm-reject count */
uint64_t icwref_hilbert(unsigned filter_no, int is_kahan, int is_reject,
                        const double *x, int64_t n, double *out_i, double *out_q)
{
    IIR_COMP_CONFIG cc;
    FP_EXCEPT_STATS fes;
    LPF_HILBERT_QUAD *h;
    uint64_t cnt;
    int64_t k;

    cc.is_kahan = is_kahan;
    cc.is_subnorm_reject = is_reject;
    cc.subnorm_thr = SBN_THR_DEF;
    memset(&fes, 0, sizeof(fes));
    h = hq_rp_create_ix(filter_no, &cc);
    for (k = 0; k < n; ++k)
        hq_rp_process(x[k], &out_i[k], &out_q[k], h, &fes);
    cnt = hq_rp_get_sncnt(h);
    hq_rp_destroy(h);
    return cnt;
}

/* one half-band LPF alone (src/hblpf.c:894,1008) */
void icwref_iir(unsigned filter_no, int is_kahan, int is_reject,
                const double *x, int64_t n, double *y)
{
    IIR_COMP_CONFIG cc;
    FP_EXCEPT_STATS fes;
    IIR_RAT_POLY *f;
    int64_t k;

    cc.is_kahan = is_kahan;
    cc.is_subnorm_reject = is_reject;
    cc.subnorm_thr = SBN_THR_DEF;
    memset(&fes, 0, sizeof(fes));
    f = iir_rp_create(&iir_hb_lpf_const_filters[filter_no], &cc);
    for (k = 0; k < n; ++k)
        y[k] = iir_rp_process(x[k], f, &fes);
    iir_rp_destroy(f);
}

/* one render channel from a fresh SOUND_RENDER; returns bytes written */
int64_t icwref_render(const icwref_cfg *c, uint32_t seed, const double *in, int64_t n,
                      char *out, unsigned *clips, double *peak)
{
    SR_VCONFIG v;
    SOUND_RENDER sr;
    FP_EXCEPT_STATS fes;
    char *p = out;
    int64_t k;

    v.dth_bits = c->dth_bits;
    v.quantz_type = c->quantz_type;
    v.render_type = c->render_type;
    v.nshape_type = c->nshape_type;
    v.sign_bits16 = c->sign_bits16;
    v.sign_bits24 = c->sign_bits24;
    memset(&fes, 0, sizeof(fes));
    memset(&sr, 0, sizeof(sr));
    sound_render_init(&v, c->need24bits, seed, &sr);
    *clips = 0;
    *peak = SR_ZERO_SIGNAL_DB;
    for (k = 0; k < n; ++k)
        sound_render_value(&p, in[k], clips, peak, &sr, &fes);
    sound_render_cleanup(&sr);
    return (int64_t)(p - out);
}

/* raw generator words after mtrnd_init_seed (the seeding the plugin uses, sound_render.c:590) */
void icwref_mt_words(uint32_t seed, int64_t skip, int64_t n, uint32_t *out)
{
    MT_JRND_STATE st;
    int64_t k;
    mtrnd_init_seed(&st, seed);
    for (k = 0; k < skip; ++k)
        (void)mtrnd_gen_ui32(&st);
    for (k = 0; k < n; ++k)
        out[k] = mtrnd_gen_ui32(&st);
}

/* the reference's own known-answer seeding: init_key then n words (test_mt_jrnd/main.c:30-60) */
void icwref_mt_words_key(const uint32_t *key, uint32_t key_len, int64_t n, uint32_t *out)
{
    MT_JRND_STATE st;
    int64_t k;
    mtrnd_init_key(&st, key, key_len);
    for (k = 0; k < n; ++k)
        out[k] = mtrnd_gen_ui32(&st);
}

void icwref_mt_dsopen(uint32_t seed, int64_t n, double *out)
{
    MT_JRND_STATE st;
    int64_t k;
    mtrnd_init_seed(&st, seed);
    for (k = 0; k < n; ++k)
        out[k] = mtrnd_gen_dsopen(&st);
}
