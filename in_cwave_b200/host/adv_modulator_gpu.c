/* adv_modulator_gpu.c -- the reference-side binding: amod_process_samples() over libicw_b200.so.
 *
 * This file is compiled TOGETHER WITH THE REFERENCE'S OWN SOURCES (it includes the reference's in_cwave.h) and takes
 * the place of two functions of src/adv_modulator.c:
 *
 *     int  amod_process_samples(char *buf, MOD_CONTEXT *mc);                    src/adv_modulator.c:587-763
 *     void amod_get_clips_peaks(unsigned *, unsigned *, double *, double *, BOOL);   src/adv_modulator.c:445-465
 *
 * (the reference's own definitions are renamed away on the compiler's command line: -Damod_process_samples=... when
 * adv_modulator.c is compiled; nothing in the reference tree is edited).  Everything else -- the reader, the transcode
 * entry points, the config file, the DSP-list editor, the reset functions -- stays the reference's code and keeps
 * working on the reference's own objects, because MOD_CONTEXT remains the owner of the state: before a block the
 * context's arithmetic content (frame counter, both half-band filters' delay lines and mixer phase, the 27-plug bus,
 * sloped-TPDF memory, noise-shaper memory) is handed to the session, after the block it is written back.  The one
 * thing that lives on the GPU side is the position of the two dither generators (seed + words drawn: the reference
 * keeps 624 state words per generator that nothing but the renderer reads).
 *
 * Built by `make -C oracle ref_gpu` into oracle/_ref_gpu/libicw_ref_gpu.so (the reference compiles here, on Linux,
 * with oracle/shim/ standing in for <windows.h>); tests/test_ref_gpu.py drives the reference's own
 * winampGetExtendedRead_* through it and compares with the pure reference.  There is no CPU fallback: a failing GPU
 * call ends the block with 0 frames, which the reference's callers treat as end of data / error.
 */
#include "in_cwave.h"
#include "icw_b200.h"

#include <stdio.h>
#include <string.h>

static icw_engine  *g_engine;
static icw_session *g_session[2];               /* [0] playback, [1] transcode: the two MOD_CONTEXTs (src/in_cwave.h:470-471) */
static unsigned     g_fp_seen[2][4][7];         /* FP-exception counters already handed to the context */
static unsigned     g_clips[2];                 /* what the reference keeps in its static `am` (src/adv_modulator.c:53-54) */
static double       g_peak[2];                  /* linear; the getter converts like src/sound_render.c:773-775 */
static char         g_err[512];

const char *amod_gpu_last_error(void) { return g_err; }

static int fail(const char *what)
{
    snprintf(g_err, sizeof g_err, "%s: %s", what, icw_last_error());
    return 0;
}

/* ---- parameters: config + reader facts + DSP list -> icw_chain_spec ------------------------------------------------ */
static void spec_from_plugin(icw_chain_spec *sp, const MOD_CONTEXT *mc)
{
    const XWAVE_READER *xr = mc->xr;
    const NODE_DSP *nd;
    int n = 0;

    icw_default_spec(sp);
    sp->fmt = xr->is_sample_complex ? 16 + (int)xr->spec.cwave.header.format    /* HCW_FMT_*, src/cwave.h:76-84 */
                                    : (int)xr->spec.rwave.format;                /* HRW_FMT_*, src/in_cwave.h:326-330 */
    sp->n_channels  = (int)xr->n_channels;
    sp->sample_rate = xr->sample_rate;
    sp->n_samples   = xr->n_samples;
    sp->n_fade_in   = xr->n_fade_in;
    sp->n_fade_out  = xr->n_fade_out;
    sp->filter_no         = (int)the.cfg.iir_filter_no;
    sp->is_kahan          = the.cfg.iir_comp_config.is_kahan;
    sp->is_subnorm_reject = the.cfg.iir_comp_config.is_subnorm_reject;
    sp->hilbert_mode      = ICW_HILBERT_EXACT;                                  /* the reference's own rounding sequence */
    sp->is_frmod_scaled   = the.cfg.is_frmod_scaled;
    sp->is_fp_check       = the.cfg.is_fp_check;
    /* the renderer's volatile parameters as the per-frame code sees them (srenders_set_vcfg writes them into every
       SOUND_RENDER, src/in_cwave.c) */
    sp->need24bits  = mc->sr_left.is24bits ? 1 : 0;
    sp->dth_bits    = mc->sr_left.config.dth_bits;
    sp->quantz_type = mc->sr_left.config.quantz_type;
    sp->render_type = mc->sr_left.config.render_type;
    sp->nshape_type = mc->sr_left.config.nshape_type;
    sp->sign_bits16 = mc->sr_left.config.sign_bits16;
    sp->sign_bits24 = mc->sr_left.config.sign_bits24;
    sp->bypass      = amod_get_bypass_list_flag();

    /* the list is walked tail -> head (src/adv_modulator.c:637): emit it in that order, master last */
    for (nd = amod_get_headdsp(); nd->next; nd = nd->next) ;
    for (; nd && n < ICW_MAX_NODES; nd = nd->prev, ++n) {
        icw_node *o = &sp->nodes[n];
        int k;
        memset(o, 0, sizeof *o);
        o->mode = nd->mode;
        for (k = 0; k < N_INPUTS; ++k)
            if (nd->inputs[k]) o->inputs_mask |= 1u << k;
        o->xch_mode = nd->xch_mode;
        o->l_iq_invert = nd->l_iq_invert;  o->r_iq_invert = nd->r_iq_invert;
        o->l_gain = nd->l_gain;            o->r_gain = nd->r_gain;
        switch (nd->mode) {
        case MODE_MASTER: o->l_tout = nd->dsp.mk_master.le.tout; o->r_tout = nd->dsp.mk_master.ri.tout; break;
        case MODE_SHIFT:  o->n_out = nd->dsp.mk_shift.n_out;
                          o->l_on = nd->dsp.mk_shift.le.is_shift; o->r_on = nd->dsp.mk_shift.ri.is_shift;
                          o->l_p[0] = nd->dsp.mk_shift.le.fr_shift; o->r_p[0] = nd->dsp.mk_shift.ri.fr_shift; break;
        case MODE_PM:     o->n_out = nd->dsp.mk_pm.n_out;
                          o->l_on = nd->dsp.mk_pm.le.is_pm; o->r_on = nd->dsp.mk_pm.ri.is_pm;
                          o->l_p[0] = nd->dsp.mk_pm.le.freq;  o->l_p[1] = nd->dsp.mk_pm.le.phase;
                          o->l_p[2] = nd->dsp.mk_pm.le.level; o->l_p[3] = nd->dsp.mk_pm.le.angle;
                          o->r_p[0] = nd->dsp.mk_pm.ri.freq;  o->r_p[1] = nd->dsp.mk_pm.ri.phase;
                          o->r_p[2] = nd->dsp.mk_pm.ri.level; o->r_p[3] = nd->dsp.mk_pm.ri.angle; break;
        case MODE_MIX:    o->n_out = nd->dsp.mk_mix.n_out; break;
        }
    }
    sp->n_nodes = n;
}

/* ---- state: MOD_CONTEXT <-> icw_stream_state ----------------------------------------------------------------------- */
/* delay line: the reference inserts at pz[ix] and reads backwards from ix - 1 (src/hblpf.c:903-921), so the value
   j + 1 frames old sits at pz[ix - 1 - j] (mod nord) -- icw_stream_state.hb[c][f][j] by definition */
static void iir_to_state(const IIR_RAT_POLY *f, double *hb, uint64_t *rejects)
{
    int j;
    for (j = 0; j < ICW_MAX_ORD; ++j) hb[j] = 0.0;
    for (j = 0; j < f->nord && j < ICW_MAX_ORD; ++j) hb[j] = f->pz[(f->ix - 1 - j + 2 * f->nord) % f->nord];
    *rejects = f->subnorm_cnt;
}
static void state_to_iir(IIR_RAT_POLY *f, const double *hb, uint64_t rejects)
{
    int j;
    f->ix = 0;
    for (j = 0; j < f->nord && j < ICW_MAX_ORD; ++j) f->pz[f->nord - 1 - j] = hb[j];
    f->subnorm_cnt = rejects;
}
/* shaper memory: the newest value sits at ns_ix_pos, older ones follow upwards (src/sound_render.c:403-489) */
static void ns_to_state(const SOUND_RENDER *sr, double *e, double *o, double *prev_err)
{
    const NS_SHAPER *ns = &sr->ns_shaper;
    unsigned a, n = ns->dsc ? ns->dsc->num_coeffs : 0;
    for (a = 0; a < ICW_NS_MAX_TAPS; ++a) e[a] = o[a] = 0.0;
    for (a = 0; a < n && a < ICW_NS_MAX_TAPS; ++a) {
        if (ns->ns_ebuffer) e[a] = ns->ns_ebuffer[((unsigned)ns->ns_ix_pos + a) % n];
        if (ns->ns_obuffer) o[a] = ns->ns_obuffer[((unsigned)ns->ns_ix_pos + a) % n];
    }
    *prev_err = sr->prev_ns_err;
}
static void state_to_ns(SOUND_RENDER *sr, const double *e, const double *o, double prev_err)
{
    NS_SHAPER *ns = &sr->ns_shaper;
    unsigned a, n = ns->dsc ? ns->dsc->num_coeffs : 0;
    ns->ns_ix_pos = 0;
    for (a = 0; a < n && a < ICW_NS_MAX_TAPS; ++a) {
        if (ns->ns_ebuffer) ns->ns_ebuffer[a] = e[a];
        if (ns->ns_obuffer) ns->ns_obuffer[a] = o[a];
    }
    sr->prev_ns_err = prev_err;
}

static void context_to_state(MOD_CONTEXT *mc, icw_stream_state *st)
{
    const XWAVE_READER *xr = mc->xr;
    LPF_HILBERT_QUAD *h[2];
    int c, k;
    h[0] = mc->h_left; h[1] = mc->h_right;
    st->n_frame = mc->n_frame;
    /* absolute index of the block's first frame in the file incl. the silence tail: the fade ramps' argument
       (src/xwave_reader.c:921-936); xwave_read_samples has already moved the reader past the block */
    st->pos = (xr->pos_samples + xr->pos_tail) - (int64_t)xr->really_readed;
    for (c = 0; c < 2; ++c) {
        iir_to_state(h[c]->iir_I, st->hb[c][0], &st->hb_rejects[c][0]);
        iir_to_state(h[c]->iir_Q, st->hb[c][1], &st->hb_rejects[c][1]);
        st->quad[c] = h[c]->sampe_ix & 3u;
    }
    st->hb_basis = 0;
    st->prev_rnd[0] = mc->sr_left.prev_rnd;
    st->prev_rnd[1] = mc->sr_right.prev_rnd;
    st->clips[0] = st->clips[1] = 0;            /* per block: summed into g_clips / g_peak afterwards */
    st->peak[0] = st->peak[1] = 0.0;
    for (k = 0; k < N_INPUTS && k < ICW_N_PLUGS; ++k) {
        st->bus[k][0] = mc->inout[k].le.re; st->bus[k][1] = mc->inout[k].le.im;
        st->bus[k][2] = mc->inout[k].ri.re; st->bus[k][3] = mc->inout[k].ri.im;
    }
    ns_to_state(&mc->sr_left, st->ns_e[0], st->ns_o[0], &st->ns_prev_err[0]);
    ns_to_state(&mc->sr_right, st->ns_e[1], st->ns_o[1], &st->ns_prev_err[1]);
    /* mt_seed / mt_drawn: left as the session holds them */
}

static void state_to_context(const icw_stream_state *st, MOD_CONTEXT *mc)
{
    LPF_HILBERT_QUAD *h[2];
    int c, k;
    h[0] = mc->h_left; h[1] = mc->h_right;
    mc->n_frame = st->n_frame;
    for (c = 0; c < 2; ++c) {
        state_to_iir(h[c]->iir_I, st->hb[c][0], st->hb_rejects[c][0]);
        state_to_iir(h[c]->iir_Q, st->hb[c][1], st->hb_rejects[c][1]);
        h[c]->sampe_ix = st->quad[c];
    }
    mc->sr_left.prev_rnd = st->prev_rnd[0];
    mc->sr_right.prev_rnd = st->prev_rnd[1];
    for (k = 0; k < N_INPUTS && k < ICW_N_PLUGS; ++k) {
        mc->inout[k].le.re = st->bus[k][0]; mc->inout[k].le.im = st->bus[k][1];
        mc->inout[k].ri.re = st->bus[k][2]; mc->inout[k].ri.im = st->bus[k][3];
    }
    state_to_ns(&mc->sr_left, st->ns_e[0], st->ns_o[0], st->ns_prev_err[0]);
    state_to_ns(&mc->sr_right, st->ns_e[1], st->ns_o[1], st->ns_prev_err[1]);
    g_clips[0] += st->clips[0]; g_clips[1] += st->clips[1];
    if (st->peak[0] > g_peak[0]) g_peak[0] = st->peak[0];
    if (st->peak[1] > g_peak[1]) g_peak[1] = st->peak[1];
}

static void add_fp_counters(FP_EXCEPT_STATS *fes, const uint32_t now[7], unsigned seen[7])
{
    volatile unsigned *dst[7];
    int k;
    dst[0] = &fes->cnt_total; dst[1] = &fes->cnt_snan; dst[2] = &fes->cnt_qnan; dst[3] = &fes->cnt_ninf;
    dst[4] = &fes->cnt_nden;  dst[5] = &fes->cnt_pden; dst[6] = &fes->cnt_pinf;
    for (k = 0; k < 7; ++k) {
        *dst[k] += now[k] - seen[k];
        seen[k] = now[k];
    }
}

/* ---- the two replaced entry points ---------------------------------------------------------------------------------- */
int amod_process_samples(char *buf, MOD_CONTEXT *mc)       /* same signature, same return value */
{
    const int which = (mc == &the.mc_transcode);
    icw_chain_spec sp;
    icw_stream_state st;
    XWAVE_READER *xr = mc->xr;

    if (!xwave_read_samples(xr))                           /* file I/O and the silence tail stay the reference's */
        return 0;                                          /* 0 == EOF or read error, as before */
    if (!g_engine && icw_engine_create(0, &g_engine) != ICW_OK)
        return fail("icw_engine_create");
    spec_from_plugin(&sp, mc);
    if (!g_session[which]) {
        if (icw_session_create(g_engine, &sp, 1, &g_session[which]) != ICW_OK)
            return fail("icw_session_create");
    } else if (icw_session_set_spec(g_session[which], &sp) != ICW_OK) {
        return fail("icw_session_set_spec");
    }
    if (icw_session_get_state(g_session[which], 0, &st) != ICW_OK)
        return fail("icw_session_get_state");
    context_to_state(mc, &st);
    if (icw_session_set_state(g_session[which], 0, &st) != ICW_OK)
        return fail("icw_session_set_state");
    /* xr->tbuff holds really_readed frames incl. the virtual silence tail (src/xwave_reader.c:838-904) */
    if (icw_session_process_host(g_session[which], xr->really_readed, xr->tbuff, 0, buf, 0) != ICW_OK)
        return fail("icw_session_process_host");
    if (icw_session_get_state(g_session[which], 0, &st) != ICW_OK)
        return fail("icw_session_get_state");
    state_to_context(&st, mc);
    if (the.cfg.is_fp_check) {
        uint32_t now[4][7];
        if (icw_session_fp_stats(g_session[which], 0, now) == ICW_OK) {
            add_fp_counters(&mc->fes_hilb_left, now[0], g_fp_seen[which][0]);
            add_fp_counters(&mc->fes_hilb_right, now[1], g_fp_seen[which][1]);
            add_fp_counters(&mc->fes_sr_left, now[2], g_fp_seen[which][2]);
            add_fp_counters(&mc->fes_sr_right, now[3], g_fp_seen[which][3]);
        }
    }
    xr->unpacked = xr->really_readed;
    xr->ptr_tbuff = xr->tbuff + (size_t)xr->really_readed * xr->sample_size;
    return (int)xr->unpacked;
}

void amod_get_clips_peaks(unsigned *lc, unsigned *rc, double *lpv, double *rpv, BOOL isReset)
{
    if (isReset) {
        g_clips[0] = g_clips[1] = 0;
        g_peak[0] = g_peak[1] = 0.0;
    }
    *lc = g_clips[0];
    *rc = g_clips[1];
    *lpv = icw_peak_db(g_peak[0]);
    *rpv = icw_peak_db(g_peak[1]);
}

/* test hook: drop the sessions (the next block starts the generators from their seeds again, like a fresh plugin
   instance -- the reference's own winampGetInModule2 re-seeds them, src/in_cwave.c:69-70) */
void amod_gpu_reset(void)
{
    int k;
    for (k = 0; k < 2; ++k)
        if (g_session[k]) { icw_session_destroy(g_session[k]); g_session[k] = NULL; }
    memset(g_fp_seen, 0, sizeof g_fp_seen);
    g_clips[0] = g_clips[1] = 0;
    g_peak[0] = g_peak[1] = 0.0;
}
