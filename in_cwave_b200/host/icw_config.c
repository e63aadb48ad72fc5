/* icw_config.c -- the reference's configuration file, read and written by the host C layer.
 *
 * A reference config (src/config.c:815-975) is a text file of "KEYWORD=value ..." lines; the DSP
 * list is one NODE_DSP line per node, head (the master) first (src/config.c:562-774).  This file
 * turns such a file into the icw_chain_spec / icwp_options pair the B200 library runs, with the
 * reference's own acceptance rules:
 *   - values are clamped to the bounds of src/config.c:118-207 and :690-750;
 *   - an unknown keyword, a line without '=' or with control characters, or a VER_CONFIG other than
 *     10 rejects the WHOLE file and leaves the defaults (src/config.c:906-913); a VALUE that does not
 *     parse is silently skipped (see the note in icwp_load_config);
 *   - the list is accepted only if its first node is the one and only master; the "lock" flags
 *     copy the left channel's parameters to the right one (amod_init, src/adv_modulator.c:216-331);
 *     a rejected list leaves the bare default master.
 * Doubles are stored as 0x<16 hex digits> of their bit pattern (src/config.c:547-560) and read
 * back either way (:520-541).  Strings escape blanks and '%' with '%' (:373-440).
 * Plain C99, no CUDA: the graph is validated again (feedback edges, plug ranges) by
 * icw_session_create / icw_session_set_spec behind the C ABI.
 */
#define _POSIX_C_SOURCE 200809L
#include "icw_plugin.h"

#include <ctype.h>
#include <inttypes.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>

#define CFG_VERSION      10u            /* src/config.c:50 */
#define CFG_LINE_MAX     2048           /* src/in_cwave.h:124 */
#define CFG_N_INPUTS     27             /* src/in_cwave.h:198 */
#define CFG_NAME_MAX     (80 + 16)      /* src/in_cwave.h:199-200 */

/* bounds, src/in_cwave.h:137-182, src/sound_render.h:56-110, src/lpf_hilbert_quad.h:70-87 */
#define MAX_GAIN 2.0
#define MAX_FSHIFT 20.0
#define MAX_PMFREQ 40.0
#define MAX_PMLEVEL 1.0

typedef struct cur { char *p; } cur;

/* next blank-separated token with the reference's '%' escapes undone; returns its start (in place) */
static char *token(cur *c)
{
    char *out, *start;
    while (*c->p && isblank((unsigned char)*c->p)) ++c->p;
    start = out = c->p;
    while (*c->p && !isblank((unsigned char)*c->p)) {
        if (*c->p == '%') {
            ++c->p;
            if (isblank((unsigned char)*c->p) || *c->p == '%') *out++ = *c->p++;
            /* '%' before anything else is dropped, the character itself is taken by the next turn */
        } else {
            *out++ = *c->p++;
        }
    }
    if (*c->p) ++c->p;
    *out = 0;
    return start;
}

static int get_int(cur *c, int *v)            { return sscanf(token(c), "%d", v) == 1; }
static int get_unsigned(cur *c, unsigned *v)  { return sscanf(token(c), "%u", v) == 1; }
static int get_bool(cur *c, int *v)           { int t; if (!get_int(c, &t)) return 0; *v = t ? 1 : 0; return 1; }
static int get_double(cur *c, double *v)
{
    const char *t = token(c);
    if (t[0] == '0' && (t[1] == 'x' || t[1] == 'X')) {
        uint64_t bits;
        if (sscanf(t + 2, "%" SCNx64, &bits) != 1) return 0;
        memcpy(v, &bits, sizeof bits);
        return 1;
    }
    return sscanf(t, "%lg", v) == 1;
}
static int clampi(int v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : v; }
static unsigned clampu(unsigned v, unsigned lo, unsigned hi) { return v < lo ? lo : v > hi ? hi : v; }
/* the reference's two comparisons (src/config.c:691): a NaN passes both and stays */
static double clampd(double v, double lo, double hi) { if (v < lo) v = lo; if (v > hi) v = hi; return v; }

/* one NODE_DSP line -> a node in FILE order, plus the lock flags amod_init acts on */
typedef struct file_node {
    icw_node n;
    int lock_gain, lock_shift, sign_lock_shift, lock_freq, lock_phase, lock_level, lock_angle;
} file_node;

static int parse_node(cur *c, file_node *fn)
{
    icw_node *n = &fn->n;
    int k, b;
    memset(fn, 0, sizeof *fn);
    (void)token(c);                                             /* the name: display only */
    if (!get_double(c, &n->l_gain)) return 0;
    n->l_gain = clampd(n->l_gain, 0.0, MAX_GAIN);
    if (!get_double(c, &n->r_gain)) return 0;
    n->r_gain = clampd(n->r_gain, 0.0, MAX_GAIN);
    if (!get_bool(c, &fn->lock_gain)) return 0;
    for (k = 0; k < CFG_N_INPUTS; ++k) {
        if (!get_bool(c, &b)) return 0;
        if (b) n->inputs_mask |= 1u << k;
    }
    if (!get_int(c, &n->xch_mode)) return 0;
    n->xch_mode = clampi(n->xch_mode, ICW_XCH_NORMAL, ICW_XCH_MIXLR);
    if (!get_bool(c, &n->l_iq_invert) || !get_bool(c, &n->r_iq_invert)) return 0;
    if (!get_int(c, &n->mode)) return 0;
    n->mode = clampi(n->mode, ICW_MODE_MASTER, ICW_MODE_MIX);
    switch (n->mode) {
    case ICW_MODE_MASTER:
        if (!get_int(c, &n->l_tout) || !get_int(c, &n->r_tout)) return 0;
        /* the reference clamps each right after reading it; a missing second value fails either way */
        n->l_tout = clampi(n->l_tout, ICW_OUT_ADD_REIM, ICW_OUT_IM);
        n->r_tout = clampi(n->r_tout, ICW_OUT_ADD_REIM, ICW_OUT_IM);
        break;
    case ICW_MODE_SHIFT:
        if (!get_double(c, &n->l_p[0])) return 0;
        n->l_p[0] = clampd(n->l_p[0], -MAX_FSHIFT, MAX_FSHIFT);
        if (!get_bool(c, &n->l_on)) return 0;
        if (!get_double(c, &n->r_p[0])) return 0;
        n->r_p[0] = clampd(n->r_p[0], -MAX_FSHIFT, MAX_FSHIFT);
        if (!get_bool(c, &n->r_on)) return 0;
        if (!get_int(c, &n->n_out)) return 0;
        n->n_out = clampi(n->n_out, 1, CFG_N_INPUTS);
        if (!get_bool(c, &fn->lock_shift) || !get_bool(c, &fn->sign_lock_shift)) return 0;
        break;
    case ICW_MODE_PM: {
        static const double lo[4] = { 0.0, -1.0, 0.0, -1.0 }, hi[4] = { MAX_PMFREQ, 1.0, MAX_PMLEVEL, 1.0 };
        for (k = 0; k < 4; ++k) {
            if (!get_double(c, &n->l_p[k])) return 0;
            n->l_p[k] = clampd(n->l_p[k], lo[k], hi[k]);
        }
        if (!get_bool(c, &n->l_on)) return 0;
        for (k = 0; k < 4; ++k) {
            if (!get_double(c, &n->r_p[k])) return 0;
            n->r_p[k] = clampd(n->r_p[k], lo[k], hi[k]);
        }
        if (!get_bool(c, &n->r_on)) return 0;
        if (!get_int(c, &n->n_out)) return 0;
        n->n_out = clampi(n->n_out, 1, CFG_N_INPUTS);
        if (!get_bool(c, &fn->lock_freq) || !get_bool(c, &fn->lock_phase) ||
            !get_bool(c, &fn->lock_level) || !get_bool(c, &fn->lock_angle)) return 0;
        break;
    }
    case ICW_MODE_MIX:
        if (!get_int(c, &n->n_out)) return 0;
        n->n_out = clampi(n->n_out, 1, CFG_N_INPUTS);
        break;
    default:
        return 0;
    }
    return 1;
}

/* what amod_init does to a list it is handed (src/adv_modulator.c:216-331); 0 = list refused */
static int adopt_list(file_node *fl, int n)
{
    int i, masters = 0;
    if (n < 1 || fl[0].n.mode != ICW_MODE_MASTER) return 0;
    for (i = 0; i < n; ++i) {
        icw_node *d = &fl[i].n;
        if (fl[i].lock_gain) { d->r_gain = d->l_gain; d->r_iq_invert = d->l_iq_invert; }
        switch (d->mode) {
        case ICW_MODE_MASTER:
            if (masters++) return 0;
            break;
        case ICW_MODE_SHIFT:
            if (fl[i].lock_shift) {
                d->r_p[0] = fl[i].sign_lock_shift ? -d->l_p[0] : d->l_p[0];
                d->r_on = d->l_on;
            }
            break;
        case ICW_MODE_PM:
            if (fl[i].lock_freq)  { d->r_p[0] = d->l_p[0]; d->r_on = d->l_on; }
            if (fl[i].lock_phase) d->r_p[1] = d->l_p[1];
            if (fl[i].lock_level) d->r_p[2] = d->l_p[2];
            if (fl[i].lock_angle) d->r_p[3] = d->l_p[3];
            break;
        default:
            break;
        }
    }
    return 1;
}

static void set_defaults(icw_chain_spec *chain, icwp_options *opt)
{
    const int device = opt->device;
    const int64_t ra = opt->readahead_frames;
    icw_default_spec(chain);
    memset(opt, 0, sizeof *opt);
    opt->device = device;                       /* not part of the reference config: kept */
    opt->readahead_frames = ra;
}

int icwp_load_config(const char *path, icw_chain_spec *chain, icwp_options *opt)
{
    FILE *fp;
    char line[CFG_LINE_MAX + 2];
    file_node *fl = NULL;
    int n_fl = 0, cap_fl = 0, ok = 1, i;
    unsigned version = 0, u;
    int b;
    double d;

    if (!path || !chain || !opt) return 0;
    set_defaults(chain, opt);
    fp = fopen(path, "r");
    if (!fp) return 0;

    while (ok) {
        size_t len = 0;
        int ch, blank = 1;
        char *eq;
        cur c;
        /* a line: '\r' dropped, tabs become blanks, other control characters are an error */
        while ((ch = fgetc(fp)) != EOF && ch != '\n') {
            if (ch == '\r') continue;
            if (len >= CFG_LINE_MAX) { ok = 0; break; }
            line[len++] = (char)(ch == '\t' ? ' ' : ch);
        }
        if (!ok) break;
        line[len] = 0;
        for (i = 0; i < (int)len; ++i) {
            if (iscntrl((unsigned char)line[i])) ok = 0;
            if (isgraph((unsigned char)line[i])) blank = 0;
        }
        if (!ok) break;
        if (blank) { if (ch == EOF) break; continue; }
        eq = strchr(line, '=');
        if (!eq) { ok = 0; break; }
        *eq = 0;
        c.p = line;
        {
            char *kw = token(&c);
            c.p = eq + 1;
#define KW(name) (strcasecmp(kw, name) == 0)
            /* A value that does not parse is NOT an error in the reference: its handler's verdict is
             * overwritten by the next line's read (src/config.c:846-903), the item just keeps its default
             * (a NODE_DSP line: the node is dropped).  Only an unknown keyword, a line without '=', a control
             * character or an over-long line reject the file. */
            if      (KW("VER_CONFIG"))    { if (get_unsigned(&c, &u)) version = u; }
            else if (KW("WAV_SUPPORT") || KW("RWAVE_SUPPORT") || KW("LAST_CHANCE") || KW("DISABLE_SLEEP") ||
                     KW("SHOW_LONGNUMB") || KW("IBOX_PARENT") || KW("PLAY_SLEEP") ||
                     KW("IIR_SUBN_THR"))
                                          { /* GUI / player plumbing, or never reaches the arithmetic (src/hblpf.c:915) */ }
            else if (KW("FP_CHECK"))      { if (get_bool(&c, &b)) chain->is_fp_check = b; }      /* src/config.c:180 */
            else if (KW("SEC_ALIGN"))     { if (get_unsigned(&c, &u)) opt->sec_align = clampu(u, 0, 20); }
            else if (KW("FADE_IN"))       { if (get_unsigned(&c, &u)) opt->fade_in_ms = clampu(u, 0, 10000); }
            else if (KW("FADE_OUT"))      { if (get_unsigned(&c, &u)) opt->fade_out_ms = clampu(u, 0, 10000); }
            else if (KW("FRMOD_SCALED"))  { if (get_bool(&c, &b)) chain->is_frmod_scaled = b; }
            else if (KW("IIR_HBLPF_IX"))  { if (get_unsigned(&c, &u)) chain->filter_no = (int)clampu(u, 0, 5); }
            else if (KW("IIR_SUM_KAHAN")) { if (get_bool(&c, &b)) chain->is_kahan = b; }
            else if (KW("IIR_SUBN_ZERO")) { if (get_bool(&c, &b)) chain->is_subnorm_reject = b; }
            else if (KW("CLR_NFRAME_PT")) { if (get_bool(&c, &b)) opt->clr_nframe_trk = b; }
            else if (KW("CLR_HILB_PT"))   { if (get_bool(&c, &b)) opt->clr_hilb_trk = b; }
            else if (KW("NEED24BITS"))    { if (get_bool(&c, &b)) chain->need24bits = b; }
            else if (KW("DITHER_BITS"))   { if (get_double(&c, &d)) chain->dth_bits = clampd(d, 0.0, 23.0); }
            else if (KW("QUANTIZE_TYPE")) { if (get_unsigned(&c, &u)) chain->quantz_type = clampu(u, 0, 1); }
            else if (KW("RENDER_TYPE"))   { if (get_unsigned(&c, &u)) chain->render_type = clampu(u, 0, 4); }
            else if (KW("NOISE_SHAPING")) { if (get_unsigned(&c, &u)) chain->nshape_type = clampu(u, 0, 17); }
            else if (KW("SIGNBITS16"))    { if (get_unsigned(&c, &u)) chain->sign_bits16 = clampu(u, 2, 16); }
            else if (KW("SIGNBITS24"))    { if (get_unsigned(&c, &u)) chain->sign_bits24 = clampu(u, 2, 24); }
            else if (KW("NODE_DSP")) {
                if (n_fl == cap_fl) {
                    file_node *t = realloc(fl, (size_t)(cap_fl = cap_fl ? 2 * cap_fl : 8) * sizeof *fl);
                    if (!t) { ok = 0; break; }
                    fl = t;
                }
                if (parse_node(&c, &fl[n_fl])) ++n_fl;
            }
            else ok = 0;                                                  /* unknown keyword */
#undef KW
        }
        if (ch == EOF) break;
    }
    fclose(fp);

    if (!ok || version != CFG_VERSION) {
        free(fl);
        set_defaults(chain, opt);
        return 0;
    }
    if (n_fl > ICW_MAX_NODES) {                 /* the reference has no limit; this library does: refuse loudly */
        free(fl);
        set_defaults(chain, opt);
        return 0;
    }
    /* the list: head first in the file, tail first in execution order */
    if (n_fl > 0 && adopt_list(fl, n_fl)) {
        for (i = 0; i < n_fl; ++i) chain->nodes[i] = fl[n_fl - 1 - i].n;
        chain->n_nodes = n_fl;
    }                                                                     /* else: the default master stays */
    free(fl);
    return 1;
}

/* ---- writing (src/config.c:924-975; value formats :373-560) ------------------------------------------- */
static void put_double_bin(FILE *fp, double v)
{
    uint64_t bits;
    memcpy(&bits, &v, sizeof bits);
    fprintf(fp, " 0x%08" PRIX64, bits);
}

int icwp_save_config(const char *path, const icw_chain_spec *chain, const icwp_options *opt)
{
    static const char *names[4] = { "Master", "Shift-", "PM-", "MIX-" };   /* src/adv_modulator.c:62-68 */
    FILE *fp;
    int i, k;
    if (!path || !chain || !opt || chain->n_nodes < 1 || chain->n_nodes > ICW_MAX_NODES) return 0;
    if (chain->nodes[chain->n_nodes - 1].mode != ICW_MODE_MASTER) return 0;
    fp = fopen(path, "w");
    if (!fp) return 0;
    fprintf(fp, "VER_CONFIG=%u\n", CFG_VERSION);
    fprintf(fp, "WAV_SUPPORT=1\nRWAVE_SUPPORT=0\nIBOX_PARENT=1\nLAST_CHANCE=0\nPLAY_SLEEP=10\nDISABLE_SLEEP=0\n");
    fprintf(fp, "SEC_ALIGN=%u\nFADE_IN=%u\nFADE_OUT=%u\n", opt->sec_align, opt->fade_in_ms, opt->fade_out_ms);
    fprintf(fp, "FRMOD_SCALED=%d\nIIR_HBLPF_IX=%u\nIIR_SUM_KAHAN=%d\nIIR_SUBN_ZERO=%d\n", chain->is_frmod_scaled ? 1 : 0,
            (unsigned)chain->filter_no, chain->is_kahan ? 1 : 0, chain->is_subnorm_reject ? 1 : 0);
    fprintf(fp, "IIR_SUBN_THR="); { uint64_t b; double t = 1.0e-150; memcpy(&b, &t, 8); fprintf(fp, "0x%08" PRIX64 "\n", b); }
    fprintf(fp, "CLR_NFRAME_PT=%d\nCLR_HILB_PT=%d\nSHOW_LONGNUMB=0\nFP_CHECK=%d\nNEED24BITS=%d\n",
            opt->clr_nframe_trk ? 1 : 0, opt->clr_hilb_trk ? 1 : 0, chain->is_fp_check ? 1 : 0, chain->need24bits ? 1 : 0);
    fprintf(fp, "DITHER_BITS="); { uint64_t b; memcpy(&b, &chain->dth_bits, 8); fprintf(fp, "0x%08" PRIX64 "\n", b); }
    fprintf(fp, "QUANTIZE_TYPE=%u\nRENDER_TYPE=%u\nNOISE_SHAPING=%u\nSIGNBITS16=%u\nSIGNBITS24=%u\n",
            chain->quantz_type, chain->render_type, chain->nshape_type, chain->sign_bits16, chain->sign_bits24);
    for (i = chain->n_nodes - 1; i >= 0; --i) {                           /* head (master) first */
        const icw_node *n = &chain->nodes[i];
        if (n->mode < ICW_MODE_MASTER || n->mode > ICW_MODE_MIX) { fclose(fp); return 0; }
        fprintf(fp, "NODE_DSP=%s", names[n->mode]);
        if (n->mode != ICW_MODE_MASTER) fprintf(fp, "%d", chain->n_nodes - 1 - i);
        put_double_bin(fp, n->l_gain);
        put_double_bin(fp, n->r_gain);
        fprintf(fp, " 0");                                                /* lock_gain */
        for (k = 0; k < CFG_N_INPUTS; ++k) fprintf(fp, " %d", (int)((n->inputs_mask >> k) & 1u));
        fprintf(fp, " %d %d %d %d", n->xch_mode, n->l_iq_invert ? 1 : 0, n->r_iq_invert ? 1 : 0, n->mode);
        switch (n->mode) {
        case ICW_MODE_MASTER:
            fprintf(fp, " %d %d", n->l_tout, n->r_tout);
            break;
        case ICW_MODE_SHIFT:
            put_double_bin(fp, n->l_p[0]); fprintf(fp, " %d", n->l_on ? 1 : 0);
            put_double_bin(fp, n->r_p[0]); fprintf(fp, " %d", n->r_on ? 1 : 0);
            fprintf(fp, " %d 0 0", n->n_out);
            break;
        case ICW_MODE_PM:
            for (k = 0; k < 4; ++k) put_double_bin(fp, n->l_p[k]);
            fprintf(fp, " %d", n->l_on ? 1 : 0);
            for (k = 0; k < 4; ++k) put_double_bin(fp, n->r_p[k]);
            fprintf(fp, " %d", n->r_on ? 1 : 0);
            fprintf(fp, " %d 0 0 0 0", n->n_out);
            break;
        default:
            fprintf(fp, " %d", n->n_out);
            break;
        }
        fputc('\n', fp);
    }
    k = ferror(fp);
    return fclose(fp) == 0 && !k;
}
