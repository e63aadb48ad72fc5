/* icw_oracle.c -- CPU restatement ("port" oracle) of the in_cwave signal chain.
 *
 * TEST INFRASTRUCTURE ONLY -- see icw_oracle.h.  Parity status: PINNED against the compiled
 * reference (oracle/_ref) and tests/golden/ by tests/test_oracle_vs_ref.py.
 *
 * This is a restatement, not a copy: the reference evaluates one frame at a time through
 * function pointers, linked lists and process-global state; here every stage runs over a whole
 * block with explicit state, which is equivalent because the stages only couple through the
 * per-frame values they hand on (unpack -> Hilbert -> graph -> render) and each stage's own
 * carried state.  What IS kept, operation for operation, is the floating-point arithmetic:
 * every product, sum and comparison below happens in the order the reference performs it, in
 * plain IEEE-754 double (build with -ffp-contract=off, no fast-math), because the half-band
 * recurrences amplify any re-association far above 1 ulp (finding 3).
 */
#include "icw_oracle.h"

#include <math.h>
#include <string.h>

#include "icwo_hb_tables.inc"
#include "icwo_ns_tables.inc"

#define ICWO_PI     3.1415926535897932384626433832795029   /* reference src/in_cwave.h:148 */
#define ICWO_SQRT2  1.4142135623730950488016887242097      /* reference src/adv_modulator.c:43 */
#define ICWO_SQRT6  2.4494897427831780981972840747059      /* reference src/sound_render.c:50 */
#define ICWO_HZ_SCALE 1000u                                /* reference src/in_cwave.h:162 */
#define ICWO_SILENCE_DB (-555.0)                           /* reference src/sound_render.h:103 */

/* ------------------------------------------------------------------------------------------
 * MT19937 (reference src/mersene_twister/mt_jrnd.c:28-134, 218-256)
 * ---------------------------------------------------------------------------------------- */
void icwo_mt_seed(icwo_mt *mt, uint32_t seed)
{
    /* Knuth-style linear fill, mt_jrnd.c:28-47 */
    uint32_t prev = seed;
    mt->w[0] = prev;
    for (uint32_t j = 1; j < ICWO_MT_N; ++j) {
        prev = 1812433253u * (prev ^ (prev >> 30)) + j;
        mt->w[j] = prev;
    }
    mt->pos = ICWO_MT_N;
    mt->drawn = 0;
}

void icwo_mt_seed_key(icwo_mt *mt, const uint32_t *key, uint32_t key_len)
{
    /* array seeding, mt_jrnd.c:51-97; only used by the reference's known-answer test */
    uint32_t i = 1, j = 0;
    uint32_t rounds = ICWO_MT_N > key_len ? ICWO_MT_N : key_len;

    icwo_mt_seed(mt, 19650218u);
    for (; rounds; --rounds) {
        uint32_t p = mt->w[i - 1];
        mt->w[i] = (mt->w[i] ^ ((p ^ (p >> 30)) * 1664525u)) + key[j] + j;
        if (++i >= ICWO_MT_N) { mt->w[0] = mt->w[ICWO_MT_N - 1]; i = 1; }
        if (++j >= key_len) j = 0;
    }
    for (rounds = ICWO_MT_N - 1; rounds; --rounds) {
        uint32_t p = mt->w[i - 1];
        mt->w[i] = (mt->w[i] ^ ((p ^ (p >> 30)) * 1566083941u)) - i;
        if (++i >= ICWO_MT_N) { mt->w[0] = mt->w[ICWO_MT_N - 1]; i = 1; }
    }
    mt->w[0] = 0x80000000u;
    mt->pos = ICWO_MT_N;
    mt->drawn = 0;
}

static void mt_regenerate(uint32_t *w)
{
    /* mt_jrnd.c:103-122: w[k] <- w[k+397] ^ twist(w[k], w[k+1]), indices mod 624, in place */
    for (int k = 0; k < ICWO_MT_N; ++k) {
        uint32_t a = w[k];
        uint32_t b = w[k + 1 < ICWO_MT_N ? k + 1 : 0];
        uint32_t mix = (a & 0x80000000u) | (b & 0x7FFFFFFFu);
        uint32_t tw = (mix >> 1) ^ ((b & 1u) ? 0x9908B0DFu : 0u);
        w[k] = w[k + 397 < ICWO_MT_N ? k + 397 : k + 397 - ICWO_MT_N] ^ tw;
    }
}

/* Test hook: a draw lands on the rejection loop of mtrnd_gen_dsopen (mt_jrnd.c:249-253) with probability 2^-53, and no
 * seed is known that gets there.  A patch replaces ONE output word of a generator's stream: word number `idx` (counted
 * from the seeding), if it comes out as `match` (which tells the two channels' generators apart), is handed out as
 * `value` instead.  Two patched words (0, 0) make a pair that the loop below rejects.  tests/ only. */
static struct { uint64_t idx; uint32_t match, value; } g_patch[16];
static int g_npatch;
void icwo_debug_patch_word(uint64_t idx, uint32_t match, uint32_t value)
{
    if (g_npatch < 16) { g_patch[g_npatch].idx = idx; g_patch[g_npatch].match = match; g_patch[g_npatch].value = value; ++g_npatch; }
}
void icwo_debug_clear_patches(void) { g_npatch = 0; }

uint32_t icwo_mt_u32(icwo_mt *mt)
{
    if (mt->pos >= ICWO_MT_N) {
        mt_regenerate(mt->w);
        mt->pos = 0;
    }
    uint32_t y = mt->w[mt->pos++];
    mt->drawn++;
    /* tempering, mt_jrnd.c:127-131 */
    y ^= y >> 11;
    y ^= (y << 7) & 0x9D2C5680u;
    y ^= (y << 15) & 0xEFC60000u;
    y ^= y >> 18;
    for (int k = 0; k < g_npatch; ++k)
        if (g_patch[k].idx == mt->drawn - 1 && g_patch[k].match == y) return g_patch[k].value;
    return y;
}

double icwo_mt_dsopen(icwo_mt *mt)
{
    /* mt_jrnd.c:218-226 (53-bit [0,1)) then :245-256 (map to (-1,1), redraw on +-1) */
    double r;
    do {
        uint32_t hi = icwo_mt_u32(mt) >> 5;
        uint32_t lo = icwo_mt_u32(mt) >> 6;
        double u = ((double)hi * 67108864.0 + (double)lo) * (1.0 / 9007199254740992.0);
        r = u * 2.0 - 1.0;
    } while (r == -1.0 || r == 1.0);
    return r;
}

/* ------------------------------------------------------------------------------------------
 * sample unpacking (reference src/unpack_lsb.h:53-125, src/xwave_reader.c:171-239)
 * ---------------------------------------------------------------------------------------- */
static inline int32_t rd_i16(const uint8_t *p) { return (int16_t)((uint16_t)p[0] | ((uint16_t)p[1] << 8)); }
static inline int32_t rd_i24(const uint8_t *p)
{
    uint32_t u = ((uint32_t)p[0] << 8) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 24);
    return (int32_t)u >> 8;
}
static inline uint32_t rd_u32(const uint8_t *p)
{
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
static inline double rd_f32(const uint8_t *p) { uint32_t u = rd_u32(p); float f; memcpy(&f, &u, 4); return (double)f; }
static inline double rd_f64(const uint8_t *p)
{
    uint64_t u = (uint64_t)rd_u32(p) | ((uint64_t)rd_u32(p + 4) << 32);
    double d; memcpy(&d, &u, 8); return d;
}

static int chan_bytes(int fmt)
{
    switch (fmt) {
    case ICWO_FMT_WAV_U8:  return 1;
    case ICWO_FMT_WAV_I16: return 2;
    case ICWO_FMT_WAV_I24: return 3;
    case ICWO_FMT_WAV_I32: return 4;
    case ICWO_FMT_WAV_F32: return 4;
    case ICWO_FMT_CW_F64:    return 16;
    case ICWO_FMT_CW_I16:    return 4;
    case ICWO_FMT_CW_I16F32: return 6;
    case ICWO_FMT_CW_F32:    return 8;
    default: return -1;
    }
}

/* real sample -> +-32768 units (xwave_reader.c:203-239) */
static double real_value(int fmt, const uint8_t *p)
{
    switch (fmt) {
    case ICWO_FMT_WAV_U8:  return 256.0 * (double)(int8_t)(uint8_t)(p[0] - 0x80u);
    case ICWO_FMT_WAV_I16: return (double)rd_i16(p);
    case ICWO_FMT_WAV_I24: return (double)rd_i24(p) / 256.0;
    case ICWO_FMT_WAV_I32: return (double)(int32_t)rd_u32(p) / 65536.0;
    default:               return 32768.0 * rd_f32(p);
    }
}

/* complex sample -> (I, Q) (xwave_reader.c:171-199) */
static void iq_value(int fmt, const uint8_t *p, double *vi, double *vq)
{
    switch (fmt) {
    case ICWO_FMT_CW_F64:    *vi = rd_f64(p);          *vq = rd_f64(p + 8);          break;
    case ICWO_FMT_CW_I16:    *vi = (double)rd_i16(p);  *vq = (double)rd_i16(p + 2);  break;
    case ICWO_FMT_CW_I16F32: *vi = (double)rd_i16(p);  *vq = rd_f32(p + 2);          break;
    default:                 *vi = rd_f32(p);          *vq = rd_f32(p + 4);          break;
    }
}

/* out4[n][4]: complex input -> (L.I, L.Q, R.I, R.Q); real input -> (L, 0, R, 0), mono copied */
void icwo_unpack(int fmt, int n_channels, const uint8_t *in, int64_t n, double *out4)
{
    int cb = chan_bytes(fmt);
    for (int64_t k = 0; k < n; ++k) {
        const uint8_t *p = in + k * (int64_t)cb * n_channels;
        double *o = out4 + 4 * k;
        if (fmt >= ICWO_FMT_CW_F64) {
            iq_value(fmt, p, &o[0], &o[1]);
            if (n_channels > 1) iq_value(fmt, p + cb, &o[2], &o[3]);
            else { o[2] = o[0]; o[3] = o[1]; }
        } else {
            o[0] = real_value(fmt, p);
            o[2] = n_channels > 1 ? real_value(fmt, p + cb) : o[0];
            o[1] = o[3] = 0.0;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * half-band elliptic LPF, direct form II (reference src/hblpf.c:828-1057)
 * ---------------------------------------------------------------------------------------- */
typedef struct hb_design { int ord; double fb[ICWO_MAX_ORD], ff[ICWO_MAX_ORD], d0; } hb_design;

static double word_as_double(unsigned long long w) { double d; memcpy(&d, &w, 8); return d; }

static void hb_load(int type, hb_design *h)
{
    /* hblpf.c:849-856: fb = -a[j]/a0, ff = b[j]/a0, d0 = b0/a0 */
    double a0 = word_as_double(ICW_HB_A[type][0]);
    h->ord = ICW_HB_ORDER[type];
    h->d0 = word_as_double(ICW_HB_B[type][0]) / a0;
    for (int i = 0; i < h->ord; ++i) {
        h->fb[i] = -word_as_double(ICW_HB_A[type][i + 1]) / a0;
        h->ff[i] =  word_as_double(ICW_HB_B[type][i + 1]) / a0;
    }
}

/* ---- FP-exception-checked twins (reference src/fp_check.c:52-99; used when cfg.is_fp_check) ----
 * FC(v): NaN and denormals become 0.0, +-Inf becomes +-65535.0, each event counted by class.  The
 * counter block in force is file-scope state of this (single-threaded, test-only) oracle: NULL = the
 * unchecked code paths. */
static unsigned *g_fpc = NULL;
static double fc(double v)
{
    unsigned *c = g_fpc;
    switch (fpclassify(v)) {
    case FP_NAN:       c[0]++; c[2]++; return 0.0;              /* every NaN counts as quiet: arithmetic results are */
    case FP_SUBNORMAL: c[0]++; c[signbit(v) ? 4 : 5]++; return 0.0;
    case FP_INFINITE:  c[0]++; if (signbit(v)) { c[3]++; return -65535.0; } c[6]++; return 65535.0;
    default:           return v;
    }
}

/* the "reject" test compares |w| with the FLAG VALUE, not with the configured threshold
 * (hblpf.c:915,1046; SURVEY.md finding 4) */
static inline double hb_reject(double w, int flag, uint64_t *count)
{
    if (flag && fabs(w) < (double)flag) { ++*count; return 0.0; }
    return w;
}

static double hb_step_plain(const hb_design *h, icwo_iir *f, double x, int flag)
{
    /* hblpf.c:894-926 */
    double acc_in = x, acc_out = 0.0;
    int k = f->ix;
    for (int i = 0; i < h->ord; ++i) {
        k = (k == 0 ? h->ord : k) - 1;          /* newest state first */
        acc_in  += f->z[k] * h->fb[i];
        acc_out += f->z[k] * h->ff[i];
    }
    acc_in = hb_reject(acc_in, flag, &f->rejects);
    f->z[f->ix] = acc_in;
    if (++f->ix >= h->ord) f->ix = 0;
    return acc_in * h->d0 + acc_out;
}

typedef struct comp_sum { double s, c; } comp_sum;
static inline void comp_add(comp_sum *a, double x)
{
    /* hblpf.c:991-997; safe under -ffp-contract=off without fast-math: gcc keeps (t - s) - y */
    double y = x - a->c;
    double t = a->s + y;
    a->c = (t - a->s) - y;
    a->s = t;
}

static double hb_step_kahan(const hb_design *h, icwo_iir *f, double x, int flag)
{
    /* hblpf.c:1008-1057.  The output is the compensated sum of z*ff[i] and (z*fb[i])*d0 only:
     * the direct term d0*x is never added (SURVEY.md finding 4). */
    comp_sum in = { x, 0.0 }, out;
    int k = f->ix;

    k = (k == 0 ? h->ord : k) - 1;
    double t = f->z[k] * h->fb[0];
    comp_add(&in, t);
    out.s = f->z[k] * h->ff[0];
    out.c = 0.0;
    comp_add(&out, t * h->d0);
    for (int i = 1; i < h->ord; ++i) {
        k = (k == 0 ? h->ord : k) - 1;
        t = f->z[k] * h->fb[i];
        comp_add(&in, t);
        comp_add(&out, f->z[k] * h->ff[i]);
        comp_add(&out, t * h->d0);
    }
    in.s = hb_reject(in.s, flag, &f->rejects);
    f->z[f->ix] = in.s;
    if (++f->ix >= h->ord) f->ix = 0;
    return out.s;
}

static double hb_step_plain_fc(const hb_design *h, icwo_iir *f, double x, int flag)
{
    /* hblpf.c:928-950 */
    double acc_in = x, acc_out = 0.0;
    int k = f->ix;
    for (int i = 0; i < h->ord; ++i) {
        k = (k == 0 ? h->ord : k) - 1;
        acc_in  = fc(acc_in  + fc(f->z[k] * h->fb[i]));
        acc_out = fc(acc_out + fc(f->z[k] * h->ff[i]));
    }
    acc_in = hb_reject(acc_in, flag, &f->rejects);
    f->z[f->ix] = acc_in;
    if (++f->ix >= h->ord) f->ix = 0;
    return fc(fc(acc_in * h->d0) + acc_out);
}

static inline void comp_add_fc(comp_sum *a, double x)
{
    /* hblpf.c:995-1003 */
    double y = fc(x - a->c);
    double t = fc(a->s + y);
    a->c = fc(fc(t - a->s) - y);
    a->s = t;
}

static double hb_step_kahan_fc(const hb_design *h, icwo_iir *f, double x, int flag)
{
    /* hblpf.c:1059-1096 */
    comp_sum in = { x, 0.0 }, out;
    int k = f->ix;

    k = (k == 0 ? h->ord : k) - 1;
    double t = fc(f->z[k] * h->fb[0]);
    comp_add_fc(&in, t);
    out.s = fc(f->z[k] * h->ff[0]);
    out.c = 0.0;
    comp_add_fc(&out, fc(t * h->d0));
    for (int i = 1; i < h->ord; ++i) {
        k = (k == 0 ? h->ord : k) - 1;
        t = fc(f->z[k] * h->fb[i]);
        comp_add_fc(&in, t);
        comp_add_fc(&out, fc(f->z[k] * h->ff[i]));
        comp_add_fc(&out, fc(t * h->d0));
    }
    in.s = hb_reject(in.s, flag, &f->rejects);
    f->z[f->ix] = in.s;
    if (++f->ix >= h->ord) f->ix = 0;
    return out.s;
}

static inline double hb_step(const hb_design *h, icwo_iir *f, double x, int kahan, int flag)
{
    if (g_fpc) return kahan ? hb_step_kahan_fc(h, f, x, flag) : hb_step_plain_fc(h, f, x, flag);
    return kahan ? hb_step_kahan(h, f, x, flag) : hb_step_plain(h, f, x, flag);
}

void icwo_iir_run(int filter_no, int is_kahan, int is_reject, icwo_iir *f,
                  const double *x, int64_t n, double *y)
{
    hb_design h;
    hb_load(filter_no, &h);
    for (int64_t k = 0; k < n; ++k)
        y[k] = hb_step(&h, f, x[k], is_kahan, is_reject);
}

/* real -> analytic, one channel (reference src/lpf_hilbert_quad.c:129-156): mix down by fs/4,
 * low-pass I and Q, mix back up, times two.  lpf[0] is the I filter, lpf[1] the Q filter. */
void icwo_hilbert(int filter_no, int is_kahan, int is_reject, icwo_iir lpf[2], unsigned *quad,
                  const double *x, int64_t n, double *out_i, double *out_q)
{
    hb_design h;
    hb_load(filter_no, &h);
    unsigned q = *quad;
    for (int64_t k = 0; k < n; ++k, q = (q + 1) & 3u) {
        double a, b;
        switch (q) {
        case 0:
            a = hb_step(&h, &lpf[0],  x[k], is_kahan, is_reject);
            b = hb_step(&h, &lpf[1],  0.0,  is_kahan, is_reject);
            out_i[k] =  a * 2.0;  out_q[k] =  b * 2.0;
            break;
        case 1:
            a = hb_step(&h, &lpf[1], -x[k], is_kahan, is_reject);
            b = hb_step(&h, &lpf[0],  0.0,  is_kahan, is_reject);
            out_i[k] = -a * 2.0;  out_q[k] =  b * 2.0;
            break;
        case 2:
            a = hb_step(&h, &lpf[0], -x[k], is_kahan, is_reject);
            b = hb_step(&h, &lpf[1],  0.0,  is_kahan, is_reject);
            out_i[k] = -a * 2.0;  out_q[k] = -b * 2.0;
            break;
        default:
            a = hb_step(&h, &lpf[1],  x[k], is_kahan, is_reject);
            b = hb_step(&h, &lpf[0],  0.0,  is_kahan, is_reject);
            out_i[k] =  a * 2.0;  out_q[k] = -b * 2.0;
            break;
        }
    }
    *quad = q;
}

/* ------------------------------------------------------------------------------------------
 * "truth": the same converter with every recurrence evaluated in IEEE binary128 (libquadmath),
 * i.e. the exact-arithmetic value of the reference's filter with its rounded coefficients, to
 * ~1e-30.  It is what the reference's FP64 evaluation approximates (to 1e-9 .. 1e-3 of RMS,
 * depending on the design) and what the GPU scan mode is held to (1e-12).  The |w|<1 zeroing
 * is not modelled (it only acts on all-zero / fully decayed state).  drop_direct selects the
 * Kahan path's output, which omits the d0*x term (hblpf.c:1056).
 * ---------------------------------------------------------------------------------------- */
typedef struct hbq_filter { __float128 z[ICWO_MAX_ORD]; } hbq_filter;

static double hbq_step(const hb_design *h, hbq_filter *f, double x, int drop_direct)
{
    __float128 w = x, o = 0;
    for (int i = 0; i < h->ord; ++i) {
        w += f->z[i] * (__float128)h->fb[i];
        o += f->z[i] * (__float128)h->ff[i];
    }
    for (int i = h->ord - 1; i > 0; --i) f->z[i] = f->z[i - 1];
    f->z[0] = w;
    __float128 y = w * (__float128)h->d0 + o;
    if (drop_direct) y -= (__float128)h->d0 * (__float128)x;
    return (double)y;
}

void icwo_hilbert_truth(int filter_no, int drop_direct, unsigned quad0, const double *x, int64_t n,
                        double *out_i, double *out_q)
{
    hb_design h;
    hbq_filter fi, fq;
    hb_load(filter_no, &h);
    memset(&fi, 0, sizeof fi);
    memset(&fq, 0, sizeof fq);
    unsigned q = quad0 & 3u;
    for (int64_t k = 0; k < n; ++k, q = (q + 1) & 3u) {
        double a, b;
        switch (q) {
        case 0:  a = hbq_step(&h, &fi,  x[k], drop_direct); b = hbq_step(&h, &fq, 0.0, drop_direct); out_i[k] =  a * 2.0; out_q[k] =  b * 2.0; break;
        case 1:  a = hbq_step(&h, &fq, -x[k], drop_direct); b = hbq_step(&h, &fi, 0.0, drop_direct); out_i[k] = -a * 2.0; out_q[k] =  b * 2.0; break;
        case 2:  a = hbq_step(&h, &fi, -x[k], drop_direct); b = hbq_step(&h, &fq, 0.0, drop_direct); out_i[k] = -a * 2.0; out_q[k] = -b * 2.0; break;
        default: a = hbq_step(&h, &fq,  x[k], drop_direct); b = hbq_step(&h, &fi, 0.0, drop_direct); out_i[k] =  a * 2.0; out_q[k] = -b * 2.0; break;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * renderer (reference src/sound_render.c:499-581 setup, :691-810 per value)
 * ---------------------------------------------------------------------------------------- */
typedef struct quant_plan {
    double dth_mul, hi, lo, norm_mul, round_off;
    int    neg_delta, shift, bytes;
} quant_plan;

static void quant_setup(const icwo_spec *sp, quant_plan *q)
{
    int64_t top;
    q->dth_mul = pow(2.0, sp->dth_bits) - 1.0;
    if (sp->quantz_type == 0) { q->round_off = 0.5; q->neg_delta = 0; }     /* mid tread */
    else                      { q->round_off = 0.0; q->neg_delta = -1; }    /* mid riser */
    if (sp->need24bits) {
        q->bytes = 3;
        q->shift = 24 - (int)sp->sign_bits24;
        top = 0x800000LL >> q->shift;
        q->norm_mul = q->shift < 8 ? (double)(0x100 >> q->shift)
                                   : 1.0 / (double)(1ULL << (q->shift - 8));
    } else {
        q->bytes = 2;
        q->shift = 16 - (int)sp->sign_bits16;
        top = 0x8000LL >> q->shift;
        q->norm_mul = 1.0 / (double)(1ULL << q->shift);
    }
    q->hi = (double)top;
    q->lo = -(double)(top + 1 + q->neg_delta);
    q->lo -= (double)q->neg_delta;
}

static double dither_draw(unsigned type, icwo_mt *mt, double *prev)
{
    /* sound_render.c:711-751 */
    double r, t;
    switch (type) {
    case ICWO_DITHER_RPDF:
        return icwo_mt_dsopen(mt) / ICWO_SQRT2;
    case ICWO_DITHER_TPDF:
        r = icwo_mt_dsopen(mt);
        r += icwo_mt_dsopen(mt);
        return r / 2.0;
    case ICWO_DITHER_STPDF:
        t = icwo_mt_dsopen(mt);
        r = (t - *prev) / 2.0;
        *prev = t;
        return r;
    case ICWO_DITHER_GAUSS:
        r = icwo_mt_dsopen(mt);
        for (int j = 1; j < 12; ++j) r += icwo_mt_dsopen(mt);
        return r / (2.0 * ICWO_SQRT6);
    default:
        return 0.0;
    }
}

/* noise shaping: the error of this sample through the shaper -> what the NEXT sample subtracts
 * (src/sound_render.c:403-489).  FIR: sum c_i * err[i samples ago], newest first, plain mul + add.
 * IIR: sum (c_i * err_i - d_i * out_i), then the sum joins the outputs. */
static double ns_step(unsigned type, icwo_ns *ns, double err)
{
    const int kind = ICW_NS_KIND[type], ord = ICW_NS_ORDER[type];
    const double *c = (const double *)ICW_NS_COEF[type];
    double res = 0.0;
    if (!kind) return 0.0;
    for (int i = ord - 1; i > 0; --i) ns->e[i] = ns->e[i - 1];
    ns->e[0] = err;
    if (kind == 1) {
        for (int i = 0; i < ord; ++i) res += c[i] * ns->e[i];
    } else {
        for (int i = 0; i < ord; ++i) res += c[i] * ns->e[i] - c[i + ord] * ns->o[i];
        for (int i = ord - 1; i > 0; --i) ns->o[i] = ns->o[i - 1];
        ns->o[0] = res;
    }
    return res;
}

static double ns_step_fc(unsigned type, icwo_ns *ns, double err)
{
    /* src/sound_render.c:415-441 (FIR), :458-489 (IIR), checked halves */
    const int kind = ICW_NS_KIND[type], ord = ICW_NS_ORDER[type];
    const double *c = (const double *)ICW_NS_COEF[type];
    double res = 0.0;
    if (!kind) return 0.0;
    for (int i = ord - 1; i > 0; --i) ns->e[i] = ns->e[i - 1];
    ns->e[0] = fc(err);
    if (kind == 1) {
        for (int i = 0; i < ord; ++i) res = fc(res + fc(c[i] * ns->e[i]));
    } else {
        for (int i = 0; i < ord; ++i) res = fc(res + fc(fc(c[i] * ns->e[i]) - fc(c[i + ord] * ns->o[i])));
        for (int i = ord - 1; i > 0; --i) ns->o[i] = ns->o[i - 1];
        ns->o[0] = res;
    }
    return res;
}

int64_t icwo_render(const icwo_spec *sp, icwo_mt *mt, double *prev_rnd, icwo_ns *ns, const double *in, int64_t n,
                    uint8_t *out, unsigned *clips, double *peak_db)
{
    quant_plan q;
    uint8_t *p = out;
    const unsigned nst = sp->nshape_type < ICW_NS_COUNT ? sp->nshape_type : 0;   /* src/sound_render.c:546 */
    quant_setup(sp, &q);
    for (int64_t k = 0; k < n; ++k) {
        /* the dither draws are sums of a few values in (-1, 1) on a 2^-52 grid: FC() can never fire on
         * them (src/sound_render.c:821-845), so the checked twin shares dither_draw() */
        double rnd = dither_draw(sp->render_type, mt, prev_rnd);
        double v, qv;
        int delta;
        if (g_fpc) {                                    /* src/sound_render.c:848-861 */
            v = fc(fc(in[k] * q.norm_mul) - (nst ? ns->prev_err : 0.0));
            qv = fc(v + fc(rnd * q.dth_mul));
            if (qv < 0.0) { qv = fc(qv - q.round_off); delta = q.neg_delta; }
            else          { qv = fc(qv + q.round_off); delta = 0; }
        } else {
            v = in[k] * q.norm_mul - (nst ? ns->prev_err : 0.0);
            qv = v + rnd * q.dth_mul;
            if (qv < 0.0) { qv -= q.round_off; delta = q.neg_delta; }
            else          { qv += q.round_off; delta = 0; }
        }

        double lvl = fabs(qv) / q.hi;                   /* peak: after rounding offset, before clip */
        lvl = lvl ? 20.0 * log10(lvl) : ICWO_SILENCE_DB;
        if (lvl > *peak_db) *peak_db = lvl;

        if (qv >= q.hi) { qv = q.hi - 1.0; ++*clips; }
        if (qv <= q.lo) { qv = q.lo + 1.0; ++*clips; }

        int val = (int)qv + delta;
        if (nst) ns->prev_err = g_fpc ? ns_step_fc(nst, ns, fc((double)val - v))   /* src/sound_render.c:897 */
                                      : ns_step(nst, ns, (double)val - v);          /* src/sound_render.c:800 */
        val = (int)((unsigned)val << q.shift);          /* same bits as the reference's signed << */
        *p++ = (uint8_t)val;
        *p++ = (uint8_t)(val >> 8);
        if (q.bytes == 3) *p++ = (uint8_t)(val >> 16);
    }
    return (int64_t)(p - out);
}

/* ------------------------------------------------------------------------------------------
 * modulator graph (reference src/adv_modulator.c:485-583 node ops, :604-760 frame loop)
 * ---------------------------------------------------------------------------------------- */
static inline double scaled_freq(double f, int scaled)
{
    /* adv_modulator.c:36-38: mHz grid as unsigned, back to double */
    return scaled ? (double)(unsigned)(f * (double)ICWO_HZ_SCALE + 0.5) : f;
}

static void rotate(double c, double s, const double *in, double *out)
{
    double re = in[0] * c - in[1] * s;
    double im = in[0] * s + in[1] * c;
    out[0] = re; out[1] = im;
}

static void node_shift(int on, double fr, int scaled, double omega, const double *in, double *out)
{
    /* adv_modulator.c:519-550 */
    if (!on) { out[0] = in[0]; out[1] = in[1]; return; }
    int neg = fr < 0.0;
    double f = scaled_freq(neg ? -fr : fr, scaled);
    double ph = fmod(omega * f, 2.0 * ICWO_PI);
    double c = cos(ph), s = sin(ph);
    if (neg) s = -s;
    rotate(c, s, in, out);
}

static void node_pm(int on, const double *p, int scaled, double omega, const double *in, double *out)
{
    /* adv_modulator.c:554-583; p = {freq, phase, level, angle} */
    if (!on) { out[0] = in[0]; out[1] = in[1]; return; }
    double f = scaled_freq(p[0], scaled);
    double ph = fmod(omega * f, 2.0 * ICWO_PI);
    double psi = p[2] * ICWO_PI * (sin(ph + p[1] * ICWO_PI) + p[3]);
    rotate(cos(psi), sin(psi), in, out);
}

static double node_master(int tout, const double *in)
{
    /* adv_modulator.c:485-507 */
    switch (tout) {
    case ICWO_OUT_RE:  return in[0];
    case ICWO_OUT_IM:  return in[1];
    case ICWO_OUT_ADD: return (in[0] + in[1]) / ICWO_SQRT2;
    case ICWO_OUT_SUB: return (in[0] - in[1]) / ICWO_SQRT2;
    default:           return 0.0;
    }
}

static void graph_frame(const icwo_spec *sp, double bus[ICWO_N_PLUGS][4], double omega, double lr[2])
{
    /* adv_modulator.c:637-751; nodes[] is already in execution order */
    int first = sp->bypass ? sp->n_nodes - 1 : 0;
    for (int n = first; n < sp->n_nodes; ++n) {
        const icwo_node *nd = &sp->nodes[n];
        double d[4], t;
        if (sp->bypass) {
            memcpy(d, bus[0], sizeof d);
        } else {
            d[0] = d[1] = d[2] = d[3] = 0.0;
            for (int k = 0; k < ICWO_N_PLUGS; ++k)
                if ((nd->inputs_mask >> k) & 1u)
                    for (int c = 0; c < 4; ++c) d[c] += bus[k][c];
        }
        switch (nd->xch_mode) {
        case ICWO_XCH_SWAP:
            t = d[0]; d[0] = d[2]; d[2] = t;
            t = d[1]; d[1] = d[3]; d[3] = t;
            break;
        case ICWO_XCH_LEFT:  d[2] = d[0]; d[3] = d[1]; break;
        case ICWO_XCH_RIGHT: d[0] = d[2]; d[1] = d[3]; break;
        case ICWO_XCH_MIXLR:
            d[0] = d[2] = (d[0] + d[2]) / 2.0;
            d[1] = d[3] = (d[1] + d[3]) / 2.0;
            break;
        default: break;
        }
        if (nd->l_iq_invert) { t = d[0]; d[0] = d[1]; d[1] = t; }
        if (nd->r_iq_invert) { t = d[2]; d[2] = d[3]; d[3] = t; }
        d[0] *= nd->l_gain; d[1] *= nd->l_gain;
        d[2] *= nd->r_gain; d[3] *= nd->r_gain;

        switch (nd->mode) {
        case ICWO_MODE_MASTER:
            lr[0] = node_master(nd->l_tout, &d[0]);
            lr[1] = node_master(nd->r_tout, &d[2]);
            break;
        case ICWO_MODE_SHIFT:
            node_shift(nd->l_on, nd->l_p[0], sp->is_frmod_scaled, omega, &d[0], &bus[nd->n_out][0]);
            node_shift(nd->r_on, nd->r_p[0], sp->is_frmod_scaled, omega, &d[2], &bus[nd->n_out][2]);
            break;
        case ICWO_MODE_PM:
            node_pm(nd->l_on, nd->l_p, sp->is_frmod_scaled, omega, &d[0], &bus[nd->n_out][0]);
            node_pm(nd->r_on, nd->r_p, sp->is_frmod_scaled, omega, &d[2], &bus[nd->n_out][2]);
            break;
        case ICWO_MODE_MIX:
            memcpy(bus[nd->n_out], d, sizeof d);
            break;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * whole chain
 * ---------------------------------------------------------------------------------------- */
void icwo_default_spec(icwo_spec *sp)
{
    /* reference defaults, src/config.c:118-207 and adv_modulator.c:112-118 */
    memset(sp, 0, sizeof *sp);
    sp->fmt = ICWO_FMT_WAV_F32;
    sp->n_channels = 2;
    sp->sample_rate = 48000;
    sp->filter_no = 1;
    sp->is_kahan = 1;
    sp->is_subnorm_reject = 1;
    sp->is_frmod_scaled = 1;
    sp->need24bits = 1;
    sp->dth_bits = 1.0;
    sp->quantz_type = 1;
    sp->render_type = ICWO_DITHER_NONE;
    sp->sign_bits16 = 16;
    sp->sign_bits24 = 24;
    sp->n_nodes = 1;
    sp->nodes[0].mode = ICWO_MODE_MASTER;
    sp->nodes[0].inputs_mask = 1u;
    sp->nodes[0].l_gain = sp->nodes[0].r_gain = 0.8;
    sp->nodes[0].l_tout = sp->nodes[0].r_tout = ICWO_OUT_ADD;
}

void icwo_state_init(icwo_state *st)
{
    memset(st, 0, sizeof *st);
    icwo_mt_seed(&st->mt[0], 0x13579BDFu);      /* src/in_cwave.c:69 */
    icwo_mt_seed(&st->mt[1], 0x479B22ABu);      /* src/in_cwave.c:70 */
    st->peak_db[0] = st->peak_db[1] = ICWO_SILENCE_DB;
}

int icwo_frame_bytes(const icwo_spec *sp)
{
    int cb = chan_bytes(sp->fmt);
    return cb < 0 ? -1 : cb * sp->n_channels;
}

int icwo_out_frame_bytes(const icwo_spec *sp) { return sp->need24bits ? 6 : 4; }

static double fade_gain(const icwo_spec *sp, int64_t ix)
{
    /* xwave_reader.c:921-936; negative = no fade */
    if (ix < sp->n_fade_in)
        return (double)ix / (double)sp->n_fade_in;
    if (ix > sp->n_samples - sp->n_fade_out && ix < sp->n_samples)
        return (double)(sp->n_samples - ix) / (double)sp->n_fade_out;
    return -1.0;
}

int icwo_process(const icwo_spec *sp, icwo_state *st, const uint8_t *in, int64_t n,
                 uint8_t *pcm, double *analytic, double *bus_tap, const int *tap_plugs,
                 int n_tap, double *lr_tap)
{
    enum { BLK = 4096 };
    static _Thread_local double x4[BLK][4], ai[BLK], aq[BLK], xr[BLK], lo[BLK], ro[BLK];
    int fb = icwo_frame_bytes(sp);
    int ob = icwo_out_frame_bytes(sp);
    int is_complex = sp->fmt >= ICWO_FMT_CW_F64;

    if (fb < 0 || sp->n_nodes < 1 || sp->n_nodes > ICWO_MAX_NODES ||
        sp->nodes[sp->n_nodes - 1].mode != ICWO_MODE_MASTER || sp->filter_no < 0 ||
        sp->filter_no >= ICW_HB_NTYPES)
        return -1;

    for (int64_t base = 0; base < n; base += BLK) {
        int m = (int)(n - base < BLK ? n - base : BLK);

        /* unpack + fade (xwave_reader.c:908-1009) */
        icwo_unpack(sp->fmt, sp->n_channels, in + base * fb, m, &x4[0][0]);
        for (int k = 0; k < m; ++k) {
            double g = fade_gain(sp, st->pos + k);
            if (g >= 0.0) {
                if (is_complex) { x4[k][0] *= g; x4[k][1] *= g; x4[k][2] *= g; x4[k][3] *= g; }
                else {
                    x4[k][0] *= g;
                    /* mono: the right channel reuses the already faded left value (:988-998) */
                    x4[k][2] = sp->n_channels > 1 ? x4[k][2] * g : x4[k][0];
                }
            }
        }
        st->pos += m;

        /* real input: per-channel analytic conversion */
        if (!is_complex) {
            for (int ch = 0; ch < 2; ++ch) {
                for (int k = 0; k < m; ++k) xr[k] = x4[k][2 * ch];
                g_fpc = sp->is_fp_check ? st->fp_cnt[ch] : NULL;        /* fes_hilb_left / _right (xwave_reader.c:980,998) */
                icwo_hilbert(sp->filter_no, sp->is_kahan, sp->is_subnorm_reject, st->lpf[ch],
                             &st->quad[ch], xr, m, ai, aq);
                g_fpc = NULL;
                for (int k = 0; k < m; ++k) { x4[k][2 * ch] = ai[k]; x4[k][2 * ch + 1] = aq[k]; }
            }
        }

        /* oscillator + graph, frame by frame (the bus persists between frames) */
        for (int k = 0; k < m; ++k) {
            double omega, lr[2] = { 0.0, 0.0 };
            if (sp->is_frmod_scaled) {
                /* adv_modulator.c:612-618 */
                unsigned scale_sr = sp->sample_rate * ICWO_HZ_SCALE;
                omega = (2.0 * ICWO_PI) * (double)st->n_frame / (double)scale_sr;
                st->n_frame = (st->n_frame + 1) % (uint64_t)scale_sr;
            } else {
                omega = (2.0 * ICWO_PI) * (double)st->n_frame / (double)sp->sample_rate;
                ++st->n_frame;
            }
            memcpy(st->bus[0], x4[k], sizeof x4[k]);
            graph_frame(sp, st->bus, omega, lr);
            lo[k] = lr[0]; ro[k] = lr[1];
            if (analytic) memcpy(analytic + 4 * (base + k), x4[k], sizeof x4[k]);
            if (bus_tap)
                for (int j = 0; j < n_tap; ++j)
                    memcpy(bus_tap + ((base + k) * n_tap + j) * 4, st->bus[tap_plugs[j]], 4 * sizeof(double));
            if (lr_tap) { lr_tap[2 * (base + k)] = lr[0]; lr_tap[2 * (base + k) + 1] = lr[1]; }
        }

        /* render L then R per frame, interleaved (adv_modulator.c:756-759) */
        if (pcm) {
            uint8_t tmpl[BLK * 3], tmpr[BLK * 3];
            int sb = ob / 2;
            g_fpc = sp->is_fp_check ? st->fp_cnt[2] : NULL;             /* fes_sr_left / _right (adv_modulator.c:757-758) */
            icwo_render(sp, &st->mt[0], &st->prev_rnd[0], &st->ns[0], lo, m, tmpl, &st->clips[0], &st->peak_db[0]);
            g_fpc = sp->is_fp_check ? st->fp_cnt[3] : NULL;
            icwo_render(sp, &st->mt[1], &st->prev_rnd[1], &st->ns[1], ro, m, tmpr, &st->clips[1], &st->peak_db[1]);
            g_fpc = NULL;
            for (int k = 0; k < m; ++k) {
                memcpy(pcm + (base + k) * ob, tmpl + k * sb, sb);
                memcpy(pcm + (base + k) * ob + sb, tmpr + k * sb, sb);
            }
        }
    }
    return 0;
}
