/* icw_plugin.c -- host C layer: the reference's transcode entry points over libicw_b200.so.
 * See include/icw_plugin.h.  Plain C99; no CUDA here -- everything numeric happens behind the C ABI. */
#define _FILE_OFFSET_BITS 64
#define _POSIX_C_SOURCE 200809L
#define _DEFAULT_SOURCE
#include <time.h>
#include "icw_plugin.h"

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <sys/types.h>
#include <unistd.h>

#define MIN_FILE_SAMPLES 2              /* reference src/in_cwave.h:315 */
#define MAX_FS_SRC 2000000u             /* reference src/in_cwave.h:160 */
#define DEFAULT_READAHEAD (1 << 20)

typedef struct reader {
    FILE    *fp;
    icwp_fileinfo fi;
    int      frame_bytes;
    int64_t  pos_samples, pos_tail;     /* like XWAVE_READER, reference src/in_cwave.h:394-395 */
} reader;

/* One read-ahead block: `to_read` frames from the file at byte offset `file_off`, then `to_fill` frames of
 * virtual silence, in a page-locked buffer (so the H2D copy of icw_session_process_host is a real
 * asynchronous DMA).  Three of them: while the host drains the PCM of one in its 4-64 KB getData calls, the render
 * thread has the next on the GPU and the reader thread reads the one after that from the file. */
typedef struct block {
    unsigned char *buf;
    size_t         cap;
    int64_t        pos_samples, pos_tail;   /* reader position the block starts at */
    int64_t        to_read, to_fill;
    int64_t        file_off;
    int            ok;                       /* the file gave every byte asked for */
    /* the block rendered: its PCM (page-locked), and the stream state its first frame started from */
    unsigned char *pcm;
    size_t         pcm_cap;
    int            rendered;
    icw_stream_state snap;
} block;

static struct {
    int             configured;
    icw_chain_spec  chain;
    icwp_options    opt;
    icw_engine     *engine;
    icw_session    *session;
    reader          rd;
    int             open;
    int             out_frame_bytes;
    /* read-ahead.  The DSP state on the device stands at the END of the last rendered block; a block's `snap` is the
     * state at its START, so that a seek or an early close can put the stream back to exactly the frames the host
     * took (the reference advances its MOD_CONTEXT only by what each getData asked for, src/transcode.c:82-100) */
    block           blk[3];                  /* cur: served; cur+1: rendered ahead; cur+2: read ahead (indices mod 3) */
    int             cur;                     /* block whose PCM is being served */
    int             pf_valid;                /* blk[cur + 1] holds (or is receiving) the rendered block after `cur` */
    int             rd_for;                  /* block index the reader thread was given last (-1: none) */
    int64_t         want;                    /* frames per block */
    int64_t         pcm_have, pcm_taken;     /* bytes of blk[cur].pcm */
    int64_t         blk_frames;              /* frames rendered from blk[cur] */
    /* two worker threads: file reads, and rendering (read k+2 || render k+1 || host drains k) */
    pthread_mutex_t mu;
    pthread_cond_t  cv_req, cv_done;
    icwp_iostats    io;
} P = { .mu = PTHREAD_MUTEX_INITIALIZER, .cv_req = PTHREAD_COND_INITIALIZER, .cv_done = PTHREAD_COND_INITIALIZER };

static void ensure_defaults(void);

/* ---- little-endian field readers ---------------------------------------------------------------- */
static unsigned rd16(const unsigned char *p) { return (unsigned)p[0] | ((unsigned)p[1] << 8); }
static uint32_t rd32(const unsigned char *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

static int64_t file_size(FILE *fp)
{
    off_t cur = ftello(fp), end;
    if (cur < 0 || fseeko(fp, 0, SEEK_END)) return 0;
    end = ftello(fp);
    fseeko(fp, cur, SEEK_SET);
    return end < 0 ? 0 : (int64_t)end;
}

/* ---- CWAVE header (reference src/cwave.h:47-84, checks src/xwave_reader.c:246-339) ------------------ */
static int parse_cwave(FILE *fp, icwp_fileinfo *fi)
{
    static const int slen[4] = { 16, 4, 6, 8 };     /* bytes per channel sample, formats 0..3 */
    unsigned char h[48];
    int64_t fsize = file_size(fp);
    uint32_t hsize, version, format, nch, nsmp, srate;

    if (fsize < 48 || fread(h, 1, 48, fp) != 48) return 0;
    if (memcmp(h, "cPLXwAVE", 8)) return 0;
    hsize = rd32(h + 8); version = rd32(h + 12); format = rd32(h + 16);
    nch = rd32(h + 20); nsmp = rd32(h + 24); srate = rd32(h + 28);
    if (hsize < 48 || (int64_t)hsize >= fsize) return 0;
    if (version != 1 && version != 2) return 0;
    if (format > 3) return 0;
    if (nch == 0 || nch > 2) return 0;
    if (nsmp < MIN_FILE_SAMPLES) return 0;
    if ((int64_t)nsmp * nch * slen[format] + hsize > fsize) return 0;
    if (srate == 0) return 0;
    fi->fmt = ICW_FMT_CW_F64 + (int)format;
    fi->n_channels = (int)nch;
    fi->sample_rate = srate;
    fi->n_samples = nsmp;
    fi->offset_data = hsize;
    return 1;
}

/* ---- RIFF/WAVE (reference src/xwave_reader.c:362-585) ------------------------------------------------ */
static int pcm_bits_to_fmt(unsigned bits)
{
    switch (bits) {
    case 8: return ICW_FMT_WAV_U8;
    case 16: return ICW_FMT_WAV_I16;
    case 24: return ICW_FMT_WAV_I24;
    case 32: return ICW_FMT_WAV_I32;
    default: return -1;
    }
}

static int parse_wav(FILE *fp, icwp_fileinfo *fi)
{
    static const unsigned char guid_tail[14] = { 0x00, 0x00, 0x00, 0x00, 0x10, 0x00, 0x80, 0x00, 0x00, 0xaa, 0x00, 0x38, 0x9b, 0x71 };
    unsigned char hb[12], ck[8], fmtb[40];
    int64_t fsize = file_size(fp), fpos = 0;
    uint32_t clen, dlen;
    unsigned tag, nch, block, bits, srate;
    size_t toread;
    int fmt;

    if (fread(hb, 1, 12, fp) != 12 || memcmp(hb, "RIFF", 4) || memcmp(hb + 8, "WAVE", 4)) return 0;
    fpos = 12;
    for (;;) {                                      /* find "fmt "; a "data" chunk before it is an error */
        if (fread(ck, 1, 8, fp) != 8) return 0;
        clen = rd32(ck + 4);
        fpos += 8;
        if (fpos + (int64_t)clen > fsize) return 0;
        if (!memcmp(ck, "fmt ", 4)) break;
        if (!memcmp(ck, "data", 4)) return 0;
        if (fseeko(fp, (off_t)clen, SEEK_CUR)) return 0;
        fpos += clen;
    }
    if (clen < 14) return 0;
    toread = clen < 40 ? clen : 40;
    if (fread(fmtb, 1, toread, fp) != toread) return 0;
    if (toread < clen && fseeko(fp, (off_t)(clen - toread), SEEK_CUR)) return 0;
    fpos += clen;
    tag = rd16(fmtb); nch = rd16(fmtb + 2); srate = rd32(fmtb + 4); block = rd16(fmtb + 12);
    bits = toread >= 16 ? rd16(fmtb + 14) : 0;
    bits = (bits + 7) & ~7u;
    if (!nch || nch > 2 || !srate) return 0;
    switch (tag) {
    case 0x0001:                                    /* WAVE_FORMAT_PCM */
        if (clen < 16) {                            /* bare WAVEFORMAT: derive the width from the block size */
            if (block % nch) return 0;
            bits = (block / nch) << 3;
        }
        if ((fmt = pcm_bits_to_fmt(bits)) < 0) return 0;
        break;
    case 0x0003:                                    /* WAVE_FORMAT_IEEE_FLOAT */
        if (clen < 16 || bits != 32) return 0;
        fmt = ICW_FMT_WAV_F32;
        break;
    case 0xFFFE:                                    /* WAVE_FORMAT_EXTENSIBLE */
        if (clen < 40 || rd16(fmtb + 16) < 22) return 0;
        if (memcmp(fmtb + 26, guid_tail, 14)) return 0;
        if (rd16(fmtb + 24) == 0x0001) { if ((fmt = pcm_bits_to_fmt(bits)) < 0) return 0; }
        else if (rd16(fmtb + 24) == 0x0003) { if (bits != 32) return 0; fmt = ICW_FMT_WAV_F32; }
        else return 0;
        break;
    default:
        return 0;
    }
    for (;;) {                                      /* find "data" */
        if (fread(ck, 1, 8, fp) != 8) return 0;
        fpos += 8;
        dlen = rd32(ck + 4);
        if (fpos + (int64_t)dlen > fsize) return 0;
        if (!memcmp(ck, "data", 4)) break;
        if (fseeko(fp, (off_t)dlen, SEEK_CUR)) return 0;
        fpos += dlen;
    }
    {
        unsigned sample_size = (bits >> 3) * nch;
        fi->n_samples = dlen / sample_size;
        if (fi->n_samples < MIN_FILE_SAMPLES || block != sample_size) return 0;
    }
    fi->fmt = fmt;
    fi->n_channels = (int)nch;
    fi->sample_rate = srate;
    fi->offset_data = fpos;
    return 1;
}

/* everything xwave_reader_create decides (reference src/xwave_reader.c:593-728) */
static FILE *open_and_parse(const char *name, const icwp_options *opt, icwp_fileinfo *fi)
{
    const char *dot;
    FILE *fp;
    int is_cw, ok;

    if (!name || !*name) return NULL;
    dot = strrchr(name, '.');
    if (!dot || dot == name) return NULL;           /* the extension decides the parser (:125-134) */
    if (!strcasecmp(dot + 1, "CWAVE")) is_cw = 1;
    else if (!strcasecmp(dot + 1, "WAV") || !strcasecmp(dot + 1, "RWAVE")) is_cw = 0;
    else return NULL;
    if (!(fp = fopen(name, "rb"))) return NULL;
    memset(fi, 0, sizeof *fi);
    ok = is_cw ? parse_cwave(fp, fi) : parse_wav(fp, fi);
    if (!ok || fi->sample_rate > MAX_FS_SRC) { fclose(fp); return NULL; }
    if (opt && opt->sec_align) {                    /* virtual zero tail up to a multiple of sec_align seconds */
        int64_t max_tail = (int64_t)fi->sample_rate * (int64_t)opt->sec_align;
        int64_t rest = fi->n_samples % max_tail;
        fi->n_tail = rest ? max_tail - rest : 0;
    }
    fi->n_fade_in = opt ? (int64_t)(((uint64_t)opt->fade_in_ms * fi->sample_rate) / 1000u) : 0;
    fi->n_fade_out = opt ? (int64_t)(((uint64_t)opt->fade_out_ms * fi->sample_rate) / 1000u) : 0;
    if (fi->n_fade_in + fi->n_fade_out >= fi->n_samples) {      /* too short for the faders (:700-716) */
        if (fi->n_samples < 300) fi->n_fade_in = fi->n_fade_out = 0;
        else {
            if (fi->n_fade_in) fi->n_fade_in = fi->n_samples / 3;
            if (fi->n_fade_out) fi->n_fade_out = fi->n_samples / 3;
        }
    }
    return fp;
}

/* CWAVE sample-data integrity (reference check_cwave, src/gui_cwave.c:82-130): CRC-32 over exactly
 * n_samples frames from offset_data, compared with the header's n_CRC32 when the file is V2. */
int icwp_check_cwave(const char *filename, uint32_t *crc_calc, uint32_t *crc_file, int *has_crc)
{
    icwp_fileinfo fi;
    unsigned char h[48];
    FILE *fp;
    int64_t left;
    uint32_t crc = 0, part;
    int first = 1, ok = 1;
    const size_t block = (size_t)32 << 20;
    unsigned char *buf;

    ensure_defaults();
    fp = open_and_parse(filename, NULL, &fi);
    if (!fp) return 0;
    if (fi.fmt < ICW_FMT_CW_F64) { fclose(fp); return 0; }          /* WAV carries no CRC */
    if (fseeko(fp, 0, SEEK_SET) || fread(h, 1, 48, fp) != 48) { fclose(fp); return 0; }
    if (has_crc) *has_crc = rd32(h + 12) > 1;                        /* V1 headers have no CRC field */
    if (crc_file) *crc_file = rd32(h + 36);
    if (!P.engine && icw_engine_create(P.opt.device, &P.engine) != ICW_OK) { fclose(fp); return 0; }
    {
        icw_chain_spec sp;
        icw_default_spec(&sp);
        sp.fmt = fi.fmt; sp.n_channels = fi.n_channels;
        left = fi.n_samples * (int64_t)icw_frame_bytes(&sp);
    }
    buf = malloc(block);
    if (!buf || fseeko(fp, (off_t)fi.offset_data, SEEK_SET)) { free(buf); fclose(fp); return 0; }
    while (left > 0 && ok) {
        size_t n = left < (int64_t)block ? (size_t)left : block;
        if (fread(buf, 1, n, fp) != n) { ok = 0; break; }           /* "Read error or file corrupted" */
        if (icw_crc32_host(P.engine, buf, n, &part) != ICW_OK) { ok = 0; break; }
        crc = first ? part : icw_crc32_combine(crc, part, n);
        first = 0;
        left -= (int64_t)n;
    }
    free(buf);
    fclose(fp);
    if (ok && crc_calc) *crc_calc = crc;
    return ok;
}

int icwp_probe(const char *filename, const icwp_options *opt, icwp_fileinfo *out)
{
    icwp_fileinfo fi;
    FILE *fp = open_and_parse(filename, opt, &fi);
    if (!fp) return 0;
    fclose(fp);
    if (out) *out = fi;
    return 1;
}

/* ---- configuration ---------------------------------------------------------------------------------- */
static void ensure_defaults(void)
{
    if (P.configured) return;
    icw_default_spec(&P.chain);
    memset(&P.opt, 0, sizeof P.opt);
    P.configured = 1;
}

int icwp_configure(const icw_chain_spec *chain, const icwp_options *opt)
{
    ensure_defaults();
    if (chain) P.chain = *chain; else icw_default_spec(&P.chain);
    if (opt) P.opt = *opt; else memset(&P.opt, 0, sizeof P.opt);
    if (P.session && !P.open) {
        /* parameters are snapshotted at the next open */
    }
    return ICW_OK;
}

static void reader_stop(void);
static void reader_wait(void);
static void settle(void);

void icwp_reset(void)
{
    if (P.open) winampGetExtendedRead_close((intptr_t)&P);
    if (P.session) { icw_session_destroy(P.session); P.session = NULL; }
    reader_stop();
    for (int i = 0; i < 3; ++i) {
        icw_pinned_free(P.blk[i].buf); P.blk[i].buf = NULL; P.blk[i].cap = 0;
        icw_pinned_free(P.blk[i].pcm); P.blk[i].pcm = NULL; P.blk[i].pcm_cap = 0;
    }
}

int icwp_io_stats(icwp_iostats *out, int reset)
{
    if (out) *out = P.io;
    if (reset) memset(&P.io, 0, sizeof P.io);
    return ICW_OK;
}

int icwp_stats(icw_stats *out)
{
    if (!P.session) return ICW_E_ARG;
    reader_wait();                                  /* the worker may be rendering the next block on this session */
    return icw_session_stats(P.session, out);
}

/* ---- the reference's entry points -------------------------------------------------------------------- */
intptr_t winampGetExtendedRead_open(const char *filename, int *size, int *bps, int *nch, int *srate)
{
    icw_chain_spec sp;
    int64_t total, sz;
    /* every mod_context_fopen re-runs sound_render_recalc on both renderers (src/in_cwave.c:231-234): sloped-TPDF memory and
     * shaper memory start from zero in every file, the dither generators carry on */
    unsigned reset = ICW_RESET_FILEPOS | ICW_RESET_RENDER_MEMORY;

    ensure_defaults();
    if (P.open) return 0;                           /* one transcode at a time, like &the.mc_transcode */
    P.rd.fp = open_and_parse(filename, &P.opt, &P.rd.fi);
    if (!P.rd.fp) return 0;
    sp = P.chain;
    sp.fmt = P.rd.fi.fmt;
    sp.n_channels = P.rd.fi.n_channels;
    sp.sample_rate = P.rd.fi.sample_rate;
    sp.n_samples = P.rd.fi.n_samples;
    sp.n_fade_in = P.rd.fi.n_fade_in;
    sp.n_fade_out = P.rd.fi.n_fade_out;
    P.rd.frame_bytes = icw_frame_bytes(&sp);
    P.out_frame_bytes = icw_out_frame_bytes(&sp);
    total = P.rd.fi.n_samples + P.rd.fi.n_tail;
    sz = total * P.out_frame_bytes;
    /* the reference refuses outputs that do not fit a 2 GiB WAV (src/transcode.c:59-68) */
    if (sz > 0x7FFFFFFFLL - 40 - 12 - 16 - 64) goto fail;
    if (!P.engine && icw_engine_create(P.opt.device, &P.engine) != ICW_OK) goto fail;
    if (!P.session) {
        if (icw_session_create(P.engine, &sp, 1, &P.session) != ICW_OK) goto fail;
    } else if (icw_session_set_spec(P.session, &sp) != ICW_OK) goto fail;
    if (P.opt.clr_nframe_trk) reset |= ICW_RESET_FRAMECNT;      /* src/in_cwave.c:226-229 */
    if (P.opt.clr_hilb_trk) reset |= ICW_RESET_HILBERT;
    if (icw_session_reset(P.session, reset) != ICW_OK) goto fail;
    P.rd.pos_samples = P.rd.pos_tail = 0;
    P.pcm_have = P.pcm_taken = 0;
    P.pf_valid = 0; P.cur = 0; P.blk_frames = 0; P.rd_for = -1;
    P.open = 1;
    *nch = 2;
    *srate = (int)P.rd.fi.sample_rate;
    *bps = P.out_frame_bytes * 4;
    *size = (int)sz;
    return (intptr_t)&P;
fail:
    fclose(P.rd.fp);
    P.rd.fp = NULL;
    return 0;
}

/* ---- overlapped file I/O (the step before the path: reference src/xwave_reader.c:838-904) ----------------
 * The reference reads read_quant frames with one ReadFile, then computes them, in series.  Here a reader
 * thread preads block k+1 into page-locked memory while block k is on the GPU and its PCM is being handed
 * to the host; only the main thread ever touches the DSP state. */
static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static int fill_block(block *b, int fd, int silence_byte, int frame_bytes)
{
    size_t want = (size_t)b->to_read * (size_t)frame_bytes, got = 0;
    while (got < want) {
        ssize_t r = pread(fd, b->buf + got, want - got, (off_t)(b->file_off + (int64_t)got));
        if (r <= 0) return 0;                       /* short read == broken file == end, as in the reference */
        got += (size_t)r;
    }
    if (b->to_fill)     /* silence: 0x80 for unsigned 8-bit, zero bits otherwise (src/xwave_reader.c:100-105) */
        memset(b->buf + want, silence_byte, (size_t)b->to_fill * (size_t)frame_bytes);
    return 1;
}

/* render blk[i] (already filled) into its PCM buffer, remembering the state it started from */
static int render_block(block *b, int64_t n)
{
    size_t need_out = (size_t)n * P.out_frame_bytes;
    double t0;
    b->rendered = 0;
    if (need_out > b->pcm_cap) {
        icw_pinned_free(b->pcm);
        b->pcm = NULL; b->pcm_cap = 0;
        if (icw_pinned_alloc(need_out, (void **)&b->pcm) != ICW_OK) return 0;
        b->pcm_cap = need_out;
    }
    if (icw_session_get_state(P.session, 0, &b->snap) != ICW_OK) return 0;
    t0 = now_s();
    if (icw_session_process_host(P.session, n, b->buf, 0, b->pcm, 0) != ICW_OK) return 0;
    pthread_mutex_lock(&P.mu);
    P.io.gpu_s += now_s() - t0;
    P.io.frames += (uint64_t)n;
    pthread_mutex_unlock(&P.mu);
    b->rendered = 1;
    return 1;
}

typedef struct worker {
    pthread_t th;
    int       up, quit, pending, arg;
    void    (*fn)(int);
} worker;

static void *worker_main(void *arg)
{
    worker *w = (worker *)arg;
    pthread_mutex_lock(&P.mu);
    for (;;) {
        while (!w->pending && !w->quit) pthread_cond_wait(&P.cv_req, &P.mu);
        if (w->quit) break;
        int a = w->arg;
        pthread_mutex_unlock(&P.mu);
        w->fn(a);
        pthread_mutex_lock(&P.mu);
        w->pending = 0;
        pthread_cond_broadcast(&P.cv_done);
    }
    pthread_mutex_unlock(&P.mu);
    return NULL;
}

static void worker_wait(worker *w)
{
    pthread_mutex_lock(&P.mu);
    while (w->pending) pthread_cond_wait(&P.cv_done, &P.mu);
    pthread_mutex_unlock(&P.mu);
}

static int worker_submit(worker *w, int arg)
{
    if (!w->up) {
        if (pthread_create(&w->th, NULL, worker_main, w)) return 0;
        w->up = 1;
    }
    pthread_mutex_lock(&P.mu);
    w->arg = arg; w->pending = 1;
    pthread_cond_broadcast(&P.cv_req);
    pthread_mutex_unlock(&P.mu);
    return 1;
}

static void worker_stop(worker *w)
{
    if (!w->up) return;
    worker_wait(w);
    pthread_mutex_lock(&P.mu);
    w->quit = 1;
    pthread_cond_broadcast(&P.cv_req);
    pthread_mutex_unlock(&P.mu);
    pthread_join(w->th, NULL);
    w->up = 0; w->quit = 0;
}

static int64_t plan_block(block *b, int64_t ps, int64_t pt, int64_t want);
static int block_reserve(block *b, size_t need);

/* reader thread: the file bytes (and the silence tail) of blk[i] */
static void do_read(int i)
{
    block *b = &P.blk[i];
    const int sil = P.rd.fi.fmt == ICW_FMT_WAV_U8 ? 0x80 : 0, fb = P.rd.frame_bytes;
    double t0 = now_s();
    int ok = fill_block(b, b->to_read > 0 ? fileno(P.rd.fp) : -1, sil, fb);
    double dt = now_s() - t0;
    pthread_mutex_lock(&P.mu);
    b->ok = ok;
    P.io.read_s += dt;
    P.io.read_bytes += (uint64_t)b->to_read * (uint64_t)fb;
    pthread_mutex_unlock(&P.mu);
}
static worker W_read = { .fn = do_read };

/* render thread: blk[i] (planned by whoever asked) -> its PCM.  The session is this thread's until the request is done:
 * every other user of the session waits for that first (workers_wait). */
static void do_render(int i)
{
    block *b = &P.blk[i], *nb = &P.blk[(i + 1) % 3];
    int64_t n_next;
    if (P.rd_for == i) worker_wait(&W_read);        /* read while the block before this one was on the GPU */
    else do_read(i);
    /* the block after this one: off to the reader thread before the GPU gets this one */
    n_next = plan_block(nb, b->pos_samples + b->to_read, b->pos_tail + b->to_fill, P.want);
    nb->rendered = 0;
    if (n_next > 0 && block_reserve(nb, (size_t)n_next * P.rd.frame_bytes) && worker_submit(&W_read, (i + 1) % 3)) P.rd_for = (i + 1) % 3;
    else P.rd_for = -1;
    if (b->ok) render_block(b, b->to_read + b->to_fill);
}
static worker W_render = { .fn = do_render };

static void reader_wait(void)                       /* both workers idle: the session and every block are the caller's */
{
    worker_wait(&W_render);
    worker_wait(&W_read);
}

static void reader_stop(void)
{
    worker_stop(&W_render);
    worker_stop(&W_read);
}

/* what xwave_read_samples would take next from position (ps, pt): frames from the file, then silence */
static int64_t plan_block(block *b, int64_t ps, int64_t pt, int64_t want)
{
    const reader *r = &P.rd;
    int64_t to_read = 0, to_fill = 0;
    if (ps < r->fi.n_samples) {
        to_read = r->fi.n_samples - ps;
        if (to_read > want) to_read = want;
    }
    if (to_read < want && r->fi.n_tail) {
        to_fill = r->fi.n_tail - pt;
        if (to_fill > want - to_read) to_fill = want - to_read;
    }
    b->pos_samples = ps; b->pos_tail = pt;
    b->to_read = to_read; b->to_fill = to_fill;
    b->file_off = r->fi.offset_data + ps * r->frame_bytes;
    b->ok = 0;
    return to_read + to_fill;
}

static int block_reserve(block *b, size_t need)
{
    if (need <= b->cap) return 1;
    icw_pinned_free(b->buf);
    b->buf = NULL; b->cap = 0;
    if (icw_pinned_alloc(need, (void **)&b->buf) != ICW_OK) return 0;
    b->cap = need;
    return 1;
}

/* make the next block of up to `want` frames from the reader position the current one; returns its frames */
static int64_t refill(int64_t want)
{
    reader *r = &P.rd;
    block *b;
    int64_t n, n_next;
    const int next = (P.cur + 1) % 3;
    double t0;

    P.want = want;
    if (P.pf_valid && P.blk[next].pos_samples == r->pos_samples && P.blk[next].pos_tail == r->pos_tail) {
        t0 = now_s();
        worker_wait(&W_render);                     /* read and rendered while the host drained the block before it */
        P.io.wait_s += now_s() - t0;
        P.cur = next;
        b = &P.blk[P.cur];
        n = b->to_read + b->to_fill;
        P.pf_valid = 0;
        if (n <= 0 || !b->ok || !b->rendered) return 0;
        P.io.blocks_prefetched++;
    } else {
        reader_wait();                              /* (settle() has already dropped blocks made for another position) */
        P.pf_valid = 0; P.rd_for = -1;
        b = &P.blk[P.cur];
        n = plan_block(b, r->pos_samples, r->pos_tail, want);
        if (n <= 0) return 0;
        if (!block_reserve(b, (size_t)n * r->frame_bytes)) return 0;
        t0 = now_s();
        b->ok = fill_block(b, fileno(r->fp), r->fi.fmt == ICW_FMT_WAV_U8 ? 0x80 : 0, r->frame_bytes);
        P.io.read_s += now_s() - t0;
        P.io.read_bytes += (uint64_t)b->to_read * (uint64_t)r->frame_bytes;
        P.io.blocks_sync++;
        if (!b->ok || !render_block(b, n)) return 0;
    }
    r->pos_samples += b->to_read;
    r->pos_tail += b->to_fill;
    P.blk_frames = n;
    /* the block after this one goes to the render thread (its bytes are usually there already: the render thread put the
     * reader thread on them before it rendered this block) */
    {
        const int nn = (P.cur + 1) % 3;
        block *nb = &P.blk[nn];
        if (P.rd_for != nn) {
            n_next = plan_block(nb, r->pos_samples, r->pos_tail, want);
            if (n_next > 0 && !block_reserve(nb, (size_t)n_next * r->frame_bytes)) n_next = 0;
        } else {
            n_next = nb->to_read + nb->to_fill;
        }
        nb->rendered = 0;
        if (n_next > 0 && worker_submit(&W_render, nn)) P.pf_valid = 1;
    }
    return n;
}

/* Put the DSP state where the HOST is: rendered-but-unserved frames are un-done by restoring the state the
 * block started from and running only the frames that were handed out again (their input is still in
 * blk[cur]).  The reference never runs ahead of getData (src/transcode.c:82-100), so after this the frame
 * counter, Hilbert memory, dither stream and clip/peak counters are the reference's. */
static void settle(void)
{
    int64_t served;
    block *b = &P.blk[P.cur];
    int next_rendered;
    if (!P.session) { P.pcm_have = P.pcm_taken = 0; return; }
    reader_wait();                                  /* the worker may be in the middle of the block after this one */
    next_rendered = P.pf_valid && P.blk[(P.cur + 1) % 3].ok && P.blk[(P.cur + 1) % 3].rendered;
    P.pf_valid = 0; P.rd_for = -1;
    if (P.pcm_taken >= P.pcm_have) {
        /* everything rendered from this block was handed out; a block rendered ahead of it is un-done by going back to
         * the state it started from (= the state after this block) */
        if (next_rendered) { icw_session_set_state(P.session, 0, &P.blk[(P.cur + 1) % 3].snap); P.io.resettles++; }
        P.pcm_have = P.pcm_taken = 0;
        return;
    }
    served = P.pcm_taken / P.out_frame_bytes;
    if (icw_session_set_state(P.session, 0, &b->snap) == ICW_OK && served > 0)
        icw_session_process_host(P.session, served, b->buf, 0, b->pcm, 0);
    /* the reader stands after the served frames too */
    {
        int64_t from_file = served < b->to_read ? served : b->to_read;
        P.rd.pos_samples = b->pos_samples + from_file;
        P.rd.pos_tail = b->pos_tail + (served - from_file);
    }
    P.io.resettles++;
    P.pcm_have = P.pcm_taken = 0;
}

intptr_t winampGetExtendedRead_getData(intptr_t handle, char *dest, int len, int *killswitch)
{
    int64_t want_bytes, done = 0;
    if ((void *)handle != (void *)&P || !P.open || len <= 0 || (killswitch && *killswitch)) return 0;
    want_bytes = (int64_t)(len / P.out_frame_bytes) * P.out_frame_bytes;       /* whole frames only */
    while (done < want_bytes) {
        int64_t avail = P.pcm_have - P.pcm_taken, take;
        if (avail == 0) {
            int64_t ra = P.opt.readahead_frames > 0 ? P.opt.readahead_frames : DEFAULT_READAHEAD;
            int64_t got = refill(ra);
            if (got <= 0) break;
            P.pcm_have = got * P.out_frame_bytes;
            P.pcm_taken = 0;
            avail = P.pcm_have;
        }
        take = want_bytes - done < avail ? want_bytes - done : avail;
        memcpy(dest + done, P.blk[P.cur].pcm + P.pcm_taken, (size_t)take);
        P.pcm_taken += take;
        done += take;
    }
    return (intptr_t)done;
}

int winampGetExtendedRead_setTime(intptr_t handle, int decode_pos_ms)
{
    reader *r = &P.rd;
    int64_t total, pos;
    icw_stream_state st;
    if ((void *)handle != (void *)&P || !P.open) return 0;
    settle();                                       /* un-served read-ahead is dropped, like the reader's buffer */
    total = r->fi.n_samples + r->fi.n_tail;
    pos = (int64_t)decode_pos_ms * (int64_t)r->fi.sample_rate / 1000;
    if (pos > total) pos = total;
    if (pos < 0) pos = 0;
    if (pos <= r->fi.n_samples) { r->pos_samples = pos; r->pos_tail = 0; }
    else { r->pos_samples = r->fi.n_samples; r->pos_tail = pos - r->fi.n_samples; }
    /* the fades follow the absolute file position (src/xwave_reader.c:923); nothing else of the context moves
     * (xwave_seek_samples, src/xwave_reader.c:782-808) */
    if (icw_session_get_state(P.session, 0, &st) != ICW_OK) return 0;
    st.pos = pos;
    return icw_session_set_state(P.session, 0, &st) == ICW_OK;
}

void winampGetExtendedRead_close(intptr_t handle)
{
    if ((void *)handle != (void *)&P || !P.open) return;
    settle();                                       /* the next file continues from what the host took */
    reader_wait();
    P.pf_valid = 0;
    fclose(P.rd.fp);
    P.rd.fp = NULL;
    P.open = 0;
    P.pcm_have = P.pcm_taken = 0;
}

/* ---- the input-module table (reference src/in_cwave.c:551-572, src/playback.c: get_playback_iface) -------------------------- */
static int g_last_length_ms;

static void im_config(void *w) { (void)w; }
static void im_about(void *w) { (void)w; }
static int  im_init(void) { ensure_defaults(); return 0; }          /* IN_INIT_SUCCESS */
static void im_quit(void) { icwp_reset(); }
static void im_getfileinfo(const char *file, char *title, int *length_in_ms)
{
    icwp_fileinfo fi;
    if (title) title[0] = 0;
    if (length_in_ms) *length_in_ms = -1000;                         /* the reference's "unknown" (src/playback.c) */
    if (!file || !*file) {                                           /* NULL / "": the file last asked about */
        if (length_in_ms && g_last_length_ms) *length_in_ms = g_last_length_ms;
        return;
    }
    ensure_defaults();
    if (!icwp_probe(file, &P.opt, &fi)) {
        if (title) snprintf(title, 2048, "-CAN'T OPEN-");                /* src/playback.c:getfileinfo */
        return;
    }
    if (title) {
        const char *b = strrchr(file, '/');
        const char *b2 = strrchr(file, '\\');
        if (b2 && (!b || b2 > b)) b = b2;
        snprintf(title, 2048, "%s", b ? b + 1 : file);               /* GETFILEINFO_TITLE_LENGTH */
    }
    g_last_length_ms = (int)(((fi.n_samples + fi.n_tail) * 1000) / (fi.sample_rate ? fi.sample_rate : 1));
    if (length_in_ms) *length_in_ms = g_last_length_ms;
}
static int  im_infobox(const char *f, void *w) { (void)f; (void)w; return 1; }   /* INFOBOX_UNCHANGED */
static int  im_isourfile(const char *fn) { (void)fn; return 0; }                  /* by extension only, like the reference */
static int  im_play(const char *fn) { (void)fn; return ICWP_PLAY_UNSUPPORTED; }
static void im_void(void) { }
static int  im_zero(void) { return 0; }
static int  im_getlength(void) { return g_last_length_ms; }
static void im_int(int a) { (void)a; }

static icwp_in_module g_in_module = {
    ICWP_IN_VER,
    (char *)"in_cwave on B200 (transcode entry points; playback out of scope)",
    0, 0,
    (char *)"cwave\0CWAVE analytic signal (*.cwave)\0wav\0WAV / RWAVE (*.wav)\0",
    1, 1,
    im_config, im_about, im_init, im_quit, im_getfileinfo, im_infobox, im_isourfile,
    im_play, im_void, im_void, im_zero, im_void,
    im_getlength, im_zero, im_int, im_int, im_int,
    0, 0, 0, 0, 0, 0, 0, 0, 0,          /* visualisation feeds: filled in by the host */
    0, 0, 0, 0, 0, 0
};

icwp_in_module *winampGetInModule2(void)
{
    ensure_defaults();
    return &g_in_module;
}
