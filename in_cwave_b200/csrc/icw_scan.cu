// icw_scan.cu -- time-parallel real -> analytic conversion ("scan mode").
//
// The reference's half-band low-pass H(z) (order 15..20 elliptic, reference src/hblpf.c:740-820)
// is evaluated in its partial-fraction form  H(z) = d0 + sum_k r_k z^-1 / (1 - p_k z^-1):
// one complex one-pole recurrence s_k[n] = p_k s_k[n-1] + u[n] per conjugate pair,
// y[n] = sum_k 2 Re(r_k s_k[n-1]) (+ d0 u[n] for the baseline summation; the reference's Kahan
// path drops that term, src/hblpf.c:1056, and so do we).  One-pole recurrences are a linear
// scan: a chunk's end state from zero (pass 1), a carry scan over chunks (pass 2), the real run
// from the carried state (pass 3).  The fs/4 mixer of the converter (reference
// src/lpf_hilbert_quad.c:129-156) feeds the I filter only on even phases and the Q filter only on
// odd phases, so each state advances two samples at a time with p^2 and its zero-input output
// uses the residue r*p: 16 FMA per mode per sample PAIR per channel instead of 24.
//
// Numerics: this is the mathematically exact filter evaluated in FP64 (within 3e-15 of a
// quad-precision evaluation of the reference's recurrence, all six designs), NOT the reference's
// rounding sequence; it differs from the reference by the reference's own rounding noise
// (1e-9 .. 1e-3 of RMS depending on the design) and ignores the |w|<1
// state zeroing.  Exact mode (icw_fused.cu) is the bit-exact path.
#include <cmath>
#include <cstring>
#include <vector>

#include "icw_dev.cuh"
#include "icw_hb.cuh"
#include "icw_kernels.h"
#include "icw_scan.h"
#include "icw_hb_modal.inc"

namespace icw {

// ---------------------------------------------------------------------------------------------
// host: modal constants in extended precision, rounded once
// ---------------------------------------------------------------------------------------------
typedef long double ld;
struct cld { ld re, im; };
static inline cld cmul(cld a, cld b) { return { a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re }; }
static inline cld cpowi(cld a, long long e)
{
    cld r = { 1.0L, 0.0L };
    while (e) { if (e & 1) r = cmul(r, a); a = cmul(a, a); e >>= 1; }
    return r;
}

void scan_make_coef(int filter_no, bool baseline, double d0, ModalCoef &mc, std::vector<double> &pw_table)
{
    memset(&mc, 0, sizeof mc);
    const int nm = ICW_HB_NMODES[filter_no];
    mc.nm = nm;
    mc.baseline = baseline;
    mc.d0 = d0;
    pw_table.assign((size_t)SCAN_CH * SCAN_NMAX * 2, 0.0);
    for (int m = 0; m < nm; ++m) {
        const icw_hb_mode &md = ICW_HB_MODES[filter_no][m];
        cld p = { (ld)md.p_re[0] + (ld)md.p_re[1], (ld)md.p_im[0] + (ld)md.p_im[1] };
        cld r = { (ld)md.r_re[0] + (ld)md.r_re[1], (ld)md.r_im[0] + (ld)md.r_im[1] };
        const ld w = md.is_real ? 1.0L : 2.0L;          // a pair contributes r s + conj(r s) = 2 Re(r s)
        cld p2 = cmul(p, p);
        ld n2 = p.re * p.re + p.im * p.im;
        cld pinv = { p.re / n2, -p.im / n2 };
        cld rp = cmul(r, p);
        cld pl = cpowi(p, SCAN_L);
        cld pt = cpowi(p, (long long)SCAN_L * SCAN_CH);
        mc.p_re[m] = (double)p.re;       mc.p_im[m] = (double)p.im;
        mc.p2_re[m] = (double)p2.re;     mc.p2_im[m] = (double)p2.im;
        mc.pinv_re[m] = (double)pinv.re; mc.pinv_im[m] = (double)pinv.im;
        mc.c_re[m] = (double)(w * r.re);   mc.c_im[m] = (double)(-w * r.im);
        mc.cp_re[m] = (double)(w * rp.re); mc.cp_im[m] = (double)(-w * rp.im);
        mc.pl_re[m] = (double)pl.re;     mc.pl_im[m] = (double)pl.im;
        mc.pt_re[m] = (double)pt.re;     mc.pt_im[m] = (double)pt.im;
        cld pj = pl;
        for (int j = 0; j < 16; ++j) {                  // p^(L*(j+1))
            mc.plp_re[j][m] = (double)pj.re; mc.plp_im[j][m] = (double)pj.im;
            pj = cmul(pj, pl);
        }
        cld acc = { 1.0L, 0.0L };
        for (int j = 0; j < SCAN_CH; ++j) {             // p^(L*j)
            pw_table[((size_t)j * SCAN_NMAX + m) * 2] = (double)acc.re;
            pw_table[((size_t)j * SCAN_NMAX + m) * 2 + 1] = (double)acc.im;
            acc = cmul(acc, pl);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// device
// ---------------------------------------------------------------------------------------------
struct Cx { double re, im; };
__device__ __forceinline__ Cx cx_mul(double ar, double ai, Cx b)
{
    Cx r;
    r.re = fma(ar, b.re, -(ai * b.im));
    r.im = fma(ar, b.im, ai * b.re);
    return r;
}
// s <- m*s + u (u real)
__device__ __forceinline__ void cx_step(Cx &s, double mr, double mi, double u)
{
    const double nr = fma(mr, s.re, fma(-mi, s.im, u));
    const double ni = fma(mr, s.im, mi * s.re);
    s.re = nr; s.im = ni;
}

__device__ __forceinline__ double scan_sample(const DevChain &ch, const uint8_t *row, int64_t frame, int chan_off,
                                              int64_t pos0, bool fading)
{
    double x = unpack_real(ch.fmt, row + frame * ch.frame_bytes + chan_off, ch.aligned);
    if (fading) {
        double g = fade_gain(ch, pos0 + frame);
        if (g >= 0.0) x *= g;
    }
    return x;
}

// E layout: [stream][comp][chan][chunk], comp = (filter * SCAN_NMAX + mode) * 2 + {re, im}
__device__ __forceinline__ size_t e_index(int stream, int comp, int chan, int64_t chunk, int64_t n_chunks)
{
    return (((size_t)stream * (4 * SCAN_NMAX) + comp) * 2 + chan) * (size_t)n_chunks + (size_t)chunk;
}

// pass 1 + 2a: end state of every full chunk started from zero, then -- inside the CTA, which is
// exactly one tile of 128 chunks -- the carry scan over chunks: warp shuffles (16 chunks per warp),
// one shared-memory hop across the 8 warps.  Writes, per chunk, its carry-in from the tile start
// (exclusive prefix) and, per tile, the tile's end state.
template <int NM>
__global__ void __launch_bounds__(256)
scan_local_kernel(const __grid_constant__ ModalCoef mc, const __grid_constant__ DevChain ch,
                  const DevStream *__restrict__ streams, int64_t n_frames, int64_t n_chunks, int64_t n_tiles,
                  const uint8_t *__restrict__ in, size_t in_stride, double *__restrict__ E,
                  double *__restrict__ Tend /* [stream][tile][chan][filter][mode][2] */)
{
    __shared__ double wtot[8][2][2][SCAN_NMAX][2];      // inclusive total of each warp's 16 chunks
    __shared__ double wcar[8][2][2][SCAN_NMAX][2];      // carry into each warp from the tile start
    const int stream = blockIdx.y;
    const int64_t tile = blockIdx.x;
    const int cl = threadIdx.x >> 1;                    // chunk inside the tile, 0..127
    const int64_t chunk = tile * SCAN_CH + cl;
    const int chan = threadIdx.x & 1;
    const int warp = threadIdx.x >> 5, cw = cl & 15;    // chunk inside the warp
    const bool have = chunk < n_chunks - 1;             // the call's last chunk has no end state to hand on
    const DevStream &st = streams[stream];
    const unsigned q0 = st.quad[chan];
    const int xiq = q0 & 1;                             // filter fed on the first sample of a pair

    Cx s[2][NM];                                        // [filter I/Q][mode]
#pragma unroll
    for (int m = 0; m < NM; ++m) { s[0][m].re = s[0][m].im = 0.0; s[1][m].re = s[1][m].im = 0.0; }
    if (have) {
        const uint8_t *row = in + (size_t)stream * in_stride;
        const int chan_off = (ch.n_channels > 1 ? chan : 0) * ch.chan_bytes;
        const bool fading = (ch.n_fade_in | ch.n_fade_out) != 0;
        const int64_t f0 = chunk * SCAN_L;
        Cx A[NM], B[NM];
#pragma unroll
        for (int m = 0; m < NM; ++m) { A[m].re = A[m].im = 0.0; B[m].re = B[m].im = 0.0; }
        for (int k = 0; k < SCAN_L; k += 2) {
            const double x0 = scan_sample(ch, row, f0 + k, chan_off, st.pos, fading);
            const double x1 = scan_sample(ch, row, f0 + k + 1, chan_off, st.pos, fading);
            const unsigned qa = (q0 + (unsigned)k) & 3u;        // f0 is a multiple of 4
            const double ux = mix_down(xiq, qa, x0);
            const double uy = mix_down(xiq ^ 1, (qa + 1) & 3u, x1);
#pragma unroll
            for (int m = 0; m < NM; ++m) {
                cx_step(A[m], mc.p2_re[m], mc.p2_im[m], ux);
                cx_step(B[m], mc.p2_re[m], mc.p2_im[m], uy);
            }
        }
        // state after the last (second-of-pair) sample: X filter p*A, Y filter B
#pragma unroll
        for (int m = 0; m < NM; ++m) {
            Cx sx = cx_mul(mc.p_re[m], mc.p_im[m], A[m]);
            if (xiq == 0) { s[0][m] = sx; s[1][m] = B[m]; }
            else          { s[1][m] = sx; s[0][m] = B[m]; }
        }
    }

    // inclusive scan over the warp's 16 chunks (lanes of one channel are 2 apart)
#pragma unroll
    for (int d = 1; d < 16; d <<= 1) {
#pragma unroll
        for (int f = 0; f < 2; ++f)
#pragma unroll
            for (int m = 0; m < NM; ++m) {
                const double ore = __shfl_up_sync(0xffffffffu, s[f][m].re, 2 * d);
                const double oim = __shfl_up_sync(0xffffffffu, s[f][m].im, 2 * d);
                if (cw >= d) {
                    Cx o; o.re = ore; o.im = oim;
                    Cx w = cx_mul(mc.plp_re[d - 1][m], mc.plp_im[d - 1][m], o);     // p^(L*d) * earlier
                    s[f][m].re += w.re; s[f][m].im += w.im;
                }
            }
    }
    if (cw == 15) {
#pragma unroll
        for (int f = 0; f < 2; ++f)
#pragma unroll
            for (int m = 0; m < NM; ++m) { wtot[warp][chan][f][m][0] = s[f][m].re; wtot[warp][chan][f][m][1] = s[f][m].im; }
    }
    __syncthreads();
    // carry into each warp: sequential over the 8 warps, one thread per (chan, filter, mode)
    if (threadIdx.x < 2 * 2 * NM) {
        int r = threadIdx.x;
        const int m = r % NM; r /= NM;
        const int f = r & 1, c = r >> 1;
        Cx acc; acc.re = acc.im = 0.0;
        for (int w = 0; w < 8; ++w) {
            wcar[w][c][f][m][0] = acc.re; wcar[w][c][f][m][1] = acc.im;
            // acc <- p^(16 L) * acc + total(w)
            const double nr = fma(mc.plp_re[15][m], acc.re, fma(-mc.plp_im[15][m], acc.im, wtot[w][c][f][m][0]));
            const double ni = fma(mc.plp_re[15][m], acc.im, fma(mc.plp_im[15][m], acc.re, wtot[w][c][f][m][1]));
            acc.re = nr; acc.im = ni;
        }
    }
    __syncthreads();
    // inclusive prefix from the tile start, then shift by one chunk for the carry-IN of this chunk
#pragma unroll
    for (int f = 0; f < 2; ++f)
#pragma unroll
        for (int m = 0; m < NM; ++m) {
            Cx c; c.re = wcar[warp][chan][f][m][0]; c.im = wcar[warp][chan][f][m][1];
            Cx w = cx_mul(mc.plp_re[cw][m], mc.plp_im[cw][m], c);                   // p^(L*(cw+1)) * warp carry
            const double inc_re = s[f][m].re + w.re, inc_im = s[f][m].im + w.im;
            double ex_re = __shfl_up_sync(0xffffffffu, inc_re, 2);
            double ex_im = __shfl_up_sync(0xffffffffu, inc_im, 2);
            if (cw == 0) { ex_re = c.re; ex_im = c.im; }
            if (chunk < n_chunks) {
                const int comp = (f * SCAN_NMAX + m) * 2;
                E[e_index(stream, comp, chan, chunk, n_chunks)] = ex_re;
                E[e_index(stream, comp + 1, chan, chunk, n_chunks)] = ex_im;
            }
            if (cl == SCAN_CH - 1) {
                double *t = Tend + ((((size_t)stream * n_tiles + tile) * 2 + chan) * 2 + f) * (SCAN_NMAX * 2) + m * 2;
                t[0] = inc_re; t[1] = inc_im;
            }
        }
}

// pass 2b: carry into every tile.  The filters forget: |p|^(tile) <= 1e-3, so eight tiles back is
// below 1e-24 and the sum over the previous SCAN_W tile end states (Horner in p^tile) is exact to
// rounding.  Index -1 stands for the stream's state before this call.
constexpr int SCAN_W = 8;
__global__ void __launch_bounds__(128)
scan_tile_carry_kernel(const __grid_constant__ ModalCoef mc, const DevStream *__restrict__ streams, int64_t n_tiles,
                       const double *__restrict__ Tend, double *__restrict__ Tin)
{
    const int stream = blockIdx.y;
    const int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int per_tile = 2 * 2 * mc.nm;
    if (id >= n_tiles * per_tile) return;
    const int64_t tile = id / per_tile;
    int r = (int)(id % per_tile);
    const int m = r % mc.nm; r /= mc.nm;
    const int f = r & 1, chan = r >> 1;
    const double tr = mc.pt_re[m], ti = mc.pt_im[m];
    Cx acc; acc.re = acc.im = 0.0;
    for (int i = SCAN_W; i >= 1; --i) {
        const int64_t src = tile - i;
        Cx e; e.re = e.im = 0.0;
        if (src >= 0) {
            const double *t = Tend + ((((size_t)stream * n_tiles + src) * 2 + chan) * 2 + f) * (SCAN_NMAX * 2) + m * 2;
            e.re = t[0]; e.im = t[1];
        } else if (src == -1) {
            e.re = streams[stream].hb[chan][f][2 * m];
            e.im = streams[stream].hb[chan][f][2 * m + 1];
        }
        const double nr = fma(tr, acc.re, fma(-ti, acc.im, e.re));
        const double ni = fma(tr, acc.im, fma(ti, acc.re, e.im));
        acc.re = nr; acc.im = ni;
    }
    double *o = Tin + ((((size_t)stream * n_tiles + tile) * 2 + chan) * 2 + f) * (SCAN_NMAX * 2) + m * 2;
    o[0] = acc.re; o[1] = acc.im;
}

// pass 3: every chunk again, from its true initial state, producing the analytic signal
template <int NM>
__global__ void __launch_bounds__(256)
scan_apply_kernel(const __grid_constant__ ModalCoef mc, const __grid_constant__ DevChain ch,
                  DevStream *__restrict__ streams, int64_t n_frames, int64_t n_chunks, int64_t n_tiles,
                  const uint8_t *__restrict__ in, size_t in_stride,
                  const double *__restrict__ E, const double *__restrict__ Tin, const double *__restrict__ pw,
                  double *__restrict__ analytic /* [stream][frame][4] */)
{
    const int stream = blockIdx.y;
    const int64_t chunk = (int64_t)blockIdx.x * (blockDim.x / 2) + (threadIdx.x >> 1);
    const int chan = threadIdx.x & 1;
    if (chunk >= n_chunks) return;
    DevStream &st = streams[stream];
    const unsigned q0 = st.quad[chan];
    const int xiq = q0 & 1, yiq = xiq ^ 1;
    const uint8_t *row = in + (size_t)stream * in_stride;
    const int chan_off = (ch.n_channels > 1 ? chan : 0) * ch.chan_bytes;
    const bool fading = (ch.n_fade_in | ch.n_fade_out) != 0;
    const int64_t f0 = chunk * SCAN_L;
    const int len = (int)((n_frames - f0 < SCAN_L) ? n_frames - f0 : SCAN_L);
    const int64_t tile = chunk / SCAN_CH;
    const int jl = (int)(chunk % SCAN_CH);
    double *dst = analytic + ((size_t)stream * (size_t)n_frames + (size_t)f0) * 4 + chan * 2;

    // carried state after the sample before this chunk: prefix inside the tile + p^(L*j) * tile carry
    Cx A[NM], B[NM];
#pragma unroll
    for (int m = 0; m < NM; ++m) {
        Cx s[2];
#pragma unroll
        for (int f = 0; f < 2; ++f) {
            const int comp = (f * SCAN_NMAX + m) * 2;
            const double *t = Tin + ((((size_t)stream * n_tiles + tile) * 2 + chan) * 2 + f) * (SCAN_NMAX * 2) + m * 2;
            Cx c; c.re = t[0]; c.im = t[1];
            const double wr = pw[((size_t)jl * SCAN_NMAX + m) * 2], wi = pw[((size_t)jl * SCAN_NMAX + m) * 2 + 1];
            Cx cw = cx_mul(wr, wi, c);
            s[f].re = E[e_index(stream, comp, chan, chunk, n_chunks)] + cw.re;
            s[f].im = E[e_index(stream, comp + 1, chan, chunk, n_chunks)] + cw.im;
        }
        // the X filter last saw input two samples ago: step it back one sample (exact algebra: A = s / p)
        A[m] = cx_mul(mc.pinv_re[m], mc.pinv_im[m], s[xiq]);
        B[m] = s[yiq];
    }

    const int npair = len >> 1;
    for (int k2 = 0; k2 < npair; ++k2) {
        const int k = 2 * k2;
        const double x0 = scan_sample(ch, row, f0 + k, chan_off, st.pos, fading);
        const double x1 = scan_sample(ch, row, f0 + k + 1, chan_off, st.pos, fading);
        const unsigned qa = (q0 + (unsigned)k) & 3u, qb = (qa + 1) & 3u;
        const double ux = mix_down(xiq, qa, x0);
        const double uy = mix_down(yiq, qb, x1);
        double yx1 = mc.baseline ? mc.d0 * ux : 0.0, yy1 = 0.0, yx2 = 0.0, yy2 = mc.baseline ? mc.d0 * uy : 0.0;
#pragma unroll
        for (int m = 0; m < NM; ++m) {
            yx1 = fma(mc.cp_re[m], A[m].re, fma(mc.cp_im[m], A[m].im, yx1));    // X one sample after its input: r*p
            yy1 = fma(mc.c_re[m], B[m].re, fma(mc.c_im[m], B[m].im, yy1));
            yy2 = fma(mc.cp_re[m], B[m].re, fma(mc.cp_im[m], B[m].im, yy2));
            cx_step(A[m], mc.p2_re[m], mc.p2_im[m], ux);
            yx2 = fma(mc.c_re[m], A[m].re, fma(mc.c_im[m], A[m].im, yx2));
            cx_step(B[m], mc.p2_re[m], mc.p2_im[m], uy);
        }
        // up-mix (reference lpf_hilbert_quad.c:132-153): X filter is I when xiq == 0
        int slot;
        double v2[2];
        double a = mix_up(xiq, qa, yx1, slot); v2[slot] = a;
        double b = mix_up(yiq, qa, yy1, slot); v2[slot] = b;
        *reinterpret_cast<double2 *>(dst + (size_t)k * 4) = make_double2(v2[0], v2[1]);
        a = mix_up(xiq, qb, yx2, slot); v2[slot] = a;
        b = mix_up(yiq, qb, yy2, slot); v2[slot] = b;
        *reinterpret_cast<double2 *>(dst + (size_t)(k + 1) * 4) = make_double2(v2[0], v2[1]);
    }
    const bool odd = len & 1;
    if (odd) {
        const int k = len - 1;
        const double x0 = scan_sample(ch, row, f0 + k, chan_off, st.pos, fading);
        const unsigned qa = (q0 + (unsigned)k) & 3u;
        const double ux = mix_down(xiq, qa, x0);
        double yx1 = mc.baseline ? mc.d0 * ux : 0.0, yy1 = 0.0;
#pragma unroll
        for (int m = 0; m < NM; ++m) {
            yx1 = fma(mc.cp_re[m], A[m].re, fma(mc.cp_im[m], A[m].im, yx1));
            yy1 = fma(mc.c_re[m], B[m].re, fma(mc.c_im[m], B[m].im, yy1));
            cx_step(A[m], mc.p2_re[m], mc.p2_im[m], ux);
        }
        int slot;
        double v2[2];
        double a = mix_up(xiq, qa, yx1, slot); v2[slot] = a;
        double b = mix_up(yiq, qa, yy1, slot); v2[slot] = b;
        *reinterpret_cast<double2 *>(dst + (size_t)k * 4) = make_double2(v2[0], v2[1]);
    }
    if (chunk == n_chunks - 1) {
        // state after the call's last sample, filed under I / Q
#pragma unroll
        for (int m = 0; m < NM; ++m) {
            Cx sx, sy;
            if (odd) { sx = A[m]; sy = cx_mul(mc.p_re[m], mc.p_im[m], B[m]); }     // last sample fed X
            else     { sx = cx_mul(mc.p_re[m], mc.p_im[m], A[m]); sy = B[m]; }     // last sample fed Y
            st.hb[chan][xiq][2 * m] = sx.re; st.hb[chan][xiq][2 * m + 1] = sx.im;
            st.hb[chan][yiq][2 * m] = sy.re; st.hb[chan][yiq][2 * m + 1] = sy.im;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// launch sequence
// ---------------------------------------------------------------------------------------------
size_t scan_scratch_doubles(int n_streams, int64_t n_frames)
{
    const int64_t n_chunks = (n_frames + SCAN_L - 1) / SCAN_L;
    const int64_t n_tiles = (n_chunks + SCAN_CH - 1) / SCAN_CH;
    const size_t e = (size_t)n_streams * (4 * SCAN_NMAX) * 2 * (size_t)n_chunks;
    const size_t t = (size_t)n_streams * (size_t)n_tiles * 2 * 2 * (SCAN_NMAX * 2);
    return e + 2 * t;
}

template <int NM>
static cudaError_t scan_launch_nm(const ModalCoef &mc, const DevChain &ch, DevStream *streams, int n_streams,
                                  int64_t n_frames, const uint8_t *in, size_t in_stride, const double *pw,
                                  double *scratch, double *analytic, cudaStream_t s, int *launches)
{
    const int64_t n_chunks = (n_frames + SCAN_L - 1) / SCAN_L;
    const int64_t n_tiles = (n_chunks + SCAN_CH - 1) / SCAN_CH;
    double *E = scratch;
    double *Tend = E + (size_t)n_streams * (4 * SCAN_NMAX) * 2 * (size_t)n_chunks;
    double *Tin = Tend + (size_t)n_streams * (size_t)n_tiles * 2 * 2 * (SCAN_NMAX * 2);
    const unsigned cgrid = (unsigned)n_tiles;                   // one CTA = one tile of 128 chunks x 2 channels
    scan_local_kernel<NM><<<dim3(cgrid, n_streams), 256, 0, s>>>(mc, ch, streams, n_frames, n_chunks, n_tiles, in, in_stride, E, Tend);
    const int64_t items = n_tiles * 2 * 2 * mc.nm;
    const unsigned tgrid = (unsigned)((items + 127) / 128);
    scan_tile_carry_kernel<<<dim3(tgrid, n_streams), 128, 0, s>>>(mc, streams, n_tiles, Tend, Tin);
    scan_apply_kernel<NM><<<dim3(cgrid, n_streams), 256, 0, s>>>(mc, ch, streams, n_frames, n_chunks, n_tiles, in, in_stride,
                                                               E, Tin, pw, analytic);
    *launches += 3;
    return cudaGetLastError();
}

cudaError_t launch_hb_scan(const ModalCoef &mc, const DevChain &ch, DevStream *streams, int n_streams,
                           int64_t n_frames, const uint8_t *in, size_t in_stride, const double *pw,
                           double *scratch, double *analytic, cudaStream_t s, int *launches)
{
    switch (mc.nm) {
    case 8:  return scan_launch_nm<8>(mc, ch, streams, n_streams, n_frames, in, in_stride, pw, scratch, analytic, s, launches);
    case 9:  return scan_launch_nm<9>(mc, ch, streams, n_streams, n_frames, in, in_stride, pw, scratch, analytic, s, launches);
    case 10: return scan_launch_nm<10>(mc, ch, streams, n_streams, n_frames, in, in_stride, pw, scratch, analytic, s, launches);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace icw
