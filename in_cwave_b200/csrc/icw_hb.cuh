// icw_hb.cuh -- the half-band recurrences in the reference's own operation order
// (reference src/hblpf.c:894-926 baseline, :974-1057 Kahan) and the fs/4 mixer around them
// (reference src/lpf_hilbert_quad.c:129-156).  Shared by the leaf, unfused and fused kernels.
#pragma once
#include "icw_dev.cuh"
#include "icw_kernels.h"

namespace icw {

struct Comp { double s, c; };
__device__ __forceinline__ void comp_add(Comp &a, double x)
{
    // hblpf.c:991-997 -- y = x - c; t = s + y; c = (t - s) - y; s = t
    double y = __dsub_rn(x, a.c);
    double t = __dadd_rn(a.s, y);
    a.c = __dsub_rn(__dsub_rn(t, a.s), y);
    a.s = t;
}
// three (two) independent compensated sums advanced in lock step, written stage by stage so that
// the instruction stream alternates between the chains: every DADD has an 8-cycle latency and
// the FP64 pipe takes one warp instruction per 2 cycles, so one chain alone idles it 3/4 of the time
__device__ __forceinline__ void comp_add3(Comp &a, double xa, Comp &b, double xb, Comp &c, double xc)
{
    double ya = __dsub_rn(xa, a.c), yb = __dsub_rn(xb, b.c), yc = __dsub_rn(xc, c.c);
    double ta = __dadd_rn(a.s, ya), tb = __dadd_rn(b.s, yb), tc = __dadd_rn(c.s, yc);
    double da = __dsub_rn(ta, a.s), db = __dsub_rn(tb, b.s), dc = __dsub_rn(tc, c.s);
    a.c = __dsub_rn(da, ya); b.c = __dsub_rn(db, yb); c.c = __dsub_rn(dc, yc);
    a.s = ta; b.s = tb; c.s = tc;
}
__device__ __forceinline__ void comp_add2(Comp &a, double xa, Comp &b, double xb)
{
    double ya = __dsub_rn(xa, a.c), yb = __dsub_rn(xb, b.c);
    double ta = __dadd_rn(a.s, ya), tb = __dadd_rn(b.s, yb);
    double da = __dsub_rn(ta, a.s), db = __dsub_rn(tb, b.s);
    a.c = __dsub_rn(da, ya); b.c = __dsub_rn(db, yb);
    a.s = ta; b.s = tb;
}

// part [I0, I1) of the output sum of the sample whose states are z[OFF], z[OFF+1], ...
template <int ORD, int I0, int I1, int OFF>
__device__ __forceinline__ void out_sum_part(Comp &o, const double (&z)[ORD + 2], const HbCoef &k)
{
#pragma unroll
    for (int i = I0; i < I1; ++i) {
        const double zi = z[i + OFF];
        const double t = __dmul_rn(zi, k.fb[i]);
        if (i == 0) { o.s = __dmul_rn(zi, k.ff[0]); o.c = 0.0; }
        else comp_add(o, __dmul_rn(zi, k.ff[i]));
        comp_add(o, __dmul_rn(t, k.d0));
    }
}

// ---------------------------------------------------------------------------------------------
// Kahan variant (the reference's default).  Per sample the reference forms two compensated sums:
// the state sum (x, z0*fb0, z1*fb1, ...: 4 dependent DADDs per term, 76 in a row for order 19)
// and the output sum (z0*ff0, (z0*fb0)*d0, z1*ff1, ...: 148 in a row).  Only the first feeds the
// next sample, so step() evaluates the output sum of sample i-1 (first half) and of sample i-2
// (second half) while the state sum of sample i runs: three independent dependency chains of
// ~75 DADDs per step instead of one of ~225.  Every operation and its order inside each sum is the
// reference's; only the interleaving in time differs, which rounding cannot see.
// z[j] = state j+1 samples ago (newest first: the order the reference's circular walk visits).
// ---------------------------------------------------------------------------------------------
template <int ORD>
struct KahanChain {
    static constexpr int LAG = 2;           // step(i) returns the output of sample i-2
    static constexpr int H = (ORD + 1) / 2;
    double z[ORD + 2];
    Comp o_half;

    __device__ __forceinline__ void load(const double *zs)
    {
#pragma unroll
        for (int i = 0; i < ORD; ++i) z[i] = zs[i];
        z[ORD] = z[ORD + 1] = 0.0;          // only reach outputs of samples before this call
        o_half.s = 0.0; o_half.c = 0.0;
    }
    __device__ __forceinline__ void store(double *zs) const
    {
#pragma unroll
        for (int i = 0; i < ORD; ++i) zs[i] = z[i];
    }

    __device__ __forceinline__ double step(double x, const HbCoef &k, int reject, double thr,
                                           unsigned long long &rejects)
    {
        // term lists in the reference's order:
        //   state sum a:                      x, z0*fb0, z1*fb1, ...                     (ORD additions)
        //   output sum of sample i-1, 1st half (o_new):  z1*ff0 | t0*d0, z2*ff1, t1*d0, ...
        //   output sum of sample i-2, 2nd half (o_half): z[H+2]*ffH, tH*d0, ...
        Comp a; a.s = x; a.c = 0.0;
        Comp o_new;
        constexpr int NA = ORD;                 // additions into a
        constexpr int NB = 2 * H - 1;           // additions into o_new (its first term initialises)
        constexpr int NC = 2 * (ORD - H);       // additions into o_half
        const double tb0 = __dmul_rn(z[1], k.fb[0]);
        o_new.s = __dmul_rn(z[1], k.ff[0]); o_new.c = 0.0;
        constexpr int NSTEP = NA > NB ? (NA > NC ? NA : NC) : (NB > NC ? NB : NC);
#pragma unroll
        for (int st = 0; st < NSTEP; ++st) {
            const double xa = st < NA ? __dmul_rn(z[st < NA ? st : 0], k.fb[st < NA ? st : 0]) : 0.0;
            const int ib = (st + 1) / 2;
            const int ibc = ib < H ? ib : 0;
            const double xb = st < NB
                ? ((st & 1) == 0 ? __dmul_rn(st == 0 ? tb0 : __dmul_rn(z[ibc + 1], k.fb[ibc]), k.d0)
                                 : __dmul_rn(z[ibc + 1], k.ff[ibc]))
                : 0.0;
            const int ic = H + st / 2;
            const int icc = ic < ORD ? ic : H;
            const double xc = st < NC
                ? ((st & 1) == 0 ? __dmul_rn(z[icc + 2], k.ff[icc])
                                 : __dmul_rn(__dmul_rn(z[icc + 2], k.fb[icc]), k.d0))
                : 0.0;
            if (st < NA && st < NB && st < NC) comp_add3(a, xa, o_new, xb, o_half, xc);
            else if (st < NA && st < NB) comp_add2(a, xa, o_new, xb);
            else if (st < NA && st < NC) comp_add2(a, xa, o_half, xc);
            else if (st < NB && st < NC) comp_add2(o_new, xb, o_half, xc);
            else if (st < NA) comp_add(a, xa);
            else if (st < NB) comp_add(o_new, xb);
            else if (st < NC) comp_add(o_half, xc);
        }
        double w = a.s;
        const bool rj = (reject != 0) & (fabs(w) < thr);    // |w| < flag value: bug-for-bug (hblpf.c:1046)
        w = rj ? 0.0 : w;
        rejects += rj ? 1ull : 0ull;
        const double y = o_half.s;                          // no d0*x term: bug-for-bug (hblpf.c:1056)
        o_half = o_new;
#pragma unroll
        for (int j = ORD + 1; j > 0; --j) z[j] = z[j - 1];
        z[0] = w;
        return y;
    }

    // after the last step(n-1): the outputs of samples n-2 and n-1 (no shift in between: both read z[1..])
    __device__ __forceinline__ void drain(const HbCoef &k, double &y_nm2, double &y_nm1)
    {
        Comp o_new; o_new.s = 0.0; o_new.c = 0.0;
        out_sum_part<ORD, 0, H, 1>(o_new, z, k);
        out_sum_part<ORD, H, ORD, 2>(o_half, z, k);
        y_nm2 = o_half.s;
        out_sum_part<ORD, H, ORD, 1>(o_new, z, k);
        y_nm1 = o_new.s;
    }
};

// baseline summation (reference src/hblpf.c:894-926): two short chains, no lag
template <int ORD>
struct PlainChain {
    static constexpr int LAG = 0;
    double z[ORD];
    __device__ __forceinline__ void load(const double *zs)
    {
#pragma unroll
        for (int i = 0; i < ORD; ++i) z[i] = zs[i];
    }
    __device__ __forceinline__ void store(double *zs) const
    {
#pragma unroll
        for (int i = 0; i < ORD; ++i) zs[i] = z[i];
    }
    __device__ __forceinline__ double step(double x, const HbCoef &k, int reject, double thr,
                                           unsigned long long &rejects)
    {
        double acc_in = x, acc_out = 0.0;
#pragma unroll
        for (int j = 0; j < ORD; ++j) {
            acc_in = __dadd_rn(acc_in, __dmul_rn(z[j], k.fb[j]));
            acc_out = __dadd_rn(acc_out, __dmul_rn(z[j], k.ff[j]));
        }
        double w = acc_in;
        const bool rj = (reject != 0) & (fabs(w) < thr);
        w = rj ? 0.0 : w;
        rejects += rj ? 1ull : 0ull;
        const double y = __dadd_rn(__dmul_rn(w, k.d0), acc_out);
#pragma unroll
        for (int j = ORD - 1; j > 0; --j) z[j] = z[j - 1];
        z[0] = w;
        return y;
    }
    __device__ __forceinline__ void drain(const HbCoef &, double &, double &) {}
};

// The FP-exception-checked twins (reference src/hblpf.c:928-950 baseline, :1059-1096 Kahan): the same sums
// with FC() around every product and every partial sum, written the way the reference walks them (no lag, no
// interleaving -- a diagnostic mode).  cnt = the FP_EXCEPT_STATS block of this channel's converter.
template <int ORD, bool KAHAN>
struct CheckedChain {
    static constexpr int LAG = 0;
    double z[ORD];
    uint32_t *cnt = nullptr;
    __device__ __forceinline__ void load(const double *zs)
    {
#pragma unroll
        for (int i = 0; i < ORD; ++i) z[i] = zs[i];
    }
    __device__ __forceinline__ void store(double *zs) const
    {
#pragma unroll
        for (int i = 0; i < ORD; ++i) zs[i] = z[i];
    }
    __device__ __forceinline__ void kadd(Comp &a, double x)
    {
        // kahan_step_fes, hblpf.c:995-1003
        const double y = fc(__dsub_rn(x, a.c), cnt);
        const double t = fc(__dadd_rn(a.s, y), cnt);
        a.c = fc(__dsub_rn(fc(__dsub_rn(t, a.s), cnt), y), cnt);
        a.s = t;
    }
    __device__ __forceinline__ double step(double x, const HbCoef &k, int reject, double thr, unsigned long long &rejects)
    {
        double w, y = 0.0, acc_out = 0.0;
        if (KAHAN) {
            Comp in, out;
            in.s = x; in.c = 0.0;
            out.s = 0.0; out.c = 0.0;
#pragma unroll
            for (int j = 0; j < ORD; ++j) {
                const double t = fc(__dmul_rn(z[j], k.fb[j]), cnt);
                kadd(in, t);
                const double f = fc(__dmul_rn(z[j], k.ff[j]), cnt);
                if (j == 0) { out.s = f; out.c = 0.0; } else kadd(out, f);
                kadd(out, fc(__dmul_rn(t, k.d0), cnt));
            }
            w = in.s;
            y = out.s;                                          // no d0*x term: bug-for-bug (hblpf.c:1095)
        } else {
            double acc_in = x;
#pragma unroll
            for (int j = 0; j < ORD; ++j) {
                acc_in = fc(__dadd_rn(acc_in, fc(__dmul_rn(z[j], k.fb[j]), cnt)), cnt);
                acc_out = fc(__dadd_rn(acc_out, fc(__dmul_rn(z[j], k.ff[j]), cnt)), cnt);
            }
            w = acc_in;
        }
        const bool rj = (reject != 0) & (fabs(w) < thr);
        w = rj ? 0.0 : w;
        rejects += rj ? 1ull : 0ull;
        if (!KAHAN) y = fc(__dadd_rn(fc(__dmul_rn(w, k.d0), cnt), acc_out), cnt);      // hblpf.c:949
#pragma unroll
        for (int j = ORD - 1; j > 0; --j) z[j] = z[j - 1];
        z[0] = w;
        return y;
    }
    __device__ __forceinline__ void drain(const HbCoef &, double &, double &) {}
};

template <int ORD, bool KAHAN, bool CHECK = false> struct ChainSel { using type = KahanChain<ORD>; };
template <int ORD> struct ChainSel<ORD, false, false> { using type = PlainChain<ORD>; };
template <int ORD, bool KAHAN> struct ChainSel<ORD, KAHAN, true> { using type = CheckedChain<ORD, KAHAN>; };

// One half-band recurrence over n samples: `in(i)` gives the mixed-down input of sample i,
// `out(i, y)` receives the filter output of sample i.
template <class Chain> __device__ __forceinline__ void chain_set_counters(Chain &, uint32_t *) {}
template <int ORD, bool KAHAN> __device__ __forceinline__ void chain_set_counters(CheckedChain<ORD, KAHAN> &c, uint32_t *cnt) { c.cnt = cnt; }

template <int ORD, bool KAHAN, bool CHECK = false, class In, class Out>
__device__ __forceinline__ void hb_run(double *zstate, unsigned long long &rejects, const HbCoef &k,
                                       int reject, int64_t n, In in, Out out, uint32_t *fp_cnt = nullptr)
{
    using Chain = typename ChainSel<ORD, KAHAN, CHECK>::type;
    const double thr = (double)reject;
    if (n <= 0) return;
    Chain c;
    chain_set_counters(c, fp_cnt);
    c.load(zstate);
    double xc = in(0);
    for (int64_t i = 0; i < n; ++i) {
        const double xn = in(i + 1 < n ? i + 1 : i);        // the next input is fetched a sample ahead
        const double y = c.step(xc, k, reject, thr, rejects);
        if (i >= Chain::LAG) out(i - Chain::LAG, y);
        xc = xn;
    }
    if (Chain::LAG) {
        double y2, y1;
        c.drain(k, y2, y1);
        if (n >= 2) out(n - 2, y2);
        out(n - 1, y1);
    }
    c.store(zstate);
}

// fs/4 mixer around one filter (reference src/lpf_hilbert_quad.c:132-153).  Down-mix: the I filter
// gets (+x, 0, -x, 0), the Q filter (0, -x, 0, +x).  Up-mix and *2: the I filter feeds
// (+re, +im, -re, -im), the Q filter (+im, -re, -im, +re); slot 0 = re, 1 = im.
__device__ __forceinline__ double mix_down(int iq, unsigned q, double x)
{
    if (iq == 0) return (q == 0) ? x : (q == 2) ? -x : 0.0;
    return (q == 1) ? -x : (q == 3) ? x : 0.0;
}
__device__ __forceinline__ double mix_up(int iq, unsigned q, double y, int &slot)
{
    double v = __dmul_rn(y, 2.0);
    if (iq == 0) { slot = q & 1; if (q >= 2) v = -v; }
    else         { slot = (q & 1) ^ 1; if (q == 1 || q == 2) v = -v; }
    return v;
}

}  // namespace icw
