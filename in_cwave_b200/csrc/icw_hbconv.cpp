// icw_hbconv.cpp -- the state of one half-band filter in its two bases, converted on the host.
//
// Exact mode carries the reference's own DF-II delay line z[j] = w[n-j], j = 0..ord-1 (reference IIR_RAT_POLY.pz,
// src/hblpf.h:115-127; w[n] = u[n] + sum_i fb_i w[n-1-i], src/hblpf.c:894-926).  Scan mode carries one complex
// one-pole state per conjugate pair, s_k[n] = p_k s_k[n-1] + u[n] (icw_scan.cu).  Both are driven by the same
// input u, and  1 / A(z) = 1 / prod_j (1 - p_j z^-1),  so
//
//       S_k(z) = U(z) / (1 - p_k z^-1) = W(z) * prod_{j != k} (1 - p_j z^-1) =: W(z) Q_k(z)
//       s_k[n] = sum_{i=0}^{ord-1} q_{k,i} z[i]                                      (delay line -> modal)
//
// and the other direction is the inverse of that ord x ord real matrix.  |w| is ~1e10 x the filter's output for
// these elliptic designs (finding 3), so the sums cancel ten digits: everything here runs in
// IEEE binary128 (__float128, plain + - * / from libgcc) on poles known to double-double accuracy, and rounds once
// at the end.  What stays is the rounding already IN a delay line (1 ulp of w ~ 1e-6 of the output): a stream that
// switches bases continues within the reference's own noise of either mode, it does not become bit-exact.
//
// This lets a live stream change hilbert_mode (the reference changes filter settings on a live stream too,
// src/in_cwave.c:135-191) and lets a time shard start from an exact-mode state.  Plain g++: nvcc's front end
// does not take __float128.
#include <cstring>
#include <mutex>

#include "../../include/icw_b200.h"
#include "icw_hb_modal.inc"

namespace {

typedef __float128 q;
struct cq { q re, im; };
inline cq cmul(cq a, cq b) { return { a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re }; }
inline q qabs(q v) { return v < 0 ? -v : v; }

const int HB_ORDER[6] = { 15, 19, 18, 19, 20, 20 };     // reference src/hblpf.c:740-820

struct Maps {
    bool ready = false;
    int ord = 0;
    q fwd[ICW_MAX_ORD][ICW_MAX_ORD];    // modal[r] = sum_i fwd[r][i] * z[i];  r indexes hb[] of the modal basis
    q inv[ICW_MAX_ORD][ICW_MAX_ORD];    // z[i] = sum_r inv[i][rows[r]] ...  (see build)
    int rows[ICW_MAX_ORD];              // the `ord` modal slots in use (a real pole has no Im slot)
};
Maps g_maps[6];
std::mutex g_mu;

bool build(int ft, Maps &M)
{
    const int nm = ICW_HB_NMODES[ft];
    cq poles[ICW_MAX_ORD];
    int owner[ICW_MAX_ORD];             // index into poles[] of mode m's own (Im > 0 or real) pole
    int np = 0;
    for (int m = 0; m < nm; ++m) {
        const icw_hb_mode &md = ICW_HB_MODES[ft][m];
        cq p = { (q)md.p_re[0] + (q)md.p_re[1], (q)md.p_im[0] + (q)md.p_im[1] };
        owner[m] = np;
        poles[np++] = p;
        if (!md.is_real) { cq c = { p.re, -p.im }; poles[np++] = c; }
    }
    if (np != HB_ORDER[ft]) return false;
    M.ord = np;
    memset(M.fwd, 0, sizeof M.fwd);
    int nrows = 0;
    for (int m = 0; m < nm; ++m) {
        // Q_m(x) = prod_{j != owner} (1 - p_j x), degree ord - 1
        cq poly[ICW_MAX_ORD + 1];
        for (int i = 0; i <= np; ++i) poly[i] = { 0, 0 };
        poly[0] = { 1, 0 };
        int deg = 0;
        for (int j = 0; j < np; ++j) {
            if (j == owner[m]) continue;
            for (int i = deg + 1; i >= 1; --i) {
                cq t = cmul(poles[j], poly[i - 1]);
                poly[i].re -= t.re; poly[i].im -= t.im;
            }
            ++deg;
        }
        const bool real_mode = ICW_HB_MODES[ft][m].is_real != 0;
        for (int i = 0; i < np; ++i) {
            M.fwd[2 * m][i] = poly[i].re;
            if (!real_mode) M.fwd[2 * m + 1][i] = poly[i].im;
        }
        M.rows[nrows++] = 2 * m;
        if (!real_mode) M.rows[nrows++] = 2 * m + 1;
    }
    if (nrows != np) return false;
    // inverse of the square system A[r][i] = fwd[rows[r]][i] by Gauss-Jordan with partial pivoting
    q a[ICW_MAX_ORD][2 * ICW_MAX_ORD];
    for (int r = 0; r < np; ++r)
        for (int i = 0; i < np; ++i) { a[r][i] = M.fwd[M.rows[r]][i]; a[r][np + i] = (r == i) ? 1 : 0; }
    for (int c = 0; c < np; ++c) {
        int piv = c;
        for (int r = c + 1; r < np; ++r) if (qabs(a[r][c]) > qabs(a[piv][c])) piv = r;
        if (a[piv][c] == 0) return false;
        if (piv != c) for (int i = 0; i < 2 * np; ++i) { q t = a[c][i]; a[c][i] = a[piv][i]; a[piv][i] = t; }
        const q d = a[c][c];
        for (int i = 0; i < 2 * np; ++i) a[c][i] /= d;
        for (int r = 0; r < np; ++r) {
            if (r == c) continue;
            const q f = a[r][c];
            if (f == 0) continue;
            for (int i = 0; i < 2 * np; ++i) a[r][i] -= f * a[c][i];
        }
    }
    memset(M.inv, 0, sizeof M.inv);
    for (int i = 0; i < np; ++i)
        for (int r = 0; r < np; ++r) M.inv[i][r] = a[i][np + r];      // z[i] = sum_r inv[i][r] * modal[rows[r]]
    M.ready = true;
    return true;
}

}  // namespace

// in / out: one filter's ICW_MAX_ORD doubles (icw_stream_state.hb[channel][filter]); to_basis 1 = delay line -> modal,
// 0 = modal -> delay line.  Host only, no GPU.  Returns 0, or -1 on a bad argument.
extern "C" int icw_host_hb_convert(int filter_no, int to_basis, const double *in, double *out)
{
    if (filter_no < 0 || filter_no >= 6 || !in || !out || (to_basis != 0 && to_basis != 1)) return -1;
    Maps &M = g_maps[filter_no];
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (!M.ready && !build(filter_no, M)) return -1;
    }
    double res[ICW_MAX_ORD];
    for (int i = 0; i < ICW_MAX_ORD; ++i) res[i] = 0.0;
    if (to_basis == 1) {
        for (int r = 0; r < M.ord; ++r) {
            q acc = 0;
            for (int i = 0; i < M.ord; ++i) acc += M.fwd[M.rows[r]][i] * (q)in[i];
            res[M.rows[r]] = (double)acc;
        }
    } else {
        for (int i = 0; i < M.ord; ++i) {
            q acc = 0;
            for (int r = 0; r < M.ord; ++r) acc += M.inv[i][r] * (q)in[M.rows[r]];
            res[i] = (double)acc;
        }
    }
    memcpy(out, res, sizeof res);
    return 0;
}
