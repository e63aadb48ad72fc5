// icw_mt.cu -- MT19937 jump-ahead: GF(2) polynomial arithmetic on the host, application on the GPU.
// See icw_mt.h for the identity this relies on.  Stream facts follow the reference generator,
// src/mersene_twister/mt_jrnd.c:28-47 (seeding) and :99-134 (block regeneration + tempering).
#include "icw_mt.h"

#include <cstdio>
#include <cstring>

#include "icw_kernels.h"
#include "../../include/icw_b200.h"

namespace icw {

// =============================================================================================
// host: generator
// =============================================================================================
void mt_seed_state(uint32_t seed, uint32_t st[MT_N])
{
    uint32_t prev = seed;
    st[0] = prev;
    for (uint32_t j = 1; j < MT_N; ++j) {
        prev = 1812433253u * (prev ^ (prev >> 30)) + j;
        st[j] = prev;
    }
}

void mt_regen_host(uint32_t st[MT_N])
{
    for (int k = 0; k < MT_N; ++k) {
        uint32_t a = st[k], b = st[k + 1 < MT_N ? k + 1 : 0];
        uint32_t mix = (a & 0x80000000u) | (b & 0x7FFFFFFFu);
        uint32_t tw = (mix >> 1) ^ ((b & 1u) ? 0x9908B0DFu : 0u);
        st[k] = st[k + 397 < MT_N ? k + 397 : k + 397 - MT_N] ^ tw;
    }
}

// =============================================================================================
// host: characteristic polynomial by Berlekamp-Massey over GF(2) on the generator's own bits
// =============================================================================================
static std::vector<int> compute_charpoly_terms()
{
    const int N = 2 * MT_DEG + 128;                 // bits examined
    const int W = (N + 63) / 64 + 1;
    std::vector<uint8_t> bits((size_t)N);
    {
        uint32_t st[MT_N];
        mt_seed_state(5489u, st);
        int n = 0;
        while (n < N) {
            mt_regen_host(st);
            for (int i = 0; i < MT_N && n < N; ++i) bits[n++] = (uint8_t)(st[i] & 1u);
        }
    }
    std::vector<uint64_t> C((size_t)W, 0), B((size_t)W, 0), T((size_t)W, 0), R((size_t)W, 0);
    C[0] = 1; B[0] = 1;
    int L = 0, m = 1;
    for (int n = 0; n < N; ++n) {
        // R bit j = s[n - j]: shift in the new bit
        uint64_t carry = bits[n];
        const int rw = n / 64 + 1 < W ? n / 64 + 1 : W;
        for (int i = 0; i < rw; ++i) {
            uint64_t nc = R[i] >> 63;
            R[i] = (R[i] << 1) | carry;
            carry = nc;
        }
        uint64_t acc = 0;
        const int lw = L / 64 + 1;
        for (int i = 0; i < lw; ++i) acc ^= C[i] & R[i];
        int d = __builtin_parityll(acc);
        if (!d) { ++m; continue; }
        const bool grow = 2 * L <= n;
        if (grow) T = C;
        // C ^= B << m
        const int ws = m / 64, bs = m % 64;
        for (int i = W - 1; i >= ws; --i) {
            uint64_t v = B[i - ws] << bs;
            if (bs && i - ws - 1 >= 0) v |= B[i - ws - 1] >> (64 - bs);
            C[i] ^= v;
        }
        if (grow) { L = n + 1 - L; B.swap(T); m = 1; }
        else ++m;
    }
    std::vector<int> terms;
    if (L != MT_DEG) return terms;                   // caller reports the failure
    // connection polynomial C(x) = sum c_i x^i  <->  phi(x) = x^L * C(1/x): phi_{L-i} = c_i
    for (int i = L; i >= 0; --i)
        if ((C[i / 64] >> (i % 64)) & 1u) terms.push_back(L - i);
    return terms;                                    // ascending exponents, last = 19937
}

const std::vector<int> &mt_charpoly_terms()
{
    static const std::vector<int> terms = compute_charpoly_terms();
    return terms;
}

// =============================================================================================
// host: polynomials modulo phi
// =============================================================================================
void mt_poly_one(MtPoly &p) { memset(&p, 0, sizeof p); p.w[0] = 1; }

static inline void flip(uint64_t *w, int bit) { w[bit >> 6] ^= 1ull << (bit & 63); }
static inline int get(const uint64_t *w, int bit) { return (int)((w[bit >> 6] >> (bit & 63)) & 1u); }

// wide[0 .. 2*MT_PW) -> reduced into its low MT_DEG bits
static void reduce_wide(uint64_t *wide)
{
    const std::vector<int> &t = mt_charpoly_terms();
    for (int pos = 2 * MT_PW * 64 - 1; pos >= MT_DEG; --pos) {
        if (!get(wide, pos)) continue;
        const int sh = pos - MT_DEG;
        for (int e : t) flip(wide, sh + e);          // the top term clears bit `pos`
    }
}

void mt_poly_square(const MtPoly &a, MtPoly &out)
{
    // bit-spreading table, built once (function-local static: initialisation is thread-safe)
    struct Spread { uint16_t t[256]; Spread() { for (int v = 0; v < 256; ++v) { uint16_t s = 0; for (int b = 0; b < 8; ++b) if (v & (1 << b)) s |= (uint16_t)(1u << (2 * b)); t[v] = s; } } };
    static const Spread table;
    const uint16_t *spread = table.t;
    uint64_t wide[2 * MT_PW];
    for (int i = 0; i < MT_PW; ++i) {
        uint64_t v = a.w[i], lo = 0, hi = 0;
        for (int b = 0; b < 4; ++b) lo |= (uint64_t)spread[(v >> (8 * b)) & 0xFF] << (16 * b);
        for (int b = 0; b < 4; ++b) hi |= (uint64_t)spread[(v >> (32 + 8 * b)) & 0xFF] << (16 * b);
        wide[2 * i] = lo; wide[2 * i + 1] = hi;
    }
    reduce_wide(wide);
    memcpy(out.w, wide, sizeof out.w);
}

void mt_poly_mul(const MtPoly &a, const MtPoly &b, MtPoly &out)
{
    uint64_t wide[2 * MT_PW];
    memset(wide, 0, sizeof wide);
    for (int i = 0; i < MT_PW * 64; ++i) {
        if (!get(a.w, i)) continue;
        const int ws = i >> 6, bs = i & 63;
        for (int j = 0; j < MT_PW; ++j) {
            wide[ws + j] ^= b.w[j] << bs;
            if (bs) wide[ws + j + 1] ^= b.w[j] >> (64 - bs);
        }
    }
    reduce_wide(wide);
    memcpy(out.w, wide, sizeof out.w);
}

static void poly_mulx(MtPoly &p)
{
    uint64_t carry = 0;
    for (int i = 0; i < MT_PW; ++i) {
        uint64_t nc = p.w[i] >> 63;
        p.w[i] = (p.w[i] << 1) | carry;
        carry = nc;
    }
    if (get(p.w, MT_DEG))
        for (int e : mt_charpoly_terms()) flip(p.w, e);
}

void mt_poly_mulx_pow(MtPoly &p, uint64_t e) { while (e--) poly_mulx(p); }

void mt_poly_xpow(uint64_t e, MtPoly &out)
{
    mt_poly_one(out);
    if (!e) return;
    int top = 63;
    while (!((e >> top) & 1u)) --top;
    for (int b = top; b >= 0; --b) {
        MtPoly sq;
        mt_poly_square(out, sq);
        out = sq;
        if ((e >> b) & 1u) poly_mulx(out);
    }
}

void mt_apply_host(const MtPoly &g, const uint32_t base[MT_N], uint32_t out[MT_N])
{
    const int blocks = 33;
    std::vector<uint32_t> seq((size_t)blocks * MT_N);
    memcpy(seq.data(), base, MT_N * sizeof(uint32_t));
    for (int k = 0; k + MT_N < blocks * MT_N; ++k) {
        uint32_t a = seq[k], b = seq[k + 1];
        uint32_t mix = (a & 0x80000000u) | (b & 0x7FFFFFFFu);
        seq[k + MT_N] = seq[k + 397] ^ (mix >> 1) ^ ((b & 1u) ? 0x9908B0DFu : 0u);
    }
    for (int m = 0; m < MT_N; ++m) {
        uint32_t acc = 0;
        for (int i = 0; i < MT_DEG; ++i)
            if (get(g.w, i)) acc ^= seq[i + m];
        out[m] = acc;
    }
}

// =============================================================================================
// device: apply a jump polynomial to many states at once
// =============================================================================================
constexpr int JUMP_WARPS = 10;
constexpr int JUMP_THREADS = JUMP_WARPS * 32;
constexpr int JUMP_R = 20;                          // output words per lane: 32 x 20 = 640 >= 624
constexpr int JUMP_BLOCKS = 33;
constexpr int JUMP_SEQ = JUMP_BLOCKS * MT_N;        // 20592 words >= 19936 + 624
constexpr int JUMP_SEQ_PAD = 20608;                 // + the window overhang of the unused outputs 624..639
constexpr size_t JUMP_SMEM = (JUMP_SEQ_PAD + MT_N) * sizeof(uint32_t);  // 85 KB: two CTAs per SM

__device__ __forceinline__ uint32_t mt_twist(uint32_t a, uint32_t b)
{
    const uint32_t mix = (a & 0x80000000u) | (b & 0x7FFFFFFFu);
    return (mix >> 1) ^ ((b & 1u) ? 0x9908B0DFu : 0u);
}

// One block of the sequence from the one before it, by 227 threads without a barrier inside the
// block: u[k+624] = u[k+397] ^ twist(u[k], u[k+1]), and u[k+397] of the second and third 227 words
// is the word this same thread made one step earlier (k+397 = (k-227)+624).
__device__ __forceinline__ void mt_next_block(const uint32_t *old, uint32_t *nw, int t)
{
    if (t < 227) {
        const uint32_t n0 = old[t + 397] ^ mt_twist(old[t], old[t + 1]);
        const uint32_t n1 = n0 ^ mt_twist(old[t + 227], old[t + 228]);
        nw[t] = n0;
        nw[t + 227] = n1;
        if (t < 170) {
            // the block's last word pairs with the NEW word 0 (reference mt_jrnd.c:121)
            const uint32_t nxt = t == 169 ? (old[397] ^ mt_twist(old[0], old[1])) : old[t + 455];
            nw[t + 454] = n1 ^ mt_twist(old[t + 454], nxt);
        }
    }
}

// states[dst_first + blockIdx.x] = jump(states[dst_first + blockIdx.x - span]) by the polynomial g:
// out[m] = XOR over the set bits i of g of seq[i + m], seq = the source state continued for 32 blocks.
//
// Register-blocked GF(2) convolution.  A lane owns 20 consecutive outputs, a warp all 640 (624
// real); the warps of a CTA (and of the gridDim.y CTAs that share a target) take the 32-bit words of
// g round-robin.  For one word of g the lane loads the 52 sequence words its 20 outputs can touch
// (13 LDS.128, conflict-free at a lane stride of 20 words) and then works from registers: the
// bits are taken in pairs, and a pair costs 20 three-input LOP3 at most.  The shared-memory pipe
// that bounded the word-per-XOR version is out of the picture; the bound is the ALU pipe.
__global__ void __launch_bounds__(JUMP_THREADS, 2)
mt_jump_kernel(uint32_t *__restrict__ states, const uint32_t *__restrict__ src_states, int dst_first, int span,
               const uint32_t *__restrict__ poly, uint32_t *__restrict__ states_b = nullptr /* blockIdx.z == 1: a second
               generator's checkpoint array, same geometry (both channels' trees in one launch) */)
{
    if (blockIdx.z) states = states_b;
    extern __shared__ __align__(16) uint32_t sm[];
    uint32_t *seq = sm;
    uint32_t *pw = sm + JUMP_SEQ_PAD;
    const int tid = threadIdx.x;
    const int dst = dst_first + blockIdx.x;
    const uint32_t *src = src_states ? src_states + (size_t)blockIdx.x * MT_N
                                     : states + (size_t)(dst - span) * MT_N;
    for (int i = tid; i < MT_N; i += JUMP_THREADS) { seq[i] = src[i]; pw[i] = poly[i]; }
    if (tid < JUMP_SEQ_PAD - JUMP_SEQ) seq[JUMP_SEQ + tid] = 0u;
    __syncthreads();
    for (int b = 1; b < JUMP_BLOCKS; ++b) {
        mt_next_block(seq + (b - 1) * MT_N, seq + b * MT_N, tid);
        __syncthreads();
    }

    const int lane = tid & 31, warp = tid >> 5;
    uint32_t acc[JUMP_R];
#pragma unroll
    for (int r = 0; r < JUMP_R; ++r) acc[r] = 0u;
    const int stride = JUMP_WARPS * gridDim.y;
    for (int wi = blockIdx.y * JUMP_WARPS + warp; wi < MT_N; wi += stride) {
        const uint32_t bits = pw[wi];
        if (!bits) continue;
        uint32_t win[32 + JUMP_R];
        const uint4 *p = reinterpret_cast<const uint4 *>(seq + wi * 32 + lane * JUMP_R);
#pragma unroll
        for (int q = 0; q < (32 + JUMP_R) / 4; ++q) {
            const uint4 v = p[q];
            win[4 * q] = v.x; win[4 * q + 1] = v.y; win[4 * q + 2] = v.z; win[4 * q + 3] = v.w;
        }
#pragma unroll
        for (int b = 0; b < 32; b += 2) {
            const uint32_t two = (bits >> b) & 3u;
            if (two == 3u) {
#pragma unroll
                for (int r = 0; r < JUMP_R; ++r) acc[r] ^= win[b + r] ^ win[b + 1 + r];
            } else if (two == 1u) {
#pragma unroll
                for (int r = 0; r < JUMP_R; ++r) acc[r] ^= win[b + r];
            } else if (two == 2u) {
#pragma unroll
                for (int r = 0; r < JUMP_R; ++r) acc[r] ^= win[b + 1 + r];
            }
        }
    }
    // XOR the warps' partial results together (the sequence is no longer needed: reuse its space)
    __syncthreads();
    uint32_t *red = sm;                              // [warp][640]
    {
        uint4 *q = reinterpret_cast<uint4 *>(red + warp * (32 * JUMP_R) + lane * JUMP_R);
#pragma unroll
        for (int r = 0; r < JUMP_R; r += 4) q[r / 4] = make_uint4(acc[r], acc[r + 1], acc[r + 2], acc[r + 3]);
    }
    __syncthreads();
    for (int m = tid; m < MT_N; m += JUMP_THREADS) {
        uint32_t v = 0u;
#pragma unroll
        for (int w = 0; w < JUMP_WARPS; ++w) v ^= red[w * (32 * JUMP_R) + m];
        if (gridDim.y == 1) states[(size_t)dst * MT_N + m] = v;
        else atomicXor(&states[(size_t)dst * MT_N + m], v);         // the target was zeroed before the launch
    }
}

// few targets: the launch is latency-bound by one CTA's walk over the polynomial -> slice it
static int jump_slices(int count, int sm_count)
{
    int s = 1;
    while (s < 8 && count * s <= sm_count) s *= 2;
    return s;
}

void MtJump::release()
{
    for (uint32_t *p : dev_poly_) if (p) cudaFree(p);
    dev_poly_.clear();
    host_poly_.clear();
    for (auto &kv : any_poly_) if (kv.second.dev) cudaFree(kv.second.dev);
    any_poly_.clear();
    for (int l = 0; l < 2; ++l) { if (d_ckpt_[l]) cudaFree(d_ckpt_[l]); d_ckpt_[l] = nullptr; ckpt_cap_[l] = 0; }
    if (d_tmp_) cudaFree(d_tmp_);
    if (d_tail_) cudaFree(d_tail_);
    d_tmp_ = d_tail_ = nullptr;
    for (auto &t : tails_) t.valid = false;
}

int MtJump::ensure_poly(int k)
{
    if (mt_charpoly_terms().empty()) { err_ = "MT19937 characteristic polynomial: Berlekamp-Massey did not reach degree 19937"; return ICW_E_ARG; }
    while ((int)host_poly_.size() <= k) {
        MtPoly p;
        if (host_poly_.empty()) mt_poly_xpow(MT_N, p);              // x^624
        else mt_poly_square(host_poly_.back(), p);                   // (x^(624*2^(k-1)))^2
        host_poly_.push_back(p);
        dev_poly_.push_back(nullptr);
    }
    if (!dev_poly_[k]) {
        if (cudaMalloc(&dev_poly_[k], MT_N * sizeof(uint32_t)) != cudaSuccess) { cudaGetLastError(); err_ = "cudaMalloc(jump polynomial) failed"; return ICW_E_NOMEM; }
        if (cudaMemcpy(dev_poly_[k], host_poly_[k].w, MT_N * sizeof(uint32_t), cudaMemcpyHostToDevice) != cudaSuccess) { err_ = "copy of jump polynomial failed"; return ICW_E_CUDA; }
    }
    if (!attr_set_) {
        if (cudaFuncSetAttribute(mt_jump_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)JUMP_SMEM) != cudaSuccess) { err_ = "cudaFuncSetAttribute(mt_jump_kernel) failed"; return ICW_E_CUDA; }
        attr_set_ = true;
    }
    return ICW_OK;
}

int MtJump::poly_for(uint64_t blocks, const uint32_t **d_poly)
{
    if (blocks && !(blocks & (blocks - 1))) {                       // a power of two: the family itself
        int k = 0;
        while ((blocks >> k) != 1u) ++k;
        int rc = ensure_poly(k);
        if (rc) return rc;
        *d_poly = dev_poly_[k];
        return ICW_OK;
    }
    auto it = any_poly_.find(blocks);
    if (it == any_poly_.end()) {
        AnyPoly ap;
        ap.dev = nullptr;
        auto half = (blocks & 1u) ? any_poly_.end() : any_poly_.find(blocks >> 1);
        if (half != any_poly_.end()) {
            mt_poly_square(half->second.host, ap.host);
        } else {
            bool first = true;
            for (int k = 0; k < 64; ++k) {
                if (!((blocks >> k) & 1u)) continue;
                int rc = ensure_poly(k);
                if (rc) return rc;
                if (first) { ap.host = host_poly_[k]; first = false; }
                else { MtPoly t; mt_poly_mul(ap.host, host_poly_[k], t); ap.host = t; }
            }
        }
        if (cudaMalloc(&ap.dev, MT_N * sizeof(uint32_t)) != cudaSuccess) { cudaGetLastError(); err_ = "cudaMalloc(jump polynomial) failed"; return ICW_E_NOMEM; }
        if (cudaMemcpy(ap.dev, ap.host.w, MT_N * sizeof(uint32_t), cudaMemcpyHostToDevice) != cudaSuccess) { err_ = "copy of jump polynomial failed"; return ICW_E_CUDA; }
        it = any_poly_.emplace(blocks, ap).first;
    }
    *d_poly = it->second.dev;
    return ICW_OK;
}

// S_block (the state array u[624*block .. 624*block+623]) into d_state
int MtJump::state_at_block(uint32_t seed, uint64_t block, uint32_t *d_state, cudaStream_t stream, uint64_t *launches)
{
    for (int i = 0; i < NTAIL * 2; ++i)
        if (tails_[i].valid && tails_[i].stream == stream && tails_[i].seed == seed && tails_[i].block == block) {
            if (cudaMemcpyAsync(d_state, d_tail_ + (size_t)i * MT_N, MT_N * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream) != cudaSuccess) { err_ = "tail state copy failed"; return ICW_E_CUDA; }
            return ICW_OK;
        }
    uint32_t st[MT_N];
    mt_seed_state(seed, st);
    uint64_t rest = 0;
    if (block >= 1) { mt_regen_host(st); rest = block - 1; }        // never jump from S_0: u[0] carries 31 junk bits
    // short distances are cheaper sequentially on the host than one jump launch
    if (rest <= 64) { for (; rest; --rest) mt_regen_host(st); }
    if (cudaMemcpyAsync(d_state, st, sizeof st, cudaMemcpyHostToDevice, stream) != cudaSuccess) { err_ = "state upload failed"; return ICW_E_CUDA; }
    if (cudaStreamSynchronize(stream) != cudaSuccess) { err_ = "sync failed"; return ICW_E_CUDA; }   // st is on our stack
    if (!rest) return ICW_OK;
    if (!d_tmp_ && cudaMalloc(&d_tmp_, 2 * MT_N * sizeof(uint32_t)) != cudaSuccess) { cudaGetLastError(); err_ = "cudaMalloc failed"; return ICW_E_NOMEM; }
    // binary decomposition of the distance, one single-CTA jump per set bit, ping-pong in d_tmp_
    uint32_t *cur = d_state;
    int flip = 0;
    for (int k = 0; k < 64; ++k) {
        if (!((rest >> k) & 1u)) continue;
        int rc = ensure_poly(k);
        if (rc) return rc;
        uint32_t *dst = d_tmp_ + (size_t)flip * MT_N;
        cudaMemsetAsync(dst, 0, MT_N * sizeof(uint32_t), stream);
        mt_jump_kernel<<<dim3(1, 8), JUMP_THREADS, JUMP_SMEM, stream>>>(dst, cur, 0, 0, dev_poly_[k]);
        if (launches) ++*launches;
        cur = dst;
        flip ^= 1;
    }
    if (cur != d_state && cudaMemcpyAsync(d_state, cur, MT_N * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream) != cudaSuccess) { err_ = "state copy failed"; return ICW_E_CUDA; }
    if (cudaGetLastError() != cudaSuccess) { err_ = "mt_jump_kernel launch failed"; return ICW_E_CUDA; }
    return ICW_OK;
}

uint64_t mt_unit_blocks(uint64_t nb, int max_units)
{
    const uint64_t max_cta = (uint64_t)(max_units < 1 ? 1 : max_units);
    uint64_t bpu = (nb + max_cta - 1) / max_cta;
    int sh = 0;
    while (bpu >= 16) { bpu = (bpu + 1) >> 1; ++sh; }
    return (bpu ? bpu : 1) << sh;
}

// geometry + checkpoint 0 of one lane; the doubling tree is run by the caller (alone or for both lanes at once)
int MtJump::plan_prepare(int lane, uint32_t seed, uint64_t skip, int64_t n, int max_units, cudaStream_t stream,
                         uint64_t *launches, MtPlan &pl, uint64_t &bpu_out)
{
    pl = MtPlan();
    bpu_out = 0;
    if (n <= 0) return ICW_OK;
    const uint64_t b0 = skip / MT_N, b1 = (skip + (uint64_t)n - 1) / MT_N;
    const uint64_t nb = b1 - b0 + 1;
    // blocks per unit: the smallest m * 2^k (m < 16) that covers the range with max_units units -- a full
    // wave of CTAs to within a few per cent, and a small set of jump distances to keep polynomials for
    const uint64_t bpu = mt_unit_blocks(nb, max_units);
    const int n_cta = (int)((nb + bpu - 1) / bpu);
    uint32_t *&ck = d_ckpt_[lane];
    if ((size_t)n_cta > ckpt_cap_[lane]) {
        if (ck) cudaFree(ck);
        ck = nullptr; ckpt_cap_[lane] = 0;
        size_t want = (size_t)std::max(n_cta, 64);
        if (cudaMalloc(&ck, want * MT_N * sizeof(uint32_t)) != cudaSuccess) { cudaGetLastError(); err_ = "cudaMalloc(checkpoints) failed"; return ICW_E_NOMEM; }
        ckpt_cap_[lane] = want;
    }
    int rc = state_at_block(seed, b0, ck, stream, launches);
    if (rc) return rc;
    if (!d_tail_ && cudaMalloc(&d_tail_, (size_t)NTAIL * 2 * MT_N * sizeof(uint32_t)) != cudaSuccess) { cudaGetLastError(); err_ = "cudaMalloc(tail states) failed"; return ICW_E_NOMEM; }
    const int slot = tail_next_;
    tail_next_ = (tail_next_ + 1) % NTAIL;
    // the consumer of the plan fills the slot in stream order; lookups on the same stream come after it
    tails_[2 * slot] = { seed, b1, true, stream };          // state that regenerates into block b1
    tails_[2 * slot + 1] = { seed, b1 + 1, true, stream };  // state after block b1
    pl.ckpt = ck;
    pl.tail = d_tail_ + (size_t)slot * 2 * MT_N;
    pl.n_units = n_cta;
    pl.blocks_per_unit = (int)bpu;
    pl.first_word = (int64_t)(b0 * MT_N);
    pl.want_lo = (int64_t)skip;
    pl.want_hi = (int64_t)(skip + (uint64_t)n);
    pl.tail_block = (int64_t)(b1 - b0);
    bpu_out = bpu;
    return ICW_OK;
}

// doubling: checkpoints [2^j, 2^(j+1)) come from [0, 2^j) by a jump of bpu * 2^j blocks; ck_b != NULL runs a second
// checkpoint array of the same geometry in the same launches (blockIdx.z)
int MtJump::plan_tree(uint32_t *ck_a, uint32_t *ck_b, int n_cta, uint64_t bpu, int sm_count, cudaStream_t stream, uint64_t *launches)
{
    if (n_cta > 1 && !attr_set_) { int rc = ensure_poly(0); if (rc) return rc; }
    const int nz = ck_b ? 2 : 1;
    for (int j = 0; (1 << j) < n_cta; ++j) {
        const uint32_t *d_poly = nullptr;
        int rc = poly_for(bpu << j, &d_poly);
        if (rc) return rc;
        const int first = 1 << j;
        const int count = std::min(n_cta, 2 << j) - first;
        const int sl = jump_slices(count * nz, sm_count);
        if (sl > 1) {
            cudaMemsetAsync(ck_a + (size_t)first * MT_N, 0, (size_t)count * MT_N * sizeof(uint32_t), stream);
            if (ck_b) cudaMemsetAsync(ck_b + (size_t)first * MT_N, 0, (size_t)count * MT_N * sizeof(uint32_t), stream);
        }
        mt_jump_kernel<<<dim3(count, sl, nz), JUMP_THREADS, JUMP_SMEM, stream>>>(ck_a, nullptr, first, first, d_poly, ck_b);
        if (launches) ++*launches;
    }
    if (cudaGetLastError() != cudaSuccess) { err_ = "mt_jump_kernel launch failed"; return ICW_E_CUDA; }
    return ICW_OK;
}

int MtJump::plan(int lane, uint32_t seed, uint64_t skip, int64_t n, int max_units, int sm_count, cudaStream_t stream,
                 uint64_t *launches, MtPlan &pl)
{
    uint64_t bpu = 0;
    int rc = plan_prepare(lane, seed, skip, n, max_units, stream, launches, pl, bpu);
    if (rc || n <= 0) return rc;
    return plan_tree(d_ckpt_[lane], nullptr, pl.n_units, bpu, sm_count, stream, launches);
}

int MtJump::plan_pair(const uint32_t seed[2], uint64_t skip, int64_t n, int max_units, int sm_count, cudaStream_t stream,
                      uint64_t *launches, MtPlan pl[2])
{
    uint64_t bpu[2] = { 0, 0 };
    for (int c = 0; c < 2; ++c) {
        int rc = plan_prepare(c, seed[c], skip, n, max_units, stream, launches, pl[c], bpu[c]);
        if (rc) return rc;
    }
    if (n <= 0) return ICW_OK;
    return plan_tree(d_ckpt_[0], d_ckpt_[1], pl[0].n_units, bpu[0], sm_count, stream, launches);
}

int MtJump::generate(uint32_t seed, uint64_t skip, int64_t n, uint32_t *d_out, int sm_count,
                     cudaStream_t stream, uint64_t *launches)
{
    if (n <= 0) return ICW_OK;
    // one CTA of 256 threads per checkpoint (mt_words_kernel), eight of them resident per SM
    MtPlan pl;
    int rc = plan(0, seed, skip, n, sm_count * 8, sm_count, stream, launches, pl);
    if (rc) return rc;
    cudaError_t e = launch_mt_words(pl.ckpt, pl.n_units, pl.blocks_per_unit, pl.first_word, pl.want_lo, pl.want_hi, d_out,
                                    pl.tail_block, pl.tail, stream);
    if (launches) ++*launches;
    if (e != cudaSuccess) { err_ = std::string("mt kernels: ") + cudaGetErrorString(e); return ICW_E_CUDA; }
    return ICW_OK;
}

}  // namespace icw

// =============================================================================================
// host-only hooks for CPU tests of the polynomial machinery (no GPU touched)
// =============================================================================================
extern "C" int icw_mt_host_charpoly(int *n_terms, int *degree)
{
    const std::vector<int> &t = icw::mt_charpoly_terms();
    if (n_terms) *n_terms = (int)t.size();
    if (degree) *degree = t.empty() ? -1 : t.back();
    return t.empty() ? ICW_E_ARG : ICW_OK;
}

// S_blocks by sequential regeneration
extern "C" void icw_mt_host_seq_state(uint32_t seed, uint64_t blocks, uint32_t *out624)
{
    icw::mt_seed_state(seed, out624);
    for (uint64_t b = 0; b < blocks; ++b) icw::mt_regen_host(out624);
}

// S_blocks through x^(624*(blocks-1)) mod phi applied to S_1 (the jump path, all on the host)
extern "C" int icw_mt_host_jump_state(uint32_t seed, uint64_t blocks, uint32_t *out624)
{
    if (icw::mt_charpoly_terms().empty()) return ICW_E_ARG;
    uint32_t st[icw::MT_N];
    icw::mt_seed_state(seed, st);
    if (blocks == 0) { memcpy(out624, st, sizeof st); return ICW_OK; }
    icw::mt_regen_host(st);
    icw::MtPoly g;
    icw::mt_poly_xpow((blocks - 1) * (uint64_t)icw::MT_N, g);
    icw::mt_apply_host(g, st, out624);
    return ICW_OK;
}

// same, but composed from the x^(624*2^k) family exactly as the GPU path composes it
extern "C" int icw_mt_host_jump_state_family(uint32_t seed, uint64_t blocks, uint32_t *out624)
{
    if (icw::mt_charpoly_terms().empty()) return ICW_E_ARG;
    uint32_t st[icw::MT_N], nx[icw::MT_N];
    icw::mt_seed_state(seed, st);
    if (blocks == 0) { memcpy(out624, st, sizeof st); return ICW_OK; }
    icw::mt_regen_host(st);
    uint64_t rest = blocks - 1;
    icw::MtPoly f, sq;
    icw::mt_poly_xpow(icw::MT_N, f);
    for (int k = 0; k < 64 && (rest >> k); ++k) {
        if ((rest >> k) & 1u) { icw::mt_apply_host(f, st, nx); memcpy(st, nx, sizeof st); }
        icw::mt_poly_square(f, sq);
        f = sq;
    }
    memcpy(out624, st, sizeof st);
    return ICW_OK;
}

// same, as one polynomial: the product of the family over the set bits of the distance (MtJump::poly_for)
extern "C" int icw_mt_host_jump_state_product(uint32_t seed, uint64_t blocks, uint32_t *out624)
{
    if (icw::mt_charpoly_terms().empty()) return ICW_E_ARG;
    uint32_t st[icw::MT_N];
    icw::mt_seed_state(seed, st);
    if (blocks == 0) { memcpy(out624, st, sizeof st); return ICW_OK; }
    icw::mt_regen_host(st);
    const uint64_t rest = blocks - 1;
    if (!rest) { memcpy(out624, st, sizeof st); return ICW_OK; }
    icw::MtPoly f, sq, g, t;
    icw::mt_poly_xpow(icw::MT_N, f);
    bool first = true;
    for (int k = 0; k < 64 && (rest >> k); ++k) {
        if ((rest >> k) & 1u) {
            if (first) { g = f; first = false; }
            else { icw::mt_poly_mul(g, f, t); g = t; }
        }
        icw::mt_poly_square(f, sq);
        f = sq;
    }
    icw::mt_apply_host(g, st, out624);
    return ICW_OK;
}

// unit length (blocks) the checkpoint planner picks for nb blocks and at most max_units units
extern "C" uint64_t icw_mt_host_unit_blocks(uint64_t nb, int max_units) { return icw::mt_unit_blocks(nb, max_units); }
