// icw_mt.h -- MT19937 jump-ahead service: lets any CTA start the reference's dither stream at an
// arbitrary 624-word block without generating what comes before it.
//
// The reference draws its dither from mtrnd_gen_ui32 (reference src/mersene_twister/mt_jrnd.c:99-134)
// after mtrnd_init_seed (:28-47); draw number j is the tempered word u[624 + j] of the linear
// recurring sequence u[k+624] = u[k+397] ^ twist(u[k], u[k+1]).  With phi(x) the characteristic
// polynomial (degree 19937) and g(x) = x^J mod phi(x),  u[p + J] = XOR_{i : g_i = 1} u[p + i]  for
// every p >= 1 -- a GF(2) convolution of g with 33 blocks of the sequence, embarrassingly parallel
// over the 624 words of the target state.  The host computes phi once (Berlekamp-Massey on the
// generator's own output), the polynomials x^(624*2^k) mod phi lazily, and the GPU applies them.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>
#include <cuda_runtime.h>

namespace icw {

constexpr int MT_N = 624;
constexpr int MT_DEG = 19937;
constexpr int MT_PW = 312;              // 64-bit words of a polynomial of degree < 19968

struct MtPoly { uint64_t w[MT_PW]; };   // bit i = coefficient of x^i

// ---- host-side GF(2) machinery (also exported for CPU tests, see icw_mt_host_* in icw_mt.cu) ----
void mt_seed_state(uint32_t seed, uint32_t st[MT_N]);               // mtrnd_init_seed
void mt_regen_host(uint32_t st[MT_N]);                              // one block regeneration, in place
const std::vector<int> &mt_charpoly_terms();                         // exponents of phi, ascending
void mt_poly_one(MtPoly &p);                                         // p = 1
void mt_poly_mulx_pow(MtPoly &p, uint64_t e);                        // p = p * x^e mod phi (e small) ...
void mt_poly_square(const MtPoly &a, MtPoly &out);                   // out = a^2 mod phi
void mt_poly_mul(const MtPoly &a, const MtPoly &b, MtPoly &out);     // out = a*b mod phi
void mt_poly_xpow(uint64_t e, MtPoly &out);                          // out = x^e mod phi
// target[m] = XOR_i g_i * seq[i+m], seq = the block sequence continued from base (host, slow)
void mt_apply_host(const MtPoly &g, const uint32_t base[MT_N], uint32_t out[MT_N]);

// where a kernel that regenerates the stream itself starts: one checkpoint per unit of
// `blocks_per_unit` consecutive 624-word blocks; unit 0's first block starts at stream word `first_word`
struct MtPlan {
    const uint32_t *ckpt = nullptr;     // [n_units][624]: the state array that PRECEDES each unit's first block
    uint32_t *tail = nullptr;           // [2][624]: to be filled with the state before and after block `tail_block`
    int n_units = 0, blocks_per_unit = 0;
    int64_t first_word = 0, want_lo = 0, want_hi = 0;   // wanted stream words [want_lo, want_hi)
    int64_t tail_block = 0;             // relative index of the block holding the last wanted word
};

// blocks per jump-ahead unit for a range of nb blocks and at most max_units units: the smallest m * 2^k, m < 16
uint64_t mt_unit_blocks(uint64_t nb, int max_units);

class MtJump {
public:
    // checkpoints for words [skip, skip + n) of the stream of `seed`, at most `max_units` of them.
    // `lane` (0/1) selects one of two checkpoint buffers so that two generators can be live at once.
    // The kernel that consumes the plan MUST write pl.tail (see mt_words_kernel).
    int plan(int lane, uint32_t seed, uint64_t skip, int64_t n, int max_units, int sm_count, cudaStream_t stream,
             uint64_t *launches, MtPlan &pl);
    // the same for two generators standing at the same draw (the two channels' dither): one doubling tree
    // of launches for both (lanes 0 and 1)
    int plan_pair(const uint32_t seed[2], uint64_t skip, int64_t n, int max_units, int sm_count, cudaStream_t stream,
                  uint64_t *launches, MtPlan pl[2]);
    // tempered words [skip, skip + n) of the stream of `seed` -> d_out (device), on `stream`
    int generate(uint32_t seed, uint64_t skip, int64_t n, uint32_t *d_out, int sm_count,
                 cudaStream_t stream, uint64_t *launches);
    const char *error() const { return err_.c_str(); }
    void release();

private:
    int plan_prepare(int lane, uint32_t seed, uint64_t skip, int64_t n, int max_units, cudaStream_t stream,
                     uint64_t *launches, MtPlan &pl, uint64_t &bpu_out);
    int plan_tree(uint32_t *ck_a, uint32_t *ck_b, int n_cta, uint64_t bpu, int sm_count, cudaStream_t stream, uint64_t *launches);
    int ensure_poly(int k);             // device copy of x^(624 * 2^k) mod phi
    // device copy of x^(624 * blocks) mod phi for any distance: the product of the 2^k family over the
    // set bits of `blocks`, or the square of the polynomial for blocks / 2 when that one is cached
    int poly_for(uint64_t blocks, const uint32_t **d_poly);
    struct AnyPoly { MtPoly host; uint32_t *dev; };
    std::map<uint64_t, AnyPoly> any_poly_;
    int state_at_block(uint32_t seed, uint64_t block, uint32_t *d_state, cudaStream_t stream, uint64_t *launches);
    std::vector<MtPoly> host_poly_;     // [k]
    std::vector<uint32_t *> dev_poly_;  // [k] -> 624 x uint32 on the device
    uint32_t *d_ckpt_[2] = { nullptr, nullptr };    // checkpoint states [lane][count][624]
    size_t ckpt_cap_[2] = { 0, 0 };
    uint32_t *d_tmp_ = nullptr;         // ping-pong for sequential jumps
    // generator states left behind by earlier calls: a stream that continues where the last call
    // stopped finds its start state here instead of jumping there from the seed
    struct Tail { uint32_t seed; uint64_t block; bool valid; cudaStream_t stream; };   // stream-ordered: only valid on the stream that wrote it
    static constexpr int NTAIL = 8;     // buffers of 2 states each
    uint32_t *d_tail_ = nullptr;        // [NTAIL][2][624]
    Tail tails_[NTAIL * 2] = {};
    int tail_next_ = 0;
    bool attr_set_ = false;
    std::string err_;
};

}  // namespace icw
