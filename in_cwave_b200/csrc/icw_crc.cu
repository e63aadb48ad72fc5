// icw_crc.cu -- CRC-32 of a device buffer: the CWAVE sample-data check.
//
// The reference checks a CWAVE file's data against the header's n_CRC32 with a bytewise table CRC
// (reference src/crc32.c:55-108, driven by src/gui_cwave.c:82-130).  Its "first four bytes inverted"
// start and final inversion make it CRC-32/ISO-HDLC (reflected polynomial 0xEDB88320, init and
// xor-out 0xFFFFFFFF) -- tests pin that against the compiled reference.
//
// A CRC register is linear over GF(2): the register after a message, started from zero, is the
// message polynomial times x^32 mod P, so  raw(A || B) = raw(A) * x^(8|B|)  xor  raw(B),  and a
// non-zero start value I adds  I * x^(8n).  That turns the serial byte loop into:
//   1. every thread: raw CRC of its own 128 contiguous bytes (slicing-by-4 tables in shared memory,
//      bytes staged through shared memory by coalesced 16-byte loads, rows padded against bank
//      conflicts);
//   2. a tree over the CTA's 256 threads: left half times x^(8 * 128 * 2^k) xor right half, the
//      multiplication by that constant being a 32 x 32 bit matrix (one column per register bit);
//   3. one CTA folds the per-tile values the same way (Horner over each thread's run, then a tree).
// Tiles start on 16-byte boundaries of the pointer.  What precedes the first boundary and what
// follows the last full tile are short tiles, right-aligned in a tile of virtual zeros (zero bytes in
// FRONT of a message leave a zero register alone); the tail joins through its own shift matrix.
// HBM traffic = the bytes, once.
#include <cstdint>
#include <cstring>
#include <vector>

#include "icw_crc.h"

namespace icw {

constexpr uint32_t CRC_POLY = 0xEDB88320u;
constexpr int CRC_THREADS = 256;
constexpr int CRC_SUB = 128;                        // bytes per thread
constexpr int CRC_TILE = CRC_THREADS * CRC_SUB;     // 32 KB per CTA
constexpr int CRC_ROW = CRC_SUB + 16;               // padded row in shared memory: 16-byte accesses at this stride hit 8 x 4 distinct banks
constexpr int CRC_LEVELS = 8;                       // log2(CRC_THREADS)

// ---- host: GF(2) arithmetic on the reflected register ------------------------------------------------
// multiply the register by x^8 (one zero byte): the table-CRC step with a zero data byte
static uint32_t host_table[256];
static void host_init_table()
{
    if (host_table[1]) return;
    for (uint32_t i = 0; i < 256; ++i) {
        uint32_t c = i;
        for (int j = 0; j < 8; ++j) c = (c >> 1) ^ ((c & 1u) ? CRC_POLY : 0u);
        host_table[i] = c;
    }
}
struct Mat { uint32_t col[32]; };                   // col[b] = image of register bit b
static uint32_t mat_apply(const Mat &m, uint32_t v)
{
    uint32_t r = 0;
    for (int b = 0; v; ++b, v >>= 1) if (v & 1u) r ^= m.col[b];
    return r;
}
static Mat mat_mul(const Mat &a, const Mat &b)      // a after b
{
    Mat r;
    for (int i = 0; i < 32; ++i) r.col[i] = mat_apply(a, b.col[i]);
    return r;
}
static Mat mat_shift_bytes(uint64_t n_bytes)        // register -> register * x^(8 n)
{
    host_init_table();
    Mat one, r;
    for (int b = 0; b < 32; ++b) { uint32_t v = 1u << b; one.col[b] = host_table[v & 0xFF] ^ (v >> 8); r.col[b] = 1u << b; }
    Mat p = one;
    for (; n_bytes; n_bytes >>= 1) { if (n_bytes & 1) r = mat_mul(p, r); p = mat_mul(p, p); }
    return r;
}

uint32_t crc32_shift(uint32_t reg, uint64_t n_bytes) { return mat_apply(mat_shift_bytes(n_bytes), reg); }

// zlib-style: crc(A || B) from crc(A), crc(B), |B| (finalised values)
uint32_t crc32_combine(uint32_t crc_a, uint32_t crc_b, uint64_t len_b)
{
    return crc32_shift(crc_a, len_b) ^ crc_b;       // the init / xor-out terms cancel pairwise
}

// ---- device ------------------------------------------------------------------------------------------
struct CrcMats { uint32_t col[CRC_LEVELS + 1][32]; };    // [k] = times x^(8 * unit * 2^k)

__device__ __forceinline__ uint32_t dev_mat_apply(const uint32_t *col, uint32_t v)
{
    uint32_t r = 0;
#pragma unroll
    for (int b = 0; b < 32; ++b) r ^= ((v >> b) & 1u) ? col[b] : 0u;
    return r;
}

// tile t = bytes [first + t * CRC_TILE, ...) of `data`; a negative first byte = virtual leading zeros
__global__ void __launch_bounds__(CRC_THREADS)
crc_tile_kernel(const uint8_t *__restrict__ data, int64_t first, int64_t n_bytes, const __grid_constant__ CrcMats mats,
                uint32_t *__restrict__ partial)
{
    __shared__ uint32_t tab[4][256];
    __shared__ __align__(16) uint8_t rows[CRC_THREADS * CRC_ROW];
    __shared__ uint32_t red[CRC_THREADS];
    __shared__ uint32_t mcol[CRC_LEVELS][32];
    const int t = threadIdx.x;
    // slicing-by-4 tables: tab[0] = the byte table, tab[k][i] = tab[k-1][i] advanced by one zero byte
    {
        uint32_t c = (uint32_t)t;
#pragma unroll
        for (int j = 0; j < 8; ++j) c = (c >> 1) ^ ((c & 1u) ? CRC_POLY : 0u);
        tab[0][t] = c;
    }
    if (t < CRC_LEVELS * 32) mcol[t >> 5][t & 31] = mats.col[t >> 5][t & 31];
    __syncthreads();
    {
        uint32_t c = tab[0][t];
#pragma unroll
        for (int k = 1; k < 4; ++k) { c = tab[0][c & 0xFF] ^ (c >> 8); tab[k][t] = c; }
    }
    // stage the tile: 2048 16-byte vectors, 8 per thread, coalesced; bytes outside [0, n_bytes) are zero
    const int64_t base = first + (int64_t)blockIdx.x * CRC_TILE;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(data + (base < 0 ? 0 : base)) & 15u) == 0 && base >= 0 && base + CRC_TILE <= n_bytes;
#pragma unroll
    for (int k = 0; k < CRC_TILE / 16 / CRC_THREADS; ++k) {
        const int v = k * CRC_THREADS + t;                              // vector index inside the tile
        const int row = v / (CRC_SUB / 16), cv = v % (CRC_SUB / 16);
        uint4 w;
        if (vec_ok) {
            w = *reinterpret_cast<const uint4 *>(data + base + (int64_t)v * 16);
        } else {
            uint8_t b[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int64_t p = base + (int64_t)v * 16 + j;
                b[j] = (p >= 0 && p < n_bytes) ? data[p] : (uint8_t)0;
            }
            memcpy(&w, b, 16);
        }
        *reinterpret_cast<uint4 *>(rows + row * CRC_ROW + cv * 16) = w;
    }
    __syncthreads();
    // raw CRC of this thread's 128 bytes
    uint32_t crc = 0;
    const uint4 *mine = reinterpret_cast<const uint4 *>(rows + t * CRC_ROW);   // 16-byte reads at a 144-byte lane stride: conflict-free
#pragma unroll
    for (int k = 0; k < CRC_SUB / 16; ++k) {
        const uint4 w = mine[k];
        const uint32_t ws[4] = { w.x, w.y, w.z, w.w };
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t c = crc ^ ws[j];
            crc = tab[3][c & 0xFF] ^ tab[2][(c >> 8) & 0xFF] ^ tab[1][(c >> 16) & 0xFF] ^ tab[0][c >> 24];
        }
    }
    // tree: at level k the left partner covers 128 * 2^k bytes that precede the right partner's
    red[t] = crc;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < CRC_LEVELS; ++k) {
        const int stride = 1 << k;
        uint32_t v = 0;
        const bool act = (t & (2 * stride - 1)) == 0;
        if (act) v = dev_mat_apply(mcol[k], red[t]) ^ red[t + stride];
        __syncthreads();
        if (act) red[t] = v;
        __syncthreads();
    }
    if (t == 0) partial[blockIdx.x] = red[0];
}

// fold n tile values (each a raw CRC of CRC_TILE bytes, in stream order) into one; mats[k] = x^(8 * CRC_TILE * run * 2^k)
// with run = tiles per thread, mats[CRC_LEVELS] = x^(8 * CRC_TILE)
// then, if has_tail, the value is advanced over the tail's bytes (tail_mat) and joined with partial[n]
__global__ void __launch_bounds__(CRC_THREADS)
crc_fold_kernel(const uint32_t *__restrict__ partial, int64_t n, int run, const __grid_constant__ CrcMats mats,
                const __grid_constant__ CrcMats tail_mat, int has_tail, uint32_t *__restrict__ out)
{
    __shared__ uint32_t red[CRC_THREADS];
    __shared__ uint32_t mcol[CRC_LEVELS + 1][32];
    const int t = threadIdx.x;
    for (int i = t; i < (CRC_LEVELS + 1) * 32; i += CRC_THREADS) mcol[i >> 5][i & 31] = mats.col[i >> 5][i & 31];
    __syncthreads();
    // the runs are right-aligned: thread 255 owns the last `run` tiles, missing tiles in front are zero
    const int64_t lo = n - (int64_t)(CRC_THREADS - t) * run;
    uint32_t acc = 0;
    for (int j = 0; j < run; ++j) {
        const int64_t idx = lo + j;
        acc = dev_mat_apply(mcol[CRC_LEVELS], acc) ^ (idx >= 0 ? partial[idx] : 0u);
    }
    red[t] = acc;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < CRC_LEVELS; ++k) {
        const int stride = 1 << k;
        uint32_t v = 0;
        const bool act = (t & (2 * stride - 1)) == 0;
        if (act) v = dev_mat_apply(mcol[k], red[t]) ^ red[t + stride];
        __syncthreads();
        if (act) red[t] = v;
        __syncthreads();
    }
    if (t == 0) *out = has_tail ? (dev_mat_apply(tail_mat.col[0], red[0]) ^ partial[n]) : red[0];
}

static void fill_mats(CrcMats &m, uint64_t unit_bytes, bool with_unit)
{
    Mat s = mat_shift_bytes(unit_bytes);
    for (int k = 0; k < CRC_LEVELS; ++k) {
        memcpy(m.col[k], s.col, sizeof s.col);
        s = mat_mul(s, s);
    }
    if (with_unit) { Mat u = mat_shift_bytes(CRC_TILE); memcpy(m.col[CRC_LEVELS], u.col, sizeof u.col); }
    else memset(m.col[CRC_LEVELS], 0, sizeof m.col[CRC_LEVELS]);
}

// raw register (start value 0) over n bytes -> *d_out (device).  d_partial: >= n / CRC_TILE + 3 words.
cudaError_t launch_crc32_raw(const uint8_t *d_data, size_t n, uint32_t *d_partial, uint32_t *d_out, cudaStream_t s, int *launches)
{
    if (n == 0) return cudaMemsetAsync(d_out, 0, sizeof(uint32_t), s);
    const size_t mis = (size_t)(reinterpret_cast<uintptr_t>(d_data) & 15u);
    size_t head = (16 - mis) & 15;
    if (head > n) head = n;
    const int64_t m = (int64_t)((n - head) / CRC_TILE);                 // aligned full tiles
    const size_t n_main = head + (size_t)m * CRC_TILE;
    const size_t n_tail = n - n_main;
    const int64_t main_tiles = m + (head ? 1 : 0);
    CrcMats tm;
    fill_mats(tm, CRC_SUB, false);
    if (main_tiles)
        crc_tile_kernel<<<(unsigned)main_tiles, CRC_THREADS, 0, s>>>(d_data, head ? (int64_t)head - CRC_TILE : 0, (int64_t)n_main, tm, d_partial);
    if (n_tail)
        crc_tile_kernel<<<1, CRC_THREADS, 0, s>>>(d_data + n_main, (int64_t)n_tail - CRC_TILE, (int64_t)n_tail, tm, d_partial + main_tiles);
    const int run = (int)((main_tiles + CRC_THREADS - 1) / CRC_THREADS);
    CrcMats fm, tail;
    fill_mats(fm, (uint64_t)CRC_TILE * (uint64_t)(run > 0 ? run : 1), true);
    memset(&tail, 0, sizeof tail);
    if (n_tail) { Mat t = mat_shift_bytes(n_tail); memcpy(tail.col[0], t.col, sizeof t.col); }
    crc_fold_kernel<<<1, CRC_THREADS, 0, s>>>(d_partial, main_tiles, run, fm, tail, n_tail ? 1 : 0, d_out);
    if (launches) *launches += 1 + (main_tiles ? 1 : 0) + (n_tail ? 1 : 0);
    return cudaGetLastError();
}

}  // namespace icw
