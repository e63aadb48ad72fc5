// icw_crc.h -- CRC-32 of device buffers (icw_crc.cu)
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace icw {

// register * x^(8 n) mod P (reflected form): the effect of n zero bytes on a CRC register
uint32_t crc32_shift(uint32_t reg, uint64_t n_bytes);
// crc(A || B) from the finished crc(A), crc(B) and |B|
uint32_t crc32_combine(uint32_t crc_a, uint32_t crc_b, uint64_t len_b);
// CRC register started from ZERO over n bytes -> *d_out; d_partial holds >= n / 32768 + 3 words
cudaError_t launch_crc32_raw(const uint8_t *d_data, size_t n, uint32_t *d_partial, uint32_t *d_out, cudaStream_t s, int *launches);

}  // namespace icw
