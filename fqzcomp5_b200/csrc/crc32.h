// crc32.h -- host-visible launcher of the CRC-32 kernels (crc32.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace b200 {

size_t crc32_scratch_bytes(uint64_t n);
// zlib's crc32(crc_in, buf, n) of d_buf[0, n) (any alignment) -> *d_out.  With d_patch, also
// writes patch_size_val at d_patch[0..3] and the CRC at d_patch[8..11] (little endian): the
// size and CRC fields of an fqzcomp5 block (fqzcomp5.c:2266-2274).
cudaError_t crc32_launch(const uint8_t *d_buf, uint64_t n, uint32_t crc_in, uint32_t *d_out, uint8_t *d_patch,
                         uint32_t patch_size_val, uint8_t *d_scratch, cudaStream_t st, int *launches);

}  // namespace b200
