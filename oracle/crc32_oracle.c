/*
 * crc32_oracle.c -- TEST INFRASTRUCTURE ONLY: bit-at-a-time CRC-32 and fqzcomp5's block framing,
 * restated in plain C to check fqzcomp5_b200/csrc/crc32.cu.  Nothing in the product path links or
 * calls this file.
 *
 * The reference takes its CRC from a third-party dependency that is not under /root/reference:
 * zlib's crc32() (fqzcomp5.c:2268-2269 and :2310-2311; `-lz` in the reference Makefile; the system
 * zlib here is what `python -c "import zlib; print(zlib.ZLIB_VERSION)"` reports).  Its published
 * algorithm is CRC-32/ISO-HDLC: reflected polynomial 0xEDB88320, register preset to ~crc, result
 * complemented; check value crc32("123456789") = 0xCBF43926.
 *
 * Parity pinned: tests/test_crc32_oracle.py compares orc_crc32 with zlib itself (Python's zlib
 * module binds the same library) and the check value, and checks the framing against a block
 * written by the unmodified reference tool (oracle/_ref/fqzcomp5_ref) and against the known answer
 * recorded in tests/golden/block_frame.json.
 */
#include <stdint.h>
#include <string.h>

uint32_t orc_crc32(uint32_t crc, const uint8_t *buf, uint64_t n) {
    uint32_t r = ~crc;
    for (uint64_t i = 0; i < n; i++) {
        r ^= buf[i];
        for (int k = 0; k < 8; k++) r = (r >> 1) ^ ((r & 1) ? 0xEDB88320u : 0u);
    }
    return ~r;
}

/* encode_block's framing (fqzcomp5.c:2147-2280): [u32 size = total - 4][u32 num_records][u32 crc]
 * [pieces...], the CRC over everything after the CRC field (:2266-2274).  Returns the block length. */
uint32_t orc_frame_block(uint32_t num_records, int n_pieces, const uint8_t *const *piece, const uint32_t *len,
                         uint8_t *out) {
    uint32_t o = 12;
    for (int i = 0; i < n_pieces; i++) { memcpy(out + o, piece[i], len[i]); o += len[i]; }
    uint32_t size = o - 4, crc = orc_crc32(0, out + 12, o - 12);
    memcpy(out, &size, 4);
    memcpy(out + 4, &num_records, 4);
    memcpy(out + 8, &crc, 4);
    return o;
}
