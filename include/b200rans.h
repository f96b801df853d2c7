/*
 * b200rans.h -- C ABI of libb200rans.so: the htscodecs rANS Nx16 codec
 * (order-0/order-1, 4 or 32 interleaved lanes, PACK / RLE / NOSZ / CAT / STRIPE)
 * computed by hand-written sm_100a CUDA kernels.
 *
 * Part 1 re-declares, unchanged, the interface fqzcomp5 binds today
 * (/root/reference/htscodecs/rANS_static4x16.h:41-64); stock fqzcomp5.o and
 * tokenise_name3.o link against this library instead of rANS_static4x16pr.o +
 * rANS_static32x16pr*.o (call sites: fqzcomp5.c:1422,1528,1550,1594,1639,1650,
 * 1996,2008,2016,2433,2447,2490; tokenise_name3.c:1243,1261).
 *
 * Part 2 is the batched extension those callers need to keep a GPU busy
 * (SURVEY H1/H9): one call carries many independent streams.
 *
 * There is no CPU implementation behind any entry point: without a usable
 * CUDA device every call fails (NULL / negative return) and says so on stderr.
 */
#ifndef B200RANS_H
#define B200RANS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------
 * Part 1: drop-in symbols.  Semantics, flag byte, ownership (out == NULL =>
 * the library malloc()s and the caller free()s) and error convention (NULL,
 * encoder also sets *out_size = 0) are the reference's.
 * ---------------------------------------------------------------------- */

/* rANS_static4x16.h:41  (rANS_static4x16pr.c:93-106) */
unsigned int rans_compress_bound_4x16(unsigned int size, int order);

/* rANS_static4x16.h:42-44  (rANS_static4x16pr.c:1224-1600) */
unsigned char *rans_compress_to_4x16(unsigned char *in, unsigned int in_size,
                                     unsigned char *out, unsigned int *out_size,
                                     int order);
/* rANS_static4x16.h:45-46  (rANS_static4x16pr.c:1602-1605) */
unsigned char *rans_compress_4x16(unsigned char *in, unsigned int in_size,
                                  unsigned int *out_size, int order);
/* rANS_static4x16.h:47-48  (rANS_static4x16pr.c:1607-1894) */
unsigned char *rans_uncompress_to_4x16(unsigned char *in, unsigned int in_size,
                                       unsigned char *out, unsigned int *out_size);
/* rANS_static4x16.h:49-50  (rANS_static4x16pr.c:1896-1899) */
unsigned char *rans_uncompress_4x16(unsigned char *in, unsigned int in_size,
                                    unsigned int *out_size);
/* rANS_static4x16.h:64  (rANS_static4x16pr.c:1212-1217): test hook selecting
 * x86 ISA variants; accepted and ignored (there is one implementation). */
void rans_set_cpu(int opts);

/* "order" bits, rANS_static4x16.h:66-103 */
#define RANS_ORDER_PACK       0x80
#define RANS_ORDER_RLE        0x40
#define RANS_ORDER_CAT        0x20
#define RANS_ORDER_NOSZ       0x10
#define RANS_ORDER_STRIPE     0x08
#define RANS_ORDER_X32        0x04
#define RANS_ORDER_STRIPE_NO0 (1 << 16)
#define RANS_ORDER_SIMD_AUTO  (1 << 17)

/* ------------------------------------------------------------------------
 * Part 2: batched extension (not in the reference).
 *
 * A batch is n independent calls.  Each produces / consumes exactly the byte
 * stream the corresponding single call would.  Calls that the reference would
 * fail (NULL) are reported per stream, the rest of the batch still completes.
 * All functions return 0 on success and a negative b200rans_status otherwise.
 * ---------------------------------------------------------------------- */

typedef enum {
    B200RANS_OK = 0,
    B200RANS_ENODEV = -1,   /* no CUDA device / driver; there is no CPU fallback */
    B200RANS_ECUDA = -2,    /* a CUDA call failed; text on stderr */
    B200RANS_ENOMEM = -3,
    B200RANS_EINVAL = -4,
    B200RANS_ESPACE = -5    /* output arena too small */
} b200rans_status;

/* Select the device used by the calling thread's context (default: device 0,
 * or $B200RANS_DEVICE).  Each host thread owns a private context (stream,
 * device + pinned scratch arenas), mirroring the per-thread scratch of
 * htscodecs/utils.c:119-208, so hts_tpool workers may call concurrently. */
int b200rans_set_device(int device);
int b200rans_device_count(void);

/* Pinned host memory helpers: buffers allocated here move at full PCIe rate. */
void *b200rans_host_alloc(size_t bytes);
void  b200rans_host_free(void *p);

/* Host buffers in, one host arena out.
 *   in[k], in_size[k], order[k]   as rans_compress_to_4x16
 *   out, out_cap                  arena receiving the n streams back to back
 *   out_off[k], out_size[k]       where stream k landed; out_size[k]==0 => that
 *                                 call failed as the reference's would
 * Capacity semantics of each call are those of out == NULL in the reference
 * (the library provides rans_compress_bound_4x16 bytes). */
int b200rans_compress_batch(int n,
                            const unsigned char *const *in, const unsigned int *in_size,
                            const int *order,
                            unsigned char *out, size_t out_cap,
                            size_t *out_off, unsigned int *out_size);

/* Batched method trial: the rANS members of fqzcomp5's compress_with_methods
 * (fqzcomp5.c:1979-2119: RANS0..RANS193 = orders {0,1,64,65,128,129,192,193} at
 * :2005-2011, RANSXN1 = (fixed_len<<8)+9 at :2013-2022) and of the learner that
 * consumes their sizes (metrics_update, fqzcomp5.c:1950-1958).  Every input is
 * staged once and encoded under each of methods[0..n_methods) -- `order` values
 * as rans_compress_to_4x16 takes them -- in one launch; only the winner leaves
 * the device.
 *   best[k]                index into methods[] of the first smallest stream, as
 *                          the reference's `if (best_sz > out_len)` walk in list
 *                          order keeps it (fqzcomp5.c:2097-2106); -1 if every
 *                          method failed (then out_size[k] == 0)
 *   csize[k*n_methods+j]   size under method j (what metrics_update receives);
 *                          0 = that call failed.  May be NULL.
 *   out, out_off, out_size the winning stream of input k, byte-identical to
 *                          rans_compress_4x16(in[k], in_size[k], ., methods[best[k]])
 * A failed method is never selected (the reference would keep a NULL buffer of
 * size 0 in that case; no caller relies on it). */
int b200rans_compress_methods_batch(int n,
                                    const unsigned char *const *in, const unsigned int *in_size,
                                    int n_methods, const int *methods,
                                    unsigned char *out, size_t out_cap,
                                    size_t *out_off, unsigned int *out_size,
                                    int *best, unsigned int *csize);

/* The same with a private method list per input -- the shape of tok3's compress()
 * (tokenise_name3.c:1268-1417), which brute-forces each token stream over the
 * 1-6 `order` values its type and level select: input k is tried under
 * methods[method_first[k] .. method_first[k+1]) (1-64 entries each); best[k]
 * indexes that sub-list and csize[] is laid out like methods[]. */
int b200rans_compress_trials(int n,
                             const unsigned char *const *in, const unsigned int *in_size,
                             const unsigned int *method_first, const int *methods,
                             unsigned char *out, size_t out_cap,
                             size_t *out_off, unsigned int *out_size,
                             int *best, unsigned int *csize);

/* One input, malloc()ed winner (caller free()s), NULL on failure: the shape of
 * the rANS arm of compress_with_methods for a single section buffer. */
unsigned char *b200rans_compress_methods(unsigned char *in, unsigned int in_size,
                                         int n_methods, const int *methods,
                                         unsigned int *out_size, int *best, unsigned int *csize);

/* Host buffers in, caller-placed outputs.
 *   out[k]        destination of stream k (must not be NULL)
 *   out_size[k]   in: capacity (exact length for NOSZ streams); out: bytes
 *                 written, or 0 with status[k] != 0 on failure
 *   status        optional per-stream result (0 = ok), may be NULL */
int b200rans_uncompress_batch(int n,
                              const unsigned char *const *in, const unsigned int *in_size,
                              unsigned char *const *out, unsigned int *out_size,
                              int *status);

/* Peek the uncompressed length stored in a stream header (host memory).
 * Returns -1 for NOSZ streams (length not stored) or malformed headers. */
int64_t b200rans_uncompressed_size(const unsigned char *in, unsigned int in_size);

/* ---- device-resident variants (inputs and outputs already in HBM) ---------
 * Descriptor arrays are host memory; data pointers are device memory.  Work is
 * enqueued on `stream` (a cudaStream_t; NULL = the thread context's stream)
 * and is asynchronous: results are valid once the stream has been
 * synchronised.  d_out_off / d_out_size / d_status are DEVICE arrays. */
size_t b200rans_compress_batch_dev_bound(int n, const unsigned int *in_size, const int *order);

int b200rans_compress_batch_dev(void *stream, int n,
                                const unsigned char *d_in,
                                const uint64_t *in_off, const unsigned int *in_size,
                                const int *order,
                                unsigned char *d_out, size_t out_cap,
                                uint64_t *d_out_off, unsigned int *d_out_size);

/* STRIPE streams are not accepted here (their sub-stream table must be read on
 * the host); everything else is.  out_size[k] is the expected length.
 * flags: optional HOST array holding the first byte of each stream (the format's
 * flag byte); it lets the library pick the lean order-0 kernel and skip
 * transform scratch.  NULL = unknown, the general kernel is used. */
int b200rans_uncompress_batch_dev(void *stream, int n,
                                  const unsigned char *d_in,
                                  const uint64_t *in_off, const unsigned int *in_size,
                                  const unsigned char *flags,
                                  unsigned char *d_out,
                                  const uint64_t *out_off, const unsigned int *out_size,
                                  unsigned int *d_out_size, int *d_status);

/* Multi-GPU block partitioning: streams are dealt round-robin by
 * `block_of[k]` (or by k when NULL) over the first `ngpu` devices, one worker
 * thread and stream per device, results gathered in call order.  No
 * collective is involved. */
int b200rans_compress_batch_multi(int ngpu, int n,
                                  const unsigned char *const *in, const unsigned int *in_size,
                                  const int *order, const int *block_of,
                                  unsigned char *out, size_t out_cap,
                                  size_t *out_off, unsigned int *out_size);
int b200rans_uncompress_batch_multi(int ngpu, int n,
                                    const unsigned char *const *in, const unsigned int *in_size,
                                    const int *block_of,
                                    unsigned char *const *out, unsigned int *out_size,
                                    int *status);

/* ------------------------------------------------------------------------
 * Part 3: the step either side of the codec -- a block of 4-line
 * FASTQ text split into the name / seq / qual buffers the codec is fed with,
 * and joined back, on the device.
 *
 * b200fq_split* restates load_seqs (fqzcomp5.c:279-410): names without '@',
 * NUL separated; bases and qualities concatenated without separators, the
 * qualities shifted by -33 (:375); len[] and flag[] per record (FQZ_FREAD2 = 128
 * when the name ends in "/2" or repeats the previous one, :319-327); fixed_len
 * as fq->fixed_len (:344-348); `consumed` is *last_offset: the record the block
 * ends in is not taken, and a last record whose qualities do not match its
 * bases yet is held back when it ends on the block's final byte (:382-386).
 * status 1 = the reference returns NULL (a record not starting with '@', a
 * third line not starting with '+', base / quality lengths differing); blocks
 * holding NUL bytes are also reported malformed (the reference would read a NUL
 * as a line end).  status 2 = max_records or a buffer capacity was too small.
 * n <= INT_MAX as in the reference (`int blk_size`).
 *
 * b200fq_join* is output_fastq (fqzcomp5.c:3440-3480) with the +33 of
 * fqzcomp5.c:2532-2533 folded in: "@name\nseq\n+[name]\nqual\n" per record.
 * ---------------------------------------------------------------------- */
typedef struct {
    int32_t  status;        /* 0 ok, 1 malformed, 2 capacity */
    uint32_t num_records;
    uint32_t name_len, seq_len, qual_len;   /* bytes used in the three buffers */
    int32_t  fixed_len;     /* -1 nothing seen, L > 0 every read L long, else 0 */
    uint32_t consumed;      /* split: offset of the first byte not consumed */
    uint32_t text_len;      /* join: bytes of text produced */
} b200fq_info;

/* Host buffers (pinned for full PCIe rate).  name_cap / seq_cap bytes are
 * available in name / seq and qual; n bytes each are always enough.  len and
 * flag hold max_records entries. */
int b200fq_split(const unsigned char *text, uint32_t n,
                 unsigned char *name, uint32_t name_cap,
                 unsigned char *seq, unsigned char *qual, uint32_t seq_cap,
                 uint32_t *len, uint32_t *flag, uint32_t max_records, b200fq_info *info);
int b200fq_join(const unsigned char *name, uint32_t name_len,
                const unsigned char *seq, const unsigned char *qual, uint32_t seq_len,
                const uint32_t *len, uint32_t num_records, int plus_name,
                unsigned char *text, uint32_t text_cap, b200fq_info *info);

/* Device-resident forms, asynchronous on `stream` (NULL = the context's).  All
 * pointers are device memory; d_text / d_name 16-byte aligned, d_scratch 256-byte
 * aligned and *_scratch_bytes() large; d_name_off / d_seq_off receive the offset
 * of every record in the name and seq/qual buffers (fq->name[], fq->seq[]). */
size_t b200fq_split_scratch_bytes(uint32_t n, uint32_t max_records);
int b200fq_split_dev(void *stream, const unsigned char *d_text, uint32_t n,
                     unsigned char *d_name, uint32_t name_cap,
                     unsigned char *d_seq, unsigned char *d_qual, uint32_t seq_cap,
                     uint32_t *d_len, uint32_t *d_flag, uint32_t *d_name_off, uint32_t *d_seq_off,
                     uint32_t max_records, void *d_scratch, size_t scratch_bytes, b200fq_info *d_info);
size_t b200fq_join_scratch_bytes(uint32_t name_len, uint32_t num_records);
int b200fq_join_dev(void *stream, const unsigned char *d_name, uint32_t name_len,
                    const unsigned char *d_seq, const unsigned char *d_qual,
                    const uint32_t *d_len, uint32_t num_records, int plus_name,
                    unsigned char *d_text, uint32_t text_cap,
                    void *d_scratch, size_t scratch_bytes, b200fq_info *d_info);

/* ------------------------------------------------------------------------
 * Part 4: the step after the codec -- fqzcomp5's block framing
 * (encode_block, fqzcomp5.c:2147-2280) with its CRC-32 computed on the device.
 *
 * b200fqz_crc32* is zlib's crc32(crc_in, buf, n) (the reference links zlib for
 * it: fqzcomp5.c:2268-2269, :2310-2311), any alignment, n up to 2^32-1 bytes.
 * b200fqz_assemble_block_dev builds
 *     [u32 block size = total-4][u32 num_records][u32 crc][piece 0][piece 1]...
 * in d_block, the CRC taken over everything after the CRC field (:2266-2274),
 * so a finished block leaves the GPU in one copy.  Pieces are the sections as
 * encode_block appends them (name stream; length bytes; 9-byte meta + seq
 * stream; 9-byte meta + qual stream) and may live in host or device memory.
 * Asynchronous on `stream`; *block_len (host) is known at once.  Like every
 * `_dev` entry point it uses scratch owned by the calling thread's context: keep
 * the device-resident calls of one thread on one stream (or order them yourself).
 * ---------------------------------------------------------------------- */
typedef struct {
    const void *ptr;
    uint32_t len;
    int on_device;          /* 0: host memory, 1: device memory */
} b200fqz_piece;

int b200fqz_crc32(uint32_t crc_in, const unsigned char *buf, uint64_t n, uint32_t *crc_out);
int b200fqz_crc32_dev(void *stream, const unsigned char *d_buf, uint64_t n, uint32_t crc_in, uint32_t *d_crc);
int b200fqz_assemble_block_dev(void *stream, uint32_t num_records, int n_pieces, const b200fqz_piece *pieces,
                               unsigned char *d_block, uint64_t block_cap, uint32_t *block_len);

/* Number of kernel launches issued by this thread's context so far. */
uint64_t b200rans_launch_count(void);

/* Measurement hooks: with profiling on, the encode / decode coder kernel of each
 * batch call is bracketed by CUDA events on the launching stream;
 * b200rans_last_kernel_ms(0|1) returns the duration of the last encode (0) or
 * decode (1) coder kernel in milliseconds (synchronises on its stop event),
 * or a negative value if none was recorded. */
int   b200rans_set_profiling(int on);
float b200rans_last_kernel_ms(int which);
const char *b200rans_version(void);

#ifdef __cplusplus
}
#endif
#endif /* B200RANS_H */
