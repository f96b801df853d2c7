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

/* ---- device-resident variants, second form -------------------------------
 * b200rans_compress_batch_dev2 adds `flags`:
 *   B200RANS_OUT_IN_SLOT   every call gets its own rans_compress_bound_4x16-sized,
 *                          256-byte aligned slot inside d_out -- the reference's
 *                          "one bound-sized buffer per call" (fqzcomp5.c:1996) -- and the
 *                          finished stream stays where the encoder built it
 *                          (d_out_off[k] = its first byte); no packing pass runs.
 *                          out_cap >= b200rans_compress_slots_bound(n, in_size, order).
 *   0                      as b200rans_compress_batch_dev (streams packed back to back).
 * b200rans_compress_trials_dev is b200rans_compress_trials with inputs and outputs in
 * HBM: d_best [n], d_csize [method_first[n]] (either may be NULL) are DEVICE arrays;
 * pack_align (a power of two, 1 = no gaps) places the winners in d_out.
 * b200rans_tok3_methods fills out[] (<= 8 entries) with the candidate `order` values
 * tok3's compress() walks for one token stream (tokenise_name3.c:1283-1357 tables with
 * the filters of :1374-1378: X32 cleared, STRIPE dropped when in_len % 4), returning
 * their number: what a C caller passes to b200rans_compress_trials per stream. */
#define B200RANS_OUT_IN_SLOT 1u
size_t b200rans_compress_slots_bound(int n, const unsigned int *in_size, const int *order);
int b200rans_compress_batch_dev2(void *stream, int n,
                                 const unsigned char *d_in,
                                 const uint64_t *in_off, const unsigned int *in_size,
                                 const int *order,
                                 unsigned char *d_out, size_t out_cap,
                                 uint64_t *d_out_off, unsigned int *d_out_size, unsigned int flags);
int b200rans_compress_trials_dev(void *stream, int n,
                                 const unsigned char *d_in,
                                 const uint64_t *in_off, const unsigned int *in_size,
                                 const unsigned int *method_first, const int *methods,
                                 unsigned char *d_out, size_t out_cap, unsigned int pack_align,
                                 uint64_t *d_out_off, unsigned int *d_out_size,
                                 int *d_best, unsigned int *d_csize);
int b200rans_tok3_methods(int level, int token_type, unsigned int in_len, int *out);

/* Multi-GPU block partitioning: streams are dealt round-robin by
 * `block_of[k]` >= 0 (or by k when NULL) over the first `ngpu` devices, results
 * gathered in call order (the hts_tpool contract, thread_pool.c:113-164).  The
 * worker threads -- up to B200RANS_WORKERS_PER_DEVICE per device, each with its own
 * context, streams and arenas -- are created on first use and kept for the life
 * of the process.  The block calls keep B200RANS_BLOCK_WORKERS of them busy per device
 * (environment, default 2: one block copying in while another is being coded).
 * No collective is involved. */
#define B200RANS_WORKERS_PER_DEVICE 4
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
    uint32_t more;          /* split, kseq mode: 1 = the block-size rule ended the block (records are left for
                               the next one), 0 = the text ran out first */
} b200fq_info;

/* The live loader, load_seqs_kseq (fqzcomp5.c:423-623; main() reaches only this one, through encode_gzip
 * :3051): B200FQ_MODE_KSEQ applies its rules to strict 4-line FASTQ text --
 *   - kseq's record (kseq.h:176-218): name = header up to its first isspace() byte, comment = the rest;
 *     the stored name is name [+ ' ' + comment unless that is empty] (a tab becomes a space, a trailing
 *     separator is dropped, fqzcomp5.c:485-510);
 *   - a block takes records while name.l + 1 + seq.l + qual.l sums to <= blk_size, one at least
 *     (:468-476); `consumed` is where the first record left for the next block starts, info.more tells
 *     whether the rule or the end of the text stopped it;
 *   - FQZ_FREAD2: stored name ends in "/2" (name longer than one byte) or equals the previous one (:512-520);
 *   - fixed_len over the records taken (:536-541).
 * Not handled, reported as status 1: multi-line records, FASTA ('>' headers or no quality line; the
 * reference switches to is_fasta, :566-571), sequence lines starting with '@', '+' or '>', CRLF line
 * ends, base / quality lengths that differ (kseq would read further quality lines or return -2), and a
 * last record without its newline.  The record behind the last one taken is checked too, as the
 * reference has parsed it before it keeps it for the next block. */
#define B200FQ_MODE_LOAD_SEQS 0
#define B200FQ_MODE_KSEQ      1
int b200fq_split_mode(int mode, uint32_t blk_size, const unsigned char *text, uint32_t n,
                      unsigned char *name, uint32_t name_cap,
                      unsigned char *seq, unsigned char *qual, uint32_t seq_cap,
                      uint32_t *len, uint32_t *flag, uint32_t max_records, b200fq_info *info);

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
int b200fq_split_dev_mode(void *stream, int mode, uint32_t blk_size, const unsigned char *d_text, uint32_t n,
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

/* ------------------------------------------------------------------------
 * Part 5: one call per fqzcomp5 block (encode_block / decode_block,
 * fqzcomp5.c:2147-2280, :2290-2547) -- FASTQ text in, framed block out, and
 * back, everything between on the device: split (part 3), method trial over
 * the section slices (part 2), framing + CRC-32 (part 4); decode: CRC check,
 * batched decode, join.
 *
 * Sections are cut into `slice_bytes` pieces (rounded down to whole reads when
 * the read length is fixed), each an independent rans_compress_to_4x16 stream
 * tried under the section's method list, first smallest kept (one
 * call has 4 or 32 serial lanes, a GPU needs thousands).  slice_bytes == 0
 * codes every section as ONE stream: the layout of a section is then exactly
 * the reference's ([u8 strat = 0][u32 ulen][u32 clen][stream], :2218-2232) and
 * a block whose name stream the caller supplies is a stock FQZ5 block.
 * With slices the section is
 *   [u8 strat = B200FQZ_STRAT_SLICED][u32 ulen][u32 clen]
 *   [u32 nslices][u32 slice_bytes][u32 csize[nslices]][streams back to back]
 * which stock fqzcomp5 does not read (its format has one stream per section).
 *
 * Methods are `order` values as rans_compress_to_4x16 takes them;
 * B200FQZ_RANSXN1 in a list stands for RANSXN1 = (fixed_len << 8) + 9 and is
 * skipped when the block's reads differ in length (fqzcomp5.c:2013-2022).
 * Names: with n_name_methods > 0 the name buffer is a third section coded like
 * the others (the codec half of TLZP3 = LZP + order 5, fqzcomp5.c:2024-2028);
 * with 0 the caller's `name_coder` turns the name buffer into the name section
 * (the tokeniser is host code) while the device runs the trials.
 * ---------------------------------------------------------------------- */
#define B200FQZ_RANSXN1      (-1)
#define B200FQZ_STRAT_SLICED 0xB2
#define B200FQZ_MAX_METHODS  16

typedef int (*b200fqz_name_coder)(void *user, const unsigned char *names, uint32_t name_len,
                                  const uint32_t *flags, uint32_t num_records,
                                  unsigned char **out, uint32_t *out_len);   /* *out: malloc()ed, freed by the library */
typedef struct {
    uint32_t slice_bytes;
    int n_name_methods, n_seq_methods, n_qual_methods;
    int name_methods[B200FQZ_MAX_METHODS], seq_methods[B200FQZ_MAX_METHODS], qual_methods[B200FQZ_MAX_METHODS];
    b200fqz_name_coder name_coder;
    void *name_user;
    uint32_t kseq_blk_size;     /* 0: load_seqs' block rule (the text given IS the block); > 0: B200FQ_MODE_KSEQ
                                   with this blk_size, rep->consumed tells where the next block starts */
} b200fqz_block_opts;

typedef struct {
    int32_t  status;                /* 0 ok; 1 malformed input (the reference returns NULL); 3 CRC mismatch */
    uint32_t num_records, consumed; /* encode: bytes of text taken (load_seqs' *last_offset) */
    int32_t  fixed_len;
    uint32_t ulen[3], clen[3];      /* name, seq, qual: section bytes in and coded */
    uint32_t nslices[3];
    uint64_t csize[3][B200FQZ_MAX_METHODS];   /* per method: summed size over the slices (metrics_update's input) */
    uint32_t wins[3][B200FQZ_MAX_METHODS];    /* per method: slices it won */
    uint32_t block_len;             /* encode: bytes of block produced; decode: bytes of text produced */
    uint32_t crc;
    float    ms[4];                 /* encode, host wall time of the call's phases: copy in + split; planning and
                                       queueing the trials; waiting for them (+ the name coder); framing + CRC +
                                       copy out */
} b200fqz_block_report;

/* The learner that decides which methods a block is coded with (metrics_method / metrics_update,
 * fqzcomp5.c:1899-1958; METRICS_TRIAL = 3 and METRICS_REVIEW = 100 at :151-152), kept per section
 * (0 names, 1 seq, 2 qual) as the reference keeps stats[sec]: while a trial is on every method of the
 * caller's list is tried and the sizes are accumulated; when it ends the method with the smallest
 * (csize + 1) / usize is used alone for the next METRICS_REVIEW blocks, then the next trial starts.
 * Plain host logic, no device involved:
 *   b200fqz_learner_methods(L, all, out)  copies *all to *out with each section's list cut down to what
 *                                         metrics_method returns for this block;
 *   b200fqz_learner_update(L, out, rep)   feeds the sizes of the finished block back (sections that were
 *                                         on trial for it); `out` is what that block was coded with. */
typedef struct {
    int32_t  review[3], trial[3], used[3];          /* used: index into the caller's full method list */
    uint64_t usize[3][B200FQZ_MAX_METHODS], csize[3][B200FQZ_MAX_METHODS];
    int32_t  on_trial[3];                           /* set by _methods: this block's sizes count */
} b200fqz_learner;
void b200fqz_learner_init(b200fqz_learner *L);
void b200fqz_learner_methods(b200fqz_learner *L, const b200fqz_block_opts *all, b200fqz_block_opts *out);
void b200fqz_learner_update(b200fqz_learner *L, const b200fqz_block_opts *out, const b200fqz_block_report *rep);

/* Host buffers (pinned for full PCIe rate).  text[0, n) holds FASTQ text; the block is written to
 * block[0, block_cap); b200fqz_block_bound(n) bytes always suffice. */
size_t b200fqz_block_bound(uint32_t n);
int b200fqz_encode_block(const unsigned char *text, uint32_t n, const b200fqz_block_opts *opts,
                         unsigned char *block, size_t block_cap, b200fqz_block_report *rep);
/* block[0, block_len) as written by b200fqz_encode_block with n_name_methods > 0; text_cap bytes at text. */
int b200fqz_decode_block(const unsigned char *block, uint32_t block_len, int plus_name,
                         unsigned char *text, size_t text_cap, b200fqz_block_report *rep);
/* nblocks independent blocks dealt round-robin over the first ngpu devices (block b on device
 * b % ngpu) by the persistent workers; results in call order. rep[b].status is per block. */
int b200fqz_encode_blocks_multi(int ngpu, int nblocks, const unsigned char *const *text, const uint32_t *n,
                                const b200fqz_block_opts *opts, unsigned char *const *block,
                                const size_t *block_cap, b200fqz_block_report *rep);
int b200fqz_decode_blocks_multi(int ngpu, int nblocks, const unsigned char *const *block,
                                const uint32_t *block_len, int plus_name, unsigned char *const *text,
                                const size_t *text_cap, b200fqz_block_report *rep);

/* Number of kernel launches issued by this thread's context so far (workers of the
 * multi-GPU calls add theirs when a call returns). */
uint64_t b200rans_launch_count(void);

/* Measurement hooks: with profiling on, the encode / decode coder kernel of each
 * batch call is bracketed by CUDA events on the launching stream;
 * b200rans_last_kernel_ms(0|1) returns the duration of the last encode (0) or
 * decode (1) coder kernel in milliseconds (synchronises on its stop event),
 * or a negative value if none was recorded. */
int   b200rans_set_profiling(int on);
float b200rans_last_kernel_ms(int which);
/* Diagnostics of the staged order-1 decode (streams behind PACK / RLE with a known flag byte): what became
 * of the streams its head stage looked at on the current device since the last reset -- out16[0] taken,
 * [1] taken and then failed, [2..9] handed back to the general kernel (header, pack meta-data, run-length
 * header, no payload, order-1 header, table peek, alphabet, alphabet of <= 64 symbols).  Synchronises. */
int   b200rans_dec_staged_stats(unsigned long long out16[16], int reset);
const char *b200rans_version(void);

#ifdef __cplusplus
}
#endif
#endif /* B200RANS_H */
