/*
 * block_pipeline.c -- the calls a fqzcomp5 maintainer would make from C,
 * end to end on one FASTQ block held in host memory:
 *
 *   load_seqs            -> b200fq_split                 (fqzcomp5.c:279-410)
 *   compress_with_methods -> b200rans_compress_methods   (fqzcomp5.c:1979-2119, rANS members)
 *   zlib crc32            -> b200fqz_crc32               (fqzcomp5.c:2268-2269)
 *   rans_uncompress_4x16  -> unchanged drop-in symbol    (rANS_static4x16.h:49-50)
 *   output_fastq          -> b200fq_join                 (fqzcomp5.c:3440-3480)
 *
 * Build:  gcc -std=c99 -Iinclude examples/block_pipeline.c -Lfqzcomp5_b200 -lb200rans \
 *             -Wl,-rpath,$PWD/fqzcomp5_b200 -o block_pipeline
 * Run:    ./block_pipeline reads.fastq        (needs a CUDA device: the library has no CPU path)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "b200rans.h"

int main(int argc, char **argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s reads.fastq\n", argv[0]); return 2; }
    FILE *fp = fopen(argv[1], "rb");
    if (!fp) { perror(argv[1]); return 1; }
    fseek(fp, 0, SEEK_END);
    long fsz = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    uint32_t n = (uint32_t)fsz;
    unsigned char *text = malloc(n + 1);
    if (fread(text, 1, n, fp) != n) { fprintf(stderr, "short read\n"); return 1; }
    fclose(fp);

    /* split the block */
    uint32_t max_records = n / 16 + 16;
    unsigned char *name = malloc(n + 16), *seq = malloc(n + 16), *qual = malloc(n + 16);
    uint32_t *len = malloc(4 * (size_t)max_records), *flag = malloc(4 * (size_t)max_records);
    b200fq_info fi;
    memset(&fi, 0, sizeof fi);
    int rc = b200fq_split(text, n, name, n + 16, seq, qual, n + 16, len, flag, max_records, &fi);
    if (rc || fi.status) {
        fprintf(stderr, "split failed (library status %d, block status %d)\n", rc, fi.status);
        return 1;
    }
    printf("%u records, %u name / %u seq / %u qual bytes, fixed_len %d, %u bytes consumed\n", fi.num_records,
           fi.name_len, fi.seq_len, fi.qual_len, fi.fixed_len, fi.consumed);

    /* trial the fast-mode rANS methods on seq and qual, keep the smallest of each */
    int methods[5] = {0 | 4, 1 | 4, 129 | 4, 193 | 4, 0}, nm = 4;
    if (fi.fixed_len > 0) methods[nm++] = (fi.fixed_len << 8) + 9;        /* RANSXN1, qual only */
    unsigned int csize[5], seq_clen = 0, qual_clen = 0;
    int best;
    unsigned char *cseq = b200rans_compress_methods(seq, fi.seq_len, 4, methods, &seq_clen, &best, csize);
    if (!cseq) { fprintf(stderr, "seq trial failed\n"); return 1; }
    printf("seq : method %#x wins, %u -> %u bytes\n", methods[best], fi.seq_len, seq_clen);
    unsigned char *cqual = b200rans_compress_methods(qual, fi.qual_len, nm, methods, &qual_clen, &best, csize);
    if (!cqual) { fprintf(stderr, "qual trial failed\n"); return 1; }
    printf("qual: method %#x wins, %u -> %u bytes\n", methods[best], fi.qual_len, qual_clen);

    /* CRC of what would follow the CRC field of the block (here: just the two streams) */
    uint32_t crc = 0;
    b200fqz_crc32(0, cseq, seq_clen, &crc);
    b200fqz_crc32(crc, cqual, qual_clen, &crc);
    printf("crc32 of the two streams: %08x\n", crc);

    /* and back: the drop-in decoder, then the join */
    unsigned int ulen = 0;
    unsigned char *seq2 = rans_uncompress_4x16(cseq, seq_clen, &ulen);
    unsigned char *qual2 = rans_uncompress_4x16(cqual, qual_clen, &ulen);
    if (!seq2 || !qual2 || memcmp(seq2, seq, fi.seq_len) || memcmp(qual2, qual, fi.qual_len)) {
        fprintf(stderr, "round trip mismatch\n");
        return 1;
    }
    unsigned char *back = malloc((size_t)n + 64);
    b200fq_info ji;
    memset(&ji, 0, sizeof ji);
    if (b200fq_join(name, fi.name_len, seq2, qual2, fi.seq_len, len, fi.num_records, 0, back, n + 64, &ji) || ji.status) {
        fprintf(stderr, "join failed\n");
        return 1;
    }
    printf("joined %u bytes of text, %s the consumed part of the input\n", ji.text_len,
           ji.text_len == fi.consumed && !memcmp(back, text, ji.text_len) ? "identical to" : "DIFFERENT from");
    free(cseq); free(cqual); free(seq2); free(qual2);
    return 0;
}
