// fastq.h -- host-visible launchers of the FASTQ split / join kernels (fastq.cu): the step
// either side of the codec in fqzcomp5 (load_seqs, fqzcomp5.c:279-410; output_fastq,
// fqzcomp5.c:3440-3480; the qual -33 / +33 shifts, :355 and :2532-2533).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace b200 {

// mirrors b200fq_info in include/b200rans.h
struct FqInfo {
    int32_t  status;        // 0 ok; 1 malformed (load_seqs returns NULL); 2 a capacity was too small
    uint32_t num_records;
    uint32_t name_len, seq_len, qual_len;
    int32_t  fixed_len;     // fq->fixed_len: -1 nothing seen, L > 0 all reads L long, else 0
    uint32_t consumed;      // *last_offset: where the first record not taken starts
    uint32_t text_len;      // join: bytes of FASTQ text produced
    uint32_t more;          // split, kseq mode: 1 = the block-size rule ended the block, 0 = the text ran out
};

size_t fq_split_scratch_bytes(uint32_t n, uint32_t max_records);
cudaError_t fq_split_launch(const uint8_t *d_text, uint32_t n, uint8_t *d_name, uint8_t *d_seq, uint8_t *d_qual,
                            uint32_t name_cap, uint32_t seq_cap, uint32_t *d_len, uint32_t *d_flag,
                            uint32_t *d_name_off, uint32_t *d_seq_off, uint32_t max_records, uint8_t *d_scratch,
                            FqInfo *d_info, cudaStream_t st, int *launches, int kseq = 0, uint32_t blk_size = 0);
// kseq != 0: load_seqs_kseq's rules (fqzcomp5.c:423-623) for strict 4-line FASTQ -- kseq's name / comment
// split, records taken while name.l + 1 + seq.l + qual.l sums to <= blk_size (one at least).

size_t fq_join_scratch_bytes(uint32_t name_len, uint32_t num_records);
cudaError_t fq_join_launch(const uint8_t *d_name, uint32_t name_len, const uint8_t *d_seq, const uint8_t *d_qual,
                           const uint32_t *d_len, uint32_t num_records, int plus_name, uint8_t *d_text,
                           uint32_t text_cap, uint8_t *d_scratch, FqInfo *d_info, cudaStream_t st, int *launches);

}  // namespace b200
