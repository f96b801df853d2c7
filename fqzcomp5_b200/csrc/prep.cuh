// prep.cuh -- everything in front of the state chains of a PACK / RLE stream, done by a whole
// CTA (8 warps) per stream instead of the stream's coder warp: PACK, the RLE symbol choice and
// split, the order-0 counts and -- for order-1 streams -- the complete model: pair counts, the
// precision decision, the 256 x normalise_freq, the serialised table and the encoder symbols.
// The coder warp (enc_stream in kernels.cu) then only assembles the header and runs the chains
// (run-length meta-data, table self-compression, payload).
//
// Reference behaviour restated here (never its code):
//   pack.c:56-147 (hts_pack), rle.c:48-138 (symbol choice, encode),
//   rANS_static16_int.h:312-421 (encode_freq1), rANS_static4x16pr.c:357-420 (rans_compute_shift),
//   utils.h:279-357 (hist1_4).
#pragma once
#include "common.cuh"
#include "rans_encode.cuh"

namespace b200 {

#ifdef B200_PREP_PROF
__device__ unsigned long long g_prep_prof[16];
#define PREP_T0() long long t_prof = clock64()
#define PREP_T(i) do { __syncthreads(); if (threadIdx.x == 0) { long long t2 = clock64(); atomicAdd(&g_prep_prof[i], (unsigned long long)(t2 - t_prof)); t_prof = t2; } } while (0)
#else
#define PREP_T0() do {} while (0)
#define PREP_T(i) do {} while (0)
#endif

constexpr int PREP_THREADS = 256, PREP_WARPS = 8;

// What the CTA leaves for the coder warp, at the head of the job's prep area (J.prep).
struct __align__(16) Prep {
    uint32_t state;         // 0: not prepared (the coder warp does everything itself), 1: ready
    uint32_t packed;        // PACK: 1 applied (pack.c returned a buffer), 0 refused (> 16 symbols)
    uint32_t pmeta, plen;   // PACK: bytes of its meta-data (already at slot + meta), packed length
    uint32_t rle_len, rmeta_len;    // RLE: literals and meta-data bytes as hts_rle_encode returns them
    uint32_t model;         // 0 none, 1 order-0 counts in F, 2 order-1 model (table, symbols, rank)
    uint32_t nsym, shift, tl;       // order 1: alphabet size, precision, bytes of the uncompressed table
    uint32_t err;           // order 1: normalise_freq failed (the reference returns NULL)
    uint32_t stream_syms;   // order 1: 1 = encoder symbols per position (E), 0 = symbol table
    uint32_t pad_[4];
    uint32_t F[256];        // order-0 counts of the data that reaches the coder
    uint8_t  rank[256];     // order 1: symbol -> rank
};
// layout of the prep area behind the header
constexpr uint32_t PREP_ROW_STRIDE = 768;      // a serialised row: at most 2 bytes per column plus a closing run token
// Alphabets whose nsym x nsym encoder symbols fit the coder warp's shared memory keep only the symbol table; for
// larger ones the CTA also looks up the encoder symbol of every POSITION while its table is still hot in L2
// (all 256 threads, 16 look-ups in flight each) and leaves them as a 4-byte stream the chains read sequentially:
// looked up from the chains, much later, a 256 x 256 table per stream is a DRAM sector per symbol.
constexpr uint32_t PREP_SYM_MAX = 40;
struct PrepPlan { uint32_t o_bkt, o_rows, o_sym, o_E, o_tbl, o_tmp, total; };
__host__ __device__ inline PrepPlan prep_plan(uint32_t isz) {
    const uint32_t m = isz + 1 < 256 ? isz + 1 : 256;                      // alphabet of a short stream
    const uint32_t hw = m * m * 4;
    const uint32_t tbl = (4 * (uint64_t)isz + 2048 < 257 * 257 * 3 + 4 ? 4 * isz + 2048 : 257 * 257 * 3 + 4) + 64;
    PrepPlan p;
    p.o_bkt = (uint32_t)((sizeof(Prep) + 255) & ~255u);                     // pairs dealt by context: isz + 32 bytes
    p.o_rows = p.o_bkt + ((isz + 64 + 255) & ~255u);                        // serialised rows before they are packed
    p.o_sym = p.o_rows + ((m * PREP_ROW_STRIDE + 255) & ~255u);             // encoder symbols (rank(ctx), rank(sym))
    p.o_E = p.o_sym + ((hw + 255) & ~255u);                                 // encoder symbol of every position (8 bytes)
    p.o_tbl = p.o_E + ((8 * (isz + 64) + 255) & ~255u);                     // the table, uncompressed
    p.o_tmp = p.o_tbl + ((tbl + 255) & ~255u);                              // scratch of the table's order-0 coder
    p.total = p.o_tmp + ((compress_bound(tbl, 0) + 64 + 255) & ~255u);
    return p;
}

struct __align__(16) PrepSmem {
    uint32_t Fw[PREP_WARPS][256];   // warp-private order-0 bins; PACK: byte flags / codes in Fw[0]; RLE: scores in Fw[1];
                                    // order 1: per-warp scratch of the greedy shave
    uint32_t Hs[256];               // bucket cursors of the pair dealing
    uint32_t bstart[260];           // first byte of every context's bucket (and the end of the last)
    uint32_t T[256];                // order-0 counts (symbol space), then row totals (rank space)
    uint32_t rowlen[256];           // serialised row lengths, then offsets
    uint16_t S16[256];              // stored total of each row
    uint8_t  rank[256], sym[256];
    uint32_t pres[8];               // alphabet bitmap (symbol space)
    uint32_t wtot[PREP_WARPS];
    double   red[2][PREP_WARPS];
    uint32_t redu[PREP_WARPS];
    // RLE chunk states
    uint32_t c_nl[PREP_WARPS], c_nr[PREP_WARPS], c_open[PREP_WARPS], c_start[PREP_WARPS], c_first[PREP_WARPS];
    uint32_t bc[8];                 // broadcast scalars
};

// rank of every set flag among 256 (thread t owns flag t); returns the number set
__device__ __forceinline__ uint32_t cta_rank256(bool pres, uint32_t *wtot, uint32_t *my_rank) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t bal = __ballot_sync(FULL, pres);
    if (lane == 0) wtot[wid] = __popc(bal);
    __syncthreads();
    uint32_t before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < PREP_WARPS; w++) { const uint32_t c = wtot[w]; if (w < wid) before += c; total += c; }
    *my_rank = before + __popc(bal & lanemask_lt());
    __syncthreads();
    return total;
}

// ------------------------------------------------------------------ PACK (pack.c:56-147)
__device__ inline bool cta_pack(const uint8_t *in, uint32_t n, uint8_t *meta, uint32_t *meta_len, uint8_t *out,
                                uint32_t *out_len, PrepSmem &S) {
    const int tid = threadIdx.x;
    uint8_t *code = (uint8_t *)S.Fw[0];
    if (tid < 64) ((uint32_t *)code)[tid] = 0;
    __syncthreads();
    uint32_t head = (uint32_t)((16 - ((uintptr_t)in & 15)) & 15);
    if (head > n) head = n;
    {   // presence flags (benign write races)
        if ((uint32_t)tid < head) code[in[tid]] = 1;
        const uint8_t *p = in + head;
        const uint32_t rest = n - head, nv = rest >> 4;
        const uint4 *v = (const uint4 *)p;
        for (uint32_t i = tid; i < nv; i += PREP_THREADS) {
            const uint4 q = __ldg(v + i);
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) code[(w[a] >> (8 * b)) & 0xff] = 1;
        }
        for (uint32_t i = (nv << 4) + tid; i < rest; i += PREP_THREADS) code[p[i]] = 1;
    }
    __syncthreads();
    const bool pres = code[tid] != 0;
    uint32_t r;
    const uint32_t nsym = cta_rank256(pres, S.wtot, &r);
    if (pres) { meta[1 + r] = (uint8_t)tid; code[tid] = (uint8_t)r; }           // pack.c:65-70 writes every listed symbol
    if (tid == 0) meta[0] = (uint8_t)nsym;                                      // 256 wraps to 0
    __syncthreads();
    if (nsym > 16) return false;
    *meta_len = nsym + 1;
    const uint32_t per = nsym > 4 ? 2 : nsym > 2 ? 4 : nsym > 1 ? 8 : 0;
    if (!per) { *out_len = 0; return true; }
    const uint32_t bits = 8 / per, olen = (n + per - 1) / per;
    uint32_t jdone = 0;
    if ((((uintptr_t)in) & 15) == 0) {
        const uint4 *v = (const uint4 *)in;
        const uint32_t nv = n >> 4, ob = 16 / per;
        for (uint32_t vi = tid; vi < nv; vi += PREP_THREADS) {
            const uint4 q = __ldg(v + vi);
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
            uint32_t c[16];
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) c[4 * a + b] = code[(w[a] >> (8 * b)) & 0xff];
            if (per == 4) {
                uint32_t x = 0;
#pragma unroll
                for (int k = 0; k < 16; k++) x |= c[k] << (2 * k);
                *(uint32_t *)(out + vi * 4) = x;
            } else if (per == 2) {
                uint32_t x0 = 0, x1 = 0;
#pragma unroll
                for (int k = 0; k < 8; k++) { x0 |= c[k] << (4 * k); x1 |= c[8 + k] << (4 * k); }
                *(uint2 *)(out + vi * 8) = make_uint2(x0, x1);
            } else {
                uint32_t x = 0;
#pragma unroll
                for (int k = 0; k < 16; k++) x |= c[k] << k;
                *(uint16_t *)(out + vi * 2) = (uint16_t)x;
            }
        }
        jdone = nv * ob;
    }
    for (uint32_t j = jdone + tid; j < olen; j += PREP_THREADS) {
        uint32_t v = 0;
        const uint32_t base = j * per;
        for (uint32_t q = 0; q < per && base + q < n; q++) v |= (uint32_t)code[in[base + q]] << (q * bits);
        out[j] = (uint8_t)v;
    }
    *out_len = olen;
    __syncthreads();
    return true;
}

// ------------------------------------------------------------------ RLE (rle.c:48-138)
// A symbol is run-length coded iff it repeats its predecessor more often than not.  Literals: one byte per
// maximal run of such a symbol, one byte per occurrence otherwise; (run length - 1) goes to a varint stream in
// run order.  Each warp owns a contiguous chunk and writes its literals and varints to staging areas in ONE pass;
// a run still open at the end of a chunk is closed by the first emitting position of a later chunk (its varint
// is the last one of the chunk that opened it); the pieces are then copied to their final, packed places.
__device__ __forceinline__ void rle_chunk(const uint8_t *in, uint32_t lo, uint32_t hi, const int32_t *score,
                                          uint8_t *lits, uint8_t *runs, uint32_t &nl, uint32_t &nr, bool &open,
                                          uint32_t &open_start, uint32_t &first, int lane) {
    const uint32_t lt = lanemask_lt();
    nl = 0; nr = 0; open = false; open_start = 0; first = 0xffffffffu;
    if (lo >= hi) return;
    uint32_t carry = lo ? in[lo - 1] : 0x200u;              // byte in front of the round
    uint32_t cn = lo + lane < hi ? in[lo + lane] : 0x100u;  // next round's bytes, one round ahead
    for (uint32_t base = lo; base < hi; base += 32) {
        const uint32_t p = base + lane;
        const bool valid = p < hi;
        const uint32_t c = cn;
        cn = base + 32 + lane < hi ? in[base + 32 + lane] : 0x100u;
        uint32_t pc = __shfl_up_sync(FULL, c, 1);
        if (lane == 0) pc = carry;
        carry = __shfl_sync(FULL, c, 31);
        const bool isr = valid && score[c & 0xff] > 0;
        const uint32_t anyr = __ballot_sync(FULL, isr);
        if (!anyr && !open) {                               // nothing run-length coded here: every byte is a literal
            if (first == 0xffffffffu) first = base;
            if (valid) lits[nl + lane] = (uint8_t)c;
            nl += min(32u, hi - base);
            continue;
        }
        const bool cont = isr && p && pc == c;              // continues a run: emits nothing
        const bool emit = valid && !cont;
        const uint32_t E = __ballot_sync(FULL, emit);
        const uint32_t R = __ballot_sync(FULL, emit && isr);
        if (!E) continue;
        if (first == 0xffffffffu) first = base + __ffs(E) - 1;
        const uint32_t below = E & lt;
        const bool prev_in = below != 0;
        const int pl = 31 - __clz(below);
        const bool prev_open = prev_in ? ((R >> pl) & 1) : open;
        const uint32_t prev_pos = prev_in ? base + pl : open_start;
        const bool closes = emit && prev_open;
        const uint32_t rl = closes ? p - prev_pos - 1 : 0;
        const uint32_t C = __ballot_sync(FULL, closes);
        if (C) {
            const uint32_t vs = closes ? var_size_u32(rl) : 0;
            const uint32_t vincl = warp_incl_scan(vs, lane);
            if (closes) var_put_u32(runs + nr + vincl - vs, rl);
            nr += __shfl_sync(FULL, vincl, 31);
        }
        if (emit) lits[nl + __popc(E & lt)] = (uint8_t)c;
        nl += __popc(E);
        const int hl = 31 - __clz(E);
        open = (R >> hl) & 1;
        open_start = base + hl;
    }
}

// lits / meta: final buffers; meta has room for 3 * n + 1024 bytes (its own n + 321 at most, then the two staging areas)
__device__ inline void cta_rle_encode(const uint8_t *in, uint32_t n, uint8_t *lits, uint32_t *lits_len, uint8_t *meta,
                                      uint32_t *meta_len, PrepSmem &S) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int32_t *score = (int32_t *)S.Fw[1];
    score[tid] = 0;
    __syncthreads();
    {   // score[c] += (in[p-1] == c) ? +1 : -1 over all positions (the first byte has no predecessor)
        uint32_t done = 0;
        if ((((uintptr_t)in) & 15) == 0) {
            const uint4 *v = (const uint4 *)in;
            const uint32_t nv = n >> 4;
            for (uint32_t i = tid; i < nv; i += PREP_THREADS) {
                const uint4 q = v[i];                       // (`in` may have been written by this kernel: plain loads)
                const uint32_t prev = i ? in[16 * i - 1] : 0x100u;
                const uint32_t w[4] = {q.x, q.y, q.z, q.w};
                uint32_t cur = w[0] & 0xff;
                int acc = cur == prev ? 1 : -1;
#pragma unroll
                for (int k = 1; k < 16; k++) {
                    const uint32_t c = (w[k >> 2] >> (8 * (k & 3))) & 0xff;
                    if (c == cur) acc++;
                    else { atomicAdd(&score[cur], acc); cur = c; acc = -1; }
                }
                atomicAdd(&score[cur], acc);
            }
            done = nv << 4;
        }
        for (uint32_t p = done + tid; p < n; p += PREP_THREADS) {
            const uint32_t c = in[p];
            atomicAdd(&score[c], (p && in[p - 1] == c) ? 1 : -1);
        }
    }
    __syncthreads();
    uint32_t r;
    const bool pr = score[tid] > 0;
    const uint32_t nsyms = cta_rank256(pr, S.wtot, &r);
    if (pr) meta[1 + r] = (uint8_t)tid;
    if (tid == 0) meta[0] = (uint8_t)nsyms;
    uint8_t *runs = meta + 1 + nsyms;
    // chunks of whole 32-byte rounds; staging: chunk w's varints at stage_r + lo + 8w, its literals at stage_l + lo
    const uint32_t CH = (((n + PREP_WARPS - 1) / PREP_WARPS) + 31) & ~31u;
    const uint32_t lo = min(n, (uint32_t)wid * CH), hi = min(n, lo + CH);
    // (both staging areas lie behind everything the final meta-data can occupy: a warp packing its piece must
    // not write over the staged piece of another)
    uint8_t *stage_r = meta + ((n + 400 + 15) & ~15u), *stage_l = stage_r + ((n + 128 + 15) & ~15u);
    uint32_t nl, nr, ostart, first;
    bool open;
    rle_chunk(in, lo, hi, score, stage_l + lo, stage_r + lo + 8 * wid, nl, nr, open, ostart, first, lane);
    if (lane == 0) { S.c_nl[wid] = nl; S.c_nr[wid] = nr; S.c_first[wid] = first; }
    __syncthreads();
    // the run open at the end of this chunk ends at the first emitter behind it (or at n)
    uint32_t next_first = n;
    for (int w = PREP_WARPS - 1; w > wid; w--) if (S.c_first[w] != 0xffffffffu) next_first = S.c_first[w];
    if (open) {
        const uint32_t tail_rl = next_first - ostart - 1;
        if (lane == 0) var_put_u32(stage_r + lo + 8 * wid + nr, tail_rl);
        nr += var_size_u32(tail_rl);
    }
    __syncthreads();
    if (lane == 0) S.c_nr[wid] = nr;
    __syncthreads();
    uint32_t lit_off = 0, run_off = 0, lit_tot = 0, run_tot = 0;
    for (int w = 0; w < PREP_WARPS; w++) {
        if (w < wid) { lit_off += S.c_nl[w]; run_off += S.c_nr[w]; }
        lit_tot += S.c_nl[w]; run_tot += S.c_nr[w];
    }
    __syncwarp();
    warp_copy(lits + lit_off, stage_l + lo, nl, lane);
    warp_copy(runs + run_off, stage_r + lo + 8 * wid, nr, lane);
    *lits_len = lit_tot;
    *meta_len = 1 + nsyms + run_tot;
    __syncthreads();
}

// ------------------------------------------------------------------ order-0 counts (utils.h:145-244)
// into S.T (symbol space); warp-private bins, equal neighbours merged before an atomic
__device__ inline void cta_hist8(const uint8_t *in, uint32_t n, PrepSmem &S) {
    const int tid = threadIdx.x, wid = tid >> 5;
    for (int j = tid; j < PREP_WARPS * 256; j += PREP_THREADS) (&S.Fw[0][0])[j] = 0;
    __syncthreads();
    uint32_t head = (uint32_t)((16 - ((uintptr_t)in & 15)) & 15);
    if (head > n) head = n;
    const uint8_t *p = in + head;
    const uint32_t rest = n - head, nv = rest >> 4;
    const uint4 *v = (const uint4 *)p;
    uint32_t *F = S.Fw[wid];
    if ((uint32_t)tid < head) atomicAdd(&F[in[tid]], 1u);
    for (uint32_t i = tid; i < nv; i += PREP_THREADS) hist16(v[i], F);
    for (uint32_t t = (nv << 4) + tid; t < rest; t += PREP_THREADS) atomicAdd(&F[p[t]], 1u);
    __syncthreads();
    uint32_t f = 0;
#pragma unroll
    for (int w = 0; w < PREP_WARPS; w++) f += S.Fw[w][tid];
    S.T[tid] = f;
    __syncthreads();
}

// ------------------------------------------------------------------ order-1 model (rANS_static16_int.h:312-421)
// Pair counts without a 256 x 256 table: the number of pairs per context follows from the order-0 counts, so
// every byte is dealt into its CONTEXT's bucket (a counting sort on the previous byte; the first byte follows
// symbol 0, utils.h:279-357).  Row i of the matrix is then the histogram of bucket i -- about n / nsym bytes --
// which a warp takes in its private 256 shared-memory bins whenever it needs the row.
//   bucket: n (+ N) bytes of symbol ranks grouped by context rank;  S.rowlen[r] = first byte of bucket r (the
//   start of bucket nsym closes the last one);  the lane starts (:325-327) are appended to the bucket of symbol 0.
__device__ inline void cta_pair_buckets(const uint8_t *in, uint32_t n, int N, uint32_t nsym, uint8_t *bucket,
                                        PrepSmem &S) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t seg = n / N;
    // bucket sizes in rank space: occurrences of the symbol except as the last byte, the virtual 0 in front of the
    // first byte, and the N-1 lane starts for symbol 0
    uint32_t sz = 0;
    if ((uint32_t)tid < nsym) {
        const uint32_t s = S.sym[tid];
        sz = S.T[s] - (s == in[n - 1] ? 1u : 0u) + (s == 0 ? (uint32_t)N : 0u);
    }
    uint32_t incl = warp_incl_scan(sz, lane);
    if (lane == 31) S.wtot[wid] = incl;
    __syncthreads();
    uint32_t before = 0;
#pragma unroll
    for (int w = 0; w < PREP_WARPS; w++) if (w < wid) before += S.wtot[w];
    const uint32_t start = before + incl - sz;
    uint32_t *cur = S.Hs;                       // cursors, then (rowlen) the bucket starts
    cur[tid] = start;
    S.rowlen[tid] = start;
    __syncthreads();
    auto deal = [&](uint32_t rp, uint32_t rc) { bucket[atomicAdd(&cur[rp], 1u)] = (uint8_t)rc; };
    const uint8_t *rank = S.rank;
    uint32_t head = (uint32_t)((16 - ((uintptr_t)in & 15)) & 15);
    if (head > n) head = n;
    if ((uint32_t)tid < head) deal(rank[tid ? in[tid - 1] : 0], rank[in[tid]]);
    const uint8_t *p = in + head;
    const uint32_t rest = n - head, nv = rest >> 4;
    const uint4 *v = (const uint4 *)p;
    for (uint32_t i = tid; i < nv; i += PREP_THREADS) {
        const uint4 q = v[i];
        const uint32_t pb = (i || head) ? p[16 * (size_t)i - 1] : 0;
        const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
        uint32_t rp = rank[pb];
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const uint32_t rc = rank[(w4[a] >> (8 * b)) & 0xff];
                deal(rp, rc);
                rp = rc;
            }
    }
    for (uint32_t t = (nv << 4) + tid; t < rest; t += PREP_THREADS) {
        const uint32_t pos = head + t;
        deal(rank[pos ? in[pos - 1] : 0], rank[in[pos]]);
    }
    if (tid >= 1 && tid < N) deal(rank[0], rank[in[(size_t)tid * seg]]);      // lanes 1..N-1 start in context 0
    __threadfence_block();
    __syncthreads();
}

// the counts of context row i, lane l getting columns 8l..8l+7: histogram of bucket i in the warp's bins
__device__ __forceinline__ void row_from_bucket(const uint8_t *bucket, uint32_t b0, uint32_t b1, uint32_t *bins,
                                                uint32_t (&f)[8], int lane) {
    uint4 *z = (uint4 *)bins;
    z[lane] = make_uint4(0, 0, 0, 0); z[lane + 32] = make_uint4(0, 0, 0, 0);
    __syncwarp();
    uint32_t k = b0 + lane;
    for (; k + 96 < b1; k += 128) {                 // four byte loads in flight per lane
        const uint32_t s0 = bucket[k], s1 = bucket[k + 32], s2 = bucket[k + 64], s3 = bucket[k + 96];
        atomicAdd(&bins[s0], 1u); atomicAdd(&bins[s1], 1u); atomicAdd(&bins[s2], 1u); atomicAdd(&bins[s3], 1u);
    }
    for (; k < b1; k += 32) atomicAdd(&bins[bucket[k]], 1u);
    __syncwarp();
    const uint4 a = z[2 * lane], b = z[2 * lane + 1];
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    __syncwarp();
}

// load / store the 8 columns a lane owns of one row
__device__ __forceinline__ void row_load8(const uint32_t *row, uint32_t nsym, uint32_t j0, uint32_t (&f)[8]) {
    if ((nsym & 3) == 0 && j0 + 8 <= nsym) {
        const uint4 a = *(const uint4 *)(row + j0), b = *(const uint4 *)(row + j0 + 4);
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    } else {
#pragma unroll
        for (int t = 0; t < 8; t++) f[t] = j0 + t < nsym ? row[j0 + t] : 0;
    }
}
__device__ __forceinline__ void row_store8(uint32_t *row, uint32_t nsym, uint32_t j0, const uint32_t (&f)[8]) {
    if ((nsym & 3) == 0 && j0 + 8 <= nsym) {
        *(uint4 *)(row + j0) = make_uint4(f[0], f[1], f[2], f[3]);
        *(uint4 *)(row + j0 + 4) = make_uint4(f[4], f[5], f[6], f[7]);
    } else {
#pragma unroll
        for (int t = 0; t < 8; t++) if (j0 + t < nsym) row[j0 + t] = f[t];
    }
}

// Bytes of one serialised row (rANS_static16_int.h:278-306: a varint per non-zero count, 0x00,(z-1) per run of z
// zeros) and, with out != null, the bytes themselves at out + off and the row's encoder symbols.  f = the row's
// normalised counts, lane l holding columns 8l..8l+7.  Returns the row's length (uniform).
__device__ __forceinline__ uint32_t row_emit(const uint32_t (&f)[8], uint32_t nsym, uint32_t mv, uint32_t shift,
                                             uint8_t *out, uint32_t off, uint32_t *srow, int lane) {
    const uint32_t j0 = (uint32_t)lane * 8;
    // zero bitmap of the row: zb[u] covers columns 32u..32u+31 (columns >= nsym read as non-zero)
    uint32_t m8 = 0;
#pragma unroll
    for (int t = 0; t < 8; t++) if (j0 + t < nsym && !f[t]) m8 |= 1u << t;
    uint32_t wz = m8 << (8 * (lane & 3));
    wz |= __shfl_xor_sync(FULL, wz, 1);
    wz |= __shfl_xor_sync(FULL, wz, 2);
    uint32_t zb[8];
#pragma unroll
    for (int u = 0; u < 8; u++) zb[u] = __shfl_sync(FULL, wz, 4 * u);
    uint32_t nz[9];                          // first non-zero column at or after the start of word u (256 if none)
    nz[8] = 256;
#pragma unroll
    for (int u = 7; u >= 0; u--) nz[u] = ~zb[u] ? 32 * u + __ffs(~zb[u]) - 1 : nz[u + 1];
    int sh = 0;
    while ((mv << sh) < (1u << shift)) sh++;
    const uint32_t myw = lane >> 2;
    uint32_t zw = 0, nzn = 256, zprev = 0;
#pragma unroll
    for (int u = 0; u < 8; u++) if (myw == (uint32_t)u) { zw = zb[u]; nzn = nz[u + 1]; }
#pragma unroll
    for (int u = 1; u < 8; u++) if (myw == (uint32_t)u) zprev = zb[u - 1] >> 31;
    uint32_t len[8], run[8], tot8 = 0;
#pragma unroll
    for (int t = 0; t < 8; t++) {
        const uint32_t j = j0 + t, bit = j & 31;
        uint32_t l = 0, r = 0;
        if (j < nsym) {
            if (f[t]) l = f[t] >= 128 ? 2 : 1;
            else {
                const uint32_t pz = bit ? (zw >> (bit - 1)) & 1 : zprev;
                if (!pz) {                   // first zero of a run: one token for the whole run
                    const uint32_t w = (~zw) >> bit;
                    const uint32_t end = w ? j + __ffs(w) - 1 : nzn;
                    l = 2;
                    r = min(end, nsym) - j;
                }
            }
        }
        len[t] = l; run[t] = r;
        tot8 += l | ((f[t] << sh) << 16);
    }
    uint32_t ex = warp_incl_scan(tot8, lane);
    const uint32_t rowbytes = __shfl_sync(FULL, ex, 31) & 0xffff;
    if (!out) return rowbytes;
    ex -= tot8;
    uint32_t o = off + (ex & 0xffff), x = ex >> 16;
    uint32_t e8[8];
#pragma unroll
    for (int t = 0; t < 8; t++) {
        const uint32_t fs = f[t] << sh;
        if (len[t]) {
            if (f[t]) {
                if (len[t] == 2) { out[o] = (uint8_t)(0x80 | (f[t] >> 7)); out[o + 1] = (uint8_t)(f[t] & 0x7f); }
                else out[o] = (uint8_t)f[t];
            } else { out[o] = 0; out[o + 1] = (uint8_t)(run[t] - 1); }
            o += len[t];
        }
        e8[t] = enc_sym_make4(x, fs, shift);
        x += fs;
    }
    row_store8(srow, nsym, j0, e8);
    return rowbytes;
}

// The whole order-1 model of `in`: counts in S.T on entry (symbol space, cta_hist8).  Leaves the uncompressed
// table (first byte = shift << 4) at tbl, the encoder symbols at symtab, the rank map in P.  N = lanes of the coder.
__device__ inline void cta_o1_model(const uint8_t *in, uint32_t n, int N, uint8_t *bucket, uint8_t *rowstage,
                                    uint32_t *symtab, uint2 *E, uint8_t *tbl, Prep &P, PrepSmem &S) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    PREP_T0();
    // ---- alphabet = symbols present, plus 0 (:357-361)
    const bool pres = S.T[tid] != 0 || tid == 0;
    {
        const uint32_t bal = __ballot_sync(FULL, pres);
        if ((tid & 31) == 0) S.pres[tid >> 5] = bal;
    }
    uint32_t r;
    const uint32_t nsym = cta_rank256(pres, S.wtot, &r);
    S.rank[tid] = pres ? (uint8_t)r : 0xff;
    if (pres) S.sym[r] = (uint8_t)tid;
    P.rank[tid] = pres ? (uint8_t)r : 0xff;
    __syncthreads();
    // ---- pairs dealt into their contexts' buckets (the lane starts of :325-327 included)
    const bool stream_syms = nsym > PREP_SYM_MAX;       // encoder symbols per position instead of a table
    cta_pair_buckets(in, n, N, nsym, bucket, S);
    PREP_T(4);
    S.bstart[tid] = S.rowlen[tid];
    if (tid == 0) S.bstart[256] = n + (uint32_t)N - 1;
    __syncthreads();
    const uint32_t last_rank = S.rank[in[n - 1]];
    const uint32_t j0 = (uint32_t)lane * 8;
    uint32_t *bins = S.Fw[wid];
    // ---- sweep 1: row totals (the last symbol's gets one extra, utils.h:311,345) and the statistics of
    // rans_compute_shift (rANS_static4x16pr.c:357-420); warp w takes rows w, w+8, ...
    double e10 = 0, e12 = 0;
    uint32_t max_tot = 0;
    for (uint32_t i = wid; i < nsym; i += PREP_WARPS) {
        uint32_t f[8];
        const uint32_t b0 = S.bstart[i], b1 = S.bstart[i + 1];
        const uint32_t Ti = b1 - b0 + (i == last_rank ? 1u : 0u);
        if (lane == 0) S.T[i] = Ti;
        if (!Ti) { if (lane == 0) S.S16[i] = 0; continue; }
        row_from_bucket(bucket, b0, b1, bins, f, lane);
        uint32_t max_val = round2(Ti);
        uint32_t cnt = 0;                           // ns | sm10 << 10 | sm12 << 20
#pragma unroll
        for (int t = 0; t < 8; t++) {
            if (!f[t]) continue;
            cnt += 1;
            if ((uint64_t)f[t] * 1025 <= max_val) cnt += 1u << 10;     // max_val / f > 1024
            if ((uint64_t)f[t] * 4097 <= max_val) cnt += 1u << 20;     // max_val / f > 4096
        }
        cnt = warp_sum(cnt);
        const uint32_t ns = cnt & 1023, sm10 = (cnt >> 10) & 1023, sm12 = cnt >> 20;
        const double l10 = log((double)(1024 + sm10)), l12 = log((double)(4096 + sm12));
        const double T_slow = (double)4096 / Ti, T_fast = (double)1024 / Ti;
#pragma unroll
        for (int t = 0; t < 8; t++) {
            if (!f[t]) continue;
            const double a = f[t] * T_fast, b = f[t] * T_slow;
            e10 -= f[t] * (fast_log(a > 1 ? a : 1) - l10);
            e12 -= f[t] * (fast_log(b > 1 ? b : 1) - l12);
            e10 += 1.3;
            e12 += 4.7;
        }
        if (ns < 64 && max_val > 128) max_val /= 2;
        if (max_val > 1024) max_val /= 2;
        if (max_val > 4096) max_val = 4096;
        if (lane == 0) S.S16[i] = (uint16_t)max_val;
        if (max_tot < max_val) max_tot = max_val;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        e10 += __shfl_xor_sync(FULL, e10, o);
        e12 += __shfl_xor_sync(FULL, e12, o);
    }
    if (lane == 0) { S.red[0][wid] = e10; S.red[1][wid] = e12; S.redu[wid] = max_tot; }
    __syncthreads();
    PREP_T(5);
    e10 = 0; e12 = 0; max_tot = 0;
#pragma unroll
    for (int w = 0; w < PREP_WARPS; w++) { e10 += S.red[0][w]; e12 += S.red[1][w]; max_tot = max(max_tot, S.redu[w]); }
    const uint32_t shift = (e10 / e12 < 1.01 || max_tot <= 1024) ? 10 : 12;
    __syncthreads();
    // ---- sweep 2: normalise_freq of every row (rANS_static16_int.h:97-146) to its stored total, serialise it
    // (encode_freq_d, :278-306) into its staging slot and build its encoder symbols (rANS_word.h:201-272)
    int err = 0;
    for (uint32_t i = wid; i < nsym; i += PREP_WARPS) {
        const uint32_t Ti = S.T[i];
        if (!Ti) { if (lane == 0) S.rowlen[i] = 0; continue; }
        uint32_t f[8];
        row_from_bucket(bucket, S.bstart[i], S.bstart[i + 1], bins, f, lane);
        uint32_t mv = S.S16[i];
        if (shift == 10 && mv > 1024) mv = 1024;
        uint32_t size = Ti;
        for (int pass = 0; pass < 2; pass++) {
            const uint64_t tr = ((uint64_t)mv << 31) / size + (uint32_t)((1 << 30) / (int)size);
            uint32_t top = 0, arg = 0, sum = 0;
#pragma unroll
            for (int t = 0; t < 8; t++) {
                uint32_t v = f[t];
                if (!v) continue;
                if (top < v) { top = v; arg = j0 + t; }
                v = (uint32_t)((v * tr) >> 31);
                if (!v) v = 1;
                f[t] = v;
                sum += v;
            }
            sum = warp_sum(sum);
#pragma unroll
            for (int o = 16; o; o >>= 1) {      // max, ties -> lowest index
                const uint32_t t2 = __shfl_xor_sync(FULL, top, o), a2 = __shfl_xor_sync(FULL, arg, o);
                if (t2 > top || (t2 == top && a2 < arg)) { top = t2; arg = a2; }
            }
            const uint32_t big = top ? arg : 0;
            uint32_t mine = 0;
#pragma unroll
            for (int t = 0; t < 8; t++) if ((big & 7) == (uint32_t)t) mine = f[t];
            const uint32_t fb = __shfl_sync(FULL, mine, big >> 3);
            int adjust = (int)mv - (int)sum;
            const bool own = (big >> 3) == (uint32_t)lane;
            if (adjust >= 0 || (fb > (uint32_t)-adjust && (pass == 1 || fb / 2 >= (uint32_t)-adjust))) {
                if (own) {
#pragma unroll
                    for (int t = 0; t < 8; t++) if ((big & 7) == (uint32_t)t) f[t] += adjust;
                }
                break;
            }
            if (pass == 0) { size = sum; continue; }
            // greedy shave (rare): serial over the row through the warp's bins
#pragma unroll
            for (int t = 0; t < 8; t++) bins[j0 + t] = f[t];
            __syncwarp();
            if (lane == 0) {
                adjust += (int)fb - 1;
                bins[big] = 1;
                for (uint32_t j = 0; adjust && j < nsym; j++) {
                    if (bins[j] < 2) continue;
                    const int d = (bins[j] > (uint32_t)-adjust) ? adjust : 1 - (int)bins[j];
                    bins[j] += d;
                    adjust -= d;
                }
                if (!bins[big]) err = 1;
            }
            __syncwarp();
#pragma unroll
            for (int t = 0; t < 8; t++) f[t] = bins[j0 + t];
            __syncwarp();
        }
        const uint32_t rb = row_emit(f, nsym, mv, shift, rowstage + (size_t)i * PREP_ROW_STRIDE, 0,
                                     symtab + (size_t)i * nsym, lane);
        if (lane == 0) S.rowlen[i] = rb;
    }
    err = __any_sync(FULL, err);
    if (err && lane == 0) P.err = 1;
    PREP_T(6);
    // ---- the alphabet of the contexts, 0 forced in (:357-361), then the rows' offsets
    if (tid == 0) {
        uint8_t *cp = tbl;
        *cp++ = (uint8_t)(shift << 4);
        auto present = [&](int s) { return (S.pres[s >> 5] >> (s & 31)) & 1; };
        int j = 0;
        while (j < 256) {
            if (!present(j)) { j++; continue; }
            *cp++ = (uint8_t)j;
            if (j && present(j - 1)) {
                int k = j + 1;
                while (k < 256 && present(k)) k++;
                *cp++ = (uint8_t)(k - (j + 1));
                j = k;
            } else j++;
        }
        *cp++ = 0;
        S.bc[0] = (uint32_t)(cp - tbl);
    }
    __threadfence_block();
    __syncthreads();
    uint32_t mylen = (uint32_t)tid < nsym ? S.rowlen[tid] : 0;      // thread t owns row t
    {
        const uint32_t incl = warp_incl_scan(mylen, lane);
        if (lane == 31) S.wtot[wid] = incl;
        __syncthreads();
        uint32_t before = S.bc[0];
#pragma unroll
        for (int w = 0; w < PREP_WARPS; w++) if (w < wid) before += S.wtot[w];
        S.bstart[tid] = before + incl - mylen;          // (the bucket starts are dead: row offsets in the table)
        if (tid == PREP_THREADS - 1) S.bc[1] = before + incl;
    }
    __syncthreads();
    for (uint32_t i = wid; i < nsym; i += PREP_WARPS)
        warp_copy(tbl + S.bstart[i], rowstage + (size_t)i * PREP_ROW_STRIDE, S.rowlen[i], lane);
    __threadfence_block();
    __syncthreads();
    PREP_T(7);
    if (stream_syms) {
        // E[p] = symbol of in[p] in the context of in[p-1] (E[0]: context 0); E[n + z] = first symbol of lane z in
        // context 0: the packed 4-byte symbol and, beside it, the reciprocal of its frequency, so that a chain step
        // needs no table at all.  16 positions per thread and trip: one 16-byte load of the data, 16 look-ups
        // (+ 16 in the 16 KiB reciprocal table, L1-resident here), eight 16-byte stores.
        auto with_rcp = [](uint32_t c) { return make_uint2(c, rcp_of_freq((c >> 13) & 0x1fff)); };
        const uint8_t *rank = S.rank;
        const uint32_t r0 = rank[0], seg = n / N;
        uint32_t done = 0;
        if ((((uintptr_t)in) & 15) == 0) {
            const uint4 *v = (const uint4 *)in;
            const uint32_t nv = n >> 4;
            for (uint32_t i = tid; i < nv; i += PREP_THREADS) {
                const uint4 q = v[i];
                const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
                uint32_t rp = i ? rank[in[16 * i - 1]] : r0, e[16];
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    const uint32_t rc = rank[(w4[k >> 2] >> (8 * (k & 3))) & 0xff];
                    e[k] = symtab[rp * nsym + rc];
                    rp = rc;
                }
                uint4 *d = (uint4 *)(E + 16 * (size_t)i);
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const uint2 a = with_rcp(e[2 * k]), b2 = with_rcp(e[2 * k + 1]);
                    d[k] = make_uint4(a.x, a.y, b2.x, b2.y);
                }
            }
            done = nv << 4;
        }
        for (uint32_t p2 = done + tid; p2 < n; p2 += PREP_THREADS)
            E[p2] = with_rcp(symtab[(p2 ? rank[in[p2 - 1]] : r0) * nsym + rank[in[p2]]]);
        if (tid >= 1 && tid < N) E[n + tid] = with_rcp(symtab[r0 * nsym + rank[in[(size_t)tid * seg]]]);
        __syncthreads();
    }
    PREP_T(8);
    if (tid == 0) { P.nsym = nsym; P.shift = shift; P.tl = S.bc[1]; P.stream_syms = stream_syms ? 1u : 0u; }
}

// ------------------------------------------------------------------ the coder warp's side
// Order-1 stream whose model the CTA built: the table goes to `out` raw or, when longer than 1000 bytes and it
// pays, through the 4-lane order-0 coder (rANS_static16_int.h:396-412); then the state chains.
// Shared memory of a warp in the chains-only kernel: [ring ORING][rank 256][encoder symbols when they fit]; the
// order-0 coder of the table (and of the run-length meta-data before it) uses the same bytes from the start.
struct __align__(16) EncPrepSmem {
    uint8_t ring[ORING];
    uint8_t rank[256];
};
// The order-1 state chains over a stream of per-position encoder symbols: E[p] codes in[p] in the context of
// in[p-1] (E[0]: context 0), E[n + z] codes the first symbol of lane z >= 1 in context 0 (rANS_static32x16pr.c:
// 457-525).  An entry is the packed 4-byte symbol (enc_sym_make4) and the reciprocal of its frequency.  Lane z owns
// [z*seg, (z+1)*seg), lane N-1 also the tail; everything runs backwards.
__device__ __forceinline__ EncSym enc_sym_unpack2(uint2 c, uint32_t bits) {
    const uint32_t f = (c.x >> 13) & 0x1fff;
    EncSym s;
    s.xlim = f << (31 - bits);
    s.rcp = c.y;
    s.bias = c.x & 0x1fff;
    s.cmpl = (1u << bits) - f;
    s.shw = c.x >> 26;
    return s;
}
template <int N>
__device__ __forceinline__ void enc_o1_payload_stream(const uint2 *E, uint32_t n, uint8_t *out, uint8_t *out_end,
                                                      uint8_t **ptr_out, uint8_t *ring, uint32_t shift, int lane) {
    const uint32_t seg = n / N;
    const bool act = lane < N;
    OutRing w;
    w.init(out, out_end, ring);
    uint32_t R = RANS_L;
    {   // tail on lane N-1, from the end down to N*seg
        const bool lastl = lane == N - 1;
        for (uint32_t p = n - 1; p >= N * seg && p > 0; p--) {
            EncSym e = make_uint4(0, 0, 0, 0);
            if (lastl) e = enc_sym_unpack2(E[p], shift);
            w.maybe_flush(lane);
            R = enc_step(R, lastl, e, w, lane);
        }
    }
    const uint2 *q = E + (size_t)(act ? lane : 0) * seg;
    uint32_t k = seg;                                        // positions q[1 .. k) are still to be coded
    __syncwarp();
    if (seg >= 8) {
        // Groups of four entries counted from the END of the lane's segment, one group ahead of the chain in registers
        // (a segment has any length -- the payload behind RLE -- so the groups are 8-byte aligned only); the seg % 4
        // entries in front of the first group go through the loop below.  Every lane walks its own segment: the
        // lines further down are asked into L2 sixteen groups (four lines) ahead.
        const uint32_t lead = seg & 3;
        const uint2 *qq = q + lead;                          // qq[4 g .. 4 g + 3] = group g
        auto ld4 = [&](uint32_t g, uint2 (&c)[4]) {
            const uint2 *p = qq + 4 * (size_t)g;
            c[0] = p[0]; c[1] = p[1]; c[2] = p[2]; c[3] = p[3];
        };
        uint32_t j = seg >> 2;
        uint2 c[4];
        ld4(j - 1, c);
        for (uint32_t a = 2; a <= 16 && a < j; a++)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(__cvta_generic_to_global(qq + 4 * (size_t)(j - a))));
        while (j > 1) {
            if (j > 17) asm volatile("prefetch.global.L2 [%0];" ::"l"(__cvta_generic_to_global(qq + 4 * (size_t)(j - 18))));
            uint2 nx[4];
            ld4(j - 2, nx);
            w.maybe_flush(lane);
            R = enc_step<N == 32>(R, act, enc_sym_unpack2(c[3], shift), w, lane);
            R = enc_step<N == 32>(R, act, enc_sym_unpack2(c[2], shift), w, lane);
            R = enc_step<N == 32>(R, act, enc_sym_unpack2(c[1], shift), w, lane);
            R = enc_step<N == 32>(R, act, enc_sym_unpack2(c[0], shift), w, lane);
            c[0] = nx[0]; c[1] = nx[1]; c[2] = nx[2]; c[3] = nx[3];
            j--;
        }
        w.maybe_flush(lane);                                 // group 0
        R = enc_step<N == 32>(R, act, enc_sym_unpack2(c[3], shift), w, lane);
        R = enc_step<N == 32>(R, act, enc_sym_unpack2(c[2], shift), w, lane);
        R = enc_step<N == 32>(R, act, enc_sym_unpack2(c[1], shift), w, lane);
        if (lead) R = enc_step<N == 32>(R, act, enc_sym_unpack2(c[0], shift), w, lane);    // else it is the lane's first symbol
        k = lead ? lead : 1;
    }
    for (; k > 1; k--) {
        EncSym e = make_uint4(0, 0, 0, 0);
        if (act) e = enc_sym_unpack2(q[k - 1], shift);
        w.maybe_flush(lane);
        R = enc_step<N == 32>(R, act, e, w, lane);
    }
    if (seg) {                                               // every lane's first symbol: context 0
        EncSym e = make_uint4(0, 0, 0, 0);
        if (act) e = enc_sym_unpack2(lane ? E[n + lane] : E[0], shift);
        w.maybe_flush(lane);
        R = enc_step<N == 32>(R, act, e, w, lane);
    }
    enc_flush(R, act, N, w, lane);
    *ptr_out = w.slot + w.off;
    __syncwarp();
}

template <int N>
__device__ int enc_o1_prepped(const uint8_t *in, uint32_t n, uint8_t *out, uint8_t *out_end, uint32_t *tab_len,
                              uint8_t **ptr_out, uint8_t *smem, uint32_t smem_bytes, const Prep &P,
                              const uint8_t *prep_base, uint32_t prep_isz, int lane) {
    *tab_len = 0;
    *ptr_out = out_end;
    if (P.err) return 1;
    const PrepPlan pl = prep_plan(prep_isz);
    const uint32_t *gsym = (const uint32_t *)(prep_base + pl.o_sym);
    const uint8_t *tbl = prep_base + pl.o_tbl;
    uint8_t *tmp = const_cast<uint8_t *>(prep_base) + pl.o_tmp;
    const uint32_t nsym = P.nsym, shift = P.shift;
    uint32_t tl = P.tl;
    bool raw = true;
    if (tl > 1000) {
        const uint32_t usz = tl - 1;
        const uint32_t cb = (compress_bound(usz, 0) - 20) & ~1u;
        uint32_t ctab = 0;
        uint8_t *cptr = nullptr;
        if (enc_o0<4>(tbl + 1, usz, tmp, tmp + cb, &ctab, &cptr, *(EncO0Smem *)smem, lane) == 0) {
            const uint32_t pay = (uint32_t)(tmp + cb - cptr), csz = ctab + pay;
            if (csz + 6 < tl) {
                uint32_t h = 1;
                if (lane == 0) {
                    out[0] = tbl[0] | 1;
                    h += var_put_u32(out + h, usz);
                    h += var_put_u32(out + h, csz);
                }
                h = __shfl_sync(FULL, h, 0);
                __syncwarp();
                warp_copy(out + h, tmp, ctab, lane);
                warp_copy(out + h + ctab, cptr, pay, lane);
                tl = h + csz;
                raw = false;
            }
        }
        __syncwarp();
    }
    if (raw) warp_copy(out, tbl, tl, lane);
    *tab_len = tl;
    // the chains: rank map and, when they fit, the encoder symbols in shared memory (the order-0 scratch is dead)
    EncPrepSmem &S = *(EncPrepSmem *)smem;
    ((uint2 *)S.rank)[lane] = ((const uint2 *)P.rank)[lane];
    if (P.stream_syms) {
        __syncwarp();
        enc_o1_payload_stream<N>((const uint2 *)(prep_base + pl.o_E), n, out, out_end, ptr_out, S.ring, shift, lane);
        return 0;
    }
    const uint32_t hw = nsym * nsym;
    const bool sym_smem = sizeof(EncPrepSmem) + hw * 4 <= smem_bytes;
    const uint32_t *symtab = gsym;
    if (sym_smem) {
        uint32_t *d = (uint32_t *)(smem + sizeof(EncPrepSmem));
        for (uint32_t j = lane; j < hw; j += 32) d[j] = gsym[j];
        symtab = d;
    }
    __syncwarp();
    enc_o1_payload<N>(in, n, out, out_end, ptr_out, S.ring, S.rank, symtab, nsym, shift, sym_smem, lane);
    return 0;
}

// The CTA's work for one stream: mirrors the decisions of enc_stream (kernels.cu) up to the coder.
__device__ inline void prep_stream(EncJob &J, PrepSmem &S) {
    Prep *Pp = (Prep *)J.prep;
    const int tid = threadIdx.x;
    int order = J.order;
    uint32_t in_size = J.in_size;
    const uint8_t *in = J.in;
    if ((order & ORDER_SIMD_AUTO) && in_size >= 50000 && !(order & X_STRIPE)) order |= X_32;
    if (in_size <= 20) order &= ~X_STRIPE;
    if (in_size <= 1000) order &= ~X_32;
    const bool skip = in_size > 0x7fffffffu || (order & (X_STRIPE | X_CAT)) || !in_size || !J.work ||
                      !(order & (X_PACK | X_RLE));
    if (skip) { if (tid == 0) Pp->state = 0; return; }
    Prep &P = *Pp;
    if (tid == 0) P.err = 0;
    const int do_pack = order & X_PACK, no_size = order & X_NOSZ, do_rle = order & X_RLE;
    int do_simd = order & X_32, o1 = order & 1;
    const uint32_t meta = 1 + (no_size ? 0 : var_size_u32(in_size));
    uint8_t *work = J.work;
    uint32_t packed = 0, pmeta = 0, plen = 0, rle_len = 0, rmeta_len = 0;
    PREP_T0();
    if (do_pack) {                                                        // rANS_static4x16pr.c:1429-1459
        packed = cta_pack(in, in_size, J.slot + meta, &pmeta, work, &plen, S) ? 1u : 0u;
        if (packed) {
            in = work; work += (plen + 15) & ~15u;
            in_size = plen;
            if (do_simd && in_size < 32) do_simd = 0;
        }
    }
    PREP_T(0);
    if (do_rle && in_size) {                                              // :1464-1533
        uint8_t *lits = work, *rmeta = work + ((in_size + 15) & ~15u);
        cta_rle_encode(in, in_size, lits, &rle_len, rmeta, &rmeta_len, S);
        if (!((double)((uint64_t)rle_len + rmeta_len) >= .99 * (double)in_size)) {
            if (do_simd && (rmeta_len < 32 || rle_len < 32)) do_simd = 0;
            in = lits; in_size = rle_len;
        }
    }
    PREP_T(1);
    if (o1 && in_size < 8) o1 = 0;                                        // :1547
    uint32_t model = 0;
    if (in_size) {
        cta_hist8(in, in_size, S);
        PREP_T(2);
        P.F[tid] = S.T[tid];
        model = 1;
        const int N = do_simd ? 32 : 4;
        if (o1 && !(N == 32 && in_size < 32)) {
            const PrepPlan pl = prep_plan(J.in_size);
            uint8_t *base = (uint8_t *)Pp;
            cta_o1_model(in, in_size, N, base + pl.o_bkt, base + pl.o_rows, (uint32_t *)(base + pl.o_sym),
                         (uint2 *)(base + pl.o_E), base + pl.o_tbl, P, S);
            model = 2;
        }
    }
    __syncthreads();
    if (tid == 0) {
        P.packed = packed; P.pmeta = pmeta; P.plen = plen; P.rle_len = rle_len; P.rmeta_len = rmeta_len;
        P.model = model;
        __threadfence();
        P.state = 1;
    }
}

}  // namespace b200
