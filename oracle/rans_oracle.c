/*
 * rans_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, single-threaded restatement of the htscodecs rANS Nx16 codec as
 * used by fqzcomp5 (order-0 / order-1, N = 4 or 32 interleaved lanes, with the
 * PACK / RLE / NOSZ / CAT / STRIPE container).  It exists only so that tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg can check the CUDA
 * path; nothing in fqzcomp5_b200/ links, loads or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle.py compares every entry point,
 * byte for byte, with the unmodified reference compiled by oracle/Makefile into
 * oracle/_ref/libref_rans.so (the reference's tests hold no golden vectors for
 * this path), and with the vectors that library produced, which
 * are committed under tests/golden/.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference/htscodecs).  The arithmetic is written from the format
 * description in its textbook form -- e.g. the encoder
 * uses x/f and x%f where the reference uses a fixed-point reciprocal
 * (rANS_word.h:101 documents the equivalence) -- so agreement with the
 * reference library is a real check and not a tautology.
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include <math.h>

#define ORC_API __attribute__((visibility("default")))

enum { X_PACK = 0x80, X_RLE = 0x40, X_CAT = 0x20, X_NOSZ = 0x10, X_STRIPE = 0x08, X_32 = 0x04 };
#define ORDER_STRIPE_NO0 (1 << 16)
#define ORDER_SIMD_AUTO  (1 << 17)
#define RANS_L   (1u << 15)      /* rANS_word.h:64 */
#define O0_BITS  12              /* rANS_static16_int.h:61 */

/* ------------------------------------------------------------------ varints
 * 7 bits per byte, most significant group first, 0x80 = more (varint.h:205-299) */
ORC_API int orc_var_put_u32(uint8_t *p, uint32_t v) {
    int n = 1, k;
    while (n < 5 && (v >> (7 * n))) n++;
    for (k = n - 1; k >= 0; k--)
        *p++ = (uint8_t)(((v >> (7 * k)) & 0x7f) | (k ? 0x80 : 0));
    return n;
}

ORC_API int orc_var_get_u32(const uint8_t *p, const uint8_t *end, uint32_t *v) {
    const uint8_t *s = p;
    uint32_t x = 0;
    int more = 1, cnt = 0;
    if (p >= end) { *v = 0; return 0; }
    while (more && p < end && cnt < 6) {
        uint8_t c = *p++;
        x = (x << 7) | (c & 0x7f);
        more = c & 0x80;
        cnt++;
    }
    *v = x;
    return (int)(p - s);
}

/* ------------------------------------------------------------ histograms
 * utils.h:145-244 (hist8/hist8e): plain byte counts. */
ORC_API void orc_hist8(const uint8_t *in, uint32_t n, uint32_t F[256]) {
    uint32_t i;
    memset(F, 0, 256 * sizeof(*F));
    for (i = 0; i < n; i++) F[in[i]]++;
}

/* utils.h:279-357 (hist1_4): F[prev][cur], prev of the first byte is 0;
 * T[i] = row sum, and the LAST symbol's total gets one extra count. */
ORC_API void orc_hist1(const uint8_t *in, uint32_t n, uint32_t (*F)[256], uint32_t T[256]) {
    uint32_t i, j;
    uint8_t prev = 0;
    for (i = 0; i < n; i++) { F[prev][in[i]]++; prev = in[i]; }
    T[prev]++;
    for (i = 0; i < 256; i++) {
        uint32_t t = 0;
        for (j = 0; j < 256; j++) t += F[i][j];
        T[i] += t;
    }
}

/* rANS_static16_int.h:86-95 */
static uint32_t round2(uint32_t v) {
    uint32_t p = 1;
    if (v == 0) return 0;              /* (0-1 | ...)+1 wraps to 0 in the reference */
    if (v > 0x80000000u) return 0;
    while (p < v) p <<= 1;
    return p;
}

/* rANS_static16_int.h:97-146.  Scale counts so that they sum to tot.
 * 31-bit fixed point multiplier, zero counts stay zero, non-zero counts stay
 * >= 1, the most frequent symbol absorbs the rounding error; if that would
 * more than halve it, rescale once more from the already scaled counts, and
 * if it still does not fit, shave the other symbols greedily. */
ORC_API int orc_normalise_freq(uint32_t F[256], int size, uint32_t tot) {
    int pass, j, big = 0;
    if (!size) return 0;
    for (pass = 0; pass < 2; pass++) {
        uint64_t tr = ((uint64_t)tot << 31) / size + (1 << 30) / size;
        uint32_t top = 0;
        int adjust;
        size = 0; big = 0;
        for (j = 0; j < 256; j++) {
            if (!F[j]) continue;
            if (top < F[j]) { top = F[j]; big = j; }
            F[j] = (uint32_t)((F[j] * tr) >> 31);
            if (F[j] == 0) F[j] = 1;
            size += F[j];
        }
        adjust = (int)tot - size;
        if (adjust >= 0) { F[big] += adjust; break; }
        if (F[big] > (uint32_t)-adjust && (pass == 1 || F[big] / 2 >= (uint32_t)-adjust)) {
            F[big] += adjust;
            break;
        }
        if (pass == 0) continue;
        adjust += F[big] - 1;
        F[big] = 1;
        for (j = 0; adjust && j < 256; j++) {
            int d;
            if (F[j] < 2) continue;
            d = (F[j] > (uint32_t)-adjust) ? adjust : 1 - (int)F[j];
            F[j] += d;
            adjust -= d;
        }
    }
    return F[big] > 0 ? 0 : -1;
}

/* rANS_static16_int.h:151-162: counts already sum to a power of two; shift up. */
static void normalise_freq_shift(uint32_t F[256], uint32_t size, uint32_t max_tot) {
    int sh = 0, i;
    if (size == 0 || size == max_tot) return;
    while (size < max_tot) { size *= 2; sh++; }
    for (i = 0; i < 256; i++) F[i] <<= sh;
}

/* --------------------------------------------------------- table (de)serialisers
 * rANS_static16_int.h:165-189.  Symbol list; a symbol that directly follows a
 * listed one is followed by a count of further consecutive symbols, which are
 * then omitted.  0 terminates. */
static int put_alphabet(uint8_t *cp, const uint32_t F[256]) {
    uint8_t *op = cp;
    int j = 0;
    while (j < 256) {
        if (!F[j]) { j++; continue; }
        *cp++ = (uint8_t)j;
        if (j && F[j - 1]) {
            int k = j + 1;
            while (k < 256 && F[k]) k++;
            *cp++ = (uint8_t)(k - (j + 1));
            j = k;
        } else {
            j++;
        }
    }
    *cp++ = 0;
    return (int)(cp - op);
}

/* rANS_static16_int.h:191-238.  The loop body runs before the "j != 0" test, so
 * a leading 0 byte lists symbol 0 (order-1 tables always start that way). */
static int get_alphabet(const uint8_t *cp, const uint8_t *end, uint32_t F[256]) {
    const uint8_t *op = cp;
    int run = 0, j;
    if (cp >= end) return 0;
    j = *cp++;
    do {
        F[j] = 1;
        if (cp >= end) return 0;
        if (!run && j + 1 == *cp) {
            if (cp + 1 >= end) return 0;
            j = *cp++;
            run = *cp++;
        } else if (run) {
            run--;
            if (++j > 255) return 0;
        } else {
            j = *cp++;
        }
    } while (j && cp < end);
    return (int)(cp - op);
}

/* rANS_static16_int.h:240-252 */
static int put_freq0(uint8_t *cp, const uint32_t F[256]) {
    uint8_t *op = cp;
    int j;
    cp += put_alphabet(cp, F);
    for (j = 0; j < 256; j++)
        if (F[j]) cp += orc_var_put_u32(cp, F[j]);
    return (int)(cp - op);
}

/* rANS_static16_int.h:254-272 */
static int get_freq0(const uint8_t *cp, const uint8_t *end, uint32_t F[256], uint32_t *tot) {
    const uint8_t *op = cp;
    int j, n;
    uint32_t t = 0;
    n = get_alphabet(cp, end, F);
    if (!n) return 0;
    cp += n;
    for (j = 0; j < 256; j++) {
        if (!F[j]) continue;
        cp += orc_var_get_u32(cp, end, &F[j]);
        t += F[j];
    }
    *tot = t;
    return (int)(cp - op);
}

/* rANS_static16_int.h:278-306: one order-1 row against the order-0 alphabet A;
 * a run of z zero counts is written as 0,(z-1). */
static int put_freq_row(uint8_t *cp, const uint32_t A[256], const uint32_t F[256]) {
    uint8_t *op = cp;
    int j = 0;
    while (j < 256) {
        if (!A[j]) { j++; continue; }
        if (F[j]) { cp += orc_var_put_u32(cp, F[j]); j++; continue; }
        {   /* count the zero run over listed symbols */
            int z = 0;
            while (j < 256 && (!A[j] || !F[j])) { if (A[j]) z++; j++; }
            *cp++ = 0;
            *cp++ = (uint8_t)(z - 1);
        }
    }
    return (int)(cp - op);
}

/* rANS_static16_int.h:425-456 */
static int get_freq_row(const uint8_t *cp, const uint8_t *end, const uint32_t A[256],
                        uint32_t F[256], uint32_t *tot) {
    const uint8_t *op = cp;
    int j, zrun = 0;
    uint32_t t = 0;
    if (cp >= end) return 0;
    for (j = 0; j < 256 && cp < end; j++) {
        uint32_t f;
        if (!A[j]) continue;
        if (zrun) { f = 0; zrun--; }
        else {
            cp += orc_var_get_u32(cp, end, &f);
            if (f == 0) {
                if (cp >= end) return 0;
                zrun = *cp++;
            }
        }
        F[j] = f;
        t += f;
    }
    *tot = t;
    return (int)(cp - op);
}

/* ------------------------------------------------------------- rANS primitives
 * One symbol, textbook form (rANS_word.h:78-102, 224, 287-336): emit the low
 * 16 bits (little endian, growing DOWNWARD) when x exceeds the bound, then
 * x = (x/f << bits) + x%f + start. */
static inline uint32_t enc_put(uint32_t x, uint8_t **pp, uint32_t start, uint32_t f, int bits) {
    uint32_t x_max = ((RANS_L >> bits) << 16) * f - 1;
    if (x > x_max) {
        uint8_t *p = *pp - 2;
        p[0] = x & 0xff; p[1] = (x >> 8) & 0xff;
        *pp = p;
        x >>= 16;
    }
    return ((x / f) << bits) + (x % f) + start;
}
static inline void enc_flush(uint32_t x, uint8_t **pp) {            /* rANS_word.h:105-117 */
    uint8_t *p = *pp - 4;
    p[0] = x; p[1] = x >> 8; p[2] = x >> 16; p[3] = x >> 24;
    *pp = p;
}
/* rANS_word.h:468-476 and the in-range fast form :414-462 collapse to this:
 * refill only when two more bytes exist. */
static inline uint32_t dec_renorm(uint32_t x, const uint8_t **pp, const uint8_t *end) {
    if (x < RANS_L && *pp + 2 <= end) {
        x = (x << 16) | (*pp)[0] | ((uint32_t)(*pp)[1] << 8);
        *pp += 2;
    }
    return x;
}

/* rANS_static4x16pr.c:93-106 */
ORC_API unsigned int orc_rans_compress_bound_4x16(unsigned int size, int order) {
    int N = (order >> 8) & 0xff;
    unsigned int sz;
    if (!N) N = 4;
    order &= 0xff;
    sz = (order == 0
          ? 1.05 * size + 257 * 3 + 4
          : 1.05 * size + 257 * 257 * 3 + 4 + 257 * 3 + 4) +
         ((order & X_PACK) ? 1 : 0) +
         ((order & X_RLE) ? 1 + 257 * 3 + 4 : 0) + 20 +
         ((order & X_32) ? (32 - 4) * 4 : 0) +
         ((order & X_STRIPE) ? 7 + 5 * N : 0);
    return sz + (sz & 1) + 2;
}

/* ----------------------------------------------------------- order-0 encode
 * rANS_static4x16pr.c:112-232 (N=4) and rANS_static32x16pr.c:67-254 (N=32).
 * Symbol i belongs to lane i%N.  Encoding walks the input backwards; within
 * one group of N symbols lane N-1 goes first.  Output = table, N states
 * (lane 0 first), then the 16-bit words in decoder order. */
static uint8_t *enc_o0(const uint8_t *in, uint32_t n, uint8_t *out, uint32_t *out_size, int N) {
    uint32_t F[256], start[256], R[32];
    uint32_t bound = orc_rans_compress_bound_4x16(n, 0) - 20, fsum, x;
    uint8_t *ptr, *end, *cp = out;
    int j, tab = 0;
    int64_t i;

    if (!out || bound > *out_size) return NULL;
    if ((uintptr_t)out & 1) bound--;
    ptr = end = out + bound;
    if (n) {
        orc_hist8(in, n, F);
        fsum = round2(n);
        if (fsum > (1u << O0_BITS)) fsum = 1u << O0_BITS;
        if (orc_normalise_freq(F, n, fsum) < 0) return NULL;
        cp += put_freq0(cp, F);
        tab = (int)(cp - out);
        if (orc_normalise_freq(F, fsum, 1u << O0_BITS) < 0) return NULL;
        for (x = 0, j = 0; j < 256; j++) { start[j] = x; x += F[j]; }
        for (j = 0; j < N; j++) R[j] = RANS_L;
        for (i = (int64_t)n - 1; i >= 0; i--) {
            int z = (int)(i % N);
            R[z] = enc_put(R[z], &ptr, start[in[i]], F[in[i]], O0_BITS);
        }
        for (j = N - 1; j >= 0; j--) enc_flush(R[j], &ptr);
    }
    *out_size = (uint32_t)(end - ptr) + tab;
    memmove(out + tab, ptr, end - ptr);
    return out;
}

/* ----------------------------------------------------------- order-0 decode
 * rANS_static4x16pr.c:234-349, rANS_static32x16pr.c:256-410. */
static uint8_t *dec_o0(const uint8_t *in, uint32_t in_size, uint8_t *out, uint32_t out_sz, int N) {
    uint32_t F[256] = {0}, fsum, R[32], x;
    uint16_t base[256];
    uint8_t  slot2sym[1 << O0_BITS];
    const uint8_t *cp = in, *end = in + in_size;
    int j, n;
    uint32_t i;

    if (in_size < 16 || out_sz >= INT_MAX) return NULL;
    /* the 4-lane decoder parses its table against end-8 (:249,:263) */
    n = get_freq0(cp, N == 4 ? end - 8 : end, F, &fsum);
    if (!n) return NULL;
    cp += n;
    normalise_freq_shift(F, fsum, 1u << O0_BITS);
    for (x = 0, j = 0; j < 256; j++) {
        if (!F[j]) continue;
        if (F[j] > (1u << O0_BITS) - x) return NULL;
        base[j] = (uint16_t)x;
        memset(slot2sym + x, j, F[j]);
        x += F[j];
    }
    if (x != (1u << O0_BITS)) return NULL;
    if (end - cp < N * 4) return NULL;
    for (j = 0; j < N; j++) {
        R[j] = cp[0] | (cp[1] << 8) | (cp[2] << 16) | ((uint32_t)cp[3] << 24);
        cp += 4;
        if (R[j] < RANS_L) return NULL;
    }
    for (i = 0; i < out_sz; i++) {
        int z = (int)(i % N);
        uint32_t m = R[z] & ((1u << O0_BITS) - 1);
        uint8_t s = slot2sym[m];
        out[i] = s;
        /* the 32-lane decoder leaves the last n%32 symbols as pure look-ups
         * (rANS_static32x16pr.c:400-401); no difference in output */
        R[z] = F[s] * (R[z] >> O0_BITS) + m - base[s];
        R[z] = dec_renorm(R[z], &cp, end);
    }
    return out;
}

/* rANS_static4x16pr.c:357-420.  Choose 10- or 12-bit order-1 precision from an
 * entropy estimate; S[i] = power-of-two total each context row is stored at.
 * Floating point, same expression order as the reference. */
static double fast_log(double a) {                                 /* utils.h:69-72 */
    union { double d; long long x; } u;
    u.d = a;
    return (u.x - 4606921278410026770LL) * 1.539095918623324e-16;
}
ORC_API int orc_rans_compute_shift(const uint32_t *F0, uint32_t (*F)[256],
                                   const uint32_t *T, uint32_t *S) {
    double e10 = 0, e12 = 0;
    int i, j, max_tot = 0;
    for (i = 0; i < 256; i++) {
        unsigned int max_val;
        int ns = 0, sm10 = 0, sm12 = 0;
        double l10, l12, T_slow, T_fast;
        if (F0[i] == 0) continue;
        max_val = round2(T[i]);
        for (j = 0; j < 256; j++) {
            if (F[i][j] && max_val / F[i][j] > 1024) sm10++;
            if (F[i][j] && max_val / F[i][j] > 4096) sm12++;
        }
        l10 = log(1024 + sm10);
        l12 = log(4096 + sm12);
        T_slow = (double)4096 / T[i];
        T_fast = (double)1024 / T[i];
        for (j = 0; j < 256; j++) {
            if (!F[i][j]) continue;
            ns++;
            {
                double a = F[i][j] * T_fast, b = F[i][j] * T_slow;
                e10 -= F[i][j] * (fast_log(a > 1 ? a : 1) - l10);
                e12 -= F[i][j] * (fast_log(b > 1 ? b : 1) - l12);
            }
            e10 += 1.3;
            e12 += 4.7;
        }
        if (ns < 64 && max_val > 128) max_val /= 2;
        if (max_val > 1024) max_val /= 2;
        if (max_val > 4096) max_val = 4096;
        S[i] = max_val;
        if (max_tot < (int)max_val) max_tot = max_val;
    }
    return (e10 / e12 < 1.01 || max_tot <= 1024) ? 10 : 12;
}

/* ------------------------------------------------------ order-1 model + table
 * rANS_static16_int.h:312-421 (encode_freq1).  Builds freq/start tables at the
 * chosen precision and writes the (possibly self-compressed) table. */
typedef struct { uint16_t f[256][256]; uint16_t s[256][256]; } o1_model;

static int enc_o1_model(const uint8_t *in, uint32_t n, int N, o1_model *M, uint8_t **cpp) {
    uint32_t (*F)[256] = calloc(256, sizeof(*F));
    uint32_t T[256] = {0}, S[256] = {0}, A[256];
    uint8_t *out = *cpp, *cp = out;
    int i, j, z, shift, seg = n / N;
    if (!F) return -1;
    orc_hist1(in, n, F, T);
    for (z = 1; z < N; z++) F[0][in[z * seg]]++;        /* lanes 1..N-1 start in context 0 */
    T[0] += N - 1;

    memcpy(A, T, sizeof(A));
    A[0] = 1;                                            /* symbol 0 always listed (:357-361) */
    *cp++ = 0;
    cp += put_alphabet(cp, A);

    shift = orc_rans_compute_shift(T, F, T, S);
    for (i = 0; i < 256; i++) {
        uint32_t mv, x;
        if (T[i] == 0) continue;
        mv = S[i];
        if (shift == 10 && mv > 1024) mv = 1024;
        if (orc_normalise_freq(F[i], T[i], mv) < 0) { free(F); return -1; }
        T[i] = mv;
        cp += put_freq_row(cp, T, F[i]);
        normalise_freq_shift(F[i], mv, 1u << shift);
        T[i] = 1u << shift;
        for (x = 0, j = 0; j < 256; j++) {
            M->f[i][j] = (uint16_t)F[i][j];
            M->s[i][j] = (uint16_t)x;
            x += F[i][j];
        }
    }
    *out = (uint8_t)(shift << 4);
    if (cp - out > 1000) {                               /* try the 4-lane o0 coder on the table */
        uint32_t usz = (uint32_t)(cp - (out + 1));
        uint32_t csz = orc_rans_compress_bound_4x16(usz, 0) - 20;
        uint8_t *c = malloc(csz);
        if (c && enc_o0(out + 1, usz, c, &csz, 4) && csz + 6 < (uint32_t)(cp - out)) {
            uint8_t *op = out;
            *op++ |= 1;
            op += orc_var_put_u32(op, usz);
            op += orc_var_put_u32(op, csz);
            memcpy(op, c, csz);
            cp = op + csz;
        }
        free(c);
    }
    *cpp = cp;
    free(F);
    return shift;
}

/* ----------------------------------------------------------- order-1 encode
 * rANS_static4x16pr.c:422-518 (N=4), rANS_static32x16pr.c:414-525 (N=32).
 * Lane z owns [z*seg,(z+1)*seg), lane N-1 also the tail.  Each symbol is coded
 * in the context of its predecessor; the first symbol of each lane in context
 * 0.  Backwards: first the tail on lane N-1, then N lanes in lock step (lane
 * N-1 first), then the N first symbols. */
static uint8_t *enc_o1(const uint8_t *in, uint32_t n, uint8_t *out, uint32_t *out_size, int N) {
    uint32_t bound = orc_rans_compress_bound_4x16(n, 1) - 20, R[32];
    uint8_t *cp = out, *ptr, *end;
    o1_model *M;
    int shift, z, tab;
    int64_t seg = n / N, k, i;

    if (N == 32 && n < 32) return NULL;
    if (!out || bound > *out_size) return NULL;
    if ((uintptr_t)out & 1) bound--;
    end = ptr = out + bound;
    if (!(M = malloc(sizeof(*M)))) return NULL;
    shift = enc_o1_model(in, n, N, M, &cp);
    if (shift < 0) { free(M); return NULL; }
    tab = (int)(cp - out);
    for (z = 0; z < N; z++) R[z] = RANS_L;

    for (i = (int64_t)n - 1; i >= N * seg; i--) {        /* tail, lane N-1 */
        uint8_t ctx = in[i - 1], s = in[i];
        if (i == 0) break;
        R[N - 1] = enc_put(R[N - 1], &ptr, M->s[ctx][s], M->f[ctx][s], shift);
    }
    for (k = seg - 1; k >= 1; k--)
        for (z = N - 1; z >= 0; z--) {
            uint8_t ctx = in[z * seg + k - 1], s = in[z * seg + k];
            R[z] = enc_put(R[z], &ptr, M->s[ctx][s], M->f[ctx][s], shift);
        }
    for (z = N - 1; z >= 0; z--) {
        uint8_t s = in[z * seg];
        R[z] = enc_put(R[z], &ptr, M->s[0][s], M->f[0][s], shift);
    }
    for (z = N - 1; z >= 0; z--) enc_flush(R[z], &ptr);
    *out_size = (uint32_t)(end - ptr) + tab;
    memmove(out + tab, ptr, end - ptr);
    free(M);
    return out;
}

/* ----------------------------------------------------------- order-1 decode
 * rANS_static4x16pr.c:524-821, rANS_static32x16pr.c:531-758,
 * table: rANS_static16_int.h:468-536. */
static uint8_t *dec_o1(const uint8_t *in, uint32_t in_size, uint8_t *out, uint32_t out_sz, int N) {
    const uint8_t *cp = in, *end = in + in_size, *tend, *after = NULL;
    uint8_t *ctab = NULL, *ok = NULL;
    uint16_t (*f)[256] = NULL, (*b)[256] = NULL;
    uint8_t *slot = NULL;                       /* [256][1<<shift] */
    uint32_t A[256] = {0}, R[32], pos[32], last[32] = {0};
    uint32_t shift, seg, step;
    int i, j, z, n;

    if (in_size < (uint32_t)(N == 4 ? 16 : N * 4) || out_sz >= INT_MAX) return NULL;
    shift = *cp >> 4;
    if (shift != 10 && shift != 12) return NULL;   /* only values the encoder writes */
    tend = end;
    if (*cp++ & 1) {
        uint32_t usz, csz;
        cp += orc_var_get_u32(cp, end, &usz);
        cp += orc_var_get_u32(cp, end, &csz);
        if (csz > (uint32_t)(end - cp)) return NULL;
        after = cp + csz;
        if (!(ctab = malloc(usz ? usz : 1))) return NULL;
        if (!dec_o0(cp, csz, ctab, usz, 4)) goto err;
        cp = ctab; tend = ctab + usz;
    }
    f = calloc(256, sizeof(*f)); b = calloc(256, sizeof(*b));
    slot = calloc(256, 1u << shift);
    if (!f || !b || !slot) goto err;
    n = get_alphabet(cp, tend, A);
    if (!n) goto err;
    cp += n;
    if (cp >= tend) goto err;
    for (i = 0; i < 256; i++) {
        uint32_t Fr[256] = {0}, T = 0, x = 0;
        if (!A[i]) continue;
        n = get_freq_row(cp, tend, A, Fr, &T);
        if (!n) goto err;
        cp += n;
        if (!T) continue;
        normalise_freq_shift(Fr, T, 1u << shift);
        for (j = 0; j < 256; j++) {
            if (!Fr[j]) continue;
            if (Fr[j] > (1u << shift) - x) goto err;
            memset(slot + ((size_t)i << shift) + x, j, Fr[j]);
            f[i][j] = (uint16_t)Fr[j]; b[i][j] = (uint16_t)x;
            x += Fr[j];
        }
        if (x != (1u << shift)) goto err;
    }
    if (after) cp = after;
    free(ctab); ctab = NULL;
    if (end - cp < N * 4) goto err;
    for (z = 0; z < N; z++) {
        R[z] = cp[0] | (cp[1] << 8) | (cp[2] << 16) | ((uint32_t)cp[3] << 24);
        cp += 4;
        if (R[z] < RANS_L) goto err;
    }
    seg = out_sz / N;
    for (z = 0; z < N; z++) pos[z] = z * seg;
    for (step = 0; step < seg; step++)
        for (z = 0; z < N; z++) {
            uint32_t m = R[z] & ((1u << shift) - 1);
            uint8_t s = slot[((size_t)last[z] << shift) + m];
            out[pos[z]++] = s;
            R[z] = f[last[z]][s] * (R[z] >> shift) + m - b[last[z]][s];
            R[z] = dec_renorm(R[z], &cp, end);
            last[z] = s;
        }
    z = N - 1;
    while (pos[z] < out_sz) {
        uint32_t m = R[z] & ((1u << shift) - 1);
        uint8_t s = slot[((size_t)last[z] << shift) + m];
        out[pos[z]++] = s;
        R[z] = f[last[z]][s] * (R[z] >> shift) + m - b[last[z]][s];
        R[z] = dec_renorm(R[z], &cp, end);
        last[z] = s;
    }
    ok = out;
err:
    free(ctab); free(f); free(b); free(slot);
    return ok;
}

/* --------------------------------------------------------------------- PACK
 * pack.c:56-147.  meta = [nsym][symbols ascending]; codes are ranks; first
 * symbol in the low bits; 8/4/2 codes per byte for nsym <=2/<=4/<=16. */
ORC_API uint8_t *orc_pack(const uint8_t *in, int64_t len, uint8_t *meta, int *meta_len, uint64_t *out_len) {
    int code[256] = {0}, n = 0, i, per, bits;
    uint8_t *out;
    uint64_t j = 0;
    int64_t k;
    for (k = 0; k < len; k++) code[in[k]] = 1;
    for (i = 0; i < 256; i++)
        if (code[i]) { code[i] = n++; meta[n] = (uint8_t)i; }
    meta[0] = (uint8_t)n;
    if (n > 16) return NULL;
    if (!(out = malloc(len + 1))) return NULL;
    *meta_len = n + 1;
    per = n > 4 ? 2 : n > 2 ? 4 : n > 1 ? 8 : 0;
    if (per) {
        bits = 8 / per;
        for (k = 0; k < len; k += per) {
            int v = 0, q;
            for (q = 0; q < per && k + q < len; q++) v |= code[in[k + q]] << (q * bits);
            out[j++] = (uint8_t)v;
        }
    }
    *out_len = j;
    return out;
}

/* pack.c:161-194 */
static int unpack_meta(const uint8_t *d, uint32_t dlen, uint8_t *map, int *per) {
    unsigned n, c;
    if (dlen == 0) return 0;
    n = d[0] ? d[0] : 256;
    if (n <= 1) *per = 0; else if (n <= 2) *per = 8; else if (n <= 4) *per = 4;
    else if (n <= 16) *per = 2; else { *per = 1; return 1; }
    if (dlen <= 1) return 0;
    for (c = 0; c < n && 1 + c < dlen; c++) map[c] = d[1 + c];
    return c < n ? 0 : (int)(1 + c);
}

/* pack.c:207-344 */
static uint8_t *unpack(const uint8_t *d, int64_t len, uint8_t *out, uint64_t out_len, int per, const uint8_t *map) {
    uint64_t i;
    int bits;
    if (per == 1) { memcpy(out, d, len); return out; }
    if (per == 0) { memset(out, map[0], out_len); return out; }
    if (per != 2 && per != 4 && per != 8) return NULL;
    if ((int64_t)((out_len + per - 1) / per) > len) return NULL;
    bits = 8 / per;
    for (i = 0; i < out_len; i++)
        out[i] = map[(d[i / per] >> ((i % per) * bits)) & ((1 << bits) - 1)];
    return out;
}

/* ---------------------------------------------------------------------- RLE
 * rle.c:48-98: a symbol is run-length coded iff it follows itself more often
 * than not.  rle.c:100-138: literals keep one byte per run of such symbols,
 * the run length minus one goes to a varint stream. */
ORC_API uint8_t *orc_rle_encode(const uint8_t *in, uint64_t len, uint8_t *run, uint64_t *run_len,
                                uint8_t *syms, int *nsyms, uint64_t *out_len) {
    int64_t score[256] = {0};
    uint8_t *out = malloc(len * 2 + 1);
    uint64_t i, j = 0, k = 0;
    int last = -1, n = 0;
    if (!out) return NULL;
    for (i = 0; i < len; i++) { score[in[i]] += (in[i] == last) ? 1 : -1; last = in[i]; }
    for (i = 0; i < 256; i++) if (score[i] > 0) syms[n++] = (uint8_t)i;
    *nsyms = n;
    for (i = 0; i < len; ) {
        uint8_t s = in[i];
        out[k++] = s;
        if (score[s] > 0) {
            uint64_t e = i + 1;
            while (e < len && in[e] == s) e++;
            j += orc_var_put_u32(run + j, (uint32_t)(e - i - 1));
            i = e;
        } else {
            i++;
        }
    }
    *run_len = j;
    *out_len = k;
    return out;
}

/* rle.c:142-189 */
static uint8_t *rle_decode(const uint8_t *lit, uint64_t lit_len, const uint8_t *run, uint64_t run_len,
                           const uint8_t *syms, int nsyms, uint8_t *out, uint64_t *out_len) {
    uint8_t is_rle[256] = {0};
    const uint8_t *run_end = run + run_len, *lit_end = lit + lit_len;
    uint8_t *o = out, *o_end = out + *out_len;
    int j;
    for (j = 0; j < nsyms; j++) is_rle[syms[j]] = 1;
    for (; lit < lit_end; lit++) {
        uint8_t s = *lit;
        if (o >= o_end) return NULL;
        if (is_rle[s]) {
            uint32_t r;
            run += orc_var_get_u32(run, run_end, &r);
            if (r) {
                if (o + r >= o_end) return NULL;
                memset(o, s, (size_t)r + 1);
                o += (size_t)r + 1;
                continue;
            }
        }
        *o++ = s;
    }
    *out_len = o - out;
    return out;
}

/* --------------------------------------------------------------- container
 * rANS_static4x16pr.c:1224-1600. */
static uint8_t *enc_payload(int N, int o1, const uint8_t *in, uint32_t n, uint8_t *out, uint32_t *out_size) {
    return o1 ? enc_o1(in, n, out, out_size, N) : enc_o0(in, n, out, out_size, N);
}
static uint8_t *dec_payload(int N, int o1, const uint8_t *in, uint32_t n, uint8_t *out, uint32_t out_sz) {
    return o1 ? dec_o1(in, n, out, out_sz, N) : dec_o0(in, n, out, out_sz, N);
}

ORC_API unsigned char *orc_rans_compress_to_4x16(unsigned char *in, unsigned int in_size,
                                                 unsigned char *out, unsigned int *out_size, int order) {
    uint8_t *out_free = NULL, *out_end, *packed = NULL, *rle = NULL;
    unsigned int meta_len;
    int do_pack, do_rle, no_size, do_simd;

    if (in_size > INT_MAX || (out && *out_size == 0)) { *out_size = 0; return NULL; }
    if (!out) {
        *out_size = orc_rans_compress_bound_4x16(in_size, order);
        if (!(out_free = out = malloc(*out_size))) { *out_size = 0; return NULL; }
    }
    out_end = out + *out_size;

    if ((order & ORDER_SIMD_AUTO) && in_size >= 50000 && !(order & X_STRIPE)) order |= X_32;   /* :1256 */
    if (in_size <= 20) order &= ~X_STRIPE;                                                     /* :1260 */
    if (in_size <= 1000) order &= ~X_32;                                                       /* :1263 */

    if (order & X_STRIPE) {                                                                    /* :1266-1393 */
        static const int methods[4] = {1, 64, 128, 0};
        unsigned int N = (order >> 8) & 0xff, part_len[256], idx[256], i, j;
        uint8_t *tr, *out2, *out2_start, *best = NULL;
        if (N == 0) N = 4;
        if (N > in_size) N = in_size;
        if (!(tr = malloc(in_size))) goto fail;
        for (i = 0; i < N; i++) {
            part_len[i] = in_size / N + ((in_size % N) > i);
            idx[i] = i ? idx[i - 1] + part_len[i - 1] : 0;
        }
        for (i = 0; i < in_size; i++) tr[idx[i % N] + i / N] = in[i];

        out[0] = (uint8_t)(order & ~X_NOSZ);
        meta_len = 1 + orc_var_put_u32(out + 1, in_size);
        if (meta_len >= *out_size) { free(tr); goto fail; }
        out[meta_len++] = (uint8_t)N;
        out2_start = out2 = out + 7 + 5 * N;
        if (!(best = malloc(orc_rans_compress_bound_4x16(part_len[0], order & 0xff) + 64))) { free(tr); goto fail; }
        for (i = 0; i < N; i++) {
            unsigned int best_sz = INT_MAX, olen2 = 0;
            for (j = 0; j < 4; j++) {
                if ((order & methods[j]) != methods[j]) continue;
                if ((order & ORDER_STRIPE_NO0) && !(methods[j] & 1)) continue;
                if (out2 - out > (ptrdiff_t)*out_size) continue;
                olen2 = *out_size - (unsigned)(out2 - out);
                if (orc_rans_compress_to_4x16(tr + idx[i], part_len[i], out2, &olen2,
                                              methods[j] | X_NOSZ | (order & X_32))
                    && olen2 && best_sz > olen2) {
                    best_sz = olen2;
                    memcpy(best, out2, olen2);
                }
            }
            if (best_sz == INT_MAX) { free(best); free(tr); goto fail; }
            memcpy(out2, best, best_sz);
            out2 += best_sz;
            meta_len += orc_var_put_u32(out + meta_len, best_sz);
        }
        free(best);
        memmove(out + meta_len, out2_start, out2 - out2_start);
        free(tr);
        *out_size = meta_len + (unsigned)(out2 - out2_start);
        return out;
    }

    if (order & X_CAT) {                                                                       /* :1395-1409 */
        out[0] = X_CAT;
        meta_len = 1 + orc_var_put_u32(out + 1, in_size);
        if (meta_len + in_size > *out_size) goto fail;
        if (in_size) memcpy(out + meta_len, in, in_size);
        *out_size = meta_len + in_size;
        return out;
    }

    do_pack = order & X_PACK; do_rle = order & X_RLE;
    no_size = order & X_NOSZ; do_simd = order & X_32;
    out[0] = (uint8_t)order;
    meta_len = 1;
    if (!no_size) meta_len += orc_var_put_u32(out + 1, in_size);
    order &= 3;

    if (do_pack && in_size) {                                                                  /* :1429-1459 */
        int pm; uint64_t plen;
        if (meta_len + 256 > *out_size) goto fail;
        packed = orc_pack(in, in_size, out + meta_len, &pm, &plen);
        if (!packed) { out[0] &= ~X_PACK; do_pack = 0; }
        else {
            int sz;
            in = packed; in_size = (unsigned)plen; meta_len += pm;
            sz = orc_var_put_u32(out + meta_len, in_size);
            meta_len += sz; *out_size -= sz;
            if (do_simd && in_size < 32) { do_simd = 0; out[0] &= ~X_32; }
        }
    } else if (do_pack) out[0] &= ~X_PACK;

    if (do_rle && in_size) {                                                                   /* :1464-1533 */
        uint8_t *meta = malloc(in_size + 257 + 8), syms[256];
        unsigned int rmeta_len, c_rmeta_len;
        uint64_t rle_len = 0, runs_len = 0;
        int nsyms = 0;
        if (!meta) goto fail;
        rle = orc_rle_encode(in, in_size, meta + 257, &runs_len, syms, &nsyms, &rle_len);
        if (rle) {
            memmove(meta + 1 + nsyms, meta + 257, runs_len);
            meta[0] = (uint8_t)nsyms;
            memcpy(meta + 1, syms, nsyms);
        }
        rmeta_len = (unsigned)runs_len + nsyms + 1;
        if (!rle || rle_len + rmeta_len >= .99 * in_size) {
            out[0] &= ~X_RLE; do_rle = 0; free(rle); rle = NULL;
        } else {
            int sz = orc_var_put_u32(out + meta_len, rmeta_len * 2), sz2;
            sz += orc_var_put_u32(out + meta_len + sz, (uint32_t)rle_len);
            if (meta_len + sz + 5 > *out_size) { free(meta); goto fail; }
            c_rmeta_len = *out_size - (meta_len + sz + 5);
            if (do_simd && (rmeta_len < 32 || rle_len < 32)) { do_simd = 0; out[0] &= ~X_32; }
            if (!enc_o0(meta, rmeta_len, out + meta_len + sz + 5, &c_rmeta_len, do_simd ? 32 : 4)) { free(meta); goto fail; }
            if (c_rmeta_len < rmeta_len) {
                sz2 = orc_var_put_u32(out + meta_len + sz, c_rmeta_len);
                memmove(out + meta_len + sz + sz2, out + meta_len + sz + 5, c_rmeta_len);
            } else {
                sz = orc_var_put_u32(out + meta_len, rmeta_len * 2 + 1);
                sz2 = orc_var_put_u32(out + meta_len + sz, (uint32_t)rle_len);
                memcpy(out + meta_len + sz + sz2, meta, rmeta_len);
                c_rmeta_len = rmeta_len;
            }
            meta_len += sz + sz2 + c_rmeta_len;
            in = rle; in_size = (unsigned)rle_len;
        }
        free(meta);
    } else if (do_rle) out[0] &= ~X_RLE;

    if (meta_len > *out_size) goto fail;
    *out_size -= meta_len;
    if (order && in_size < 8) { out[0] &= ~1; order &= ~1; }                                   /* :1547 */
    if (!enc_payload(do_simd ? 32 : 4, order & 1, in, in_size, out + meta_len, out_size)) goto fail;
    if (*out_size >= in_size) {                                                                /* :1560-1574 */
        out[0] &= ~3;
        out[0] |= X_CAT | no_size;
        if (out + meta_len + in_size > out_end) goto fail;
        if (in_size) memcpy(out + meta_len, in, in_size);
        *out_size = in_size;
    }
    free(rle); free(packed);
    *out_size += meta_len;
    return out;
fail:
    free(out_free); free(rle); free(packed);
    *out_size = 0;
    return NULL;
}

ORC_API unsigned char *orc_rans_compress_4x16(unsigned char *in, unsigned int in_size,
                                              unsigned int *out_size, int order) {
    return orc_rans_compress_to_4x16(in, in_size, NULL, out_size, order);
}

/* rANS_static4x16pr.c:1607-1894 */
ORC_API unsigned char *orc_rans_uncompress_to_4x16(unsigned char *in, unsigned int in_size,
                                                   unsigned char *out, unsigned int *out_size) {
    const uint8_t *in_end = in + in_size;
    uint8_t *out_free = NULL, *tmp = NULL, *meta_free = NULL, *t1, *t2, *t3, *ret = NULL;
    const uint8_t *meta = NULL;
    uint8_t map[16] = {0};
    unsigned int osz, t1_size, u_meta = 0;
    int order, do_pack, do_rle, do_cat, no_size, do_simd, sz, per = 0;
    uint64_t unpacked_sz = 0;

    if (in_size == 0) return NULL;
    if (*in & X_STRIPE) {                                                                      /* :1615-1694 */
        unsigned int ulen, meta_len = 1, N, i, clen[256], ulenN[256], idx[256];
        uint64_t ctot = 0;
        uint8_t *outN;
        meta_len += orc_var_get_u32(in + meta_len, in_end, &ulen);
        if (meta_len >= in_size) return NULL;
        N = in[meta_len++];
        if (N < 1) return NULL;
        if (!out) {
            if (ulen >= INT_MAX) return NULL;
            if (!(out_free = out = malloc(ulen ? ulen : 1))) return NULL;
            *out_size = ulen;
        }
        if (ulen != *out_size) { free(out_free); return NULL; }
        for (i = 0; i < N; i++) {
            ulenN[i] = ulen / N + ((ulen % N) > i);
            idx[i] = i ? idx[i - 1] + ulenN[i - 1] : 0;
            meta_len += orc_var_get_u32(in + meta_len, in_end, &clen[i]);
            ctot += clen[i];
            if (meta_len > in_size || clen[i] > in_size || clen[i] < 1) { free(out_free); return NULL; }
        }
        if (meta_len + ctot > in_size) { free(out_free); return NULL; }
        in_size = (unsigned)(meta_len + ctot);
        if (!(outN = malloc(ulen ? ulen : 1))) { free(out_free); return NULL; }
        for (i = 0; i < N; i++) {
            unsigned int olen = ulenN[i];
            if (in_size < meta_len ||
                !orc_rans_uncompress_to_4x16(in + meta_len, in_size - meta_len, outN + idx[i], &olen) ||
                olen != ulenN[i]) { free(out_free); free(outN); return NULL; }
            meta_len += clen[i];
        }
        for (i = 0; i < ulen; i++) out[i] = outN[idx[i % N] + i / N];   /* utils.h:79-138 */
        free(outN);
        *out_size = ulen;
        return out;
    }

    order = *in++; in_size--;
    do_pack = order & X_PACK; do_rle = order & X_RLE; do_cat = order & X_CAT;
    no_size = order & X_NOSZ; do_simd = order & X_32;
    order &= 1;
    if (!no_size) sz = orc_var_get_u32(in, in_end, &osz);
    else { sz = 0; osz = *out_size; }
    in += sz; in_size -= sz;
    if (no_size && !out) return NULL;
    if (!out) {
        *out_size = osz;
        if (!(out = out_free = malloc(osz ? osz : 1))) return NULL;
    } else {
        if (*out_size < osz) return NULL;
        *out_size = osz;
    }
    t1_size = *out_size;
    if (do_pack || do_rle) {                                                                   /* :1760-1782 */
        if (!(tmp = malloc(*out_size ? *out_size : 1))) goto err;
        if (do_pack && do_rle) { t1 = out; t2 = tmp; t3 = out; }
        else if (do_pack)      { t1 = tmp; t2 = tmp; t3 = out; }
        else                   { t1 = tmp; t2 = out; t3 = out; }
    } else t1 = t2 = t3 = out;

    if (do_pack) {                                                                             /* :1788-1806 */
        unsigned int psz;
        int c = unpack_meta(in, in_size, map, &per);
        if (c == 0) goto err;
        unpacked_sz = osz;
        in += c; in_size -= c;
        sz = orc_var_get_u32(in, in_end, &psz);
        in += sz; in_size -= sz;
        if (psz > t1_size) goto err;
        t1_size = psz;
    }
    if (do_rle) {                                                                              /* :1810-1834 */
        uint32_t c_meta, rle_len, s2;
        s2 = orc_var_get_u32(in, in_end, &u_meta);
        s2 += orc_var_get_u32(in + s2, in_end, &rle_len);
        if (rle_len > t1_size) goto err;
        if (u_meta & 1) {
            meta = in + s2;
            u_meta = u_meta / 2 > (uint32_t)(in_end - meta) ? (uint32_t)(in_end - meta) : u_meta / 2;
            c_meta = u_meta;
        } else {
            s2 += orc_var_get_u32(in + s2, in_end, &c_meta);
            u_meta /= 2;
            if (!(meta_free = malloc(u_meta ? u_meta : 1))) goto err;
            if (in_size < s2 || !dec_o0(in + s2, in_size - s2, meta_free, u_meta, do_simd ? 32 : 4)) goto err;
            meta = meta_free;
        }
        if (c_meta + s2 > in_size) goto err;
        in += c_meta + s2; in_size -= c_meta + s2;
        t1_size = rle_len;
    }
    if (in_size) {                                                                             /* :1838-1853 */
        if (do_cat) {
            if (t1_size > in_size || t1_size > *out_size) goto err;
            memcpy(t1, in, t1_size);
        } else if (!dec_payload(do_simd ? 32 : 4, order, in, in_size, t1, t1_size)) goto err;
    } else t1_size = 0;

    {
        uint64_t t2_size = t1_size, t3_size = t1_size;
        if (do_rle) {                                                                          /* :1856-1871 */
            uint64_t unrle = *out_size;
            int nsyms;
            if (u_meta == 0) goto err;
            nsyms = *meta ? *meta : 256;
            if (u_meta < (uint32_t)(1 + nsyms)) goto err;
            if (!rle_decode(t1, t1_size, meta + 1 + nsyms, u_meta - (1 + nsyms), meta + 1, nsyms, t2, &unrle)) goto err;
            t3_size = t2_size = unrle;
        }
        if (do_pack) {                                                                         /* :1872-1881 */
            if (per == 1) unpacked_sz = t2_size;     /* ">16 symbols": bytes were stored unpacked */
            if (!unpack(t2, t2_size, t3, unpacked_sz, per, map)) goto err;
            t3_size = unpacked_sz;
        }
        *out_size = (unsigned)t3_size;
    }
    ret = t3;
err:
    free(meta_free); free(tmp);
    if (!ret) free(out_free);
    return ret;
}

ORC_API unsigned char *orc_rans_uncompress_4x16(unsigned char *in, unsigned int in_size, unsigned int *out_size) {
    return orc_rans_uncompress_to_4x16(in, in_size, NULL, out_size);
}
