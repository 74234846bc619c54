/*
 * bz3_oracle.c -- CPU restatement of the bzip3 per-block hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under bzip3_b200/ may include, link or
 * call this file; it exists so that tests/, __graft_entry__.smoke() and the
 * cpu_baseline leg of bench.py can check the CUDA path against an independent
 * statement of the algorithm.  Parity is PINNED: tests/test_oracle.py checks
 * every function here against the reference itself compiled from
 * /root/reference (oracle/_ref/libbz3_ref.so, see oracle/Makefile) and against
 * the reference's golden vector examples/shakespeare.txt.bz3
 * (copied as a data fixture to tests/golden/).
 *
 * Each function cites the reference lines whose behaviour it restates
 * (paths relative to /root/reference).  The suffix sorter is NOT libsais: the
 * BWT is a pure function of its input, so a plain
 * prefix-doubling sorter is used and checked against libsais_bwt through the
 * reference build.
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* error codes, include/libbz3.h:47-55 */
enum {
    ORC_OK = 0, ORC_ERR_OOB = -1, ORC_ERR_BWT = -2, ORC_ERR_CRC = -3, ORC_ERR_MALFORMED = -4,
    ORC_ERR_TRUNC = -5, ORC_ERR_TOO_BIG = -6, ORC_ERR_INIT = -7, ORC_ERR_TOO_SMALL = -8
};

/* ------------------------------------------------------------------ */
/* little-endian helpers (include/common.h:39-48)                      */
static uint32_t ld32le(const uint8_t *p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
static void st32le(uint8_t *p, uint32_t v) {
    p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24);
}

ORC_API size_t orc_bound(size_t n) { return n + n / 50 + 32; } /* src/libbz3.c:510 */

/* ------------------------------------------------------------------ */
/* CRC: reflected CRC-32C polynomial 0x82F63B78, caller-chosen init,   */
/* no final xor (src/libbz3.c:37-72; table[1] = 0xF26B8303).           */
static uint32_t crc_tab[256];
static int crc_tab_ready = 0;
static void crc_build(void) {
    for (uint32_t b = 0; b < 256; b++) {
        uint32_t r = b;
        for (int k = 0; k < 8; k++) r = (r >> 1) ^ ((r & 1) ? 0x82F63B78u : 0u);
        crc_tab[b] = r;
    }
    crc_tab_ready = 1;
}
ORC_API uint32_t orc_crc32(uint32_t crc, const uint8_t *buf, size_t n) {
    if (!crc_tab_ready) crc_build();
    for (size_t i = 0; i < n; i++) crc = crc_tab[(crc ^ buf[i]) & 0xff] ^ (crc >> 8);
    return crc;
}

/* ------------------------------------------------------------------ */
/* mRLE (src/libbz3.c:264-329).                                        */
/* A run of L equal bytes of a flagged symbol becomes                  */
/*   sym, 255 x floor((L-1)/255), (L-1) mod 255                        */
/* a symbol is flagged iff sum over its runs of                        */
/*   (L-1) - floor((L-1)/255) - 1  is positive.                        */
ORC_API int32_t orc_mrle_encode(const uint8_t *in, int32_t n, uint8_t *out) {
    int64_t gain[256];
    memset(gain, 0, sizeof gain);
    for (int32_t i = 0; i < n;) {
        int32_t j = i + 1;
        while (j < n && in[j] == in[i]) j++;
        int64_t L = j - i;
        gain[in[i]] += (L - 1) - (L - 1) / 255 - 1;
        i = j;
    }
    int32_t op = 0;
    for (int b = 0; b < 32; b++) {
        int v = 0;
        for (int k = 0; k < 8; k++) v |= (gain[b * 8 + k] > 0) << k;
        out[op++] = (uint8_t)v;
    }
    for (int32_t i = 0; i < n;) {
        int32_t j = i + 1;
        while (j < n && in[j] == in[i]) j++;
        int32_t L = j - i;
        uint8_t c = in[i];
        if (gain[c] > 0) {
            out[op++] = c;
            int32_t rem = L; /* 255*q + k + 1 */
            while (rem > 255) { out[op++] = 255; rem -= 255; }
            out[op++] = (uint8_t)(rem - 1);
        } else {
            for (int32_t k = 0; k < L; k++) out[op++] = c;
        }
        i = j;
    }
    return op;
}

/* returns nonzero on failure (src/libbz3.c:303-329) */
ORC_API int orc_mrle_decode(const uint8_t *in, uint8_t *out, int32_t outlen, int32_t maxin) {
    if (maxin < 32) return 1;
    int flagged[256];
    for (int b = 0; b < 32; b++)
        for (int k = 0; k < 8; k++) flagged[b * 8 + k] = (in[b] >> k) & 1;
    int32_t ip = 32, op = 0;
    int last = -1; /* the reference's `pc` survives across tokens when input is truncated */
    while (op < outlen && ip < maxin) {
        int c = in[ip++];
        if (!flagged[c]) { out[op++] = (uint8_t)c; continue; }
        int32_t run = 0;
        while (ip < maxin) {
            last = in[ip++];
            if (last != 255) break;
            run += 255;
        }
        run += last + 1;
        while (run > 0 && op < outlen) { out[op++] = (uint8_t)c; run--; }
    }
    return op != outlen;
}

/* ------------------------------------------------------------------ */
/* LZP (src/libbz3.c:84-257): order-4 hashed prediction, 2^18 slots,   */
/* minimum match 40, escape byte 0xF2.                                 */
#define LZP_SLOTS (1 << 18)
#define LZP_MIN 40
#define LZP_ESC 0xF2

static uint32_t ld32ne(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static uint32_t lzp_hash(uint32_t ctx) { return ((ctx >> 15) ^ ctx ^ (ctx >> 3)) & (LZP_SLOTS - 1); }
static uint32_t lzp_ctx_at(const uint8_t *base, int32_t pos) {
    return (uint32_t)base[pos - 1] | ((uint32_t)base[pos - 2] << 8) | ((uint32_t)base[pos - 3] << 16) |
           ((uint32_t)base[pos - 4] << 24);
}

/* returns encoded size, or -1 when the input is too small / output would not shrink by 8 */
ORC_API int32_t orc_lzp_encode(const uint8_t *in, int32_t n, uint8_t *out, int32_t *lut /* LZP_SLOTS */) {
    if (n < LZP_MIN + 32) return -1; /* :244 */
    memset(lut, 0, sizeof(int32_t) * LZP_SLOTS);
    const int32_t out_stop = n - 8;       /* out_end - 8, out_end = out + n (:128,:248) */
    const int32_t scan_end = n - LZP_MIN - 32; /* :137 */
    int32_t ip = 0, op = 0, heur = 0;
    while (ip < 4) out[op++] = in[ip++];
    uint32_t ctx = lzp_ctx_at(in, ip);
    while (ip < scan_end && op < out_stop) {
        uint32_t h = lzp_hash(ctx);
        int32_t cand = lut[h];
        lut[h] = ip;
        int take = 0;
        int32_t len = 0;
        if (cand > 0 && ld32ne(in + ip + LZP_MIN - 4) == ld32ne(in + cand + LZP_MIN - 4) &&
            ld32ne(in + ip) == ld32ne(in + cand)) {
            int vetoed = heur > ip && ld32ne(in + heur) != ld32ne(in + cand + (heur - ip)); /* :145 */
            if (!vetoed) {
                len = 4;
                while (ip + len < scan_end && ld32ne(in + ip + len) == ld32ne(in + cand + len)) len += 4;
                if (len < LZP_MIN) {
                    if (heur < ip + len) heur = ip + len;
                } else {
                    take = 1;
                }
            }
        }
        if (take) {
            /* up to three more single bytes, each test depending on the previous (:157-159) */
            for (int k = 0; k < 3; k++) len += in[ip + len] == in[cand + len];
            ip += len;
            ctx = lzp_ctx_at(in, ip);
            out[op++] = LZP_ESC;
            len -= LZP_MIN;
            while (len >= 254) {
                len -= 254;
                out[op++] = 254;
                if (op >= out_stop) break;
            }
            out[op++] = (uint8_t)len;
        } else {
            uint8_t b = in[ip++];
            out[op++] = b;
            ctx = (ctx << 8) | b;
            if (cand > 0 && b == LZP_ESC) out[op++] = 255; /* escape only when the slot was live (:176-181) */
        }
    }
    ctx = lzp_ctx_at(in, ip);
    while (ip < n && op < out_stop) { /* tail: literals only (:187-195) */
        uint32_t h = lzp_hash(ctx);
        int32_t cand = lut[h];
        lut[h] = ip;
        uint8_t b = in[ip++];
        out[op++] = b;
        ctx = (ctx << 8) | b;
        if (b == LZP_ESC && cand > 0) out[op++] = 255;
    }
    return op >= out_stop ? -1 : op;
}

/* returns decoded size or -1 on a truncated token (src/libbz3.c:200-257) */
ORC_API int32_t orc_lzp_decode(const uint8_t *in, int32_t n, uint8_t *out, int32_t max, int32_t *lut) {
    if (n < 4) return -1;
    memset(lut, 0, sizeof(int32_t) * LZP_SLOTS);
    int32_t ip = 0, op = 0;
    while (ip < 4) out[op++] = in[ip++];
    uint32_t ctx = lzp_ctx_at(out, op);
    while (ip < n && op < max) {
        uint32_t h = lzp_hash(ctx);
        int32_t cand = lut[h];
        lut[h] = op;
        if (in[ip] == LZP_ESC && cand > 0) {
            ip++;
            if (ip == n) return -1;
            if (in[ip] == 255) { /* escaped literal */
                ip++;
                out[op++] = LZP_ESC;
                ctx = (ctx << 8) | LZP_ESC;
            } else {
                int32_t len = LZP_MIN;
                for (;;) {
                    if (ip == n) return -1;
                    uint8_t b = in[ip++];
                    len += b;
                    if (b != 254) break;
                }
                int32_t stop = op + len;
                if (stop > max) stop = max;
                int32_t src = cand;
                while (op < stop) out[op++] = out[src++]; /* forward byte copy, may overlap */
                ctx = lzp_ctx_at(out, op);
            }
        } else {
            uint8_t b = in[ip++];
            out[op++] = b;
            ctx = (ctx << 8) | b;
        }
    }
    return op;
}

/* ------------------------------------------------------------------ */
/* BWT.  Definition restated from include/libsais.h:4095-4121 (and the */
/* final induce pass :2789/:3039): with SA the suffix array of T under */
/* plain lexicographic order where a proper prefix sorts first,        */
/*   U = T[n-1] ++ [ T[SA[i]-1] : i = 0..n-1, SA[i] != 0 ]             */
/*   idx = 1 + (i such that SA[i] == 0)                                */
/* Suffix sorting here: prefix doubling with per-group quicksort.      */
static int cmp_u64(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return (x > y) - (x < y);
}

static int32_t *orc_suffix_array(const uint8_t *T, int32_t n) {
    int32_t *sa = malloc(sizeof(int32_t) * (size_t)(n + 1));
    int32_t *rk = malloc(sizeof(int32_t) * (size_t)(n + 1));
    int32_t *nr = malloc(sizeof(int32_t) * (size_t)(n + 1));
    uint64_t *tmp = malloc(sizeof(uint64_t) * (size_t)(n + 1));
    if (!sa || !rk || !nr || !tmp) { free(sa); free(rk); free(nr); free(tmp); return NULL; }
    /* round 0: bucket by first byte; rank = index of first element of the group */
    int32_t cnt[257];
    memset(cnt, 0, sizeof cnt);
    for (int32_t i = 0; i < n; i++) cnt[T[i] + 1]++;
    for (int c = 0; c < 256; c++) cnt[c + 1] += cnt[c];
    {
        int32_t pos[256];
        for (int c = 0; c < 256; c++) pos[c] = cnt[c];
        for (int32_t i = 0; i < n; i++) sa[pos[T[i]]++] = i;
    }
    for (int32_t i = 0; i < n; i++) rk[i] = cnt[T[i]];
    for (int32_t h = 1;; h *= 2) {
        int any = 0;
        int32_t i = 0;
        while (i < n) {
            int32_t g = rk[sa[i]];
            int32_t j = i + 1;
            while (j < n && rk[sa[j]] == g) j++;
            if (j - i > 1) {
                any = 1;
                for (int32_t k = i; k < j; k++) {
                    int32_t s = sa[k];
                    /* past-the-end rank is 0, real ranks are shifted by one: shorter suffix sorts first */
                    uint64_t key = (s + h < n) ? (uint64_t)rk[s + h] + 1 : 0;
                    tmp[k] = (key << 32) | (uint32_t)s;
                }
                qsort(tmp + i, (size_t)(j - i), sizeof(uint64_t), cmp_u64);
                int32_t head = i;
                for (int32_t k = i; k < j; k++) {
                    if (k > i && (tmp[k] >> 32) != (tmp[k - 1] >> 32)) head = k;
                    sa[k] = (int32_t)(uint32_t)tmp[k];
                    nr[sa[k]] = head;
                }
            } else {
                nr[sa[i]] = i;
            }
            i = j;
        }
        int32_t *t = rk; rk = nr; nr = t;
        if (!any || h > n) break;
    }
    free(rk); free(nr); free(tmp);
    return sa;
}

/* returns primary index (1..n), n for n<=1, or -1 on allocation failure */
ORC_API int32_t orc_bwt(const uint8_t *T, uint8_t *U, int32_t n) {
    if (n < 0) return -1;
    if (n <= 1) { if (n == 1) U[0] = T[0]; return n; } /* include/libsais.h:4098-4108 */
    int32_t *sa = orc_suffix_array(T, n);
    if (!sa) return -1;
    int32_t idx = -1, op = 0;
    U[op++] = T[n - 1];
    for (int32_t i = 0; i < n; i++) {
        if (sa[i] == 0) idx = i + 1;
        else U[op++] = T[sa[i] - 1];
    }
    free(sa);
    return idx;
}

/* Inverse BWT; returns 0 or -1 (argument validation of include/libsais.h:5210-5232).
 *
 * Rows 0..n are the sorted suffixes of T$ (row 0 = the empty suffix); row `idx` is the one preceded
 * by $ and owns no byte of L.  psi maps the row of suffix T[k..] to the row of T[k+1..]; the text is
 * read off by walking psi forwards from row idx.  libsais walks two characters per step through a
 * bigram table (include/libsais.h:4555-4633, :5133-5156); for a VALID transform that is the same
 * walk.  For an INVALID one (corrupt stream) the psi graph is a path idx -> ... -> row 0 plus stray
 * cycles, and what libsais emits once the path runs out is restated here step for step so that the
 * block decoder reaches the same CRC verdict as the reference on hostile input:
 *   - after the path reaches row 0 every further pair is the smallest bigram present,
 *   - a pair that starts exactly on the last path row q* (the length-1 suffix) is looked up in the
 *     bigram table through the 2^17-entry "fastbits" accelerator, whose rounding is visible there,
 *   - the final byte is always L[0].
 */
ORC_API int32_t orc_unbwt(const uint8_t *L, uint8_t *T, int32_t n, int32_t idx) {
    if (n < 0) return -1;
    if (n <= 1) { if (idx != n) return -1; if (n == 1) T[0] = L[0]; return 0; }
    if (idx <= 0 || idx > n) return -1;
    uint32_t start[257];
    uint32_t cnt[256];
    memset(cnt, 0, sizeof cnt);
    for (int32_t i = 0; i < n; i++) cnt[L[i]]++;
    start[0] = 1;
    for (int c = 0; c < 256; c++) start[c + 1] = start[c] + cnt[c];
    uint32_t *psi = calloc((size_t)n + 2, sizeof(uint32_t));
    uint8_t *F = malloc((size_t)n + 2);
    uint32_t *big = calloc(65536, sizeof(uint32_t));
    if (!psi || !F || !big) { free(psi); free(F); free(big); return -1; }
    {
        uint32_t fill[256];
        for (int c = 0; c < 256; c++) fill[c] = start[c];
        for (int32_t r = 0; r <= n; r++) {
            if (r == idx) continue;
            uint8_t c = L[r < idx ? r : r - 1];
            psi[fill[c]++] = (uint32_t)r;
        }
        F[0] = 0;
        for (int c = 0; c < 256; c++)
            for (uint32_t q = start[c]; q < start[c + 1]; q++) F[q] = (uint8_t)c;
    }
    const uint32_t lastc = L[0];
    const uint32_t qstar = start[lastc]; /* row of the one-byte suffix: first row of its bucket */
    /* bigram census over rows 1..n except q* : (F(q), F(psi(q))) */
    for (uint32_t q = 1; q <= (uint32_t)n; q++)
        if (q != qstar) big[((uint32_t)F[q] << 8) | F[psi[q]]]++;
    uint32_t wmin = 0;
    while (wmin < 65536 && big[wmin] == 0) wmin++;
    /* pair emitted when a step starts on q*: bigram-table lookup with the fastbits hint */
    uint32_t wstar;
    {
        int shift = 0;
        while ((n >> shift) > (1 << 17)) shift++;
        uint32_t sum = 1, hint = 65536; /* 65536 = not assigned yet; unassigned fastbits slots read as 0 */
        uint32_t *end = malloc(65537 * sizeof(uint32_t));
        if (!end) { free(psi); free(F); free(big); return -1; }
        for (uint32_t w = 0; w < 65536; w++) {
            if ((w & 255) == 0 && (w >> 8) == lastc) sum += 1;
            sum += big[w];
            end[w] = sum;
            if (big[w] && hint == 65536 && ((sum - 1) >> shift) >= (qstar >> shift)) hint = w;
        }
        end[65536] = 0xFFFFFFFFu;
        wstar = hint == 65536 ? 0 : hint;
        while (end[wstar] <= qstar) wstar++;
        free(end);
    }
    uint32_t p = (uint32_t)idx;
    for (int32_t i = 0; i < (n >> 1); i++) {
        uint32_t w;
        if (p == 0) {
            w = wmin;
        } else if (p == qstar) {
            w = wstar;
            p = 0;
        } else {
            uint32_t r = psi[p];
            w = ((uint32_t)F[p] << 8) | F[r];
            p = (r == qstar) ? 0 : psi[r];
        }
        T[2 * i] = (uint8_t)(w >> 8);
        T[2 * i + 1] = (uint8_t)w;
    }
    T[n - 1] = (uint8_t)lastc;
    free(psi); free(F); free(big);
    return 0;
}

/* ------------------------------------------------------------------ */
/* Context-mixing model + 32-bit carry-less binary range coder         */
/* (src/libbz3.c:333-494).                                             */
typedef struct {
    uint16_t c0[256];      /* order-0, rate 2 */
    uint16_t c1[256][256]; /* order-1 keyed by a previous byte, rate 4 */
    uint16_t c2[512][17];  /* SSE/APM rows keyed by 2*node + runflag, rate 6 */
} orc_model;

static void model_reset(orc_model *m) { /* :350-358 */
    for (int i = 0; i < 256; i++) m->c0[i] = 32768;
    for (int i = 0; i < 256; i++)
        for (int j = 0; j < 256; j++) m->c1[i][j] = 32768;
    for (int r = 0; r < 512; r++)
        for (int k = 0; k < 17; k++) m->c2[r][k] = (uint16_t)((k << 12) - (k == 16));
}

static inline void adapt(uint16_t *p, int bit, int rate) { /* :347-348 */
    unsigned v = *p;
    if (bit) v += (v ^ 65535u) >> rate; else v -= v >> rate;
    *p = (uint16_t)v;
}

/* 18-bit probability that the next bit is 1, plus the SSE cell index (:377-385) */
static inline uint32_t predict(const orc_model *m, int node, int prev1, int prev2, int flag, int *cell) {
    int a = m->c0[node], b = m->c1[prev1][node], c = m->c1[prev2][node];
    int p = ((a + b) * 7 + c + c) >> 4;
    int j = p >> 12;
    int lo = m->c2[2 * node + flag][j], hi = m->c2[2 * node + flag][j + 1];
    int sse = lo + (((hi - lo) * (p & 4095)) >> 12); /* arithmetic shift of a possibly negative product */
    *cell = j;
    return (uint32_t)(sse * 3 + p);
}

static inline void learn(orc_model *m, int node, int prev1, int flag, int cell, int bit) { /* :396-399 */
    adapt(&m->c0[node], bit, 2);
    adapt(&m->c1[prev1][node], bit, 4);
    adapt(&m->c2[2 * node + flag][cell], bit, 6);
    adapt(&m->c2[2 * node + flag][cell + 1], bit, 6);
}

/* returns number of bytes written */
ORC_API int32_t orc_cm_encode(const uint8_t *in, int32_t n, uint8_t *out) {
    orc_model *m = malloc(sizeof *m);
    if (!m) return -1;
    model_reset(m);
    uint32_t low = 0, high = 0xFFFFFFFFu;
    int prev1 = 0, prev2 = 0;
    unsigned run = 0;
    int32_t op = 0;
    for (int32_t i = 0; i < n; i++) {
        run = (prev1 == prev2) ? run + 1 : 0;
        int flag = run > 2;
        int node = 1;
        unsigned sym = in[i];
        for (int k = 7; k >= 0; k--) {
            int bit = (sym >> k) & 1, cell;
            uint32_t P = predict(m, node, prev1, prev2, flag, &cell);
            uint32_t split = low + (uint32_t)(((uint64_t)(high - low) * P) >> 18);
            if (bit) high = split; else low = split + 1;
            while ((low ^ high) < (1u << 24)) {
                out[op++] = (uint8_t)(low >> 24);
                low <<= 8;
                high = (high << 8) | 0xFF;
            }
            learn(m, node, prev1, flag, cell, bit);
            node = node * 2 + bit;
        }
        prev2 = prev1;
        prev1 = node & 255;
    }
    for (int k = 0; k < 4; k++) { out[op++] = (uint8_t)(low >> 24); low <<= 8; } /* :425-432 */
    free(m);
    return op;
}

ORC_API int32_t orc_cm_decode(const uint8_t *in, int32_t insize, uint8_t *out, int32_t n) {
    orc_model *m = malloc(sizeof *m);
    if (!m) return -1;
    model_reset(m);
    uint32_t low = 0, high = 0xFFFFFFFFu, code = 0;
    int32_t ip = 0;
    /* reads past the end yield -1, i.e. 0xFFFFFFFF added in (:345, :438-441) */
#define NEXT_BYTE() ((ip < insize) ? (uint32_t)in[ip++] : 0xFFFFFFFFu)
    for (int k = 0; k < 4; k++) code = (code << 8) + NEXT_BYTE();
    int prev1 = 0, prev2 = 0;
    unsigned run = 0;
    for (int32_t i = 0; i < n; i++) {
        run = (prev1 == prev2) ? run + 1 : 0;
        int flag = run > 2;
        int node = 1;
        while (node < 256) {
            int cell;
            uint32_t P = predict(m, node, prev1, prev2, flag, &cell);
            uint32_t split = low + (uint32_t)(((uint64_t)(high - low) * P) >> 18);
            int bit = code <= split;
            if (bit) high = split; else low = split + 1;
            while ((low ^ high) < (1u << 24)) {
                low <<= 8;
                high = (high << 8) | 0xFF;
                code = (code << 8) + NEXT_BYTE();
            }
            learn(m, node, prev1, flag, cell, bit);
            node = node * 2 + bit;
        }
        prev2 = prev1;
        prev1 = node & 255;
        out[i] = (uint8_t)prev1;
    }
#undef NEXT_BYTE
    free(m);
    return 0;
}

/* ------------------------------------------------------------------ */
/* Block codec (src/libbz3.c:585-809).  `buf` is transformed in place. */
/* *err receives the value bz3_last_error() would report, or is left   */
/* untouched on the <64-byte raw paths exactly like the reference.     */
ORC_API int32_t orc_encode_block(int32_t block_size, uint8_t *buf, int32_t size, int8_t *err) {
    if (size > block_size) { *err = ORC_ERR_TOO_BIG; return -1; }
    uint32_t crc = orc_crc32(1, buf, (size_t)size);
    if (size < 64) { /* :596-601 */
        memmove(buf + 8, buf, (size_t)size);
        st32le(buf, crc);
        st32le(buf + 4, 0xFFFFFFFFu);
        return size + 8;
    }
    size_t cap = orc_bound((size_t)block_size);
    uint8_t *a = malloc(cap + 64), *b = malloc(cap + 64);
    int32_t *lut = malloc(sizeof(int32_t) * LZP_SLOTS);
    if (!a || !b || !lut) { free(a); free(b); free(lut); *err = ORC_ERR_INIT; return -1; }
    memcpy(a, buf, (size_t)size);
    int model = 0;
    int32_t cur = size, lzp_size = -1, rle_size;
    rle_size = orc_mrle_encode(a, cur, b);
    if (rle_size < cur) { uint8_t *t = a; a = b; b = t; cur = rle_size; model |= 4; }
    lzp_size = orc_lzp_encode(a, cur, b, lut);
    if (lzp_size > 0 && lzp_size < cur) { uint8_t *t = a; a = b; b = t; cur = lzp_size; model |= 2; }
    int32_t idx = orc_bwt(a, b, cur);
    if (idx < 0) { free(a); free(b); free(lut); *err = ORC_ERR_BWT; return -1; }
    int hdr = 9 + ((model & 2) ? 4 : 0) + ((model & 4) ? 4 : 0);
    int32_t payload = orc_cm_encode(b, cur, buf + hdr);
    st32le(buf, crc);
    st32le(buf + 4, (uint32_t)idx);
    buf[8] = (uint8_t)model;
    int at = 9;
    if (model & 2) { st32le(buf + at, (uint32_t)lzp_size); at += 4; }
    if (model & 4) { st32le(buf + at, (uint32_t)rle_size); at += 4; }
    free(a); free(b); free(lut);
    *err = ORC_OK;
    return payload + hdr;
}

ORC_API int32_t orc_decode_block(int32_t block_size, uint8_t *buf, size_t buf_size, int32_t csize, int32_t osize,
                                 int8_t *err) {
    const int64_t bound = (int64_t)orc_bound((size_t)block_size);
    /* :658 compares size_t with a converted signed value: negative csize becomes huge */
    if (buf_size < 9 || buf_size < (size_t)(int64_t)csize) { *err = ORC_ERR_TOO_SMALL; return -1; }
    uint32_t crc = ld32le(buf);
    int32_t idx = (int32_t)ld32le(buf + 4);
    if (csize < 0 || (int64_t)csize > bound) { *err = ORC_ERR_MALFORMED; return -1; }
    if (idx == -1) { /* raw block :672-692 */
        if (csize - 8 > 64 || csize < 8) { *err = ORC_ERR_MALFORMED; return -1; }
        if ((size_t)(csize - 8) > buf_size) { *err = ORC_ERR_TOO_SMALL; return -1; }
        memmove(buf, buf + 8, (size_t)(csize - 8));
        if (orc_crc32(1, buf, (size_t)(csize - 8)) != crc) { *err = ORC_ERR_CRC; return -1; }
        return csize - 8;
    }
    int model = (int8_t)buf[8];
    size_t need = 9 + (size_t)((model & 2) * 4) + (size_t)((model & 4) * 4); /* sic :697 */
    if (buf_size < need) { *err = ORC_ERR_TOO_SMALL; return -1; }
    int32_t lzp_size = -1, rle_size = -1;
    int at = 9;
    if (model & 2) { lzp_size = (int32_t)ld32le(buf + at); at += 4; }
    if (model & 4) { rle_size = (int32_t)ld32le(buf + at); at += 4; }
    int32_t payload = csize - at;
    if (((model & 2) && (lzp_size < 0 || lzp_size > bound)) || ((model & 4) && (rle_size < 0 || rle_size > bound))) {
        *err = ORC_ERR_MALFORMED; return -1;
    }
    if (osize < 0 || osize > bound) { *err = ORC_ERR_MALFORMED; return -1; }
    int32_t n = (model & 2) ? lzp_size : (model & 4) ? rle_size : osize;
    {
        size_t l = lzp_size < 0 ? 0 : (size_t)lzp_size, r = rle_size < 0 ? 0 : (size_t)rle_size;
        if (l > buf_size || r > buf_size || (size_t)osize > buf_size) { *err = ORC_ERR_TOO_SMALL; return -1; }
    }
    size_t cap = (size_t)bound + 64;
    uint8_t *a = malloc(cap), *b = malloc(cap);
    int32_t *lut = malloc(sizeof(int32_t) * LZP_SLOTS);
    if (!a || !b || !lut) { free(a); free(b); free(lut); *err = ORC_ERR_INIT; return -1; }
    int32_t ret = -1;
    orc_cm_decode(buf + at, payload, a, n);
    if (idx > n) { *err = ORC_ERR_MALFORMED; goto done; }
    if (orc_unbwt(a, b, n, idx) < 0) { *err = ORC_ERR_BWT; goto done; }
    { uint8_t *t = a; a = b; b = t; }
    int32_t cur = n;
    if (model & 2) {
        cur = orc_lzp_decode(a, lzp_size, b, (int32_t)bound, lut);
        if (cur == -1) { *err = ORC_ERR_CRC; goto done; }
        if ((size_t)cur > buf_size) { *err = ORC_ERR_TOO_SMALL; goto done; }
        uint8_t *t = a; a = b; b = t;
    }
    if (model & 4) {
        if (orc_mrle_decode(a, b, osize, cur)) { *err = ORC_ERR_CRC; goto done; }
        cur = osize;
        uint8_t *t = a; a = b; b = t;
    }
    *err = ORC_OK;
    if (cur > block_size || cur < 0) { *err = ORC_ERR_MALFORMED; goto done; }
    memcpy(buf, a, (size_t)cur);
    if (orc_crc32(1, buf, (size_t)cur) != crc) { *err = ORC_ERR_CRC; goto done; }
    ret = cur;
done:
    free(a); free(b); free(lut);
    return ret;
}
