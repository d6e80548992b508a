/*
 * oracle/rans_oracle.c -- CPU restatement of the entropy coder on the bottleneck path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is product code: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * PARITY UNPINNED: the arithmetic below lives in the third-party CompressAI package
 * (compressai>=1.2.3, /root/reference/setup.py:28), which is neither vendored in the
 * reference nor installable here.  This file restates its published algorithm
 *   - compressai/cpp_exts/rans/rans_interface.cpp  (RansEncoder / RansDecoder, 16-bit
 *     precision, 4-bit bypass escape coding) on top of ryg_rans' rans64.h, and
 *   - compressai/cpp_exts/ops/ops.cpp              (pmf_to_quantized_cdf)
 * following SURVEY.md Appendix A.3 / A.5; the reference reaches it from
 *   sc2bench/models/layer.py:506 (entropy_bottleneck.compress)
 *   sc2bench/models/layer.py:520 (entropy_bottleneck.decompress)
 *   sc2bench/models/layer.py:441 (update -> pmf_to_quantized_cdf).
 * The reference ships no tests or golden vectors for it.
 *
 * Plain C, no dependencies.  Build: see oracle/Makefile.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_PRECISION 16
#define ORC_BYPASS_PRECISION 4
#define ORC_MAX_BYPASS_VAL 15
#define ORC_RANS64_L (1ull << 31)

typedef struct {
    uint32_t start;
    uint32_t range; /* for bypass entries: unused */
    uint32_t bypass;
} orc_sym_t;

/* ---- symbol expansion (forward pass of encode_with_indexes) ---------------- */

/* Upper bound on coder entries produced by one input symbol: 1 regular + unary count of
 * nibbles (raw < 2^32 -> n_bypass <= 8 -> one unary digit) + 8 nibbles. */
#define ORC_MAX_ENTRIES_PER_SYMBOL 10

static size_t orc_expand(const int32_t *symbols, const int32_t *indexes, size_t n,
                         const int32_t *cdfs, int cdf_stride, const int32_t *cdf_sizes,
                         const int32_t *offsets, orc_sym_t *out) {
    size_t m = 0;
    for (size_t i = 0; i < n; ++i) {
        const int32_t row = indexes[i];
        const int32_t *cdf = cdfs + (size_t)row * cdf_stride;
        const int32_t max_value = cdf_sizes[row] - 2;
        int32_t value = symbols[i] - offsets[row];
        uint32_t raw = 0;
        if (value < 0) {
            raw = (uint32_t)(-2 * value - 1);
            value = max_value;
        } else if (value >= max_value) {
            raw = (uint32_t)(2 * (value - max_value));
            value = max_value;
        }
        out[m].start = (uint32_t)cdf[value];
        out[m].range = (uint32_t)(cdf[value + 1] - cdf[value]);
        out[m].bypass = 0;
        ++m;
        if (value == max_value) {
            int32_t n_bypass = 0;
            while (n_bypass < 8 && (raw >> (n_bypass * ORC_BYPASS_PRECISION)) != 0) ++n_bypass;
            int32_t val = n_bypass;
            while (val >= ORC_MAX_BYPASS_VAL) {
                out[m].start = ORC_MAX_BYPASS_VAL; out[m].range = 0; out[m].bypass = 1; ++m;
                val -= ORC_MAX_BYPASS_VAL;
            }
            out[m].start = (uint32_t)val; out[m].range = 0; out[m].bypass = 1; ++m;
            for (int32_t j = 0; j < n_bypass; ++j) {
                out[m].start = (uint32_t)((raw >> (j * ORC_BYPASS_PRECISION)) & ORC_MAX_BYPASS_VAL);
                out[m].range = 0; out[m].bypass = 1; ++m;
            }
        }
    }
    return m;
}

/* ---- encode_with_indexes: returns number of bytes written, or -1 on error ---------------
 * out must hold at least orc_rans_max_bytes(n) bytes.  The stream is the tail of a buffer
 * of (entries + 2) u32 words exactly as the reverse pass fills it (SURVEY A.5 "Flush"). */
long orc_rans_max_bytes(long n_symbols) {
    return 4 * ((long)ORC_MAX_ENTRIES_PER_SYMBOL * n_symbols + 2);
}

long orc_rans_encode(const int32_t *symbols, const int32_t *indexes, long n,
                     const int32_t *cdfs, int cdf_stride, const int32_t *cdf_sizes,
                     const int32_t *offsets, uint8_t *out, long out_capacity) {
    orc_sym_t *syms = (orc_sym_t *)malloc(sizeof(orc_sym_t) * (size_t)(ORC_MAX_ENTRIES_PER_SYMBOL * (n > 0 ? n : 1)));
    if (!syms) return -1;
    const size_t m = orc_expand(symbols, indexes, (size_t)n, cdfs, cdf_stride, cdf_sizes, offsets, syms);
    const size_t n_words = m + 2;
    uint32_t *buf = (uint32_t *)malloc(sizeof(uint32_t) * n_words);
    if (!buf) { free(syms); return -1; }
    uint32_t *ptr = buf + n_words;
    uint64_t x = ORC_RANS64_L;
    for (size_t k = m; k-- > 0;) {
        const orc_sym_t s = syms[k];
        if (!s.bypass) {
            const uint64_t x_max = ((ORC_RANS64_L >> ORC_PRECISION) << 32) * (uint64_t)s.range;
            if (x >= x_max) { *--ptr = (uint32_t)x; x >>= 32; }
            x = ((x / s.range) << ORC_PRECISION) + (x % s.range) + s.start;
        } else {
            const uint64_t freq = 1ull << (ORC_PRECISION - ORC_BYPASS_PRECISION);
            const uint64_t x_max = ((ORC_RANS64_L >> ORC_PRECISION) << 32) * freq;
            if (x >= x_max) { *--ptr = (uint32_t)x; x >>= 32; }
            x = (x << ORC_BYPASS_PRECISION) | s.start;
        }
    }
    ptr -= 2;
    ptr[0] = (uint32_t)x;
    ptr[1] = (uint32_t)(x >> 32);
    const long nbytes = (long)((buf + n_words) - ptr) * 4;
    long rv = nbytes;
    if (nbytes > out_capacity) rv = -2; else memcpy(out, ptr, (size_t)nbytes);
    free(buf);
    free(syms);
    return rv;
}

/* ---- decode_with_indexes ------------------------------------------------------------------
 * Returns 0, or -1 if the stream would be read past its end (the C++ original has no such
 * check; a well-formed stream never triggers it). */
static inline uint32_t orc_getbits(uint64_t *x, const uint32_t **pp, const uint32_t *end, int nbits, int *err) {
    const uint32_t val = (uint32_t)(*x & ((1u << nbits) - 1));
    *x >>= nbits;
    if (*x < ORC_RANS64_L) {
        if (*pp >= end) { *err = 1; return val; }
        *x = (*x << 32) | *(*pp)++;
    }
    return val;
}

int orc_rans_decode(const uint8_t *stream, long nbytes, const int32_t *indexes, long n,
                    const int32_t *cdfs, int cdf_stride, const int32_t *cdf_sizes,
                    const int32_t *offsets, int32_t *out) {
    if (nbytes < 8 || (nbytes & 3)) return -1;
    uint32_t *words = (uint32_t *)malloc((size_t)nbytes);
    if (!words) return -1;
    memcpy(words, stream, (size_t)nbytes);
    const uint32_t *p = words, *end = words + nbytes / 4;
    uint64_t x = (uint64_t)p[0] | ((uint64_t)p[1] << 32);
    p += 2;
    int err = 0;
    for (long i = 0; i < n && !err; ++i) {
        const int32_t row = indexes[i];
        const int32_t *cdf = cdfs + (size_t)row * cdf_stride;
        const int32_t size = cdf_sizes[row];
        const int32_t max_value = size - 2;
        const uint32_t cum = (uint32_t)(x & 0xFFFFu);
        int32_t k = 0;
        while (k < size && !((uint32_t)cdf[k] > cum)) ++k; /* first entry > cum (linear find_if) */
        const int32_t s = k - 1;
        const uint64_t start = (uint64_t)cdf[s], freq = (uint64_t)(cdf[s + 1] - cdf[s]);
        x = freq * (x >> ORC_PRECISION) + (x & 0xFFFFu) - start;
        if (x < ORC_RANS64_L) {
            if (p >= end) { err = 1; break; }
            x = (x << 32) | *p++;
        }
        int32_t value = s;
        if (value == max_value) {
            int32_t val = (int32_t)orc_getbits(&x, &p, end, ORC_BYPASS_PRECISION, &err);
            int32_t n_bypass = val;
            while (val == ORC_MAX_BYPASS_VAL && !err) {
                val = (int32_t)orc_getbits(&x, &p, end, ORC_BYPASS_PRECISION, &err);
                n_bypass += val;
            }
            uint32_t raw = 0;
            for (int32_t j = 0; j < n_bypass && !err; ++j) {
                val = (int32_t)orc_getbits(&x, &p, end, ORC_BYPASS_PRECISION, &err);
                if (j < 8) raw |= (uint32_t)val << (j * ORC_BYPASS_PRECISION);
            }
            value = (int32_t)(raw >> 1);
            if (raw & 1u) value = -value - 1; else value += max_value;
        }
        out[i] = value + offsets[row];
    }
    free(words);
    return err ? -1 : 0;
}

/* ---- pmf_to_quantized_cdf ------------------------------------------------------
 * cdf_out has n + 1 entries.  Returns 0; -1 for a negative / non-finite pmf entry;
 * -2 if the pmf sums to zero; -3 if no frequency can be stolen. */
int orc_pmf_to_quantized_cdf(const float *pmf, int n, int precision, uint32_t *cdf_out) {
    for (int i = 0; i < n; ++i)
        if (pmf[i] < 0.0f || !isfinite(pmf[i])) return -1;
    const int m = n + 1;
    cdf_out[0] = 0;
    for (int i = 0; i < n; ++i) cdf_out[i + 1] = (uint32_t)roundf(pmf[i] * (float)(1 << precision));
    int32_t total = 0; /* std::accumulate(..., 0) sums in int */
    for (int i = 0; i < m; ++i) total += (int32_t)cdf_out[i];
    if (total == 0) return -2;
    for (int i = 0; i < m; ++i)
        cdf_out[i] = (uint32_t)((((uint64_t)1 << precision) * cdf_out[i]) / (uint32_t)total);
    for (int i = 1; i < m; ++i) cdf_out[i] += cdf_out[i - 1];
    cdf_out[m - 1] = 1u << precision;
    for (int i = 0; i < m - 1; ++i) {
        if (cdf_out[i] == cdf_out[i + 1]) {
            uint32_t best_freq = ~0u;
            int best = -1;
            for (int j = 0; j < m - 1; ++j) {
                const uint32_t f = cdf_out[j + 1] - cdf_out[j];
                if (f > 1 && f < best_freq) { best_freq = f; best = j; }
            }
            if (best < 0) return -3;
            if (best < i) { for (int j = best + 1; j <= i; ++j) cdf_out[j]--; }
            else { for (int j = i + 1; j <= best; ++j) cdf_out[j]++; }
        }
    }
    return 0;
}
