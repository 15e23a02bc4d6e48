/*
 * oracle/rans_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of the integer arithmetic the reference's encode/decode
 * path delegates to compressai==1.1.5 (pinned in /root/reference/requirements/
 * environment.yaml:105; the package is NOT vendored under /root/reference and is not
 * installable offline, so its published algorithm is restated here):
 *
 *   - compressai/cpp_exts/ops/ops.cpp           pmf_to_quantized_cdf
 *   - compressai/cpp_exts/rans/rans_interface.cpp  RansEncoder::encode_with_indexes,
 *                                                RansDecoder::decode_with_indexes
 *   - third_party/ryg_rans/rans64.h             Rans64{EncInit,EncPut,EncFlush,DecInit,
 *                                                DecGet,DecAdvance} (64-bit state,
 *                                                32-bit renormalisation words)
 *
 * Reference call sites this stands in for: hub/compressor.py:63 (update -> tables),
 * hub/compressor.py:98 (compress), hub/compressor.py:124 (decompress),
 * lossyless/rates.py:64-71, lossyless/rates.py:556-564.
 *
 * PARITY UNPINNED: the reference ships no golden bitstream for this path (it has no
 * tests at all).  The only reproducible pin is results/exp_banana_viz_VIC/
 * summarized_metrics_featurizer.csv n_bits == 64.0 (z_dim = 2 -> bare 8-byte flush),
 * checked in tests/test_oracle.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this file's shared object.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define RANS64_L (1ull << 31)
#define PRECISION 16
#define BYPASS_PRECISION 4
#define MAX_BYPASS_VAL ((1 << BYPASS_PRECISION) - 1)

/* ------------------------------------------------------------------ ops.cpp */

/* pmf[n] (float32) -> cdf[n+1] (uint32).  Returns 0 on success, <0 on the conditions
 * the C++ original only guards with assert()/throws. */
int oracle_pmf_to_quantized_cdf(const float *pmf, int n, int precision, uint32_t *cdf)
{
    int i, j;
    for (i = 0; i < n; ++i)
        if (!(pmf[i] >= 0.0f) || !isfinite(pmf[i])) return -1;

    cdf[0] = 0;
    for (i = 0; i < n; ++i)
        /* std::round(p * (1 << precision)) : float * int -> float, round half away */
        cdf[i + 1] = (uint32_t)roundf(pmf[i] * (float)(1 << precision));

    uint32_t total = 0;
    for (i = 0; i <= n; ++i) total += cdf[i];
    if (total == 0) return -2;

    for (i = 0; i <= n; ++i)
        cdf[i] = (uint32_t)((((uint64_t)1 << precision) * (uint64_t)cdf[i]) / total);

    for (i = 1; i <= n; ++i) cdf[i] += cdf[i - 1];
    cdf[n] = 1u << precision;

    for (i = 0; i < n; ++i) {
        if (cdf[i] == cdf[i + 1]) {
            /* steal one count from the symbol with the smallest frequency > 1 */
            uint32_t best_freq = ~0u;
            int best_steal = -1;
            for (j = 0; j < n; ++j) {
                uint32_t freq = cdf[j + 1] - cdf[j];
                if (freq > 1 && freq < best_freq) {
                    best_freq = freq;
                    best_steal = j;
                }
            }
            if (best_steal == -1) return -3;
            if (best_steal < i) {
                for (j = best_steal + 1; j <= i; ++j) cdf[j]--;
            } else {
                for (j = i + 1; j <= best_steal; ++j) cdf[j]++;
            }
        }
    }
    if (cdf[0] != 0 || cdf[n] != (1u << precision)) return -4;
    for (i = 0; i < n; ++i)
        if (!(cdf[i + 1] > cdf[i])) return -5;
    return 0;
}

/* ------------------------------------------------------------ rans64.h core */

typedef struct {
    uint16_t start;
    uint16_t range;
    uint8_t bypass;
} sym_t;

static inline void enc_put(uint64_t *r, uint32_t **pptr, uint32_t start, uint32_t freq,
                           uint32_t scale_bits)
{
    uint64_t x = *r;
    uint64_t x_max = ((RANS64_L >> scale_bits) << 32) * freq;
    if (x >= x_max) {
        *pptr -= 1;
        **pptr = (uint32_t)x;
        x >>= 32;
    }
    *r = ((x / freq) << scale_bits) + (x % freq) + start;
}

static inline void enc_put_bits(uint64_t *r, uint32_t **pptr, uint32_t val, uint32_t nbits)
{
    uint64_t x = *r;
    uint32_t freq = 1u << (16 - nbits);
    uint64_t x_max = ((RANS64_L >> 16) << 32) * freq;
    if (x >= x_max) {
        *pptr -= 1;
        **pptr = (uint32_t)x;
        x >>= 32;
    }
    *r = (x << nbits) | val;
}

static inline uint32_t dec_get_bits(uint64_t *r, const uint32_t **pptr, uint32_t nbits)
{
    uint64_t x = *r;
    uint32_t val = (uint32_t)(x & ((1u << nbits) - 1));
    x >>= nbits;
    if (x < RANS64_L) {
        x = (x << 32) | **pptr;
        *pptr += 1;
    }
    *r = x;
    return val;
}

/* ------------------------------------------------- rans_interface.cpp: encode */

/* Worst-case number of pushed entries for n symbols: 1 + (1 count nibble, since
 * n_bypass <= 8 < 15) + 8 raw nibbles. */
#define MAX_PUSH_PER_SYMBOL 10

/* Encodes n symbols; writes the byte string to out (capacity out_cap bytes).
 * cdfs is a dense [n_cdfs, cdf_stride] int32 table.  Returns the number of bytes,
 * or <0 on error (-1 alloc, -2 capacity). */
int64_t oracle_rans_encode(const int32_t *symbols, const int32_t *indexes, int n,
                           const int32_t *cdfs, int cdf_stride, const int32_t *cdf_sizes,
                           const int32_t *offsets, uint8_t *out, int64_t out_cap)
{
    int64_t cap_syms = (int64_t)n * MAX_PUSH_PER_SYMBOL + 4;
    sym_t *syms = (sym_t *)malloc(sizeof(sym_t) * (size_t)cap_syms);
    if (!syms) return -1;
    int64_t ns = 0;
    int i, j;

    for (i = 0; i < n; ++i) {
        const int32_t cdf_idx = indexes[i];
        const int32_t *cdf = cdfs + (int64_t)cdf_idx * cdf_stride;
        const int32_t max_value = cdf_sizes[cdf_idx] - 2;
        int32_t value = symbols[i] - offsets[cdf_idx];
        uint32_t raw_val = 0;
        if (value < 0) {
            raw_val = (uint32_t)(-2 * value - 1);
            value = max_value;
        } else if (value >= max_value) {
            raw_val = (uint32_t)(2 * (value - max_value));
            value = max_value;
        }
        syms[ns].start = (uint16_t)cdf[value];
        syms[ns].range = (uint16_t)(cdf[value + 1] - cdf[value]);
        syms[ns].bypass = 0;
        ns++;

        if (value == max_value) {
            int32_t n_bypass = 0;
            while ((raw_val >> (n_bypass * BYPASS_PRECISION)) != 0) ++n_bypass;

            int32_t val = n_bypass;
            while (val >= MAX_BYPASS_VAL) {
                syms[ns].start = MAX_BYPASS_VAL;
                syms[ns].range = MAX_BYPASS_VAL + 1;
                syms[ns].bypass = 1;
                ns++;
                val -= MAX_BYPASS_VAL;
            }
            syms[ns].start = (uint16_t)val;
            syms[ns].range = (uint16_t)(val + 1);
            syms[ns].bypass = 1;
            ns++;

            for (j = 0; j < n_bypass; ++j) {
                const int32_t v = (int32_t)((raw_val >> (j * BYPASS_PRECISION)) & MAX_BYPASS_VAL);
                syms[ns].start = (uint16_t)v;
                syms[ns].range = (uint16_t)(v + 1);
                syms[ns].bypass = 1;
                ns++;
            }
        }
    }

    /* flush(): walk the pushed list backwards, words written back to front */
    int64_t nwords = ns + 2;
    uint32_t *buf = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)nwords);
    if (!buf) {
        free(syms);
        return -1;
    }
    uint32_t *ptr = buf + nwords;
    uint64_t rans = RANS64_L;
    for (int64_t k = ns - 1; k >= 0; --k) {
        if (!syms[k].bypass)
            enc_put(&rans, &ptr, syms[k].start, syms[k].range, PRECISION);
        else
            enc_put_bits(&rans, &ptr, syms[k].start, BYPASS_PRECISION);
    }
    ptr -= 2;
    ptr[0] = (uint32_t)(rans >> 0);
    ptr[1] = (uint32_t)(rans >> 32);

    int64_t nbytes = (int64_t)((buf + nwords) - ptr) * 4;
    int64_t rv = nbytes;
    if (nbytes > out_cap)
        rv = -2;
    else
        memcpy(out, ptr, (size_t)nbytes);
    free(buf);
    free(syms);
    return rv;
}

/* ------------------------------------------------- rans_interface.cpp: decode */

/* Returns 0, or -2 if the decoder would read past the end of the string. */
int oracle_rans_decode(const uint8_t *enc, int64_t nbytes, const int32_t *indexes, int n,
                       const int32_t *cdfs, int cdf_stride, const int32_t *cdf_sizes,
                       const int32_t *offsets, int32_t *out)
{
    if (nbytes < 8) return -2;
    /* copy to an aligned, zero-padded buffer: the original reads uint32 from the string */
    int64_t nwords = nbytes / 4;
    uint32_t *w = (uint32_t *)calloc((size_t)nwords + 16, 4);
    if (!w) return -1;
    memcpy(w, enc, (size_t)nwords * 4);
    const uint32_t *ptr = w;
    const uint32_t *end = w + nwords + 16;
    uint64_t rans = (uint64_t)ptr[0] | ((uint64_t)ptr[1] << 32);
    ptr += 2;

    for (int i = 0; i < n; ++i) {
        const int32_t cdf_idx = indexes[i];
        const int32_t *cdf = cdfs + (int64_t)cdf_idx * cdf_stride;
        const int32_t size = cdf_sizes[cdf_idx];
        const int32_t max_value = size - 2;
        const int32_t offset = offsets[cdf_idx];
        const uint32_t cum_freq = (uint32_t)(rans & ((1u << PRECISION) - 1));

        int k = 0;
        while (k < size && !((uint32_t)cdf[k] > cum_freq)) ++k; /* std::find_if */
        const uint32_t s = (uint32_t)(k - 1);

        {
            const uint32_t start = (uint32_t)cdf[s];
            const uint32_t freq = (uint32_t)(cdf[s + 1] - cdf[s]);
            uint64_t x = rans;
            x = freq * (x >> PRECISION) + (x & ((1ull << PRECISION) - 1)) - start;
            if (x < RANS64_L) {
                x = (x << 32) | *ptr;
                ptr += 1;
            }
            rans = x;
        }

        int32_t value = (int32_t)s;
        if (value == max_value) {
            int32_t val = (int32_t)dec_get_bits(&rans, &ptr, BYPASS_PRECISION);
            int32_t n_bypass = val;
            while (val == MAX_BYPASS_VAL) {
                val = (int32_t)dec_get_bits(&rans, &ptr, BYPASS_PRECISION);
                n_bypass += val;
                if (ptr >= end) { free(w); return -2; }
            }
            int32_t raw_val = 0;
            for (int j = 0; j < n_bypass; ++j) {
                val = (int32_t)dec_get_bits(&rans, &ptr, BYPASS_PRECISION);
                raw_val |= val << (j * BYPASS_PRECISION);
                if (ptr >= end) { free(w); return -2; }
            }
            value = raw_val >> 1;
            if (raw_val & 1)
                value = -value - 1;
            else
                value += max_value;
        }
        out[i] = value + offset;
        if (ptr >= end) { free(w); return -2; }
    }
    free(w);
    return 0;
}

/* --------------------------------------------------- quantise (float32, exact) */

/* compressai EntropyModel.quantize(inputs, "symbols", means) applied to
 * hub/compressor.py:105-109 process_z_in:  sym = int(rint((z + b) * exp_s - med)).
 * z is given as float32 (the .float() of the fp16 encoder output); exp_s = exp(scaling)
 * is passed in (it is computed by torch at load time, see oracle/eb.py).  volatile
 * temporaries forbid FMA contraction so every op rounds to float32 separately. */
void oracle_quantize_symbols(const float *z, int64_t rows, int c, const float *bias,
                             const float *exp_s, const float *medians, int32_t *sym)
{
    for (int64_t r = 0; r < rows; ++r)
        for (int k = 0; k < c; ++k) {
            volatile float a = z[r * c + k] + bias[k];
            volatile float m = a * exp_s[k];
            volatile float d = m - medians[k];
            sym[r * c + k] = (int32_t)rintf(d); /* torch.round = half-to-even */
        }
}

/* ------------------------------------ batch drivers for the CPU baseline legs */

/* Encode `rows` images of `c` symbols each (indexes = arange(c)), sequentially on the
 * calling thread.  out is [rows, cap] bytes, lens[rows].  Returns 0 or first error. */
int oracle_rans_encode_batch(const int32_t *symbols, int64_t rows, int c, const int32_t *cdfs,
                             int cdf_stride, const int32_t *cdf_sizes, const int32_t *offsets,
                             uint8_t *out, int64_t cap, int32_t *lens)
{
    int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * (size_t)c);
    if (!idx) return -1;
    for (int k = 0; k < c; ++k) idx[k] = k;
    int rc = 0;
    for (int64_t r = 0; r < rows; ++r) {
        int64_t nb = oracle_rans_encode(symbols + r * c, idx, c, cdfs, cdf_stride, cdf_sizes,
                                        offsets, out + r * cap, cap);
        if (nb < 0) { rc = (int)nb; break; }
        lens[r] = (int32_t)nb;
    }
    free(idx);
    return rc;
}

int oracle_rans_decode_batch(const uint8_t *enc, int64_t cap, const int32_t *lens, int64_t rows,
                             int c, const int32_t *cdfs, int cdf_stride,
                             const int32_t *cdf_sizes, const int32_t *offsets, int32_t *out)
{
    int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * (size_t)c);
    if (!idx) return -1;
    for (int k = 0; k < c; ++k) idx[k] = k;
    int rc = 0;
    for (int64_t r = 0; r < rows; ++r) {
        rc = oracle_rans_decode(enc + r * cap, lens[r], idx, c, cdfs, cdf_stride, cdf_sizes,
                                offsets, out + r * c);
        if (rc) break;
    }
    free(idx);
    return rc;
}
