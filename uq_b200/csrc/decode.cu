// Stage 5: uQ tables -> FASTQ text (replaces split_bits uq.py:1002-1007, convert_qname
// uq.py:1010-1026 and the emit loop uq.py:1030-1058).
//
// Two phases: (1) one thread per record computes the byte length of its four lines (QNAME text
// length from the column values, read length from the marker bit for variable-length files);
// an exclusive scan turns lengths into output offsets; (2) one warp per record writes the text:
// lane 0 formats the QNAME, all lanes unpack DNA / QUAL symbols.
#include "common.cuh"

#define DC 256

struct dec_col {
    const uint8_t* data;       // column values [n][itemsize], little endian
    const uint8_t* dict;       // mapping strings, zero padded rows (device)
    const uint32_t* dict_len;  // strlen of every dictionary row (device)
    uint64_t dict_count;
    uint32_t dict_width;
    uint32_t itemsize;
    uint32_t format, offset;
    long long min_val;
};

struct dec_params {
    uint8_t base_char[256];
    uint8_t qual_char[256];
    int16_t qual_to_base[256];
    uint32_t bb, bq, variable, dna_max, wd, wq;
    uint32_t prefix_len, suffix_len, nseps, ncols;
    uint8_t prefix[UQB_HDR_MAX];
    uint8_t suffix[UQB_HDR_MAX];
    uint8_t seps[UQB_MAX_COLS];
    dec_col cols[UQB_MAX_COLS];
};

__device__ __forceinline__ unsigned long long load_le(const uint8_t* p, uint32_t size) {
    unsigned long long v = 0;
    for (uint32_t i = 0; i < size; i++) v |= (unsigned long long)p[i] << (8 * i);
    return v;
}

__device__ __forceinline__ uint32_t dec_digits(long long v) {
    uint32_t n = v < 0 ? 1u : 0u;
    unsigned long long a = v < 0 ? (unsigned long long)(-(v + 1)) + 1ull : (unsigned long long)v;
    do { n++; a /= 10ull; } while (a);
    return n;
}

// read length of a variable-length row: the first set bit is the marker's
__device__ __forceinline__ uint32_t row_read_len(const uint8_t* row, uint32_t w, uint32_t bits, uint32_t dna_max, uint32_t variable) {
    if (!variable) return dna_max;
    for (uint32_t j = 0; j < w; j++) {
        unsigned b = row[j];
        if (b) {
            uint32_t p = 8 * j + (__clz(b) - 24);
            return (8 * w - 1 - p) / bits;
        }
    }
    return 0;
}

__device__ __forceinline__ uint32_t qname_text_len(const dec_params& P, uint64_t r) {
    uint32_t n = P.prefix_len + P.suffix_len;
    for (uint32_t c = 0; c < P.ncols; c++) {
        const dec_col& dc = P.cols[c];
        unsigned long long raw = load_le(dc.data + r * dc.itemsize, dc.itemsize);
        if (dc.format == 0) n += raw < dc.dict_count ? dc.dict_len[raw] : 0u;
        else n += dec_digits((long long)raw + (dc.offset ? dc.min_val : 0ll));
        if (c < P.nseps) n += 1;
    }
    return n;
}

__global__ void __launch_bounds__(DC) k_decode_len(const dec_params* __restrict__ Pp, const uint8_t* __restrict__ dna, uint64_t n,
                                                  uint32_t* __restrict__ rec_len) {
    const uint64_t r = (uint64_t)blockIdx.x * DC + threadIdx.x;
    if (r >= n) return;
    const dec_params& P = *Pp;
    const uint32_t len = row_read_len(dna + r * P.wd, P.wd, P.bb, P.dna_max, P.variable);
    rec_len[r] = qname_text_len(P, r) + 1 + len + 1 + 2 + len + 1;
}

__device__ __forceinline__ unsigned row_symbol(const uint8_t* row, uint32_t w, uint32_t bits, uint32_t pos) {
    const uint32_t j = pos >> 3;
    unsigned v = (unsigned)row[j] << 8;
    if (j + 1 < w) v |= row[j + 1];
    return (v >> (16 - (pos & 7u) - bits)) & ((1u << bits) - 1u);
}

__global__ void __launch_bounds__(DC) k_decode_write(const dec_params* __restrict__ Pp, const uint8_t* __restrict__ dna,
                                                    const uint8_t* __restrict__ qual, uint64_t n, const uint64_t* __restrict__ rec_off,
                                                    uint8_t* __restrict__ out) {
    __shared__ uint8_t base_char[256], qual_char[256];
    __shared__ int16_t qual_to_base[256];
    const dec_params& P = *Pp;
    for (unsigned i = threadIdx.x; i < 256; i += DC) { base_char[i] = P.base_char[i]; qual_char[i] = P.qual_char[i]; qual_to_base[i] = P.qual_to_base[i]; }
    __syncthreads();
    const unsigned lane = threadIdx.x & 31u;
    const uint64_t wstride = (uint64_t)gridDim.x * (DC / 32);
    const uint32_t nsym_row = P.dna_max + P.variable;
    const uint32_t pad_d = 8 * P.wd - P.bb * nsym_row, pad_q = 8 * P.wq - P.bq * nsym_row;
    for (uint64_t r = (uint64_t)blockIdx.x * (DC / 32) + (threadIdx.x >> 5); r < n; r += wstride) {
        const uint8_t* drow = dna + r * P.wd;
        const uint8_t* qrow = qual + r * P.wq;
        uint8_t* o = out + rec_off[r];
        // ---- QNAME: prefix + sum(column text + separator) + suffix (uq.py:1010-1024) ----
        // lane c formats column c: text length -> warp prefix sum over the columns -> every lane writes
        // its own column text (and the separator behind it) straight to its place in the output line
        uint32_t run = P.prefix_len;                // offset of the next column's text within the line
        for (uint32_t c0 = 0; c0 < P.ncols; c0 += 32) {
            const uint32_t c = c0 + lane;
            uint32_t tl = 0;
            unsigned long long raw = 0;
            long long v = 0;
            if (c < P.ncols) {
                const dec_col& dc = P.cols[c];
                raw = load_le(dc.data + r * dc.itemsize, dc.itemsize);
                if (dc.format == 0) {
                    tl = raw < dc.dict_count ? dc.dict_len[raw] : 0u;
                } else {
                    v = (long long)raw + (dc.offset ? dc.min_val : 0ll);
                    tl = dec_digits(v);
                }
                if (c < P.nseps) tl += 1;           // the separator that follows this column
            }
            uint32_t incl = tl;
#pragma unroll
            for (int sh = 1; sh < 32; sh <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, sh);
                if (lane >= (unsigned)sh) incl += t;
            }
            if (c < P.ncols) {
                const dec_col& dc = P.cols[c];
                uint8_t* w = o + run + incl - tl;
                uint32_t body = tl - (c < P.nseps ? 1u : 0u);
                if (dc.format == 0) {
                    if (raw < dc.dict_count) {
                        const uint8_t* sp = dc.dict + raw * dc.dict_width;
                        for (uint32_t i = 0; i < body; i++) w[i] = sp[i];
                    }
                } else {
                    unsigned long long a = v < 0 ? (unsigned long long)(-(v + 1)) + 1ull : (unsigned long long)v;
                    if (v < 0) w[0] = '-';
                    for (uint32_t i = 0; i < body - (v < 0 ? 1u : 0u); i++) { w[body - 1 - i] = (uint8_t)('0' + a % 10ull); a /= 10ull; }
                }
                if (c < P.nseps) w[body] = P.seps[c];
            }
            run += __shfl_sync(0xffffffffu, incl, 31);
        }
        for (uint32_t i = lane; i < P.prefix_len; i += 32) o[i] = P.prefix[i];
        for (uint32_t i = lane; i < P.suffix_len; i += 32) o[run + i] = P.suffix[i];
        if (lane == 0) o[run + P.suffix_len] = '\n';
        const uint32_t hl = run + P.suffix_len + 1;
        uint32_t len = 0;
        if (lane == 0) len = row_read_len(drow, P.wd, P.bb, P.dna_max, P.variable);
        len = __shfl_sync(0xffffffffu, len, 0);
        uint8_t* od = o + hl;                  // DNA line
        uint8_t* oq = od + len + 3;            // after "\n+\n"
        const uint32_t s_first = nsym_row - len;     // skips the padding symbols and the marker
        for (uint32_t i = lane; i < len; i += 32) {
            const uint32_t s = s_first + i;
            const unsigned dcode = row_symbol(drow, P.wd, P.bb, pad_d + P.bb * s);
            const unsigned qcode = row_symbol(qrow, P.wq, P.bq, pad_q + P.bq * s);
            const int restored = qual_to_base[qcode];          // qual_N (uq.py:1036)
            od[i] = restored >= 0 ? (uint8_t)restored : base_char[dcode];
            oq[i] = qual_char[qcode];
        }
        if (lane == 0) { od[len] = '\n'; od[len + 1] = '+'; od[len + 2] = '\n'; oq[len] = '\n'; }
    }
}

extern "C" int uqb_decode(uqb_ctx* ctx, const uqb_array* dna, const uqb_array* qual, uqb_array* const* cols,
                          const uqb_decode_params* p, uqb_array** fastq) {
    const uint64_t n = dna->n;
    if (qual->n != n) return uqb_fail(ctx, "decode: DNA has %llu rows, QUAL %llu", (unsigned long long)n, (unsigned long long)qual->n);
    if (p->ncols > UQB_MAX_COLS || p->nseps > UQB_MAX_COLS) return uqb_fail(ctx, "decode: too many QNAME columns");
    if (p->prefix_len > UQB_HDR_MAX || p->suffix_len > UQB_HDR_MAX) return uqb_fail(ctx, "decode: prefix/suffix too long");
    if (p->bits_per_base < 1 || p->bits_per_base > 8 || p->bits_per_quality < 1 || p->bits_per_quality > 8) return uqb_fail(ctx, "decode: bad bit widths");
    const uint64_t nsym = (uint64_t)p->dna_max + p->variable;
    if ((uint64_t)dna->width * 8 < nsym * p->bits_per_base || (uint64_t)qual->width * 8 < nsym * p->bits_per_quality)
        return uqb_fail(ctx, "decode: rows too narrow for dna_max");
    dec_params* hp = new dec_params();
    memset(hp, 0, sizeof(*hp));
    memcpy(hp->base_char, p->base_char, 256);
    memcpy(hp->qual_char, p->qual_char, 256);
    memcpy(hp->qual_to_base, p->qual_to_base, 512);
    hp->bb = p->bits_per_base; hp->bq = p->bits_per_quality; hp->variable = p->variable; hp->dna_max = p->dna_max;
    hp->wd = dna->width; hp->wq = qual->width;
    hp->prefix_len = p->prefix_len; hp->suffix_len = p->suffix_len; hp->nseps = p->nseps; hp->ncols = p->ncols;
    if (p->prefix_len) memcpy(hp->prefix, p->prefix, p->prefix_len);
    if (p->suffix_len) memcpy(hp->suffix, p->suffix, p->suffix_len);
    if (p->nseps) memcpy(hp->seps, p->seps, p->nseps);
    std::vector<void*> temps;
    std::vector<size_t> temp_sizes;
    int rc = 0;
    for (uint32_t c = 0; c < p->ncols && rc == 0; c++) {
        const uqb_decode_col& sc = p->cols[c];
        dec_col& dc = hp->cols[c];
        if (cols[c]->n != n || cols[c]->width != sc.itemsize) { rc = uqb_fail(ctx, "decode: column %u shape mismatch", c); break; }
        dc.data = (const uint8_t*)cols[c]->d;
        dc.itemsize = sc.itemsize; dc.format = sc.format; dc.offset = sc.offset; dc.min_val = sc.min_val;
        dc.dict_count = sc.dict_count; dc.dict_width = sc.dict_width;
        if (sc.format == 0) {
            const size_t nb = (size_t)sc.dict_count * sc.dict_width;
            std::vector<uint32_t> lens(sc.dict_count ? sc.dict_count : 1);
            for (uint64_t k = 0; k < sc.dict_count; k++) {
                uint32_t l = 0;
                while (l < sc.dict_width && sc.dict[k * sc.dict_width + l]) l++;
                lens[k] = l;
            }
            void *dd, *dl;
            if ((rc = uqb_dalloc(ctx, &dd, nb + 16))) break;
            temps.push_back(dd); temp_sizes.push_back(nb + 16);
            if ((rc = uqb_dalloc(ctx, &dl, lens.size() * 4))) break;
            temps.push_back(dl); temp_sizes.push_back(lens.size() * 4);
            if (nb) cudaMemcpyAsync(dd, sc.dict, nb, cudaMemcpyHostToDevice, ctx->stream);
            cudaMemcpyAsync(dl, lens.data(), lens.size() * 4, cudaMemcpyHostToDevice, ctx->stream);
            cudaStreamSynchronize(ctx->stream);
            dc.dict = (const uint8_t*)dd; dc.dict_len = (const uint32_t*)dl;
        }
    }
    dec_params* dp = nullptr;
    uint32_t* rec_len = nullptr;
    uint64_t* rec_off = nullptr;
    uint64_t* d_total = nullptr;
    uint64_t total = 0;
    if (rc == 0) rc = uqb_dalloc_t(ctx, &dp, 1);
    if (rc == 0) {
        cudaMemcpyAsync(dp, hp, sizeof(dec_params), cudaMemcpyHostToDevice, ctx->stream);
        cudaStreamSynchronize(ctx->stream);
    }
    delete hp;
    if (rc) return rc;
    UQB_TRY(uqb_dalloc_t(ctx, &rec_len, n));
    UQB_TRY(uqb_dalloc_t(ctx, &rec_off, n));
    UQB_TRY(uqb_dalloc_t(ctx, &d_total, 1));
    if (n) UQB_LAUNCH(k_decode_len, uqb_blocks(n, DC), DC, 0, dp, (const uint8_t*)dna->d, n, rec_len);
    UQB_TRY(uqb_scan_u32_to_u64(ctx, rec_len, rec_off, n, d_total));
    UQB_TRY(uqb_readback(ctx, &total, d_total, 8));
    UQB_TRY(uqb_new_array(ctx, total, 1, fastq));
    if (n) UQB_LAUNCH(k_decode_write, uqb_grid(ctx, n, DC / 32, 16), DC, 0, dp, (const uint8_t*)dna->d, (const uint8_t*)qual->d, n, rec_off, (uint8_t*)(*fastq)->d);
    UQB_TRY(uqb_dfree(ctx, rec_len, n * 4));
    UQB_TRY(uqb_dfree(ctx, rec_off, n * 8));
    UQB_TRY(uqb_dfree(ctx, d_total, 8));
    UQB_TRY(uqb_dfree(ctx, dp, sizeof(dec_params)));
    for (size_t i = 0; i < temps.size(); i++) UQB_TRY(uqb_dfree(ctx, temps[i], temp_sizes[i]));
    return 0;
}
