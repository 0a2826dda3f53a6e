// Stage 2: symbol mapping + bit packing of DNA and QUAL rows.
// Replaces encoder_fixed (uq.py:108-182) and encoder_variable (uq.py:188-254).
//
// Closed form: with nsym = len + variable symbols of `bits` bits (for variable
// length files symbol 0 is the marker, code 1), a row is the symbols concatenated most significant
// first and right-aligned in `width` bytes; everything in front is zero.
//
// v1 mapping: one warp per read; lanes walk the Bd + Bq output bytes of the read, each output byte
// assembles the (at most 8/bits + 1) symbols that overlap it.  The symbol LUTs live in shared
// memory.  Output rows are contiguous, so a warp stores runs of consecutive bytes.
#include "common.cuh"

#define PK_THREADS 256

struct pack_lut {
    uint8_t base_code[256];
    uint8_t qual_code[256];
    int16_t trick_qual[256];
};

__device__ __forceinline__ unsigned pack_byte(const uint8_t* __restrict__ sym_src, const uint8_t* __restrict__ base_src, bool is_qual,
                                              const pack_lut* lut, uint32_t bits, uint32_t variable, uint32_t len,
                                              uint32_t width, uint32_t j) {
    // bit coordinates: position 0 is the MSB of row byte 0 (all row-local, 32-bit arithmetic)
    const uint32_t nsym = len + variable;
    const uint32_t total_bits = nsym * bits;
    const uint32_t pad = width * 8u - total_bits;
    const uint32_t lo = j * 8u, hi = lo + 7u;
    if (hi < pad) return 0u;
    uint32_t s0 = lo > pad ? (lo - pad) / bits : 0u;
    uint32_t s1 = (hi - pad) / bits;
    if (s1 >= nsym) s1 = nsym - 1;
    unsigned v = 0;
    for (uint32_t s = s0; s <= s1; s++) {
        unsigned code;
        if (variable && s == 0) {
            code = 1u;                                   // the marker (uq.py:242-243)
        } else {
            const uint32_t i = s - variable;
            const unsigned b = __ldg(base_src + i);
            if (is_qual) {
                const int t = lut->trick_qual[b];        // N_qual[base] for tricked bases (uq.py:153)
                code = t >= 0 ? (unsigned)t : lut->qual_code[__ldg(sym_src + i)];
            } else {
                code = lut->base_code[b];                // tricked bases -> 0 (uq.py:152)
            }
        }
        const int rel = (int)(pad + s * bits) - (int)lo;                  // symbol MSB relative to byte MSB
        const int sh = 8 - rel - (int)bits;
        v |= sh >= 0 ? (code << sh) : (code >> (-sh));
    }
    return v & 255u;
}

__global__ void __launch_bounds__(PK_THREADS) k_pack_rows(const uint8_t* __restrict__ d, const uint64_t* __restrict__ line_off,
                                                         uint64_t n_reads, pack_lut lut_in, uint32_t bb, uint32_t bq,
                                                         uint32_t wd, uint32_t wq, uint32_t variable, uint32_t dna_max,
                                                         uint8_t* __restrict__ dna_out, uint8_t* __restrict__ qual_out,
                                                         unsigned long long* __restrict__ err_record) {
    __shared__ pack_lut lut;
    for (unsigned i = threadIdx.x; i < 256; i += PK_THREADS) {
        lut.base_code[i] = lut_in.base_code[i];
        lut.qual_code[i] = lut_in.qual_code[i];
        lut.trick_qual[i] = lut_in.trick_qual[i];
    }
    __syncthreads();
    const unsigned lane = threadIdx.x & 31u;
    const uint64_t wstride = (uint64_t)gridDim.x * (PK_THREADS / 32);
    for (uint64_t r = (uint64_t)blockIdx.x * (PK_THREADS / 32) + (threadIdx.x >> 5); r < n_reads; r += wstride) {
        const uint64_t o1 = line_off[4 * r + 1], o2 = line_off[4 * r + 2], o3 = line_off[4 * r + 3];
        uint64_t len64 = o2 - o1 - 1;
        uint32_t len = (uint32_t)len64;
        if (len64 > dna_max) {                              // cannot happen after uqb_analyze; never write out of row
            if (lane == 0) atomicMin(err_record, (unsigned long long)r);
            len = dna_max;
        }
        const uint8_t* dna = d + o1;
        const uint8_t* qual = d + o3;
        uint8_t* drow = dna_out + r * wd;
        uint8_t* qrow = qual_out + r * wq;
        for (uint32_t j = lane; j < wd + wq; j += 32) {
            if (j < wd) drow[j] = (uint8_t)pack_byte(dna, dna, false, &lut, bb, variable, len, wd, j);
            else        qrow[j - wd] = (uint8_t)pack_byte(qual, dna, true, &lut, bq, variable, len, wq, j - wd);
        }
    }
}

extern "C" int uqb_pack(uqb_ctx* ctx, uqb_fastq* fq, const uqb_pack_params* p, uqb_array** dna, uqb_array** qual) {
    if (!fq->line_off) return uqb_fail(ctx, "uqb_pack: call uqb_split first");
    if (p->bits_per_base < 1 || p->bits_per_base > 8 || p->bits_per_quality < 1 || p->bits_per_quality > 8)
        return uqb_fail(ctx, "uqb_pack: bits per symbol must be 1..8");
    const uint64_t need_d = ((uint64_t)p->bits_per_base * (p->dna_max + p->variable) + 7) / 8;
    const uint64_t need_q = ((uint64_t)p->bits_per_quality * (p->dna_max + p->variable) + 7) / 8;
    if (p->dna_bytes < need_d || p->qual_bytes < need_q) return uqb_fail(ctx, "uqb_pack: row widths too small for dna_max");
    for (int i = 0; i < 256; i++) {
        if (p->base_code[i] >> p->bits_per_base) return uqb_fail(ctx, "uqb_pack: base code %d does not fit %u bits", p->base_code[i], p->bits_per_base);
        if (p->qual_code[i] >> p->bits_per_quality) return uqb_fail(ctx, "uqb_pack: quality code %d does not fit %u bits", p->qual_code[i], p->bits_per_quality);
        // codes are ADDED by the reference, a wider code would carry into the neighbouring symbol (Q4)
        if (p->trick_qual[i] >= 0 && (p->trick_qual[i] >> p->bits_per_quality))
            return uqb_fail(ctx, "uqb_pack: N-trick quality code %d does not fit %u bits (reference quirk Q4)", p->trick_qual[i], p->bits_per_quality);
    }
    const uint64_t N = fq->n_reads;
    UQB_TRY(uqb_new_array(ctx, N, p->dna_bytes, dna));
    UQB_TRY(uqb_new_array(ctx, N, p->qual_bytes, qual));
    if (N == 0) return 0;
    pack_lut lut;
    memcpy(lut.base_code, p->base_code, 256);
    memcpy(lut.qual_code, p->qual_code, 256);
    memcpy(lut.trick_qual, p->trick_qual, 512);
    unsigned long long* d_err;
    UQB_TRY(uqb_dalloc_t(ctx, &d_err, 1));
    UQB_CUDA(cudaMemsetAsync(d_err, 0xFF, 8, ctx->stream));
    // algorithmic bytes: base + quality bytes in, 4 line offsets per record in, both packed tables out
    const uint64_t abytes = 2 * fq->total_bases + 32 * N + N * ((uint64_t)p->dna_bytes + p->qual_bytes);
    UQB_LAUNCH_B(abytes, k_pack_rows, uqb_grid(ctx, N, PK_THREADS / 32, 16), PK_THREADS, 0, fq->d, fq->line_off, N, lut,
               p->bits_per_base, p->bits_per_quality, p->dna_bytes, p->qual_bytes, p->variable, p->dna_max,
               (uint8_t*)(*dna)->d, (uint8_t*)(*qual)->d, d_err);
    unsigned long long err;
    UQB_TRY(uqb_readback(ctx, &err, d_err, 8));
    UQB_TRY(uqb_dfree(ctx, d_err, 8));
    if (err != ~0ull) return uqb_fail(ctx, "uqb_pack: record %llu is longer than dna_max", err);
    return 0;
}
