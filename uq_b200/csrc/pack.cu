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
#include "tile.cuh"

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
                                                         uint32_t wd, uint32_t wq, uint32_t variable, uint32_t dna_max, uint32_t prefilled,
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
        // Variable-length rows are right aligned: everything before the marker's byte is zero.  The host has zeroed
        // both tables with one memset, so only the significant bytes are produced here.
        const uint32_t jd0 = prefilled ? (wd * 8u - (len + variable) * bb) >> 3 : 0u;
        const uint32_t jq0 = prefilled ? (wq * 8u - (len + variable) * bq) >> 3 : 0u;
        for (uint32_t j = jd0 + lane; j < wd; j += 32) drow[j] = (uint8_t)pack_byte(dna, dna, false, &lut, bb, variable, len, wd, j);
        for (uint32_t j = jq0 + lane; j < wq; j += 32) qrow[j] = (uint8_t)pack_byte(qual, dna, true, &lut, bq, variable, len, wq, j);
    }
}

// ================================================================================================
// v2 (fixed-length reads): record tiles in shared memory (tile.cuh).  One work item = 16 consecutive
// positions of one read: the thread maps the 16 bases and qualities through the shared-memory LUTs,
// builds their bit strings most-significant-bit first and ORs them into the tile's output staging
// area, which holds the packed rows as big-endian 32-bit words.  The staging area is then written out
// with 128-bit stores: a tile of TL_R = 128 rows starts at a multiple of 16 bytes in both tables.
// ================================================================================================
#define PKT_THREADS 256
#define PKT_STAGE_MAX (64 * 1024)

struct pkt_lut { uint16_t base[256]; uint8_t qual[256]; };      // base: low byte = code, high byte = N-trick quality code or 0xFF

// append `nbits` (<= 32) bits, given right-aligned in `v`, at bit position `gbit` of a big-endian word array
__device__ __forceinline__ void or_bits(uint32_t* words, uint32_t gbit, uint32_t v, uint32_t nbits) {
    const uint32_t w = gbit >> 5, o = gbit & 31u;
    if (o + nbits <= 32u) {
        atomicOr(&words[w], v << (32u - o - nbits));
    } else {
        const uint32_t lo_bits = o + nbits - 32u;
        atomicOr(&words[w], v >> lo_bits);
        atomicOr(&words[w + 1], v << (32u - lo_bits));
    }
}

// 16 symbols of BITS bits -> a bit string (first symbol most significant) of NW words, ORed into the
// staging words at bit position gbit.  Everything but the final placement is compile-time indexed.
template <int BITS>
__device__ __forceinline__ void place_chunk16(const uint32_t (&code)[16], uint32_t* stage, uint32_t gbit) {
    constexpr int NW = (16 * BITS + 31) / 32;
    uint32_t w[NW + 1];
#pragma unroll
    for (int k = 0; k <= NW; k++) w[k] = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const int off = k * BITS, wi = off >> 5, o = off & 31;
        if (o + BITS <= 32) {
            w[wi] |= code[k] << (32 - o - BITS);
        } else {
            w[wi] |= code[k] >> (o + BITS - 32);
            w[wi + 1] |= code[k] << (64 - o - BITS);
        }
    }
    const uint32_t sh = gbit & 31u, w0 = gbit >> 5;
    uint32_t prev = 0;
#pragma unroll
    for (int k = 0; k < NW; k++) {
        atomicOr(&stage[w0 + k], __funnelshift_r(w[k], prev, sh));
        prev = w[k];
    }
    if (sh) {
        const uint32_t out = prev << (32u - sh);
        if (out) atomicOr(&stage[w0 + NW], out);
    }
}

__device__ __forceinline__ void place_chunk16_any(uint32_t bits, const uint32_t (&code)[16], uint32_t* stage, uint32_t gbit) {
    switch (bits) {          // warp-uniform
        case 1: place_chunk16<1>(code, stage, gbit); break;
        case 2: place_chunk16<2>(code, stage, gbit); break;
        case 3: place_chunk16<3>(code, stage, gbit); break;
        case 4: place_chunk16<4>(code, stage, gbit); break;
        case 5: place_chunk16<5>(code, stage, gbit); break;
        case 6: place_chunk16<6>(code, stage, gbit); break;
        case 7: place_chunk16<7>(code, stage, gbit); break;
        default: place_chunk16<8>(code, stage, gbit); break;
    }
}

// Long / variable-length reads: one CTA per read.  The significant part of both packed rows (from the marker's byte
// to the end of the row - everything before it is zero and was set by the host's memset) is assembled in shared
// memory with the same 16-symbol work items as the tile kernel, reading the symbols straight from global memory,
// and then written out.  Slot 0 of a variable-length row is the marker (code 1, uq.py:242-243).
__global__ void __launch_bounds__(PKT_THREADS) k_pack_long(const uint8_t* __restrict__ d, const uint64_t* __restrict__ line_off,
                                                          uint64_t n_reads, pack_lut lut_in, uint32_t bb, uint32_t bq, uint32_t wd, uint32_t wq,
                                                          uint32_t variable, uint32_t dna_max, uint8_t* __restrict__ dna_out,
                                                          uint8_t* __restrict__ qual_out, unsigned long long* __restrict__ err_record) {
    extern __shared__ __align__(16) uint8_t pkl_raw[];
    pkt_lut* lut = reinterpret_cast<pkt_lut*>(pkl_raw);
    uint32_t* stage_d = reinterpret_cast<uint32_t*>(pkl_raw + sizeof(pkt_lut));
    uint32_t* stage_q = stage_d + (wd + 3) / 4 + 8;
    const unsigned tid = threadIdx.x;
    for (unsigned i = tid; i < 256; i += PKT_THREADS) {
        const int t = lut_in.trick_qual[i];
        lut->base[i] = (uint16_t)(lut_in.base_code[i] | ((t >= 0 ? (unsigned)t : 0xFFu) << 8));
        lut->qual[i] = lut_in.qual_code[i];
    }
    for (uint64_t r = blockIdx.x; r < n_reads; r += gridDim.x) {
        const uint64_t o1 = line_off[4 * r + 1], o2 = line_off[4 * r + 2], o3 = line_off[4 * r + 3];
        uint64_t len64 = o2 - o1 - 1;
        uint32_t len = (uint32_t)len64;
        if (len64 > dna_max) {                              // cannot happen after uqb_analyze; never write out of row
            if (tid == 0) atomicMin(err_record, (unsigned long long)r);
            len = dna_max;
        }
        const uint8_t* dna = d + o1;
        const uint8_t* qual = d + o3;
        const uint32_t nsym = len + variable;
        const uint32_t pad_d = wd * 8u - nsym * bb, pad_q = wq * 8u - nsym * bq;       // bit position of slot 0 in its row
        const uint32_t jd0 = pad_d >> 3, jq0 = pad_q >> 3;                             // first significant byte
        const uint32_t wjd = jd0 & ~3u, wjq = jq0 & ~3u;                               // staging word 0 = row bytes wj .. wj + 3
        const uint32_t nwd = (wd - wjd + 3) / 4 + 8, nwq = (wq - wjq + 3) / 4 + 8;
        __syncthreads();                                     // the previous read has left the staging areas (and the LUT is set)
        for (uint32_t i = tid; i < nwd; i += PKT_THREADS) stage_d[i] = 0;
        for (uint32_t i = tid; i < nwq; i += PKT_THREADS) stage_q[i] = 0;
        __syncthreads();
        const uint32_t chunks = (nsym + 15) / 16;
        for (uint32_t item = tid; item < chunks; item += PKT_THREADS) {
            const uint32_t s0 = item * 16;
            uint32_t cdv[16], cqv[16];
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const uint32_t slot = s0 + k;
                uint32_t cd = 0, cq = 0;
                if (slot < nsym) {
                    if (variable && slot == 0) {
                        cd = 1; cq = 1;
                    } else {
                        const uint32_t i = slot - variable;
                        const uint32_t be = lut->base[__ldg(dna + i)];
                        const uint32_t tq = be >> 8;
                        cd = be & 0xFFu;
                        cq = tq != 0xFFu ? tq : lut->qual[__ldg(qual + i)];
                    }
                }
                cdv[k] = cd; cqv[k] = cq;
            }
            place_chunk16_any(bb, cdv, stage_d, pad_d - 8u * wjd + s0 * bb);
            place_chunk16_any(bq, cqv, stage_q, pad_q - 8u * wjq + s0 * bq);
        }
        __syncthreads();
        uint8_t* drow = dna_out + r * wd;
        uint8_t* qrow = qual_out + r * wq;
        for (uint32_t j = jd0 + tid; j < wd; j += PKT_THREADS) drow[j] = (uint8_t)(stage_d[(j - wjd) >> 2] >> (24u - 8u * ((j - wjd) & 3u)));
        for (uint32_t j = jq0 + tid; j < wq; j += PKT_THREADS) qrow[j] = (uint8_t)(stage_q[(j - wjq) >> 2] >> (24u - 8u * ((j - wjq) & 3u)));
    }
}

__global__ void __launch_bounds__(PKT_THREADS) k_pack_tiles(const uint8_t* __restrict__ d, uint64_t n_bytes,
                                                           const uint64_t* __restrict__ line_off, uint64_t n_reads, pack_lut lut_in,
                                                           uint32_t bb, uint32_t bq, uint32_t wd, uint32_t wq, uint32_t L,
                                                           uint8_t* __restrict__ dna_out, uint8_t* __restrict__ qual_out,
                                                           uint64_t* __restrict__ key_d, uint64_t* __restrict__ key_q,
                                                           unsigned int* __restrict__ fallback) {
    // key_d / key_q (optional): be64 of the first 8 bytes of every packed row = the round-0 key of the row sort, taken
    // from the staging area so that the sort does not have to re-read the strided rows
    extern __shared__ __align__(128) uint8_t pkt_raw[];
    tile_smem* T = reinterpret_cast<tile_smem*>(pkt_raw);
    pkt_lut* lut = reinterpret_cast<pkt_lut*>(pkt_raw + sizeof(tile_smem));
    uint32_t* stage_d = reinterpret_cast<uint32_t*>(pkt_raw + sizeof(tile_smem) + sizeof(pkt_lut));
    const uint32_t words_d = (TL_R * wd + 3) / 4 + 4, words_q = (TL_R * wq + 3) / 4 + 4;
    uint32_t* stage_q = stage_d + ((words_d + 3) & ~3u);
    const unsigned tid = threadIdx.x;
    for (unsigned i = tid; i < 256; i += PKT_THREADS) {
        const int t = lut_in.trick_qual[i];
        lut->base[i] = (uint16_t)(lut_in.base_code[i] | ((t >= 0 ? (unsigned)t : 0xFFu) << 8));
        lut->qual[i] = lut_in.qual_code[i];
    }
    tile_init(T);
    __syncthreads();
    const uint32_t pad_d = wd * 8u - L * bb, pad_q = wq * 8u - L * bq;
    const uint32_t fullc = L / 16, tailn = L % 16;
    const uint64_t ntiles = (n_reads + TL_R - 1) / TL_R;
    // item -> (record, chunk) without an integer division: i = item / fullc through a 32-bit reciprocal (exact for the
    // item counts of a tile: item < 2^16, 2 <= fullc < 2^10; fullc = 1 needs none)
    const uint32_t fullc_rcp = fullc ? (uint32_t)((0x100000000ull + fullc - 1) / fullc) : 0u;
    unsigned phase = 0;
    // The line offsets of the NEXT tile are loaded while this one is packed (three per thread, plus the byte bounds), so that
    // the tile load is one exposed round trip - the bulk copy - instead of two dependent ones.
    constexpr int OPT = (4 * TL_R + 1 + PKT_THREADS - 1) / PKT_THREADS;
    uint64_t pre_off[OPT], pre_b0 = 0, pre_b1 = 0;
    auto prefetch = [&](uint64_t t) {
        const uint64_t r0 = t * TL_R, r1 = (r0 + TL_R < n_reads) ? r0 + TL_R : n_reads;
        const uint32_t nrec = (uint32_t)(r1 - r0);
        pre_b0 = __ldg(line_off + 4 * r0); pre_b1 = __ldg(line_off + 4 * r1);
#pragma unroll
        for (int k = 0; k < OPT; k++) {
            const uint32_t i = tid + k * PKT_THREADS;
            pre_off[k] = i <= 4 * nrec ? __ldg(line_off + 4 * r0 + i) : 0ull;
        }
    };
    if (blockIdx.x < ntiles) prefetch(blockIdx.x);
    for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const uint64_t r0 = t * TL_R, r1 = (r0 + TL_R < n_reads) ? r0 + TL_R : n_reads;
        // ---- tile_load (tile.cuh) with the prefetched offsets ----
        __syncthreads();                                     // every thread is done with the previous tile
        const uint32_t nrec = (uint32_t)(r1 - r0);
        const uint64_t b0 = pre_b0, b1 = pre_b1, a0 = b0 & ~15ull;
        const bool fits = b1 - a0 <= TL_CAP;
        if (fits) {
            uint64_t a1 = (b1 + 15) & ~15ull;
            const uint64_t lim = n_bytes & ~15ull;
            if (a1 > lim) a1 = lim;
            if (a1 < a0) a1 = a0;
            const uint32_t bulk = (uint32_t)(a1 - a0);
            if (tid == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&T->bar, bulk);
                if (bulk) bulk_g2s(T->bytes, d + a0, bulk, &T->bar);
            }
#pragma unroll
            for (int k = 0; k < OPT; k++) {
                const uint32_t i = tid + k * PKT_THREADS;
                if (i <= 4 * nrec) T->loff[i] = (uint32_t)(pre_off[k] - a0);
            }
            for (uint64_t q = a1 + tid; q < b1; q += PKT_THREADS) T->bytes[q - a0] = d[q];
        }
        if (t + gridDim.x < ntiles) prefetch(t + gridDim.x);   // in flight during the wait and the whole tile
        if (!fits) { if (tid == 0) atomicOr(fallback, 1u); continue; }
        mbar_wait(&T->bar, phase);
        __syncthreads();
        phase ^= 1u;
        for (uint32_t i = tid; i < words_d; i += PKT_THREADS) stage_d[i] = 0;
        for (uint32_t i = tid; i < words_q; i += PKT_THREADS) stage_q[i] = 0;
        __syncthreads();
        // work items: one chunk of 16 positions of one record.  All full chunks come first, then the (shorter) last
        // chunks of the records, so that only the warps of that last group pay for masking the unused codes.
        const uint32_t items_full = nrec * fullc, items = items_full + (tailn ? nrec : 0u);
        for (uint32_t item = tid; item < items; item += PKT_THREADS) {
            uint32_t i, c, nsym;
            if (item < items_full) { i = fullc > 1u ? __umulhi(item, fullc_rcp) : item; c = item - i * fullc; nsym = 16; }
            else { i = item - items_full; c = fullc; nsym = tailn; }
            const uint32_t o1 = T->loff[4 * i + 1], o2 = T->loff[4 * i + 2], o3 = T->loff[4 * i + 3];
            if (o2 - o1 - 1 != L) { atomicOr(fallback, 2u); continue; }      // not a fixed-length file after all
            const uint32_t p0 = c * 16;
            const uint32_t gd = (i * wd) * 8u + pad_d + p0 * bb;      // bit position in the tile's DNA staging area
            const uint32_t gq = (i * wq) * 8u + pad_q + p0 * bq;
            // 2 x 16 bytes through aligned 32-bit shared loads + funnel shifts, fully unrolled.  A last chunk reads
            // past the end of its line (still inside the tile); those codes are zeroed, and OR-ing zero bits into
            // the staging area changes nothing.
            uint32_t xd[5], xq[5];
            const uint32_t ad = o1 + p0, aq = o3 + p0;
            const uint32_t* wdp = reinterpret_cast<const uint32_t*>(T->bytes + (ad & ~3u));
            const uint32_t* wqp = reinterpret_cast<const uint32_t*>(T->bytes + (aq & ~3u));
#pragma unroll
            for (int k = 0; k < 5; k++) { xd[k] = wdp[k]; xq[k] = wqp[k]; }
            const uint32_t sd = (ad & 3u) * 8u, sq = (aq & 3u) * 8u;
            uint32_t cdv[16], cqv[16];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t vd = __funnelshift_r(xd[k], xd[k + 1], sd), vq = __funnelshift_r(xq[k], xq[k + 1], sq);
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    // branch-free: both LUTs are always read, the N-trick code is a select
                    const uint32_t be = lut->base[__byte_perm(vd, 0, 0x4440 + b)];
                    const uint32_t ql = lut->qual[__byte_perm(vq, 0, 0x4440 + b)];
                    const uint32_t tq = be >> 8;
                    cdv[4 * k + b] = be & 0xFFu;
                    cqv[4 * k + b] = (tq != 0xFFu) ? tq : ql;
                }
            }
            if (nsym < 16u) {
#pragma unroll
                for (int k = 0; k < 16; k++) if ((uint32_t)k >= nsym) { cdv[k] = 0; cqv[k] = 0; }
            }
            place_chunk16_any(bb, cdv, stage_d, gd);
            place_chunk16_any(bq, cqv, stage_q, gq);
        }
        __syncthreads();
        if (tid < nrec) {
            if (key_d) {
                const uint32_t B = tid * wd, sh = (B & 3u) * 8u;
                const uint32_t w0 = stage_d[B >> 2], w1 = stage_d[(B >> 2) + 1], w2 = stage_d[(B >> 2) + 2];
                key_d[r0 + tid] = ((uint64_t)__funnelshift_l(w1, w0, sh) << 32) | __funnelshift_l(w2, w1, sh);
            }
            if (key_q) {
                const uint32_t B = tid * wq, sh = (B & 3u) * 8u;
                const uint32_t w0 = stage_q[B >> 2], w1 = stage_q[(B >> 2) + 1], w2 = stage_q[(B >> 2) + 2];
                key_q[r0 + tid] = ((uint64_t)__funnelshift_l(w1, w0, sh) << 32) | __funnelshift_l(w2, w1, sh);
            }
        }
        // staging (big-endian words) -> global bytes
        {
            const uint64_t out0 = r0 * wd, nb = (uint64_t)nrec * wd;
            const uint32_t vec = (uint32_t)(nb / 16);
            uint4* o = reinterpret_cast<uint4*>(dna_out + out0);
            for (uint32_t v = tid; v < vec; v += PKT_THREADS) {
                uint4 x;
                x.x = __byte_perm(stage_d[4 * v], 0, 0x0123); x.y = __byte_perm(stage_d[4 * v + 1], 0, 0x0123);
                x.z = __byte_perm(stage_d[4 * v + 2], 0, 0x0123); x.w = __byte_perm(stage_d[4 * v + 3], 0, 0x0123);
                o[v] = x;
            }
            for (uint32_t b = vec * 16 + tid; b < nb; b += PKT_THREADS) dna_out[out0 + b] = (uint8_t)(stage_d[b >> 2] >> (24 - 8 * (b & 3u)));
        }
        {
            const uint64_t out0 = r0 * wq, nb = (uint64_t)nrec * wq;
            const uint32_t vec = (uint32_t)(nb / 16);
            uint4* o = reinterpret_cast<uint4*>(qual_out + out0);
            for (uint32_t v = tid; v < vec; v += PKT_THREADS) {
                uint4 x;
                x.x = __byte_perm(stage_q[4 * v], 0, 0x0123); x.y = __byte_perm(stage_q[4 * v + 1], 0, 0x0123);
                x.z = __byte_perm(stage_q[4 * v + 2], 0, 0x0123); x.w = __byte_perm(stage_q[4 * v + 3], 0, 0x0123);
                o[v] = x;
            }
            for (uint32_t b = vec * 16 + tid; b < nb; b += PKT_THREADS) qual_out[out0 + b] = (uint8_t)(stage_q[b >> 2] >> (24 - 8 * (b & 3u)));
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
    // algorithmic bytes: base + quality bytes in, 4 line offsets per record in, both packed tables out
    const uint64_t abytes = 2 * fq->total_bases + 32 * N + N * ((uint64_t)p->dna_bytes + p->qual_bytes);
    // v2: shared-memory record tiles for fixed-length reads that fit
    const size_t stage_bytes = (size_t)((((TL_R * p->dna_bytes + 3) / 4 + 4 + 3) & ~3u) + (TL_R * p->qual_bytes + 3) / 4 + 4) * 4;
    bool trick_ok = true;                       // 0xFF marks "no trick" in the packed LUT
    for (int i = 0; i < 256; i++) if (p->trick_qual[i] >= 255) trick_ok = false;
    if (!p->variable && trick_ok && fq->n / N <= (TL_CAP - 64) / TL_R && stage_bytes <= PKT_STAGE_MAX && p->dna_max > 0) {
        unsigned int* d_fb;
        UQB_TRY(uqb_dalloc_t(ctx, &d_fb, 1));
        UQB_CUDA(cudaMemsetAsync(d_fb, 0, 4, ctx->stream));
        const size_t smem = sizeof(tile_smem) + sizeof(pkt_lut) + stage_bytes;
        UQB_CUDA(cudaFuncSetAttribute(k_pack_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const uint64_t ntiles = (N + TL_R - 1) / TL_R;
        const unsigned per_sm = (unsigned)(220 * 1024 / (smem + 1024)) ? (unsigned)(220 * 1024 / (smem + 1024)) : 1u;
        const uint64_t cap = (uint64_t)ctx->sm_count * per_sm;
        // round-0 sort keys of rows of 8+ bytes come for free here (16 bytes per record more to write)
        if (p->dna_bytes >= 8) UQB_TRY(uqb_dalloc_t(ctx, &(*dna)->key0, N));
        if (p->qual_bytes >= 8) UQB_TRY(uqb_dalloc_t(ctx, &(*qual)->key0, N));
        UQB_LAUNCH_B(abytes + 8 * N * (((*dna)->key0 ? 1 : 0) + ((*qual)->key0 ? 1 : 0)), k_pack_tiles, (unsigned)(ntiles < cap ? ntiles : cap), PKT_THREADS, smem,
                     fq->d, fq->n, fq->line_off, N, lut,
                     p->bits_per_base, p->bits_per_quality, p->dna_bytes, p->qual_bytes, p->dna_max,
                     (uint8_t*)(*dna)->d, (uint8_t*)(*qual)->d, (*dna)->key0, (*qual)->key0, d_fb);
        unsigned int fb = 0;
        UQB_TRY(uqb_readback(ctx, &fb, d_fb, 4));
        UQB_TRY(uqb_dfree(ctx, d_fb, 4));
        if (fb == 0) return 0;
        if ((*dna)->key0) { UQB_TRY(uqb_dfree(ctx, (*dna)->key0, N * 8)); (*dna)->key0 = nullptr; }
        if ((*qual)->key0) { UQB_TRY(uqb_dfree(ctx, (*qual)->key0, N * 8)); (*qual)->key0 = nullptr; }
        if (fb & 2u) return uqb_fail(ctx, "uqb_pack: a record is not dna_max long although variable == 0");
        // a tile did not fit: fall through to the direct-from-global kernel
    }
    unsigned long long* d_err;
    UQB_TRY(uqb_dalloc_t(ctx, &d_err, 1));
    UQB_CUDA(cudaMemsetAsync(d_err, 0xFF, 8, ctx->stream));
    const uint32_t prefilled = p->variable ? 1u : 0u;
    if (prefilled) {                                   // right-aligned rows: zero both tables at memset speed first
        UQB_CUDA(cudaMemsetAsync((*dna)->d, 0, N * (uint64_t)p->dna_bytes, ctx->stream));
        UQB_CUDA(cudaMemsetAsync((*qual)->d, 0, N * (uint64_t)p->qual_bytes, ctx->stream));
    }
    const size_t long_smem = sizeof(pkt_lut) + (((size_t)p->dna_bytes + 3) / 4 + 8 + ((size_t)p->qual_bytes + 3) / 4 + 8) * 4;
    if (trick_ok && long_smem <= 160 * 1024 && p->dna_max >= 64) {
        // rows assembled in shared memory, one CTA per read
        UQB_CUDA(cudaFuncSetAttribute(k_pack_long, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)long_smem));
        const unsigned per_sm = (unsigned)(200 * 1024 / (long_smem + 1024)) > 8u ? 8u : ((unsigned)(200 * 1024 / (long_smem + 1024)) ? (unsigned)(200 * 1024 / (long_smem + 1024)) : 1u);
        const uint64_t cap = (uint64_t)ctx->sm_count * per_sm;
        UQB_LAUNCH_B(abytes, k_pack_long, (unsigned)(N < cap ? N : cap), PKT_THREADS, long_smem, fq->d, fq->line_off, N, lut,
                     p->bits_per_base, p->bits_per_quality, p->dna_bytes, p->qual_bytes, p->variable, p->dna_max,
                     (uint8_t*)(*dna)->d, (uint8_t*)(*qual)->d, d_err);
    } else {
        UQB_LAUNCH_B(abytes, k_pack_rows, uqb_grid(ctx, N, PK_THREADS / 32, 16), PK_THREADS, 0, fq->d, fq->line_off, N, lut,
                   p->bits_per_base, p->bits_per_quality, p->dna_bytes, p->qual_bytes, p->variable, p->dna_max, prefilled,
                   (uint8_t*)(*dna)->d, (uint8_t*)(*qual)->d, d_err);
    }
    unsigned long long err;
    UQB_TRY(uqb_readback(ctx, &err, d_err, 8));
    UQB_TRY(uqb_dfree(ctx, d_err, 8));
    if (err != ~0ull) return uqb_fail(ctx, "uqb_pack: record %llu is longer than dna_max", err);
    return 0;
}
