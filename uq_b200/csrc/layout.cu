// Stage 4: the eight --pattern physical layouts (write_pattern uq.py:257-270) and their inverse
// (load_from_tar uq.py:943-945).
//
// For the logical table m[N][B] the byte stream numpy.save emits is one of
//     row-major    stream[r' * B + b']     ids 0.1 (0), 1.2 (5), 3.2 (7), 2.1 (2)
//     column-major stream[b' * N + r']     ids 0.2 (4), 1.1 (1), 3.1 (3), 2.2 (6)
// with r' = r or N-1-r and b' = b or B-1-b.  Row-major forms are streaming copies (one warp per
// row); column-major forms are shared-memory tiled transposes (64 x 64 byte tiles, loads coalesced
// along b, stores coalesced along r).
#include "common.cuh"

#define LT 256
#define TILE 64

struct layout_desc { int transposed, rev_r, rev_b; };

static int pattern_desc(int pattern, layout_desc* d) {
    switch (pattern) {
        case 0: *d = {0, 0, 0}; return 0;   // 0.1
        case 5: *d = {0, 0, 1}; return 0;   // 1.2
        case 7: *d = {0, 1, 0}; return 0;   // 3.2
        case 2: *d = {0, 1, 1}; return 0;   // 2.1
        case 4: *d = {1, 0, 0}; return 0;   // 0.2
        case 1: *d = {1, 0, 1}; return 0;   // 1.1
        case 3: *d = {1, 1, 0}; return 0;   // 3.1
        case 6: *d = {1, 1, 1}; return 0;   // 2.2
    }
    return 1;
}

// INVERSE = false: stream <- table ; INVERSE = true: table <- stream
template <bool INVERSE>
__global__ void __launch_bounds__(LT) k_layout_rows(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, uint64_t n, uint32_t width,
                                                   int rev_r, int rev_b) {
    const unsigned lane = threadIdx.x & 31u;
    const uint64_t wstride = (uint64_t)gridDim.x * (LT / 32);
    for (uint64_t r = (uint64_t)blockIdx.x * (LT / 32) + (threadIdx.x >> 5); r < n; r += wstride) {
        const uint64_t rp = rev_r ? n - 1 - r : r;
        const uint8_t* s = src + (INVERSE ? rp : r) * width;
        uint8_t* d = dst + (INVERSE ? r : rp) * width;
        for (uint32_t b = lane; b < width; b += 32) {
            const uint32_t bp = rev_b ? width - 1 - b : b;
            if (INVERSE) d[b] = __ldg(s + bp); else d[bp] = __ldg(s + b);
        }
    }
}

template <bool INVERSE>
__global__ void __launch_bounds__(LT) k_layout_transpose(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, uint64_t n, uint32_t width,
                                                        int rev_r, int rev_b) {
    __shared__ uint8_t tile[TILE][TILE + 4];
    const uint64_t r0 = (uint64_t)blockIdx.x * TILE;
    const uint32_t b0 = blockIdx.y * TILE;
    // table side: j (byte within row) fastest; stream side: i (row) fastest
    for (unsigned t = threadIdx.x; t < TILE * TILE; t += LT) {
        unsigned i, j;
        if (!INVERSE) { i = t / TILE; j = t % TILE; } else { j = t / TILE; i = t % TILE; }
        const uint64_t r = r0 + i;
        const uint32_t b = b0 + j;
        if (r < n && b < width) {
            if (!INVERSE) {
                tile[i][j] = __ldg(src + r * width + b);
            } else {
                const uint64_t rp = rev_r ? n - 1 - r : r;
                const uint32_t bp = rev_b ? width - 1 - b : b;
                tile[i][j] = __ldg(src + (uint64_t)bp * n + rp);
            }
        }
    }
    __syncthreads();
    for (unsigned t = threadIdx.x; t < TILE * TILE; t += LT) {
        unsigned i, j;
        if (!INVERSE) { j = t / TILE; i = t % TILE; } else { i = t / TILE; j = t % TILE; }
        const uint64_t r = r0 + i;
        const uint32_t b = b0 + j;
        if (r < n && b < width) {
            if (!INVERSE) {
                const uint64_t rp = rev_r ? n - 1 - r : r;
                const uint32_t bp = rev_b ? width - 1 - b : b;
                dst[(uint64_t)bp * n + rp] = tile[i][j];
            } else {
                dst[r * width + b] = tile[i][j];
            }
        }
    }
}

// Column-major forms, tall tiles: a CTA takes LR_ROWS consecutive rows of ALL columns.  On the table side that is one
// contiguous, 16-byte aligned range (vector loads / stores through shared memory); on the stream side it is one run
// of LR_ROWS bytes per column, moved as 32-bit words.  The 64 x 64 tiles above write 64-byte runs scattered over
// `width` streams, which DRAM does not like; here a run is 512 bytes.
#define LR_ROWS 512
#define LR_SMEM_MAX (96 * 1024)

template <bool INVERSE>
__global__ void __launch_bounds__(LT) k_layout_transpose_tall(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, uint64_t n,
                                                             uint32_t width, int rev_r, int rev_b) {
    extern __shared__ __align__(16) uint8_t lr_tile[];                 // [rows][width] as in the table
    const uint64_t nblk = (n + LR_ROWS - 1) / LR_ROWS;
    for (uint64_t blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        const uint64_t r0 = blk * LR_ROWS;
        const uint32_t nr = (uint32_t)(n - r0 < LR_ROWS ? n - r0 : LR_ROWS);
        const uint32_t total = nr * width, q = (nr + 3) / 4;
        const uint64_t lo = rev_r ? n - r0 - nr : r0;                  // first stream index (within a column) of this tile
        const uint8_t* table_c = INVERSE ? nullptr : src + r0 * width;
        uint8_t* table_m = INVERSE ? dst + r0 * width : nullptr;
        __syncthreads();
        if (!INVERSE) {
            for (uint32_t v = threadIdx.x; v < (total + 15) / 16; v += LT)   // the table has 64 bytes of slack behind its last row
                reinterpret_cast<uint4*>(lr_tile)[v] = __ldg(reinterpret_cast<const uint4*>(table_c) + v);
            __syncthreads();
        }
        // stream side: work item = 4 consecutive bytes of one column's run
        for (uint32_t t = threadIdx.x; t < width * q; t += LT) {
            const uint32_t b = t / q, i4 = (t - b * q) * 4;
            const uint32_t bp = rev_b ? width - 1 - b : b;
            const uint64_t a = (uint64_t)bp * n + lo + i4;             // stream address of the first of the 4 bytes
            const uint32_t cnt = nr - i4 < 4 ? nr - i4 : 4;
            if (!INVERSE) {
                uint32_t w = 0;
#pragma unroll
                for (uint32_t k = 0; k < 4; k++) {
                    if (k < cnt) {
                        const uint32_t i = rev_r ? nr - 1 - (i4 + k) : i4 + k;       // tile row of stream byte i4 + k
                        w |= (uint32_t)lr_tile[i * width + b] << (8 * k);
                    }
                }
                if (cnt == 4 && (a & 3) == 0) *reinterpret_cast<uint32_t*>(dst + a) = w;
                else for (uint32_t k = 0; k < cnt; k++) dst[a + k] = (uint8_t)(w >> (8 * k));
            } else {
                uint32_t w = 0;
                if (cnt == 4 && (a & 3) == 0) w = __ldg(reinterpret_cast<const uint32_t*>(src + a));
                else for (uint32_t k = 0; k < cnt; k++) w |= (uint32_t)__ldg(src + a + k) << (8 * k);
#pragma unroll
                for (uint32_t k = 0; k < 4; k++) {
                    if (k < cnt) {
                        const uint32_t i = rev_r ? nr - 1 - (i4 + k) : i4 + k;
                        lr_tile[i * width + b] = (uint8_t)(w >> (8 * k));
                    }
                }
            }
        }
        if (INVERSE) {
            __syncthreads();
            for (uint32_t v = threadIdx.x; v < total / 16; v += LT)
                reinterpret_cast<uint4*>(table_m)[v] = reinterpret_cast<const uint4*>(lr_tile)[v];
            for (uint32_t x = (total & ~15u) + threadIdx.x; x < total; x += LT) table_m[x] = lr_tile[x];
        }
    }
}

// Row-major forms with reversed rows and / or bytes, tall tiles: LR_ROWS consecutive rows are one contiguous range on
// both sides (the rows of the block just appear in reverse order), so the block is read with 16-byte loads into shared
// memory and written as 32-bit words.  Reversing is its own inverse: the same kernel serves layout and unlayout.
__global__ void __launch_bounds__(LT) k_layout_rows_tall(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, uint64_t n, uint32_t width,
                                                        int rev_r, int rev_b) {
    extern __shared__ __align__(16) uint8_t lr_tile[];
    const uint64_t nblk = (n + LR_ROWS - 1) / LR_ROWS;
    for (uint64_t blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        const uint64_t r0 = blk * LR_ROWS;
        const uint32_t nr = (uint32_t)(n - r0 < LR_ROWS ? n - r0 : LR_ROWS);
        const uint32_t total = nr * width;
        __syncthreads();
        for (uint32_t v = threadIdx.x; v < (total + 15) / 16; v += LT)        // the source has 64 bytes of slack behind its last row
            reinterpret_cast<uint4*>(lr_tile)[v] = __ldg(reinterpret_cast<const uint4*>(src + r0 * width) + v);
        __syncthreads();
        uint8_t* out = dst + (rev_r ? n - r0 - nr : r0) * width;              // the block's place on the other side
        const uint32_t head = (uint32_t)((4u - ((uintptr_t)out & 3u)) & 3u);   // bytes before the first aligned word
        const uint32_t nwords = total > head ? (total - head) / 4 : 0;
        for (uint32_t x = threadIdx.x; x < nwords + 8; x += LT) {
            // items 0 .. nwords-1: aligned words; the last 8 items: the (at most 3 + 3) edge bytes
            uint32_t first, cnt;
            if (x < nwords) { first = head + 4 * x; cnt = 4; }
            else { const uint32_t e = x - nwords; first = e < 4 ? e : head + 4 * nwords + (e - 4); cnt = (e < 4 ? (e < head && e < total) : (first < total && first >= head)) ? 1u : 0u; }
            uint32_t w = 0;
            uint32_t io = first / width, bo = first - io * width;             // output row / byte of the first byte
            for (uint32_t k = 0; k < cnt; k++) {
                const uint32_t i = rev_r ? nr - 1 - io : io, b = rev_b ? width - 1 - bo : bo;
                w |= (uint32_t)lr_tile[i * width + b] << (8 * k);
                if (++bo == width) { bo = 0; io++; }
            }
            if (cnt == 4) *reinterpret_cast<uint32_t*>(out + first) = w;
            else if (cnt == 1) out[first] = (uint8_t)w;
        }
    }
}

// ---- wide rows (long reads: 5 - 17 kB per row) ------------------------------------------------------
// Column-major forms, panels: a CTA takes PN_R rows x PN_C columns.  The panel lives in shared memory row by row with a
// pitch of PN_C + 1 BYTES: rows 4 k (k = lane) then fall into different banks, so the stream side can collect the four
// bytes of a 32-bit word of one column's run (rows 4 k .. 4 k + 3) without bank conflicts, and the table side writes a
// row's bytes with conflict-free byte stores.  Table side: one coalesced 128-byte segment per row (aligned 32-bit
// loads + funnel shift, rows start at any byte); stream side: one run of PN_R bytes per column, moved as aligned
// 32-bit words (the first / last bytes of a run one by one, runs start at any byte because n is arbitrary).
#define PN_R 256
#define PN_C 128
#define PN_P (PN_C + 1)

template <bool INVERSE>
__global__ void __launch_bounds__(LT) k_layout_transpose_panel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, uint64_t n,
                                                              uint32_t width, int rev_r, int rev_b) {
    __shared__ uint8_t pan[PN_R * PN_P + 8];
    const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    const uint32_t npc = (width + PN_C - 1) / PN_C;
    const uint64_t npr = (n + PN_R - 1) / PN_R, npan = npr * npc;
    for (uint64_t pidx = blockIdx.x; pidx < npan; pidx += gridDim.x) {
        const uint64_t pr = pidx / npc;
        const uint32_t pc = (uint32_t)(pidx - pr * npc);
        const uint64_t r0 = pr * PN_R;
        const uint32_t b0 = pc * PN_C;
        const uint32_t nr = (uint32_t)(n - r0 < PN_R ? n - r0 : PN_R), nc = width - b0 < PN_C ? width - b0 : PN_C;
        const uint64_t lo = rev_r ? n - r0 - nr : r0;                  // first stream index (within a column) of this panel
        __syncthreads();
        if (!INVERSE) {
            // table -> panel: warp per row, lane l takes bytes 4 l .. 4 l + 3 of the row's segment
            for (uint32_t i = w; i < nr; i += LT / 32) {
                if (4u * lane < nc) {
                    const uint64_t x = (uint64_t)(uintptr_t)src + (r0 + i) * width + b0 + 4u * lane;
                    const uint32_t ph = (uint32_t)x & 3u;
                    const uint32_t* a = reinterpret_cast<const uint32_t*>(x - ph);
                    const uint32_t v = __funnelshift_r(__ldg(a), __ldg(a + 1), ph * 8u);     // may read up to 3 bytes past the row: table slack
                    uint8_t* p = pan + i * PN_P + 4u * lane;
                    p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24);
                }
            }
            __syncthreads();
        }
        // stream side: warp per column run
        for (uint32_t c = w; c < nc; c += LT / 32) {
            const uint32_t b = b0 + c, bp = rev_b ? width - 1 - b : b;
            const uint64_t a0 = (uint64_t)bp * n + lo;                 // stream address of the run
            const uint8_t* sp = src + a0;
            uint8_t* dp = dst + a0;
            const uint32_t head = min(nr, (uint32_t)((4u - ((uintptr_t)(INVERSE ? (const void*)sp : (const void*)dp) & 3u)) & 3u));
            const uint32_t nwords = (nr - head) >> 2, tail0 = head + 4u * nwords;
            // stream offset t <-> panel row: t (or nr - 1 - t when the rows are reversed)
            if (!INVERSE) {
                if (lane < head) dp[lane] = pan[(rev_r ? nr - 1 - lane : lane) * PN_P + c];
                for (uint32_t q = lane; q < nwords; q += 32) {
                    const uint32_t t = head + 4u * q;
                    uint32_t v = 0;
#pragma unroll
                    for (uint32_t j = 0; j < 4; j++) v |= (uint32_t)pan[(rev_r ? nr - 1 - (t + j) : t + j) * PN_P + c] << (8 * j);
                    reinterpret_cast<uint32_t*>(dp + head)[q] = v;
                }
                if (lane < nr - tail0) dp[tail0 + lane] = pan[(rev_r ? nr - 1 - (tail0 + lane) : tail0 + lane) * PN_P + c];
            } else {
                if (lane < head) pan[(rev_r ? nr - 1 - lane : lane) * PN_P + c] = __ldg(sp + lane);
                for (uint32_t q = lane; q < nwords; q += 32) {
                    const uint32_t t = head + 4u * q;
                    const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(sp + head) + q);
#pragma unroll
                    for (uint32_t j = 0; j < 4; j++) pan[(rev_r ? nr - 1 - (t + j) : t + j) * PN_P + c] = (uint8_t)(v >> (8 * j));
                }
                if (lane < nr - tail0) pan[(rev_r ? nr - 1 - (tail0 + lane) : tail0 + lane) * PN_P + c] = __ldg(sp + tail0 + lane);
            }
        }
        if (INVERSE) {
            __syncthreads();
            // panel -> table: warp per row, byte stores (a row's segment starts at any byte)
            for (uint32_t i = w; i < nr; i += LT / 32) {
                uint8_t* o = dst + (r0 + i) * width + b0;
                const uint8_t* p = pan + i * PN_P;
                for (uint32_t x = lane; x < nc; x += 32) o[x] = p[x];
            }
        }
    }
}

// Row-major forms with reversed rows and / or bytes, any width: warp per row; the output row is written as aligned 32-bit
// words (its first / last bytes one by one), each taken from the source row with two aligned loads and a funnel shift
// (byte-swapped when the bytes are reversed).  Reversing is its own inverse.
__global__ void __launch_bounds__(LT) k_layout_rows_wide(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, uint64_t n, uint32_t width,
                                                        int rev_r, int rev_b) {
    const unsigned lane = threadIdx.x & 31u;
    const uint64_t wstride = (uint64_t)gridDim.x * (LT / 32);
    for (uint64_t r = (uint64_t)blockIdx.x * (LT / 32) + (threadIdx.x >> 5); r < n; r += wstride) {
        const uint8_t* s = src + r * width;
        uint8_t* d = dst + (rev_r ? n - 1 - r : r) * width;
        const uint32_t head = min(width, (uint32_t)((4u - ((uintptr_t)d & 3u)) & 3u));
        const uint32_t nwords = (width - head) >> 2, tail0 = head + 4u * nwords;
        if (lane < head) d[lane] = __ldg(s + (rev_b ? width - 1 - lane : lane));
#pragma unroll 4
        for (uint32_t q = lane; q < nwords; q += 32) {
            const uint32_t t = head + 4u * q;                          // output bytes t .. t + 3
            const uint32_t so = rev_b ? width - 4u - t : t;            // they come from source bytes so .. so + 3 (reversed when rev_b)
            const uint64_t x = (uint64_t)(uintptr_t)s + so;
            const uint32_t ph = (uint32_t)x & 3u;
            const uint32_t* a = reinterpret_cast<const uint32_t*>(x - ph);
            uint32_t v = __funnelshift_r(__ldg(a), ph ? __ldg(a + 1) : 0u, ph * 8u);
            if (rev_b) v = __byte_perm(v, 0u, 0x0123);
            reinterpret_cast<uint32_t*>(d + head)[q] = v;
        }
        if (lane < width - tail0) d[tail0 + lane] = __ldg(s + (rev_b ? width - 1 - (tail0 + lane) : tail0 + lane));
    }
}

static int layout_impl(uqb_ctx* ctx, const uint8_t* src, uint8_t* dst, uint64_t n, uint32_t width, int pattern, bool inverse) {
    layout_desc ld;
    if (pattern_desc(pattern, &ld)) return uqb_fail(ctx, "layout: pattern id %d out of range", pattern);
    if (n == 0 || width == 0) return 0;
    if (!ld.transposed && (size_t)LR_ROWS * width + 16 <= LR_SMEM_MAX && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const size_t smem = (size_t)LR_ROWS * width + 16;
        UQB_CUDA(cudaFuncSetAttribute(k_layout_rows_tall, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        UQB_LAUNCH_B(2 * n * width, k_layout_rows_tall, uqb_grid(ctx, n, LR_ROWS, 16), LT, smem, src, dst, n, width, ld.rev_r, ld.rev_b);
    } else if (!ld.transposed && width >= 64) {
        UQB_LAUNCH_B(2 * n * width, k_layout_rows_wide, uqb_grid(ctx, n, LT / 32, 16), LT, 0, src, dst, n, width, ld.rev_r, ld.rev_b);
    } else if (!ld.transposed) {
        unsigned g = uqb_grid(ctx, n, LT / 32, 16);
        if (inverse) UQB_LAUNCH_B(2 * n * width, k_layout_rows<true>, g, LT, 0, src, dst, n, width, ld.rev_r, ld.rev_b);
        else         UQB_LAUNCH_B(2 * n * width, k_layout_rows<false>, g, LT, 0, src, dst, n, width, ld.rev_r, ld.rev_b);
    } else if ((size_t)LR_ROWS * width + 16 <= LR_SMEM_MAX && ((reinterpret_cast<uintptr_t>(inverse ? dst : src)) & 15) == 0) {
        const size_t smem = (size_t)LR_ROWS * width + 16;
        const unsigned g = uqb_grid(ctx, n, LR_ROWS, 16);
        if (inverse) {
            UQB_CUDA(cudaFuncSetAttribute(k_layout_transpose_tall<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            UQB_LAUNCH_B(2 * n * width, k_layout_transpose_tall<true>, g, LT, smem, src, dst, n, width, ld.rev_r, ld.rev_b);
        } else {
            UQB_CUDA(cudaFuncSetAttribute(k_layout_transpose_tall<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            UQB_LAUNCH_B(2 * n * width, k_layout_transpose_tall<false>, g, LT, smem, src, dst, n, width, ld.rev_r, ld.rev_b);
        }
    } else if (width >= 64) {
        const uint64_t npan = ((n + PN_R - 1) / PN_R) * ((width + PN_C - 1) / PN_C);
        const uint64_t cap = (uint64_t)ctx->sm_count * 6;
        const unsigned g = (unsigned)(npan < cap ? npan : cap);
        if (inverse) UQB_LAUNCH_B(2 * n * width, k_layout_transpose_panel<true>, g, LT, 0, src, dst, n, width, ld.rev_r, ld.rev_b);
        else         UQB_LAUNCH_B(2 * n * width, k_layout_transpose_panel<false>, g, LT, 0, src, dst, n, width, ld.rev_r, ld.rev_b);
    } else {
        uint64_t gx = (n + TILE - 1) / TILE;
        if (gx > 0x7fffffffull) return uqb_fail(ctx, "layout: too many rows");
        dim3 grid((unsigned)gx, (width + TILE - 1) / TILE);
        if (grid.y > 65535) return uqb_fail(ctx, "layout: rows wider than %d bytes are not supported", 65535 * TILE);
        if (inverse) UQB_LAUNCH_B(2 * n * width, k_layout_transpose<true>, grid, LT, 0, src, dst, n, width, ld.rev_r, ld.rev_b);
        else         UQB_LAUNCH_B(2 * n * width, k_layout_transpose<false>, grid, LT, 0, src, dst, n, width, ld.rev_r, ld.rev_b);
    }
    return 0;
}

extern "C" int uqb_layout(uqb_ctx* ctx, const uqb_array* table, int pattern, uqb_array** stream) {
    UQB_TRY(uqb_new_array(ctx, table->n * table->width, 1, stream));
    return layout_impl(ctx, (const uint8_t*)table->d, (uint8_t*)(*stream)->d, table->n, table->width, pattern, false);
}

extern "C" int uqb_unlayout(uqb_ctx* ctx, const uqb_array* stream, uint64_t n, uint32_t width, int pattern, uqb_array** table) {
    if (stream->nbytes() != n * width) return uqb_fail(ctx, "unlayout: stream holds %llu bytes, expected %llu", (unsigned long long)stream->nbytes(), (unsigned long long)(n * width));
    UQB_TRY(uqb_new_array(ctx, n, width, table));
    return layout_impl(ctx, (const uint8_t*)stream->d, (uint8_t*)(*table)->d, n, width, pattern, true);
}
