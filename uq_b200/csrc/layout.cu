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

static int layout_impl(uqb_ctx* ctx, const uint8_t* src, uint8_t* dst, uint64_t n, uint32_t width, int pattern, bool inverse) {
    layout_desc ld;
    if (pattern_desc(pattern, &ld)) return uqb_fail(ctx, "layout: pattern id %d out of range", pattern);
    if (n == 0 || width == 0) return 0;
    if (!ld.transposed && (size_t)LR_ROWS * width + 16 <= LR_SMEM_MAX && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const size_t smem = (size_t)LR_ROWS * width + 16;
        UQB_CUDA(cudaFuncSetAttribute(k_layout_rows_tall, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        UQB_LAUNCH_B(2 * n * width, k_layout_rows_tall, uqb_grid(ctx, n, LR_ROWS, 16), LT, smem, src, dst, n, width, ld.rev_r, ld.rev_b);
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
