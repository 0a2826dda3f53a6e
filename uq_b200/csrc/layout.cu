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

static int layout_impl(uqb_ctx* ctx, const uint8_t* src, uint8_t* dst, uint64_t n, uint32_t width, int pattern, bool inverse) {
    layout_desc ld;
    if (pattern_desc(pattern, &ld)) return uqb_fail(ctx, "layout: pattern id %d out of range", pattern);
    if (n == 0 || width == 0) return 0;
    if (!ld.transposed) {
        unsigned g = uqb_grid(ctx, n, LT / 32, 16);
        if (inverse) UQB_LAUNCH_B(2 * n * width, k_layout_rows<true>, g, LT, 0, src, dst, n, width, ld.rev_r, ld.rev_b);
        else         UQB_LAUNCH_B(2 * n * width, k_layout_rows<false>, g, LT, 0, src, dst, n, width, ld.rev_r, ld.rev_b);
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
