// Stage 1a: record splitting.  Newline scan over the raw FASTQ byte stream -> uint64 line offsets.
// Replaces the four `for line in file` sweeps of the reference (uq.py:132-139, 205-211, 378-385,
// 563-569) and its `wc -l` record count (uq.py:85-87).
//
// HBM traffic: the byte stream is read twice (count, then write) with 128-bit loads; line
// offsets (8 B per line) are written once.  Each thread owns 64 consecutive bytes so that the
// rank of a newline inside the tile follows byte order.
#include "common.cuh"

#define SP_THREADS 256
#define SP_BYTES_PER_THREAD 64
#define SP_TILE (SP_THREADS * SP_BYTES_PER_THREAD)
#define SP_STAGE 1024          // line offsets staged per tile (a 16 KB tile of FASTQ holds ~200)

__device__ __forceinline__ unsigned nl_mask(unsigned w) { return __vcmpeq4(w, 0x0A0A0A0Au); }   // 0xFF per '\n' byte

// load the 16 words of this thread's 64-byte chunk; bytes at or beyond `n` read as zero
__device__ __forceinline__ void load_chunk(const uint8_t* __restrict__ d, uint64_t n, uint64_t pos, unsigned (&w)[16]) {
    if (pos + SP_BYTES_PER_THREAD <= n) {
        const uint4* p = reinterpret_cast<const uint4*>(d + pos);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            uint4 q = __ldg(p + i);
            w[4 * i] = q.x; w[4 * i + 1] = q.y; w[4 * i + 2] = q.z; w[4 * i + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            unsigned x = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                uint64_t q = pos + 4 * i + b;
                if (q < n) x |= (unsigned)d[q] << (8 * b);
            }
            w[i] = x;
        }
    }
}

__global__ void __launch_bounds__(SP_THREADS) k_newline_count(const uint8_t* __restrict__ d, uint64_t n, uint64_t tile0, uint32_t* __restrict__ tile_counts) {
    uint64_t pos = (tile0 + blockIdx.x) * SP_TILE + (uint64_t)threadIdx.x * SP_BYTES_PER_THREAD;
    unsigned c = 0;
    if (pos < n) {
        unsigned w[16];
        load_chunk(d, n, pos, w);
#pragma unroll
        for (int i = 0; i < 16; i++) c += __popc(nl_mask(w[i]));
        c >>= 3;
    }
    // block reduce
    __shared__ unsigned ws[SP_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane_id() == 0) ws[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
#pragma unroll
        for (int i = 0; i < SP_THREADS / 32; i++) t += ws[i];
        tile_counts[tile0 + blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(SP_THREADS) k_newline_write(const uint8_t* __restrict__ d, uint64_t n, uint64_t tile0, uint64_t line_base,
                                                            const uint64_t* __restrict__ tile_base, uint64_t* __restrict__ line_off) {
    uint64_t pos = (tile0 + blockIdx.x) * SP_TILE + (uint64_t)threadIdx.x * SP_BYTES_PER_THREAD;
    // bit b of `nl` = byte b of this thread's 64-byte chunk is a newline: four mask bits per word, gathered with
    // one multiply (the 0x01 bits of the byte mask land in bits 24..27)
    unsigned long long nl = 0;
    if (pos < n) {
        unsigned w[16];
        load_chunk(d, n, pos, w);
        unsigned lo = 0, hi = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            lo |= (((nl_mask(w[i]) & 0x01010101u) * 0x01020408u) >> 24) << (4 * i);
            hi |= (((nl_mask(w[8 + i]) & 0x01010101u) * 0x01020408u) >> 24) << (4 * i);
        }
        nl = ((unsigned long long)hi << 32) | lo;
    }
    const unsigned c = __popcll(nl);
    // exclusive scan of c over the CTA
    __shared__ unsigned ws[SP_THREADS / 32];
    unsigned lane = lane_id(), wid = threadIdx.x >> 5;
    unsigned incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (unsigned)o) incl += t;
    }
    if (lane == 31) ws[wid] = incl;
    __syncthreads();
    unsigned wbase = 0;
#pragma unroll
    for (int i = 0; i < SP_THREADS / 32; i++)
        if ((unsigned)i < wid) wbase += ws[i];
    // the tile's offsets are staged in shared memory in order and stored as one contiguous run
    __shared__ uint64_t stage[SP_STAGE];
    __shared__ unsigned tile_total;
    if (threadIdx.x == SP_THREADS - 1) tile_total = wbase + incl;
    __syncthreads();
    const unsigned total = tile_total;
    const uint64_t out0 = line_base + tile_base[tile0 + blockIdx.x] + 1;      // +1: line_off[0] = 0 is the first line
    unsigned k = wbase + incl - c;
    const bool staged = total <= SP_STAGE;
    while (nl) {
        const int b = __ffsll((long long)nl) - 1;
        nl &= nl - 1;
        const uint64_t v = pos + b + 1;                                      // the next line starts after the newline
        if (staged) stage[k] = v; else line_off[out0 + k] = v;
        k++;
    }
    if (staged) {
        __syncthreads();
        for (unsigned i = threadIdx.x; i < total; i += SP_THREADS) line_off[out0 + i] = stage[i];
    }
}

__global__ void k_set_u64(uint64_t* p, uint64_t v) { *p = v; }

extern "C" int uqb_fastq_load(uqb_ctx* ctx, const uint8_t* host, uint64_t nbytes, uqb_fastq** out) {
    uqb_fastq* fq = new uqb_fastq();
    uint8_t* d;
    int r = uqb_dalloc(ctx, (void**)&d, nbytes + 64);
    if (r) { delete fq; return r; }
    fq->d = d;
    fq->n = nbytes;
    fq->owned = true;
    UQB_CUDA(cudaMemsetAsync(d + nbytes, 0, 64, ctx->stream));
    if (nbytes) UQB_CUDA(cudaMemcpyAsync(d, host, nbytes, cudaMemcpyHostToDevice, ctx->stream));
    UQB_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = fq;
    return 0;
}

extern "C" int uqb_fastq_adopt(uqb_ctx* ctx, const uint8_t* dev, uint64_t nbytes, uqb_fastq** out) {
    if (((uintptr_t)dev) & 15) return uqb_fail(ctx, "adopted FASTQ buffer must be 16-byte aligned");
    uqb_fastq* fq = new uqb_fastq();
    fq->d = dev;
    fq->n = nbytes;
    fq->owned = false;
    *out = fq;
    return 0;
}

static int free_qcols(uqb_ctx* ctx, uqb_fastq* fq) {
    for (auto& c : fq->qcols) {
        UQB_TRY(uqb_dfree(ctx, c.val, fq->n_reads * 8));
        UQB_TRY(uqb_dfree(ctx, c.span, fq->n_reads * 4));
        UQB_TRY(uqb_dfree(ctx, c.rank, fq->n_reads * 4));
        UQB_TRY(uqb_dfree(ctx, c.first_occ, 0));
        UQB_TRY(uqb_dfree(ctx, c.dict, c.dict_count * c.dict_width + 64));
    }
    fq->qcols.clear();
    return 0;
}

int uqb_fastq_free_qcols(uqb_ctx* ctx, uqb_fastq* fq) { return free_qcols(ctx, fq); }

extern "C" int uqb_fastq_free(uqb_ctx* ctx, uqb_fastq* fq) {
    if (!fq) return 0;
    UQB_TRY(free_qcols(ctx, fq));
    UQB_TRY(uqb_scan_release(ctx, fq));
    if (fq->line_off) UQB_TRY(uqb_dfree(ctx, fq->line_off, (fq->n_lines + 1) * 8));
    delete fq->cached_stats;
    if (fq->ref_name) UQB_TRY(uqb_dfree(ctx, fq->ref_name, 0));
    if (fq->owned) UQB_TRY(uqb_dfree(ctx, (void*)fq->d, fq->n + 64));
    delete fq;
    return 0;
}

extern "C" int uqb_fastq_download(uqb_ctx* ctx, const uqb_fastq* fq, uint64_t offset, uint8_t* host, uint64_t nbytes) {
    if (offset + nbytes > fq->n) return uqb_fail(ctx, "fastq download out of range");
    if (nbytes) UQB_CUDA(cudaMemcpyAsync(host, fq->d + offset, nbytes, cudaMemcpyDeviceToHost, ctx->stream));
    UQB_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// chunk-wise building blocks for the streamed load (analyze.cu)
int uqb_split_count_tiles(uqb_ctx* ctx, const uint8_t* d, uint64_t n_end, uint64_t tile0, uint64_t ntiles, uint32_t* counts) {
    UQB_LAUNCH_B(ntiles * SP_TILE, k_newline_count, (unsigned)ntiles, SP_THREADS, 0, d, n_end, tile0, counts);
    return 0;
}
int uqb_split_write_tiles(uqb_ctx* ctx, const uint8_t* d, uint64_t n_end, uint64_t tile0, uint64_t ntiles, uint64_t line_base,
                          const uint64_t* bases, uint64_t* line_off) {
    UQB_LAUNCH_B(ntiles * SP_TILE, k_newline_write, (unsigned)ntiles, SP_THREADS, 0, d, n_end, tile0, line_base, bases, line_off);
    return 0;
}
int uqb_split_set_first(uqb_ctx* ctx, uint64_t* line_off) {
    UQB_LAUNCH(k_set_u64, 1, 1, 0, line_off, 0ull);
    return 0;
}

extern "C" int uqb_split(uqb_ctx* ctx, uqb_fastq* fq, uqb_split_info* info) {
    memset(info, 0, sizeof(*info));
    if (fq->streamed && fq->line_off) {          // already split while the bytes were streaming in
        info->n_bytes = fq->n; info->n_lines = fq->n_lines; info->n_reads = fq->n_reads;
        info->status = (fq->n_lines % 4 == 0) ? 0 : 1;
        return 0;
    }
    info->n_bytes = fq->n;
    if (fq->line_off) { UQB_TRY(uqb_dfree(ctx, fq->line_off, (fq->n_lines + 1) * 8)); fq->line_off = nullptr; }
    UQB_TRY(uqb_scan_release(ctx, fq));
    {   // sweep A: newline scan, line offsets, record checks, histograms and the compact QNAME array in ONE pass
        bool done = false;
        UQB_TRY(uqb_scan_file(ctx, fq, &done));
        if (done) {
            info->n_lines = fq->n_lines;
            info->n_reads = fq->n_reads;
            info->status = (fq->n_lines % 4 == 0) ? 0 : 1;
            return 0;
        }
    }
    uint64_t ntiles = (fq->n + SP_TILE - 1) / SP_TILE;
    uint64_t total = 0;
    uint32_t* counts = nullptr;
    uint64_t* bases = nullptr;
    uint64_t* d_total = nullptr;
    if (ntiles) {
        UQB_TRY(uqb_dalloc_t(ctx, &counts, ntiles));
        UQB_TRY(uqb_dalloc_t(ctx, &bases, ntiles));
        UQB_TRY(uqb_dalloc_t(ctx, &d_total, 1));
        UQB_LAUNCH_B(fq->n, k_newline_count, (unsigned)ntiles, SP_THREADS, 0, fq->d, fq->n, 0ull, counts);
        UQB_TRY(uqb_scan_u32_to_u64(ctx, counts, bases, ntiles, d_total));
        UQB_TRY(uqb_readback(ctx, &total, d_total, 8));
    }
    fq->n_lines = total;
    fq->n_reads = total / 4;
    UQB_TRY(uqb_dalloc_t(ctx, &fq->line_off, total + 1));
    UQB_LAUNCH(k_set_u64, 1, 1, 0, fq->line_off, 0ull);
    if (ntiles) {
        UQB_LAUNCH_B(fq->n + 8 * total, k_newline_write, (unsigned)ntiles, SP_THREADS, 0, fq->d, fq->n, 0ull, 0ull, bases, fq->line_off);
        UQB_TRY(uqb_dfree(ctx, counts, ntiles * 4));
        UQB_TRY(uqb_dfree(ctx, bases, ntiles * 8));
        UQB_TRY(uqb_dfree(ctx, d_total, 8));
    }
    info->n_lines = total;
    info->n_reads = total / 4;
    info->status = (total % 4 == 0) ? 0 : 1;
    return 0;
}

extern "C" int uqb_fastq_line_offsets(uqb_ctx* ctx, const uqb_fastq* fq, uint64_t first, uint64_t count, uint64_t* host) {
    if (!fq->line_off || first + count > fq->n_lines + 1) return uqb_fail(ctx, "line offset range out of bounds");
    if (count) UQB_CUDA(cudaMemcpyAsync(host, fq->line_off + first, count * 8, cudaMemcpyDeviceToHost, ctx->stream));
    UQB_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}
