// Sweep A: ONE pass over the raw FASTQ bytes that does what k_newline_count, k_newline_write, the per-record checks of
// k_record_stats_names and k_pair_hist_tiles did in four sweeps (included by analyze.cu, after the pt_* helpers).
//
//   tiles      a CTA takes byte ranges of SC_T bytes ("own range") in ticket order; the staged range is SC_LOAD bytes,
//              i.e. it runs SC_LOAD - SC_T bytes into the next tile, so that every record that STARTS behind a newline of
//              the own range is completely in shared memory (longer records raise `fallback`; the host then uses the
//              line-offset based kernels).  One TMA bulk copy per tile, double buffered.
//   newlines   every thread scans SC_UNIT staged bytes (3 x 16-byte shared loads, exact zero-byte test on w ^ 0x0A..),
//              a CTA scan ranks the newlines, their positions go to nlp[] in order.
//   numbering  the number of newlines in front of the tile comes from a decoupled look-back over one 64-bit status word
//              per tile (aggregate / inclusive prefix, Merrill & Garland); tickets make the protocol independent of which
//              CTAs are resident.  line_off[] is written from nlp[] as contiguous runs.
//   records    newline number g ends line g; record r starts behind newline 4r-1 and belongs to the tile that holds that
//              newline (record 0: tile 0).  Eight lanes work on one record (four records per warp): per-record checks
//              (uq.py:360, 366, 382, 388), min / max read length, the QNAME line copied to a compact side array
//              (`names`, one row of name_pitch bytes per record: length byte + text) for the name statistics and the
//              tokeniser, and the base / quality histograms:
//   histogram  a lane takes ALIGNED 32-bit words of the DNA line and of the QUAL line (the two lines are counted
//              independently; only a base that still carries the CHECK bit looks at its own quality).  Bases: pending
//              counters in registers behind a 64-bit LUT entry (pt_* scheme of k_pair_hist_tiles).  Qualities: private
//              8-bit counters in shared memory, byte granular ([value / 4][thread] words, so every lane stays in its own
//              bank), addressed through a 32-bit LUT; flushed by a skewed column sum before a counter can wrap.
#pragma once

#ifdef SC_BIG                                   // one CTA of 1024 threads per SM, 48 KB tiles
#define SC_THREADS 1024
#define SC_CTAS_PER_SM 1
#define SC_OWN_THREADS 960
#define SC_NLCAP 2048
#define SC_SMEM_LIMIT (226 * 1024)
#else                                           // two CTAs of 512 threads per SM: one computes while the other is at a barrier
#define SC_THREADS 512
#define SC_CTAS_PER_SM 2
#define SC_OWN_THREADS 480
#define SC_NLCAP 1024                           // newlines per staged range
#define SC_SMEM_LIMIT (112 * 1024)
#endif
#define SC_UNIT 48
#define SC_LOAD (SC_THREADS * SC_UNIT)          // staged bytes per tile
#define SC_T (SC_OWN_THREADS * SC_UNIT)         // bytes of own range; records of up to SC_LOAD - SC_T bytes fit behind it
#define SC_PAD 16                               // addressable bytes in front of the staged range (position -1)
#define SC_QWORDS 25                            // 24 words for byte values 32..127 + 1 word of dummy counters
#define SC_QROW (SC_THREADS * 4)
#define SC_FLUSH 252u

#define SC_FB_LONG 1u          // a record does not fit the staged range / too many newlines in a tile
#define SC_FB_BYTES 2u         // byte outside 32..127 in a DNA or QUAL line
#define SC_FB_CAP 4u           // line_off / names capacity exceeded
#define SC_FB_NAME 8u          // QNAME line longer than the row of the side array
#define SC_FB_PHASE 16u        // the record phase a tile guessed from its text was wrong

struct sc_smem {
    alignas(128) uint8_t buf[2][SC_PAD + SC_LOAD + 16];
    uint32_t priv_q[SC_QWORDS * SC_THREADS];
    uint2 lutb[256];                // base byte -> {increment of fields 0-3, (group | CHECK) << 24 | increment of fields 4-6}
    uint32_t lutq[256];             // quality byte -> byte offset of its private counter (row * SC_QROW + field)
    unsigned hist_b[256], hist_q[4 * SC_QWORDS];
    int state[256];
    uint8_t rev[PT_GROUPS * 8];
    uint16_t nlp[SC_NLCAP + 8];     // nlp[1 + i] = staged position of newline i; nlp[0] = 0xFFFF (the newline in front of the tile)
    uint2 rect[SC_NLCAP / 4 + 4];   // record table of the tile: {DNA start | QUAL start << 16, length | QNAME start << 16}
    uint32_t wtot[32];
    alignas(8) uint64_t bar[2];
    uint64_t L0;                    // newlines in front of the tile
    uint32_t tile[2];               // tickets: current / next tile
    uint32_t total_own, total_all;
    int spec;                       // guessed index of the first newline of the own range that ends a record, -1: none
};
static_assert(sizeof(sc_smem) <= SC_SMEM_LIMIT, "sweep A shared memory");

struct sc_params {
    const uint8_t* d;
    uint64_t n;                     // bytes readable (file size)
    uint64_t n_avail;               // bytes that have arrived (streamed load); == n for a resident file
    uint32_t tile_end;              // tiles [ticket start, tile_end)
    uint64_t* status;               // one word per tile: flag << 62 | value
    unsigned int* ticket;
    uint64_t* line_off;
    uint64_t cap_lines;             // line_off holds cap_lines + 1 entries
    uint8_t* names;
    uint32_t name_pitch;
    uint64_t cap_records;
    an_dev* s;
    unsigned int* fallback;
    unsigned long long* lines_total;   // written by the CTA of the last tile of the FILE
};

__device__ __forceinline__ uint64_t sc_ld_volatile(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void sc_st_volatile(uint64_t* p, uint64_t v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds_u8m(uint32_t a) {        // with memory clobber: the private counters are read back
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u8(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 lds_u128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
// 0x80 in every byte of w that equals '\n' (exact for all byte values)
__device__ __forceinline__ uint32_t sc_nl_flags(uint32_t w) {
    const uint32_t x = w ^ 0x0A0A0A0Au;
    const uint32_t t = (x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
    return ~(t | x) & 0x80808080u;
}
// low x bytes set, x in 0..4
__device__ __forceinline__ uint32_t sc_byte_prefix(uint32_t x) { return __funnelshift_rc(0xFFFFFFFFu, 0u, 32u - 8u * x); }

__device__ __forceinline__ uint64_t sc_warp_sum64(uint64_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// private quality counters of one warp -> CTA histogram: lane l sums word row l over the warp's 32 columns (skewed, so
// that the 25 lanes hit 25 different banks), every lane clears its own column.  All 32 lanes must call it.
__device__ __forceinline__ void sc_flush_q(sc_smem* S, unsigned tid) {
    const unsigned lane = tid & 31u, wbase = tid & ~31u;
    __syncwarp();
    unsigned e = 0, o = 0;
    if (lane < SC_QWORDS) {
        const unsigned* row = S->priv_q + lane * SC_THREADS + wbase;
#pragma unroll 8
        for (unsigned i = 0; i < 32; i++) {
            const unsigned x = row[(lane + i) & 31u];
            e += x & 0x00FF00FFu;
            o += (x >> 8) & 0x00FF00FFu;
        }
    }
    __syncwarp();
#pragma unroll
    for (int w = 0; w < SC_QWORDS; w++) S->priv_q[w * SC_THREADS + tid] = 0;
    if (lane < SC_QWORDS) {
        if (e & 0xFFFFu) atomicAdd(&S->hist_q[4 * lane], e & 0xFFFFu);
        if (o & 0xFFFFu) atomicAdd(&S->hist_q[4 * lane + 1], o & 0xFFFFu);
        if (e >> 16) atomicAdd(&S->hist_q[4 * lane + 2], e >> 16);
        if (o >> 16) atomicAdd(&S->hist_q[4 * lane + 3], o >> 16);
    }
    __syncwarp();
}

// pending base counters -> CTA histogram (same scheme as pt_base_flush, on sc_smem)
__device__ __noinline__ void sc_base_flush(sc_smem* S, unsigned cur, unsigned pa, unsigned pb) {
    const unsigned g = cur >> 24;
    if (g >= PT_GROUPS) return;
#pragma unroll
    for (unsigned f = 0; f < 4; f++) {
        const unsigned ca = (pa >> (8 * f)) & 255u;
        if (ca) atomicAdd(&S->hist_b[S->rev[g * 8 + f]], ca);
    }
#pragma unroll
    for (unsigned f = 0; f < 3; f++) {
        const unsigned cb = (pb >> (8 * f)) & 255u;
        if (cb) atomicAdd(&S->hist_b[S->rev[g * 8 + 4 + f]], cb);
    }
}

// slow path of one base: bytes outside 32..126 (group 0: reported), single-quality state, change of the pending group
__device__ __noinline__ uint3 sc_base_slow(sc_smem* S, unsigned b, unsigned q, unsigned ey, unsigned cur, unsigned pa, unsigned pb,
                                           unsigned int* fallback) {
    if ((ey & (PT_GMASK & ~PT_CHECK)) == 0u) {                   // not a byte this kernel counts
        atomicOr(fallback, SC_FB_BYTES);
        return make_uint3(cur, pa, pb);
    }
    if (ey & PT_CHECK) {
        int f = S->state[b];
        if (f != 256 && f != (int)q) {
            if (f < 0) {
                const int old = atomicCAS(&S->state[b], -1, (int)q);
                if (old >= 0 && old != (int)q) { S->state[b] = 256; f = 256; }
            } else {
                S->state[b] = 256; f = 256;
            }
        }
        if (f == 256) S->lutb[b].y = ey & ~PT_CHECK;             // several qualities: no more checks for this base
    }
    const unsigned grp = ey & (PT_GMASK & ~PT_CHECK);
    if (grp != cur) {
        sc_base_flush(S, cur, pa, pb);
        cur = grp; pa = 0; pb = 0;
    }
    return make_uint3(cur, pa, pb);
}

// one base byte b at staged address `pos_a` (its quality sits at pos_a + qdelta)
#define SC_BASE(b, pos_a)                                                                                         \
    do {                                                                                                          \
        const uint2 e_ = lds_u64(lutb_a + ((b) << 3));                                                            \
        const unsigned dd_ = (e_.y ^ cur) & PT_GMASK;                                                             \
        if (dd_) {                                                                                                \
            const unsigned q_ = lds_u8((pos_a) + qdelta);                                                         \
            if (dd_ != PT_CHECK || lds_u32(state_a + ((b) << 2)) != q_) {                                         \
                const uint3 r_ = sc_base_slow(S, (b), q_, e_.y, cur, pa, pb, P.fallback);                         \
                cur = r_.x; pa = r_.y; pb = r_.z;                                                                 \
                if ((e_.y & (PT_GMASK & ~PT_CHECK)) == 0u) break;                                                 \
            }                                                                                                     \
        }                                                                                                         \
        pa += e_.x;                                                                                               \
        pb += e_.y;                                                                                               \
    } while (0)

// one quality byte q: private 8-bit counter, byte granular
#define SC_QUAL(q)                                                                                                \
    do {                                                                                                          \
        const uint32_t a_ = pq_a + lds_u32(lutq_a + ((q) << 2));                                                  \
        sts_u8(a_, lds_u8m(a_) + 1u);                                                                             \
    } while (0)

__global__ void __launch_bounds__(SC_THREADS, SC_CTAS_PER_SM) k_scan_hist(const sc_params P) {
    extern __shared__ __align__(128) uint8_t sc_raw[];
    sc_smem* S = reinterpret_cast<sc_smem*>(sc_raw);
    const unsigned tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    // ---- one-time set-up ----
    for (unsigned i = tid; i < SC_QWORDS * SC_THREADS; i += SC_THREADS) S->priv_q[i] = 0;
    for (unsigned i = tid; i < 4 * SC_QWORDS; i += SC_THREADS) S->hist_q[i] = 0;
    for (unsigned i = tid; i < PT_GROUPS * 8; i += SC_THREADS) S->rev[i] = 0;
    if (tid < 256) { S->hist_b[tid] = 0; S->state[tid] = -1; }
    if (tid < 32) S->wtot[tid] = 0;
    if (tid == 0) {
        mbar_init(&S->bar[0], 1);
        mbar_init(&S->bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        S->tile[0] = atomicAdd(P.ticket, 1u);
        S->nlp[0] = 0xFFFFu;
    }
    __syncthreads();
    if (tid < 256) {
        const unsigned v = tid;
        unsigned grp, field;
        pt_base_group(v, &grp, &field);
        if (v < 32u || v > 126u) { grp = 0; field = 0; }            // group 0: never equals a pending group -> slow path -> reported
        uint2 eb = make_uint2((grp && field < 4) ? 1u << (8 * field) : 0u,
                              (grp ? ((grp << 24) | PT_CHECK) : 0u) | ((grp && field >= 4) ? 1u << (8 * (field - 4)) : 0u));
        // '\n' cannot occur inside a line: it stands for "no byte here" in the partial words at the two ends of a line
        // (bases: group of A C G T N with no increment; qualities: an ignored counter)
        if (v == 10u) eb = make_uint2(0u, 1u << 24);
        S->lutb[v] = eb;
        if (grp) S->rev[grp * 8 + field] = (uint8_t)v;
        const unsigned idx = v == 10u ? 97u : ((v >= 32u && v < 128u) ? v - 32u : 96u);   // out of range -> dummy word 24, field 0
        S->lutq[v] = (idx >> 2) * SC_QROW + (idx & 3u);
    }
    __syncthreads();
    const uint32_t pq_a = smem_u32(S->priv_q) + tid * 4u;
    const uint32_t lutb_a = smem_u32(S->lutb), lutq_a = smem_u32(S->lutq), state_a = smem_u32(S->state), nlp_a = smem_u32(S->nlp), rect_a = smem_u32(S->rect);
    unsigned cur = PT_NONE, pa = 0, pb = 0, since_flush = 0;
    unsigned long long mn = ~0ull, mx = 0ull;
    unsigned nm = 0;
    long long bad_plus = LLONG_MAX, bad_len = LLONG_MAX;
    const uint32_t ntiles_file = (uint32_t)((P.n + SC_T - 1) / SC_T);

    // thread 0: start the bulk copy of tile t into buffer bf
    auto issue = [&](uint32_t t, unsigned bf) {
        const uint64_t s0 = (uint64_t)t * SC_T;
        uint64_t end = s0 + SC_LOAD;
        const uint64_t lim = P.n_avail & ~15ull;
        if (end > lim) end = lim;
        const uint32_t bulk = end > s0 ? (uint32_t)(end - s0) : 0u;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&S->bar[bf], bulk);
        if (bulk) bulk_g2s(S->buf[bf] + SC_PAD, P.d + s0, bulk, &S->bar[bf]);
    };
    unsigned buf = 0, phases = 0;
    uint32_t t = S->tile[0];
    if (tid == 0 && t < P.tile_end) issue(t, 0);
    while (t < P.tile_end) {
        // ---- prefetch the next tile, wait for this one ----
        if (tid == 0) {
            const uint32_t tn = atomicAdd(P.ticket, 1u);
            S->tile[1] = tn;
            if (tn < P.tile_end) issue(tn, buf ^ 1u);
        }
        const uint64_t s0 = (uint64_t)t * SC_T;
        const uint32_t valid = (uint32_t)(P.n_avail - s0 < SC_LOAD ? P.n_avail - s0 : SC_LOAD);
        {   // tail of the stream that is not a full 16-byte unit (last tile of the file)
            const uint64_t lim = P.n_avail & ~15ull;
            const uint64_t a1 = lim > s0 ? (lim - s0 < SC_LOAD ? lim - s0 : SC_LOAD) : 0;
            for (uint32_t q = (uint32_t)a1 + tid; q < valid; q += SC_THREADS) S->buf[buf][SC_PAD + q] = P.d[s0 + q];
        }
        mbar_wait(&S->bar[buf], (phases >> buf) & 1u);
        phases ^= 1u << buf;
        __syncthreads();                                     // tail bytes visible
        const uint32_t data_a = smem_u32(S->buf[buf]) + SC_PAD;
        // ---- newline scan: bit 8b + w of pk[v] = byte b of word w of the thread's v-th 16-byte unit is a newline ----
        uint32_t pk[SC_UNIT / 16];
        const uint32_t ub = tid * SC_UNIT;
        unsigned c = 0;
        if (ub + SC_UNIT <= valid) {
#pragma unroll
            for (int v = 0; v < SC_UNIT / 16; v++) {
                const uint4 x = lds_u128(data_a + ub + 16 * v);
                pk[v] = (sc_nl_flags(x.x) >> 7) | (sc_nl_flags(x.y) >> 6) | (sc_nl_flags(x.z) >> 5) | (sc_nl_flags(x.w) >> 4);
            }
        } else {
#pragma unroll
            for (int v = 0; v < SC_UNIT / 16; v++) {
                uint32_t f = 0;
                for (int i = 0; i < 16; i++) {
                    const uint32_t p = ub + 16 * v + i;
                    if (p < valid && lds_u8(data_a + p) == 10u) f |= 1u << (8 * (i & 3) + (i >> 2));
                }
                pk[v] = f;
            }
        }
#pragma unroll
        for (int v = 0; v < SC_UNIT / 16; v++) c += __popc(pk[v]);
        unsigned incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned x = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (unsigned)o) incl += x;
        }
        if (lane == 31) S->wtot[wid] = incl;
        __syncthreads();
        unsigned wt = S->wtot[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned x = __shfl_up_sync(0xffffffffu, wt, o);
            if (lane >= (unsigned)o) wt += x;
        }
        const unsigned wbase = wid ? __shfl_sync(0xffffffffu, wt, (int)wid - 1) : 0u;
        const unsigned total_own = __shfl_sync(0xffffffffu, wt, SC_OWN_THREADS / 32 - 1);
        const unsigned total_all = __shfl_sync(0xffffffffu, wt, 31);
        {   // newline positions, in order: every lane hands out its next newline per round, the warp runs as many rounds
            // as its busiest lane has newlines (two or three); several newlines in ONE 16-byte unit are rare and are
            // taken by byte position
            static_assert(SC_UNIT == 48, "three 16-byte units per thread");
            unsigned k = wbase + incl - c;
            const unsigned cmax = __reduce_max_sync(0xffffffffu, c);
            uint32_t a0 = pk[0], a1 = pk[1], a2 = pk[2];
            for (unsigned it = 0; it < cmax; it++) {
                const uint32_t f = a0 ? a0 : (a1 ? a1 : a2);
                if (f) {
                    const uint32_t vb = a0 ? 0u : (a1 ? 16u : 32u);
                    uint32_t p = (uint32_t)__ffs((int)f) - 1u;
                    uint32_t rest = f & (f - 1u);
                    if (rest) {
                        uint32_t best = 4u * (p & 7u) + (p >> 3);
                        while (rest) {
                            const uint32_t q = (uint32_t)__ffs((int)rest) - 1u;
                            rest &= rest - 1u;
                            const uint32_t key = 4u * (q & 7u) + (q >> 3);
                            if (key < best) { best = key; p = q; }
                        }
                    }
                    const uint32_t cleared = f & ~(1u << p);
                    if (a0) a0 = cleared; else if (a1) a1 = cleared; else a2 = cleared;
                    if (k < SC_NLCAP) S->nlp[1 + k] = (uint16_t)(ub + vb + 4u * (p & 7u) + (p >> 3));
                    k++;
                }
            }
        }
        // ---- line numbering, part 1: publish the aggregate and START the look-back (warp 0: one window of status words
        //      is requested now and looked at after the histograms - the round trip hides behind them) ----
        uint64_t lb_v = 2ull << 62;
        if (wid == 0) {
            if (lane == 0) {
                __threadfence();
                sc_st_volatile(P.status + t, (1ull << 62) | total_own);
                S->total_own = total_own; S->total_all = total_all;
            }
            const int64_t my = (int64_t)t - 1 - (int64_t)lane;
            if (my >= 0) lb_v = sc_ld_volatile(P.status + my);
        }
        // ---- which newline of the own range ends a record?  Known exactly once L0 is (newline number 4r-1 precedes
        //      record r).  Until then the tile SPECULATES from its own text: the phase for which the first records start
        //      with '@' and have a '+' line two lines further - used only if exactly one phase qualifies, and checked
        //      against L0 afterwards (a wrong guess invalidates the scan: `fallback`).  Tile 0 knows (L0 = 0). ----
        __syncthreads();                                     // nlp[] is complete
        if (tid < 4) {
            bool good = false;
            if (t > 0 && total_all <= SC_NLCAP) {
                good = true;
                int seen = 0;
                for (int k = (int)tid; k + 4 < (int)total_all && seen < 3; k += 4, seen++) {
                    const uint32_t h = S->nlp[1 + k] + 1u, pl = S->nlp[1 + k + 2] + 1u;
                    if (lds_u8(data_a + h) != (unsigned)'@' || lds_u8(data_a + pl) != (unsigned)'+') good = false;
                }
                if (seen == 0) good = false;
            }
            const unsigned m = __ballot_sync(0xFu, good);
            if (tid == 0) S->spec = (__popc(m) == 1) ? (int)(__ffs((int)m) - 1) : -1;
        }
        __syncthreads();
        const int spec = t == 0 ? -4 : S->spec;               // -4: tile 0 (kfirst = -1, exact); -1: no guess
        const bool at_eof = s0 + valid >= P.n;
        const unsigned sub = lane >> 3, sl = lane & 7u;
        int tb_plus = INT_MAX, tb_len = INT_MAX;                // first bad record of THIS tile (index inside the tile)
        uint64_t L0 = 0;
        for (int pass = 0; pass < 2; pass++) {
            // pass 0: histograms + per-record checks with the speculated phase (skipped without a guess)
            // pass 1: after the look-back - the same with the exact phase if pass 0 was skipped
            if (pass == 1) {
                // ---- line numbering, part 2: finish the look-back ----
                if (wid == 0) {
                    uint64_t run = 0;
                    if (t > 0) {
                        int64_t idx = (int64_t)t - 1;
                        uint64_t v = lb_v;
                        while (true) {
                            const int64_t my = idx - (int64_t)lane;
                            if (my >= 0) {
                                while ((v >> 62) == 0) { __nanosleep(20); v = sc_ld_volatile(P.status + my); }
                            } else {
                                v = 2ull << 62;               // in front of tile 0: inclusive prefix 0
                            }
                            const unsigned inc_mask = __ballot_sync(0xffffffffu, (v >> 62) == 2ull);
                            if (inc_mask) {
                                const int first = __ffs((int)inc_mask) - 1;
                                run += sc_warp_sum64((int)lane <= first ? (v & ((1ull << 62) - 1)) : 0ull);
                                break;
                            }
                            run += sc_warp_sum64(v & ((1ull << 62) - 1));
                            idx -= 32;
                            const int64_t nx = idx - (int64_t)lane;
                            v = nx >= 0 ? sc_ld_volatile(P.status + nx) : (2ull << 62);
                        }
                    }
                    if (lane == 0) {
                        __threadfence();
                        sc_st_volatile(P.status + t, (2ull << 62) | (run + total_own));
                        S->L0 = run;
                        if (t + 1 == ntiles_file) *P.lines_total = run + total_own;
                    }
                }
                __syncthreads();
                L0 = S->L0;
                if (spec >= 0 && (int)((3u - (unsigned)(L0 & 3ull)) & 3u) != spec && tid == 0) atomicOr(P.fallback, SC_FB_PHASE);
            }
            if (total_all > SC_NLCAP) {
                if (tid == 0 && pass == 1) atomicOr(P.fallback, SC_FB_LONG);
                continue;
            }
            const bool do_hist = pass == 0 ? spec != -1 : spec == -1;
            if (!do_hist) continue;
            const int kfirst = pass == 0 ? (spec == -4 ? -1 : spec) : (int)((3u - (unsigned)(L0 & 3ull)) & 3u) - (t == 0 ? 4 : 0);
            const int nrec = kfirst < (int)total_own ? ((int)total_own - kfirst + 3) / 4 : 0;
            // ---- record table: geometry and per-record checks (uq.py:360, 366, 382, 388), one thread per record ----
            for (int rec = (int)tid; rec < nrec; rec += SC_THREADS) {
                const int k = kfirst + 4 * rec;
                uint2 e = make_uint2(0u, 0xFFFF0000u);               // length 0, QNAME start 0xFFFF: not a complete record
                if (k + 4 < (int)total_all) {
                    const uint32_t a = nlp_a + 2u * (uint32_t)(1 + k);
                    const uint32_t p0 = lds_u16(a), p1 = lds_u16(a + 2), p2 = lds_u16(a + 4), p3 = lds_u16(a + 6), p4 = lds_u16(a + 8);
                    const uint32_t ns = (p0 + 1u) & 0xFFFFu;         // 0xFFFF + 1 wraps to 0: the record that starts the file
                    const uint32_t name_len = p1 - ns, dlen = p2 - p1 - 1u, qlen = p4 - p3 - 1u;
                    const bool plus_ok = p3 - p2 >= 2u && lds_u8(data_a + p2 + 1u) == (unsigned)'+';
                    if (!plus_ok) tb_plus = tb_plus < rec ? tb_plus : rec;
                    if (dlen != qlen) tb_len = tb_len < rec ? tb_len : rec;
                    mn = dlen < mn ? dlen : mn; mx = dlen > mx ? dlen : mx;
                    nm = name_len > nm ? name_len : nm;
                    e = make_uint2((p1 + 1u) | ((p3 + 1u) << 16), (dlen < qlen ? dlen : qlen) | (ns << 16));
                } else if (!at_eof) {
                    atomicOr(P.fallback, SC_FB_LONG);
                }
                S->rect[rec] = e;
            }
            __syncthreads();
            for (int g0 = (int)wid * 4; g0 < nrec; g0 += 4 * (SC_THREADS / 32)) {
                const int rec = g0 + (int)sub;
                const uint2 R = rec < nrec ? lds_u64(rect_a + 8u * (uint32_t)rec) : make_uint2(0u, 0xFFFF0000u);
                // ---- histograms: aligned words of the DNA line and of the QUAL line ----
                const uint32_t len = R.y & 0xFFFFu;
                const uint32_t da = R.x & 0xFFFFu, qa = R.x >> 16;
                const uint32_t qdelta = qa - da;
                const uint32_t fwd = da >> 2, fwq = qa >> 2;
                const uint32_t nwd = len ? ((da + len - 1u) >> 2) - fwd + 1u : 0u, nwq = len ? ((qa + len - 1u) >> 2) - fwq + 1u : 0u;
                const uint32_t nwmax = nwd > nwq ? nwd : nwq, nwmin = nwd < nwq ? nwd : nwq;
                const unsigned iters = (__reduce_max_sync(0xffffffffu, nwmax) + 7u) >> 3;
                // iterations 1 .. full_hi touch interior words only (no masks) for every record of the warp
                const int full_hi = __reduce_min_sync(0xffffffffu, ((int)nwmin - 9) >> 3);
                if (since_flush + 4u * iters > SC_FLUSH) {        // warp uniform
                    sc_flush_q(S, tid);
                    sc_base_flush(S, cur, pa, pb);
                    pa = 0; pb = 0;
                    since_flush = 0;
                }
                since_flush += 4u * iters;
                const uint32_t mfd = 0xFFFFFFFFu << (8u * (da & 3u)), mld = 0xFFFFFFFFu >> (8u * (3u - ((da + len - 1u) & 3u)));
                const uint32_t mfq = 0xFFFFFFFFu << (8u * (qa & 3u)), mlq = 0xFFFFFFFFu >> (8u * (3u - ((qa + len - 1u) & 3u)));
                for (unsigned it = 0; it < iters; it++) {
                    const uint32_t od = sl + 8u * it;                          // word offset inside both lines
                    const uint32_t wa_d = data_a + 4u * (fwd + od), wa_q = data_a + 4u * (fwq + od);
                    uint32_t wd = lds_u32(wa_d), wq = lds_u32(wa_q);
                    if (!((int)it >= 1 && (int)it <= full_hi)) {
                        // a word at an end of a line (or behind it): bytes outside the line become '\n'
                        uint32_t md = od < nwd ? 0xFFFFFFFFu : 0u, mq = od < nwq ? 0xFFFFFFFFu : 0u;
                        if (od == 0) { md &= mfd; mq &= mfq; }
                        if (od + 1u == nwd) md &= mld;
                        if (od + 1u == nwq) mq &= mlq;
                        wd = (wd & md) | (0x0A0A0A0Au & ~md);
                        wq = (wq & mq) | (0x0A0A0A0Au & ~mq);
                    }
                    // ---- qualities: the four counter addresses first (independent LUT loads), then the four updates in
                    //      order (two bytes of one word may hit the same counter) ----
                    uint32_t qa4[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) qa4[j] = lds_u32(lutq_a + (__byte_perm(wq, 0, 0x4440 + j) << 2));
                    // ---- bases: four LUT entries, ONE test for "all four in the pending group" ----
                    uint2 eb[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) eb[j] = lds_u64(lutb_a + (__byte_perm(wd, 0, 0x4440 + j) << 3));
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const uint32_t a_ = pq_a + qa4[j];
                        sts_u8(a_, lds_u8m(a_) + 1u);
                    }
                    const uint32_t gx = ((eb[0].y ^ cur) | (eb[1].y ^ cur)) | ((eb[2].y ^ cur) | (eb[3].y ^ cur));
                    bool fast = (gx & (PT_GMASK & ~PT_CHECK)) == 0u;
                    if (fast && (gx & PT_CHECK)) {
                        // bases that had one single quality so far (N, typically): still the same one?
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            if (eb[j].y & PT_CHECK) {
                                const unsigned b = __byte_perm(wd, 0, 0x4440 + j);
                                if (lds_u32(state_a + (b << 2)) != lds_u8(wa_d + j + qdelta)) fast = false;
                            }
                        }
                    }
                    if (fast) {
                        pa += (eb[0].x + eb[1].x) + (eb[2].x + eb[3].x);
                        pb += (eb[0].y + eb[1].y) + (eb[2].y + eb[3].y);
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const unsigned b = __byte_perm(wd, 0, 0x4440 + j);
                            if (b != 10u) SC_BASE(b, wa_d + j);
                        }
                    }
                }
            }
        }
        // ---- everything that needs the global line number: line offsets, QNAME rows, record indices ----
        if (total_all <= SC_NLCAP) {
            for (unsigned k = tid; k < total_own; k += SC_THREADS) {
                const uint64_t li = L0 + k + 1;
                if (li <= P.cap_lines) P.line_off[li] = s0 + S->nlp[1 + k] + 1;
                else atomicOr(P.fallback, SC_FB_CAP);
            }
            const int kfirst = (int)((3u - (unsigned)(L0 & 3ull)) & 3u) - (t == 0 ? 4 : 0);
            const int nrec = kfirst < (int)total_own ? ((int)total_own - kfirst + 3) / 4 : 0;
            const uint64_t rec0 = (L0 + (uint64_t)(kfirst + 1)) >> 2;
            if (tb_plus != INT_MAX) { const long long r = (long long)rec0 + tb_plus; bad_plus = bad_plus < r ? bad_plus : r; }
            if (tb_len != INT_MAX) { const long long r = (long long)rec0 + tb_len; bad_len = bad_len < r ? bad_len : r; }
            for (int g0 = (int)wid * 4; g0 < nrec; g0 += 4 * (SC_THREADS / 32)) {
                const int rec = g0 + (int)sub;
                if (rec >= nrec) continue;
                const uint2 R = lds_u64(rect_a + 8u * (uint32_t)rec);
                const uint32_t ns = R.y >> 16;
                if (ns == 0xFFFFu) continue;                         // not a complete record
                const uint32_t name_len = (R.x & 0xFFFFu) - 1u - ns;
                const uint64_t r = rec0 + (uint64_t)rec;
                if (name_len + 1u > P.name_pitch || name_len > 255u) {
                    if (sl == 0) atomicOr(P.fallback, SC_FB_NAME);
                } else if (r >= P.cap_records) {
                    if (sl == 0) atomicOr(P.fallback, SC_FB_CAP);
                } else {
                    uint8_t* row = P.names + r * P.name_pitch;
                    for (uint32_t o8 = 8u * sl; o8 < P.name_pitch; o8 += 64u) {
                        // row bytes o8 .. o8+7 = staged bytes (ns - 1 + o8) ..; byte 0 of the row is the length
                        const uint32_t src = data_a + ns + o8 - 1u;
                        const uint32_t aw = src & ~3u, sh = (src & 3u) * 8u;
                        const uint32_t x0 = lds_u32(aw), x1 = lds_u32(aw + 4), x2 = lds_u32(aw + 8);
                        uint32_t lo = __funnelshift_r(x0, x1, sh), hi = __funnelshift_r(x1, x2, sh);
                        const uint32_t have = name_len + 1u > o8 ? name_len + 1u - o8 : 0u;      // valid bytes of this unit
                        lo &= sc_byte_prefix(have < 4u ? have : 4u);
                        hi &= sc_byte_prefix(have > 4u ? (have - 4u < 4u ? have - 4u : 4u) : 0u);
                        if (o8 == 0) lo = (lo & ~0xFFu) | name_len;
                        *reinterpret_cast<uint2*>(row + o8) = make_uint2(lo, hi);
                    }
                }
            }
        }
        __syncthreads();                                     // everybody is done with this buffer and with nlp[]
        t = S->tile[1];
        __syncthreads();
        buf ^= 1u;
    }
    // ---- CTA results -> global ----
    sc_flush_q(S, tid);
    sc_base_flush(S, cur, pa, pb);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long a = __shfl_xor_sync(0xffffffffu, mn, o), b = __shfl_xor_sync(0xffffffffu, mx, o);
        unsigned cc = __shfl_xor_sync(0xffffffffu, nm, o);
        long long e = __shfl_xor_sync(0xffffffffu, bad_plus, o), f = __shfl_xor_sync(0xffffffffu, bad_len, o);
        mn = a < mn ? a : mn; mx = b > mx ? b : mx; nm = cc > nm ? cc : nm;
        bad_plus = e < bad_plus ? e : bad_plus; bad_len = f < bad_len ? f : bad_len;
    }
    if (lane == 0) {
        if (mn != ~0ull) atomicMin(&P.s->dna_min, mn);
        atomicMax(&P.s->dna_max, mx);
        atomicMax(&P.s->max_name_len, nm);
        if (bad_plus != LLONG_MAX) atomicMin(&P.s->bad_plus, bad_plus);
        if (bad_len != LLONG_MAX) atomicMin(&P.s->bad_len, bad_len);
    }
    __syncthreads();
    if (tid < 256) {
        if (S->hist_b[tid]) atomicAdd(&P.s->base_count[tid], (unsigned long long)S->hist_b[tid]);
        if (tid >= 32 && tid < 128 && S->hist_q[tid - 32]) atomicAdd(&P.s->qual_count[tid], (unsigned long long)S->hist_q[tid - 32]);
        if (tid == 0 && S->hist_q[96]) atomicOr(P.fallback, SC_FB_BYTES);          // [97] counts the "no byte here" fillers
        const int f = S->state[tid];
        if (f >= 0) {
            if (f == 256) {
                P.s->multi[tid] = 1;
                atomicCAS(&P.s->first_q[tid], -1, 0);                 // mark the base as present
            } else {
                const int old = atomicCAS(&P.s->first_q[tid], -1, f);
                if (old >= 0 && old != f) P.s->multi[tid] = 1;
            }
        }
    }
}
