"""Synthetic FASTQ generators (host / numpy) for the BASELINE.json configurations.

TEST INFRASTRUCTURE.  The same counter-based generator is implemented on the device
(uq_b200/csrc/synth.cu) so that bench.py can create the 100 M-read inputs directly in HBM;
tests check that the two produce identical bytes.

Everything is a pure function of (seed, stream, index) through splitmix64-style mixing, so any
record can be produced independently (this is what lets the device generate records in parallel
and lets every rank of a multi-GPU run generate its own contiguous read range).

kinds:
  illumina : @SIM001:1:FCX123:<lane 1-8>:<tile 1101-2678>:<x 1000-29999>:<y 1000-199999>
             i.i.d. ACGT, N with p=0.005 carrying the exclusive quality '#'
  genome   : same header; DNA = L-mers of a hash-defined random genome (duplicates exist),
             0.1 % substitutions; qualities drawn per read from a pool (duplicates exist)
  casava   : @EAS139:136:FC706VJ:<lane>:<tile>:<x>:<y> 1:<Y|N>:<even 0-38>:<one of 4 barcodes>
  offset   : @HWI-ST:<990-1009>:<7+3r> <A|B|C>/1  (host only; offset columns, suffix, mapping)
  ont      : @ONT7:<run 1-4>:<read index+1>:<channel 1-512>, variable length (log-uniform),
             ACGT only, 70 quality symbols (7 bit)
"""
import numpy as np

M64 = np.uint64(0xFFFFFFFFFFFFFFFF)
GOLD = 0x9E3779B97F4A7C15
IDXK = 0x632BE59BD9B4E019

# stream ids (shared with synth.cu)
ST_LANE, ST_TILE, ST_X, ST_Y = 1, 2, 3, 4
ST_NPOS, ST_BASE, ST_QA, ST_QB = 5, 6, 7, 8
ST_FLAG, ST_EVEN, ST_BARCODE = 9, 10, 11
ST_OFF, ST_SUB, ST_SUBV, ST_POOL = 12, 13, 14, 15
ST_GENOME, ST_LEN, ST_RUN, ST_CH = 100, 16, 17, 18
ST_NQ = 19

BARCODES = [b"ATCACG", b"CGATGT", b"TTAGGC", b"TGACCA"]


def mix64(x):
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = x ^ (x >> np.uint64(30))
        x = x * np.uint64(0xBF58476D1CE4E5B9)
        x = x ^ (x >> np.uint64(27))
        x = x * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return x


def rnd(seed, stream, idx):
    """64 random bits for (seed, stream, idx); idx may be an array."""
    idx = np.asarray(idx, dtype=np.uint64)
    with np.errstate(over="ignore"):
        k = np.uint64((int(seed) + int(stream) * GOLD) & 0xFFFFFFFFFFFFFFFF)
        return mix64(k ^ mix64(idx + np.uint64(IDXK)))


def _skew(a, b, nsym):
    """Skewed symbol index in [0, nsym): product of two uniforms, clipped."""
    m = np.uint64(nsym + 1)
    v = ((a % m) * (b % m)) // m
    return np.minimum(v, np.uint64(nsym - 1)).astype(np.int64)


def ont_length_table(lo, hi):
    """Integer table used by both host and device so that no float pow runs on the device:
    4096 log-spaced lengths; record picks entry (rnd >> 52)."""
    k = np.arange(4096, dtype=np.float64) / 4096.0
    t = np.floor(lo * np.power(float(hi) / float(lo), k)).astype(np.int64)
    return np.clip(t, lo, hi)


def make_records(kind="illumina", n=1000, length=100, seed=1, first=0, genome=None, pool=None,
                 n_two_quals=False):
    """Return list of (header, dna, qual) byte strings for records first..first+n-1."""
    r = np.arange(first, first + n, dtype=np.uint64)
    if kind == "ont":
        lo, hi = length
        tab = ont_length_table(lo, hi)
        lens = tab[(rnd(seed, ST_LEN, r) >> np.uint64(52)).astype(np.int64)]
    else:
        lens = np.full(n, int(length), dtype=np.int64)
    lane = 1 + (rnd(seed, ST_LANE, r) % np.uint64(8)).astype(np.int64)
    tile = 1101 + (rnd(seed, ST_TILE, r) % np.uint64(1578)).astype(np.int64)
    x = 1000 + (rnd(seed, ST_X, r) % np.uint64(29000)).astype(np.int64)
    y = 1000 + (rnd(seed, ST_Y, r) % np.uint64(199000)).astype(np.int64)
    out = []
    if kind == "genome":
        L = int(length)
        if genome is None:
            genome = max(4 * L, (first + n) // 4 + L + 1)
        if pool is None:
            pool = max(1, (first + n) // 5)
        off = (rnd(seed, ST_OFF, r) % np.uint64(genome - L)).astype(np.uint64)
        pidx = (rnd(seed, ST_POOL, r) % np.uint64(pool)).astype(np.uint64)
    if kind == "casava":
        flag = (rnd(seed, ST_FLAG, r) % np.uint64(8) == 0)
        even = 2 * (rnd(seed, ST_EVEN, r) % np.uint64(20)).astype(np.int64)
        bc = (rnd(seed, ST_BARCODE, r) % np.uint64(4)).astype(np.int64)
    if kind == "ont":
        run = 1 + (rnd(seed, ST_RUN, r) % np.uint64(4)).astype(np.int64)
        ch = 1 + (rnd(seed, ST_CH, r) % np.uint64(512)).astype(np.int64)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    for i in range(n):
        L = int(lens[i])
        ri = int(r[i])
        pos = np.arange(L, dtype=np.uint64)
        with np.errstate(over="ignore"):
            gidx = np.uint64(ri) * np.uint64(1 << 20) + pos if kind != "ont" else np.uint64(ri) * np.uint64(1 << 24) + pos
        if kind == "genome":
            g = off[i] + pos
            b = (rnd(seed, ST_GENOME, g) & np.uint64(3)).astype(np.int64)
            sub = (rnd(seed, ST_SUB, gidx) % np.uint64(1000)) == 0
            sv = (rnd(seed, ST_SUBV, gidx) % np.uint64(3)).astype(np.int64)
            b = np.where(sub, (b + 1 + sv) & 3, b)
            with np.errstate(over="ignore"):
                qi = pidx[i] * np.uint64(L) + pos
            qidx = _skew(rnd(seed, ST_QA, qi), rnd(seed, ST_QB, qi), 38)
        else:
            b = (rnd(seed, ST_BASE, gidx) & np.uint64(3)).astype(np.int64)
            if kind == "ont":
                qidx = _skew(rnd(seed, ST_QA, gidx), rnd(seed, ST_QB, gidx), 70)
            else:
                qidx = _skew(rnd(seed, ST_QA, gidx), rnd(seed, ST_QB, gidx), 38)
        dna = acgt[b].copy()
        if kind == "ont":
            qual = (33 + qidx).astype(np.uint8)
        else:
            qual = (74 - qidx).astype(np.uint8)          # 'J' downwards to '%'
            isn = (rnd(seed, ST_NPOS, gidx) % np.uint64(200)) == 0
            dna[isn] = ord("N")
            if n_two_quals:
                nq = np.where((rnd(seed, ST_NQ, gidx) & np.uint64(1)) == 1, 35, 36).astype(np.uint8)
                qual[isn] = nq[isn]
            else:
                qual[isn] = ord("#")
        if kind in ("illumina", "genome"):
            h = b"@SIM001:1:FCX123:%d:%d:%d:%d" % (lane[i], tile[i], x[i], y[i])
        elif kind == "casava":
            h = b"@EAS139:136:FC706VJ:%d:%d:%d:%d 1:%s:%d:%s" % (
                lane[i], tile[i], x[i], y[i], b"Y" if flag[i] else b"N", even[i], BARCODES[bc[i]])
        elif kind == "ont":
            h = b"@ONT7:%d:%d:%d" % (run[i], ri + 1, ch[i])
        elif kind == "offset":
            # small-range integers above the dtype maximum (offset=True), a suffix, a non-integer mapping
            h = b"@HWI-ST:%d:%d %s/1" % (990 + int(lane[i] - 1) * 2 + int(tile[i] & 1) + 3 * int(x[i] % 2),
                                         7 + 3 * ri, b"ABC"[int(y[i] % 3):int(y[i] % 3) + 1])
        else:
            raise ValueError(kind)
        out.append((h, dna.tobytes(), qual.tobytes()))
    return out


def make_fastq(**kw):
    recs = make_records(**kw)
    return b"".join(h + b"\n" + d + b"\n+\n" + q + b"\n" for h, d, q in recs)
