"""Mid-size GPU parity (hundreds of thousands to a million reads, the BASELINE.json config shapes) against the
vectorised oracle and through size-independent properties (round trips, sortedness, uniqueness)."""
import numpy as np
import pytest

from conftest import records_multiset

pytestmark = pytest.mark.gpu
PATTERNS = ['0.1', '1.1', '2.1', '3.1', '0.2', '1.2', '2.2', '3.2']


@pytest.fixture(scope="module")
def ctx():
    from uq_b200.device import Context
    c = Context(0)
    yield c
    c.close()


def _encode(ctx, dev, stages=None, **kw):
    from uq_b200 import host
    fq = ctx.adopt_fastq(dev)
    members, cfg = host.encode_device(ctx, fq, stages=stages, **kw)
    out = members.download()
    members.free()
    fq.free()
    return out, cfg


def _dec(cfg, st):
    d = dict(st["dec"])
    return d


def test_scatter_u32_through_l2_windows(ctx):
    """out[idx[j]] = src[j] above 25 M items goes through the window-regrouped scatter (k_pairs_hist / k_pairs_regroup /
    k_pairs_apply, the kernels behind every `.key` member of the 100 M-read encode): 30 M items over four 32 MB windows"""
    n = 30_000_007
    rng = np.random.default_rng(5)
    idx = rng.permutation(n).astype(np.uint32)
    src = rng.integers(0, 1 << 32, size=n, dtype=np.uint32)
    d_idx, d_src = ctx.upload(idx), ctx.upload(src)
    out = ctx.scatter_u32(d_src, d_idx)
    got = out.download(np.uint32)
    want = np.empty(n, np.uint32)
    want[idx] = src
    assert np.array_equal(got, want)
    for a in (d_idx, d_src, out):
        a.free()


def test_config1_shape_tables_equal_closed_form(ctx):
    """1 M x 100 bp, raw tables: packed DNA / QUAL rows bit-exact vs the vectorised oracle."""
    from oracle import uq_vec as vec
    dev = ctx.synth("illumina", 1_000_000, 100, 1001)
    fastq = dev.download().tobytes()
    st = {}
    out, cfg = _encode(ctx, dev, stages=st, sort="None", raw=["DNA", "QUAL", "QNAME"], pattern=["0.1", "0.1"])
    _, dna, qual = vec.parse_fixed(fastq, 100)
    d, q = vec.pack_tables(dna, qual, st["dec"])
    assert np.array_equal(out["DNA.raw"], d)
    assert np.array_equal(out["QUAL.raw"], q)
    # QNAME columns against a numpy parse of the headers
    heads = [l.split(b":") for l in fastq.split(b"\n")[0::4][:-1]]
    for j, name in enumerate(["QNAME_1.raw", "QNAME_2.raw", "QNAME_3.raw", "QNAME_4.raw"]):
        want = np.array([int(h[3 + j]) for h in heads], dtype=np.uint64)
        assert np.array_equal(out[name].astype(np.uint64), want), name
    dev.free()


def test_config2_shape_keyed_sort_dna_equals_numpy(ctx):
    """600 k x 150 bp genome reads, --sort DNA keyed: every DNA / QUAL member bit-exact vs numpy stable sort/unique."""
    from oracle import uq_vec as vec
    n = 600_000
    dev = ctx.synth("genome", n, 150, 1002, genome=60_000, pool=100_000)
    fastq = dev.download().tobytes()
    st = {}
    out, cfg = _encode(ctx, dev, stages=st, sort="DNA")
    _, dna, qual = vec.parse_fixed(fastq, 150)
    d, q = vec.pack_tables(dna, qual, st["dec"])
    perm, key, uniq = vec.sort_unique(d)
    assert np.array_equal(out["DNA"], uniq)
    assert np.array_equal(out["DNA.key"].astype(np.int64), key[perm])
    _, qkey, quniq = vec.sort_unique(q)
    assert np.array_equal(out["QUAL"], quniq)
    assert np.array_equal(out["QUAL.key"].astype(np.int64), qkey[perm])
    dev.free()


def test_config3_shape_casava_sort_qname(ctx):
    """500 k CASAVA-1.8 reads, --sort QNAME: typed columns, sorted rows, byte-exact records after decode."""
    from uq_b200 import host
    n = 500_000
    dev = ctx.synth("casava", n, 100, 1003)
    fastq = dev.download().tobytes()
    out, cfg = _encode(ctx, dev, sort="QNAME")
    assert cfg["QNAME_prefix"] == "@EAS139:136:FC706VJ:" and cfg["QNAME_separators"] == "::: :::"
    fmts = [(c["format"], c["dtype"]) for c in cfg["QNAME_columns"]]
    assert fmts == [("integers", "uint8"), ("integers", "uint16"), ("integers", "uint16"), ("integers", "uint32"),
                    ("integers", "uint8"), ("mapping", "uint8"), ("integers", "uint8"), ("mapping", "uint8")]
    assert cfg["QNAME_columns"][5]["map"] == ["N", "Y"] and cfg["QNAME_columns"][7]["map"] == ["ATCACG", "CGATGT", "TGACCA", "TTAGGC"]
    key = out["QNAME.key"].astype(np.int64)
    assert np.all(np.diff(key) >= 0)
    cols = np.stack([out["QNAME_%d" % (i + 1)].astype(np.int64) for i in range(8)], axis=1)
    assert len(np.unique(cols, axis=0)) == len(cols)                     # unique table has no duplicates
    assert np.array_equal(cols, cols[np.lexsort(cols.T[::-1])])         # and is sorted lexicographically by column
    decoded = host.decode(out, cfg, ctx=ctx).tobytes()
    assert records_multiset(decoded) == records_multiset(fastq)
    dev.free()


@pytest.mark.parametrize("pad", [False, True])
@pytest.mark.parametrize("notricks", [False, True])
def test_config4_shape_pattern_sweep(ctx, pad, notricks):
    """200 k reads: all 8 layouts x --pad x --notricks; streams bit-exact vs numpy rot90 of the closed-form tables."""
    from oracle import uq_vec as vec
    from oracle.uq_literal import apply_pattern
    from uq_b200 import host
    dev = ctx.synth("illumina", 200_000, 100, 1004)
    fastq = dev.download().tobytes()
    _, dna, qual = vec.parse_fixed(fastq, 100)
    want_bits = {(False, False): (2, 6), (False, True): (3, 6), (True, False): (2, 8), (True, True): (4, 8)}[(pad, notricks)]
    d = q = None
    for i, p in enumerate(PATTERNS):
        pq = PATTERNS[(i + 5) % 8]
        st = {}
        out, cfg = _encode(ctx, dev, stages=st, sort="None", raw=["DNA", "QUAL", "QNAME"], pattern=[p, pq], pad=pad, notricks=notricks)
        assert (cfg["bits_per_base"], cfg["bits_per_quality"]) == want_bits
        if d is None:
            d, q = vec.pack_tables(dna, qual, st["dec"])
        for name, tab, pat in (("DNA.raw", d, p), ("QUAL.raw", q, pq)):
            want = apply_pattern(tab, pat)
            got = out[name]
            assert got.shape == want.shape and got.flags.f_contiguous == want.flags.f_contiguous
            assert np.array_equal(got, want), (name, pat)
        if i in (3, 6):
            assert host.decode(out, cfg, ctx=ctx).tobytes() == fastq
    dev.free()


def test_config5_shape_variable_length_roundtrip_20k(ctx):
    from oracle import synth
    from uq_b200 import host
    tab = synth.ont_length_table(1000, 20000)
    dev = ctx.synth("ont", 20_000, (1000, 20000), 1005, len_table=tab)
    fastq = dev.download().tobytes()
    out, cfg = _encode(ctx, dev, sort="QUAL")
    assert cfg["variable_read_lengths"] is True
    width = out["QUAL"].shape[1]
    assert width == -(-(cfg["bits_per_quality"] * (cfg["dna_max"] + 1)) // 8)
    v = np.ascontiguousarray(out["QUAL"]).view("V%d" % width).reshape(-1)
    assert np.array_equal(np.sort(v), v) and len(np.unique(v)) == len(v)
    assert np.all(np.diff(out["QUAL.key"].astype(np.int64)) >= 0)
    decoded = host.decode(out, cfg, ctx=ctx).tobytes()
    assert records_multiset(decoded) == records_multiset(fastq)
    dev.free()


def test_full_size_roundtrip_on_device(ctx):
    """BASELINE full size (100 M reads x 150 bp = 34 GB, offsets beyond 4 GiB): encode -> decode entirely in HBM,
    the decoded text must equal the input byte for byte (compared on the device)."""
    from uq_b200 import host
    n = 100_000_000
    used, free, total = ctx.mem_info()
    if free < 120 << 30:
        pytest.skip("needs ~100 GB of free HBM")
    dev = ctx.synth("genome", n, 150, 1002, genome=10_000_000, pool=n // 5)
    assert dev.nbytes > 8 << 30          # far beyond 32-bit offsets
    fq = ctx.adopt_fastq(dev)
    st = {}
    members, cfg = host.encode_device(ctx, fq, sort="None", raw=["DNA", "QUAL", "QNAME"], stages=st)
    assert cfg["reads"] == n and cfg["dna_max"] == 150 and cfg["N_qual"] == {"N": 0}
    text = host.decode_device(ctx, st["dna"], st["qual"], st["cols"], cfg)
    assert text.nbytes == dev.nbytes
    assert text.first_difference(dev) == -1
    # and the comparison itself detects a difference
    other = ctx.synth("genome", 1000, 150, 1003, genome=10_000_000, pool=200)
    ref = ctx.synth("genome", 1000, 150, 1002, genome=10_000_000, pool=200)
    assert other.first_difference(ref) >= 0
    for a in (text, other, ref):
        a.free()
    members.free(); fq.free(); dev.free()


# ---- BASELINE.json full sizes through size-independent properties (everything stays in HBM) ----------------------
def _logical_tables(ctx, members, cfg):
    """keyed / raw members in HBM -> logical DNA and QUAL tables and QNAME columns in record order (uq.py:947-973)"""
    it = members.items
    def table(name):
        if name + ".raw" in it:
            return it[name + ".raw"][0], False
        return ctx.gather_rows(it[name][0], it[name + ".key"][0]), True
    dna, fd = table("DNA")
    qual, fq_ = table("QUAL")
    cols, fc = [], []
    for meta in cfg["QNAME_columns"]:
        nm = meta["name"]
        if nm + ".raw" in it:
            cols.append(it[nm + ".raw"][0]); fc.append(False)
        else:
            cols.append(ctx.gather_rows(it[nm][0], it["QNAME.key"][0])); fc.append(True)
    return dna, qual, cols, [a for a, f in [(dna, fd), (qual, fq_)] + list(zip(cols, fc)) if f]


def _assert_idempotent(ctx, dev, **opts):
    """encode -> decode -> encode gives the same container, member by member, compared on the device; the decoded
    text has the size of the input.  (With --sort the decoded records come out in sorted order, so the second encode
    sees a different file with the same multiset of records - its sorted / keyed output must not change.)"""
    from uq_b200 import host
    fq = ctx.adopt_fastq(dev)
    m1, cfg1 = host.encode_device(ctx, fq, **opts)
    fq.free()
    dna, qual, cols, tmp = _logical_tables(ctx, m1, cfg1)
    text = host.decode_device(ctx, dna, qual, cols, cfg1)
    for a in tmp:
        a.free()
    assert text.nbytes == dev.nbytes
    fq2 = ctx.adopt_fastq(text)
    m2, cfg2 = host.encode_device(ctx, fq2, **opts)
    fq2.free()
    for k in ("reads", "bases", "qualities", "N_qual", "bits_per_base", "bits_per_quality", "dna_max", "variable_read_lengths",
              "QNAME_prefix", "QNAME_suffix", "QNAME_separators", "base_distribution", "qual_distribution"):
        assert cfg2[k] == cfg1[k], k
    assert [(c["format"], c["dtype"]) for c in cfg2["QNAME_columns"]] == [(c["format"], c["dtype"]) for c in cfg1["QNAME_columns"]]
    assert sorted(m1.items) == sorted(m2.items)
    for name in m1.items:
        a, b = m1.items[name][0], m2.items[name][0]
        assert (a.n, a.width) == (b.n, b.width), name
        assert a.first_difference(b) == -1, name
    return m1, cfg1, m2, text


def rows_strictly_ascending(t, chunk=4_000_000):
    """memcmp order of the rows of a uint8 table, strictly increasing (numpy has no ordering for void dtypes: the rows
    are compared as big-endian 64-bit words, most significant first)"""
    n, w = t.shape
    pad = (-w) % 8
    for s0 in range(0, n, chunk):
        blk = t[max(s0 - 1, 0):s0 + chunk]
        a = np.zeros((len(blk), w + pad), dtype=np.uint8)
        a[:, :w] = blk
        k = a.view(">u8")
        lt = np.zeros(len(blk) - 1, dtype=bool)
        eq = np.ones(len(blk) - 1, dtype=bool)
        for j in range(k.shape[1]):
            lt |= eq & (k[:-1, j] < k[1:, j])
            eq &= k[:-1, j] == k[1:, j]
        if not bool(lt.all()):
            return False
    return True


def row_hashes(t, chunk=8_000_000):
    """order-independent fingerprint material: one 64-bit hash per row"""
    n, w = t.shape
    pad = (-w) % 8
    mult = np.array([0x9E3779B97F4A7C15, 0xC2B2AE3D27D4EB4F, 0x165667B19E3779F9, 0x27D4EB2F165667C5] * 8, dtype=np.uint64)
    out = np.empty(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        for s0 in range(0, n, chunk):
            blk = t[s0:s0 + chunk]
            a = np.zeros((len(blk), w + pad), dtype=np.uint8)
            a[:, :w] = blk
            k = a.view("<u8")
            h = np.zeros(len(blk), dtype=np.uint64)
            for j in range(k.shape[1]):
                h = (h ^ k[:, j]) * mult[j % len(mult)]
                h ^= h >> np.uint64(29)
            out[s0:s0 + len(blk)] = h
    return out


def test_config1_full_size_sort_dna_keyed(ctx):
    """BASELINE configs[1] - the benched workload - at its full size: 100 M reads x 150 bp, --sort DNA, keyed tables.
      * the packed DNA / QUAL rows of the first 5 M reads equal the vectorised oracle (closed form, bit exact);
      * both unique tables are strictly ascending in memcmp order, DNA.key is non-decreasing, every key < table size;
      * unique[key] (the sorted rows) is a permutation of the packed table (multiset of 64-bit row hashes);
      * encode -> decode -> encode reproduces every member (compared on the device)."""
    from oracle import uq_vec as vec
    from uq_b200 import host
    n = 100_000_000
    used, free, total = ctx.mem_info()
    if total - used < 150 << 30:
        pytest.skip("needs ~150 GB of free HBM")
    dev = ctx.synth("genome", n, 150, 1002, genome=10_000_000, pool=n // 5)
    # ---- pack parity on the head of the file ----
    fq = ctx.adopt_fastq(dev)
    st = {}
    m0, cfg0 = host.encode_device(ctx, fq, sort="DNA", stages=st)
    head_reads = 5_000_000
    offs = fq.line_offsets(4 * head_reads, 1)
    head = fq.download(0, int(offs[0])).tobytes()
    dna_rows = st["dna"].download()                      # 3.8 GB, record order
    qual_head = ctx.gather_rows(st["qual"], ctx.upload(np.arange(head_reads, dtype=np.uint32))).download()
    lines_per, start = 4 * 1_000_000, 0
    for s0 in range(0, head_reads, 1_000_000):           # 1 M reads at a time (the bit expansion of the oracle is 48x)
        end = _nth_newline(head, start, lines_per)
        _, dna, qual = vec.parse_fixed(head[start:end], 150)
        d, q = vec.pack_tables(dna, qual, st["dec"])
        assert np.array_equal(dna_rows[s0:s0 + 1_000_000], d), "DNA rows %d.." % s0
        assert np.array_equal(qual_head[s0:s0 + 1_000_000], q), "QUAL rows %d.." % s0
        start = end
    del head, qual_head
    # ---- sort / unique properties ----
    out = {k: m0.items[k][0].download(dtype=np.dtype(m0.items[k][2]) if m0.items[k][1] == "vector" else np.uint8)
           for k in ("DNA", "QUAL", "DNA.key", "QUAL.key")}
    assert rows_strictly_ascending(out["DNA"]) and rows_strictly_ascending(out["QUAL"])
    key = out["DNA.key"]
    assert key.dtype == np.uint32 and len(key) == n
    assert bool(np.all(key[1:] >= key[:-1])) and int(key[-1]) == len(out["DNA"]) - 1 and int(key[0]) == 0
    assert int(out["QUAL.key"].max()) == len(out["QUAL"]) - 1
    assert np.array_equal(np.unique(key), np.arange(len(out["DNA"]), dtype=np.uint32))      # every unique row is used
    hs = np.sort(row_hashes(out["DNA"])[key])
    hp = np.sort(row_hashes(dna_rows))
    assert np.array_equal(hs, hp), "DNA[DNA.key] is not a permutation of the packed table"
    del dna_rows, out, hs, hp
    keep = {id(a) for a, _, _ in m0.items.values()}
    for a in [st["dna"], st["qual"]] + st["cols"]:
        if id(a) not in keep:
            a.free()
    m0.free(); fq.free()
    # ---- idempotence ----
    m1, cfg, m2, text = _assert_idempotent(ctx, dev, sort="DNA")
    assert cfg["reads"] == n and cfg["bits_per_base"] == 2 and cfg["bits_per_quality"] == 6
    m1.free(); m2.free(); text.free(); dev.free()


def _nth_newline(buf, start, count):
    """offset just behind the count-th newline at or after `start`"""
    a = np.frombuffer(buf, dtype=np.uint8, offset=start)
    nl = np.flatnonzero(a[:count * 200] == 10)           # 4 lines of a 150 bp record: about 85 bytes per line
    return start + int(nl[count - 1]) + 1


def test_config3_full_size_sort_qname(ctx):
    """50 M CASAVA-1.8 reads, --sort QNAME: column typing, sorted unique QNAME table, monotone key, idempotence."""
    n = 50_000_000
    used, free, total = ctx.mem_info()
    if total - used < 100 << 30:                 # the arena keeps what it mapped: room = total - bytes in use
        pytest.skip("needs ~100 GB of free HBM")
    dev = ctx.synth("casava", n, 100, 1003)
    m1, cfg, m2, text = _assert_idempotent(ctx, dev, sort="QNAME")
    assert cfg["reads"] == n and cfg["QNAME_prefix"] == "@EAS139:136:FC706VJ:" and cfg["QNAME_separators"] == "::: :::"
    assert [(c["format"], c["dtype"]) for c in cfg["QNAME_columns"]] == [
        ("integers", "uint8"), ("integers", "uint16"), ("integers", "uint16"), ("integers", "uint32"),
        ("integers", "uint8"), ("mapping", "uint8"), ("integers", "uint8"), ("mapping", "uint8")]
    out = {k: m1.items[k][0].download(dtype=np.dtype(m1.items[k][2]) if m1.items[k][1] == "vector" else np.uint8)
           for k in ["QNAME.key"] + ["QNAME_%d" % (i + 1) for i in range(8)]}
    key = out["QNAME.key"]
    assert np.all(key[1:] >= key[:-1])                                   # --sort QNAME: keys in sorted order
    # the unique table is strictly increasing in column-lexicographic order
    lt = np.zeros(len(out["QNAME_1"]) - 1, dtype=bool)
    eq = np.ones(len(lt), dtype=bool)
    for i in range(8):
        c = out["QNAME_%d" % (i + 1)].astype(np.int64)
        lt |= eq & (c[:-1] < c[1:])
        eq &= c[:-1] == c[1:]
    assert bool(np.all(lt)) and not bool(np.any(eq))
    m1.free(); m2.free(); text.free(); dev.free()


@pytest.mark.parametrize("pad,notricks", [(False, False), (True, True)])
def test_config4_full_size_layouts(ctx, pad, notricks):
    """20 M reads: each of the 8 layouts written by the encode is undone on the device and compared with the packed
    table; two of them are decoded back to the input text."""
    from uq_b200 import host
    n = 20_000_000
    dev = ctx.synth("illumina", n, 100, 1004)
    for i, p in enumerate(PATTERNS):
        pq = PATTERNS[(i + 3) % 8]
        fq = ctx.adopt_fastq(dev)
        st = {}
        m, cfg = host.encode_device(ctx, fq, sort="None", raw=["DNA", "QUAL", "QNAME"], pattern=[p, pq], pad=pad, notricks=notricks, stages=st)
        fq.free()
        for name, tab, pat in (("DNA.raw", st["dna"], p), ("QUAL.raw", st["qual"], pq)):
            back = ctx.unlayout(m.items[name][0], tab.n, tab.width, pat)
            assert back.first_difference(tab) == -1, (name, pat)
            back.free()
        if i in (2, 5):
            text = host.decode_device(ctx, st["dna"], st["qual"], st["cols"], cfg)
            assert text.nbytes == dev.nbytes and text.first_difference(dev) == -1
            text.free()
        keep = {id(a) for a, _, _ in m.items.values()}
        for a in [st["dna"], st["qual"]] + st["cols"]:
            if id(a) not in keep:
                a.free()
        m.free()
    dev.free()


def test_config5_full_size_variable_length(ctx):
    """1 M reads of 1-20 kb (12.7 GB of FASTQ, 5 GB + 17.5 GB tables), --sort QUAL: idempotence on the device."""
    from oracle import synth
    used, free, total = ctx.mem_info()
    if total - used < 150 << 30:
        pytest.skip("needs ~150 GB of free HBM")
    tab = synth.ont_length_table(1000, 20000)
    dev = ctx.synth("ont", 1_000_000, (1000, 20000), 1005, len_table=tab)
    assert dev.nbytes > 8 << 30              # log-uniform lengths: about 12.7 GB of FASTQ, offsets far beyond 32 bits
    m1, cfg, m2, text = _assert_idempotent(ctx, dev, sort="QUAL")
    assert cfg["variable_read_lengths"] is True and cfg["reads"] == 1_000_000
    key = m1.items["QUAL.key"][0].download(dtype=np.uint32)
    assert np.all(key[1:] >= key[:-1])
    m1.free(); m2.free(); text.free(); dev.free()
