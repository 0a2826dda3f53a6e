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
