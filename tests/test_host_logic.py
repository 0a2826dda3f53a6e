"""Host decision logic (uq_b200/host.py) driven by the CPU emulation of the device contract,
checked against the literal oracle - no GPU needed."""
import random

import pytest

from conftest import golden_case, load_manifest
from emu import emu_colstats, emu_stats
from oracle import uq_literal as lit
from uq_b200 import host

CASES = sorted(load_manifest())


def host_decisions(fq, pad=False, notricks=False):
    st, n = emu_stats(fq)
    prefix, suffix, seps = host.derive_qname_layout(st, n)
    dec = host.decide_alphabets(st, notricks=notricks, pad=pad)
    cols, bad, dicts = emu_colstats(fq, len(prefix), len(suffix), seps)
    assert bad < 0
    columns = host.decide_columns(cols, n, lambda i: dicts[i])
    return prefix, suffix, seps, dec, columns


@pytest.mark.parametrize("name", CASES)
def test_decisions_match_oracle_on_golden(name):
    fq, _, kw = golden_case(name)
    stages = {}
    lit.encode(fq, stages=stages, **kw)
    prefix, suffix, seps, dec, columns = host_decisions(fq, pad=kw["pad"], notricks=kw["notricks"])
    assert (prefix, suffix, seps) == (stages["p1"]["prefix"], stages["p1"]["suffix"], stages["dec"]["separators"])
    for k in ("bases", "qualities", "N_qual", "bits_per_base", "bits_per_quality", "variable_read_lengths",
              "dna_bytes", "qual_bytes", "dna_max", "base_distribution", "qual_distribution"):
        assert dec[k] == stages["dec"][k], k
    assert columns == stages["columns"]


def _fastq(names):
    return b"".join(n + b"\nACGT\n+\nIIII\n" for n in names)


def _rand_names(rng, style):
    n = rng.randint(2, 40)
    out = []
    for i in range(n):
        if style == 0:
            out.append(b"@M:%d:%d %d/1" % (rng.randint(1, 3), rng.randint(0, 3000), rng.randint(5, 15)))
        elif style == 1:      # separator-like characters that are only sometimes constant
            out.append(b"@r%s_%d:%s" % (b"_" * rng.randint(0, 1), rng.randint(0, 99), rng.choice([b"a:b", b"ab", b"a_b"])))
        elif style == 2:      # shared long prefixes shrinking over time, adversarial for the order dependence
            out.append(b"@abc:def:" [: rng.randint(3, 9)] + rng.choice([b"x:1", b"y_2:", b":z", b"q"]) + b"#%d" % rng.randint(0, 5))
        else:
            out.append(b"@" + bytes(rng.choice(b"ab:_ 12") for _ in range(rng.randint(1, 8))))
    return out


def test_separator_class_that_changes_meaning_is_refused():
    """uq.py:433-436 pastes the separators into a regex character class unescaped: ':-_' is a RANGE that takes in the
    capitals.  A class that does not behave like the set it was meant to be is refused (Q11); a harmless one passes."""
    bad = _fastq([b"@x%d:%d-%d_%d" % (i, i, i + 1, i + 2) for i in range(1, 30)])       # separators ':', '-', '_' in that order
    st, n = emu_stats(bad)
    with pytest.raises(host.UQError, match="Q11|unusual"):
        host.derive_qname_layout(st, n)
    ok = _fastq([b"@ab_%d-%d" % (i, 2 * i) for i in range(1, 30)])                       # class '_-' : the dash is last, a literal
    st, n = emu_stats(ok)
    prefix, suffix, seps = host.derive_qname_layout(st, n)
    assert seps in ("-", "_-") and prefix.startswith("@ab")


@pytest.mark.parametrize("style", [0, 1, 2, 3])
def test_separator_logic_equals_sequential_reference(style):
    """The closed form derived from (first_lcp_eq, last_count_mismatch) reproduces the reference's
    order-dependent loop (uq.py:395-413) on random, deliberately ill-formed QNAME sets."""
    rng = random.Random(1234 + style)
    agree = errors = 0
    for _ in range(400):
        names = _rand_names(rng, style)
        fq = _fastq(names)
        lines, total = lit.split_lines(fq)
        try:
            p1 = lit.pass1(lines, total)
            want = (p1["prefix"], p1["suffix"], lit.decide(p1)["separators"])
        except (lit.UQError, IndexError, Exception) as e:   # IndexError = Q8, re.error = Q6/Q11
            want = type(e)
        st, n = emu_stats(fq)
        try:
            got = host.derive_qname_layout(st, n)
        except host.UQError:
            got = host.UQError
        if isinstance(want, type):
            assert got is host.UQError, (names, want)
            errors += 1
        else:
            assert got == want, names
            agree += 1
    assert agree > (20 if style < 3 else 3) and agree + errors == 400


def test_error_messages():
    with pytest.raises(host.UQError):
        host.normalise_options(sort="banana")
    with pytest.raises(host.UQError):
        host.normalise_options(pattern=["0.1"])
    assert host.normalise_options(sort="none") == ((None,), (None,), ["0.1", "0.1"])
    assert host.key_itemsize(1) == 1 and host.key_itemsize(256) == 1 and host.key_itemsize(257) == 2
    assert host.key_itemsize(65537) == 4


def test_ntrick_new_quality_branch_records_the_symbol():
    """Q3 (uq.py:489-494): N always has '#', '#' also occurs on other bases -> code len(qualities)+1, and the host
    records which character that code stands for so that the container can be decoded."""
    from conftest import hiseq_like_fastq
    fq = hiseq_like_fastq(n=400)
    st, n = emu_stats(fq)
    dec = host.decide_alphabets(st)
    stages = {}
    lit.encode(fq, stages=stages)
    assert dec["N_qual"] == stages["dec"]["N_qual"] == {"N": len(dec["qualities"]) + 1}
    assert dec["N_qual_symbol"] == {"N": "#"}
    assert dec["bases"] == "ACGT"
    cfg = host.config_of(dec, n, "@HS:", "", ":", [], (None,), (None,), ["0.1", "0.1"])
    assert cfg["N_qual_symbol"] == {"N": "#"}
    fq2, _, _ = golden_case("c1_raw_p01_31")               # the usual branch: N owns its quality -> no extra key
    plain = host.decide_alphabets(emu_stats(fq2)[0])
    assert plain["N_qual"] and plain["N_qual_symbol"] == {}
    assert "N_qual_symbol" not in host.config_of(plain, 2, "@a:", "", ":", [], (None,), (None,), ["0.1", "0.1"])
