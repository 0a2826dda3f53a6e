"""The oracle restatement against containers written by the reference program itself."""
import pytest

from conftest import (assert_config_equal, assert_members_equal, golden_case, load_manifest,
                      records_multiset)
from oracle import uq_literal as lit

CASES = sorted(load_manifest())


@pytest.mark.parametrize("name", CASES)
def test_literal_encode_matches_reference(name):
    fq, uq, kw = golden_case(name)
    want_members, want_cfg = lit.read_container(uq)
    got_members, got_cfg = lit.encode(fq, **kw)
    assert_members_equal(got_members, want_members, name)
    assert_config_equal(got_cfg, want_cfg, name)


@pytest.mark.parametrize("name", CASES)
def test_literal_decode_of_reference_container(name):
    fq, uq, kw = golden_case(name)
    members, cfg = lit.read_container(uq)
    out = lit.decode(members, cfg)
    if kw["sort"] in (None, "None"):
        assert out == fq
    else:
        assert records_multiset(out) == records_multiset(fq)


def test_container_roundtrip(tmp_path):
    fq, uq, kw = golden_case("c2_keyed_sortDNA")
    members, cfg = lit.encode(fq, **kw)
    p = tmp_path / "x.uQ"
    lit.write_container(str(p), members, cfg)
    m2, c2 = lit.read_container(str(p))
    assert_members_equal(m2, members)
    assert records_multiset(lit.decode(m2, c2)) == records_multiset(fq)


def test_readme_known_answers():
    """Sizing known-answers from the reference README transcript (README.md:158-167, 287-316)."""
    assert lit.bits_for(5, False) == 3 and -(-(3 * 36) // 8) == 14
    assert lit.bits_for(26, False) == 5 and -(-(5 * 36) // 8) == 23
    assert lit.bits_for(5, True) == 4 and lit.bits_for(26, True) == 8
