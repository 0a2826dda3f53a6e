"""The vectorised oracle against the literal one (which is pinned to the reference's golden containers)."""
import numpy as np
import pytest

from conftest import golden_case, load_manifest
from oracle import uq_literal as lit, uq_vec as vec

FIXED = [c for c in sorted(load_manifest()) if not c.startswith("c5_")]


@pytest.mark.parametrize("name", FIXED)
def test_vec_pack_equals_literal(name):
    fq, _, kw = golden_case(name)
    st = {}
    lit.encode(fq, stages=st, **kw)
    _, dna, qual = vec.parse_fixed(fq, st["dec"]["dna_max"])
    d, q = vec.pack_tables(dna, qual, st["dec"])
    assert np.array_equal(d, st["dna"]) and np.array_equal(q, st["qual"])
    perm, key, uniq = vec.sort_unique(st["dna"])
    out = {}
    lit.mix_dna_qual(out, st["dna"], "DNA", False, False, "0.1")
    assert np.array_equal(out["DNA"], uniq) and np.array_equal(out["DNA.key"], key[perm].astype(out["DNA.key"].dtype))
