import io
import json
import lzma
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_manifest():
    with open(os.path.join(GOLDEN, "manifest.json")) as f:
        return json.load(f)["cases"]


def golden_case(name):
    """-> (fastq bytes, uQ container bytes, kwargs for encode())."""
    fq = lzma.open(os.path.join(GOLDEN, name + ".fastq.xz")).read()
    uq = lzma.open(os.path.join(GOLDEN, name + ".uQ.xz")).read()
    return fq, uq, cli_to_kwargs(load_manifest()[name]["options"])


def cli_to_kwargs(opts):
    """Reference CLI option list -> keyword arguments of encode()."""
    kw = dict(sort=None, raw=None, pattern=None, pad=False, notricks=False)
    i = 0
    while i < len(opts):
        o = opts[i]
        if o == "--sort":
            kw["sort"] = opts[i + 1]; i += 2
        elif o == "--pattern":
            kw["pattern"] = [opts[i + 1], opts[i + 2]]; i += 3
        elif o == "--raw":
            j = i + 1
            vals = []
            while j < len(opts) and not opts[j].startswith("--"):
                vals.append(opts[j]); j += 1
            kw["raw"] = vals; i = j
        elif o == "--pad":
            kw["pad"] = True; i += 1
        elif o == "--notricks":
            kw["notricks"] = True; i += 1
        else:
            raise ValueError(o)
    return kw


def records_multiset(fastq):
    ls = fastq.split(b"\n")[:-1]
    return sorted(tuple(ls[i:i + 4]) for i in range(0, len(ls), 4))


def assert_members_equal(got, want, context=""):
    """Bit-exact comparison of two name->ndarray member dicts, including dtype, shape and
    memory order (what numpy.save would put in the NPY header)."""
    import numpy as np
    assert sorted(got) == sorted(want), "%s member names differ: %s vs %s" % (context, sorted(got), sorted(want))
    for k in sorted(want):
        a, b = got[k], want[k]
        assert a.dtype == b.dtype, "%s %s dtype %s vs %s" % (context, k, a.dtype, b.dtype)
        assert a.shape == b.shape, "%s %s shape %s vs %s" % (context, k, a.shape, b.shape)
        fa = a.flags.f_contiguous and not a.flags.c_contiguous
        fb = b.flags.f_contiguous and not b.flags.c_contiguous
        assert fa == fb, "%s %s fortran_order %s vs %s" % (context, k, fa, fb)
        if not np.array_equal(a, b):
            bad = np.argwhere(np.asarray(a) != np.asarray(b))
            raise AssertionError("%s %s: %d elements differ, first at %s (got %r want %r)" % (
                context, k, len(bad), bad[0], a[tuple(bad[0])], b[tuple(bad[0])]))


def assert_config_equal(got, want, context=""):
    g, w = json.loads(json.dumps(got)), json.loads(json.dumps(want))
    assert sorted(map(str, g.pop("raw"))) == sorted(map(str, w.pop("raw"))), context + " raw"
    for k in sorted(set(g) | set(w)):
        assert g.get(k) == w.get(k), "%s config[%s]: %r vs %r" % (context, k, g.get(k), w.get(k))


def hiseq_like_fastq(n=3000, length=36, seed=4):
    """N always carries '#', and '#' also sits on other (low quality) bases: the N-trick's "new quality" branch
    (uq.py:489-494)."""
    import random
    rng = random.Random(seed)
    out = []
    for i in range(n):
        dna = [rng.choice("ACGT") for _ in range(length)]
        qual = [rng.choice("IH5A#") if rng.random() < 0.9 else "#" for _ in range(length)]
        for p in range(length):
            if rng.random() < 0.03:
                dna[p] = "N"; qual[p] = "#"
        out.append("@HS:%d:%d\n%s\n+\n%s\n" % (1 + i % 4, 1000 + i, "".join(dna), "".join(qual)))
    return "".join(out).encode()
