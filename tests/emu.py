"""CPU emulation of the *device contract* (include/uqb200.h structs) - test infrastructure.

emu_stats / emu_colstats compute, with plain Python, exactly what uqb_analyze / uqb_qname_scan are
specified to return.  They serve two purposes: the host decision logic (uq_b200/host.py) can be tested
against the oracle without a GPU, and on the GPU the device's structs are compared field by field
with these, which localises a failure to one kernel."""
import re

from uq_b200 import _lib as L

NONE = L.NONE_I64


def records(fastq):
    lines = fastq.split(b"\n")
    nl = fastq.count(b"\n")
    lines = lines[:nl]
    return [lines[i:i + 4] for i in range(0, nl - nl % 4, 4)], nl


def emu_stats(fastq, ref=None, base=0):
    """ref / base: a multi-GPU shard measured against the global first QNAME line, holding records base.."""
    recs, nl = records(fastq)
    st = L.Stats()
    for i in range(256):
        st.base_single_qual[i] = -1
        st.last_count_mismatch[i] = -1
    for j in range(L.HDR_MAX + 1):
        st.first_lcp_eq[j] = st.first_lcs_eq[j] = st.first_short_prefix[j] = st.first_short_suffix[j] = NONE
    own_first = recs[0][0]
    first = own_first if ref is None else ref
    last = recs[-1][0]
    st.first_len, st.last_len = len(own_first), len(last)
    for i, b in enumerate(own_first): st.first_name[i] = b
    for i, b in enumerate(last): st.last_name[i] = b
    st.bad_first_char = -1 if own_first[:1] == b"@" else 0
    st.bad_plus_record = st.bad_len_record = -1
    dmin, dmax, maxname = None, 0, 0
    seen = {}
    first_cnt = {c: first.count(bytes([c])) for c in set(first)}
    for r, (name, seq, plus, qual) in enumerate(recs):
        if plus[:1] != b"+" and st.bad_plus_record < 0: st.bad_plus_record = r
        if len(seq) != len(qual) and st.bad_len_record < 0: st.bad_len_record = r
        dmin = len(seq) if dmin is None else min(dmin, len(seq))
        dmax = max(dmax, len(seq))
        maxname = max(maxname, len(name))
        for b, q in zip(seq, qual):
            st.base_count[b] += 1
            st.qual_count[q] += 1
            s = seen.setdefault(b, set())
            s.add(q)
        lim = min(len(name), len(first))
        lcp = 0
        while lcp < lim and name[lcp] == first[lcp]: lcp += 1
        lcs = 0
        while lcs < lim and name[len(name) - 1 - lcs] == first[len(first) - 1 - lcs]: lcs += 1
        if r + base >= 1:
            st.first_lcp_eq[lcp] = min(st.first_lcp_eq[lcp], r)
            st.first_lcs_eq[lcs] = min(st.first_lcs_eq[lcs], r)
            if lcp == len(name) and len(name) < len(first):
                st.first_short_prefix[len(name)] = min(st.first_short_prefix[len(name)], r)
            if lcs == len(name) and len(name) < len(first):
                st.first_short_suffix[len(name)] = min(st.first_short_suffix[len(name)], r)
        for c, k in first_cnt.items():
            if name.count(bytes([c])) != k:
                st.last_count_mismatch[c] = r
    for b, s in seen.items():
        st.base_single_qual[b] = next(iter(s)) if len(s) == 1 else 256
    st.dna_min, st.dna_max, st.max_name_len = dmin, dmax, maxname
    pl = sl = len(first)
    for j in range(len(first) + 1):
        if st.first_lcp_eq[j] != NONE: pl = min(pl, j)
        if st.first_lcs_eq[j] != NONE: sl = min(sl, j)
    st.prefix_len, st.suffix_len = pl, sl
    return st, len(recs)


_INT = re.compile(rb"^[+-]?[0-9]+$")


def emu_colstats(fastq, prefix_len, suffix_len, separators):
    recs, _ = records(fastq)
    seps = separators.encode("latin-1")
    ncols = len(seps) + 1
    sepset = set(seps)
    toks = [[] for _ in range(ncols)]
    bad = -1
    for r, (name, _, _, _) in enumerate(recs):
        mid = name[prefix_len:max(len(name) - suffix_len, 0)] if len(name) - suffix_len > prefix_len else b""
        cur, col, out, ok = bytearray(), 0, [], True
        for ch in mid:
            if ch in sepset:
                if col >= len(seps) or ch != seps[col]:
                    ok = False
                    break
                out.append(bytes(cur)); cur = bytearray(); col += 1
            else:
                cur.append(ch)
        out.append(bytes(cur))
        if not ok or len(out) != ncols:
            bad = r
            break
        for c in range(ncols):
            toks[c].append(out[c])
    cols = (L.ColStats * ncols)()
    dicts = {}
    if bad >= 0:
        return cols, bad, dicts
    n = len(recs)
    ncheck = 0
    t = 10000
    while t <= n - 1:
        ncheck += 1
        t *= 2
    for c in range(ncols):
        cs = cols[c]
        ints = [_INT.match(t) is not None for t in toks[c]]
        cs.all_int = 1 if all(ints) else 0
        cs.all_canonical = 1 if all(i and str(int(t)).encode() == t for i, t in zip(ints, toks[c])) else 0
        vals = [int(t) for i, t in zip(ints, toks[c]) if i]
        cs.min_val = min(vals) if vals else 2 ** 63 - 1
        cs.max_val = max(vals) if vals else -2 ** 63
        cs.min_len = min(len(t) for t in toks[c])
        cs.max_len = max(len(t) for t in toks[c])
        cs.n_checkpoints = ncheck
        first_occ = {}
        for r, t in enumerate(toks[c]):
            first_occ.setdefault(t, r)
        t = 10000
        early = False
        for k in range(ncheck):
            cs.distinct_at[k] = sum(1 for f in first_occ.values() if f <= t)
            if k == 0 and cs.distinct_at[0] > 1000:
                early = True
                break
            t *= 2
        if early:
            for k in range(1, L.MAX_CHECKPOINTS): cs.distinct_at[k] = L.U64_MAX
            cs.n_distinct = L.U64_MAX
        else:
            cs.n_distinct = len(first_occ)
            dicts[c] = [t.decode("latin-1") for t in sorted(first_occ)]
    return cols, bad, dicts
