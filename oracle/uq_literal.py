"""CPU restatement of the uQ encode/decode algorithm (JohnLonginotto/uq, uq.py).

TEST INFRASTRUCTURE - the checker, never the product.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module.  The product
(uq_b200/) never does and fails loudly when its CUDA library is missing.

Parity status: PINNED.  tests/test_oracle_golden.py checks every array this module produces
against containers written by the reference program itself (oracle/make_golden.py runs the
reference source with the mechanical py2->py3 shims S1-S11 of SURVEY.md A.7 and stores its
outputs under tests/golden/).  Deviations from the *unshimmed* Python-2 program are exactly
those shims: stable argsort (S10/D1), zero-initialised rows (S8), marker carry (S9), sorted tar
members (S11).

This is a restatement organised as functions over `bytes`, one sequential Python pass per
reference pass, record at a time and symbol at a time like the original - it is deliberately
slow, it *is* the reference's algorithm, and it is what bench.py times as the CPU baseline.
Every function cites the reference lines (uq.py:<line>) it follows.
"""
import bisect
import collections
import io
import json
import re
import tarfile

import numpy

PATTERNS = ['0.1', '1.1', '2.1', '3.1', '0.2', '1.2', '2.2', '3.2']


class UQError(Exception):
    """The reference prints a message and exit()s (uq.py:48-50); the oracle raises instead."""


# ----------------------------------------------------------------------------------------
# record iteration  (uq.py:132-139, 378-385, 563-569; count via wc -l uq.py:85-87)
# ----------------------------------------------------------------------------------------

def split_lines(fastq):
    """All lines without their trailing newline, as latin-1 str (py2 str semantics, shim S4)."""
    text = fastq.decode('latin-1')
    nl = text.count('\n')
    if nl % 4 != 0:                                         # uq.py:86-87
        raise UQError('ERROR: The FASTQ file provided contains %d rows, which is not divisible by 4!' % nl)
    lines = text.split('\n')
    # anything after the last newline is invisible to `wc -l` based record counting
    return lines[:nl], nl // 4


# ----------------------------------------------------------------------------------------
# Pass 1  (uq.py:338-425) and the decisions that follow it (uq.py:427-545)
# ----------------------------------------------------------------------------------------

def pass1(lines, total):
    first = lines[0]
    if not first.startswith('@'):                           # uq.py:346
        raise UQError('ERROR: This does not look like a FASTA/FASTQ file! (first line does not start with @)')
    if not lines[2].startswith('+'):                        # uq.py:360
        raise UQError('ERROR: This does not look like a FASTA/FASTQ file! (third line does not start with +)')
    if len(lines[1]) != len(lines[3]):                      # uq.py:366
        raise UQError('ERROR: This does not look like a FASTA/FASTQ file! (SEQ and QUAL lines are not the same length)')

    prefix = first                                          # uq.py:349-352
    suffix = first
    sep_count = collections.defaultdict(int)
    rejected = set()
    lo = hi = len(lines[1])                                 # uq.py:356-357
    joint = {}                                              # static_qualities, uq.py:369-375

    def tally(seq, qual):
        for b, q in zip(seq, qual):
            per_base = joint.get(b)
            if per_base is None:
                per_base = joint[b] = collections.defaultdict(int)
            per_base[q] += 1

    tally(lines[1], lines[3])
    last_name = first
    for r in range(1, total):                               # uq.py:378-425
        name, seq, plus, qual = lines[4 * r], lines[4 * r + 1], lines[4 * r + 2], lines[4 * r + 3]
        if plus[:1] != '+':                                 # uq.py:382 (an empty line raises there)
            raise UQError('ERROR: For entry %d the third line does not start with +' % r)
        if len(seq) != len(qual):                           # uq.py:388-392
            raise UQError('ERROR: Length of DNA does not match the length of the quality scores for entry %d' % (r + 1))
        last_name = name
        if not name.startswith(prefix):                     # uq.py:395-401
            for i, ch in enumerate(prefix):
                if ch != name[i]:                           # IndexError for too-short names = Q8
                    for c in prefix[i:]:
                        if c not in rejected:
                            sep_count[c] += 1
                    prefix = prefix[:i]
                    break
        if not name.endswith(suffix):                       # uq.py:403-408
            for i, ch in enumerate(reversed(suffix)):
                if ch != name[-1 - i]:
                    suffix = '' if i == 0 else suffix[-i:]
                    break
        tail = name[len(prefix):]                           # uq.py:410-413
        for c in list(sep_count):
            if tail.count(c) != sep_count[c]:
                del sep_count[c]
                rejected.add(c)
        if len(seq) > hi: hi = len(seq)                     # uq.py:416-417
        if len(seq) < lo: lo = len(seq)
        tally(seq, qual)

    for c in list(sep_count):                               # uq.py:428-431
        k = suffix.count(c)
        if k:
            sep_count[c] -= k
            if sep_count[c] == 0:
                del sep_count[c]
    return dict(prefix=prefix, suffix=suffix, sep_count=dict(sep_count), first_name=first,
                last_name=last_name, dna_min=lo, dna_max=hi, joint=joint)


def order_separators(name, prefix, suffix, seps):
    """uq.py:433-436 - note the unescaped character class and the extra -1."""
    runs = re.findall('([' + ''.join(seps) + ']+)', name[len(prefix):-1 - len(suffix)])
    return ''.join(runs)


def bits_for(n, pad):
    """uq.py:497-503 / uq.py:534-540."""
    if n <= 4: return 2
    if n <= 8 and not pad: return 3
    if n <= 16: return 4
    if n <= 32 and not pad: return 5
    if n <= 64 and not pad: return 6
    if n <= 128 and not pad: return 7
    return 8


def decide(p1, notricks=False, pad=False):
    """Separator ordering, alphabets, N-trick, bit widths, row sizes (uq.py:433-545)."""
    seps = p1['sep_count']
    if len(seps) == 0:
        # uq.py:435 builds the regex '([]+)' which raises re.error (Q6)
        raise UQError('ERROR: no QNAME separators (the reference crashes on such files, Q6)')
    a = order_separators(p1['last_name'], p1['prefix'], p1['suffix'], seps)
    b = order_separators(p1['first_name'], p1['prefix'], p1['suffix'], seps)
    if a != b:                                              # uq.py:438-444
        raise UQError("ERROR: Sorry, the separators used in this file's QNAME/headers are so unusual/improbable ...")
    separators = a

    joint = p1['joint']
    base_graph = collections.defaultdict(int)               # uq.py:448-457
    qual_graph = collections.defaultdict(int)
    for base, per_q in joint.items():
        base_graph[base] = sum(per_q.values())
        for q, c in per_q.items():
            qual_graph[q] += c
    bases = sorted(base_graph)
    quals = sorted(qual_graph)

    n_qual = {}                                             # uq.py:476-494
    total_quals = len(quals)
    if not notricks:
        for base, per_q in joint.items():
            if len(bases) == 1:
                continue
            if len(per_q) == 1:
                bases.remove(base)
                for q, c in per_q.items():
                    if c == qual_graph[q]:
                        n_qual[base] = quals.index(q)
                    else:
                        total_quals += 1
                        n_qual[base] = total_quals          # pre-incremented on purpose (Q3)

    bpb = bits_for(len(bases), pad)
    variable = p1['dna_min'] != p1['dna_max']               # uq.py:512-513
    dna_cols = -(-(bpb * (p1['dna_max'] + variable)) // 8)  # uq.py:514-515
    bpq = bits_for(total_quals, pad)
    qual_cols = -(-(bpq * (p1['dna_max'] + variable)) // 8) # uq.py:543-544
    return dict(separators=separators, bases=''.join(bases), qualities=''.join(quals), N_qual=n_qual,
                bits_per_base=bpb, bits_per_quality=bpq, variable_read_lengths=variable,
                dna_bytes=dna_cols, qual_bytes=qual_cols, dna_max=p1['dna_max'],
                base_distribution=dict(base_graph), qual_distribution=dict(qual_graph))


# ----------------------------------------------------------------------------------------
# Pass 2  (uq.py:555-676)
# ----------------------------------------------------------------------------------------

def qname_tokens(lines, total, prefix, suffix, separators):
    """qname_reader, uq.py:557-570.  The reference slices the raw line *including* its newline
    with [len(prefix) : -1-len(suffix)]; `lines` here are newline-free, hence the +'\\n'."""
    start, end = len(prefix), -1 - len(suffix)
    rx = re.compile('(.*)'.join(separators))
    for r in range(total):
        yield re.split(rx, (lines[4 * r] + '\n')[start:end])


def _demote(columns, seen):
    """check_format, uq.py:586-602."""
    for col in columns:
        if col['format'] == 'mapping' and len(col['map']) > seen // 10:
            try:
                vals = list(map(int, col['map']))
                col['min'], col['max'] = min(vals), max(vals)
                col['format'] = 'integers'
                del col['map']
            except ValueError:
                col['format'] = 'strings'
                col['longest'] = max(map(len, col['map']))
                del col['map']


_DT_MAX = [(255, 'uint8'), (65535, 'uint16'), (4294967295, 'uint32'), (18446744073709551615, 'uint64')]


def pass2(lines, total, prefix, suffix, separators):
    ncols = None
    target = 10000
    columns = []
    last = -1
    for last, toks in enumerate(qname_tokens(lines, total, prefix, suffix, separators)):
        if ncols is None:                                   # uq.py:605-608
            ncols = len(toks)
            columns = [{'name': 'QNAME_%d' % (i + 1), 'format': 'mapping', 'map': set()} for i in range(ncols)]
        elif len(toks) != ncols:                            # uq.py:609-613, 637
            raise UQError('Encoding QNAMEs as strings has not been implimented yet.')
        for i, tok in enumerate(toks):                      # uq.py:614-633
            col = columns[i]
            if col['format'] == 'mapping':
                col['map'].add(tok)
            elif col['format'] == 'integers':
                try:
                    v = int(tok)
                    if v < col['min']: col['min'] = v
                    elif v > col['max']: col['max'] = v
                except ValueError:
                    col['format'] = 'strings'
                    col['longest'] = len(str(col['max']))
                    del col['min'], col['max']
            elif col['format'] == 'strings':
                if len(tok) > col['longest']: col['longest'] = len(tok)
        if last == target:                                  # uq.py:634-636
            _demote(columns, last)
            target *= 2
    _demote(columns, last)                                  # uq.py:638

    for col in columns:                                     # uq.py:641-673
        if col['format'] == 'mapping':
            n = len(col['map'])
            cap, col['dtype'] = next((m, d) for m, d in _DT_MAX if n <= m)
            try:
                vals = list(map(int, col['map']))
                if max(vals) - min(vals) <= cap:
                    col['format'] = 'integers'
                    col['max'], col['min'] = max(vals), min(vals)
                    col['offset'] = bool(min(vals) < 0 or max(vals) > cap)
                    del col['map']
                else:
                    col['map'] = sorted(col['map'])
            except Exception:
                col['map'] = sorted(col['map'])
        elif col['format'] == 'integers':
            span = col['max'] - col['min']
            cap, col['dtype'] = next((m, d) for m, d in _DT_MAX if span <= m)
            col['offset'] = bool(col['min'] < 0 or col['max'] > cap)
        elif col['format'] == 'strings':
            raise UQError('I havent implimented this yet')  # uq.py:672-673
    return columns


# ----------------------------------------------------------------------------------------
# Pass 3  (encoder_fixed uq.py:108-182, encoder_variable uq.py:188-254)
# ----------------------------------------------------------------------------------------

def pass3(lines, total, dec):
    """Symbol-at-a-time bit packing, walking each read backwards and flushing low bytes to
    decreasing byte positions.  Rows start zeroed (S8); the marker is added to the running
    integer (S9) instead of being forced into one uint8."""
    bases, quals, n_qual = dec['bases'], dec['qualities'], dec['N_qual']
    bb, bq = dec['bits_per_base'], dec['bits_per_quality']
    wd, wq = dec['dna_bytes'], dec['qual_bytes']
    marker = 1 if dec['variable_read_lengths'] else 0
    dna_tab = numpy.zeros((total, wd), dtype=numpy.uint8)
    qual_tab = numpy.zeros((total, wq), dtype=numpy.uint8)
    for r in range(total):
        seq = lines[4 * r + 1][::-1]                        # uq.py:135, 137
        qv = lines[4 * r + 3][::-1]
        acc_d = acc_q = 0
        nd = nq = 0
        pd, pq = wd - 1, wq - 1
        drow, qrow = dna_tab[r], qual_tab[r]
        for i in range(len(seq)):                           # uq.py:147-167
            ch = seq[i]
            cd = bases.find(ch)
            if cd >= 0:
                cq = quals.index(qv[i])
            else:                                           # tricked base: uq.py:151-153
                cd = 0
                cq = n_qual[ch]
            acc_d += cd << nd
            acc_q += cq << nq
            nd += bb
            nq += bq
            while nd > 8:                                   # strictly greater: uq.py:157
                nd -= 8
                drow[pd] = acc_d & 255
                acc_d >>= 8
                pd -= 1
            while nq > 8:                                   # uq.py:163
                nq -= 8
                qrow[pq] = acc_q & 255
                acc_q >>= 8
                pq -= 1
        acc_d += marker << nd                               # uq.py:242-243 with S9
        acc_q += marker << nq
        while True:                                         # uq.py:170-171 (+ carry byte, S9)
            drow[pd] = acc_d & 255
            pd -= 1
            acc_d >>= 8
            if acc_d == 0: break
        while True:
            qrow[pq] = acc_q & 255
            pq -= 1
            acc_q >>= 8
            if acc_q == 0: break
    return dna_tab, qual_tab


# ----------------------------------------------------------------------------------------
# Pass 4  (uq.py:717-735)
# ----------------------------------------------------------------------------------------

def pass4(lines, total, prefix, suffix, separators, columns):
    out = [numpy.zeros(total, dtype=c['dtype']) for c in columns]
    for r, toks in enumerate(qname_tokens(lines, total, prefix, suffix, separators)):
        for i, tok in enumerate(toks):
            c = columns[i]
            if c['format'] == 'mapping':
                out[i][r] = bisect.bisect_left(c['map'], tok)          # uq.py:724
            elif c['offset']:
                out[i][r] = int(tok) - c['min']                         # uq.py:725
            else:
                out[i][r] = int(tok)                                    # uq.py:726
    return out


# ----------------------------------------------------------------------------------------
# run_mix  (uq.py:739-851) and the layouts (uq.py:257-270)
# ----------------------------------------------------------------------------------------

def apply_pattern(table, pattern):
    """write_pattern, uq.py:263-270: the array object numpy.save() is handed."""
    k, order = int(pattern[0]), pattern[2]
    rot = table if k == 0 else numpy.rot90(table, k)
    return numpy.ascontiguousarray(rot) if order == '1' else numpy.asfortranarray(rot)


def mix_dna_qual(out, table, name, order, raw, pattern):
    """encode_dna_qual, uq.py:765-805.  `order`: None / False / ndarray as in the reference."""
    width = table.shape[1]
    as_void = numpy.ascontiguousarray(table).view('V%d' % width)       # uq.py:774, 785
    if raw:
        if order is not None:
            if order is False:
                order = numpy.argsort(as_void, axis=0, kind='stable').reshape(-1)   # uq.py:775 + S10
            table = table[order]
        out[name + '.raw'] = apply_pattern(table, pattern)
    else:
        uniq, key = numpy.unique(as_void, return_inverse=True)          # uq.py:786
        key = key.reshape(-1)                                            # S6
        uniq = uniq.reshape(-1, 1).view(numpy.uint8).reshape(len(uniq), width)
        key = key.astype(numpy.min_scalar_type(int(key.max())))        # uq.py:790
        if order is not None:
            if order is False:
                order = numpy.argsort(key, kind='stable')               # uq.py:796 + S10
            key = key[order]
        out[name + '.key'] = key
        out[name] = apply_pattern(uniq, pattern)
    return order


def mix_qname(out, cols, columns, order, raw):
    """encode_qname, uq.py:808-851."""
    def stacked():
        t = numpy.dstack(cols)[0]                                       # widest dtype, uq.py:814/828
        return t, t.view(','.join([str(t.dtype)] * t.shape[1]))         # structured rows
    if raw:
        if order is False:
            _, rows = stacked()
            order = numpy.argsort(rows, axis=0, kind='stable').reshape(-1)     # uq.py:816 + S10
        for c, meta in zip(cols, columns):
            out[meta['name'] + '.raw'] = c[order] if isinstance(order, numpy.ndarray) else c
    else:
        t, rows = stacked()
        uniq, key = numpy.unique(rows, return_inverse=True)             # uq.py:830
        key = key.reshape(-1)
        key = key.astype(numpy.min_scalar_type(int(key.max())))        # uq.py:832
        if order is False:
            order = numpy.argsort(key, kind='stable')                   # uq.py:833 + S10
        out['QNAME.key'] = key[order] if isinstance(order, numpy.ndarray) else key
        flat = uniq.view(t.dtype).reshape(len(uniq), t.shape[1])        # uq.py:842-844
        for i, meta in enumerate(columns):
            out[meta['name']] = flat[:, i].astype(meta['dtype'])        # uq.py:846-847
    return order if isinstance(order, numpy.ndarray) else None


def run_mix(dna, qual, cols, columns, sorted_on, raw_tables, pattern):
    """uq.py:739-753: the sorted-on table goes first and yields the permutation."""
    out = {}
    pd, pq = pattern
    if sorted_on in ('DNA', 'QUAL'):
        first = (dna, 'DNA', pd) if sorted_on == 'DNA' else (qual, 'QUAL', pq)
        second = (qual, 'QUAL', pq) if sorted_on == 'DNA' else (dna, 'DNA', pd)
        order = mix_dna_qual(out, first[0], first[1], False, first[1] in raw_tables, first[2])
        mix_dna_qual(out, second[0], second[1], order, second[1] in raw_tables, second[2])
        mix_qname(out, cols, columns, order, 'QNAME' in raw_tables)
    elif sorted_on == 'QNAME':
        order = mix_qname(out, cols, columns, False, 'QNAME' in raw_tables)
        mix_dna_qual(out, dna, 'DNA', order, 'DNA' in raw_tables, pd)
        mix_dna_qual(out, qual, 'QUAL', order, 'QUAL' in raw_tables, pq)
    else:
        mix_qname(out, cols, columns, None, 'QNAME' in raw_tables)
        mix_dna_qual(out, dna, 'DNA', None, 'DNA' in raw_tables, pd)
        mix_dna_qual(out, qual, 'QUAL', None, 'QUAL' in raw_tables, pq)
    return out


# ----------------------------------------------------------------------------------------
# whole encode, container (uq.py:681-696, 893-917)
# ----------------------------------------------------------------------------------------

def normalise_options(sort=None, raw=None, pattern=None):
    """uq.py:52-69 and 893-894.  Comparisons later are case-sensitive (Q12)."""
    if pattern is not None:
        if len(pattern) != 2: raise UQError('ERROR: There must be 2 values for --pattern!')
        if not all(p in PATTERNS for p in pattern): raise UQError('ERROR: Pattern values are incorrect!')
    if sort is not None:
        if sort.lower() not in ('dna', 'qual', 'qname', 'none'): raise UQError('ERROR: --sort value is incorrect!')
        if sort.lower() == 'none': sort = (None,)
    if raw is not None:
        if not all(x.lower() in ('dna', 'qual', 'qname', 'none') for x in raw):
            raise UQError('ERROR: --raw values are incorrect!')
        raw = set(raw)
        if 'none' in raw:
            raw.add(None); raw.discard('none')
    if sort is None: sort = (None,)
    if raw is None: raw = (None,)
    if pattern is None: pattern = ['0.1', '0.1']                        # uq.py:258
    return sort, raw, list(pattern)


def encode(fastq, sort=None, raw=None, pattern=None, pad=False, notricks=False, stages=None):
    """FASTQ bytes -> (members: name -> ndarray exactly as handed to numpy.save, config dict).
    `stages`, if a dict, receives the intermediate products (for kernel-level parity tests)."""
    sort, raw, pattern = normalise_options(sort, raw, pattern)
    lines, total = split_lines(fastq)
    p1 = pass1(lines, total)
    dec = decide(p1, notricks=notricks, pad=pad)
    columns = pass2(lines, total, p1['prefix'], p1['suffix'], dec['separators'])
    dna, qual = pass3(lines, total, dec)
    cols = pass4(lines, total, p1['prefix'], p1['suffix'], dec['separators'], columns)
    if stages is not None:
        stages.update(p1=p1, dec=dec, columns=columns, dna=dna, qual=qual, cols=cols, total=total)
    members = run_mix(dna, qual, cols, columns, sort, raw, pattern)
    config = {                                                            # uq.py:681-696, 898-900
        'base_distribution': dec['base_distribution'], 'qual_distribution': dec['qual_distribution'],
        'reads': total, 'bases': dec['bases'], 'qualities': dec['qualities'],
        'variable_read_lengths': dec['variable_read_lengths'], 'bits_per_base': dec['bits_per_base'],
        'bits_per_quality': dec['bits_per_quality'], 'N_qual': dec['N_qual'], 'dna_max': dec['dna_max'],
        'QNAME_prefix': p1['prefix'], 'QNAME_suffix': p1['suffix'], 'QNAME_separators': dec['separators'],
        'QNAME_columns': columns, 'sort': sort, 'raw': list(raw), 'pattern': pattern,
    }
    return members, config


def npy_bytes(arr):
    buf = io.BytesIO()
    numpy.save(buf, arr)
    return buf.getvalue()


def write_container(path, members, config):
    """Uncompressed tar of NPY members (no .npy suffix) + config.json, sorted names (S11)."""
    blobs = {k: npy_bytes(v) for k, v in members.items()}
    blobs['config.json'] = json.dumps(config, indent=4, sort_keys=True).encode()
    with tarfile.open(path, mode='w') as tar:
        for name in sorted(blobs):
            info = tarfile.TarInfo(name)
            info.size = len(blobs[name])
            tar.addfile(info, io.BytesIO(blobs[name]))


def read_container(path_or_bytes):
    """-> (members name -> ndarray as numpy.load returns them, config)."""
    if isinstance(path_or_bytes, (bytes, bytearray)):
        tar = tarfile.open(fileobj=io.BytesIO(path_or_bytes))
    else:
        tar = tarfile.open(path_or_bytes)
    members, config = {}, None
    for name in tar.getnames():
        data = tar.extractfile(name).read()
        if name == 'config.json':
            config = json.loads(data.decode())
        else:
            members[name] = numpy.load(io.BytesIO(data))                  # S7
    return members, config


# ----------------------------------------------------------------------------------------
# decode  (uq.py:926-1060)
# ----------------------------------------------------------------------------------------

def undo_pattern(arr, pattern):
    """load_from_tar, uq.py:943-945."""
    return arr if pattern.startswith('0.') else numpy.rot90(arr, -int(pattern[0]))


def _symbols(row, total_bits, bits):
    """split_bits, uq.py:1002-1007: the row as one big-endian integer, most significant symbol first."""
    number = 0
    for byte in row:
        number = (number << 8) | int(byte)
    mask = (1 << bits) - 1
    return [(number >> s) & mask for s in range(total_bits - bits, -bits, -bits)]


def decode(members, config):
    """members/config as read_container returns them -> FASTQ bytes."""
    pat = config['pattern']
    def table(name, p):
        if name + '.raw' in members:
            return undo_pattern(members[name + '.raw'], p)
        if name in members and name + '.key' in members:
            return undo_pattern(members[name], p)[members[name + '.key']]     # uq.py:953, 957
        raise UQError('ERROR: No %s data was found in this uQ file?!' % name)
    dna_t, qual_t = table('DNA', pat[0]), table('QUAL', pat[1])
    ncol = len(config['QNAME_columns'])
    if 'QNAME.key' in members:                                              # uq.py:962-973 (+S11)
        stack = numpy.dstack([members['QNAME_%d' % (i + 1)] for i in range(ncol)])[0][members['QNAME.key']]
    else:
        stack = numpy.dstack([members['QNAME_%d.raw' % (i + 1)] for i in range(ncol)])[0]

    bases, quals = config['bases'], config['qualities']
    back = dict((v, k) for k, v in config['N_qual'].items())                # uq.py:999
    var = config['variable_read_lengths']
    nsym = var + config['dna_max']
    bb, bq = config['bits_per_base'], config['bits_per_quality']
    seps, cols = config['QNAME_separators'], config['QNAME_columns']
    out = []
    for d_row, q_row, n_row in zip(dna_t, qual_t, stack):
        ds = _symbols(d_row, nsym * bb, bb)
        qs = _symbols(q_row, nsym * bq, bq)
        dna, qual = [], []
        for d, q in zip(ds, qs):                                            # uq.py:1034-1037
            qual.append(quals[q])          # IndexError for "new quality" codes = Q3, as the reference
            dna.append(back[q] if q in back else bases[d])
        if var:                                                             # uq.py:1039-1041
            dna = dna[1 + dna.index(bases[1]):]
            qual = qual[1 + qual.index(quals[1]):]
        name = config['QNAME_prefix']                                       # uq.py:1010-1024
        for i, c in enumerate(cols):
            if c['format'] == 'mapping':
                name += c['map'][int(n_row[i])]
            elif c['offset']:
                name += str(int(n_row[i]) + c['min'])
            else:
                name += str(int(n_row[i]))
            if i < len(seps):
                name += seps[i]
        name += config['QNAME_suffix']
        out.append(name + '\n' + ''.join(dna) + '\n+\n' + ''.join(qual) + '\n')   # uq.py:1042-1045
    return ''.join(out).encode('latin-1')
