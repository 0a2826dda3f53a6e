#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ from the REAL reference program.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (uq_b200/) imports this file.

The reference (/root/reference/uq.py) is a Python-2 script and cannot run under the
Python 3 of this image.  This script reads the reference source *at run time* (it is never
copied into the repository), applies the enumerated, purely mechanical text substitutions
listed in SHIMS below, writes the result to a scratch directory,
and runs it as a subprocess on small synthetic FASTQ files.  The inputs and the .uQ
containers the reference wrote for them are stored as tests/golden/<case>.fastq.xz and
tests/golden/<case>.uQ.xz, together with tests/golden/manifest.json (options per case).

The fixtures travel to the GPU box; /root/reference does not and is never read by tests.

Run:  python oracle/make_golden.py            (only in the build container)
"""
import io
import json
import lzma
import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFERENCE = "/root/reference/uq.py"
GOLDEN = os.path.join(ROOT, "tests", "golden")

sys.path.insert(0, ROOT)

# --------------------------------------------------------------------------------------
# The shim list.  Each entry: (id, description, function(text) -> text).  Every function
# asserts that it changed the expected number of places so that a different reference
# revision cannot silently slip through.
# --------------------------------------------------------------------------------------

def _sub(text, pattern, repl, expect, flags=0):
    new, n = re.subn(pattern, repl, text, flags=flags)
    if expect is not None and n != expect:
        raise RuntimeError("shim %r: expected %s substitutions, made %d" % (pattern, expect, n))
    return new


def _split_comment(line):
    """Split a source line at the first '#' that is outside a string literal."""
    q = None
    i = 0
    while i < len(line):
        c = line[i]
        if q:
            if c == "\\":
                i += 1
            elif c == q:
                q = None
        elif c in "'\"":
            q = c
        elif c == "#":
            return line[:i].rstrip(), "   " + line[i:]
        i += 1
    return line, ""


def s1_print(text):
    """S1a: py2 print statements -> print() calls (trailing comma -> end=' ')."""
    out = []
    pat = re.compile(r"^(\s*(?:(?:if|elif) .*?:\s+|else:\s+)?)print (.*?)((?:\s*;\s*exit\(\))?)\s*$")
    n = 0
    for line in text.split("\n"):
        if line.lstrip().startswith("#"):
            out.append(line)
            continue
        code, comment = _split_comment(line)
        m = pat.match(code)
        if m:
            head, body, tail = m.groups()
            tail = tail + comment
            body = body.rstrip()
            if body.endswith(","):
                body = body[:-1] + ", end=' '"
            line = "%sprint(%s)%s" % (head, body, tail)
            n += 1
        out.append(line)
    if n < 80:
        raise RuntimeError("S1: only %d print statements converted" % n)
    return "\n".join(out)


def s1_iter(text):
    """S1b: xrange/izip/iteritems."""
    text = _sub(text, r"\bxrange\(", "range(", 3)
    text = _sub(text, r"itertools\.izip\(", "zip(", 3)
    text = _sub(text, r"\.iteritems\(\)", ".items()", 1)
    return text


def s2_intdiv(text):
    """S2: py2 integer division."""
    text = _sub(text, r"self\.total /= 4", "self.total //= 4", 1)
    text = _sub(text, r"entries_read/10", "entries_read//10", 1)
    return text


def s3_map(text):
    """S3: map() is lazy in py3 and is consumed twice."""
    return _sub(text, r"_ = map\(int,column\['map'\]\)", "_ = list(map(int,column['map']))", 2)


def s4_textmode(text):
    """S4: py2 str == bytes.  Read the FASTQ as latin-1 text with '\\n' newlines only."""
    for name, cnt in (("file_path", 2), ("args.input", 1), ("inFile", 1)):
        text = _sub(text, r"open\(%s,'rb'\)" % re.escape(name),
                    "open(%s,'r',encoding='latin-1',newline='\\\\n')" % name, cnt)
    return text


def s5_json(text):
    """S5: bytes vs str for config.json."""
    text = _sub(text, r"with open\(path,'wb'\) as f: f\.write\(json\.dumps",
                "with open(path,'w') as f: f.write(json.dumps", 1)
    return text


def s6_unique_inverse(text):
    """S6: numpy>=2 returns the inverse with the input's shape; flatten it."""
    text = _sub(text, r"(\n(\s*)table,key = numpy\.unique\(table, return_inverse=True\)[^\n]*)",
                r"\1\n\2key = key.reshape(-1)", 1)
    text = _sub(text, r"(\n(\s*)common_dtype_columns_data, columns_key = numpy\.unique\(common_dtype_columns_data, return_inverse=True\)[^\n]*)",
                r"\1\n\2columns_key = columns_key.reshape(-1)", 1)
    return text


def s7_tarload(text):
    """S7: numpy.load on tarfile members needs a real file object under numpy 2."""
    text = _sub(text, r"numpy\.load\(uq\.extractfile\(file_name\)\)",
                "numpy.load(__import__('io').BytesIO(uq.extractfile(file_name).read()))", 2)
    return text


def s8_calloc(text):
    """S8: zero-initialised row buffers (Q2: the zero-fill loop skips one byte)."""
    text = _sub(text, r"uint8_t \*malloc\(size_t size\);",
                "uint8_t *malloc(size_t size); uint8_t *calloc(size_t n, size_t size);", 2)
    text = _sub(text, r"lib\.malloc\(", "lib.calloc(1,", 4)
    return text


def s9_marker(text):
    """S9: write the variable-length marker as part of the big integer (carry into next byte)."""
    for arr, tmp, pos, done in (("dna_array", "temp_dna", "dna_byte_position", "dna_bits_done"),
                                ("qual_array", "temp_qual", "qual_byte_position", "qual_bits_done")):
        pat = (r"%s\[row\]\[%s\]\s*=\s*%s\s*\+\s*\(variable_read_lengths << \(%s\)\)\s*;\s*%s -= 1"
               % (arr, pos, tmp, done, pos))
        rep = ("_v = %s + (1 << %s)\n"
               "                while True:\n"
               "                    %s[row][%s] = _v & 255 ; %s -= 1 ; _v >>= 8\n"
               "                    if _v == 0: break\n"
               "               " % (tmp, done, arr, pos, pos))
        text = _sub(text, pat, rep, 1)
    return text


def s10_stable(text):
    """S10 (D1): pin the tie order of the four argsorts."""
    text = _sub(text, r"numpy\.argsort\(table,axis=0\)", "numpy.argsort(table,axis=0,kind='stable')", 1)
    text = _sub(text, r"numpy\.argsort\(key\)", "numpy.argsort(key,kind='stable')", 1)
    text = _sub(text, r"numpy\.argsort\(common_dtype_columns_data,axis=0\)",
                "numpy.argsort(common_dtype_columns_data,axis=0,kind='stable')", 1)
    text = _sub(text, r"numpy\.argsort\(columns_key\)", "numpy.argsort(columns_key,kind='stable')", 1)
    return text


def s11_tar_order(text):
    """S11 (Q7): deterministic member order when writing, numeric QNAME order when reading."""
    text = _sub(text, r"for f in os\.listdir\(temp_directory\): temp_out\.add\(",
                "for f in sorted(os.listdir(temp_directory)): temp_out.add(", 1)
    text = _sub(text, r"for file_name in uq\.getnames\(\):",
                "for file_name in sorted(uq.getnames(), key=lambda n: (len(n), n)):", 2)
    return text


SHIMS = [
    ("S1a", s1_print), ("S1b", s1_iter), ("S2", s2_intdiv), ("S3", s3_map), ("S4", s4_textmode),
    ("S5", s5_json), ("S6", s6_unique_inverse), ("S7", s7_tarload), ("S8", s8_calloc),
    ("S9", s9_marker), ("S10", s10_stable), ("S11", s11_tar_order),
]


def transliterate(src_text):
    for _, fn in SHIMS:
        src_text = fn(src_text)
    return src_text


def reference_program(scratch):
    """Write the shimmed reference into `scratch` (never into the repo) and return its path."""
    with open(REFERENCE, "r", encoding="latin-1") as f:
        text = f.read()
    path = os.path.join(scratch, "uq_ref_py3.py")
    with open(path, "w", encoding="latin-1") as f:
        f.write(transliterate(text))
    return path


def run_reference_encode(prog, fastq_path, out_path, temp_dir, options):
    cmd = [sys.executable, prog, "-i", fastq_path, "-o", out_path, "--temp", temp_dir] + options
    env = dict(os.environ, PYTHONIOENCODING="latin-1")
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
    return r


def run_reference_decode(prog, uq_path):
    cmd = [sys.executable, prog, "-i", uq_path, "--decode"]
    env = dict(os.environ, PYTHONIOENCODING="latin-1")
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
    return r


# --------------------------------------------------------------------------------------
# Cases
# --------------------------------------------------------------------------------------

def cases():
    """(name, generator kwargs, reference CLI options, expect_decode_roundtrip)."""
    from oracle import synth
    all_pat = ['0.1', '1.1', '2.1', '3.1', '0.2', '1.2', '2.2', '3.2']
    out = []
    # cfg-1 style: N-trick fires, raw tables, each layout once (DNA pattern p, QUAL pattern rotated)
    for i, p in enumerate(all_pat):
        q = all_pat[(i + 3) % 8]
        out.append(("c1_raw_p%s_%s" % (p.replace('.', ''), q.replace('.', '')),
                    dict(kind="illumina", n=257, length=50, seed=1001 + i),
                    ["--sort", "None", "--raw", "DNA", "QUAL", "QNAME", "--pattern", p, q], True))
    # keyed tables, every sort
    for s in ("DNA", "QUAL", "QNAME"):
        out.append(("c2_keyed_sort%s" % s, dict(kind="genome", n=600, length=36, seed=1002),
                    ["--sort", s], True))
        out.append(("c2_raw_sort%s" % s, dict(kind="genome", n=600, length=36, seed=1012),
                    ["--sort", s, "--raw", "DNA", "QUAL", "QNAME"], True))
    out.append(("c2_keyed_nosort", dict(kind="genome", n=600, length=36, seed=1022), [], True))
    out.append(("c2_mixed_rawDNA_sortQUAL", dict(kind="genome", n=500, length=36, seed=1032),
                ["--sort", "QUAL", "--raw", "DNA", "--pattern", "2.2", "1.1"], True))
    out.append(("c2_mixed_rawQNAME_sortDNA", dict(kind="genome", n=500, length=36, seed=1042),
                ["--sort", "DNA", "--raw", "QNAME", "--pattern", "0.2", "3.2"], True))
    # CASAVA-1.8 headers
    out.append(("c3_casava_sortQNAME", dict(kind="casava", n=700, length=40, seed=1003),
                ["--sort", "QNAME"], True))
    out.append(("c3_casava_raw", dict(kind="casava", n=700, length=40, seed=1013),
                ["--sort", "None", "--raw", "DNA", "QUAL", "QNAME"], True))
    # pad / notricks sweep
    for pad in (False, True):
        for nt in (False, True):
            opts = ["--sort", "None", "--raw", "DNA", "QUAL", "QNAME", "--pattern", "1.2", "3.1"]
            if pad: opts.append("--pad")
            if nt: opts.append("--notricks")
            out.append(("c4_pad%d_notricks%d" % (pad, nt), dict(kind="illumina", n=300, length=50, seed=1004), opts, True))
    # N with two qualities: the trick never fires (README.md:282-284)
    out.append(("c4_twoNquals", dict(kind="illumina", n=300, length=50, seed=1014, n_two_quals=True),
                ["--sort", "None", "--raw", "DNA", "QUAL", "QNAME"], True))
    # variable length (needs S8+S9), short reads so the fixture stays small
    out.append(("c5_variable_sortQUAL", dict(kind="ont", n=120, length=(20, 90), seed=1005),
                ["--sort", "QUAL"], True))
    out.append(("c5_variable_raw", dict(kind="ont", n=120, length=(20, 90), seed=1015),
                ["--sort", "None", "--raw", "DNA", "QUAL", "QNAME", "--pattern", "3.2", "0.2"], True))
    # crosses two Pass-2 checkpoints (10000, 20000): exercises mapping->integers demotion
    out.append(("c6_checkpoints", dict(kind="illumina", n=20500, length=12, seed=1006),
                ["--sort", "None", "--raw", "DNA", "QUAL", "QNAME"], True))
    out.append(("c6_casava_checkpoints", dict(kind="casava", n=10500, length=10, seed=1016),
                ["--sort", "QNAME", "--raw", "DNA"], True))
    # offset=True integer column, QNAME suffix, non-integer mapping column
    out.append(("c7_offset_suffix_keyed", dict(kind="offset", n=400, length=30, seed=1007), ["--sort", "QNAME"], True))
    out.append(("c7_offset_suffix_raw", dict(kind="offset", n=400, length=30, seed=1017),
                ["--sort", "None", "--raw", "DNA", "QUAL", "QNAME"], True))
    return out


def main():
    from oracle import synth
    os.makedirs(GOLDEN, exist_ok=True)
    manifest = {}
    with tempfile.TemporaryDirectory(prefix="uq_golden_") as scratch:
        prog = reference_program(scratch)
        for name, gen, opts, want_rt in cases():
            fq = synth.make_fastq(**gen)
            fq_path = os.path.join(scratch, name + ".fastq")
            with open(fq_path, "wb") as f:
                f.write(fq)
            tdir = os.path.join(scratch, name + "_tmp")
            os.makedirs(tdir)
            uq_path = os.path.join(scratch, name + ".uQ")
            r = run_reference_encode(prog, fq_path, uq_path, tdir, opts)
            if r.returncode != 0 or not os.path.isfile(uq_path):
                sys.stderr.write(r.stdout.decode("latin-1")[-2000:])
                sys.stderr.write(r.stderr.decode("latin-1")[-4000:])
                raise SystemExit("reference failed on case %s" % name)
            d = run_reference_decode(prog, uq_path)
            rt = (d.returncode == 0 and d.stdout == fq) if "--sort" not in opts or opts[opts.index("--sort") + 1] == "None" else None
            if rt is None:
                # sorted: compare as multisets of records
                def recs(b):
                    ls = b.split(b"\n")[:-1]
                    return sorted(tuple(ls[i:i + 4]) for i in range(0, len(ls), 4))
                rt = d.returncode == 0 and recs(d.stdout) == recs(fq)
            with lzma.open(os.path.join(GOLDEN, name + ".fastq.xz"), "wb") as f:
                f.write(fq)
            with open(uq_path, "rb") as f, lzma.open(os.path.join(GOLDEN, name + ".uQ.xz"), "wb") as g:
                g.write(f.read())
            manifest[name] = {"gen": gen, "options": opts, "reference_decode_roundtrip": bool(rt)}
            print("%-32s encode ok, reference decode round trip: %s" % (name, rt))
            if want_rt and not rt:
                sys.stderr.write(d.stderr.decode("latin-1")[-3000:])
                raise SystemExit("reference decode did not round-trip on %s" % name)
    with open(os.path.join(GOLDEN, "manifest.json"), "w") as f:
        json.dump({"reference": "JohnLonginotto/uq uq.py (read at generation time, shims S1-S11)",
                   "shims": [s for s, _ in SHIMS], "cases": manifest}, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
