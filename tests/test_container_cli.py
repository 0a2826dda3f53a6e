"""Container format and the command line.  CPU tests check that the NPY byte streams this package
writes are identical to the reference-written tar members; the GPU test drives the CLI end to end."""
import io
import json
import lzma
import os
import subprocess
import sys
import tarfile

import pytest

from conftest import GOLDEN, ROOT, golden_case, load_manifest, records_multiset
from uq_b200 import container

CASES = sorted(load_manifest())


def _raw_members(uq_bytes):
    tar = tarfile.open(fileobj=io.BytesIO(uq_bytes))
    return {n: tar.extractfile(n).read() for n in tar.getnames()}


@pytest.mark.parametrize("name", CASES)
def test_npy_streams_are_byte_identical_to_the_reference(name):
    _, uq, _ = golden_case(name)
    raw = _raw_members(uq)
    members, config = container.read_container(uq)
    for k, arr in members.items():
        assert container.npy_bytes(arr) == raw[k], k
    assert json.loads(raw["config.json"]) == config


def test_write_then_read(tmp_path):
    _, uq, _ = golden_case("c3_casava_sortQNAME")
    members, config = container.read_container(uq)
    p = str(tmp_path / "x.uQ")
    container.write_container(p, members, config)
    assert tarfile.open(p).getnames() == sorted(list(members) + ["config.json"])
    m2, c2 = container.read_container(p)
    assert c2 == config and sorted(m2) == sorted(members)
    for k in members:
        assert container.npy_bytes(m2[k]) == container.npy_bytes(members[k])


def test_cli_parser_matches_reference_options():
    from uq_b200 import uq
    a = uq.build_parser().parse_args(["-i", "x.fastq", "--sort", "QUAL", "--raw", "DNA", "QNAME", "--pattern", "2.2", "1.1",
                                      "--pad", "--notricks", "--test", "--compressor", "xz", "--temp", "/tmp", "--peek"])
    assert (a.sort, a.raw, a.pattern, a.pad, a.notricks, a.test, a.compressor, a.peek, a.decode) == \
        ("QUAL", ["DNA", "QNAME"], ["2.2", "1.1"], True, True, True, "xz", True, False)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c2_keyed_sortDNA", "c1_raw_p12_01", "c5_variable_sortQUAL", "c7_offset_suffix_keyed", "c4_pad1_notricks1"])
def test_cli_end_to_end(tmp_path, name):
    fq, uq, kw = golden_case(name)
    opts = load_manifest()[name]["options"]
    src = tmp_path / "in.fastq"
    src.write_bytes(fq)
    out = tmp_path / "out.uQ"
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-m", "uq_b200.uq", "-i", str(src), "-o", str(out)] + opts, env=env, cwd=ROOT,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0 and out.is_file(), r.stdout.decode() + r.stderr.decode()
    got, want = _raw_members(out.read_bytes()), _raw_members(uq)
    assert sorted(got) == sorted(want)
    for k in want:
        if k != "config.json":
            assert got[k] == want[k], k                    # NPY member bytes identical to the reference's
    gc, wc = json.loads(got["config.json"]), json.loads(want["config.json"])
    gc.pop("raw"); wc.pop("raw")
    assert gc == wc
    d = subprocess.run([sys.executable, "-m", "uq_b200.uq", "-i", str(out), "--decode"], env=env, cwd=ROOT,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert d.returncode == 0, d.stderr.decode()
    if kw["sort"] in (None, "None"):
        assert d.stdout == fq
    else:
        assert records_multiset(d.stdout) == records_multiset(fq)


@pytest.mark.parametrize("name", ["c2_keyed_sortDNA", "c1_raw_p22_11", "c1_raw_p12_01", "c3_casava_raw", "c5_variable_sortQUAL"])
def test_planned_writer_is_byte_identical_to_tarfile_plus_numpy_save(tmp_path, name):
    """container.write_container (offsets planned up front, payloads written with pwrite straight from the buffers)
    produces exactly the bytes of tarfile + numpy.save - headers, padding, end blocks and all."""
    _, uq, _ = golden_case(name)
    members, config = container.read_container(uq)
    a, b = str(tmp_path / "planned.uQ"), str(tmp_path / "tarfile.uQ")
    plan = container.write_container(a, members, config)
    container.write_container_tarfile(b, members, config)
    da, db = open(a, "rb").read(), open(b, "rb").read()
    assert len(da) == len(db) == plan.total
    assert da == db
    # members read back through the memory map are the arrays themselves, in their memory order
    mm, cfg = container.read_container(a, mmap=True)
    assert cfg == config and sorted(mm) == sorted(members)
    for k in members:
        assert mm[k].dtype == members[k].dtype and mm[k].shape == members[k].shape
        assert mm[k].flags.f_contiguous == members[k].flags.f_contiguous
        assert (mm[k] == members[k]).all()


def test_plan_offsets_and_degenerate_shapes():
    import numpy as np
    ents = {"DNA.raw": (np.uint8, (1, 7), False), "QNAME_1.raw": (np.uint32, (0,), False), "QUAL.raw": (np.uint8, (5, 1), False)}
    plan = container.Plan(ents, {"reads": 1})
    assert plan.names == sorted(list(ents) + ["config.json"])
    for name in ents:
        pos, pay, nb = plan.offsets[name]
        assert pos % 512 == 0 and pay == pos + 512 + 128                   # NPY v1.0 headers of these shapes are 128 bytes
    assert plan.total % 10240 == 0
    assert container.npy_header(np.uint8, (3, 5), True) == container.npy_bytes(np.asfortranarray(np.zeros((3, 5), np.uint8)))[:128]
