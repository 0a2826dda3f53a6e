"""The C-ABI boundary without a GPU: libuqb200.so loads, exports every function include/uqb200.h declares,
the ctypes binding declares the same set, struct sizes agree with the header, and context creation fails
loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "uqb200.h")
LIB = os.path.join(ROOT, "uq_b200", "libuqb200.so")


@pytest.fixture(scope="module")
def lib():
    if not os.path.isfile(LIB):
        import __graft_entry__
        __graft_entry__.build()
    from uq_b200 import _lib
    return _lib.load()


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(uqb_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from uq_b200 import _lib
    names = declared_functions()
    assert len(names) >= 40
    exported = subprocess.run(["nm", "-D", "--defined-only", LIB], stdout=subprocess.PIPE, text=True).stdout
    for n in names:
        assert re.search(r"\bT %s\b" % n, exported), "%s is declared in uqb200.h but not exported" % n
        assert n in _lib.SIGNATURES, "%s is declared in uqb200.h but missing from the ctypes binding" % n
        getattr(lib, n)
    assert sorted(_lib.SIGNATURES) == names, "binding declares symbols the header does not"


def test_struct_sizes_match_the_header(tmp_path):
    """Compile a tiny C program against the header and compare sizeof() with the ctypes structures."""
    from uq_b200 import _lib
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "uqb200.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(uqb_split_info),sizeof(uqb_stats),sizeof(uqb_colstats),sizeof(uqb_colspec),sizeof(uqb_pack_params),'
                   'sizeof(uqb_decode_col),sizeof(uqb_decode_params),sizeof(uqb_synth_params));return 0;}\n')
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = list(map(int, subprocess.run([str(exe)], stdout=subprocess.PIPE, text=True, check=True).stdout.split()))
    want = [ctypes.sizeof(t) for t in (_lib.SplitInfo, _lib.Stats, _lib.ColStats, _lib.ColSpec, _lib.PackParams,
                                       _lib.DecodeCol, _lib.DecodeParams, _lib.SynthParams)]
    assert got == want


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from uq_b200.device import Context, DeviceError
    with pytest.raises(DeviceError, match="no CPU fallback"):
        Context(0)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "uq_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            text = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), fn
