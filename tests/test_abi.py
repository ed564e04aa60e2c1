"""CPU tests (no GPU): the C-ABI library builds for sm_100a, loads, and exports every symbol the
headers in include/ declare.  No compute call is made here."""
import os
import re
import subprocess

import pytest

import ropebwt2_b200
from ropebwt2_b200 import binding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"static inline[^{]*\{.*?\n\}", "", src, flags=re.S)
    src = re.sub(r"#define[^\n]*(\\\n[^\n]*)*", "", src)
    names = re.findall(r"\b((?:mr|rope|rle|rb2)_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_library_builds_and_loads():
    L = ropebwt2_b200.load()
    assert os.path.exists(ropebwt2_b200.lib_path())
    assert L.rb2_device_count() >= 0


@pytest.mark.parametrize("header", ["mrope.h", "rope.h", "rle.h", "ropebwt2_b200.h"])
def test_exports_every_declared_symbol(header):
    L = ropebwt2_b200.load()
    names = declared_functions(header)
    assert names, header
    for n in names:
        assert hasattr(L, n), f"{header} declares {n} but the library does not export it"


def test_binding_lists_match_headers():
    decl = set(declared_functions("mrope.h")) | set(declared_functions("rope.h")) | \
        set(declared_functions("rle.h")) | set(declared_functions("ropebwt2_b200.h"))
    listed = set(binding.MROPE_SYMBOLS + binding.ROPE_SYMBOLS + binding.RLE_SYMBOLS + binding.RB2_SYMBOLS)
    assert decl == listed, (decl - listed, listed - decl)


def test_sass_is_sm100a():
    out = subprocess.run(["cuobjdump", "-lelf", ropebwt2_b200.lib_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_product_does_not_touch_the_oracle():
    """The product path must never import / link anything under oracle/."""
    pkg = os.path.join(ROOT, "ropebwt2_b200")
    for dirpath, _, files in os.walk(pkg):
        if "_build" in dirpath or "__pycache__" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".c", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower() or f == "synth.py", f
    out = subprocess.run(["ldd", ropebwt2_b200.lib_path()], capture_output=True, text=True).stdout
    assert "oracle" not in out and "libref" not in out


def test_mrope_struct_layout():
    """mr_get_c (inline in mrope.h) reads mrope_t::r[a]->c at offset 8 of rope_t (reference rope.h:17-20)."""
    import ctypes as C
    assert binding._MRopeStruct.r.offset == 8
    assert binding._MRopeStruct.priv.offset == 56  # the reference's mrope_t ends here (mrope.h:10-14)
    assert C.sizeof(binding._MRopeStruct) == 64


@pytest.mark.parametrize("header", ["mrope.h", "rope.h", "rle.h", "ropebwt2_b200.h"])
def test_headers_are_plain_c(header, tmp_path):
    """A maintainer of the reference includes these from C99 sources (main.c): they must compile as C,
    on their own, without warnings."""
    src = tmp_path / "t.c"
    src.write_text('#include <stdio.h>\n#include "%s"\nint main(void) { return 0; }\n' % header)
    r = subprocess.run(["gcc", "-std=gnu99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
