"""CPU checks of the drop-in boundary: the shared library loads, exports every symbol the header
declares, the ctypes mirrors match the C struct sizes, and the product package never imports the oracle."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mocat_b200.h")
LIB = os.path.join(ROOT, "mocat_b200", "libmocat_b200.so")


@pytest.fixture(scope="module")
def built():
    if not os.path.exists(LIB):
        subprocess.run(["bash", os.path.join(ROOT, "mocat_b200", "csrc", "build.sh")], check=True)
    return LIB


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built):
    dll = ctypes.CDLL(built)
    names = _declared()
    assert len(names) >= 15
    missing = [n for n in names if not hasattr(dll, n)]
    assert not missing, f"declared in include/mocat_b200.h but not exported: {missing}"


def test_python_signatures_cover_header(built):
    from mocat_b200 import _lib
    names = set(_declared())
    assert names == set(_lib.SIGNATURES), names ^ set(_lib.SIGNATURES)


def test_struct_sizes_match_c(built, tmp_path):
    from mocat_b200 import _lib
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "mocat_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(mb_control),sizeof(mb_hist),sizeof(mb_target),sizeof(mb_move),sizeof(mb_temper),'
                   'sizeof(mb_ssm),sizeof(mb_gk),sizeof(mb_teki),sizeof(mb_teki_prm));return 0;}')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(s) for s in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    py = [ctypes.sizeof(c) for c in (_lib.Control, _lib.Hist, _lib.Target, _lib.Move, _lib.Temper, _lib.SSM, _lib.GK, _lib.Teki, _lib.TekiPrm)]
    assert sizes == py


def test_no_cpu_fallback_and_no_oracle_in_product(built):
    import torch
    from mocat_b200 import _lib
    lib = _lib.Library()
    assert lib.dll.mb_abi_version() == 1
    if not torch.cuda.is_available():
        with pytest.raises(_lib.MocatB200Error):
            lib.ctx()                                   # fails loudly without a GPU
    for dirpath, _, files in os.walk(os.path.join(ROOT, "mocat_b200")):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports the oracle"


def test_integration_doc_names_every_entry_point():
    """INTEGRATION.md must say, for every exported symbol, which reference function it stands in for"""
    import os
    from mocat_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    missing = [s for s in _lib.SIGNATURES if s not in text]
    assert not missing, missing
