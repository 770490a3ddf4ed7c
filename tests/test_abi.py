"""CPU tier: the C-ABI library builds, loads without a GPU, and exports every symbol include/tetwild_gpu.h declares.
No compute call is made here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "tetwild_gpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(twg_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_header():
    from tetwild_b200 import build
    build.build()
    import tetwild_b200 as tw
    L = tw.load_library()
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(L, s), "libtetwild_gpu.so does not export %s" % s
    assert b"sm_100a" in L.twg_version()


def test_no_cpu_fallback_without_gpu():
    """Without a usable device the context must refuse loudly (never compute on the CPU)."""
    import torch
    import tetwild_b200 as tw
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tier")
    with pytest.raises(tw.TetWildGPUError):
        tw.Context(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "tetwild_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in src and "from oracle" not in src and "tw_oracle.h" not in src and "liboracle" not in src, f


def test_docs_name_only_declared_entry_points():
    """INTEGRATION.md and DESIGN.md must not drift from the header: every twg_* name they mention is declared (wildcards like
    twg_mesh_* and bracketed suffixes like twg_x[_dev] are expanded against the header)."""
    syms = set(declared_symbols())
    types = {"twg_ctx", "twg_surface", "twg_winding", "twg_mesh"}
    for doc in ("INTEGRATION.md", "DESIGN.md", "README.md"):
        txt = open(os.path.join(ROOT, doc)).read()
        for m in re.finditer(r"\b(twg_[a-z0-9_]+)(\[(_[a-z]+)\])?(\*)?", txt):
            name, opt, star = m.group(1), m.group(3), m.group(4)
            if name in types or name.rstrip("_") in types:
                continue
            if star or name.endswith("_"):
                assert any(s.startswith(name) for s in syms), "%s: nothing declared matches %s*" % (doc, name)
                continue
            assert name in syms or any(s.startswith(name + "_") for s in syms), "%s mentions %s, which include/tetwild_gpu.h does not declare" % (doc, name)
            if opt:
                assert name + opt in syms, "%s mentions %s%s, not declared" % (doc, name, opt)
