"""The C-ABI library loads and exports every symbol include/polychase_b200.h declares; no
compute calls (there is no GPU here).  Also checks the library fails loudly without a device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "polychase_b200.h")).read()
    return sorted(set(re.findall(r"PC_API\s+[\w\s\*]+?\b(pc_\w+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for must in ("pc_create", "pc_frame_upload_rgb8", "pc_detect", "pc_lk_pair", "pc_analyze_push_frame",
                 "pc_mesh_set", "pc_track_frame", "pc_ba_load", "pc_ba_solve", "pc_ba_normal_equations"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from polychase_b200 import build, capi
    build.build()
    lib = ctypes.CDLL(capi.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    # and the ctypes table covers the header
    assert sorted(capi.SIGNATURES) == declared_symbols()


def test_no_cpu_fallback():
    import torch
    from polychase_b200 import capi
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.PcError) as e:
        capi.Context()
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_import_oracle():
    """oracle/ is test infrastructure: nothing under polychase_b200/ may import it."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "polychase_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cc", ".h", ".cuh")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", text, re.M) or "oracle/" in text and f.endswith(".py"):
                    bad.append(f)
    assert not bad, bad
