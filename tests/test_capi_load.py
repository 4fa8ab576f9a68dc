"""The C-ABI library loads and exports every symbol include/nupgcm_b200.h declares; the ctypes
prototypes cover all of them; without a GPU, context creation fails loudly (no fallback)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from nupgcm_b200 import lib


def _declared():
    text = open(os.path.join(ROOT, "include", "nupgcm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nupgcm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = _declared()
    assert len(names) >= 35
    dll = ctypes.CDLL(lib.LIB_PATH)
    for n in names:
        assert hasattr(dll, n), f"{n} declared in the header but not exported"


def test_ctypes_prototypes_cover_the_header():
    assert sorted(list(lib.SIGNATURES) + lib.OTHER_SYMBOLS) == _declared()


def test_version_and_error_string():
    l = lib.load()
    assert l.nupgcm_version() == 100
    assert isinstance(l.nupgcm_last_error(None), bytes)


def test_no_cpu_fallback():
    """On a box without a GPU the context must refuse to exist; toolkits must refuse CPU()."""
    import nupgcm_b200 as n
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        with pytest.raises(lib.NupgcmError, match="no CPU fallback"):
            lib.Context(0)
    with pytest.raises(NotImplementedError):
        n.InversionToolkit(n.CPU(), None, None, None)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "nupgcm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
