"""The C-ABI shared library loads without a GPU and exports every symbol include/bevgen_b200.h declares."""
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from bevgen_b200 import _lib, build
    build.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    from bevgen_b200 import _lib
    header = (ROOT / "include" / "bevgen_b200.h").read_text()
    declared = set(re.findall(r"BEVGEN_API\s+[\w\s\*]+?\b(bevgen_\w+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), f"header vs binding mismatch: {declared ^ set(_lib.SIGNATURES)}"
    for name in declared:
        assert getattr(lib, name) is not None


def test_version_and_no_gpu_error(lib):
    import torch
    from bevgen_b200 import _lib
    assert lib.bevgen_version() >= 100
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            _lib.check(lib.bevgen_init(-1), "bevgen_init")      # fails loudly, no CPU fallback


def test_struct_layout_matches_header(lib):
    """GemmArgs (ctypes) must mirror bevgen_gemm_args field for field."""
    from bevgen_b200 import _lib
    header = (ROOT / "include" / "bevgen_b200.h").read_text()
    body = header[header.index("typedef struct {"):header.index("} bevgen_gemm_args;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for stmt in body.split(";"):
        stmt = stmt.replace("typedef struct {", "").strip()
        if not stmt:
            continue
        stmt = re.sub(r"^(const\s+)?(void|float|unsigned int|int|long long)\s*\*?", "", stmt).strip()
        for part in stmt.split(","):
            names.append(re.sub(r"\[.*\]", "", part).replace("*", "").strip())
    assert names == [f[0] for f in _lib.GemmArgs._fields_]
