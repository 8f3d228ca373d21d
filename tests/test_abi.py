"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol that
include/bof_b200.h declares, and refuses to compute without a GPU (no fallback)."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "bof_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bof_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(bof):
    lib = bof.load()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/bof_b200.h but not exported"


def test_binding_covers_header(bof):
    from bof_b200 import _capi

    assert sorted(_capi.PROTOTYPES) == declared_symbols()


def test_abi_version(bof):
    assert bof.load().bof_abi_version() == 1


def test_struct_layouts_match_header(bof):
    # bof_config: i32 i32 u64 i32 (pad) u64 u64 i32 i32 i32 i32 i32 (pad) ; bof_stats: 8 doubles + i64
    assert C.sizeof(bof.BofConfig) == 64  # 13 x 4-byte + 3 x 8-byte fields with padding
    assert C.sizeof(bof.BofStats) == 72


def test_no_cpu_fallback(bof):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(bof.BofError) as ei:
        bof.Context(device=0)
    assert ei.value.code == -4  # BOF_ENODEV
    assert "no CUDA device" in str(ei.value)


def test_null_ctx_is_rejected(bof):
    lib = bof.load()
    assert lib.bof_host_gemm(None, b"R", b"N", b"N", 1, 1, 1, 1.0, 0.0, None, None, None, 0, 0, 0) == -1
    assert lib.bof_ctx_destroy(None) == 0
    assert lib.bof_launch_count(None) == 0


def test_product_does_not_reference_oracle():
    """The product path must not import, link or execute anything under oracle/."""
    pkg = ROOT / "blas-on-flash_b200"
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.cpp")) \
            + list((ROOT / "include").rglob("*.h")):
        text = p.read_text()
        assert "import oracle" not in text and "liboracle" not in text and "orc_" not in text, p
