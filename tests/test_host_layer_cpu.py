"""CPU-side tests of the C++ host layer above the C ABI (include/*.h, libflashblas_b200.so, drivers/): the
reference-facing types and helpers behave as the reference's (include/pointers/pointer.h:15-60,
include/pointers/allocator.h:19-59, include/lib_funcs.h:24-128), the drivers keep the positional CLIs, and the
kernels fail with -1 instead of computing on the CPU or exiting."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "blas-on-flash_b200"


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    if not (ROOT / "build" / "gemm").exists() or not g.HOSTLIB.exists():
        g.build()
    return g


def test_flash_ptr_allocator_and_sync_helpers(built, tmp_path):
    exe = tmp_path / "host_layer_check"
    cmd = [built._host_cxx(), "-O1", "-std=c++17", "-I", str(ROOT / "include"), "-o", str(exe),
           str(ROOT / "tests" / "cpp" / "host_layer_check.cpp"), f"-L{PKG}", "-lflashblas_b200", "-lbof_b200",
           f"-Wl,-rpath,{PKG}", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    work = tmp_path / "mnt"; work.mkdir()
    r = subprocess.run([str(exe), str(work)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "HOST_LAYER_OK" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("exe,nargs", [("gemm", 14), ("csrmm", 12), ("csrmm_pmem", 12), ("csrgemv", 8), ("csrcsc", 8),
                                       ("kmeans", 5)])
def test_driver_usage(built, exe, nargs):
    """A wrong argument count prints the reference's positional usage and exits 2 before touching the GPU."""
    r = subprocess.run([str(ROOT / "build" / exe), "only-one-argument"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 2
    assert r.stderr.startswith("usage : " + exe) or ("usage : " in r.stderr and exe.split("_")[0] in r.stderr)
    # the usage line names as many positional arguments as the reference's CLI takes
    assert r.stderr.count("<") >= nargs
