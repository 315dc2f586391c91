"""The C-ABI library loads and exports every entry point include/supersonic_b200.h declares
(no compute calls: this runs without a GPU)."""
import os

import pytest

from supersonic_b200 import capi


def test_header_symbols_exported(built):
    lib = capi.load()
    names = capi.declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), n
    assert lib.ssb_abi_version() == 2


def test_no_device_fails_loudly(built):
    """Without a B200 the product must refuse to run; there is no CPU path to fall back to."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.SsbError):
        capi.Context(0)


def test_plan_driver_reports_missing_device(b200):
    import numpy as np
    import torch
    from supersonic_b200 import ssplan as sp
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    cols = [sp.Column("a", sp.INT64, np.arange(10))]
    r = b200.run("(compute (plus (col a) (i64 1)) (scan 0))", [cols])
    assert r.code != 0 and "no usable B200" in r.error


def test_host_generator_matches_definition(built):
    import ctypes as C
    import numpy as np
    lib = capi.load()
    out = np.zeros(16, dtype=np.uint64)
    lib.ssb_generate_host(out.ctypes.data, 16, 5, 42, 3, 0, 0, 0)

    def splitmix(x):
        x = (x + 0x9E3779B97F4A7C15) & (2**64 - 1)
        x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & (2**64 - 1)
        x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & (2**64 - 1)
        return x ^ (x >> 31)
    want = [splitmix(((42 ^ ((3 * 0x9E3779B97F4A7C15) & (2**64 - 1))) + 5 + i) & (2**64 - 1)) for i in range(16)]
    assert [int(v) for v in out] == want


@pytest.mark.parametrize("max_rows", [0, 1, 7, 1024, 100000])
def test_view_cursor_slices_like_the_reference(ref, b200, max_rows):
    """ScanView alone is host work (view_cursor.cc:47-75): the same rows, NULLs and number of Next() calls."""
    import numpy as np
    from cases import same_results
    from supersonic_b200 import ssplan as sp
    rng = np.random.default_rng(1)
    n = 5000
    t = [sp.Column("a", sp.INT64, rng.integers(0, 100, n)), sp.Column("b", sp.DOUBLE, rng.random(n), is_null=rng.random(n) < 0.1),
         sp.Column("c", sp.BOOL, rng.integers(0, 2, n))]
    a, b = ref.run("(scan 0)", [t], next_max_rows=max_rows), b200.run("(scan 0)", [t], next_max_rows=max_rows)
    same_results(a, b)
    assert a.next_calls == b.next_calls


def test_library_sass_carries_the_blackwell_data_movement(built):
    """The fused Compute / Filter kernel stages its column tiles with TMA bulk copies completed on mbarriers
    (DESIGN.md 4.1): the sm_100a SASS of libssb200.so must contain UBLKCP and SYNCS, and -- none of the operators being a
    dense contraction -- no tensor-core (UTC*MMA / HMMA) instruction."""
    import shutil
    import subprocess
    from supersonic_b200 import capi
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("no cuobjdump on this box")
    sass = subprocess.run([cuobjdump, "-sass", capi.LIB_PATH], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    assert "sm_100a" in sass or "SM100a" in sass.upper() or "EF_CUDA_SM100" in sass, sass[:500]
    assert sass.count("UBLKCP") > 0 and sass.count("SYNCS") > 0
    assert "UTCHMMA" not in sass and "UTCQMMA" not in sass and "HMMA" not in sass
