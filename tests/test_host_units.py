"""Host-only pieces of the supersonic.h mirror (Arena, ViewCopier, Block / Table with STRING cells, TableRowWriter, Limit over a
host scan, ParseString* over literals + GetConstantExpressionValue, File / FileOutput / FileInput): tests/cpp/host_units.cc is
compiled against supersonic_b200/host/include, linked with libssb200_plan.so and run here -- none of it touches the device."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_units(built, tmp_path):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++ on this box")
    lib = os.path.join(ROOT, "supersonic_b200", "lib")
    exe = str(tmp_path / "host_units")
    subprocess.run([gxx, "-O1", "-std=c++17", "-Wall", "-I" + os.path.join(ROOT, "supersonic_b200", "host", "include"),
                    os.path.join(ROOT, "tests", "cpp", "host_units.cc"), "-o", exe, "-L" + lib, "-lssb200_plan", "-lssb200",
                    "-Wl,-rpath," + lib, "-lpthread"], check=True)
    r = subprocess.run([exe, str(tmp_path / "rows.ssb")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0 and "OK host units" in r.stdout, r.stdout[-3000:]
