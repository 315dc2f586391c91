"""Transfer narrowing, host side (supersonic_b200/host/src/narrow.h): the vectorised range check-and-pack against the
scalar one on random ranges, alignments and the edges of the 32-bit range. Pure host code, compiled and run here."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_narrow_range_forms_agree(tmp_path):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++ on this box")
    exe = str(tmp_path / "narrow_check")
    subprocess.run([gxx, "-O2", "-std=c++17", "-Wall", os.path.join(ROOT, "tests", "cpp", "narrow_check.cc"), "-o", exe], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True, timeout=300).stdout
    assert "OK 2000 trials" in out, out
