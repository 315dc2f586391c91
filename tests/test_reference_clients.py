"""The reference's own client programs against this repo's supersonic.h (the boundary is the
C++ API surface; "drop-in" means the reference's clients compile unchanged and give the same answers).

test/guide/primer.cc is compiled UNMODIFIED from /root/reference (googletest replaced by the stand-in under
tests/cpp/gtest_stub) by supersonic_b200/host/Makefile into supersonic_b200/lib/guide_primer; the binary
travels to the GPU box, where it must reproduce primer.cc's golden values (Expression::Bind +
BoundExpressionTree::Evaluate, then GroupAggregate over ScanView)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRIMER_SRC = "/root/reference/test/guide/primer.cc"
PRIMER_BIN = os.path.join(ROOT, "supersonic_b200", "lib", "guide_primer")


def test_primer_compiles_unmodified_against_the_mirror():
    if not os.path.exists(PRIMER_SRC):
        pytest.skip("the reference tree is not on this machine")
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I" + os.path.join(ROOT, "tests", "cpp", "gtest_stub"),
                        "-I" + os.path.join(ROOT, "supersonic_b200", "host", "include"), PRIMER_SRC],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-3000:]


@pytest.mark.gpu
def test_primer_runs_on_the_gpu_and_reproduces_its_golden_values(built):
    assert os.path.exists(PRIMER_BIN), "supersonic_b200/lib/guide_primer was not built (needs /root/reference at build time)"
    r = subprocess.run([PRIMER_BIN], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0 and "[ 0 failed checks ]" in r.stdout, r.stdout[-3000:]
    assert "PrimerExample1.ColumnAddTest" in r.stdout and r.stdout.count("      OK ]") >= 2, r.stdout[-3000:]
