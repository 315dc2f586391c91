"""The reference's own client programs against this repo's supersonic.h (the boundary is the
C++ API surface; "drop-in" means the reference's clients compile unchanged and give the same answers).

test/guide/primer.cc, group_sort.cc and join.cc are compiled UNMODIFIED from /root/reference (googletest
replaced by the stand-in under tests/cpp/gtest_stub) by supersonic_b200/host/Makefile into
supersonic_b200/lib/guide_*; the binaries travel to the GPU box, where every EXPECT_ / ASSERT_ of the
programs must hold: primer.cc's golden values (Expression::Bind + BoundExpressionTree::Evaluate, GroupAggregate
over ScanView), group_sort.cc's GroupAggregate by (BOOL, STRING) keys and Sort of 100000 DOUBLE rows checked
against its own std::map recomputation, join.cc's HashJoin of two Tables with STRING payloads, NULL foreign keys
and DATE values parsed by ParseStringNulling."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GUIDE = "/root/reference/test/guide"
PROGRAMS = {"primer": ["PrimerExample1.ColumnAddTest", "PrimerExample2.GroupAggregateTest"],
            "group_sort": ["GroupingTest.SmallGroupingTest", "GroupingTest.LargeRandomGroupingTest", "SortingTest.SmallSortingTest",
                           "SortingTest.LargeSortingTest"],
            "join": ["HashJoinTest.SmallHashJoinTest"]}


@pytest.mark.parametrize("program", sorted(PROGRAMS))
def test_guide_program_compiles_unmodified_against_the_mirror(program):
    src = os.path.join(GUIDE, program + ".cc")
    if not os.path.exists(src):
        pytest.skip("the reference tree is not on this machine")
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I" + os.path.join(ROOT, "tests", "cpp", "gtest_stub"),
                        "-I" + os.path.join(ROOT, "supersonic_b200", "host", "include"), src],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-3000:]


@pytest.mark.gpu
@pytest.mark.parametrize("program", sorted(PROGRAMS))
def test_guide_program_runs_on_the_gpu_and_its_own_checks_hold(built, program):
    binary = os.path.join(ROOT, "supersonic_b200", "lib", "guide_" + program)
    assert os.path.exists(binary), binary + " was not built (needs /root/reference at build time)"
    r = subprocess.run([binary], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0 and "[ 0 failed checks ]" in r.stdout, r.stdout[-3000:]
    for name in PROGRAMS[program]:
        assert "[       OK ] " + name in r.stdout, r.stdout[-3000:]
