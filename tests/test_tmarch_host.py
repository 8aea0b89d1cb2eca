"""CPU check of the t-marching kernel's shared-memory geometry (csrc/tmarch_geom.h): tests/host/tmarch_check.cpp replays
the producer copies and the consumer operand reads of csrc/tmarch.cu on symbolic link ids and verifies that every operand
is the link the six-staple stencil names, for several lattices, tile positions (periodic wraps) and t-segment lengths."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("tm") / "tmarch_check")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "host", "tmarch_check.cpp")])
    return exe


@pytest.mark.parametrize("args", [(8, 4, 2, 3, 3), (8, 4, 2, 4, 1), (16, 8, 4, 6, 4), (8, 8, 6, 5, 2), (24, 4, 4, 7, 7)])
def test_tmarch_operands_are_the_stencil_links(checker, args):
    r = subprocess.run([checker] + [str(a) for a in args], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("ok")
