"""CPU model of REFERENCE mode's parallel ordered squares sum (quickrank_b200/csrc/qr_exact_kernels.cuh,
ordered_squares_*): the per-addend (parity -> increment) functions, their composition per 256-addend chunk and the
final chain over chunks, restated in C (tests/models/ordered_squares_model.c) and compared with the plain sequential
loop of rtnode_histogram.cc:65-69 (fused multiply-add) and :199-203 (multiply, then add) — bit for bit.  The GPU
kernels themselves are checked the same way by tests/test_gpu_parity.py::test_ordered_squares_scheme_equals_the_sequential_chain."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def model(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("sqmodel") / "model")
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-o", exe, os.path.join(HERE, "models", "ordered_squares_model.c"), "-lm"])
    return exe


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4])   # pseudo-response-like, wide exponents, exact eighths, shrinking, exact ties
@pytest.mark.parametrize("n,seed", [(200000, 1), (5000, 7), (33333, 3)])
def test_model_equals_the_sequential_chain(model, mode, n, seed):
    out = subprocess.run([model, str(n), str(mode), str(seed)], capture_output=True, text=True, check=True).stdout
    lines = [l for l in out.splitlines() if l.startswith("n=")]
    assert len(lines) == 2 and all("EQUAL" in l for l in lines), out
