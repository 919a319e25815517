"""-m gpu: the reference's own known-answer tests and the committed reference-generated fixtures, replayed through
the C ABI on the B200."""
import os

import numpy as np
import pytest

import hierarchical_block_sparse_lib_b200 as hb
import known_answers
import golden_cases
from helpers import GpuBackend

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _init():
    hb.init(0)


def test_reference_known_answers_on_gpu():
    assert known_answers.run_all(GpuBackend(np.float64)) >= 30


@pytest.mark.parametrize("case", golden_cases.CASES, ids=[c["id"] for c in golden_cases.CASES])
def test_gpu_matches_reference_fixture(case):
    want = dict(np.load(os.path.join(golden_cases.GOLDEN_DIR, case["id"] + ".npz")))
    got = golden_cases.run_case(GpuBackend(case["dtype"]), case)
    golden_cases.compare(got, want, case)
