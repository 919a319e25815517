#!/usr/bin/env python
"""Generates tests/golden/*.npz by running tests/golden_cases.py through the UNMODIFIED reference
(/root/reference compiled in place into oracle/_ref by oracle/Makefile).  Run where /root/reference exists:

    python tests/golden/make_golden.py

The fixtures travel with the repo; the machine running `-m gpu` tests has no /root/reference."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import numpy as np  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
import golden_cases  # noqa: E402
from helpers import OracleBackend  # noqa: E402

if __name__ == "__main__":
    po.build(ref=True)
    assert os.path.exists(po.REF_SO), "oracle/_ref did not build: is /root/reference present?"
    for case in golden_cases.CASES:
        out = golden_cases.run_case(OracleBackend(po.RefMatrix, case["dtype"]), case)
        path = os.path.join(HERE, case["id"] + ".npz")
        np.savez_compressed(path, **out)
        print("%-22s %3d arrays  %7.1f KiB" % (case["id"], len(out), os.path.getsize(path) / 1024))
