"""The C++ drop-in class (include/hbsm/HierarchicalBlockSparseMatrix.h): compiles as plain C++11 on the CPU box,
and the reference's known-answer tests written against it (tests/cpp/test_dropin.cc) pass on the B200."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "_build", "test_dropin")


def test_dropin_header_compiles_cxx11_and_links():
    from hierarchical_block_sparse_lib_b200 import _capi
    _capi.build()
    r = subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert os.path.exists(BIN)
    # both instantiations the reference supports (gblas.h:85-143) plus every throwing stub must compile
    src = ('#include "hbsm/hierarchical_block_sparse_lib.h"\n'
           "template class hbsm::HierarchicalBlockSparseMatrix<double>;\n"
           "template class hbsm::HierarchicalBlockSparseMatrix<float>;\nint main(){return 0;}\n")
    r = subprocess.run(["/usr/bin/g++", "-std=c++11", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"),
                        "-x", "c++", "-"], input=src, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


@pytest.mark.gpu
def test_reference_known_answers_through_cpp_dropin():
    if not os.path.exists(BIN):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "cpp")], check=True)
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "checks passed" in r.stdout


REF_TESTS = [("ref_test_matrix_creation", []), ("ref_test_matrix_operations", []),
             ("ref_test_matrix_alloc_problem", ["32", "4", "20", "20", "1"])]   # arguments of the reference's `make check`


@pytest.mark.gpu
@pytest.mark.parametrize("name,argv", REF_TESTS, ids=[t[0] for t in REF_TESTS])
def test_reference_own_test_programs_pass_against_the_dropin(name, argv):
    """The reference's test_source/*.cc, compiled UNMODIFIED against include/hbsm/ (tests/cpp/Makefile `reftests`, built
    where /root/reference exists), run on the B200: failure = uncaught exception / assert, as in the reference."""
    exe = os.path.join(ROOT, "tests", "cpp", "_build", name)
    if not os.path.exists(exe):
        pytest.skip("reference test binaries were not built (no /root/reference at build time)")
    r = subprocess.run([exe] + argv, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-2000:] + r.stderr[-2000:])


@pytest.mark.gpu
def test_sharded_product_through_the_cpp_class():
    """tests/cpp/test_sharded.cc: forks one process per GPU, NCCL id handed over a pipe, hbsm::comm::init, publish() and the
    sharded_multiply / sharded_spamm / sharded_symm_square statics of the drop-in class against the single-GPU product."""
    import torch
    exe = os.path.join(ROOT, "tests", "cpp", "_build", "test_sharded")
    assert os.path.exists(exe), "tests/cpp/_build/test_sharded missing (run __graft_entry__.build())"
    ngpu = torch.cuda.device_count()
    world = 1
    for wsz in (8, 4, 2):
        if ngpu >= wsz:
            world = wsz
            break
    r = subprocess.run([exe, str(world)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "sharded c++ ok world=%d" % world in r.stdout
