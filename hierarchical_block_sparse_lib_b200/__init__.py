"""hierarchical_block_sparse_lib_b200 -- B200-native engine for the quadtree multiply / SpAMM / add hot path of
toxaart/hierarchical_block_sparse_lib.  Product path: include/hbsm_b200.h (C ABI) -> csrc/*.cu (sm_100a CUDA).
This package is the Python host-side mirror of the reference class; it never imports oracle/."""
from .matrix import (HierarchicalBlockSparseMatrix, Params, HbsmError, stage_times, set_gemm_variant, init,
                     kernel_launch_count, device_info, decay_table)
from . import generators  # noqa: F401

__all__ = ["HierarchicalBlockSparseMatrix", "Params", "HbsmError", "stage_times", "set_gemm_variant", "init",
           "kernel_launch_count", "device_info", "decay_table", "generators"]
