/* Umbrella include, same role as the reference's source/hierarchical_block_sparse_lib.h:1-12. */
#ifndef HBSM_B200_HIERARCHICAL_BLOCK_SPARSE_LIB_H
#define HBSM_B200_HIERARCHICAL_BLOCK_SPARSE_LIB_H
#include "HierarchicalBlockSparseMatrix.h"
#endif
