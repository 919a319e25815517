/* TEST INFRASTRUCTURE ONLY -- never imported, linked or executed by the product path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 *
 * CPU oracle: plain-C restatement of the reference's quadtree multiply / SpAMM / add family
 * (see hbsm_oracle_impl.h for the per-function reference citations), instantiated for double (_d)
 * and float (_s).  Build: oracle/Makefile -> oracle/_build/libhbsm_oracle.so
 * (-ffp-contract=off so x*x and the running sum round separately, as in the reference's g++ -O3 build
 * without -march).
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define REAL double
#define SUF _d
#include "hbsm_oracle_impl.h"
#undef REAL
#undef SUF

#define REAL float
#define SUF _s
#include "hbsm_oracle_impl.h"
#undef REAL
#undef SUF
