"""-m gpu: the sharded path (hbsm_comm_init / hbsm_publish / hbsm_sharded_product, NCCL inside the library) end to end on
however many GPUs the box has (world_size = 1 runs the same entry points with an empty exchange), compared with the
single-GPU engine result; equal and balanced slabs."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
import hierarchical_block_sparse_lib_b200 as hb
from hierarchical_block_sparse_lib_b200 import sharded as S, generators as G
H = hb.HierarchicalBlockSparseMatrix
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dist.init_process_group("nccl", device_id=torch.device("cuda", lr)); hb.init(lr); S.comm_init()
assert S.comm_info()[:2] == (rank, world)
DT = np.float32 if os.environ.get("HBSM_TEST_DTYPE") == "float32" else np.float64
n, b, lam = (2048, 32, 0.04) if DT == np.float32 else (4096, 64, 0.02)     # fp32: the grouped tcgen05 kernels on own/halo tile lists
W = G.decay_width(lam); g = n // b
lo, hi = S.slab_bounds(g, world, rank)
if os.environ.get("HBSM_TEST_BALANCED") == "1":     # uneven slabs (what balanced_bounds produces for a clipped band)
    bounds = S.balanced_bounds(10 + np.minimum(np.minimum(np.arange(g), np.arange(g)[::-1]), 9), world)
    lo, hi = bounds[rank], bounds[rank + 1]
for (tA, tB, spamm, tau) in [(0, 0, True, 1e-6), (0, 0, False, 0.0), (0, 1, True, 1e-4), (1, 0, True, 1e-4)]:
    # full matrices (every rank, for the check) and the slabs (what a rank really holds)
    Af = H(DT, b); Af.generate_decay(n, lam, W, 1); Af.update_internal_info()
    Bf = H(DT, b); Bf.generate_decay(n, lam, W, 2); Bf.update_internal_info()
    def slab(Mf, by_col):
        bi, bj, nr, t = Mf.export_leaves()
        line = bj if by_col else bi
        m = (line >= lo) & (line < hi)
        X = H(DT, b); X.resize(n, n); X.assign_tiles(bi[m], bj[m], t[m]); X.update_internal_info(); return X
    Al = slab(Af, bool(tA)); Bl = slab(Bf, bool(tB))
    S.publish(Bl)
    Cl, nm, nb = S.sharded_product(Al, tA, Bl, tB, spamm, tau)
    wts = S.row_weights(Al, tA, Bl, tB, spamm, tau)
    Cf = H(DT)
    nmf, nbf = (H.spamm(Af, tA, Bf, tB, Cf, tau, True) if spamm else H.multiply(Af, tA, Bf, tB, Cf))
    tot = torch.tensor([nm, nb], dtype=torch.int64, device="cuda"); dist.all_reduce(tot)
    assert (int(tot[0]), int(tot[1])) == (nmf, nbf), (tot.tolist(), nmf, nbf)
    assert S.allreduce([nm, nb]) == [float(nmf), float(nbf)]          # the library's own all-reduce
    tl = Cl.export_tasks(); tf = Cf.export_tasks()
    assert np.array_equal(wts, np.bincount(tf[:, 0], minlength=g).astype(np.uint64)), "row weights != products per C block row"
    mine = tf[(tf[:, 0] >= lo) & (tf[:, 0] < hi)]
    key = lambda t: t[np.lexsort((t[:, 2], t[:, 1], t[:, 0]))]
    assert np.array_equal(key(tl), key(mine)), "per-rank executed set differs from the single-GPU set restricted to the slab"
    bi, bj, _, t = Cl.export_leaves(norms=False); fbi, fbj, _, ft = Cf.export_leaves(norms=False)
    m = (fbi >= lo) & (fbi < hi)
    assert np.array_equal(bi, fbi[m]) and np.array_equal(bj, fbj[m])
    if DT == np.float64:
        assert np.array_equal(t, ft[m])   # same kernel, same k order: bitwise
    else:   # fp32: a 2x2 group that straddles the own / halo split is chained differently in TMEM -- same sums to rounding
        assert np.linalg.norm(t.astype(np.float64) - ft[m]) <= 2e-6 * np.linalg.norm(ft[m])
if rank == 0: print("sharded ok world=%%d" %% world)
S.comm_finalize(); dist.destroy_process_group()
'''


@pytest.mark.parametrize("slabs,dtype", [("equal", "float64"), ("balanced", "float64"), ("equal", "float32")])
def test_sharded_matches_single_gpu(tmp_path, slabs, dtype):
    import torch
    ngpu = torch.cuda.device_count()
    world = 1
    for wsz in (8, 4, 2):
        if ngpu >= wsz:
            world = wsz
            break
    script = tmp_path / "sharded_check.py"
    script.write_text(SCRIPT % {"root": ROOT})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)]
    env = dict(os.environ, HBSM_TEST_BALANCED="1" if slabs == "balanced" else "0", HBSM_TEST_DTYPE=dtype)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "sharded ok" in r.stdout


SYMM_SCRIPT = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
import hierarchical_block_sparse_lib_b200 as hb
from hierarchical_block_sparse_lib_b200 import sharded as S, generators as G
H = hb.HierarchicalBlockSparseMatrix
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dist.init_process_group("nccl", device_id=torch.device("cuda", lr)); hb.init(lr); S.comm_init()
n, b, lam = 4096, 64, 0.02
W = G.decay_width(lam); g = n // b
lo, hi = S.slab_bounds(g, world, rank)
F = H(np.float64, b); F.generate_decay(n, lam, W, 3, symmetric=True); F.update_internal_info()   # full symmetric storage
U = H(np.float64); F.get_upper_triangle(U); U.update_internal_info()
bi, bj, _, t = F.export_leaves(norms=False)
m = (bi >= lo) & (bi < hi)
Fl = H(np.float64, b); Fl.resize(n, n); Fl.assign_tiles(bi[m], bj[m], t[m]); Fl.update_internal_info()
S.publish(Fl)
for tau in (1e-6, None):
    Cl, nm, nb = S.sharded_symm_square_spamm(Fl, tau)
    Cf = H(np.float64)
    if tau is None: H.symm_square(U, Cf); nmf = Cf.get_n_block_multiplications(); nbf = Cf.get_n_blocks()
    else: nmf, nbf = H.symm_square_spamm(U, Cf, tau)
    tot = torch.tensor([nm, nb], dtype=torch.int64, device="cuda"); dist.all_reduce(tot)
    assert (int(tot[0]), int(tot[1])) == (nmf, nbf), (tot.tolist(), nmf, nbf)
    ci, cj, _, tl = Cl.export_leaves(norms=False); fi, fj, _, tf = Cf.export_leaves(norms=False)
    mm = (fi >= lo) & (fi < hi)
    assert np.all(ci <= cj)
    assert np.array_equal(ci, fi[mm]) and np.array_equal(cj, fj[mm]) and np.array_equal(tl, tf[mm])   # same kernel, same k order: bitwise
if rank == 0: print("sharded symm ok world=%%d" %% world)
S.comm_finalize(); dist.destroy_process_group()
'''


def test_sharded_symm_square_matches_single_gpu(tmp_path):
    """BASELINE config 3 sharded: triu(spamm(F,F,tau)) by block rows == symm_square_spamm of the upper storage on one GPU."""
    import torch
    ngpu = torch.cuda.device_count()
    world = 1
    for wsz in (8, 4, 2):
        if ngpu >= wsz:
            world = wsz
            break
    script = tmp_path / "sharded_symm_check.py"
    script.write_text(SYMM_SCRIPT % {"root": ROOT})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29518", str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "sharded symm ok" in r.stdout
