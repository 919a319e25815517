"""world_size-2 (and 4) `gloo` tests of the multi-GPU host logic on CPU: the request/select/exchange plan of
hierarchical_block_sparse_lib_b200/sharded.py delivers to every rank exactly the op(B) tiles its products touch,
and the union of the per-rank executed-product sets is bit-identical to the oracle's single-process set."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hierarchical_block_sparse_lib_b200 import generators as G
from hierarchical_block_sparse_lib_b200 import sharded as S
from oracle import pyoracle as po
from helpers import sort_tasks


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _global_tables(n, b, lam, dtype, seeds=(1, 2)):
    """(keys, norms, tiles) of A and B in ascending Morton order, norms from the oracle (bit-exact leaf sums)."""
    W = min(G.decay_width(lam), n - 1)
    out = []
    for s in seeds:
        r, c, v = G.decay_coo(n, lam, W, s, dtype=dtype)
        M = po.from_coo(po.OrcMatrix, b, n, n, r, c, v, dtype)
        bi, bj, nrm, t = M.leaves()
        keys = S.morton_encode(torch.from_numpy(bi), torch.from_numpy(bj))
        out.append((keys, torch.from_numpy(nrm.copy()), torch.from_numpy(t.copy()), M))
    return out


def _worker(rank, world, port, n, b, lam, tA, tB, spamm, tau, dtype_name, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dtype = np.dtype(dtype_name).type
        (ak, an, at, _), (bk, bn, bt, _) = _global_tables(n, b, lam, dtype)
        g = 1 << max(1, int(np.ceil(np.log2(-(-n // b)))))
        lo, hi = S.slab_bounds(g, world, rank)
        ar, ac = S.morton_decode(ak); br, bc = S.morton_decode(bk)
        a_line = ac if tA else ar            # C row of an op(A) tile
        b_line = bc if tB else br            # k of an op(B) tile
        am = (a_line >= lo) & (a_line < hi); bm = (b_line >= lo) & (b_line < hi)
        timers = {}
        keys, norms, tiles = S.exchange_b(ak[am], an[am], tA, bk[bm], bn[bm], bt[bm], tB, g, spamm, tau, None, timers)
        # the all-gather protocol csrc/sharded.cu runs (published table + gathered thresholds, requester and owner evaluate
        # the same predicate, one all-to-all of tiles) must deliver exactly the same tiles
        table = S.publish_table(bk[bm], bn[bm])
        thr = S.request_thresholds(ak[am], an[am], tA, g)
        k2, n2, t2 = S.exchange_b_allgather(thr, table, bt[bm], tB, spamm, tau)
        o1 = torch.argsort(keys); o2 = torch.argsort(k2)
        assert torch.equal(keys[o1], k2[o2]) and torch.equal(norms[o1], n2[o2]) and torch.equal(tiles[o1], t2[o2])
        # received (remote) tiles are the global ones, bit for bit, and none of them is owned by this rank
        pos = torch.searchsorted(bk, keys)
        assert torch.equal(bk[pos], keys) and torch.equal(bt[pos], tiles) and torch.equal(bn[pos], norms)
        assert keys.unique().numel() == keys.numel()
        assert not bm[pos].any()
        n_remote = keys.numel()
        # what the engine multiplies with: own tiles followed by the halo tail
        keys = torch.cat([bk[bm], keys]); norms = torch.cat([bn[bm], norms])
        # local executed set by the flat leaf-pair rule on (A_r, received B)
        rr, rc = S.morton_decode(keys)
        rk, rj = (rc, rr) if tB else (rr, rc)
        la_i = a_line[am].numpy(); la_k = (ar if tA else ac)[am].numpy(); la_n = an[am].numpy()
        tau2 = dtype(tau) * dtype(tau)
        tasks = []
        used = np.zeros(keys.numel(), bool)
        rkn, rjn, rnn = rk.numpy(), rj.numpy(), norms.numpy()
        for i, k, na in zip(la_i, la_k, la_n):
            m = rkn == k
            if spamm:
                m = m & ((na * rnn) > tau2)
            used |= m
            for j in rjn[m]:
                tasks.append((int(i), int(j), int(k)))
        assert used[len(used) - n_remote:].all(), "a tile was shipped that no product of this rank touches"
        ret[rank] = (np.array(tasks, np.int64).reshape(-1, 3), timers)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("tA,tB,spamm,tau,dtype", [(0, 0, True, 1e-4, "float64"), (0, 0, False, 0.0, "float64"),
                                                   (1, 0, True, 1e-3, "float64"), (0, 1, True, 1e-3, "float32"),
                                                   (1, 1, True, 1e-6, "float64")])
def test_exchange_plan_gloo(world, tA, tB, spamm, tau, dtype):
    n, b, lam = 256, 8, 0.15
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, n, b, lam, tA, tB, spamm, tau, dtype, ret), nprocs=world, join=True)
    assert len(ret) == world
    union = np.concatenate([ret[r][0] for r in range(world)])
    dt = np.dtype(dtype).type
    (_, _, _, Ao), (_, _, _, Bo) = _global_tables(n, b, lam, dt)
    _, nm, _, want = po.OrcMatrix.product(Ao, tA, Bo, tB, spamm=spamm, tau=tau, want_tasks=True)
    assert len(union) == nm
    assert np.array_equal(sort_tasks(union), sort_tasks(want))      # disjoint union == single-process executed set
    if spamm and tau >= 1e-4:
        assert sum(ret[r][1]["recv_tiles"] for r in range(world)) > 0   # the halo really crossed ranks


def test_slab_partition_is_top_level_quadtree_rows():
    for world in (1, 2, 4, 8):
        g = 64
        cover = []
        for r in range(world):
            lo, hi = S.slab_bounds(g, world, r)
            cover += list(range(lo, hi))
            # top-level quadtree block rows: the leading log2(world) bits of the block row are the rank
            assert all((bi >> (6 - int(np.log2(world)))) == r for bi in range(lo, hi)) if world > 1 else True
        assert cover == list(range(g))
    with pytest.raises(ValueError):
        S.slab_bounds(4, 8, 0)


def test_balanced_bounds_follow_the_weights():
    """Slab boundaries on prefix sums of per-row weights (python model == hbsm_shard_rows_balanced): monotone, cover every
    row, equal slabs for equal weights, and a clipped-band profile moves rows from the middle ranks to the edge ranks."""
    import ctypes as C
    from hierarchical_block_sparse_lib_b200 import _capi
    L = _capi.lib()
    for g, world in ((64, 4), (1024, 8), (16, 8), (8, 8)):
        assert S.balanced_bounds(np.ones(g), world) == [g * r // world for r in range(world + 1)]
        d = np.minimum(np.arange(g), np.arange(g)[::-1])
        w = (10 + np.minimum(d, 9)).astype(np.uint64)                 # band clipped at both matrix edges
        b = S.balanced_bounds(w, world)
        out = (C.c_int * (world + 1))()
        _capi.check(L.hbsm_shard_rows_balanced(w.ctypes.data_as(C.POINTER(C.c_uint64)), g, world, out))
        assert list(out) == b
        assert b[0] == 0 and b[-1] == g and all(x <= y for x, y in zip(b, b[1:]))
        if g >= 64:
            loads = [float(w[b[r]:b[r + 1]].sum()) for r in range(world)]
            eq = [float(w[g * r // world:g * (r + 1) // world].sum()) for r in range(world)]
            assert max(loads) - min(loads) <= max(eq) - min(eq)
            assert b[1] - b[0] >= g // world                                # the edge slab grew


def test_morton_roundtrip_matches_engine_convention():
    bi = torch.tensor([0, 1, 5, 1023, 123456]); bj = torch.tensor([0, 2, 3, 1, 654321])
    k = S.morton_encode(bi, bj)
    assert int(S.morton_encode(torch.tensor([0b101]), torch.tensor([0b011]))[0]) == 0b011011   # digit = 2*col + row
    r, c = S.morton_decode(k)
    assert torch.equal(r, bi) and torch.equal(c, bj)
