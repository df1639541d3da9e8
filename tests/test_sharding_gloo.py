"""CPU, world_size 2 over gloo: the host-side logic of the N>1 paths -- partitioning,
the min/sum all-reduces of the slab-sharded AML, the int64 WTA key merge and the
soft-argmin partial merge.  Per-rank partials are produced by the oracle here (tests may
use it); on the GPU box the same merges run on kernel outputs (tests/test_gpu_sharding.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests._synth import synth_pair


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        import msnets_b200
        from msnets_b200 import sharding
        from oracle import ms_oracle as O
        H, W, D, border = 44, 70, 24, 10
        L, R = synth_pair(H, W, 4242, shift=5)
        costs = O.get_costs(L, R, D, 11, 3, 5, 5, border, border, border)      # [h,w,D] x4 (same on all ranks)
        d0, dn = sharding.shard_range(D, rank, world)
        sig = (128.0, 0.02, 20000.0, 20000.0)
        # --- slab AML: local min -> all-reduce(min) -> local partial den -> all-reduce(sum)
        mins = torch.from_numpy(np.stack([c[:, :, d0:d0 + dn].min(-1) for c in costs]))
        sharding.merge_min(mins)
        dens = []
        for c, s, m in zip(costs, sig, mins.numpy()):
            t = c[:, :, d0:d0 + dn] - m[:, :, None]
            dens.append(np.exp(-(t * t) / np.float32(s)).astype(np.float32).sum(-1, dtype=np.float32))
        den = torch.from_numpy(np.stack(dens))
        sharding.merge_sum(den)
        full = O.extract_features_left(*costs)                                   # [8,D,h,w]
        err = 0.0
        for k, (c, s) in enumerate(zip(costs, sig)):
            m, dd = mins.numpy()[k], den.numpy()[k]
            t = c[:, :, d0:d0 + dn] - m[:, :, None]
            aml = np.where(m[:, :, None] == O.FILL, 0, np.exp(-(t * t) / np.float32(s)) / dd[:, :, None])
            err = max(err, float(np.abs(aml.transpose(2, 0, 1) - full[4 + k, d0:d0 + dn]).max()))
        # --- WTA key merge on the NCC costs (negative values exercise the monotonic map)
        ncc = costs[1]
        lm = torch.from_numpy(np.ascontiguousarray(ncc[:, :, d0:d0 + dn].min(-1)))
        la = torch.from_numpy(np.ascontiguousarray(ncc[:, :, d0:d0 + dn].argmin(-1))) + d0
        keys = sharding.wta_key_pack(lm, la)
        sharding.merge_wta_keys(keys)
        am, m1 = sharding.wta_key_unpack(keys)
        wta_ok = bool(np.array_equal(am.numpy(), ncc.argmin(-1).astype(np.int32)) and
                      np.array_equal(m1.numpy(), ncc.min(-1)))
        # --- soft-argmin partial merge
        x = np.random.default_rng(5).standard_normal((2, D, 6, 9)).astype(np.float32) * 3
        xs = torch.from_numpy(x[:, d0:d0 + dn])
        mx = xs.max(1).values
        e = torch.exp(xs - mx.unsqueeze(1))
        dvec = torch.arange(d0, d0 + dn, dtype=torch.float32).view(1, dn, 1, 1)
        part = torch.stack([mx, e.sum(1), (e * dvec).sum(1)], dim=1)
        parts = sharding.gather_parts(part)
        disp = sharding.softargmin_merge_reference(parts).numpy()
        sa_err = float(np.abs(disp - O.soft_argmin(x)).max())
        # --- (argmin, min1, min2) triple merge: all-gather + local reduce on census channel 0,
        #     whose integer values tie across slabs
        ch0 = np.ascontiguousarray(full[0].transpose(1, 2, 0))                    # [h,w,D]
        li, l1, l2 = O.wta(ch0[:, :, d0:d0 + dn])
        gp = sharding.gather_wta_parts(torch.from_numpy(li + d0)[None], torch.from_numpy(l1)[None],
                                       torch.from_numpy(l2)[None])
        gi, g1, g2 = sharding.wta_merge_reference(*gp)
        wi, w1, w2 = O.wta(ch0)
        wta_ok = wta_ok and bool(np.array_equal(gi.numpy(), wi) and np.array_equal(g1.numpy(), w1)
                                 and np.array_equal(g2.numpy(), w2))
        ret[rank] = (err, wta_ok, sa_err, tuple(parts.shape))
    finally:
        dist.destroy_process_group()


def test_slab_merges_world2_gloo():
    world = 2
    mgr = mp.get_context("spawn").Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert len(ret) == world
    for rank in range(world):
        err, wta_ok, sa_err, shape = ret[rank]
        assert err <= 2e-6, "slab-merged AML off by %g on rank %d" % (err, rank)
        assert wta_ok, "WTA key merge disagrees with np.argmin on rank %d" % rank
        assert sa_err <= 1e-3
        assert shape == (2, 2, 3, 6, 9)


def test_partitions():
    import msnets_b200
    from msnets_b200 import sharding
    for total, world in ((192, 8), (640, 8), (10, 4), (7, 3), (5, 5)):
        spans = [sharding.shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == total
        for (b0, c0), (b1, _) in zip(spans, spans[1:]):
            assert b0 + c0 == b1
        assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    assert sharding.shard_batch(8, 1, 4) == [1, 5]
    assert sorted(sum((sharding.shard_batch(11, r, 3) for r in range(3)), [])) == list(range(11))
    with pytest.raises(ValueError):
        sharding.shard_range(8, 4, 4)


def test_wta_key_roundtrip_and_order():
    import msnets_b200
    from msnets_b200 import sharding
    v = torch.tensor([-1.0, -0.5, -0.0, 0.0, 1e-20, 0.3, 120.0, 2147483648.0], dtype=torch.float32)
    d = torch.arange(8, dtype=torch.int64)
    k = sharding.wta_key_pack(v, d)
    assert torch.all(k[1:] >= k[:-1])                 # monotonic in the cost
    dd, vv = sharding.wta_key_unpack(k)
    assert torch.equal(dd, d.to(torch.int32)) and torch.equal(vv.view(torch.int32), v.view(torch.int32))
    same = sharding.wta_key_pack(torch.tensor([5.0, 5.0]), torch.tensor([9, 3]))
    assert same.min() == same[1]                      # ties resolve to the lower disparity
