"""N>1 host logic on CPU: two gloo ranks shard a batch by the reference's stride rule, exchange the
packed statistics with ONE all-reduce and must reproduce the single-process statistics of the oracle
(feat_norm.py:237-266); group statistics travel through the padded all_gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import frontend_np as O
from speechain_b200.host.dist import all_reduce_packed, batch_stats_from_packed, shard_indices


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make_batch(seed=5, B=7, F=40, D=12):
    rng = np.random.default_rng(seed)
    lens = np.sort(rng.integers(5, F + 1, size=B))[::-1].copy()
    lens[0] = F
    x = rng.normal(size=(B, F, D)).astype(np.float64) * 2 - 5
    return x, lens


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x, lens = _make_batch()
        idx = shard_indices(len(lens), rank, world)
        D = x.shape[2]
        m, s = O.utterance_stats(x[idx], lens[idx])
        packed = torch.from_numpy(np.concatenate([m.sum(0), s.sum(0), [float(len(idx))]]))
        all_reduce_packed(packed)
        bm, bs = batch_stats_from_packed(packed, D)
        # group path: the module's padded all_gather on CPU tensors
        from speechain_b200.module.norm.feat_norm import FeatureNormalization
        fn = FeatureNormalization(input_size=D, norm_type="group", distributed=True)
        gm, gs, gids = fn._gather_group(torch.from_numpy(m).float(), torch.from_numpy(s).float(), [int(i) % 3 for i in idx])
        q.put((rank, idx, bm.numpy(), bs.numpy(), gm.numpy(), gids))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_statistics_match_single_process():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=100) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    x, lens = _make_batch()
    m, s = O.utterance_stats(x, lens)
    # the stride rule partitions the batch
    all_idx = sorted(res[0][1] + res[1][1])
    assert all_idx == list(range(len(lens)))
    assert res[0][1] == list(range(1, len(lens), 2)) and res[1][1] == list(range(0, len(lens), 2))
    for _, _, bm, bs, gm, gids in res:
        assert np.allclose(bm, m.mean(0), rtol=1e-12) and np.allclose(bs, s.mean(0), rtol=1e-12)
        # gathered per-utterance statistics arrive in rank order with their group ids
        order = res[0][1] + res[1][1]
        assert np.allclose(gm, m[order].astype(np.float32), rtol=1e-6)
        assert gids == [int(i) % 3 for i in order]


def test_shard_rule_balances_samples():
    lens = np.sort(np.random.default_rng(0).integers(32000, 320001, size=64))[::-1]
    tot = [int(lens[shard_indices(64, r, 8)].sum()) for r in range(8)]
    assert max(tot) / min(tot) < 1.25
    # rank 0 gets the lightest shard (most free memory on GPU 0), like iterator/abs.py:231-236
    assert tot[0] == min(tot)
