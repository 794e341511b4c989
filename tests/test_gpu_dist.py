"""Two-GPU DDP check of global normalisation (BASELINE config 4 in miniature): each rank extracts
its utterance shard, the packed statistics are all-reduced over NCCL, and every rank must end up with
the statistics / features of the single-process oracle over the whole batch.  Skipped with < 2 GPUs."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from speechain_b200.host.dist import shard_indices
        from speechain_b200.module.acoustic_frontend import AcousticFrontend
        g = torch.Generator().manual_seed(77)
        lens = sorted(torch.randint(8000, 48001, (10,), generator=g).tolist(), reverse=True)
        wav = 0.1 * torch.randn(len(lens), max(lens), generator=g)
        for i, n in enumerate(lens):
            wav[i, n:] = 0
        idx = shard_indices(len(lens), rank, world)
        conf = dict(n_mels=80, n_fft=512, win_length=0.025, hop_length=0.010, sr=16000, preemphasis=0.97)
        chain = AcousticFrontend(frontend=dict(conf=conf), normalize=dict(norm_type="global"), distributed=True).to(rank)
        chain.train()
        # pad the shard to the global T_max so the batch-level reflect padding equals the single-process run
        y, yl = chain(wav[idx].to(rank)[..., None], torch.tensor([lens[i] for i in idx], device=rank), epoch=None)
        q.put((rank, idx, y.cpu().numpy(), chain.normalize.global_mean.cpu().numpy(), chain.normalize.global_std.cpu().numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_global_norm_matches_oracle():
    import torch.multiprocessing as mp
    from oracle import frontend_np as O
    from sfe_testutil import assert_close
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(77)
    lens = sorted(torch.randint(8000, 48001, (10,), generator=g).tolist(), reverse=True)
    wav = 0.1 * torch.randn(len(lens), max(lens), generator=g)
    for i, n in enumerate(lens):
        wav[i, n:] = 0
    conf = dict(n_mels=80, n_fft=512, win_length=0.025, hop_length=0.010, sr=16000, preemphasis=0.97)
    shards = []
    for _, idx, _, _, _ in res:
        f, fl, _ = O.speech2mel(wav[idx].numpy().astype(np.float64), [lens[i] for i in idx], O.FrontendCfg(**conf))
        shards.append((f, fl))
    st = O.RunningStats()
    ref = O.feat_norm_global(shards, st, None)
    for (rank, idx, y, gm, gs), r in zip(res, ref):
        assert_close(gm, st.mean, rtol=1e-5, atol=1e-5, what=f"rank {rank} global_mean")
        assert_close(gs, st.std, rtol=1e-5, atol=1e-5, what=f"rank {rank} global_std")
        assert_close(y, r, what=f"rank {rank} features")
    assert np.array_equal(res[0][3], res[1][3]) and np.array_equal(res[0][4], res[1][4]), "ranks must agree bit for bit"


def _group_batch():
    g = torch.Generator().manual_seed(91)
    B, F, D = 9, 120, 80
    lens = sorted(torch.randint(20, F + 1, (B,), generator=g).tolist(), reverse=True)
    lens[0] = F
    feat = torch.randn(B, F, D, generator=g) * 2.0 - 5.0
    ids = torch.tensor([i % 3 for i in range(B)], dtype=torch.int64)
    return feat, torch.tensor(lens, dtype=torch.int64), ids


def _group_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from speechain_b200.host.dist import shard_indices
        from speechain_b200.module.norm.feat_norm import FeatureNormalization
        feat, lens, ids = _group_batch()
        idx = shard_indices(feat.size(0), rank, world)
        fn = FeatureNormalization(input_size=feat.size(2), norm_type="group", distributed=True).to(rank)
        fn.train()
        y, _ = fn(feat[idx].to(rank), lens[idx].to(rank), group_ids=ids[idx].to(rank), epoch=None)
        stats = {g: (fn.get_buffer(f"{g}_mean").cpu().numpy(), fn.get_buffer(f"{g}_std").cpu().numpy()) for g in range(3)}
        q.put((rank, idx, y.cpu().numpy(), stats))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_group_norm_gathers_over_nccl():
    """norm_type="group" with tensor ids: the per-utterance statistics and ids of every rank travel through the
    padded all_gather on NCCL (feat_norm.py:397-453); each rank must register the statistics of the WHOLE batch's
    groups and normalise its shard like a single process does."""
    import torch.multiprocessing as mp
    from sfe_testutil import assert_close
    from speechain_b200.module.norm.feat_norm import FeatureNormalization
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_group_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    feat, lens, ids = _group_batch()
    single = FeatureNormalization(input_size=feat.size(2), norm_type="group").to(0)
    single.train()
    want, _ = single(feat.to(0), lens.to(0), group_ids=ids.to(0), epoch=None)
    want = want.cpu().numpy()
    for rank, idx, y, stats in res:
        for g in range(3):
            assert_close(stats[g][0], single.get_buffer(f"{g}_mean").cpu().numpy(), rtol=1e-5, atol=1e-5, what=f"rank {rank} group {g} mean")
            assert_close(stats[g][1], single.get_buffer(f"{g}_std").cpu().numpy(), rtol=1e-5, atol=1e-5, what=f"rank {rank} group {g} std")
        assert_close(y, want[idx], what=f"rank {rank} features")
    for g in range(3):  # every rank saw the same gathered rows in the same order
        assert np.array_equal(res[0][3][g][0], res[1][3][g][0]) and np.array_equal(res[0][3][g][1], res[1][3][g][1])
