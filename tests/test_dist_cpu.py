"""The multi-GPU plumbing (hyperseg_b200/dist.py) on CPU with the gloo backend, world_size 2."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hyperseg_b200 import dist as hdist


def test_shard_bounds_cover_the_batch():
    for n in (1, 7, 8, 64, 65):
        for w in (1, 2, 3, 8):
            spans = [hdist.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_confusion_matrix_and_miou_single_process():
    target = torch.tensor([[0, 0, 1, 1], [2, 2, 255, 1]])
    pred = torch.tensor([[0, 1, 1, 1], [2, 0, 2, 1]])
    mat = hdist.confusion_matrix(pred, target, 3)
    assert mat.tolist() == [[1, 1, 0], [0, 3, 0], [1, 0, 1]]
    m, iou = hdist.miou(mat)
    assert torch.allclose(iou, torch.tensor([1 / 3, 3 / 4, 1 / 2], dtype=torch.float64))
    assert abs(m - (1 / 3 + 3 / 4 + 1 / 2) / 3) < 1e-12


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world_size, port, n_items, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        g = torch.Generator().manual_seed(0)
        batch = torch.randn(n_items, 3, 4, 5, generator=g)                # same "frames" on every rank
        labels = torch.randint(0, 4, (n_items, 4, 5), generator=g)
        mine = hdist.shard_batch(batch)
        lo, hi = hdist.shard_bounds(n_items, rank, world_size)
        assert torch.equal(mine, batch[lo:hi])
        logits = torch.stack([mine.sum(1) * (c + 1) for c in range(4)], 1)   # a deterministic per-image "model"
        gathered = hdist.gather_logits(logits, n_items)
        expect = torch.stack([batch.sum(1) * (c + 1) for c in range(4)], 1)
        assert torch.equal(gathered, expect)                                # == single-process result, in order
        local = hdist.confusion_matrix(logits.argmax(1), labels[lo:hi], 4)
        total = hdist.all_reduce_confusion(local.clone())
        assert torch.equal(total, hdist.confusion_matrix(expect.argmax(1), labels, 4))
        if rank == 0:
            torch.save({"miou": hdist.miou(total)[0]}, tmp)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [8, 5])
def test_gather_and_confusion_allreduce_gloo_world2(tmp_path, n_items):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), n_items, out), nprocs=2, join=True)
    assert 0.0 <= torch.load(out)["miou"] <= 1.0


def _train_worker(rank, world_size, port, tmp):
    """Data-parallel training step (reference: DataParallel replicas, hyperseg/train.py; here one process per rank)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from hyperseg_b200.synthetic import build_model, synthetic_frames
        from oracle.hyperseg_oracle import use_oracle_ops
        torch.set_num_threads(4)
        model = build_model("hyperseg-m", seed=0).train()
        ddp = torch.nn.parallel.DistributedDataParallel(model)
        frames = synthetic_frames(4, 64, 128, seed=3)
        labels = torch.randint(0, 19, (4, 64, 128), generator=torch.Generator().manual_seed(1))
        lo, hi = hdist.shard_bounds(4, rank, world_size)
        torch.manual_seed(11 + rank)                       # drop-connect masks differ per rank, as per replica
        with use_oracle_ops():
            loss = torch.nn.functional.cross_entropy(ddp(frames[lo:hi]), labels[lo:hi])
            loss.backward()
        probe = ["decoder.level_4.0.signal2weights.weight", "decoder.level_0.0.0.signal2weights.weight",
                 "weight_mapper.in_conv.0.weight", "backbone._conv_stem.weight"]
        named = dict(model.named_parameters())
        sums = torch.stack([named[n].grad.double().abs().sum() for n in probe])
        assert all(p.grad is not None for p in model.parameters())
        gathered = [torch.zeros_like(sums) for _ in range(world_size)]
        dist.all_gather(gathered, sums)
        assert all(torch.equal(g, gathered[0]) for g in gathered)      # gradients were all-reduced: identical replicas
        assert torch.isfinite(sums).all() and (sums > 0).all()
        if rank == 0:
            torch.save({"loss": loss.item()}, tmp)
    finally:
        dist.destroy_process_group()


def test_data_parallel_training_step_gloo_world2(tmp_path):
    out = str(tmp_path / "train.pt")
    mp.spawn(_train_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert 0.0 < torch.load(out)["loss"] < 10.0
