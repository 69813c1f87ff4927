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
