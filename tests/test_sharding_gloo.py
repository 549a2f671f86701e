"""N > 1 host logic on the CPU: world_size-2 gloo process group, shard bounds, the single final all-gather,
and global-index noise keys (the device work is stubbed: what is tested is WHICH rows / offsets each rank uses)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import ldm_b200
from ldm_b200.sharding import gather_rows, generate_sharded, shard_bounds
from oracle import philox


def test_shard_bounds_partition_exactly():
    for total in (0, 1, 7, 256, 2048, 2049):
        for ws in (1, 2, 3, 4, 8):
            spans = [shard_bounds(total, ws, r) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert [shard_bounds(2048, 8, r) for r in (0, 7)] == [(0, 256), (1792, 2048)]


class _FakeAE(torch.nn.Module):
    latent_dim = 256

    def __init__(self):
        super().__init__()
        self.p = torch.nn.Parameter(torch.zeros(1))

    def decode(self, z):            # stands in for the device decoder: a deterministic per-row function
        return z[:, :12].reshape(-1, 3, 2, 2).contiguous()


class _FakeDiffusion:
    def sample(self, shape, device, c, *, seed, sample_offset):
        # stands in for the device chain: the x_T rows of the global stream plus the class id
        z = torch.from_numpy(philox.normal_rows(seed, sample_offset, shape[0], 1000, shape[1]))
        return z + c[:, None].float()


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        classes = (torch.arange(total) * 7) % 102
        images, latents, (lo, hi) = generate_sharded(_FakeAE(), _FakeDiffusion(), classes, seed=5, device="cpu")
        x = torch.arange(lo, hi, dtype=torch.float32)[:, None].repeat(1, 3)
        q.put((rank, lo, hi, images.numpy(), gather_rows(x, total).numpy()))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(total, world=2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    return sorted(res, key=lambda r: r[0])


def _expected(total):
    classes = (torch.arange(total) * 7) % 102
    z = torch.from_numpy(philox.normal_rows(5, 0, total, 1000, 256)) + classes[:, None].float()
    return _FakeAE().decode(z).numpy()


def test_two_ranks_reproduce_the_single_rank_result_even_and_ragged():
    for total in (8, 7):
        want = _expected(total)
        res = _run(total)
        assert [(r[1], r[2]) for r in res] == [shard_bounds(total, 2, 0), shard_bounds(total, 2, 1)]
        for _, _, _, images, rows in res:
            assert np.array_equal(images, want)              # every rank holds the full gathered batch
            assert np.array_equal(rows[:, 0], np.arange(total, dtype=np.float32))


def test_single_process_path_needs_no_process_group():
    classes = torch.arange(5)
    images, latents, span = generate_sharded(_FakeAE(), _FakeDiffusion(), classes, seed=5, device="cpu")
    assert span == (0, 5) and images.shape == (5, 3, 2, 2)
