"""world_size-2 gloo test of the multi-rank host logic: sharding is a partition, throughput reduction is
SUM(audio) / MAX(time), optional blob broadcast delivers identical bytes."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import covomix_b200  # noqa: F401


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import covomix_b200  # noqa: F401
    from covomix_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lengths = [1650] * 6 + [650] * 3
    mine = sharding.assign_batches(lengths, world, batch=2)[rank]
    audio = sum(len(idx) * (n - 150) / 50.0 for n, idx in mine)
    total_audio, slowest = sharding.reduce_throughput(audio, 1.0 + rank, device="cpu")
    blob = torch.arange(1000, dtype=torch.uint8) if rank == 0 else None
    got = sharding.broadcast_blob(blob, src=0)
    q.put((rank, sorted(i for _, idx in mine for i in idx), total_audio, slowest, int(got.sum())))
    dist.destroy_process_group()


def test_two_rank_sharding_and_reduction():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    idx0, idx1 = res[0][1], res[1][1]
    assert sorted(idx0 + idx1) == list(range(9)) and not set(idx0) & set(idx1)
    expect_audio = 6 * 30.0 + 3 * 10.0
    for _, _, a, w, s in res:
        assert abs(a - expect_audio) < 1e-9 and w == 2.0
        assert s == int(torch.arange(1000, dtype=torch.uint8).sum())
