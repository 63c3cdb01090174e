"""world_size-2 gloo test (CPU) of the frame sharding + code all-gather around the encode call."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from quantization_b200.dist import shard_rows, sharded_encode, sharded_round_trip_error


def _fake_encode(x):
    # stand-in for Quantizer.encode: deterministic uint8 "codes" from the frame contents
    s = (x * 7.0).round().to(torch.int64)
    return torch.stack([(s[:, 0] + s[:, 1]) % 256, (s[:, 2] * 3) % 256], dim=1).to(torch.uint8)


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(total, 4, generator=g)
    a, b = shard_rows(total, world, rank)
    codes = sharded_encode(None, x[a:b], total, encode_fn=_fake_encode)
    ok = torch.equal(codes, _fake_encode(x))
    q.put((rank, bool(ok), tuple(codes.shape)))
    dist.destroy_process_group()


def _fake_decode(c):
    return torch.stack([c[:, 0].float() / 7.0, c[:, 1].float() / 21.0, c[:, 0].float() * 0.0, c[:, 1].float() * 0.0],
                       dim=1)


def _worker_rt(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(total, 4, generator=g)
    a, b = shard_rows(total, world, rank)
    rel, codes = sharded_round_trip_error(None, x[a:b], encode_fn=_fake_encode, decode_fn=_fake_decode)
    full = ((_fake_decode(_fake_encode(x)) - x).double() ** 2).sum() / (x.double() ** 2).sum()
    q.put((rank, abs(rel - float(full)) <= 1e-12 * float(full), tuple(codes.shape)))
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, shape in res:
        assert ok and shape == (total, 2), (rank, ok, shape)


def test_shard_rows_partition():
    for total in (0, 1, 7, 8, 1000, (1 << 20) + 3):
        for world in (1, 2, 4, 8):
            spans = [shard_rows(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def test_sharded_encode_even_gloo():
    _run(64)


def test_sharded_encode_ragged_gloo():
    _run(37)


def test_sharded_round_trip_error_gloo():
    """Config-5 style round trip: every rank encodes/decodes its shard, one all-reduce of (sum err^2, sum x^2)."""
    total = 41
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_rt, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, shape in res:
        assert ok, rank
    assert sorted(shape[0] for _, _, shape in res) == [20, 21]
