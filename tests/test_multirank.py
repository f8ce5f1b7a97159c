"""The N > 1 path on CPU: two `gloo` ranks shard a batch by stream with the product's own
partition function (no data-path collective), decode their shards (here with the CPU-side kernel
simulator standing in for the GPU), and gather one checksum per rank -- the only collective the
design has.  Rank 0 checks that the shards are disjoint, cover the batch, are balanced, and that
the gathered checksums add up to the whole-batch result."""
import os
import socket
import sys
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import dcsexplorer_b200 as dx
    import dcsfuzz
    import simutil
    # every rank builds the same global batch description (seeded) and the same partition
    rng = np.random.default_rng(1234)
    streams = []
    for i in range(36):
        nf = int(rng.integers(1, 60))
        streams.append((dcsfuzz.fuzz94(rng, nf, type1=i & 1), 0x9400, 255, 0x64, 2))
    frames = [dx.stream_frames(s[0]) for s in streams]
    part, load = dx.partition_streams(frames, world)
    mine = [i for i in range(len(streams)) if part[i] == rank]
    pcm, offs, res, _, _ = simutil.decode_streams([streams[i] for i in mine])
    local = np.array([sum(r["checksum"] for r in res) & 0x7FFFFFFFFFFFFFFF, sum(r["frames"] for r in res), len(mine)], dtype=np.int64)
    gathered = [torch.zeros(3, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(local))          # the checksum gather
    idx = [torch.zeros(len(streams), dtype=torch.int64) for _ in range(world)]
    own = torch.zeros(len(streams), dtype=torch.int64)
    own[mine] = 1
    dist.all_gather(idx, own)
    if rank == 0:
        cover = torch.stack(idx).sum(0)
        assert bool((cover == 1).all()), "shards must be disjoint and cover the batch"
        _, _, res_all, _, _ = simutil.decode_streams(streams)
        total = sum(r["checksum"] for r in res_all)
        got = 0
        for r_, i_ in zip(gathered, idx):
            sel = [k for k in range(len(streams)) if int(i_[k])]
            got += sum(res_all[k]["checksum"] for k in sel)
            assert int(r_[0]) == sum(res_all[k]["checksum"] for k in sel) & 0x7FFFFFFFFFFFFFFF
            assert int(r_[1]) == sum(res_all[k]["frames"] for k in sel)
        assert got == total
        loads = [int(x) for x in load]
        assert max(loads) - min(loads) <= max(frames) + 1, loads         # LPT bound
        open(os.path.join(out_dir, "ok"), "w").write("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_shard_by_stream_and_gather_checksums(built, tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()


def test_partition_is_deterministic_and_balanced(built):
    import dcsexplorer_b200 as dx
    rng = np.random.default_rng(5)
    frames = rng.integers(0, 65536, 5000).astype(np.uint32)
    for parts in (1, 2, 4, 8):
        p1, l1 = dx.partition_streams(frames, parts)
        p2, l2 = dx.partition_streams(frames, parts)
        assert np.array_equal(p1, p2) and p1.max() == parts - 1
        assert int(l1.sum()) == int(frames.astype(np.uint64).sum()) + frames.size
        assert int(l1.max()) - int(l1.min()) <= 65536
    p, l = dx.partition_streams(np.full(4096 * 8, 1303, dtype=np.uint32), 8)
    assert all(int(x) == 4096 * 1304 for x in l)
    p, l = dx.partition_streams(np.zeros(0, dtype=np.uint32), 4)
    assert p.size == 0
