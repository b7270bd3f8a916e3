"""world_size-2 gloo test (CPU) of the N>1 path's host logic: instances sharded across ranks with no
collective during simulation, then one all-gather of the per-rank runs of 32-byte records and a
4x expansion on every rank. The CPU oracle stands in for the device here; the real thing is
bench.py --gpus N."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, num_instances, steps, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import scenarios as S
    from oracle import orc
    from spear_b200 import abi, sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.shard_instances(num_instances, world, rank)
    comps = [S.comp_c2(k, burst=100 + 13 * k, churn=True) for k in range(4)]
    systems = []
    for i in mine:
        o = orc.OracleSystem("port", (1, 2 + i, 3, 4))
        e = o.add_effect(comps[i % 4], abi.make_transform((i, 0, 0)))
        systems.append((o, e))
    for f in range(steps):
        for o, e in systems:
            o.update([e], abi.make_frame_args(0.0175, f * 0.0175, f))
    runs = [o.read_stream(e)[::4] for o, e in systems]
    run = np.concatenate(runs) if runs else np.zeros(0, dtype=orc.GPU_PARTICLE_DTYPE)
    count = torch.tensor([len(run)], dtype=torch.int64)
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, count)
    counts = [int(c) for c in counts]
    cap = max(counts) + 3
    padded = np.zeros((cap, 8), dtype=np.float32)
    padded[:len(run)] = run.view(np.float32).reshape(-1, 8)
    gathered = [torch.zeros(cap, 8) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(padded))
    allrec = torch.cat(gathered).numpy()
    stream = sharding.expand4(sharding.compact_gathered(allrec, counts, cap))
    np.save(os.path.join(out_dir, "stream_%d.npy" % rank), stream)
    dist.destroy_process_group()


def test_two_rank_gather_equals_single_process(tmp_path, oracle_port):
    sys.path.insert(0, ROOT)
    import scenarios as S
    from oracle import orc
    from spear_b200 import abi, sharding
    num_instances, steps, world = 7, 12, 2
    port = 29500 + os.getpid() % 1000
    mp.spawn(_worker, args=(world, port, num_instances, steps, str(tmp_path)), nprocs=world, join=True)
    # single process reference: all instances in instance order
    comps = [S.comp_c2(k, burst=100 + 13 * k, churn=True) for k in range(4)]
    runs = []
    for i in range(num_instances):
        o = orc.OracleSystem("port", (1, 2 + i, 3, 4))
        e = o.add_effect(comps[i % 4], abi.make_transform((i, 0, 0)))
        for f in range(steps):
            o.update([e], abi.make_frame_args(0.0175, f * 0.0175, f))
        runs.append(o.read_stream(e).view(np.float32).reshape(-1, 8))
    expect = np.concatenate(runs)
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), "stream_%d.npy" % r))
        assert got.tobytes() == expect.tobytes()
    assert sharding.shard_instances(7, 2, 0) == [0, 1, 2, 3] and sharding.shard_instances(7, 2, 1) == [4, 5, 6]
    assert sum(len(sharding.shard_instances(1024, 8, r)) for r in range(8)) == 1024
