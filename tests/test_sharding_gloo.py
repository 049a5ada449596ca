"""CPU, world_size 2 over gloo: the multi-GPU path is replicas + independent sentence shards, so the only
cross-rank logic is the shard assignment and the max/sum reductions bench.py performs."""
import os
import socket
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    from slimt_b200 import capi, synth
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # every rank owns its own sentences (weak scaling), seeded by rank exactly as bench.py does
    sentences = synth.make_sentences(64, (2, 20), seed=1000 + rank)
    plan = capi.batcher_plan([len(s) for s in sentences], 256)
    covered = sorted(np.concatenate([ids for ids, _ in plan]).tolist())
    assert covered == list(range(64))
    tokens = float(sum(len(s) for s in sentences))
    ms = 10.0 + rank  # stand-in for the device-timed step
    t = torch.tensor([ms], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    n = torch.tensor([tokens], dtype=torch.float64)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.array([float(t[0]), float(n[0]), tokens, hash(tuple(sentences[0].tolist())) % 1000003]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    assert r0[0] == r1[0] == 11.0              # max over ranks
    assert r0[1] == r1[1] == r0[2] + r1[2]      # whole-job token count = sum over ranks
    assert r0[3] != r1[3]                       # shards differ
