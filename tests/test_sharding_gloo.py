"""N > 1 host logic on CPU: two gloo ranks shard a list of families, run them independently (here with the CPU
oracle standing in for the device call) and gather in input order; the result must equal the single-process run."""
import os
import sys

import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run_local(fastas):
    import helpers as H
    out = []
    for fa in fastas:
        rc, msa, scores, cells = H.oracle_msa(fa)
        assert rc == 0
        out.append((msa, [int(s) for s in scores], cells))
    return out


def _worker(rank, world, port, fastas, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from poasta_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = sharding.sharded_msa(fastas, _run_local)
    lo, hi = sharding.shard_bounds(len(fastas), rank, world)
    q.put((rank, lo, hi, res))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process():
    from poasta_b200 import sharding, synth
    fastas = [synth.config2_family(900 + f, n_seqs=4, root_len=80, rate=0.05) for f in range(7)]
    expect = _run_local(fastas)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29517, fastas, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    blocks = sorted((lo, hi) for _, lo, hi, _ in got)
    assert blocks == [(0, 3), (3, 7)]  # disjoint, contiguous, covering
    for _, _, _, res in got:
        assert res == expect          # every rank holds the full result in input order


def test_shard_arithmetic():
    from poasta_b200 import sharding
    for n in (0, 1, 7, 10000):
        for world in (1, 2, 4, 8):
            blocks = [sharding.shard_bounds(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in blocks) - min(h - l for l, h in blocks) <= 1
    ranges = [sharding.step_family_range(k, r, 4, 100) for k in range(3) for r in range(4)]
    assert sorted(ranges) == [(i * 100, (i + 1) * 100) for i in range(12)]
