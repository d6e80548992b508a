"""CPU, world_size 2 over gloo: the data-parallel plumbing (sharding + the single counter all-reduce)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, results):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    from sc2bench_b200 import parallel
    r, w, lr = parallel.init_distributed('gloo')
    assert (r, w) == (rank, world)
    lo, hi = parallel.shard_bounds(257, rank, world)
    torch.manual_seed(2)
    logits = torch.randn(257, 1000)
    target = torch.randint(0, 1000, (257,))
    c1, c5 = parallel.topk_correct(logits[lo:hi], target[lo:hi])
    counters = parallel.EvalCounters('cpu')
    counters.add(images=hi - lo, correct_top1=c1, correct_top5=c5, bytes=1000 * (hi - lo) + rank, symbols=72600 * (hi - lo))
    out = counters.all_reduce().as_dict()
    full1, full5 = parallel.topk_correct(logits, target)
    assert out['images'] == 257 and out['correct_top1'] == float(full1) and out['correct_top5'] == float(full5)
    assert out['bytes'] == 1000 * 257 + sum(range(world)) and out['symbols'] == 72600 * 257
    assert abs(out['bits_per_symbol'] - 8.0 * out['bytes'] / out['symbols']) < 1e-12
    results[rank] = (lo, hi)
    dist.destroy_process_group()


def test_two_rank_counter_allreduce_and_sharding():
    world = 2
    port = 29500 + os.getpid() % 2000
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
        bounds = [results[r] for r in range(world)]
    assert bounds == [(0, 129), (129, 257)]


def test_shard_bounds_cover_everything_once():
    from sc2bench_b200 import parallel
    for n in (0, 1, 7, 256, 257):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
