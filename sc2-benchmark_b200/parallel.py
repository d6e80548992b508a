"""Data-parallel evaluation plumbing: one process per GPU, images sharded by contiguous batch slices, no tensor
ever crosses GPUs; the only collective is ONE all-reduce of a small counter vector per evaluation
(the reference reduces two float64 per meter at script/task/image_classification.py:139 and never reduces the
byte counts, sc2bench/analysis.py:136-142 -- here bytes and symbols ride in the same vector).
"""
import os

import torch
import torch.distributed as dist

COUNTER_NAMES = ('images', 'correct_top1', 'correct_top5', 'bytes', 'symbols')


def init_distributed(backend=None):
    """Reads RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* from the environment (torchrun).  Returns (rank, world, local_rank)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        kwargs = {}
        if backend == 'nccl':
            torch.cuda.set_device(local_rank)
            kwargs['device_id'] = torch.device('cuda', local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kwargs)
    return rank, world, local_rank


def shard_bounds(n_items, rank, world):
    """Contiguous slice [lo, hi) of rank `rank`: sizes differ by at most one, earlier ranks take the remainder."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class EvalCounters:
    """[images, correct@1, correct@5, bytes, symbols] accumulated on the device, reduced once."""

    def __init__(self, device):
        self.values = torch.zeros(len(COUNTER_NAMES), dtype=torch.float64, device=device)

    def add(self, images=0, correct_top1=0, correct_top5=0, bytes=0, symbols=0):
        upd = [images, correct_top1, correct_top5, bytes, symbols]
        if any(isinstance(u, torch.Tensor) for u in upd):
            self.values += torch.stack([u.to(self.values) if isinstance(u, torch.Tensor) else
                                        torch.tensor(float(u), dtype=torch.float64, device=self.values.device) for u in upd])
        else:
            self.values += torch.tensor(upd, dtype=torch.float64, device=self.values.device)

    def all_reduce(self):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.values, op=dist.ReduceOp.SUM)
        return self

    def as_dict(self):
        v = self.values.tolist()
        out = dict(zip(COUNTER_NAMES, v))
        n, sym = max(out['images'], 1.0), max(out['symbols'], 1.0)
        out['top1'] = out['correct_top1'] / n
        out['top5'] = out['correct_top5'] / n
        out['bytes_per_image'] = out['bytes'] / n
        out['bits_per_symbol'] = 8.0 * out['bytes'] / sym
        return out


def topk_correct(logits, target, ks=(1, 5)):
    """Number of samples whose target is within the top-k logits (script/task/image_classification.py:91-103), on device."""
    maxk = max(ks)
    pred = logits.topk(maxk, dim=1).indices
    hit = pred.eq(target.view(-1, 1))
    return [hit[:, :k].any(dim=1).sum() for k in ks]
