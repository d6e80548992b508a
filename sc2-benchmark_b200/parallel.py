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


def compute_accuracy(outputs, targets, topk=(1,)):
    """Top-k accuracies in PERCENT of the batch as device tensors (script/task/image_classification.py:91-103: same arithmetic,
    float32 counts times 100 / batch_size), without leaving the device."""
    with torch.no_grad():
        batch_size = targets.size(0)
        return [c.to(torch.float32) * (100.0 / batch_size) for c in topk_correct(outputs, targets, ks=tuple(topk))]


@torch.inference_mode()
def evaluate(model, data_loader, device, log_freq=1000, title=None, header='Test:', logger=None):
    """The evaluation loop of script/task/image_classification.py:106-145 for a model built from this package, one process per
    GPU: images -> model -> top-1 / top-5.  Differences from the reference loop, none of them visible in the result:
      * no DataParallel / DistributedDataParallel wrapper (inference only: the ranks share nothing but the final counters);
      * the correct-counts accumulate ON THE DEVICE (EvalCounters) instead of `acc1.item()` per batch, so the host never waits for
        a batch; ONE all-reduce of [images, correct@1, correct@5] replaces the two float64 all-reduces per meter
        (SmoothedValue.synchronize_between_processes) -- global_avg = sum(acc_b * n_b) / sum(n_b) = 100 * correct / images.
    Returns the global top-1 accuracy in percent, like `metric_logger.acc1.global_avg`; `.top5` / `.images` ride on the result."""
    model = model.to(device)
    model.eval()
    counters = EvalCounters(device)
    if title is not None and logger is not None:
        logger.info(title)
    for it, (image, target) in enumerate(data_loader):
        if isinstance(image, torch.Tensor):
            image = image.to(device, non_blocking=True)
        if isinstance(target, torch.Tensor):
            target = target.to(device, non_blocking=True)
        output = model(image)
        c1, c5 = topk_correct(output, target, ks=(1, 5))
        counters.add(images=len(image), correct_top1=c1, correct_top5=c5)
        if logger is not None and log_freq and it % log_freq == 0:
            logger.info('%s [%d]', header, it)
    counters.all_reduce()  # gather the stats from all processes
    d = counters.as_dict()
    result = EvalResult(100.0 * d['top1'])
    result.top5, result.images = 100.0 * d['top5'], int(d['images'])
    if logger is not None:
        logger.info(' * Acc@1 {:.4f}\tAcc@5 {:.4f}\n'.format(float(result), result.top5))
    if getattr(model, 'activated_analysis', False) and hasattr(model, 'summarize'):
        model.summarize()
    return result


class EvalResult(float):
    """top-1 accuracy in percent (what the reference's evaluate returns) carrying .top5 and .images"""
