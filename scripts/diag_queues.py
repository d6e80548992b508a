"""Do two CUDA streams share a hardware queue?  For every ordered pair (A, B): A gets a 5 ms sleep kernel that waits on an event
nobody has recorded yet... (simpler) A runs a long sleep kernel, B a tiny one; if B finishes while A is still running they are
independent.  python scripts/diag_queues.py [n_streams] [priority]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sc2bench_b200  # noqa: F401  (sets CUDA_DEVICE_MAX_CONNECTIONS like the product does)
import torch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
prio = int(sys.argv[2]) if len(sys.argv) > 2 else 0
dev = torch.device('cuda:0')
print('CUDA_DEVICE_MAX_CONNECTIONS =', os.environ.get('CUDA_DEVICE_MAX_CONNECTIONS'))
streams = [torch.cuda.Stream(device=dev, priority=prio) for _ in range(n)]
print('stream handles:', [hex(s.cuda_stream) for s in streams])
x = torch.zeros(1, device=dev)
torch.cuda.synchronize()
# A waits for an event that is recorded LATER on a helper stream (like a coder waiting for its batch's g_a); B must not be held up
helper = torch.cuda.Stream(device=dev)
blocked = []
for a in range(n):
    row = []
    for b in range(n):
        if a == b:
            row.append('.')
            continue
        gate = torch.cuda.Event()
        with torch.cuda.stream(helper):
            torch.cuda._sleep(4_000_000)  # ~2 ms
            gate.record(helper)
        streams[a].wait_event(gate)
        with torch.cuda.stream(streams[a]):
            x.add_(1)
        done_b = torch.cuda.Event(enable_timing=True)
        t0 = torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(streams[b]):
            t0.record()
            x.mul_(1)
            done_b.record()
        done_b.synchronize()
        open_gate = gate.query()  # True: the gate had already opened when B finished -> B was held behind A
        row.append('X' if open_gate else 'o')
        torch.cuda.synchronize()
    blocked.append(row)
print('rows: stream A waiting on an event; columns: stream B.  X = B could not run until A\'s wait was over (shared queue)')
for r in blocked:
    print(' '.join(r))
