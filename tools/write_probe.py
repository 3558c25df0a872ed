#!/usr/bin/env python
"""Write-only HBM ceiling for the trajectory layout of one bench step: fills the same trajectory tensors a rollout
launch writes (torch fill kernels, no compute), CUDA events, best of N.  Calibrates the roofline fraction of the
write-only rollout kernels against the copy-measured peak in MEASURED_PEAKS.json.
  python tools/write_probe.py [game] [envs] [T]"""
import json
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rlcard_b200

game = sys.argv[1] if len(sys.argv) > 1 else 'leduc-holdem'
envs = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
T = int(sys.argv[3]) if len(sys.argv) > 3 else 128
env = rlcard_b200.VecEnv(game, envs, device='cuda:0', seed=1)
tr = env.alloc_trajectory(T)
tensors = [v for v in tr.values() if torch.is_tensor(v)]
nbytes = sum(t.numel() * t.element_size() for t in tensors)
big = torch.empty(nbytes, dtype=torch.uint8, device='cuda')
res = {}
for name, fn in (('streams', lambda: [t.fill_(1) for t in tensors]), ('one_buffer', lambda: big.fill_(1))):
    best = 1e9
    for it in range(30):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    res[name] = {'ms': best, 'gbs': nbytes / best / 1e6}
print(json.dumps({'game': game, 'envs': envs, 'T': T, 'bytes': nbytes, **res}))
