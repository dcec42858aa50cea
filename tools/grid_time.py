import sys, time, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from carmpc_b200.batch import TerminalSetEvaluator
from carmpc_b200.grids import config2_axes
ev = TerminalSetEvaluator(np.load('/root/repo/terminal_sets/RoadMultipleCarsEnv_30_1.5_0_0.npy'))
axes = config2_axes()
n = 10**8
bits = torch.empty((n + 31)//32, dtype=torch.int32, device='cuda'); count = torch.zeros(1, dtype=torch.int64, device='cuda')
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ev.contains_grid_bits(axes, bits=bits, count=count)
    torch.cuda.synchronize(); print('grid call', (time.perf_counter() - t0) * 1e3, 'ms', int(count.item()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ev.contains_grid_bits(axes, bits=bits, count=count); e1.record(); e1.synchronize(); print('events', e0.elapsed_time(e1), 'ms')
