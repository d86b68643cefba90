import os, sys, json, torch
sys.path.insert(0, os.getcwd())
from stylerenderer_b200 import op
from benchmarks.kernel_bench import time_ms
k1 = torch.tensor([1., 3., 3., 1.]); k = (torch.outer(k1, k1) / 64).cuda()
x = torch.randn(128, 32, 257, 257, device='cuda')
for st in ['4', '8']:
    os.environ['SR_FIR_PLANES_STAGES'] = st
    for dbg in ['0', '1', '2', '3']:
        os.environ['SR_FIR_PLANES_DEBUG'] = dbg
        ms = time_ms(lambda: op.upfirdn2d(x, k, pad=(1, 1)))
        print('stages', st, 'debug', dbg, round(ms, 4), flush=True)
