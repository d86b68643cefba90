import os, sys, json, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, os.getcwd())
from stylerenderer_b200 import op
from benchmarks.kernel_bench import time_ms
k1 = torch.tensor([1., 3., 3., 1.]); k = (torch.outer(k1, k1) / 64).cuda()
peak = 6540.0
for (major, h, w, pad) in [(16384, 65, 65, 1), (16384, 64, 64, 2), (8192, 129, 129, 1), (8192, 128, 128, 2), (4096, 257, 257, 1), (4096, 256, 256, 2), (1024, 513, 513, 1), (1024, 512, 512, 2)]:
    x = torch.randn(major // 32, 32, h, w, device='cuda')
    oh = h + 2 * pad - 3
    by = 4 * major * (h * w + oh * oh)
    row = {}
    for cfg in ['off', 'scalar', 'vec', 'vecS4', 'vecS12']:
        for e in ('SR_FIR_PLANES_STREAM', 'SR_FIR_PLANES_J', 'SR_FIR_PLANES_STAGES', 'SR_FIR_PLANES_VEC'): os.environ.pop(e, None)
        if cfg == 'off': os.environ['SR_FIR_PLANES_STREAM'] = '0'
        elif cfg == 'scalar': os.environ['SR_FIR_PLANES_VEC'] = '0'
        elif cfg.startswith('vecS'): os.environ['SR_FIR_PLANES_STAGES'] = cfg[4:]
        ms = time_ms(lambda: op.upfirdn2d(x, k, pad=(pad, pad)))
        row[cfg] = (round(ms, 4), round(by / ms / 1e6 / peak, 3))
    print(json.dumps({"shape": [major, h, w], "pad": pad, "ms,frac": row}), flush=True)
