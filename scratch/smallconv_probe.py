import os, sys, torch, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stylerenderer_b200 import layers as L
import torch.nn.functional as F
dev = "cuda"
def tm(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / n
B = 16
for cl in (False, True):
    x = torch.randn(B, 3, 256, 256, device=dev, requires_grad=True)
    rb = L.ResBlock(3, 4, downsample=False).to(dev)
    if cl:
        x = x.detach().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    def fb():
        y = rb(x); y.sum().backward()
    print("ResBlock(3,4) 256^2 fwd+bwd channels_last=%s: %.3f ms" % (cl, tm(fb)))
    for ci, co, k in ((3, 3, 3), (3, 4, 3), (3, 4, 1), (3, 64, 1)):
        w = torch.randn(co, ci, k, k, device=dev, requires_grad=True)
        def f(): return F.conv2d(x, w, padding=k // 2)
        def fb2():
            y = F.conv2d(x, w, padding=k // 2); y.sum().backward()
        print("  conv %d->%d k%d fwd %.3f ms, fwd+bwd %.3f ms" % (ci, co, k, tm(f), tm(fb2)))
