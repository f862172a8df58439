"""Fused BatchNorm kernels: time per call and achieved bandwidth on the level shapes of the C2 scene."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from panopticsegforlargescalepointcloud_b200 import _lib
from panopticsegforlargescalepointcloud_b200._lib import ptr, check, stream_ptr
dev = torch.device("cuda:0")
lib = _lib.load()
out = []
for n, C in [(800000, 16), (800000, 64), (430000, 32), (430000, 96), (123000, 48), (123000, 128), (25000, 64), (4800, 80), (1100, 96), (230, 112)]:
    X = torch.randn(n, C, device=dev); dY = torch.randn(n, C, device=dev)
    w = torch.ones(C, device=dev); b = torch.zeros(C, device=dev); rm = torch.zeros(C, device=dev); rv = torch.ones(C, device=dev)
    Y = torch.empty_like(X); dX = torch.empty_like(X)
    sums = torch.empty(2 * C, dtype=torch.float64, device=dev); st = torch.empty(2, C, device=dev); dwb = torch.empty(2, C, device=dev)
    def fwd():
        check(lib.pgs_bn_forward(ptr(X), n, C, ptr(w), ptr(b), ptr(rm), ptr(rv), 1, 0.1, 1e-5, 1, ptr(sums), ptr(st[0]), ptr(st[1]), ptr(Y), stream_ptr()))
    def bwd():
        check(lib.pgs_bn_backward(ptr(X), ptr(Y), ptr(dY), n, C, ptr(w), ptr(st[0]), ptr(st[1]), 1, 1, ptr(sums), ptr(dX), ptr(dwb[0]), ptr(dwb[1]), stream_ptr()))
    def bwd_x():   # ReLU mask recomputed from x (what the executor calls)
        check(lib.pgs_bn_backward_ex(ptr(X), None, ptr(dY), n, C, ptr(w), ptr(b), ptr(st[0]), ptr(st[1]), 1, 1, 0, ptr(sums), ptr(dX), ptr(dwb[0]), ptr(dwb[1]), stream_ptr()))
    rec = {"n": n, "C": C}
    for name, fn, nbytes in (("fwd", fwd, 3 * 4 * n * C), ("bwd", bwd, 7 * 4 * n * C), ("bwd_x", bwd_x, 5 * 4 * n * C)):
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 50
        rec[name + "_us"] = round(us, 1); rec[name + "_gbps"] = round(nbytes / us / 1e3)
    print(json.dumps(rec), flush=True)
    out.append(rec)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bn_shapes.json"), "w"), indent=1)
