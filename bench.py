"""bench.py v0 -- backbone forward+backward on one NPM3D-shape synthetic cylinder (configs[1]).
Replaced by the full step (heads, losses, clustering, e2e, roofline, cpu_baseline) as those land."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--n", type=int, default=200000)
    args = ap.parse_args()
    from panopticsegforlargescalepointcloud_b200 import _lib, scenes, backbone as bb
    dev = torch.device("cuda:0")
    torch.manual_seed(2022)
    net = bb.Minkowski("unet", input_nc=4, config=bb.paper_backbone_config(16)).to(dev)
    s = scenes.make_scene("urban", args.n, 0.12, 16.0, seed=0)

    class D:
        pass
    d = D()
    d.batch = torch.zeros(len(s.pos), dtype=torch.int64, device=dev)
    d.coords = torch.from_numpy(s.coords).to(dev)
    d.x = torch.from_numpy(s.x).to(dev)
    d.pos = torch.from_numpy(s.pos).to(dev)

    def step():
        net.zero_grad(set_to_none=True)
        out = net(d).x
        out.square().mean().backward()

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(json.dumps({"metric": "scenes/sec (backbone fwd+bwd only, v0)", "value": 1000.0 / ms, "unit": "scenes/s",
                      "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                      "gpu_launches": _lib.launch_count() - l0, "n": len(s.pos)}))


if __name__ == "__main__":
    main()
