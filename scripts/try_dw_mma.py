"""Weight-gradient kernels: tensor-core (pgs_conv_bwd_weight_mma) vs fp32 FFMA (PGS_DW_IMPL=ffma) vs fp64 torch."""
import os, sys, json
os.environ["PGS_DW_IMPL"] = "ffma"     # pgs_conv_bwd_weight = FFMA kernel in this process; the mma entry is called directly
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from panopticsegforlargescalepointcloud_b200 import me, _lib, scenes
from panopticsegforlargescalepointcloud_b200._lib import ptr, check, stream_ptr
dev = torch.device("cuda:0")
lib = _lib.load()
N = 200000
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4      # cylinders in the batch (bench.py's step has 4)
coords = np.concatenate([np.concatenate([np.full((N, 1), b, np.int32), scenes.make_scene("urban", N, 0.12, 16.0, seed=b).coords], 1)
                         for b in range(B)], 0)
mgr = me.CoordinateManager(torch.from_numpy(coords).to(dev))
kms = {0: mgr.kernel_map(1, 1, 1, 1, 3)}
mgr.stride(1, 2); kms[1] = mgr.kernel_map(2, 2, 2, 1, 3)
mgr.stride(2, 4); kms[2] = mgr.kernel_map(4, 4, 4, 1, 3)
mgr.stride(4, 8); kms[3] = mgr.kernel_map(8, 8, 8, 1, 3)
rng = np.random.default_rng(0)
out = []
for lv, cin, cout in [(0, 16, 16), (0, 64, 16), (1, 32, 32), (1, 96, 32), (2, 48, 48), (2, 128, 48), (3, 64, 64)]:
    km = kms[lv]
    n = km.n_q
    X = torch.from_numpy(rng.standard_normal((n, cin)).astype(np.float32)).to(dev)
    dY = torch.from_numpy(rng.standard_normal((n, cout)).astype(np.float32)).to(dev)
    in_idx, out_idx, offs, max_pairs = km.pairs()
    res = {}
    for name, fn in (("ffma", lib.pgs_conv_bwd_weight), ("mma", lib.pgs_conv_bwd_weight_mma)):
        def run():
            dW = torch.zeros(27, cin, cout, device=dev)
            check(fn(ptr(X), ptr(dY), ptr(in_idx), ptr(out_idx), ptr(offs), max_pairs, 27, cin, cout, 0, ptr(dW), stream_ptr()))
            return dW
        res[name] = run()
        for _ in range(3): run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(10): run()
        e1.record(); torch.cuda.synchronize()
        res[name + "_us"] = e0.elapsed_time(e1) * 100
    ref = torch.zeros(27, cin, cout, dtype=torch.float64, device=dev)
    for k in range(27):
        idx = km.nbr[k].long(); m = idx >= 0
        ref[k] = X.double()[idx[m]].t() @ dY.double()[m]
    sc = float(ref.abs().max())
    rec = {"level": lv, "n": n, "pairs": int(offs[-1]), "c_in": cin, "c_out": cout, "ffma_us": res["ffma_us"], "mma_us": res["mma_us"],
           "err_ffma": float((res["ffma"].double() - ref).abs().max()) / sc, "err_mma": float((res["mma"].double() - ref).abs().max()) / sc}
    print(json.dumps(rec), flush=True)
    out.append(rec)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "dw_mma.json"), "w"), indent=1)
