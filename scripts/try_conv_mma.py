"""Bring-up + per-shape timing of the register-operand mma conv (csrc/conv_mma.cu) against the tcgen05 kernel and the
fp64 numpy oracle, on the levels of a real C2 scene (200 k voxels, random row order).

    python scripts/try_conv_mma.py [json_out]        env: SORT=1 -> Morton-sorted row order (locality experiment)
"""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from panopticsegforlargescalepointcloud_b200 import me, _lib, scenes
from oracle import sparse_ref as sr

dev = torch.device("cuda:0")
N = int(os.environ.get("NPTS", "200000"))
s = scenes.make_scene("urban", N, 0.12, 16.0, seed=0)
coords = np.concatenate([np.zeros((N, 1), np.int32), s.coords], 1)
if os.environ.get("SORT") == "1":
    c = (s.coords - s.coords.min(0)).astype(np.uint64)
    key = np.zeros(N, np.uint64)
    for b in range(12):
        for a in range(3):
            key |= ((c[:, a] >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b + a)
    coords = coords[np.argsort(key, kind="stable")]
mgr = me.CoordinateManager(torch.from_numpy(coords).to(dev))
levels = {}
km0 = mgr.kernel_map(1, 1, 1, 1, 3)
levels[0] = (km0, km0.n_q, km0.n_q)
mgr.stride(1, 2)
km1 = mgr.kernel_map(2, 2, 2, 1, 3)
levels[1] = (km1, km1.n_q, km1.n_q)
mgr.stride(2, 4)
km2 = mgr.kernel_map(4, 4, 4, 1, 3)
levels[2] = (km2, km2.n_q, km2.n_q)
kdn = mgr.kernel_map(2, 1, 1, +1, 3)   # stride-2 conv level 0 -> 1
levels["0>1"] = (kdn, km0.n_q, kdn.n_q)
kup = mgr.kernel_map(1, 2, 1, -1, 3)   # transposed conv level 1 -> 0
levels["1>0"] = (kup, km1.n_q, kup.n_q)
ts = 4
for lv in (3, 4, 5, 6):
    mgr.stride(ts, 2 * ts)
    ts *= 2
    kml = mgr.kernel_map(ts, ts, ts, 1, 3)
    levels[lv] = (kml, kml.n_q, kml.n_q)
if os.environ.get("SORT") == "mask":
    # rows of each table re-ordered by their 27-bit neighbour-occupancy mask: rows of a 16-row tile then share
    # their empty offsets and the kernel skips them (the output rows are permuted the same way)
    for k, (km, n_in, n_out) in list(levels.items()):
        bits = (km.nbr >= 0).to(torch.int64)
        pc = bits.sum(1)                                    # pairs per offset
        rank = torch.argsort(torch.argsort(pc))             # rarest offset -> most significant bit
        mask = (bits << rank.view(-1, 1)).sum(0)
        order = torch.argsort(mask)
        km2 = me.KernelMap(km.nbr[:, order].contiguous(), km.K, km.n_q)
        levels[k] = (km2, n_in, n_out)
        print("level", k, "distinct masks", int(torch.unique(mask).numel()))
for k, (km, n_in, n_out) in levels.items():
    print("level", k, "n_in", n_in, "n_out", n_out, "pairs/row %.2f" % (float((km.nbr >= 0).sum()) / n_out), flush=True)

cases = [(0, 16, 16), (0, 32, 16), (0, 64, 16), (0, 16, 32), (0, 16, 64), (1, 32, 32), (1, 16, 32), (1, 32, 16),
         (2, 48, 48), (2, 32, 48), (2, 48, 32), ("0>1", 16, 16), ("1>0", 64, 64), (1, 64, 64), (0, 48, 48), (1, 48, 64)]
small = [(3, 64, 64), (3, 160, 64), (4, 80, 80), (4, 192, 80), (5, 96, 96), (5, 192, 192), (6, 112, 112), (2, 48, 48),
         (2, 128, 48), (3, 128, 128)]
cases = [c + ("mma",) for c in cases + [(3, 64, 64), (3, 48, 64), (3, 64, 48)]] + [c + ("split",) for c in small]
if os.environ.get("CASES"):
    want = {tuple(c.split(":")) for c in os.environ["CASES"].split(",")}
    cases = [c for c in cases if (str(c[0]), str(c[1]), str(c[2])) in want]
rng = np.random.default_rng(0)
out = []
for lv, cin, cout, alt in cases:
    km, n_in, n_out = levels[lv]
    X = torch.from_numpy(rng.standard_normal((n_in, cin)).astype(np.float32)).to(dev)
    W = torch.from_numpy((rng.standard_normal((27, cin, cout)) / np.sqrt(9 * cin)).astype(np.float32)).to(dev)
    Wt = W.transpose(1, 2).contiguous()
    rec = {"level": str(lv), "n_in": n_in, "n_out": n_out, "c_in": cin, "c_out": cout}
    ref = None
    if lv in (0, 2, "0>1", 3, 4, 5, 6):   # fp64 oracle on a row subset is enough: compare the first 20000 output rows
        sub = min(n_out, 20000)
        ref = sr.conv_fwd(X.cpu().numpy(), W.cpu().numpy(), km.nbr[:, :sub].cpu().numpy())
    res = {}
    rec["alt"] = alt
    for impl in ("tc", alt):
        me.CONV_IMPL = impl
        Y = me._conv_fwd_raw(X, W, km, n_out, 0, 0)
        Ym = me._conv_fwd_raw(X, Wt, km, n_out, 1, 1)   # mirrored + transposed-weights launch (input-gradient form)
        torch.cuda.synchronize()
        res[impl] = (Y, Ym)
        if ref is not None:
            rec["err_%s_vs_fp64" % ("tc" if impl == "tc" else "alt")] = float(np.abs(Y[:ref.shape[0]].cpu().numpy() - ref).max())
        for _ in range(3):
            me._conv_fwd_raw(X, W, km, n_out, 0, 0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            me._conv_fwd_raw(X, W, km, n_out, 0, 0)
        e1.record(); torch.cuda.synchronize()
        rec["us_%s" % ("tc" if impl == "tc" else "alt")] = e0.elapsed_time(e1) * 100
    rec["scale"] = float(res["tc"][0].abs().max())
    rec["max_tc_minus_alt"] = float((res["tc"][0] - res[alt][0]).abs().max())
    rec["max_tc_minus_alt_mirrorT"] = float((res["tc"][1] - res[alt][1]).abs().max())
    rec["gbps_alt"] = me.conv_algorithmic_bytes(n_in, n_out, 27, cin, cout, True) / rec["us_alt"] / 1e3
    print(json.dumps(rec), flush=True)
    out.append(rec)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
