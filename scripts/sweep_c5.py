"""BASELINE config C5: dense urban tile, N in {0.25, 0.5, 1, 2} M voxels at grid 0.12 m (square tile, no crop).
Micro-sweeps of the integer kernels (coordinate hash build, stride-2 map, rulebooks k3/s1 and k3/s2, occupancy sort) and
of a single convolution (forward, input gradient, weight gradient) at C in {16, 32, 64, 96, 128, 192} square on the
stride-1 map; time per call (CUDA events, median of 5 after 2 warm-ups) and algorithmic GB/s (SURVEY 8d formulas)
against the measured HBM peak.  Single GPU ("replicas only").

    python scripts/sweep_c5.py [out.json] [N ...]
"""
import os, sys, json, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from panopticsegforlargescalepointcloud_b200 import me, _lib, scenes
from panopticsegforlargescalepointcloud_b200._lib import ptr, check, stream_ptr

def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts))


def sweep(sizes, PEAK, verbose=False):
    """-> list of rows (one per kernel x tile size); also used by `bench.py --config C5`."""
    dev = torch.device("cuda", torch.cuda.current_device())
    lib = _lib.load()

    def rec(rows, N, what, us, nbytes, **kw):
        r = dict(N=N, kernel=what, us=round(us, 1), alg_bytes=int(nbytes), gbps=round(nbytes / us / 1e3, 1),
                 frac_of_hbm_peak=round(nbytes / us / 1e3 / PEAK, 4), **kw)
        if verbose:
            print(json.dumps(r), flush=True)
        rows.append(r)

    rows = []
    for N in sizes:
        half = 8.0 * math.sqrt(math.pi * N / 200000.0)          # same areal density as the 200 k-voxel R = 16 m cylinder
        s = scenes.make_scene("urban", N, 0.12, half, seed=1, shape="square")
        n = len(s.coords)
        coords = torch.from_numpy(np.concatenate([np.zeros((n, 1), np.int32), s.coords], 1)).to(dev)
        mgr = me.CoordinateManager(coords)
        rec(rows, n, "cmap_build (hash insert + first-occurrence compaction)", timed(lambda: mgr._build(coords, 1)), 44 * n)
        m2, _ = mgr._build(coords, 2)
        rec(rows, n, "cmap_build stride 2", timed(lambda: mgr._build(coords, 2)), 16 * n + 16 * m2.n + 4 * n + 24 * n, n_out=m2.n)
        mgr.stride(1, 2)
        q, p = mgr.maps[1], mgr.maps[1]
        nbr = torch.empty((27, n), dtype=torch.int32, device=dev)
        f = lambda: check(lib.pgs_kmap_build(ptr(q.coords), q.n, ptr(p.tkeys), ptr(p.tvals), p.cap, 1, 1, 3, ptr(nbr), stream_ptr()))
        rec(rows, n, "kmap_build k3 s1 (27 probes per row)", timed(f), 16 * n + 4 * 27 * n + 12 * 27 * n)
        qc = mgr.maps[2]
        nbr2 = torch.empty((27, qc.n), dtype=torch.int32, device=dev)
        f = lambda: check(lib.pgs_kmap_build(ptr(qc.coords), qc.n, ptr(p.tkeys), ptr(p.tvals), p.cap, 1, 1, 3, ptr(nbr2), stream_ptr()))
        rec(rows, n, "kmap_build k3 s2 (coarse rows probe the fine map)", timed(f), 16 * qc.n + 4 * 27 * qc.n + 12 * 27 * qc.n, n_out=qc.n)
        km = mgr.kernel_map(1, 1, 1, 1, 3)
        def resort():
            km._sorted = None
            km.sorted()
        rec(rows, n, "occupancy sort (masks + 32-bit radix sort + permute)", timed(resort), 2 * 4 * 27 * n + 4 * 27 * n + 12 * n)
        pairs = int((km.nbr >= 0).sum())
        km.sorted(); km.pairs()
        g = torch.Generator(device="cpu").manual_seed(0)
        for C in (16, 32, 64, 96, 128, 192):
            X = torch.randn(n, C, device=dev)
            W = (torch.randn(27, C, C, device=dev) / math.sqrt(9 * C)).contiguous()
            Wt = W.transpose(1, 2).contiguous()
            by = me.conv_algorithmic_bytes(n, n, 27, C, C, True)
            kind = me._conv_kernel_choice(lib, 27, C, C, n, True)
            rec(rows, n, "conv fwd", timed(lambda: me._conv_fwd_raw(X, W, km, n, 0, 0)), by, C=C, impl=kind, pairs=pairs)
            rec(rows, n, "conv bwd-input", timed(lambda: me._conv_fwd_raw(X, Wt, km, n, 1, 1)), by, C=C, impl=kind, pairs=pairs)
            in_idx, out_idx, offs, max_pairs = km.pairs()
            dW = torch.zeros(27, C, C, device=dev)
            f = lambda: check(lib.pgs_conv_bwd_weight(ptr(X), ptr(X), ptr(in_idx), ptr(out_idx), ptr(offs), max_pairs, 27, C, C, 0,
                                                      ptr(dW), stream_ptr()))
            rec(rows, n, "conv bwd-weight", timed(f), 4 * pairs * 2 * C + 8 * pairs + 4 * 27 * C * C, C=C, pairs=pairs)
            del X, W, Wt, dW
        del mgr, km, nbr, nbr2
        torch.cuda.empty_cache()
    return rows


if __name__ == "__main__":
    try:
        PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        PEAK = 6650.0
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "c5_sweep.json")
    sizes = [int(a) for a in sys.argv[2:]] or [250000, 500000, 1000000, 2000000]
    rows = sweep(sizes, PEAK, verbose=True)
    json.dump({"hbm_peak_gbs": PEAK, "rows": rows}, open(out_path, "w"), indent=1)
