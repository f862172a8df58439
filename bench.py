"""Headline benchmark: scenes/sec of the hot path on synthetic NPM3D-shape cylinders (BASELINE.json configs[1]).

One step = one cylinder sample per GPU through
  (1) host->device copy of the batch (e2e arm only),
  (2) PointGroup.set_input / optimize_parameters2: 7-level sparse ResUNet (82 sparse convs) forward + backward,
      semantic + offset heads, NLL + offset losses, [N>1: one NCCL all-reduce of the flat gradient bucket], Adam,
  (3) offset-shifted instance clustering: region_grow(pos + offset, nsample=200, r=1.5*grid, min 10) on
      synthetic "trained" head outputs (offset = centre - pos + noise, 2 % semantic label noise; SURVEY 8d --
      an untrained net has no meaningful votes),
  (4) device->host read of the loss and of the instance partition (e2e arm only).

`python bench.py --gpus N --steps K --warmup W` prints ONE JSON line on rank 0 (see the task contract).
`--impl reference` times the reference's CPU formulation of the same step (oracle/: per-offset gather -> GEMM ->
scatter-add sparse conv on torch CPU threads + grid ball query + sequential BFS) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_POINTS = 200000
GRID = 0.12
RADIUS = 16.0
SCENE_POOL = 4          # distinct scenes rotated through the timed steps
METRIC = "scenes/sec (cylinder samples) + PQ at 1/2/4/8 B200 vs reference CPU"


def _cfg_workload(n):
    return {"workload": "C2: NPM3D-shape synthetic cylinder, %d voxels, grid %.2f m, R %.0f m; 7-level sparse ResUNet "
                        "(82 convs) fwd+bwd + semantic/offset heads + Adam; region_grow(pos+offset, r=%.2f, nsample=200, "
                        "min 10) on synthetic head outputs" % (n, GRID, RADIUS, 1.5 * GRID),
            "scenes_per_gpu_per_step": 1, "scene_pool": SCENE_POOL, "parallelism": "dp (scene-sharded)"}


def make_inputs(seed, n=N_POINTS):
    from panopticsegforlargescalepointcloud_b200 import scenes
    s = scenes.make_scene("urban", n, GRID, RADIUS, seed=seed)
    off, _, logits = scenes.synthetic_head_outputs(s, seed=seed)
    b = scenes.collate([s])
    b.syn_shifted = (s.pos + off).astype(np.float32)
    b.syn_pred = logits.argmax(1).astype(np.int64)
    return b


HOST_KEYS = ("pos", "coords", "x", "batch", "y", "instance_labels", "instance_mask", "vote_label", "center_label",
             "num_instances", "syn_shifted", "syn_pred")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            if t < t0 or t > t1 + 0.3:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except Exception:
                continue
            for nme, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    from panopticsegforlargescalepointcloud_b200 import _lib, me, panoptic, parallel, scenes, tpk, metrics
    import torch.distributed as dist
    rank, world, local = parallel.init_from_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    n = args.n
    torch.manual_seed(2022)
    opt = panoptic.paper_options("urban", cluster_type=1, grid=GRID, use_score_net=True, prepare_epoch=30, scorer=False)
    model = panoptic.PointGroup(opt, "dummy", panoptic.DatasetProperties("urban"), None).to(dev)
    model.instantiate_optimizers({})
    model.train()
    dp = parallel.DataParallelStep(model)
    ignore = [-1] + list(scenes.stuff_classes("urban"))

    # each rank owns its own scenes (seeds disjoint across ranks): weak scaling, 1 scene / GPU / step
    host = []
    for i in range(SCENE_POOL):
        b = make_inputs(seed=rank * SCENE_POOL + i, n=n)
        host.append({k: torch.as_tensor(getattr(b, k)).pin_memory() for k in HOST_KEYS})
    resident = [{k: v.to(dev) for k, v in h.items()} for h in host]
    h2d_bytes = sum(v.numel() * v.element_size() for v in host[0].values())

    class View:
        def __init__(self, d):
            self.__dict__.update(d)

        def __getitem__(self, k):
            return self.__dict__[k]

    def step(i, e2e):
        src = host[i % SCENE_POOL] if e2e else resident[i % SCENE_POOL]
        d = {k: v.to(dev, non_blocking=True) for k, v in src.items()} if e2e else src
        dp.step(View(d), epoch=1, step=i, batch_size=1)
        clusters = tpk.region_grow(d["syn_shifted"], d["syn_pred"], d["batch"], ignore_labels=ignore, nsample=200,
                                   radius=1.5 * GRID, min_cluster_size=10)
        if e2e:
            loss = float(model.loss)                                   # D2H
            flat = torch.cat(clusters).cpu() if clusters else torch.zeros(0, dtype=torch.long)
            sizes = [c.shape[0] for c in clusters]
            return loss, flat, sizes
        return None, clusters, None

    dw_prof = []

    def timed(k, e2e, profile=None):
        me.PROFILE = profile
        me.PROFILE_DW = dw_prof if profile is not None else None
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        last = None
        for i in range(k):
            last = step(i, e2e)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t1 = time.time()
        me.PROFILE = None
        me.PROFILE_DW = None
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / k, last, (t0, t1)

    sampler = ClockSampler(local) if rank == 0 else None   # started before the warm-up: nvidia-smi needs ~1 s to come up
    for i in range(args.warmup):
        step(i, False)
    for i in range(min(args.warmup, 2)):
        step(i, True)
    l0 = _lib.launch_count()
    prof = []
    ms_dev, last_dev, (t0, t1) = timed(args.steps, False, profile=prof)
    launches = _lib.launch_count() - l0
    clocks = sampler.stop(t0, t1) if sampler else None
    ms_e2e, last_e2e, _ = timed(args.steps, True)
    d2h_bytes = 4 + (last_e2e[1].numel() * 8 + len(last_e2e[2]) * 8)

    lt = torch.tensor([launches], device=dev, dtype=torch.long)
    if world > 1:
        dist.all_reduce(lt)
    if rank != 0:
        dist.barrier()
        dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (sparse conv gather-GEMM, fwd and bwd-input launches) ----
    torch.cuda.synchronize()
    tot_b = sum(p[2] for p in prof)
    tot_f = 0
    tot_ms = sum(p[0].elapsed_time(p[1]) for p in prof)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = tot_b / (tot_ms * 1e-3) / 1e9 if tot_ms > 0 else 0.0
    traffic, traffic_note = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "conv_traffic.json")))
        traffic = tj.get("dram_bytes_per_launch")
        traffic_note = {k: tj[k] for k in ("kernel", "algorithmic_bytes_same_launch") if k in tj}
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "sparse-conv gather-GEMM launches, forward + input-gradient (pgs::conv_mma_kernel "
                "[register-operand tf32 mma, narrow layers], pgs::conv_tc_kernel [tcgen05], pgs::conv_mma_split_kernel "
                "[few-row layers]); time per kernel kind in by_kernel_ms",
                "achieved": achieved,
                "peak": peak, "peak_source": "measured" if peaks else "fallback", "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_of": traffic_note if traffic is not None else None,
                "launches": len(prof),
                "alg_bytes_per_launch": tot_b / max(len(prof), 1), "avg_launch_ms": tot_ms / max(len(prof), 1),
                "conv_share_of_step": tot_ms / (ms_dev * args.steps)}
    by_kind = {}
    for p in prof:
        by_kind[p[5]] = by_kind.get(p[5], 0.0) + p[0].elapsed_time(p[1])
    roofline["by_kernel_ms_per_step"] = {k: v / args.steps for k, v in sorted(by_kind.items())}

    # per-shape table of the conv launches (evidence for DESIGN.md section 5; not part of the JSON line)
    try:
        shapes = {}
        for p in prof:
            d = shapes.setdefault(p[4] + (p[5],), [0, 0.0, 0])
            d[0] += 1
            d[1] += p[0].elapsed_time(p[1])
            d[2] += p[2]
        rows = [{"n_in": k[0], "n_out": k[1], "K": k[2], "c_in": k[3], "c_out": k[4], "kernel": k[5], "launches": v[0],
                 "avg_us": 1e3 * v[1] / v[0], "gbps": v[2] / (v[1] * 1e-3) / 1e9} for k, v in shapes.items()]
        rows.sort(key=lambda r: -r["avg_us"] * r["launches"])
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "conv_shapes.json"), "w"), indent=1)
        shapes = {}
        for p in dw_prof:
            d = shapes.setdefault(p[2], [0, 0.0])
            d[0] += 1
            d[1] += p[0].elapsed_time(p[1])
        rows = [{"n_in": k[0], "n_out": k[1], "K": k[2], "c_in": k[3], "c_out": k[4], "launches": v[0],
                 "avg_us": 1e3 * v[1] / v[0]} for k, v in shapes.items()]
        rows.sort(key=lambda r: -r["avg_us"] * r["launches"])
        json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "dw_shapes.json"), "w"), indent=1)
    except Exception:
        pass

    # ---- PQ of the product's instance partition against the synthetic ground truth ----
    # (seed 0 scene, the one the CPU arm clusters too, so the two PQ values are comparable: SURVEY 8d "matched PQ")
    b0 = make_inputs(seed=0, n=n)
    r0 = resident[0]
    got = [c.cpu().numpy() for c in tpk.region_grow(r0["syn_shifted"], r0["syn_pred"], r0["batch"], ignore_labels=ignore,
                                                    nsample=200, radius=1.5 * GRID, min_cluster_size=10)]
    pq = metrics.panoptic_quality(b0.syn_pred, got, b0.y, b0.instance_labels, 9, list(scenes.URBAN_THINGS))

    out = {"metric": METRIC, "value": world * 1000.0 / ms_dev, "unit": "scenes/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": dict(_cfg_workload(n), l2="per-step working set (activations + gradients of 82 convs, neighbour "
                          "tables) is several GB >> 126 MB L2; inputs rotate over %d scenes" % SCENE_POOL),
           "e2e": {"value": world * 1000.0 / ms_e2e, "unit": "scenes/s", "h2d_bytes_per_step": h2d_bytes,
                   "d2h_bytes_per_step": d2h_bytes, "ms_per_step": ms_e2e},
           "gpu_launches": int(lt), "clocks": clocks, "roofline": roofline,
           "pq": {"b200": pq, "instances": len(got), "scene_seed": 0}}
    if world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_arm(n, budget_s=25.0, steps=1, warmup=0, pq_check=pq)
        if out["cpu_baseline"]["n_sample"] == n:
            out["pq"]["cpu"] = out["cpu_baseline"]["pq_sample"]
            out["pq"]["matched"] = out["cpu_baseline"]["pq_sample"] == pq
    print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# CPU arm (oracle port of the reference's CPU formulation)
# ------------------------------------------------------------------------------------------------
def _cpu_state(n_feat=4):
    """Same architecture / init as the GPU arm, as a plain state_dict (CPU modules are never constructed: the
    product package has no CPU path), created from the shapes of the reference config."""
    from panopticsegforlargescalepointcloud_b200 import panoptic
    torch.manual_seed(2022)
    opt = panoptic.paper_options("urban", cluster_type=1, grid=GRID, scorer=False)
    m = panoptic.PointGroup(opt, "dummy", panoptic.DatasetProperties("urban"), None)   # parameters only, never run
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def cpu_step(sd, cfg, b, weights, params, opt_state):
    from oracle import cpu_path, tpk_ref
    loss, sem, off, emb = cpu_path.step_loss(sd, cfg, b, weights, training=True, has_offset=True, has_embed=False)
    grads = torch.autograd.grad(loss, params, allow_unused=True)
    with torch.no_grad():                                   # Adam, lr 1e-3
        opt_state["t"] += 1
        t = opt_state["t"]
        for p, g, m, v in zip(params, grads, opt_state["m"], opt_state["v"]):
            if g is None:
                continue
            m.mul_(0.9).add_(g, alpha=0.1)
            v.mul_(0.999).addcmul_(g, g, value=0.001)
            p.addcdiv_(m / (1 - 0.9 ** t), (v / (1 - 0.999 ** t)).sqrt_().add_(1e-8), value=-1e-3)
    clusters = tpk_ref.region_grow(b.syn_shifted, b.syn_pred, b.batch.numpy(), [-1, 0, 1, 5], 200, 1.5 * GRID, 10,
                                   method="grid")
    return float(loss), clusters


def cpu_arm(n_full, budget_s, steps, warmup, pq_check=None):
    """Times `steps` CPU steps on a bounded sample: a cylinder of n_sample <= n_full voxels chosen so that the run
    fits the budget; value is scaled to full-size scenes/s by n_sample / n_full (work is linear in voxels)."""
    from oracle import cpu_path
    from panopticsegforlargescalepointcloud_b200 import scenes, panoptic, metrics
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    opt = panoptic.paper_options("urban", cluster_type=1, grid=GRID, scorer=False)
    cfg = cpu_path.resolve_cfg(opt.backbone.config, 4)
    sd = _cpu_state()
    params = [v.requires_grad_(True) for k, v in sd.items() if v.dtype.is_floating_point and "running" not in k]
    state = {"t": 0, "m": [torch.zeros_like(p) for p in params], "v": [torch.zeros_like(p) for p in params]}

    def load(nn_, seed):
        b = make_inputs(seed, nn_)
        for k in HOST_KEYS:
            if k not in ("syn_shifted", "syn_pred"):
                setattr(b, k, torch.as_tensor(getattr(b, k)))
        return b

    # calibrate on a small cylinder (same density: radius scaled with sqrt(n))
    n_cal = min(20000, n_full)
    b = load(n_cal, 100)
    t = time.time()
    cpu_step(sd, cfg, b, opt.loss_weights, params, state)
    per_pt = (time.time() - t) / n_cal
    total_steps = max(steps + warmup, 1)
    n_sample = int(min(n_full, max(n_cal, budget_s / total_steps / per_pt)))
    b = load(n_sample, 0)
    for _ in range(warmup):
        cpu_step(sd, cfg, b, opt.loss_weights, params, state)
    t = time.time()
    for _ in range(steps):
        loss, clusters = cpu_step(sd, cfg, b, opt.loss_weights, params, state)
    dt = (time.time() - t) / steps
    value = 1.0 / (dt * n_full / n_sample)
    pq = metrics.panoptic_quality(b.syn_pred, clusters, b.y.numpy(), b.instance_labels.numpy(), 9,
                                  list(scenes.URBAN_THINGS))
    return {"value": value, "unit": "scenes/s", "cores": cores, "kind": "port",
            "sample": "%d step(s) on one %d-voxel cylinder (same generator and density as the %d-voxel workload), "
                      "%.1f s/step, scaled by voxels to full-size scenes/s" % (steps, n_sample, n_full, dt),
            "ms_per_step_sample": dt * 1e3, "n_sample": n_sample, "pq_sample": pq,
            "threads": torch.get_num_threads()}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_arm(args.n, budget_s=150.0, steps=args.steps, warmup=args.warmup)
    ms = 1000.0 / cb["value"]
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "scenes/s",
                      "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                      "data": "synthetic", "config": _cfg_workload(args.n), "cpu_baseline": cb,
                      "e2e": {"value": cb["value"], "unit": "scenes/s", "h2d_bytes_per_step": 0,
                              "d2h_bytes_per_step": 0},
                      "gpu_launches": 0}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=N_POINTS)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_b200(args)


if __name__ == "__main__":
    main()
