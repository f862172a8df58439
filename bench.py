"""Headline benchmark: scenes/sec of the hot path on synthetic cylinder samples (BASELINE.json configs).

`python bench.py --gpus N --steps K --warmup W` prints ONE JSON line on rank 0 (contract in the task statement).

--config C2 (default; BASELINE configs[1], the configuration the metric is quoted on): NPM3D-shape 200 k-voxel cylinders,
    one training step = one batch of 4 cylinders per GPU (the reference's `batch_size: 4`, conf/training/*.yaml:5; the
    one-cylinder-per-step number is measured in the same run and printed as `single_scene_per_step`) through
      (1) host->device copy of the batch (e2e arm only),
      (2) PointGroup.set_input / optimize_parameters2: 7-level sparse ResUNet (82 sparse convs) forward + backward,
          semantic + offset heads, NLL + offset losses, [N>1: one NCCL all-reduce of the flat gradient bucket], Adam,
      (3) offset-shifted instance clustering: region_grow(pos + offset, nsample=200, r=1.5*grid, min 10) on synthetic
          "trained" head outputs (offset = centre - pos + noise, 2 % label noise; SURVEY 8d -- an untrained net has no
          meaningful votes),
      (4) device->host read of the loss and of the instance partition (e2e arm only).
    --scenes-per-gpu B collates B cylinders into one sparse tensor per step (default 4).
--config C4 (configs[3]): 8 x C2 cylinders per step in total -- 1 GPU = one batch of 8, N GPUs = 8/N per rank (strong
    scaling), all-reduce time reported.
--config C3 (configs[2]): FOR-instance-shape cylinder, R 8 m, ~500 k voxels (grid 0.04 m), PointGroupEmbed (semantic +
    embedding heads) forward + backward + Adam, HDBSCAN(15, 5, eps 0.006) on the synthetic embeddings of the thing points.
--config C5 (configs[4]): 0.25 .. 2 M-voxel tile sweep of hash build / stride map / rulebooks / single convolutions.

`--impl reference` times the reference's CPU formulation of the same C2 step (oracle/: per-offset gather -> GEMM ->
scatter-add sparse conv on all host cores + grid ball query + sequential BFS) on the SAME batch of 200 k-voxel cylinders
(the number of measured steps is bounded by time and printed; the scene is never shrunk).
"""
import argparse
import contextlib
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
SCENE_POOL = 4          # distinct batches rotated through the timed steps
REF_BATCH_SIZE = 4      # conf/training/*.yaml:5 `batch_size: 4` -- one training step of the reference = 4 cylinders
METRIC = "scenes/sec (cylinder samples) + PQ at 1/2/4/8 B200 vs reference CPU"
PARITY = ("CUDA path == CPU oracle (bit-exact maps / rulebooks / neighbour tables / partitions incl. HDBSCAN at 350 k, "
          "fp32 1e-4) and runs under the reference's own unmodified model code (tests/test_gpu_reference_binding.py); "
          "oracle semantics of MinkowskiEngine / torch-points-kernels 0.7.0 / hdbscan 0.8.27 are restated from their "
          "published behaviour and pinned to torch conv3d / scipy cKDTree / scikit-learn goldens -- the packages "
          "themselves are absent here: PARITY UNPINNED against them")


def _cfg_workload(n, spr=REF_BATCH_SIZE, config="C2"):
    return {"workload": "%s: NPM3D-shape synthetic cylinders, %d voxels each, grid %.2f m, R %.0f m; one step = one batch of "
                        "%d cylinders per GPU (the reference trains with batch_size 4, conf/training/*.yaml:5) through the "
                        "7-level sparse ResUNet (82 convs) fwd+bwd + semantic/offset heads + Adam; region_grow(pos+offset, "
                        "r=%.2f, nsample=200, min 10) on synthetic head outputs" % (config, n, GRID, RADIUS, spr, 1.5 * GRID),
            "scenes_per_gpu_per_step": spr, "scene_pool": SCENE_POOL, "parallelism": "dp (scene-sharded)"}


def make_inputs(seeds, n=N_POINTS, kind="urban", grid=GRID, radius=RADIUS):
    """One collated batch of len(seeds) cylinders + the synthetic head outputs that drive the clustering stage."""
    from panopticsegforlargescalepointcloud_b200 import scenes
    if isinstance(seeds, int):
        seeds = [seeds]
    ss = [scenes.make_scene(kind, n, grid, radius, seed=s) for s in seeds]
    heads = [scenes.synthetic_head_outputs(s, seed=sd) for s, sd in zip(ss, seeds)]
    b = scenes.collate(ss)
    b.syn_shifted = np.concatenate([s.pos + h[0] for s, h in zip(ss, heads)]).astype(np.float32)
    b.syn_embed = np.concatenate([h[1] for h in heads]).astype(np.float32)
    b.syn_pred = np.concatenate([h[2].argmax(1) for h in heads]).astype(np.int64)
    return b


HOST_KEYS = ("pos", "coords", "x", "batch", "y", "instance_labels", "instance_mask", "vote_label", "center_label",
             "num_instances", "syn_shifted", "syn_pred")


class ClockSampler:
    """SM clock and throttle reasons during the timed region (B200_PROFILING.md's clocks line).  Read through NVML
    in-process every 100 ms (two cheap calls); an `nvidia-smi -lms 50` child polling six fields perturbed the step it was
    watching (device-resident steps 60.5 ms with it, 53.7 ms in the e2e region that runs after it is stopped).
    PGS_BENCH_SAMPLER=smi selects the nvidia-smi loop."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    MASKS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index):
        self.rows, self.proc, self.nvml, self.alive = [], None, None, True
        if os.environ.get("PGS_BENCH_SAMPLER", "nvml") != "smi":
            try:
                import pynvml
                pynvml.nvmlInit()
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
                self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
                self.nvml = pynvml
                self.t = threading.Thread(target=self._poll, daemon=True)
                self.t.start()
                return
            except Exception:
                self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nvml
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while self.alive:
            try:
                self.rows.append((time.time(), float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)), int(get_reasons(self.h))))
            except Exception:
                pass
            time.sleep(0.1)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.nvml is not None:
            self.alive = False
            sm, reasons = [], set()
            for t, mhz, mask in self.rows:
                if t0 <= t <= t1 + 0.3:
                    sm.append(mhz)
                    reasons.update(n for n, b in self.MASKS if mask & b)
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.mx, "reasons": sorted(reasons),
                    "samples": len(sm), "source": "nvml"}
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
                "samples": len(sm), "source": "nvidia-smi"}


def _pin_rank_to_cores(local, world):
    """One launch-bound Python process per GPU: give each rank its own slice of the host cores (the driver's 8-GPU
    boxes expose one 32-core NUMA node to all ranks; unpinned ranks migrate and share caches)."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = max(1, len(cores) // max(world, 1))
        mine = cores[local * per:(local + 1) * per] or cores
        os.sched_setaffinity(0, mine)
        torch.set_num_threads(max(1, min(per, 8)))
        return len(mine)
    except Exception:
        return None


def _peak():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def _timed_loop(step, k, e2e, world, dev):
    """EXACTLY k steps between barrier + synchronize on both sides; device time (CUDA events), max over ranks."""
    import torch.distributed as dist
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
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms) / k, last, (t0, t1)


# ------------------------------------------------------------------------------------------------
# B200 arm: C2 / C4
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    from panopticsegforlargescalepointcloud_b200 import _lib, me, panoptic, parallel, scenes, tpk, metrics
    import torch.distributed as dist
    rank, world, local = parallel.init_from_env()
    cores = _pin_rank_to_cores(local, world)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    n = args.n
    if args.config == "C4":
        if 8 % world:
            raise SystemExit("C4 shards 8 cylinders per step: --gpus must divide 8")
        spr = 8 // world
    else:
        spr = args.scenes_per_gpu
    torch.manual_seed(2022)
    opt = panoptic.paper_options("urban", cluster_type=1, grid=GRID, use_score_net=True, prepare_epoch=30, scorer=False)
    model = panoptic.PointGroup(opt, "dummy", panoptic.DatasetProperties("urban"), None).to(dev)
    model.instantiate_optimizers({})
    model.train()
    dp = parallel.DataParallelStep(model)
    ignore = [-1] + list(scenes.stuff_classes("urban"))

    # each rank owns its own scenes (seeds disjoint across ranks and pool slots)
    host = []
    for i in range(SCENE_POOL):
        base = (rank * SCENE_POOL + i) * spr
        b = make_inputs(list(range(base, base + spr)), n=n)
        host.append({k: torch.as_tensor(getattr(b, k)).pin_memory() for k in HOST_KEYS})
    resident = [{k: v.to(dev) for k, v in h.items()} for h in host]
    h2d_bytes = sum(v.numel() * v.element_size() for v in host[0].values())

    class View:
        def __init__(self, d):
            self.__dict__.update(d)

        def __getitem__(self, k):
            return self.__dict__[k]

    # e2e arm: the batch of step i + 1 is copied from pinned host memory on a second stream while step i computes (what a
    # data loader with pin_memory + non_blocking does); every step still pays its own host->device copy inside the timed
    # region, the copy engine just runs beside the SMs.
    copy_stream = torch.cuda.Stream(device=dev)
    staging = [{k: torch.empty_like(v, device=dev) for k, v in h.items()} for h in host]   # one device buffer set per pool slot
    inflight = {}
    rg_stream = torch.cuda.Stream(device=dev, priority=-1) if os.environ.get("PGS_BENCH_RG_STREAM", "1") == "1" else None
    PREFETCH_MAPS = os.environ.get("PGS_BENCH_PREFETCH_MAPS", "1") == "1"
    prebuilt = {}

    def upload(i):
        d = staging[i % SCENE_POOL]
        copy_stream.wait_stream(torch.cuda.current_stream(dev))     # the slot's previous consumer (SCENE_POOL steps ago) is done
        with torch.cuda.stream(copy_stream):
            for k, v in host[i % SCENE_POOL].items():
                d[k].copy_(v, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        inflight[i] = (d, ev)

    def step(i, e2e):
        if e2e:
            if i not in inflight:
                upload(i)
            d, ev = inflight.pop(i)
            torch.cuda.current_stream(dev).wait_event(ev)
            upload(i + 1)
        else:
            d = resident[i % SCENE_POOL]
        main = torch.cuda.current_stream(dev)
        if rg_stream is not None:
            rg_stream.wait_stream(main)      # the inputs (and the previous step's consumers of the cluster buffers)
        view = View(d)
        pm = prebuilt.pop((i, e2e), None)
        if pm is not None:
            view.coordinate_manager = pm
        dp.step(view, epoch=1, step=i, batch_size=spr)
        # The clustering stage reads the synthetic head outputs, not the network's: it runs on a second (high-priority)
        # stream next to the backward pass, so its convergence read-backs drain that stream only instead of the whole
        # step.  Both streams are joined before the step returns its results / before the timed region ends.
        with (torch.cuda.stream(rg_stream) if rg_stream is not None else contextlib.nullcontext()):
            clusters = tpk.region_grow(d["syn_shifted"], d["syn_pred"], d["batch"], ignore_labels=ignore, nsample=200,
                                       radius=1.5 * GRID, min_cluster_size=10)
        if rg_stream is not None:
            main.wait_stream(rg_stream)
        if PREFETCH_MAPS and rg_stream is not None:
            # the NEXT batch's coordinate maps (level-0 hash + strided maps: the forward's only host read-backs) are built
            # on the second stream while this step's backward pass runs; every step still builds one set of maps
            if e2e:
                nd, nev = inflight[i + 1]
            else:
                nd, nev = resident[(i + 1) % SCENE_POOL], None
            prebuilt.clear()
            prebuilt[(i + 1, e2e)] = dp.prefetch(View(nd), stream=rg_stream, wait_event=nev)
        if e2e:
            loss = float(model.loss)                                   # D2H
            flat = torch.cat(clusters).cpu() if clusters else torch.zeros(0, dtype=torch.long)
            sizes = [c.shape[0] for c in clusters]
            return loss, flat, sizes
        return None, clusters, None

    sampler = ClockSampler(local) if rank == 0 else None   # started before the warm-up: nvidia-smi needs ~1 s to come up
    # every batch of the pool once before the W warm-up steps: the level sizes differ from batch to batch, and a batch
    # first seen inside the timed region would make the caching allocator cudaMalloc / cudaFree multi-GB arenas there
    for i in range(SCENE_POOL):
        step(i, False)
    for i in range(args.warmup):
        step(i, False)
    for i in range(min(args.warmup, 2)):
        step(i, True)
    # ---- timed region 1: inputs resident in HBM ----
    dp.allreduce_events = [] if world > 1 else None
    l0 = _lib.launch_count()
    ms_dev, _, (t0, t1) = _timed_loop(step, args.steps, False, world, dev)
    launches = _lib.launch_count() - l0
    clocks = sampler.stop(t0, t1) if sampler else None
    ar_us = None
    if dp.allreduce_events:
        ar_us = float(np.mean([a.elapsed_time(b) for a, b in dp.allreduce_events])) * 1e3
    dp.allreduce_events = None
    # ---- timed region 2: end to end (pinned host buffers in, loss + partition out) ----
    inflight.clear()
    ms_e2e, last_e2e, _ = _timed_loop(step, args.steps, True, world, dev)
    inflight.clear()
    d2h_bytes = 4 + (last_e2e[1].numel() * 8 + len(last_e2e[2]) * 8)

    # ---- secondary: latency mode, ONE cylinder per GPU and step (round 1's headline definition) ----
    single = None
    if spr > 1 and args.config == "C2":
        one = []
        for h in host[:SCENE_POOL]:
            nb = int((h["batch"] == 0).sum())
            d = {}
            for k, v in h.items():
                if v.dim() > 0 and v.shape[0] == h["batch"].shape[0]:
                    d[k] = v[:nb].to(dev)
                elif k == "center_label":
                    d[k] = v[:v.shape[0] // spr].to(dev)
                elif k == "num_instances":
                    d[k] = v[:1].to(dev)
                else:
                    d[k] = v.to(dev)
            one.append(d)

        def step1(i, e2e):
            d = one[i % len(one)]
            dp.step(View(d), epoch=1, step=i, batch_size=1)
            return None, tpk.region_grow(d["syn_shifted"], d["syn_pred"], d["batch"], ignore_labels=ignore, nsample=200,
                                         radius=1.5 * GRID, min_cluster_size=10), None
        for i in range(max(3, len(one))):
            step1(i, False)
        ms_one, _, _ = _timed_loop(step1, args.steps, False, world, dev)
        single = {"scenes_per_gpu_per_step": 1, "ms_per_step": ms_one, "value": world * 1000.0 / ms_one, "unit": "scenes/s",
                  "note": "same step with ONE cylinder per GPU (latency mode; the coarse U-Net levels are launch-latency "
                          "bound at this size)"}

    lt = torch.tensor([launches], device=dev, dtype=torch.long)
    art = torch.tensor([ar_us or 0.0], device=dev)
    if world > 1:
        dist.all_reduce(lt)
        dist.all_reduce(art, op=dist.ReduceOp.MAX)
    if rank != 0:
        dist.barrier()
        dist.destroy_process_group()
        return

    # ---- untimed pass: per-launch CUDA events around every conv forward / input-gradient launch ----
    # (per-op Python executor, same kernels and order as the native executor; events would perturb the timed region)
    prof, dw_prof = [], []
    dp.local_only = True      # the other ranks have left: no collective in these extra steps
    me.PROFILE, me.PROFILE_DW, me.PROFILE_COUNT_PAIRS = prof, dw_prof, True
    prof_steps = 3
    for i in range(prof_steps):
        step(i, False)
    torch.cuda.synchronize()
    me.PROFILE = me.PROFILE_DW = None
    me.PROFILE_COUNT_PAIRS = False
    peak, peak_src = _peak()
    tot_ms = sum(p[0].elapsed_time(p[1]) for p in prof)
    bytes_8d = bytes_table = 0
    for p in prof:
        n_in, n_q, K, c_in, c_out = p[4]
        pairs = p[3] // (2 * c_in * c_out)
        bytes_8d += 4 * (n_in * c_in + n_q * c_out) + 4 * K * c_in * c_out + 8 * pairs          # SURVEY 8d "compulsory"
        bytes_table += p[2]                                                                     # with our 4*K*N_out table
    achieved = bytes_8d / (tot_ms * 1e-3) / 1e9 if tot_ms > 0 else 0.0
    traffic, traffic_note = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "conv_traffic.json")))
        traffic = tj.get("dram_bytes_per_launch")
        traffic_note = dict({k: tj[k] for k in ("kernel", "algorithmic_bytes_same_launch", "capture") if k in tj},
                            source="profiles/conv_traffic.json: ncu --set full capture of one launch of the dominant "
                                   "shape (committed; ncu cannot run inside the timed bench)")
    except Exception:
        pass
    by_kind, by_kind_bytes = {}, {}
    for p in prof:
        by_kind[p[5]] = by_kind.get(p[5], 0.0) + p[0].elapsed_time(p[1])
    roofline = {"bound": "hbm", "kernel": "sparse-conv gather-GEMM launches, forward + input-gradient (pgs::conv_mma(q)_kernel "
                "[register-operand tf32 mma], pgs::conv_tc_kernel [tcgen05], pgs::conv_mma_split_kernel [few-row "
                "layers], pgs::conv_fwd_kernel [4->16 input conv])",
                "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                "numerator": "SURVEY 8d compulsory bytes: 4(N_in C_in + N_out C_out) + 4 K C_in C_out + 8 P",
                "achieved_with_gather_table_bytes": bytes_table / (tot_ms * 1e-3) / 1e9 if tot_ms > 0 else 0.0,
                "traffic": traffic, "traffic_of": traffic_note,
                "launches_per_step": len(prof) / prof_steps,
                "alg_bytes_per_launch": bytes_8d / max(len(prof), 1), "avg_launch_ms": tot_ms / max(len(prof), 1),
                "conv_ms_per_step": tot_ms / prof_steps, "conv_share_of_step": (tot_ms / prof_steps) / ms_dev,
                "measured_in": "separate untimed pass of %d steps, one CUDA-event pair per launch" % prof_steps,
                "by_kernel_ms_per_step": {k: v / prof_steps for k, v in sorted(by_kind.items())}}
    _dump_shapes(prof, dw_prof, prof_steps)

    # ---- PQ of the product's instance partition against the synthetic ground truth (seed-0 scene) ----
    b0 = make_inputs(0, n=n)
    r0 = {k: torch.as_tensor(getattr(b0, k)).to(dev) for k in ("syn_shifted", "syn_pred", "batch")}
    got = [c.cpu().numpy() for c in tpk.region_grow(r0["syn_shifted"], r0["syn_pred"], r0["batch"], ignore_labels=ignore,
                                                    nsample=200, radius=1.5 * GRID, min_cluster_size=10)]
    pq = metrics.panoptic_quality(b0.syn_pred, got, b0.y, b0.instance_labels, 9, list(scenes.URBAN_THINGS))

    scenes_per_step = spr * world
    out = {"metric": METRIC, "value": scenes_per_step * 1000.0 / ms_dev, "unit": "scenes/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True,
           "scaling": "strong" if args.config == "C4" else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": dict(_cfg_workload(n, spr, args.config), l2="per-step working set (activations + gradients of 82 "
                          "convs, neighbour tables) is several GB >> 126 MB L2; inputs rotate over %d batches (each run once, untimed, "
                          "before the W warm-up steps so that the allocator has seen every shape)" % SCENE_POOL,
                          host_cores_per_rank=cores, executor=os.environ.get("PGS_EXECUTOR", "native"),
                          clustering="region_grow on a second CUDA stream next to the backward pass, joined every step"
                          if rg_stream is not None else "region_grow on the step's stream",
                          coordinate_maps="next batch's maps prefetched on the second stream (DataParallelStep.prefetch)"
                          if (PREFETCH_MAPS and rg_stream is not None) else "built inside the forward pass"),
           "e2e": {"value": scenes_per_step * 1000.0 / ms_e2e, "unit": "scenes/s", "h2d_bytes_per_step": h2d_bytes,
                   "d2h_bytes_per_step": d2h_bytes, "ms_per_step": ms_e2e},
           "gpu_launches": int(lt), "gpu_launches_per_step_per_gpu": int(lt) / world / args.steps, "clocks": clocks,
           "roofline": roofline, "parity": PARITY,
           "pq": {"b200": pq, "instances": len(got), "scene_seed": 0}}
    if single is not None:
        out["single_scene_per_step"] = single
    if world > 1:
        out["allreduce_us"] = float(art)
        out["allreduce"] = "one ncclAllReduce(sum) over the flat fp32 gradient bucket (%d floats) per step, timed with " \
                           "CUDA events on its stream, max over ranks" % dp.bucket.flat.numel()
    if world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_arm(n, steps=1, warmup=0, max_seconds=30.0, spr=spr)
        out["pq"]["cpu"] = out["cpu_baseline"]["pq_sample"]
        out["pq"]["matched"] = out["cpu_baseline"]["pq_sample"] == pq
    _emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _dump_shapes(prof, dw_prof, steps):
    """Per-shape table of the conv launches (evidence for DESIGN.md section 5; not part of the JSON line)."""
    try:
        shapes = {}
        for p in prof:
            d = shapes.setdefault(p[4] + (p[5],), [0, 0.0, 0, 0])
            d[0] += 1
            d[1] += p[0].elapsed_time(p[1])
            n_in, n_q, K, c_in, c_out = p[4]
            pairs = p[3] // (2 * c_in * c_out)
            d[2] += 4 * (n_in * c_in + n_q * c_out) + 4 * K * c_in * c_out + 8 * pairs
            d[3] = pairs
        rows = [{"n_in": k[0], "n_out": k[1], "K": k[2], "c_in": k[3], "c_out": k[4], "kernel": k[5], "launches_per_step":
                 v[0] / steps, "pairs": v[3], "avg_us": 1e3 * v[1] / v[0], "gbps_8d": v[2] / (v[1] * 1e-3) / 1e9}
                for k, v in shapes.items()]
        rows.sort(key=lambda r: -r["avg_us"] * r["launches_per_step"])
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "conv_shapes.json"), "w"), indent=1)
        shapes = {}
        for p in dw_prof:
            d = shapes.setdefault(p[2], [0, 0.0])
            d[0] += 1
            d[1] += p[0].elapsed_time(p[1])
        rows = [{"n_in": k[0], "n_out": k[1], "K": k[2], "c_in": k[3], "c_out": k[4], "launches_per_step": v[0] / steps,
                 "avg_us": 1e3 * v[1] / v[0]} for k, v in shapes.items()]
        rows.sort(key=lambda r: -r["avg_us"] * r["launches_per_step"])
        json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "dw_shapes.json"), "w"), indent=1)
    except Exception:
        pass


# ------------------------------------------------------------------------------------------------
# B200 arm: C3 (FOR-instance cylinder, embedding head + HDBSCAN)
# ------------------------------------------------------------------------------------------------
def run_c3(args):
    from panopticsegforlargescalepointcloud_b200 import _lib, panoptic, parallel, scenes, hdbscan
    rank, world, local = parallel.init_from_env()
    _pin_rank_to_cores(local, world)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    n, grid, radius = args.n if args.n != N_POINTS else 500000, 0.04, 8.0
    torch.manual_seed(2022)
    opt = panoptic.paper_options("forest", cluster_type=14, grid=grid, use_score_net=True, prepare_epoch=30, scorer=False)
    model = panoptic.PointGroupEmbed(opt, "dummy", panoptic.DatasetProperties("forest"), None).to(dev)
    model.instantiate_optimizers({})
    model.train()
    dp = parallel.DataParallelStep(model)
    pool = 2
    keys = [k for k in HOST_KEYS if k != "syn_shifted"] + ["syn_embed"]
    host = []
    for i in range(pool):
        b = make_inputs(rank * pool + i, n=n, kind="forest", grid=grid, radius=radius)
        host.append({k: torch.as_tensor(getattr(b, k)).pin_memory() for k in keys})
    resident = [{k: v.to(dev) for k, v in h.items()} for h in host]
    h2d_bytes = sum(v.numel() * v.element_size() for v in host[0].values())
    stuff = torch.tensor([-1] + list(scenes.stuff_classes("forest")), device=dev)
    hdb_ms, tree_ms, n_thing, rounds = [], [], [], []

    class View:
        def __init__(self, d):
            self.__dict__.update(d)

        def __getitem__(self, k):
            return self.__dict__[k]

    # As in the C2 step, the clustering stage reads the synthetic head outputs, not the network's: it is issued on a
    # second stream after the training step has been enqueued, so the device MST runs next to the forward / backward
    # pass and the HOST tree stage (~40 ms of CPU) overlaps the GPU work instead of idling the device.
    hdb_stream = torch.cuda.Stream(device=dev, priority=-1) if os.environ.get("PGS_BENCH_RG_STREAM", "1") == "1" else None

    def cluster(d, record):
        thing = ~torch.isin(d["syn_pred"], stuff)
        X = d["syn_embed"][thing].contiguous()
        m = hdbscan.HDBSCAN(min_cluster_size=15, min_samples=5, cluster_selection_epsilon=0.006)
        t = time.time()
        lab = m.fit_predict(X)           # device kNN + Boruvka MST, host tree stage, labels back on the device
        if record:
            hdb_ms.append((time.time() - t) * 1e3)
            n_thing.append(int(X.shape[0]))
            rounds.append(m.boruvka_rounds_)
        return lab

    def step(i, e2e):
        src = host[i % pool] if e2e else resident[i % pool]
        d = {k: v.to(dev, non_blocking=True) for k, v in src.items()} if e2e else src
        main = torch.cuda.current_stream(dev)
        if hdb_stream is not None:
            hdb_stream.wait_stream(main)         # this step's inputs
        dp.step(View(d), epoch=1, step=i, batch_size=1)
        with (torch.cuda.stream(hdb_stream) if hdb_stream is not None else contextlib.nullcontext()):
            lab = cluster(d, hdb_stream is None)
        if hdb_stream is not None:
            main.wait_stream(hdb_stream)
        if e2e:
            return float(model.loss), lab.cpu()
        return None, lab

    sampler = ClockSampler(local) if rank == 0 else None
    for i in range(args.warmup):
        step(i, False)
    step(0, True)
    del hdb_ms[:], n_thing[:], rounds[:]
    l0 = _lib.launch_count()
    ms_dev, last, (t0, t1) = _timed_loop(step, args.steps, False, world, dev)
    launches = _lib.launch_count() - l0
    clocks = sampler.stop(t0, t1) if sampler else None
    ms_e2e, last_e2e, _ = _timed_loop(step, args.steps, True, world, dev)
    if hdb_stream is not None:
        # HDBSCAN alone (nothing else on the device), untimed: the figure behind `hdbscan.ms_per_scene` and the roofline
        torch.cuda.synchronize()
        for i in range(3):
            cluster(resident[i % pool], True)
    hd = float(np.mean(hdb_ms))
    nt, rd = int(np.mean(n_thing)), float(np.mean(rounds))
    if rank != 0:
        return
    peak, peak_src = _peak()
    D = 5
    # kNN core distances 4nD + 4n; Boruvka per round n(4D + 8) read + 12n candidates; edges out 12(n-1)
    alg = (4 * nt * D + 4 * nt) + rd * (nt * (4 * D + 8) + 12 * nt) + 12 * (nt - 1)
    labels = last[1].cpu().numpy()
    out = {"metric": METRIC, "value": world * 1000.0 / ms_dev, "unit": "scenes/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32 (conv) / f64 (HDBSCAN distances)", "data": "synthetic",
           "config": {"workload": "C3: FOR-instance-shape synthetic cylinder, %d voxels, grid %.2f m (the shipped 0.2 m grid "
                      "gives ~50 k voxels; 0.04 m reaches the named size), R %.0f m; 7-level sparse ResUNet fwd+bwd + "
                      "semantic/embedding heads + Adam; HDBSCAN(15, 5, eps=0.006) on the synthetic 5-D embeddings of the "
                      "%d thing points" % (n, grid, radius, nt), "scenes_per_gpu_per_step": 1, "scene_pool": pool,
                      "parallelism": "dp (scene-sharded)",
                      "clustering": "HDBSCAN on a second CUDA stream next to the training step (its host tree stage overlaps "
                                    "the GPU work), joined every step; hdbscan.ms_per_scene is measured alone, untimed"
                      if hdb_stream is not None else "HDBSCAN on the step's stream"},
           "e2e": {"value": world * 1000.0 / ms_e2e, "unit": "scenes/s", "h2d_bytes_per_step": h2d_bytes,
                   "d2h_bytes_per_step": 4 + labels.nbytes, "ms_per_step": ms_e2e},
           "gpu_launches": int(launches), "clocks": clocks, "parity": PARITY,
           "hdbscan": {"thing_points": nt, "ms_per_scene": hd, "boruvka_rounds": rd, "clusters": int(labels.max()) + 1,
                       "noise_points": int((labels < 0).sum()), "share_of_step": hd / ms_dev},
           "roofline": {"bound": "hbm", "kernel": "pgs_hdb_mst (hdb_knn_kernel + hdb_search_kernel rounds + edge sort) + host "
                        "tree stage, timed as one call", "achieved": alg / (hd * 1e-3) / 1e9, "peak": peak,
                        "peak_source": peak_src, "unit": "GB/s", "frac": alg / (hd * 1e-3) / 1e9 / peak, "traffic": None,
                        "numerator": "SURVEY 8d: kNN 4nD+4n, Boruvka rounds x (n(4D+8)+12n), edges 12(n-1)"}}
    _emit(out)


# ------------------------------------------------------------------------------------------------
# B200 arm: C5 (tile sweep)
# ------------------------------------------------------------------------------------------------
def run_c5(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import sweep_c5
    sizes = [args.n] if args.n != N_POINTS else [250000, 500000, 1000000, 2000000]
    peak, peak_src = _peak()
    t = time.time()
    rows = sweep_c5.sweep(sizes, peak)
    big = [r for r in rows if r["N"] == max(r["N"] for r in rows)]
    conv16 = [r for r in big if r["kernel"] == "conv fwd" and r.get("C") == 16]
    per_tile_us = sum(r["us"] for r in big if r["kernel"].startswith(("cmap_build", "kmap_build")))
    out = {"metric": METRIC, "value": 1e6 / per_tile_us if per_tile_us else None, "unit": "tiles/s (hash + stride map + both "
           "rulebooks of the largest tile)", "n_gpus": 1, "steps": 5, "warmup": 2, "ms_per_step": per_tile_us / 1e3,
           "higher_is_better": True, "scaling": "replicas only", "vs_baseline": None, "dtype": "i32 / f32", "data": "synthetic",
           "config": {"workload": "C5: dense urban tile sweep, N in %s voxels at grid 0.12 m: coordinate hash, stride-2 map, "
                      "k3/s1 + k3/s2 rulebooks, occupancy sort, single conv fwd / bwd-input / bwd-weight at C in "
                      "{16,32,64,96,128,192}" % sizes},
           "roofline": ({"bound": "hbm", "kernel": "conv fwd 16->16 on the largest tile", "achieved": conv16[0]["gbps"],
                         "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": conv16[0]["frac_of_hbm_peak"],
                         "traffic": None} if conv16 else None),
           "sweep": rows, "wall_s": time.time() - t, "parity": PARITY}
    _emit(out)


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


def cpu_arm(n_full, steps, warmup, max_seconds, spr=REF_BATCH_SIZE):
    """Times the CPU restatement on the SAME workload (one batch of `spr` n_full-voxel cylinders, seeds 0.., all host
    threads): `warmup` untimed steps, then up to `steps` timed steps, stopping early once `max_seconds` of timed work are
    spent (at least one step is always measured; the count is reported as steps_measured)."""
    from oracle import cpu_path
    from panopticsegforlargescalepointcloud_b200 import scenes, panoptic, metrics
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    torch.set_num_threads(cores)
    opt = panoptic.paper_options("urban", cluster_type=1, grid=GRID, scorer=False)
    cfg = cpu_path.resolve_cfg(opt.backbone.config, 4)
    sd = _cpu_state()
    params = [v.requires_grad_(True) for k, v in sd.items() if v.dtype.is_floating_point and "running" not in k]
    state = {"t": 0, "m": [torch.zeros_like(p) for p in params], "v": [torch.zeros_like(p) for p in params]}
    b = make_inputs(list(range(spr)), n_full)
    for k in HOST_KEYS:
        if k not in ("syn_shifted", "syn_pred"):
            setattr(b, k, torch.as_tensor(getattr(b, k)))
    for _ in range(warmup):
        cpu_step(sd, cfg, b, opt.loss_weights, params, state)
    done, t0 = 0, time.time()
    while done < max(steps, 1):
        loss, clusters = cpu_step(sd, cfg, b, opt.loss_weights, params, state)
        done += 1
        if time.time() - t0 > max_seconds:
            break
    dt = (time.time() - t0) / done
    # PQ on the seed-0 cylinder of the batch (the scene the B200 arm scores too)
    m0 = b.batch.numpy() == 0
    n0 = int(m0.sum())
    c0 = [c for c in clusters if len(c) and c[0] < n0]
    pq = metrics.panoptic_quality(b.syn_pred[:n0], c0, b.y.numpy()[:n0], b.instance_labels.numpy()[:n0], 9,
                                  list(scenes.URBAN_THINGS))
    return {"value": spr / dt, "unit": "scenes/s", "cores": cores, "kind": "port",
            "sample": "%d step(s) (%d warm-up) on the workload's own batch: %d cylinders x %d voxels (seeds 0..%d), "
                      "%.2f s/step; no scaling" % (done, warmup, spr, n_full, spr - 1, dt),
            "same_config": True, "steps_measured": done, "ms_per_step": dt * 1e3, "pq_sample": pq,
            "threads": torch.get_num_threads()}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.config not in ("C2", "C4"):
        _emit({"impl": "reference", "unavailable": "the CPU arm times the C2 step only (config %s)" % args.config})
        return
    # the same 200 k-voxel cylinder as the B200 arm; ~6-8 s per step on 16-32 host cores, so the number of measured steps
    # is bounded by time (never by shrinking the scene) and printed
    spr = 8 if args.config == "C4" else args.scenes_per_gpu
    cb = cpu_arm(args.n, steps=args.steps, warmup=0, max_seconds=150.0, spr=spr)
    ms = 1000.0 / cb["value"]
    _emit({"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "scenes/s",
                      "n_gpus": args.gpus, "steps": args.steps, "steps_measured": cb["steps_measured"],
                      "warmup": 0, "ms_per_step": ms * spr,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                      "data": "synthetic", "config": _cfg_workload(args.n, spr, args.config), "same_config": True,
                      "cpu_baseline": cb,
                      "e2e": {"value": cb["value"], "unit": "scenes/s", "h2d_bytes_per_step": 0,
                              "d2h_bytes_per_step": 0},
                      "gpu_launches": 0, "parity": PARITY})


_OUT = None


def _emit(obj):
    """The bench line, on the process's original stdout (main() points descriptor 1 at stderr for everybody else)."""
    f = _OUT if _OUT is not None else sys.stdout
    f.write(json.dumps(obj) + "\n")
    f.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C2", choices=["C2", "C3", "C4", "C5"])
    ap.add_argument("--scenes-per-gpu", type=int, default=REF_BATCH_SIZE,
                    help="C2: cylinders collated into one step per GPU (default: the reference's training batch size)")
    ap.add_argument("--n", type=int, default=N_POINTS)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    # stdout carries the ONE JSON line and nothing else: libraries that write to file descriptor 1 (NCCL prints its
    # version banner there) are sent to stderr, the line itself goes to the saved descriptor
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
        return
    if args.warmup < 3:
        args.warmup = 3
    if args.config == "C3":
        run_c3(args)
    elif args.config == "C5":
        run_c5(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
