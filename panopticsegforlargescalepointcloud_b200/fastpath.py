"""Fused executor for the sparse ResUNet: the whole backbone as ONE autograd node.

Why: the module-by-module mirror of the reference (`backbone.py` on top of `me.py`, same classes and call order as
torch_points3d/modules/MinkowskiEngine/api_modules.py:9-82,235-311 and applications/minkowski.py:160-196) costs
~50 us of Python / autograd bookkeeping per layer call: 82 convolutions + 82 batch norms + ReLUs, sums and
concatenations, forward and backward, add up to ~25 ms of host time per step -- more than the B200 needs for the
kernels (profiles/r1_host_profile.txt).  The network structure is static, so it is compiled once into a flat tape
(conv / bn(+relu) / add / cat over numbered feature slots); a step then is

    forward : one pass over the tape, one C-ABI call per op, activations in one arena allocation
    backward: the tape in reverse -- input gradient = the same conv kernel on the sibling table with W^T, weight
              gradient accumulated straight into `param.grad` (when it exists), fused BN(+ReLU) backward, gradients
              of multiply-used tensors (residual inputs, skip connections) merged with one add kernel

The kernels, their dispatch (`me._conv_launch`) and therefore the results are the ones of the module path; the
module path stays as the reference implementation of this file (tests/test_gpu_fastpath.py compares outputs,
input / parameter gradients and BN running statistics) and is used whenever the module tree is not the plain
ResNetDown / ResNetUp / ResBlock structure (`Unsupported`).  PGS_FASTPATH=0 disables it.
"""
import contextlib
import ctypes
import os

import numpy as np
import torch

from . import _lib
from . import me as ME
from ._lib import check

ENABLED = os.environ.get("PGS_FASTPATH", "1") == "1"
# weight-gradient kernels on a second stream: they depend only on a layer's input activations and output gradient, not on
# the input-gradient chain, and the small / mid-size layers of that chain leave most SMs idle
DW_SIDE_STREAM = os.environ.get("PGS_DW_SIDE_STREAM", "1") == "1"
# kernel maps / sorted tables / pair lists of the deeper levels built on the second stream while the first layers run:
# measured no gain (24.9 vs 24.6 ms per step) because the forward pass is bound by the host's launch rate, so it is off
MAPS_SIDE_STREAM = os.environ.get("PGS_MAPS_SIDE_STREAM", "0") == "1"
_SIDE = {}


def _side_stream(dev):
    s = _SIDE.get(dev.index)
    if s is None:
        s = _SIDE[dev.index] = torch.cuda.Stream(device=dev)
    return s

OP_CONV, OP_BN, OP_ADD, OP_CAT = 0, 1, 2, 3
BN_ACC, BN_ZEROED = 1, 2   # include/pgs_b200.h: PGS_BN_ACCUMULATE_PARAM_GRADS, PGS_BN_SUMS_ZEROED

# "native": pass 2 of the forward and the whole backward walk the tape inside libpgs_b200.so (pgs_unet_forward /
# pgs_unet_backward, csrc/unet_exec.cpp) -- one ctypes call per direction instead of one per op (~500 per step);
# "python": the per-op loops below (same kernels, same order; used when bench.py records per-launch events).
EXECUTOR = os.environ.get("PGS_EXECUTOR", "native")

# host mirrors of the records in include/pgs_b200.h (8-byte fields first: no padding)
OP_DT = np.dtype([("kind", "i4"), ("a", "i4"), ("b", "i4"), ("dst", "i4"), ("idx", "i4"), ("relu", "i4")])
CONV_DT = np.dtype([("W", "u8"), ("dW", "u8"), ("nbr_f", "u8"), ("order_f", "u8"), ("nbr_b", "u8"), ("order_b", "u8"),
                    ("pair_in", "u8"), ("pair_out", "u8"), ("pair_offs", "u8"), ("wprep_f", "u8"), ("wprep_b", "u8"),
                    ("wprep_bytes", "u8"), ("max_pairs", "i8"), ("K", "i4"), ("c_in", "i4"), ("c_out", "i4"),
                    ("kind_f", "i4"), ("kind_b", "i4"), ("mirror_f", "i4"), ("mirror_b", "i4"), ("need_dx", "i4")])
BN_DT = np.dtype([("weight", "u8"), ("bias", "u8"), ("running_mean", "u8"), ("running_var", "u8"), ("dweight", "u8"),
                  ("dbias", "u8"), ("momentum", "f4"), ("eps", "f4"), ("training", "i4"), ("accumulate", "i4")])
KIND_CODE = {"ffma": 0, "tc": 1, "mma": 2, "split": 3}


def _vp(arr):
    return ctypes.c_void_p(arr.ctypes.data)


def _table_ptrs(km, kind, n_q):
    """(table, order) device pointers a conv kernel of `kind` reads for kernel map `km` (me._conv_launch's rule)."""
    if km is None:
        return 0, 0
    if kind != "ffma" and ME.SORT_TABLES and n_q >= ME.SORT_MIN_ROWS:
        ns, order = km.sorted()
        return ns.data_ptr(), order.data_ptr()
    return km.nbr.data_ptr(), 0


class Unsupported(Exception):
    pass


class Program:
    """Static tape of a MinkowskiUnet / MinkowskiEncoder (built once per model)."""

    def __init__(self, net):
        self.ops = []       # (kind, a, b, dst, index into convs / bns, relu)
        self.convs = []
        self.bns = []
        self.n_slots = 1    # slot 0 = network input
        x = 0
        # duck-typed on purpose: the same tape serves backbone.py's mirror classes AND the reference's own
        # MinkowskiUnet / MinkowskiEncoder (applications/minkowski.py:129-196) built from its own ResNetDown / ResNetUp /
        # ResBlock (modules/MinkowskiEngine/api_modules.py) over me.py -- see bind.install(fuse_unet=True)
        downs = getattr(net, "down_modules", None)
        ups = getattr(net, "up_modules", None)
        if downs is None or len(downs) == 0:
            raise Unsupported("no down_modules")
        if ups is not None and len(ups) > 0:
            stack = []
            for i in range(len(downs) - 1):
                x = self._resnet(downs[i], x)
                stack.append(x)
            x = self._resnet(downs[-1], x)
            stack.append(None)
            for m in ups:
                skip = stack.pop()
                if skip is not None:
                    x = self._emit(OP_CAT, x, skip)
                x = self._resnet(m, x)
        else:
            for m in downs:
                x = self._resnet(m, x)
        self.out_slot = x
        self.params = ([c.kernel for c in self.convs] + [b.bn.weight for b in self.bns] + [b.bn.bias for b in self.bns])
        self.consumers = [0] * self.n_slots
        for kind, a, b, dst, idx, relu in self.ops:
            self.consumers[a] += 1
            if b >= 0:
                self.consumers[b] += 1
        self.wprep = None        # arranged weights: per conv a forward and a backward slot
        self.wprep_off = None
        self.desc_cache = {}     # (kernel kinds, weight pointers, backward needed) -> device descriptor table
        self.prep_gen = 0        # bumped by every forward that re-arranges the weights (see _UNetFn.backward)
        self.ops_np = np.array([(k, a, b, d, i, int(r)) for k, a, b, d, i, r in self.ops], dtype=OP_DT)

    def out_tensor_stride(self, ts0):
        ts = {0: ts0}
        for kind, a, b, dst, idx, relu in self.ops:
            t = ts[a]
            if kind == OP_CONV:
                mod = self.convs[idx]
                t = t // mod.stride if mod.TRANSPOSE else t * mod.stride
            ts[dst] = t
        return ts[self.out_slot]

    # ---- tape construction -------------------------------------------------------------------
    def _emit(self, kind, a, b=-1, idx=-1, relu=False):
        dst = self.n_slots
        self.n_slots += 1
        self.ops.append((kind, a, b, dst, idx, relu))
        return dst

    def _conv(self, mod, src):
        if type(mod) not in (ME.MinkowskiConvolution, ME.MinkowskiConvolutionTranspose) or mod.bias is not None:
            raise Unsupported("convolution %r" % (mod,))
        self.convs.append(mod)
        return self._emit(OP_CONV, src, idx=len(self.convs) - 1)

    def _bn(self, mod, src, relu):
        if type(mod) is not ME.MinkowskiBatchNorm:
            raise Unsupported("normalisation %r" % (mod,))
        bn = mod.bn
        if not (bn.affine and bn.track_running_stats and bn.num_features % 4 == 0 and bn.num_features <= 1024):
            raise Unsupported("batch norm configuration")
        self.bns.append(mod)
        return self._emit(OP_BN, src, idx=len(self.bns) - 1, relu=relu)

    def _conv_bn(self, seq, src):
        mods = list(seq)
        if len(mods) == 3 and type(mods[2]) is ME.MinkowskiReLU:
            relu = True
        elif len(mods) == 2:
            relu = False
        else:
            raise Unsupported("conv block %r" % (seq,))
        return self._bn(mods[1], self._conv(mods[0], src), relu)

    def _resnet(self, m, src):
        if not (hasattr(m, "conv_in") and hasattr(m, "blocks")) or type(m).__name__ not in ("ResNetDown", "ResNetUp"):
            raise Unsupported("module %r" % type(m))
        y = self._conv_bn(m.conv_in, src)
        for blk in (m.blocks if m.blocks is not None else ()):
            if type(blk).__name__ != "ResBlock" or not hasattr(blk, "block") or not hasattr(blk, "downsample"):
                raise Unsupported("block %r" % type(blk))
            mods = list(blk.block)
            if len(mods) != 6 or type(mods[2]) is not ME.MinkowskiReLU or type(mods[5]) is not ME.MinkowskiReLU:
                raise Unsupported("residual block layout")
            a = self._bn(mods[1], self._conv(mods[0], y), True)
            b = self._bn(mods[4], self._conv(mods[3], a), True)
            s = y if blk.downsample is None else self._conv_bn(blk.downsample, y)
            y = self._emit(OP_ADD, b, s)
        return y


def _stream():
    return _lib.stream_ptr()


class _UNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, prog, cm, ts0, X, *params):
        lib = _lib.load()
        sp = _stream()
        X = X.contiguous()
        dev = X.device
        S = prog.n_slots
        n, C, ts = [0] * S, [0] * S, [0] * S
        n[0], C[0], ts[0] = X.shape[0], X.shape[1], ts0
        # ---- pass 1: shapes, coordinate / kernel maps (the only data-dependent part) ----
        # The strided coordinate maps (whose sizes the host must read back) are built first on the main stream; the
        # kernel maps, their occupancy-sorted copies and the pair lists of the weight gradient then go to a second
        # stream in order of first use, so that the tables of the deeper levels are built while the first layers
        # already run (each convolution waits for the event of its own tables).
        need_bwd = any(ctx.needs_input_grad)
        torch.cuda.nvtx.range_push("pgs.unet.maps")       # pass 1: coordinate / kernel maps (closed before pass 2)
        try:
            maps_side = MAPS_SIDE_STREAM
            if maps_side:
                t = {0: ts0}
                for kind, a, b, dst, idx, relu in prog.ops:
                    t[dst] = t[a]
                    if kind == OP_CONV:
                        mod = prog.convs[idx]
                        if mod.stride > 1:
                            if mod.TRANSPOSE:
                                t[dst] = t[a] // mod.stride
                            else:
                                t[dst] = t[a] * mod.stride
                                cm.stride(t[a], t[dst])
                main = torch.cuda.current_stream(dev)
                side = _side_stream(dev)
                side.wait_stream(main)   # the coordinate maps; and everything the previous step still reads from old tables
            cinfo = [None] * len(prog.convs)
            cevent = [None] * len(prog.convs)
            with (torch.cuda.stream(side) if maps_side else contextlib.nullcontext()):
                for kind, a, b, dst, idx, relu in prog.ops:
                    if kind == OP_CONV:
                        mod = prog.convs[idx]
                        if C[a] != mod.in_channels:
                            raise ValueError("expected %d input channels, got %d" % (mod.in_channels, C[a]))
                        K = mod.kernel_size ** 3
                        before = len(cm.kmaps)
                        km_f, km_b, mf, mb, ts_out, n_out = mod.maps_for(cm, ts[a], n[a])
                        kind_f = ME._conv_kernel_choice(lib, K, mod.in_channels, mod.out_channels, n_out, km_f is not None)
                        kind_b = ME._conv_kernel_choice(lib, K, mod.out_channels, mod.in_channels, n[a], km_b is not None)
                        if maps_side and km_f is not None:
                            fresh = len(cm.kmaps) != before
                            for km, kd, nq in ((km_f, kind_f, n_out), (km_b, kind_b, n[a])):
                                if (kd != "ffma" and ME.SORT_TABLES and nq >= ME.SORT_MIN_ROWS and km._sorted is None
                                        and (km is km_f or need_bwd)):
                                    km.sorted()
                                    fresh = True
                            if need_bwd and km_f._pairs is None:
                                km_f.pairs()
                                fresh = True
                            if fresh:
                                cevent[idx] = torch.cuda.Event()
                                cevent[idx].record(side)
                        cinfo[idx] = (km_f, km_b, mf, mb, K, kind_f, kind_b)
                        n[dst], C[dst], ts[dst] = n_out, mod.out_channels, ts_out
                    elif kind == OP_BN:
                        if prog.bns[idx].bn.momentum is None:
                            raise Unsupported("cumulative-average batch norm")
                        n[dst], C[dst], ts[dst] = n[a], C[a], ts[a]
                    elif kind == OP_ADD:
                        if (n[a], C[a], ts[a]) != (n[b], C[b], ts[b]):
                            raise ValueError("sum of sparse tensors on different maps")
                        n[dst], C[dst], ts[dst] = n[a], C[a], ts[a]
                    else:
                        if (n[a], ts[a]) != (n[b], ts[b]):
                            raise ValueError("concatenation of sparse tensors on different maps")
                        n[dst], C[dst], ts[dst] = n[a], C[a] + C[b], ts[a]
        finally:
            torch.cuda.nvtx.range_pop()
        if min(n) <= 0:
            raise Unsupported("empty level")
        # ---- arenas ----
        off = [0] * (S + 1)
        for s in range(1, S):
            off[s + 1] = off[s] + n[s] * C[s]
        arena = torch.empty(off[S], dtype=torch.float32, device=dev)
        base = arena.data_ptr()
        ptrs = [base + 4 * o for o in off[:S]]
        ptrs[0] = X.data_ptr()
        nbn = len(prog.bns)
        soff = [0] * (nbn + 1)
        for i, m in enumerate(prog.bns):
            soff[i + 1] = soff[i] + 2 * m.bn.num_features
        stats = torch.empty(soff[nbn], dtype=torch.float32, device=dev)
        sums = torch.zeros(soff[nbn], dtype=torch.float64, device=dev)
        stats_p, sums_p = stats.data_ptr(), sums.data_ptr()
        nconv = len(prog.convs)
        # ---- arranged weights: forward and backward layout of every convolution in ONE launch ----
        if prog.wprep is None or prog.wprep.device != dev:
            offs, tot = [], 0
            for mod in prog.convs:
                sz = (mod.kernel_size ** 3 * mod.in_channels * mod.out_channels * 8 + 255) // 256 * 256
                offs.append((tot, tot + sz, sz))
                tot += 2 * sz
            prog.wprep = torch.empty(max(tot, 1), dtype=torch.uint8, device=dev)
            prog.wprep_off = offs
            prog.desc_cache = {}
        key = (tuple((c[5], c[6]) for c in cinfo), tuple(p.data_ptr() for p in params[:nconv]), need_bwd)
        ent = prog.desc_cache.get(key)
        if ent is None:
            wbase = prog.wprep.data_ptr()
            rows, mx = [], 1
            for i, mod in enumerate(prog.convs):
                K, kf, kb = cinfo[i][4], cinfo[i][5], cinfo[i][6]
                of, ob, _ = prog.wprep_off[i]
                wp = params[i].data_ptr()
                if kf != "ffma":
                    rows.append([wp, wbase + of, K, mod.in_channels, mod.out_channels, 0, 0 if kf == "tc" else 1, 0])
                if need_bwd and kb != "ffma":
                    rows.append([wp, wbase + ob, K, mod.out_channels, mod.in_channels, 1, 0 if kb == "tc" else 1, 0])
                mx = max(mx, K * mod.in_channels * mod.out_channels)
            if len(prog.desc_cache) > 16:
                prog.desc_cache.clear()
            ent = (torch.tensor(rows, dtype=torch.int64).to(dev) if rows else None, len(rows), mx)
            prog.desc_cache[key] = ent
        if ent[1]:
            rc = lib.pgs_conv_prep_weights_batch(ent[0].data_ptr(), ent[1], ent[2], sp)
            if rc:
                check(rc)
        prog.prep_gen += 1
        wbase = prog.wprep.data_ptr()
        training = [False] * nbn
        tracked = []
        for i, m in enumerate(prog.bns):
            bn = m.bn
            training[i] = bool(bn.training)
            if training[i] and bn.num_batches_tracked is not None:
                tracked.append(bn.num_batches_tracked)
        native = EXECUTOR == "native" and ME.PROFILE is None
        rec = None
        if native:
            if maps_side:
                main.wait_stream(side)
            # ---- pass 2, native: this step's pointers into the host records, then ONE call walks the tape ----
            cv = np.zeros(nconv, CONV_DT)
            for i, mod in enumerate(prog.convs):
                km_f, km_b, mf, mb, K, kind_f, kind_b = cinfo[i]
                of, ob, sz = prog.wprep_off[i]
                r = cv[i]
                r["W"] = params[i].data_ptr()
                r["wprep_f"], r["wprep_b"], r["wprep_bytes"] = wbase + of, wbase + ob, sz
                r["K"], r["c_in"], r["c_out"] = K, mod.in_channels, mod.out_channels
                r["kind_f"], r["kind_b"], r["mirror_f"], r["mirror_b"] = KIND_CODE[kind_f], KIND_CODE[kind_b], mf, mb
            for op in prog.ops:
                if op[0] == OP_CONV:
                    i, a, dst = op[4], op[1], op[3]
                    km_f, km_b, mf, mb, K, kind_f, kind_b = cinfo[i]
                    r = cv[i]
                    r["nbr_f"], r["order_f"] = _table_ptrs(km_f, kind_f, n[dst])
                    if need_bwd:
                        r["nbr_b"], r["order_b"] = _table_ptrs(km_b, kind_b, n[a])
            bv = np.zeros(nbn, BN_DT)
            for i, m in enumerate(prog.bns):
                bn = m.bn
                r = bv[i]
                r["weight"], r["bias"] = params[nconv + i].data_ptr(), params[nconv + nbn + i].data_ptr()
                r["running_mean"], r["running_var"] = bn.running_mean.data_ptr(), bn.running_var.data_ptr()
                r["momentum"], r["eps"], r["training"] = bn.momentum, bn.eps, int(training[i])
            rec = (cv, bv, np.array(ptrs, np.uint64), np.array(n, np.int64), np.array(C, np.int32),
                   np.array(soff[:nbn] if nbn else [0], np.int64))
            cv, bv, slot_ptr, slot_n, slot_c, soff_np = rec
            rc = lib.pgs_unet_forward(_vp(prog.ops_np), len(prog.ops), S, _vp(slot_ptr), _vp(slot_n), _vp(slot_c), _vp(cv),
                                      _vp(bv), sums_p, stats_p, _vp(soff_np), sp)
            if rc:
                check(rc)
        # ---- pass 2, per-op Python loop (bench.py's per-launch event pass; PGS_EXECUTOR=python) ----
        for kind, a, b, dst, idx, relu in (prog.ops if not native else ()):
            if kind == OP_CONV:
                mod = prog.convs[idx]
                km_f, km_b, mf, mb, K, kind_f, kind_b = cinfo[idx]
                of, ob, sz = prog.wprep_off[idx]
                if cevent[idx] is not None:
                    main.wait_event(cevent[idx])   # this conv's tables were built on the side stream
                ME._conv_launch(lib, ptrs[a], n[a], params[idx].data_ptr(), K, mod.in_channels, mod.out_channels, km_f,
                                n[dst], mf, False, ptrs[dst], wbase + of, sz, sp, kind=kind_f, prepped=True)
            elif kind == OP_BN:
                bn = prog.bns[idx].bn
                Cb = C[a]
                rc = lib.pgs_bn_forward_ex(ptrs[a], n[a], Cb, params[nconv + idx].data_ptr(),
                                           params[nconv + nbn + idx].data_ptr(), bn.running_mean.data_ptr(),
                                           bn.running_var.data_ptr(), int(training[idx]), float(bn.momentum), float(bn.eps),
                                           int(relu), BN_ZEROED, sums_p + 8 * soff[idx], stats_p + 4 * soff[idx],
                                           stats_p + 4 * (soff[idx] + Cb), ptrs[dst], sp)
                if rc:
                    check(rc)
            elif kind == OP_ADD:
                rc = lib.pgs_add2(ptrs[a], ptrs[b], ptrs[dst], n[a] * C[a], sp)
                if rc:
                    check(rc)
            else:
                rc = lib.pgs_cat2(ptrs[a], C[a], ptrs[b], C[b], ptrs[dst], n[a], 0, sp)
                if rc:
                    check(rc)
        if tracked:
            torch._foreach_add_(tracked, 1)
        o = prog.out_slot
        out = arena[off[o]:off[o] + n[o] * C[o]].view(n[o], C[o])
        ctx.prog, ctx.rt = prog, (n, C, ptrs, cinfo, soff, training, arena, stats)
        ctx.rec, ctx.prep = rec, (prog.prep_gen, ent)
        ctx.save_for_backward(X, *params)
        return out

    @staticmethod
    def backward(ctx, dOut):
        lib = _lib.load()
        sp = _stream()
        prog = ctx.prog
        n, C, ptrs, cinfo, soff, training, arena, stats = ctx.rt
        saved = ctx.saved_tensors
        X, params = saved[0], saved[1:]
        dev = X.device
        dOut = dOut.contiguous()
        nconv, nbn = len(prog.convs), len(prog.bns)
        need_x = ctx.needs_input_grad[3]
        # ---- gradient arena: one buffer per conv / bn input, per cat input, per merge of a multiply-used slot ----
        total = 0
        for kind, a, b, dst, idx, relu in prog.ops:
            if kind in (OP_CONV, OP_BN):
                total += n[a] * C[a]
            elif kind == OP_CAT:
                total += n[a] * (C[a] + C[b])
        for s in range(prog.n_slots):
            if prog.consumers[s] > 1:
                total += (prog.consumers[s] - 1) * n[s] * C[s]
        garena = torch.empty(total, dtype=torch.float32, device=dev)
        gbase, gused = garena.data_ptr(), 0
        sums = torch.zeros(soff[nbn], dtype=torch.float64, device=dev)
        sums_p, stats_p = sums.data_ptr(), stats.data_ptr()
        wbase = prog.wprep.data_ptr()
        # parameter gradients: straight into param.grad where it exists (accumulate), else into one zeroed buffer
        P = prog.params
        needs = ctx.needs_input_grad[4:]
        direct = [needs[i] and P[i].grad is not None and P[i].grad.is_contiguous() and P[i].grad.dtype == torch.float32
                  and P[i].grad.device == dev for i in range(len(P))]
        goff = [0] * (len(P) + 1)
        for i, p in enumerate(P):
            goff[i + 1] = goff[i] + (p.numel() if (needs[i] and not direct[i]) else 0)
        gflat = torch.zeros(goff[-1], dtype=torch.float32, device=dev) if goff[-1] else None
        gflat_p = gflat.data_ptr() if gflat is not None else 0

        def pgrad(i):
            if not needs[i]:
                return None
            return P[i].grad.data_ptr() if direct[i] else gflat_p + 4 * goff[i]

        # the arranged-weight buffer belongs to the Program and is rewritten by every forward: if another forward ran
        # since ours (a second batch, an eval pass), put this graph's backward layouts back first
        gen, ent = ctx.prep
        if prog.prep_gen != gen and ent[1]:
            rc = lib.pgs_conv_prep_weights_batch(ent[0].data_ptr(), ent[1], ent[2], sp)
            if rc:
                check(rc)
            prog.prep_gen += 1
            ctx.prep = (prog.prep_gen, ent)
        if ctx.rec is not None and ME.PROFILE is None and ME.PROFILE_DW is None:
            return _UNetFn._backward_native(ctx, lib, sp, dOut, garena, total, sums_p, stats_p, pgrad, direct, needs,
                                            need_x, gflat, goff)
        side = sp_side = None
        if DW_SIDE_STREAM:
            main = torch.cuda.current_stream(dev)
            side = _side_stream(dev)
            side.wait_stream(main)          # activations, zeroed gradient buffers, dOut
            sp_side = _lib.c_void_p(side.cuda_stream)
        glist = [[] for _ in range(prog.n_slots)]
        glist[prog.out_slot].append(dOut.data_ptr())
        for kind, a, b, dst, idx, relu in reversed(prog.ops):
            gl = glist[dst]
            if not gl:
                continue
            g = gl[0]
            for extra in gl[1:]:   # tensor with several consumers: sum their gradients
                buf = gbase + 4 * gused
                gused += n[dst] * C[dst]
                rc = lib.pgs_add2(g, extra, buf, n[dst] * C[dst], sp)
                if rc:
                    check(rc)
                g = buf
            if kind == OP_CONV:
                mod = prog.convs[idx]
                km_f, km_b, mf, mb, K, kind_f, kind_b = cinfo[idx]
                cin, cout = mod.in_channels, mod.out_channels
                wp = params[idx].data_ptr()
                if a != 0 or need_x:
                    dx = gbase + 4 * gused
                    gused += n[a] * C[a]
                    of, ob, sz = prog.wprep_off[idx]
                    ME._conv_launch(lib, g, n[dst], wp, K, cout, cin, km_b, n[a], mb, True, dx, wbase + ob, sz, sp,
                                    kind=kind_b, prepped=True)
                    glist[a].append(dx)
                dw = pgrad(idx)
                if dw is not None:
                    spw = sp
                    pl = km_f.pairs() if km_f is not None else None   # (built on the main stream on first use)
                    if side is not None:        # g and the pair lists are complete on the main stream here
                        ev = torch.cuda.Event()
                        ev.record(main)
                        side.wait_event(ev)
                        spw = sp_side
                    if pl is not None:
                        in_idx, out_idx, offs, max_pairs = pl
                        rc = lib.pgs_conv_bwd_weight(ptrs[a], g, in_idx.data_ptr(), out_idx.data_ptr(), offs.data_ptr(),
                                                     max_pairs, K, cin, cout, int(mf), dw, spw)
                    else:
                        rc = lib.pgs_conv_bwd_weight(ptrs[a], g, None, None, None, n[a], 1, cin, cout, 0, dw, spw)
                    if rc:
                        check(rc)
            elif kind == OP_BN:
                Cb = C[a]
                dx = gbase + 4 * gused
                gused += n[a] * Cb
                iw, ib = nconv + idx, nconv + nbn + idx
                flags = BN_ZEROED | (BN_ACC if (direct[iw] or direct[ib]) else 0)
                if direct[iw] != direct[ib] and needs[iw] and needs[ib]:
                    raise RuntimeError("batch-norm weight and bias must both have (or both lack) a .grad buffer")
                rc = lib.pgs_bn_backward_ex(ptrs[a], None, g, n[a], Cb, params[iw].data_ptr(), params[ib].data_ptr(),
                                            stats_p + 4 * soff[idx], stats_p + 4 * (soff[idx] + Cb), int(training[idx]),
                                            int(relu), flags, sums_p + 8 * soff[idx], dx, pgrad(iw), pgrad(ib), sp)
                if rc:
                    check(rc)
                glist[a].append(dx)
            elif kind == OP_ADD:
                glist[a].append(g)
                glist[b].append(g)
            else:
                ga = gbase + 4 * gused
                gused += n[a] * C[a]
                gb = gbase + 4 * gused
                gused += n[b] * C[b]
                rc = lib.pgs_cat2(ga, C[a], gb, C[b], g, n[a], 1, sp)
                if rc:
                    check(rc)
                glist[a].append(ga)
                glist[b].append(gb)
        assert gused <= total
        if side is not None:
            main.wait_stream(side)      # gradients complete (and the arenas reusable) for whatever follows on the main stream
        dX = None
        if need_x:
            g0 = glist[0]
            off0 = (g0[0] - gbase) // 4
            dX = garena[off0:off0 + n[0] * C[0]].view(n[0], C[0])
            for extra in g0[1:]:
                o2 = (extra - gbase) // 4
                dX = dX + garena[o2:o2 + n[0] * C[0]].view(n[0], C[0])
        grads = []
        for i, p in enumerate(P):
            if not needs[i] or direct[i]:
                grads.append(None)
            else:
                grads.append(gflat[goff[i]:goff[i + 1]].view(p.shape))
        return (None, None, None, dX) + tuple(grads)

    @staticmethod
    def _backward_native(ctx, lib, sp, dOut, garena, total, sums_p, stats_p, pgrad, direct, needs, need_x, gflat, goff):
        prog = ctx.prog
        n, C, ptrs, cinfo, soff, training, arena, stats = ctx.rt
        cv, bv, slot_ptr, slot_n, slot_c, soff_np = ctx.rec
        dev = dOut.device
        nconv, nbn = len(prog.convs), len(prog.bns)
        first_ops = {}
        for op in prog.ops:
            if op[0] == OP_CONV:
                first_ops[op[4]] = op
        for i in range(nconv):
            km_f = cinfo[i][0]
            a = first_ops[i][1]
            r = cv[i]
            dw = pgrad(i)
            r["dW"] = dw or 0
            r["need_dx"] = int(a != 0 or need_x)
            if dw and km_f is not None:
                in_idx, out_idx, offs, max_pairs = km_f.pairs()
                r["pair_in"], r["pair_out"], r["pair_offs"] = in_idx.data_ptr(), out_idx.data_ptr(), offs.data_ptr()
                r["max_pairs"] = max_pairs
        for i in range(nbn):
            iw, ib = nconv + i, nconv + nbn + i
            if direct[iw] != direct[ib] and needs[iw] and needs[ib]:
                raise RuntimeError("batch-norm weight and bias must both have (or both lack) a .grad buffer")
            r = bv[i]
            r["dweight"], r["dbias"] = pgrad(iw) or 0, pgrad(ib) or 0
            r["accumulate"] = int(direct[iw] or direct[ib])
        side_p = None
        if DW_SIDE_STREAM:
            side_p = _lib.c_void_p(_side_stream(dev).cuda_stream)
        g_in = (ctypes.c_void_p * 8)()
        n_in = ctypes.c_int32(0)
        rc = lib.pgs_unet_backward(_vp(prog.ops_np), len(prog.ops), prog.n_slots, prog.out_slot, _vp(slot_ptr), _vp(slot_n),
                                   _vp(slot_c), _vp(cv), _vp(bv), sums_p, stats_p, _vp(soff_np), dOut.data_ptr(),
                                   garena.data_ptr(), total, g_in, ctypes.byref(n_in), sp, side_p)
        if rc:
            check(rc)
        dX = None
        if need_x:
            gbase = garena.data_ptr()
            for j in range(n_in.value):
                o = (g_in[j] - gbase) // 4
                part = garena[o:o + n[0] * C[0]].view(n[0], C[0])
                dX = part if dX is None else dX + part
        P = prog.params
        grads = []
        for i, p in enumerate(P):
            if not needs[i] or direct[i]:
                grads.append(None)
            else:
                grads.append(gflat[goff[i]:goff[i + 1]].view(p.shape))
        return (None, None, None, dX) + tuple(grads)


def build_strided_maps(prog, cm, ts0=1):
    """Every strided coordinate map `prog` will ask `cm` for (the part of pass 1 that reads sizes back to the host)."""
    t = {0: ts0}
    for kind, a, b, dst, idx, relu in prog.ops:
        t[dst] = t[a]
        if kind == OP_CONV:
            mod = prog.convs[idx]
            if mod.stride > 1:
                if mod.TRANSPOSE:
                    t[dst] = t[a] // mod.stride
                else:
                    t[dst] = t[a] * mod.stride
                    cm.stride(t[a], t[dst])


def program_for(net):
    """Compiled tape of `net` (cached on the module), or None when its structure is not supported."""
    prog = net.__dict__.get("_pgs_program", False)
    if prog is False:
        try:
            prog = Program(net)
        except Unsupported:
            prog = None
        net.__dict__["_pgs_program"] = prog
    return prog


def run(net, x):
    """Output SparseTensor of the network for the input SparseTensor `x`, or None when the fast path does not apply."""
    if not ENABLED:
        return None
    prog = program_for(net)
    if prog is None:
        return None
    try:
        F = _UNetFn.apply(prog, x.coordinate_manager, x.tensor_stride, x.F, *prog.params)
    except Unsupported:
        return None
    return ME.SparseTensor(F, coordinate_manager=x.coordinate_manager, tensor_stride=prog.out_tensor_stride(x.tensor_stride))
