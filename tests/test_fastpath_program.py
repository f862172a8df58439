"""Host logic of the fused executor that needs no GPU: the tape compiled from the module tree, and the kernel dispatch
table (CPU; the kernels themselves are covered by the -m gpu tests)."""
import pytest
import torch

from panopticsegforlargescalepointcloud_b200 import backbone as bb, fastpath, me


def test_paper_backbone_compiles_to_the_expected_tape():
    """82 sparse convs + 82 BatchNorms, 28 residual sums, 6 skip concatenations, output on the input map."""
    net = bb.Minkowski("unet", input_nc=4, config=bb.paper_backbone_config(16))
    p = fastpath.Program(net)
    kinds = [op[0] for op in p.ops]
    assert kinds.count(fastpath.OP_CONV) == 82 and kinds.count(fastpath.OP_BN) == 82
    assert kinds.count(fastpath.OP_ADD) == 28 and kinds.count(fastpath.OP_CAT) == 6
    assert len(p.params) == 82 * 3 and p.out_tensor_stride(1) == 1
    # parameter order = conv kernels, BN weights, BN biases, each in tape order; every module parameter exactly once
    assert {id(q) for q in p.params} == {id(q) for q in net.parameters()}
    assert all(q is c.kernel for q, c in zip(p.params[:82], p.convs))
    # a BN absorbs the ReLU exactly where the reference has conv-BN-ReLU (not on the 1x1 shortcut branch)
    relu = [op[5] for op in p.ops if op[0] == fastpath.OP_BN]
    assert relu.count(False) == 12            # 12 blocks change their channel count -> 12 shortcut conv-BN pairs
    # slots used more than once: residual inputs and skip connections (their gradients get merged in backward)
    assert sum(1 for c in p.consumers if c > 1) == 34
    # every slot is written exactly once, in order
    assert [op[3] for op in p.ops] == list(range(1, p.n_slots))


def test_encoder_and_two_level_unet_compile():
    p = fastpath.Program(bb.Minkowski("unet", input_nc=4, config=bb.two_level_config(16)))
    assert p.out_tensor_stride(1) == 1
    cfg = {"down_conv": bb.paper_backbone_config(16)["down_conv"], "define_constants": {"in_feat": 16}}
    enc = bb.Minkowski("encoder", input_nc=4, config=cfg)
    assert fastpath.Program(enc).out_tensor_stride(1) == 64


def test_unknown_module_trees_are_not_compiled():
    net = bb.Minkowski("unet", input_nc=4, config=bb.two_level_config(16))
    net.down_modules[0].conv_in[2] = me.MinkowskiLeakyReLU(0.1)
    with pytest.raises(fastpath.Unsupported):
        fastpath.Program(net)
    assert fastpath.program_for(net) is None
    net2 = bb.Minkowski("unet", input_nc=4, config=bb.two_level_config(16))
    net2.down_modules[1].blocks[0].block[1].bn = torch.nn.BatchNorm1d(32, affine=False)
    assert fastpath.program_for(net2) is None


class _FakeLib:
    def pgs_conv_mma_supported(self, ci, co):
        return int(ci % 16 == 0 and co % 16 == 0 and 16 <= ci <= 64 and 16 <= co <= 64)

    def pgs_conv_mma_split_supported(self, ci, co):
        return int(ci % 16 == 0 and co % 16 == 0 and ci >= 16 and co >= 16)

    def pgs_conv_tc_supported(self, ci, co):
        return int(ci % 16 == 0 and co % 16 == 0 and 16 <= co <= 192 and ci >= 16)


@pytest.mark.parametrize("K,ci,co,n,table,want", [
    (27, 16, 16, 200000, True, "mma"),     # narrow + tall: register-operand mma
    (27, 32, 48, 30000, True, "mma"),
    (27, 64, 64, 100000, True, "tc"),      # 64 output channels: tcgen05
    (27, 96, 32, 100000, True, "tc"),      # c_in above the mma bound
    (27, 64, 64, 6000, True, "tc"),        # mid-size level
    (27, 112, 112, 57, True, "split"),     # few rows: warp-item kernel
    (27, 4, 16, 200000, True, "ffma"),     # input conv: channels not a multiple of 16
    (1, 64, 16, 200000, False, "tc"),      # 1x1 shortcut (identity map)
    (1, 4, 16, 5000, False, "ffma"),
])
def test_conv_dispatch_table(monkeypatch, K, ci, co, n, table, want):
    monkeypatch.setattr(me, "CONV_IMPL", "auto")
    assert me._conv_kernel_choice(_FakeLib(), K, ci, co, n, table) == want


def test_forced_kernels_fall_back_where_unsupported(monkeypatch):
    lib = _FakeLib()
    monkeypatch.setattr(me, "CONV_IMPL", "mma")
    assert me._conv_kernel_choice(lib, 27, 16, 16, 100, True) == "mma"
    assert me._conv_kernel_choice(lib, 27, 96, 112, 9000, True) == "tc"
    monkeypatch.setattr(me, "CONV_IMPL", "ffma")
    assert me._conv_kernel_choice(lib, 27, 16, 16, 200000, True) == "ffma"


def test_native_executor_records_match_the_header():
    """The numpy mirrors of pgs_unet_op / pgs_unet_conv / pgs_unet_bn (fastpath.OP_DT / CONV_DT / BN_DT) have the
    C structs' sizes, and the host-only parts of the executor (scratch sizing, argument checks) agree with Python."""
    import ctypes
    import numpy as np
    from panopticsegforlargescalepointcloud_b200 import _lib
    lib = _lib.load()
    out = (ctypes.c_int32 * 3)()
    lib.pgs_unet_record_bytes(out)
    assert list(out) == [fastpath.OP_DT.itemsize, fastpath.CONV_DT.itemsize, fastpath.BN_DT.itemsize]
    net = bb.Minkowski("unet", input_nc=4, config=bb.paper_backbone_config(16))
    p = fastpath.Program(net)
    assert p.ops_np.dtype == fastpath.OP_DT and len(p.ops_np) == len(p.ops)
    # shapes of a made-up hierarchy: level sizes shrink by ~4x per stride-2 conv
    n, C, ts = [0] * p.n_slots, [0] * p.n_slots, [1] * p.n_slots
    n[0], C[0] = 1000, 4
    size = {1: 1000, 2: 260, 4: 70, 8: 20, 16: 7, 32: 3, 64: 1}
    for kind, a, b, dst, idx, relu in p.ops:
        ts[dst], n[dst], C[dst] = ts[a], n[a], C[a]
        if kind == fastpath.OP_CONV:
            m = p.convs[idx]
            ts[dst] = ts[a] // m.stride if m.TRANSPOSE else ts[a] * m.stride
            n[dst], C[dst] = size[ts[dst]], m.out_channels
        elif kind == fastpath.OP_CAT:
            C[dst] = C[a] + C[b]
    total = 0
    for kind, a, b, dst, idx, relu in p.ops:
        if kind in (fastpath.OP_CONV, fastpath.OP_BN):
            total += n[a] * C[a]
        elif kind == fastpath.OP_CAT:
            total += n[a] * (C[a] + C[b])
    total += sum((c - 1) * n[s] * C[s] for s, c in enumerate(p.consumers) if c > 1)
    sn, sc = np.array(n, np.int64), np.array(C, np.int32)
    got = lib.pgs_unet_backward_scratch_elems(p.ops_np.ctypes.data, len(p.ops), p.n_slots, sn.ctypes.data, sc.ctypes.data)
    assert got == total
    # argument errors are reported before anything is launched (no GPU needed)
    bad = p.ops_np.copy()
    bad["dst"][0] = p.n_slots + 5
    ptrs = np.zeros(p.n_slots, np.uint64)
    rc = lib.pgs_unet_forward(bad.ctypes.data, len(bad), p.n_slots, ptrs.ctypes.data, sn.ctypes.data, sc.ctypes.data,
                              None, None, None, None, None, None)
    assert rc != 0 and b"slot out of range" in lib.pgs_last_error()


def test_strided_maps_requested_ahead_of_the_forward_pass():
    """fastpath.build_strided_maps (used by backbone.prefetch_maps): the down path asks for each doubling once, in
    order; the transposed convolutions of the up path reuse the encoder's maps and ask for nothing."""
    class Recorder:
        def __init__(self):
            self.calls = []

        def stride(self, ts_in, ts_out):
            self.calls.append((ts_in, ts_out))

    rec = Recorder()
    fastpath.build_strided_maps(fastpath.Program(bb.Minkowski("unet", input_nc=4, config=bb.paper_backbone_config(16))), rec, 1)
    assert rec.calls == [(1, 2), (2, 4), (4, 8), (8, 16), (16, 32), (32, 64)]
    rec = Recorder()
    fastpath.build_strided_maps(fastpath.Program(bb.Minkowski("unet", input_nc=4, config=bb.two_level_config(16))), rec, 1)
    assert rec.calls == [(1, 2)]

