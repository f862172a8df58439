"""Sparse ResUNet backbone assembled from the ME-shaped modules (`me.py`).

Host-side mirror of
  torch_points3d/modules/MinkowskiEngine/api_modules.py:9-82   ResBlock
  torch_points3d/modules/MinkowskiEngine/api_modules.py:235-311 ResNetDown / ResNetUp
  torch_points3d/applications/minkowski.py:25-126,160-196      Minkowski(...) factory, MinkowskiUnet/Encoder
  torch_points3d/models/base_architectures/unet.py:400-487     compact config -> module lists
  torch_points3d/utils/model_building_utils/model_definition_resolver.py:29-58  string expressions in YAML

Module and parameter names follow the reference so that state_dict keys are interchangeable
(`down_modules.{i}.conv_in.0.kernel`, `...blocks.{j}.block.{0,3}.kernel`, `...downsample.0.kernel`,
`....bn.{weight,bias,running_mean,running_var,num_batches_tracked}`; SURVEY section 5, checkpoint row).
When the reference's own python files are available they can be used unchanged on top of `me`
(INTEGRATION.md); this file exists so that the path runs without torch_points3d's other dependencies.
"""
import copy
import sys

import torch
import torch.nn as nn

from . import me as ME
from . import fastpath


class ResBlock(ME.MinkowskiNetwork):
    """conv3-BN-ReLU-conv3-BN-ReLU plus identity / (conv1-BN) shortcut, no ReLU after the sum."""

    def __init__(self, input_nc, output_nc, convolution, dimension=3):
        super().__init__(dimension)
        self.block = nn.Sequential(
            convolution(in_channels=input_nc, out_channels=output_nc, kernel_size=3, stride=1, dilation=1,
                        bias=False, dimension=dimension),
            ME.MinkowskiBatchNorm(output_nc),
            ME.MinkowskiReLU(),
            convolution(in_channels=output_nc, out_channels=output_nc, kernel_size=3, stride=1, dilation=1,
                        bias=False, dimension=dimension),
            ME.MinkowskiBatchNorm(output_nc),
            ME.MinkowskiReLU(),
        )
        if input_nc != output_nc:
            self.downsample = nn.Sequential(
                convolution(in_channels=input_nc, out_channels=output_nc, kernel_size=1, stride=1, dilation=1,
                            bias=False, dimension=dimension),
                ME.MinkowskiBatchNorm(output_nc),
            )
        else:
            self.downsample = None

    def forward(self, x):
        out = self.block(x)
        return out + (self.downsample(x) if self.downsample is not None else x)


_BLOCKS = {"ResBlock": ResBlock}


class ResNetDown(ME.MinkowskiNetwork):
    """strided conv_in (+BN+ReLU) followed by N residual blocks."""

    CONVOLUTION = ME.MinkowskiConvolution

    def __init__(self, down_conv_nn=(), kernel_size=2, dilation=1, dimension=3, stride=2, N=1, block="ResBlock",
                 **kwargs):
        super().__init__(dimension)
        if block not in _BLOCKS:
            raise NotImplementedError("block %r (only ResBlock is used by the shipped panoptic configs)" % block)
        block_cls = _BLOCKS[block]
        conv1_output = down_conv_nn[0] if stride > 1 else down_conv_nn[1]
        self.conv_in = nn.Sequential(
            self.CONVOLUTION(in_channels=down_conv_nn[0], out_channels=conv1_output, kernel_size=kernel_size,
                             stride=stride, dilation=dilation, bias=False, dimension=dimension),
            ME.MinkowskiBatchNorm(conv1_output),
            ME.MinkowskiReLU(),
        )
        if N > 0:
            blocks = []
            for _ in range(N):
                blocks.append(block_cls(conv1_output, down_conv_nn[1], self.CONVOLUTION, dimension=dimension))
                conv1_output = down_conv_nn[1]
            self.blocks = nn.Sequential(*blocks)
        else:
            self.blocks = None

    def forward(self, x):
        out = self.conv_in(x)
        if self.blocks is not None:
            out = self.blocks(out)
        return out


class ResNetUp(ResNetDown):
    """Decoder twin: every convolution (also the stride-1 ones) is a transposed convolution."""

    CONVOLUTION = ME.MinkowskiConvolutionTranspose

    def __init__(self, up_conv_nn=(), kernel_size=2, dilation=1, dimension=3, stride=2, N=1, **kwargs):
        super().__init__(down_conv_nn=up_conv_nn, kernel_size=kernel_size, dilation=dilation, dimension=dimension,
                         stride=stride, N=N, **kwargs)

    def forward(self, x, skip):
        inp = ME.cat(x, skip) if skip is not None else x
        return super().forward(inp)


# --------------------------------------------------------------------------------------------
# config handling
# --------------------------------------------------------------------------------------------
def _is_list(v):
    return isinstance(v, (list, tuple)) or type(v).__name__ == "ListConfig"


def _items(cfg):
    return cfg.items() if hasattr(cfg, "items") else vars(cfg).items()


def resolve(obj, constants):
    """Evaluate string expressions ("2*in_feat", "1.5 * 0.12", "None") against the constants, recursively."""
    if isinstance(obj, str):
        try:
            return eval(obj, {"__builtins__": {}}, dict(constants))
        except Exception:
            return obj
    if _is_list(obj):
        return [resolve(v, constants) for v in obj]
    if hasattr(obj, "items"):
        return {k: resolve(v, constants) for k, v in obj.items()}
    return obj


def _fetch_arguments(opt, index):
    """One conv's kwargs from the compact dict-of-lists format (unet.py:450-468 semantics)."""
    special = ("radius", "max_num_neighbors", "block_names")
    args = {}
    for name, v in _items(opt):
        name = str(name)
        if _is_list(v) and len(v) > 0:
            if name.endswith("s") and name not in special:
                name = name[:-1]
            v_index = v[index]
            args[name] = list(v_index) if _is_list(v_index) else v_index
        else:
            args[name] = list(v) if _is_list(v) else v
    return args


def extract_output_nc(cfg):
    up = cfg.get("up_conv") if hasattr(cfg, "get") else None
    if up is not None:
        nn_ = up.get("up_conv_nn")
        return nn_[-1][-1]
    inner = cfg.get("innermost") if hasattr(cfg, "get") else None
    if inner is not None:
        return inner["nn"][-1]
    return cfg["down_conv"]["down_conv_nn"][-1][-1]


class BaseMinkowski(nn.Module):
    CONV_TYPE = "sparse"

    def __init__(self, model_config, input_nc):
        super().__init__()
        constants = {"FEAT": max(int(input_nc), 0)}
        dc = model_config.get("define_constants") if hasattr(model_config, "get") else None
        if dc:
            constants.update(resolve(dict(_items(dc)), constants))
        cfg = resolve(copy.deepcopy(dict(_items(model_config))), constants)
        self.down_modules = nn.ModuleList()
        self.inner_modules = nn.ModuleList()
        self.up_modules = nn.ModuleList()
        lib = sys.modules[__name__]
        down = cfg["down_conv"]
        down_cls = getattr(lib, down["module_name"])
        for i in range(len(down["down_conv_nn"])):
            args = _fetch_arguments(down, i)
            args.pop("module_name", None)
            self.down_modules.append(down_cls(**args))
        up = cfg.get("up_conv")
        if up is not None:
            up_cls = getattr(lib, up["module_name"])
            for i in range(len(up["up_conv_nn"])):
                args = _fetch_arguments(up, i)
                args.pop("module_name", None)
                self.up_modules.append(up_cls(**args))
        if cfg.get("innermost") is not None:
            raise NotImplementedError("innermost modules (GlobalBaseModule) are not on the hot path")
        self._output_nc = extract_output_nc(cfg)
        self.weight_initialization()

    @property
    def output_nc(self):
        return self._output_nc

    @property
    def device(self):
        return next(self.parameters()).device

    def weight_initialization(self):
        # applications/minkowski.py:104-111: kaiming only on MinkowskiConvolution (not ...Transpose)
        for m in self.modules():
            if isinstance(m, ME.MinkowskiConvolution):
                ME.utils.kaiming_normal_(m.kernel, mode="fan_out", nonlinearity="relu")
            if isinstance(m, ME.MinkowskiBatchNorm):
                nn.init.constant_(m.bn.weight, 1)
                nn.init.constant_(m.bn.bias, 0)

    def _set_input(self, data):
        pm = getattr(data, "coordinate_manager", None)
        if isinstance(pm, PrebuiltMaps) and pm.n == data.x.shape[0]:
            self.input = ME.SparseTensor(features=data.x.to(self.device), coordinate_manager=pm.adopt(), tensor_stride=1)
        else:
            coords = torch.cat([data.batch.unsqueeze(-1).int(), data.coords.int()], -1)
            self.input = ME.SparseTensor(features=data.x, coordinates=coords, device=self.device)
        self.xyz = (data.pos if getattr(data, "pos", None) is not None else data.coords).to(self.device)

    def prefetch_maps(self, batch, coords, stream=None, wait_event=None):
        """Coordinate manager of the NEXT batch -- level-0 hash and every strided map this network will ask for, i.e.
        the part of the forward pass that reads sizes back to the host -- built ahead of time on `stream` (default: the
        current one), e.g. while the backward pass of the current batch runs: the read-backs then drain only `stream`.
        `batch` int [N] and `coords` int [N,3] must be device tensors that are ready on `stream` (or after
        `wait_event`).  Attach the returned handle to the batch as `coordinate_manager`; `forward` adopts it."""
        cur = torch.cuda.current_stream(batch.device)
        stream = stream or cur
        if wait_event is not None:
            stream.wait_event(wait_event)
        with torch.cuda.stream(stream):
            c4 = torch.cat([batch.unsqueeze(-1).int(), coords.int()], -1)
            cm = ME.CoordinateManager(c4)
            prog = fastpath.program_for(self) if fastpath.ENABLED else None
            if prog is not None:
                fastpath.build_strided_maps(prog, cm, 1)
            ev = torch.cuda.Event()
            ev.record(stream)
        return PrebuiltMaps(cm, ev, stream, c4.shape[0])


class PrebuiltMaps:
    """Handle returned by `prefetch_maps`: the manager, the event that marks it complete and the stream that built it."""

    def __init__(self, manager, event, stream, n):
        self.manager, self.event, self.stream, self.n = manager, event, stream, n

    def adopt(self):
        """Make the current stream wait for the build and keep the caching allocator from recycling the manager's
        buffers (allocated on the prefetch stream) while kernels of the current stream may still read them."""
        cm = self.manager
        cur = torch.cuda.current_stream(cm.device)
        if self.stream != cur:
            cur.wait_event(self.event)
            for m in cm.maps.values():
                for t in (m.coords, m.tkeys, m.tvals):
                    t.record_stream(cur)
            for t in cm.in2out.values():
                t.record_stream(cur)
        return cm


class Output:
    """Stand-in for torch_geometric.data.Data(x=, pos=, batch=) returned by the reference backbone."""

    def __init__(self, x, pos=None, batch=None):
        self.x, self.pos, self.batch = x, pos, batch


class MinkowskiUnet(BaseMinkowski):
    def forward(self, data, *args, **kwargs):
        self._set_input(data)
        x = fastpath.run(self, self.input)   # whole backbone as one autograd node (same kernels, no per-layer host cost)
        if x is not None:
            return Output(x=x.F, pos=self.xyz, batch=x.C[:, 0])
        x = self.input
        stack_down = []
        for i in range(len(self.down_modules) - 1):
            x = self.down_modules[i](x)
            stack_down.append(x)
        x = self.down_modules[-1](x)
        stack_down.append(None)
        for i in range(len(self.up_modules)):
            x = self.up_modules[i](x, stack_down.pop())
        return Output(x=x.F, pos=self.xyz, batch=x.C[:, 0])


class MinkowskiEncoder(BaseMinkowski):
    def forward(self, data, *args, **kwargs):
        self._set_input(data)
        x = fastpath.run(self, self.input)
        if x is None:
            x = self.input
            for m in self.down_modules:
                x = m(x)
        return Output(x=x.F, batch=x.C[:, 0].long())


def Minkowski(architecture="unet", input_nc=None, num_layers=None, config=None, *args, **kwargs):
    """Factory with the reference's signature (applications/minkowski.py:25-54); `config` is required
    (the default applications/conf/sparseconv3d/*.yaml files are not shipped here)."""
    if config is None:
        raise ValueError("a backbone config (compact down_conv/up_conv format) is required")
    arch = (architecture or "unet").lower()
    if arch == "unet":
        return MinkowskiUnet(config, input_nc)
    if arch == "encoder":
        return MinkowskiEncoder(config, input_nc)
    raise NotImplementedError("architecture %r" % architecture)


def paper_backbone_config(in_feat=16):
    """conf/models/panoptic/area4_ablation_3heads_5.yaml:88-126 (7-level ResUNet, 82 sparse convs)."""
    return {
        "define_constants": {"in_feat": in_feat},
        "down_conv": {
            "module_name": "ResNetDown", "dimension": 3,
            "down_conv_nn": [["FEAT", "in_feat"], ["in_feat", "2*in_feat"], ["2*in_feat", "3*in_feat"],
                             ["3*in_feat", "4*in_feat"], ["4*in_feat", "5*in_feat"], ["5*in_feat", "6*in_feat"],
                             ["6*in_feat", "7*in_feat"]],
            "kernel_size": 3, "stride": [1, 2, 2, 2, 2, 2, 2], "N": 2,
        },
        "up_conv": {
            "module_name": "ResNetUp", "dimension": 3,
            "up_conv_nn": [["7*in_feat", "6*in_feat"], ["2*6*in_feat", "5*in_feat"], ["2*5*in_feat", "4*in_feat"],
                           ["2*4*in_feat", "3*in_feat"], ["2*3*in_feat", "2*in_feat"], ["2*2*in_feat", "in_feat"],
                           ["2*in_feat", "in_feat"]],
            "kernel_size": 3, "stride": [2, 2, 2, 2, 2, 2, 1], "N": 2,
        },
    }


def two_level_config(in_feat=16):
    """BASELINE config C1: 2-level U-Net in the PointGroup schema."""
    return {
        "define_constants": {"in_feat": in_feat},
        "down_conv": {"module_name": "ResNetDown", "dimension": 3,
                      "down_conv_nn": [["FEAT", "in_feat"], ["in_feat", "2*in_feat"]],
                      "kernel_size": 3, "stride": [1, 2], "N": 1},
        "up_conv": {"module_name": "ResNetUp", "dimension": 3,
                    "up_conv_nn": [["2*in_feat", "in_feat"], ["2*in_feat", "in_feat"]],
                    "kernel_size": 3, "stride": [2, 1], "N": 1},
    }


def scorer_unet_config(in_feat=16):
    """conf/models/panoptic/area4_ablation_3heads_5.yaml:128-146."""
    return {
        "define_constants": {"in_feat": in_feat},
        "down_conv": {"module_name": "ResNetDown", "dimension": 3,
                      "down_conv_nn": [["in_feat", "2*in_feat"], ["2*in_feat", "4*in_feat"]],
                      "kernel_size": 3, "stride": 2, "N": 2},
        "up_conv": {"module_name": "ResNetUp", "dimension": 3,
                    "up_conv_nn": [["4*in_feat", "2*in_feat"], ["4*in_feat", "in_feat"]],
                    "kernel_size": 3, "stride": 2, "N": 2},
    }
