"""Inference engine: the call a user makes to segment frames on one GPU.

    engine = SegmentationEngine(model, batch=8, height=512, width=1024)      # bf16, CUDA-graph captured
    labels = engine(frames)          # frames: pinned host float32 (B,3,H,W) -> pinned host uint8 (B,H,W)

    engine = SegmentationEngine(model, 8, 512, 1024, input_dtype=torch.uint8)    # raw 8-bit frames, normalised on the device

What it adds around the nn.Module mirror:
  * the stock-PyTorch parts (EfficientNet encoder, weight-mapper trunk) are put in inference form on a private
    copy of the model: eval-mode BatchNorms folded into the preceding convolutions, parameters cast to the
    compute dtype once (instead of autocast re-casting them every call);
  * the whole forward -- stock convolutions, decoder glue and the libhsb200 kernels -- is captured in one CUDA
    graph, so a step is one graph launch (the decoder alone is ~25 small launches per frame batch);
  * host<->device transfers use pinned staging buffers and the engine's stream.
The model handed in is not modified, and outputs equal model(frames) up to rounding.
"""
from __future__ import annotations

import copy
import os

import torch
import torch.nn as nn

from . import ops


def _fold_conv_bn(conv: nn.Conv2d, bn: nn.BatchNorm2d, keep_shift: bool = False):
    """Scale of the eval-mode BatchNorm into the convolution's weights; its shift into the convolution's bias or,
    with ``keep_shift``, returned so that it can be applied together with the activation that follows."""
    scale, shift = ops.fold_bn(bn)
    w = conv.weight.detach().float() * scale.view(-1, 1, 1, 1).to(conv.weight.device)
    b = shift.to(conv.weight.device)
    if conv.bias is not None:
        b = b + conv.bias.detach().float() * scale.to(conv.weight.device)
    conv.weight = nn.Parameter(w.to(conv.weight.dtype), requires_grad=False)
    if keep_shift:
        conv.bias = None
        return b
    conv.bias = nn.Parameter(b.to(conv.weight.dtype), requires_grad=False)
    return None


def fold_static_batchnorms(model: nn.Module, fused_epilogues: bool = True) -> int:
    """Fold eval-mode BatchNorm2d into the static convolution that feeds it, in the encoder and the weight mapper.

    In the encoder's MBConv blocks / stem / head the shift is kept apart (``FoldedBatchNorm``) so that shift, swish,
    squeeze-and-excitation mean and skip add run as one channels-last pass (hsb_bias_act_nhwc_fwd) instead of
    cuDNN's separate bias kernel plus three or four elementwise kernels.  The decoder's BatchNorms belong to the
    dynamic layers and are fused inside the CUDA kernels instead."""
    from .nn.efficientnet import EfficientNet, FoldedBatchNorm, MBConvBlock
    folded = 0

    def fold(owner, conv_name, bn_name, keep_shift=False):
        nonlocal folded
        conv, bn = getattr(owner, conv_name, None), getattr(owner, bn_name, None)
        if isinstance(conv, nn.Conv2d) and isinstance(bn, nn.BatchNorm2d) and not bn.training:
            shift = _fold_conv_bn(conv, bn, keep_shift)
            setattr(owner, bn_name, FoldedBatchNorm(shift) if keep_shift else nn.Identity())
            folded += 1

    for root_name in ("backbone", "weight_mapper"):
        root = getattr(model, root_name, None)
        if root is None:
            continue
        for m in root.modules():
            if isinstance(m, EfficientNet):
                fold(m, "_conv_stem", "_bn0", fused_epilogues)
                fold(m, "_conv_head", "_bn1", fused_epilogues)
            elif isinstance(m, MBConvBlock):
                fold(m, "_expand_conv", "_bn0", fused_epilogues)
                fold(m, "_depthwise_conv", "_bn1", fused_epilogues)
                fold(m, "_project_conv", "_bn2", fused_epilogues)
            elif isinstance(m, nn.Sequential):
                names = [n for n, _ in m.named_children()]
                for a, b in zip(names, names[1:]):
                    fold(m, a, b)
    return folded


class SegmentationEngine:
    def __init__(self, model: nn.Module, batch: int, height: int, width: int, device="cuda",
                 dtype=torch.bfloat16, use_graph: bool = True, fold_bn: bool = True, warmup: int = 3,
                 channels_last: bool = True, input_dtype=torch.float32,
                 mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)):
        """``input_dtype=torch.float32``: frames arrive normalised, as the reference's transforms leave them
        (hyperseg/test.py: ToTensor + Normalize on the host).  ``torch.uint8``: frames arrive as raw 0..255 RGB and
        ``(x / 255 - mean) / std`` runs on the device in front of the encoder -- a quarter of the upload."""
        if input_dtype not in (torch.float32, torch.uint8):
            raise ValueError("input_dtype must be torch.float32 or torch.uint8")
        self.input_dtype = input_dtype
        self.device = torch.device(device)
        self.dtype = dtype
        self.shape = (batch, 3, height, width)
        net = copy.deepcopy(model).eval()
        for p in net.parameters():
            p.requires_grad_(False)
        self.folded = fold_static_batchnorms(net) if fold_bn else 0
        # the decoder's BatchNorms are fused into the CUDA kernels: fold them once, in fp32, before the cast
        pinned = []
        for m in net.decoder.modules():
            if isinstance(m, nn.BatchNorm2d) and m.running_mean is not None:
                pinned.append((m, ops.fold_bn(m)))
        self.net = net.to(self.device, dtype)
        self.channels_last = channels_last
        if channels_last:
            # cuDNN's bf16 kernels want NHWC; the decoder kernels want NCHW, so the (small) feature maps the
            # encoder hands over are converted back where the decoder concatenates them
            self.net.backbone.to(memory_format=torch.channels_last)
            self.net.weight_mapper.to(memory_format=torch.channels_last)
        for m, (scale, shift) in pinned:
            m._hsb_folded = (scale.to(self.device, torch.float32).contiguous(),
                             shift.to(self.device, torch.float32).contiguous())
        if channels_last:           # the encoder epilogue kernels walk channels-last rows
            from .nn.efficientnet import FoldedBatchNorm
            for m in self.net.modules():
                if isinstance(m, FoldedBatchNorm):
                    m.shift32 = m.shift_master.to(self.device, torch.float32).contiguous()
        self.stream = torch.cuda.Stream(self.device)
        self.frames_dev = torch.zeros(self.shape, device=self.device, dtype=input_dtype)
        std_t = torch.tensor(std, dtype=torch.float32, device=self.device).view(1, 3, 1, 1)
        self._in_scale = 1.0 / (255.0 * std_t)
        self._in_bias = -torch.tensor(mean, dtype=torch.float32, device=self.device).view(1, 3, 1, 1) / std_t
        self.host_out = torch.empty((batch, height, width), dtype=torch.uint8).pin_memory()
        self.logits = None
        self.labels = None
        self.graph = None
        self.launches_per_step = 0
        with torch.cuda.stream(self.stream), torch.no_grad():
            for _ in range(max(1, warmup)):
                self._forward_static()
            self.stream.synchronize()
            before = ops.launch_count()
            self._forward_static()
            self.launches_per_step = ops.launch_count() - before
            if use_graph:
                self.stream.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=self.stream):
                    self._forward_static()
                self.graph = graph
        self.stream.synchronize()

    def _forward_static(self):
        net = self.net
        # one pass: fp32 NCHW frames -> compute dtype, channels-last (the decoder's glue kernel takes any strides)
        fmt = torch.channels_last if self.channels_last else torch.contiguous_format
        if self.input_dtype == torch.uint8:       # normalise in fp32 (one rounding into the compute dtype, as on the host path)
            x = torch.addcmul(self._in_bias, self.frames_dev.float(), self._in_scale).to(self.dtype, memory_format=fmt)
        else:
            x = self.frames_dev.to(self.dtype, memory_format=fmt)
        xin = x
        feats = net.backbone(xin)                  # NHWC feature maps are consumed as they are by the glue kernel
        signal = net.weight_mapper(feats[-1]).contiguous()      # NCHW once: every head reads position-contiguous rows
        self.logits = net.decoder.forward_features([x] + feats[:-1], signal)      # at the last decoder level's size
        # final bilinear upsample + argmax fused: full-resolution logits are never written
        self.labels = ops.upsample_argmax(self.logits, self.shape[-2:])

    def _check_frames(self, frames):
        if tuple(frames.shape) != self.shape:
            raise ValueError(f"engine was built for frames of shape {self.shape}, got {tuple(frames.shape)}")
        if frames.dtype != self.input_dtype:
            raise ValueError(f"engine was built for {self.input_dtype} frames, got {frames.dtype}")

    @torch.no_grad()
    def step(self):
        """One forward over the frames currently in ``frames_dev`` (enqueued on the engine's stream)."""
        with torch.cuda.stream(self.stream):
            if self.graph is not None:
                self.graph.replay()
            else:
                self._forward_static()

    @torch.no_grad()
    def __call__(self, frames: torch.Tensor) -> torch.Tensor:
        """Host frames (pinned, ``input_dtype``, B x 3 x H x W) -> host labels (pinned uint8, B x H x W)."""
        self._check_frames(frames)
        with torch.cuda.stream(self.stream):
            self.frames_dev.copy_(frames, non_blocking=True)
            if self.graph is not None:
                self.graph.replay()
            else:
                self._forward_static()
            self.host_out.copy_(self.labels, non_blocking=True)
        self.stream.synchronize()
        return self.host_out

    # ---- pipelined host -> host stream: the upload of batch k+1 overlaps the forward of batch k -------------------------
    def _pipeline(self):
        if getattr(self, "_pipe", None) is None:
            self._pipe = dict(
                copy_stream=torch.cuda.Stream(self.device),
                staging=[torch.empty(self.shape, device=self.device, dtype=self.input_dtype) for _ in range(2)],
                host_out=[torch.empty(self.host_out.shape, dtype=torch.uint8).pin_memory() for _ in range(2)],
                copied=[torch.cuda.Event() for _ in range(2)], consumed=[torch.cuda.Event() for _ in range(2)],
                done=[torch.cuda.Event() for _ in range(2)], submitted=0, collected=0)
        return self._pipe

    @torch.no_grad()
    def submit(self, frames: torch.Tensor) -> None:
        """Enqueue one batch of host frames (pinned, ``input_dtype``); at most two batches may be in flight.  The H2D copy
        runs on its own stream into a staging buffer, so it overlaps the forward of the batch submitted before."""
        self._check_frames(frames)
        p = self._pipeline()
        if p["submitted"] - p["collected"] >= 2:
            raise RuntimeError("two batches are already in flight: collect() one first")
        i = p["submitted"] % 2
        with torch.cuda.stream(p["copy_stream"]):
            p["copy_stream"].wait_event(p["consumed"][i])          # the forward two batches ago has read this buffer
            p["staging"][i].copy_(frames, non_blocking=True)
            p["copied"][i].record(p["copy_stream"])
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(p["copied"][i])
            self.frames_dev.copy_(p["staging"][i], non_blocking=True)
            p["consumed"][i].record(self.stream)
            if self.graph is not None:
                self.graph.replay()
            else:
                self._forward_static()
            p["host_out"][i].copy_(self.labels, non_blocking=True)
            p["done"][i].record(self.stream)
        p["submitted"] += 1

    def collect(self) -> torch.Tensor:
        """Host labels (pinned uint8, B x H x W) of the oldest batch in flight; valid until two more are submitted."""
        p = self._pipeline()
        if p["collected"] >= p["submitted"]:
            raise RuntimeError("nothing in flight")
        j = p["collected"] % 2
        p["done"][j].synchronize()
        p["collected"] += 1
        return p["host_out"][j]

    @torch.no_grad()
    def full_logits(self) -> torch.Tensor:
        """Logits of the last step at frame resolution (what model(frames) returns), computed on demand."""
        import torch.nn.functional as F
        with torch.cuda.stream(self.stream):
            out = self.logits
            if tuple(out.shape[-2:]) != tuple(self.shape[-2:]):
                out = F.interpolate(out, self.shape[-2:], mode='bilinear', align_corners=False)
        self.stream.synchronize()
        return out
