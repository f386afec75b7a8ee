"""SatMAERestoration -- drop-in for the reference's ``src.models.SatMAERestoration``
(src/models/satmae_restoration.py:22-313): same constructor (the YAML config dict), same parameter / buffer names
(284 state-dict entries), same helper API (count_parameters, get_optimizer_param_groups, profile_memory).

``forward`` hands the whole network to libmsd_b200 (csrc/model.cu): weights are packed once per parameter version
into bf16 [N,K] matrices with eval-mode BatchNorm folded in, activations live in a cached workspace in HBM, and
one C call enqueues the ~110 kernels of the pass on the current stream.  There is no PyTorch-op fallback: a
non-CUDA input or a missing library raises.
"""
from __future__ import annotations

import ctypes as C
import logging
import weakref
from typing import Dict

import torch
import torch.nn as nn

from .. import _lib
from .decoder import LightweightDecoder
from .encoder import SatMAEEncoder
from .export_graph import export_forward, export_graph_requested

logger = logging.getLogger(__name__)


def _bf16(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).contiguous()


def _pad_k(w2d: torch.Tensor, kpad: int) -> torch.Tensor:
    if w2d.shape[-1] == kpad:
        return w2d
    out = w2d.new_zeros(*w2d.shape[:-1], kpad)
    out[..., : w2d.shape[-1]] = w2d
    return out


def _bn_fold(bn: nn.BatchNorm2d):
    scale = bn.weight.float() / torch.sqrt(bn.running_var.float() + bn.eps)
    shift = bn.bias.float() - bn.running_mean.float() * scale
    return scale, shift


def pack_ln_folded_linear(weight: torch.Tensor, bias: torch.Tensor, ln_weight: torch.Tensor, ln_bias: torch.Tensor):
    """LayerNorm folded into the Linear behind it (encoder.py:154-176: qkv(norm1(x)), fc1(norm2(x))):
        Linear(LN(x))[n] = rstd * (sum_k x_k g_k W_nk - mean * sum_k g_k W_nk) + (b_n + sum_k beta_k W_nk)
    -> (bf16 [N,K] gamma-scaled weight, fp32 [N] column sums OF THE ROUNDED weight (a constant row then cancels exactly),
        fp32 [N] bias with the beta term).  The GEMM epilogue applies rstd / mean per row (csrc/gemm.cuh, Problem)."""
    w = weight.detach().float()
    wg = _bf16(w * ln_weight.detach().float()[None, :])
    s = wg.float().sum(dim=1).contiguous()
    b = (bias.detach().float() + w @ ln_bias.detach().float()).contiguous()
    return wg, s, b


def pack_conv3x3(conv: nn.Conv2d, bn: nn.BatchNorm2d, with_skip: bool = False):
    """[Cout,Cin,3,3] (+BN) -> bf16 [Cout, 9*Kpad] tap-major (ky,kx,ci), fp32 bias [Cout].

    with_skip: append a tenth tap holding the identity, so that the ResidualBlock's `x + block(x)` (blocks.py:85) is
    accumulated by the same tensor-core loop (A operand of that tap = the skip tensor)."""
    scale, shift = _bn_fold(bn)
    w = conv.weight.float() * scale.view(-1, 1, 1, 1)
    cout, cin = w.shape[:2]
    kpad = (cin + 63) // 64 * 64
    w = _pad_k(w.permute(0, 2, 3, 1).contiguous(), kpad).reshape(cout, 9 * kpad)
    if with_skip:
        assert cin == cout, "skip connection needs Cin == Cout"
        w = torch.cat([w, _pad_k(torch.eye(cout, device=w.device, dtype=w.dtype), kpad)], dim=1)
    bias = shift if conv.bias is None else shift + conv.bias.float() * scale
    return _bf16(w), bias.contiguous()


def pack_convt2x2(convt: nn.ConvTranspose2d, bn: nn.BatchNorm2d):
    """[Cin,Cout,2,2] (+BN) -> bf16 [4*Cout, Kpad] rows (di,dj,o), fp32 bias [4*Cout]."""
    scale, shift = _bn_fold(bn)
    w = convt.weight.float() * scale.view(1, -1, 1, 1)
    cin, cout = w.shape[:2]
    kpad = (cin + 63) // 64 * 64
    w = _pad_k(w.permute(2, 3, 1, 0).contiguous().reshape(4 * cout, cin), kpad)
    bias = (shift + convt.bias.float() * scale).repeat(4)
    return _bf16(w), bias.contiguous()


def pack_conv3x3_raw(weight: torch.Tensor, with_identity: bool = False) -> torch.Tensor:
    """[Cout,Cin,3,3] -> bf16 [Cout, 9*Kpad] tap-major, no BatchNorm fold (training step); with_identity appends the identity tap."""
    w = weight.detach().float()
    cout, cin = w.shape[:2]
    kpad = (cin + 63) // 64 * 64
    w = _pad_k(w.permute(0, 2, 3, 1).contiguous(), kpad).reshape(cout, 9 * kpad)
    if with_identity:
        assert cin == cout
        w = torch.cat([w, _pad_k(torch.eye(cout, device=w.device, dtype=w.dtype), kpad)], dim=1)
    return _bf16(w)


def pack_conv3x3_dgrad(weight: torch.Tensor, with_identity: bool = False) -> torch.Tensor:
    """Data-gradient operand of a 3x3 conv: the conv with flipped taps and swapped channel roles,
    Wd[ci, tap, co] = W[co, ci, 8 - tap]  (dX = conv3x3(dY, Wd)).  with_identity: tenth tap = identity, adds the gradient
    that reaches the same tensor through the ResidualBlock skip connection (blocks.py:85)."""
    return pack_conv3x3_raw(weight.detach().flip(2, 3).permute(1, 0, 2, 3), with_identity)


def pack_convt2x2_raw(weight: torch.Tensor, bias: torch.Tensor):
    """[Cin,Cout,2,2] -> (bf16 [4*Cout, Kpad] rows (di,dj,o), fp32 bias [4*Cout], bf16 W^T [Cin, 4*Cout])."""
    w = weight.detach().float()
    cin, cout = w.shape[:2]
    kpad = (cin + 63) // 64 * 64
    m = w.permute(2, 3, 1, 0).contiguous().reshape(4 * cout, cin)
    return _bf16(_pad_k(m, kpad)), bias.detach().float().repeat(4).contiguous(), _bf16(m.t())


class _RestoreTrainFn(torch.autograd.Function):
    """One training forward / backward of the whole network inside libmsd_b200 (csrc/model.cu, training step).  The
    parameters are passed as inputs only so that autograd routes their gradients; the arithmetic reads the packed copies."""

    @staticmethod
    def forward(ctx, model, x, *params):
        y, ws = model._forward_train(x)
        ctx.model, ctx.ws, ctx.B = model, ws, x.shape[0]
        ctx.needs = [p.requires_grad for p in params]
        return y

    @staticmethod
    def backward(ctx, grad_y):
        grads = ctx.model._backward_train(grad_y.detach().float().contiguous(), ctx.ws, ctx.B)
        names = [n for n, _ in ctx.model.named_parameters()]
        out = [grads.get(n) if need else None for n, need in zip(names, ctx.needs)]
        ctx.ws = None
        return (None, None, *out)


class SatMAERestoration(nn.Module):
    """noisy [B,13,S,S] -> SatMAE ViT-B encoder -> CNN decoder -> restored [B,13,S,S] (fp32)."""

    def __init__(self, config: Dict, block_layout: str | None = None, gemm_backend: int = 0):
        super().__init__()
        self.config = config
        enc, dec = config["model"]["encoder"], config["model"]["decoder"]
        self.patch_size = enc["patch_size"]
        self.image_size = config["data"]["image_size"]
        self.grid_size = self.image_size // self.patch_size
        self.encoder = SatMAEEncoder(
            in_channels=enc["input_channels"], image_size=self.image_size, patch_size=enc["patch_size"],
            embed_dim=enc["embed_dim"], depth=enc["depth"], num_heads=enc["num_heads"],
            pretrained_path=enc.get("pretrained_path"), freeze_layers=enc.get("freeze_layers"),
            use_checkpointing=enc["gradient_checkpointing"], block_layout=block_layout)
        self.decoder = LightweightDecoder(in_channels=enc["embed_dim"], channels=dec["channels"],
                                          out_channels=dec["output_channels"])
        self.gemm_backend = gemm_backend  # 0 = tcgen05 (product), 1 = legacy mma.sync (A/B baseline)
        self._handle = None
        self._packed = None
        self._packed_key = None
        self._workspaces: Dict[tuple, torch.Tensor] = {}
        self._log_architecture()

    # ------------------------------------------------------------------ library plumbing
    def _version_key(self, device):
        return (str(device), self.gemm_backend, self.training,
                tuple(t._version for t in list(self.parameters()) + list(self.buffers())),
                tuple(t.data_ptr() for t in self.parameters()))

    def _desc(self) -> _lib.ModelDesc:
        enc, dec = self.encoder, self.decoder
        if len(dec.channels) != 4 or dec.num_residual_blocks != 2:
            raise _lib.MsdError("libmsd_b200 builds the 4-stage / 2-residual-block decoder of the reference configs only")
        d = _lib.ModelDesc()
        d.in_channels = enc.patch_embed.proj.in_channels
        d.image_size, d.patch_size, d.embed_dim = self.image_size, self.patch_size, enc.embed_dim
        d.depth, d.num_heads = enc.depth, enc.num_heads
        d.mlp_hidden = int(enc.embed_dim * enc.mlp_ratio)
        for i, c in enumerate(dec.channels):
            d.dec_channels[i] = c
        d.out_channels = dec.out_channels
        d.ln_eps_block = enc.ln_eps_block
        d.ln_eps_outer = enc.norm.eps
        d.gemm_backend = self.gemm_backend
        return d

    @torch.no_grad()
    def _pack(self, device, encoder_only: bool = False):
        """bf16 / BN-folded copies of the parameters in the layouts include/msd.h documents.
        encoder_only: skip the BatchNorm-folded decoder (the training step packs its own un-folded decoder weights)."""
        enc, dec = self.encoder, self.decoder
        keep = []  # tensors that must outlive the C struct

        def f32(t):
            t = t.detach().float().contiguous()
            keep.append(t)
            return _lib.ptr(t)

        def b16(t):
            t = self._cached_cast(t, "b16", lambda a: _bf16(a.detach()))
            keep.append(t)
            return _lib.ptr(t)

        w = _lib.ModelWeights()
        pe = enc.patch_embed
        w.w_pe, w.b_pe = b16(pe.proj.weight.reshape(pe.proj.out_channels, -1)), f32(pe.proj.bias)
        w.pe_ln_g, w.pe_ln_b = f32(pe.norm.weight), f32(pe.norm.bias)
        w.pos_embed = f32(enc.pos_embed[0])
        blocks = (_lib.BlockWeights * enc.depth)()
        for i, blk in enumerate(enc.blocks):
            t = blk.tensors()
            bw = blocks[i]
            bw.ln1_g, bw.ln1_b = f32(t["ln1"].weight), f32(t["ln1"].bias)
            bw.ln2_g, bw.ln2_b = f32(t["ln2"].weight), f32(t["ln2"].bias)
            bw.w_qkv, bw.b_qkv = b16(t["qkv_w"]), f32(t["qkv_b"])
            bw.w_proj, bw.b_proj = b16(t["proj_w"]), f32(t["proj_b"])
            bw.w_fc1, bw.b_fc1 = b16(t["fc1"].weight), f32(t["fc1"].bias)
            bw.w_fc2, bw.b_fc2 = b16(t["fc2"].weight), f32(t["fc2"].bias)
            if not encoder_only and self.gemm_backend == 0:
                # eval forward on the tcgen05 path: norm1 / norm2 ride in the qkv / fc1 GEMMs (no per-block LayerNorm kernel)
                for wname, lnname, fields in (("qkv", "ln1", ("w_qkv_ln", "s_qkv", "b_qkv_ln")), ("fc1", "ln2", ("w_fc1_ln", "s_fc1", "b_fc1_ln"))):
                    wt, bs = (t["qkv_w"], t["qkv_b"]) if wname == "qkv" else (t["fc1"].weight, t["fc1"].bias)
                    packed = pack_ln_folded_linear(wt, bs, t[lnname].weight, t[lnname].bias)
                    keep.extend(packed)
                    for field, tensor in zip(fields, packed):
                        setattr(bw, field, _lib.ptr(tensor))
        keep.append(blocks)
        w.blocks = C.cast(blocks, C.POINTER(_lib.BlockWeights))
        w.norm_g, w.norm_b = f32(enc.norm.weight), f32(enc.norm.bias)
        if encoder_only:
            return w, keep
        mods = list(dec.upsample_blocks)
        per_stage = 1 + dec.num_residual_blocks
        for s in range(4):
            up = mods[s * per_stage]
            sw = w.stages[s]
            wt, bias = pack_convt2x2(up.upsample[0], up.upsample[1])
            keep.extend([wt, bias])
            sw.w_up, sw.b_up = _lib.ptr(wt), _lib.ptr(bias)
            k = 0
            for rb in mods[s * per_stage + 1:(s + 1) * per_stage]:
                for conv, bn, skip in ((rb.block[0], rb.block[1], False), (rb.block[3], rb.block[4], True)):
                    wt, bias = pack_conv3x3(conv, bn, with_skip=skip)
                    keep.extend([wt, bias])
                    sw.w_conv[k], sw.b_conv[k] = _lib.ptr(wt).value, _lib.ptr(bias).value
                    k += 1
        wt, bias = pack_conv3x3(dec.head[0], dec.head[1])
        keep.extend([wt, bias])
        w.w_head3, w.b_head3 = _lib.ptr(wt), _lib.ptr(bias)
        w.w_head1 = f32(dec.head[3].weight.reshape(dec.out_channels, -1))
        w.b_head1 = f32(dec.head[3].bias)
        return w, keep

    def _cached_cast(self, t: torch.Tensor, tag: str, fn):
        """Packed copy of `t`; copies of frozen parameters (requires_grad False) are made once and reused.
        An entry belongs to one parameter OBJECT (held by weak reference) at one address and version: after
        `.cpu()` / `.cuda()` the allocator may hand a parameter the address another same-shaped one had before, so the
        address alone does not identify the tensor (twelve blocks share every shape)."""
        if t.requires_grad:
            return fn(t)
        src = t._base if t._is_view() and t._base is not None else t
        cache = self.__dict__.setdefault("_frozen_cache", {})
        key = (id(src), t.data_ptr(), tuple(t.shape), tuple(t.stride()), src._version, tag)
        hit = cache.get(key)
        if hit is not None and hit[0]() is src:
            return hit[1]
        out = fn(t)
        cache[key] = (weakref.ref(src), out)
        return out

    def _apply(self, fn, recurse=True):
        """Moving / casting the module re-homes every parameter: drop the packed copies made from the old storage."""
        sig = lambda: [(t.data_ptr(), t.device, t.dtype) for t in list(self.parameters()) + list(self.buffers())]
        before = sig()
        out = super()._apply(fn, recurse)
        if sig() == before:                     # e.g. .to(device) on a model that already lives there (evaluate_model does that)
            return out
        # the C handle owns a side stream and events created on the device it first ran on: a moved model gets a fresh handle
        if self.__dict__.get("_handle") is not None:
            try:
                _lib.load().msd_model_destroy(self._handle)
            except Exception:  # pragma: no cover
                pass
            self._handle = None
        self.__dict__.pop("_frozen_cache", None)
        self._packed, self._packed_key = None, None
        for k in ("_train_packed", "_train_static_key", "_train_versions"):
            if k in self.__dict__:
                self.__dict__[k] = None
        self._workspaces.clear()
        return out

    def _ensure_ready(self, device):
        lib = _lib.load()
        _lib.require_device(device)
        if self._handle is None:
            h = C.c_void_p()
            d = self._desc()
            _lib.check(lib.msd_model_create(C.byref(d), C.byref(h)))
            self._handle = h
        key = self._version_key(device)
        if key != self._packed_key:
            w, keep = self._pack(device)
            _lib.check(lib.msd_model_set_weights(self._handle, C.byref(w)))
            self._packed, self._packed_key = (w, keep), key
        return lib

    def __del__(self):
        try:
            if self._handle is not None:
                _lib.load().msd_model_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    def launches_per_forward(self) -> int:
        self._ensure_ready(next(self.parameters()).device)
        return int(_lib.load().msd_model_launch_count(self._handle))

    # ------------------------------------------------------------------ forward
    def forward(self, x: torch.Tensor, patches: torch.Tensor | None = None) -> torch.Tensor:
        """x: noisy [B, num_bands, S, S] fp32 CUDA -> restored [B, num_bands, S, S] fp32.
        patches: optional bf16 patch-embed operand of x produced by AddSensorNoise(..., patch_out=...)."""
        B, C_, H, W = x.shape
        assert H == self.image_size and W == self.image_size, \
            f"Expected {self.image_size}x{self.image_size}, got {H}x{W}"
        assert C_ == self.config["data"]["num_bands"], \
            f"Expected {self.config['data']['num_bands']} bands, got {C_}"
        if not x.is_cuda:
            if export_graph_requested():
                # a tracer is recording (torch.onnx.export / torch.jit.trace, reference scripts/export.py:45-66) or the caller
                # opened export_mode(): the graph DEFINITION in stock torch ops (models/export_graph.py), never a product path
                return export_forward(self, x)
            raise _lib.MsdError("SatMAERestoration.forward: CUDA tensor required (libmsd_b200 has no CPU path; the torch-op export "
                                "graph is only evaluated under a tracer or msd_b200.models.export_mode())")
        if next(self.parameters()).device != x.device:
            raise _lib.MsdError("SatMAERestoration: model and input are on different devices")
        if B == 0:                              # empty batch: nothing to launch
            return x.new_zeros(0, self.decoder.out_channels, H, W, dtype=torch.float32)
        if self.training:
            # training mode: BatchNorm batch statistics + the backward pass, both inside the library (csrc/model.cu)
            if torch.is_grad_enabled():
                return _RestoreTrainFn.apply(self, x.detach().float().contiguous(), *self.parameters())
            return self._forward_train(x.detach().float().contiguous())[0]
        if torch.is_grad_enabled() and x.requires_grad:
            raise _lib.MsdError("SatMAERestoration: gradients with respect to the input image are not produced in eval mode")
        x = x.detach().float().contiguous()
        lib = self._ensure_ready(x.device)
        with torch.cuda.device(x.device):
            need = lib.msd_model_workspace_bytes(self._handle, B)
            key = (str(x.device), B)
            ws = self._workspaces.get(key)
            if ws is None or ws.numel() < need:
                self._workspaces.clear()  # one live workspace: sized for 180 GB HBM, not for many batch sizes
                ws = torch.empty(need, dtype=torch.uint8, device=x.device)
                self._workspaces[key] = ws
            y = torch.empty(B, self.decoder.out_channels, H, W, dtype=torch.float32, device=x.device)
            _lib.check(lib.msd_model_forward(self._handle, _lib.ptr(x), _lib.ptr(patches), B, _lib.ptr(y), _lib.ptr(ws),
                                             ws.numel(), _lib.stream_ptr()))
        return y

    # ------------------------------------------------------------------ training step
    def _bn_params(self, bn: nn.BatchNorm2d, keep) -> _lib.BnParams:
        b = _lib.BnParams()
        g, be = bn.weight.detach().float().contiguous(), bn.bias.detach().float().contiguous()
        keep.extend([g, be])
        b.gamma, b.beta = _lib.ptr(g), _lib.ptr(be)
        b.running_mean, b.running_var = _lib.ptr(bn.running_mean), _lib.ptr(bn.running_var)   # updated in place by the kernel
        return b

    @torch.no_grad()
    def _build_train_pack(self, device):
        """The training step's operands: un-folded bf16 weights + the transposed / tap-flipped copies the data-gradient GEMMs read
        (include/msd.h, msd_train_weights).  The destination buffers are allocated ONCE; what comes back is the C struct that
        points at them and two op tables for msd_repack_weights (csrc/optim.cu): every op (first fill) and the ops of trainable
        parameters only (re-run after each optimizer step -- one kernel launch instead of ~100 cast / permute / flip launches).
        fp32 vectors (LayerNorm / BatchNorm / biases) are passed as aliases of the parameters themselves."""
        enc, dec = self.encoder, self.decoder
        keep, ops = [], []          # ops: (RepackOp, source parameter)

        def alias(t):
            a = t.detach()
            if a.dtype != torch.float32 or not a.is_contiguous():
                raise _lib.MsdError("training step: contiguous fp32 parameters required")
            keep.append(a)
            return _lib.ptr(a)

        def packed(param, kind, shape, dtype=torch.bfloat16, d0=0, d1=0, kpad=0, flag=0):
            dst = torch.empty(shape, dtype=dtype, device=device)
            keep.append(dst)
            op = _lib.RepackOp()
            op.src, op.dst = param.data_ptr(), dst.data_ptr()
            op.kind, op.d0, op.d1, op.kpad, op.flag = kind, d0, d1, kpad, flag
            ops.append((op, param))
            return _lib.ptr(dst)

        def pad64(c):
            return (c + 63) // 64 * 64

        tw = _lib.TrainWeights()
        w = tw.enc
        pe = enc.patch_embed
        kpe = pe.proj.weight[0].numel()
        w.w_pe = packed(pe.proj.weight, _lib.REPACK_CAST, (pe.proj.out_channels, kpe), d0=pe.proj.out_channels, d1=kpe)
        w.b_pe, w.pe_ln_g, w.pe_ln_b = alias(pe.proj.bias), alias(pe.norm.weight), alias(pe.norm.bias)
        w.pos_embed = alias(enc.pos_embed[0])
        blocks = (_lib.BlockWeights * enc.depth)()
        tb = (_lib.TrainBlock * enc.depth)()
        for i, blk in enumerate(enc.blocks):
            t = blk.tensors()
            bw = blocks[i]
            bw.ln1_g, bw.ln1_b, bw.ln2_g, bw.ln2_b = alias(t["ln1"].weight), alias(t["ln1"].bias), alias(t["ln2"].weight), alias(t["ln2"].bias)
            for name, wt, bias in (("qkv", t["qkv_w"], t["qkv_b"]), ("proj", t["proj_w"], t["proj_b"]),
                                   ("fc1", t["fc1"].weight, t["fc1"].bias), ("fc2", t["fc2"].weight, t["fc2"].bias)):
                n_, k_ = wt.shape
                setattr(bw, "w_" + name, packed(wt, _lib.REPACK_CAST, (n_, k_), d0=n_, d1=k_))
                setattr(bw, "b_" + name, alias(bias))
                setattr(tb[i], "wT_" + name, packed(wt, _lib.REPACK_TRANSPOSE, (k_, n_), d0=n_, d1=k_))
            tb[i].trainable = int(any(p.requires_grad for p in blk.parameters()))
        keep.extend([blocks, tb])
        w.blocks = C.cast(blocks, C.POINTER(_lib.BlockWeights))
        tw.blocks_t = C.cast(tb, C.POINTER(_lib.TrainBlock))
        w.norm_g, w.norm_b = alias(enc.norm.weight), alias(enc.norm.bias)
        mods = list(dec.upsample_blocks)
        per_stage = 1 + dec.num_residual_blocks
        for s in range(4):
            up = mods[s * per_stage]
            ts = tw.stages[s]
            ct, ctb = up.upsample[0].weight, up.upsample[0].bias
            cin, cout = ct.shape[:2]
            ts.w_up = packed(ct, _lib.REPACK_CONVT, (4 * cout, pad64(cin)), d0=cin, d1=cout, kpad=pad64(cin))
            ts.wT_up = packed(ct, _lib.REPACK_CONVT_T, (cin, 4 * cout), d0=cin, d1=cout)
            ts.b_up = packed(ctb, _lib.REPACK_REPEAT_F32, (4 * cout,), dtype=torch.float32, d0=cout, flag=4)
            ts.bn_up = self._bn_params(up.upsample[1], keep)
            k = 0
            for rb in mods[s * per_stage + 1:(s + 1) * per_stage]:
                for conv, bn, first in ((rb.block[0], rb.block[1], True), (rb.block[3], rb.block[4], False)):
                    co, ci = conv.weight.shape[:2]
                    ts.w_conv[k] = packed(conv.weight, _lib.REPACK_CONV3, (co, 9 * pad64(ci)), d0=co, d1=ci, kpad=pad64(ci)).value
                    ts.wd_conv[k] = packed(conv.weight, _lib.REPACK_CONV3_DGRAD, (ci, (10 if first else 9) * pad64(co)), d0=co, d1=ci,
                                           kpad=pad64(co), flag=int(first)).value
                    ts.bn_conv[k] = self._bn_params(bn, keep)
                    k += 1
        h0, hbn, h3 = dec.head[0], dec.head[1], dec.head[3]
        co, ci = h0.weight.shape[:2]
        tw.w_head3 = packed(h0.weight, _lib.REPACK_CONV3, (co, 9 * pad64(ci)), d0=co, d1=ci, kpad=pad64(ci))
        tw.wd_head3 = packed(h0.weight, _lib.REPACK_CONV3_DGRAD, (ci, 9 * pad64(co)), d0=co, d1=ci, kpad=pad64(co))
        tw.b_head3 = alias(h0.bias)
        tw.bn_head = self._bn_params(hbn, keep)
        tw.w_head1, tw.b_head1 = alias(h3.weight.reshape(dec.out_channels, -1)), alias(h3.bias)
        tw.bn_eps, tw.bn_momentum = float(hbn.eps), float(hbn.momentum if hbn.momentum is not None else 0.1)
        all_ops = (_lib.RepackOp * len(ops))(*[o for o, _ in ops])
        hot = [o for o, p in ops if p.requires_grad]
        hot_ops = (_lib.RepackOp * max(len(hot), 1))(*hot)
        return tw, keep, (all_ops, len(ops)), (hot_ops, len(hot))

    def _ensure_train_ready(self, device):
        lib = _lib.load()
        _lib.require_device(device)
        if self._handle is None:
            h = C.c_void_p()
            d = self._desc()
            _lib.check(lib.msd_model_create(C.byref(d), C.byref(h)))
            self._handle = h
        # The packed copies must follow every optimizer step.  Their buffers and the C struct are built once per parameter
        # placement; afterwards one msd_repack_weights launch over the trainable parameters refreshes them whenever a parameter
        # version moved (torch optimizers bump it, FusedAdamW does so explicitly).
        params, bufs = list(self.parameters()), list(self.buffers())
        static_key = (str(device), self.gemm_backend, tuple(p.data_ptr() for p in params), tuple(p.requires_grad for p in params),
                      tuple(b.data_ptr() for b in bufs))
        versions = (tuple(p._version for p in params if p.requires_grad), tuple(p._version for p in params if not p.requires_grad))
        with torch.cuda.device(device):
            if static_key != getattr(self, "_train_static_key", None):
                tw, keep, all_ops, hot_ops = self._build_train_pack(device)
                _lib.check(lib.msd_repack_weights(all_ops[0], all_ops[1], _lib.stream_ptr()))
                _lib.check(lib.msd_model_set_train_weights(self._handle, C.byref(tw)))
                self._train_packed, self._train_static_key, self._train_versions = (tw, keep, all_ops, hot_ops), static_key, versions
            elif versions != self._train_versions:
                # frozen parameters only change through load_state_dict / manual edits: then everything is repacked
                table = self._train_packed[3] if versions[1] == self._train_versions[1] else self._train_packed[2]
                if table[1]:
                    _lib.check(lib.msd_repack_weights(table[0], table[1], _lib.stream_ptr()))
                self._train_versions = versions
        return lib

    def _forward_train(self, x: torch.Tensor):
        lib = self._ensure_train_ready(x.device)
        B = x.shape[0]
        with torch.cuda.device(x.device):
            need = lib.msd_model_train_workspace_bytes(self._handle, B)
            ws = torch.empty(need, dtype=torch.uint8, device=x.device)   # owned by the autograd node until backward ran
            y = torch.empty(B, self.decoder.out_channels, self.image_size, self.image_size, dtype=torch.float32, device=x.device)
            _lib.check(lib.msd_model_forward_train(self._handle, _lib.ptr(x), B, _lib.ptr(y), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
        for mod in self.decoder.modules():
            if isinstance(mod, nn.BatchNorm2d) and mod.num_batches_tracked is not None:
                mod.num_batches_tracked += 1
        # the kernel changed the running statistics in place: the eval-mode packing must be rebuilt from them
        self._packed_key = None
        return y, ws

    def enable_data_parallel(self, group=None, enabled: bool = True):
        """Data-parallel training (BASELINE config C4): average every gradient over the ranks of `group` INSIDE backward.  The
        library's backward writes all gradients into one flat fp32 arena ordered by when they become ready (decoder, encoder
        blocks last to first, patch embedding) and records an event per group; each group's all-reduce is enqueued on a
        communication stream behind its event, so NCCL runs under the rest of the backward, on the arena itself (no concatenation,
        no copy back).  `loss.backward()` then leaves averaged gradients in `.grad`; do not all-reduce them again."""
        self._dp = (group,) if enabled else None
        return self

    def _backward_train(self, grad_y: torch.Tensor, ws: torch.Tensor, B: int):
        """Runs msd_model_backward and returns {parameter name: fp32 gradient in the parameter's own layout}."""
        from ..training.distributed import reduce_gradient_groups
        lib = _lib.load()
        enc, dec = self.encoder, self.decoder
        dev = grad_y.device
        E, depth = enc.embed_dim, enc.depth
        # every gradient buffer is a slice of ONE arena, laid out in the order the backward completes them (gradient groups of
        # msd_model_backward: 0 = decoder, 1 + j = encoder block depth-1-j (+ encoder.norm), depth + 1 = patch embedding)
        req = []           # (group, name, shape, setter)

        def buf(group, name, setter, *shape):
            req.append((group, name, tuple(shape), setter))

        def want(param):
            return param.requires_grad

        g = _lib.ModelGrads()
        last = depth + 1
        pe = enc.patch_embed
        kpe = pe.proj.weight[0].numel()
        if want(pe.proj.weight): buf(last, "encoder.patch_embed.proj.weight", lambda p_: setattr(g, "w_pe", p_), E, kpe)
        if want(pe.proj.bias): buf(last, "encoder.patch_embed.proj.bias", lambda p_: setattr(g, "b_pe", p_), E)
        if want(pe.norm.weight):
            buf(last, "encoder.patch_embed.norm.weight", lambda p_: setattr(g, "pe_ln_g", p_), E)
            buf(last, "encoder.patch_embed.norm.bias", lambda p_: setattr(g, "pe_ln_b", p_), E)
        if want(enc.pos_embed): buf(last, "encoder.pos_embed", lambda p_: setattr(g, "pos_embed", p_), enc.pos_embed.shape[1], E)
        bg = (_lib.BlockGrads * depth)()
        names = {v: k for k, v in ((n, id(p)) for n, p in self.named_parameters())}
        for i, blk in enumerate(enc.blocks):
            if not any(p.requires_grad for p in blk.parameters()):
                continue
            t = blk.tensors()
            hid = t["fc1"].weight.shape[0]
            for field, param, shape in (
                    ("w_fc2", t["fc2"].weight, (E, hid)), ("b_fc2", t["fc2"].bias, (E,)), ("w_fc1", t["fc1"].weight, (hid, E)),
                    ("b_fc1", t["fc1"].bias, (hid,)), ("ln2_g", t["ln2"].weight, (E,)), ("ln2_b", t["ln2"].bias, (E,)),
                    ("w_proj", t["proj_w"], (E, E)), ("b_proj", t["proj_b"], (E,)), ("w_qkv", t["qkv_w"], (3 * E, E)),
                    ("b_qkv", t["qkv_b"], (3 * E,)), ("ln1_g", t["ln1"].weight, (E,)), ("ln1_b", t["ln1"].bias, (E,))):
                buf(depth - i, names[id(param)], lambda p_, i=i, field=field: setattr(bg[i], field, p_), *shape)
        g.blocks = C.cast(bg, C.POINTER(_lib.BlockGrads))
        buf(1, "encoder.norm.weight", lambda p_: setattr(g, "norm_g", p_), E)
        buf(1, "encoder.norm.bias", lambda p_: setattr(g, "norm_b", p_), E)
        mods = list(dec.upsample_blocks)
        per_stage = 1 + dec.num_residual_blocks
        post = []          # (parameter name, unpack function): packed gradient layout -> the parameter's layout
        zero_names = []    # a bias in front of BatchNorm has zero gradient
        c = dec.channels[-1]
        kpad = (c + 63) // 64 * 64
        buf(0, "decoder.head.3.weight", lambda p_: setattr(g, "w_head1", p_), dec.out_channels, c)
        buf(0, "decoder.head.3.bias", lambda p_: setattr(g, "b_head1", p_), dec.out_channels)
        buf(0, "decoder.head.1.weight", lambda p_: setattr(g, "bn_head_g", p_), c)
        buf(0, "decoder.head.1.bias", lambda p_: setattr(g, "bn_head_b", p_), c)
        buf(0, "decoder.head.0.weight", lambda p_: setattr(g, "w_head3", p_), 9, kpad, c)
        post.append(("decoder.head.0.weight", lambda t, c=c: t[:, :c, :].permute(2, 1, 0).reshape(c, c, 3, 3).contiguous()))
        zero_names.append(("decoder.head.0.bias", c))
        for s in reversed(range(4)):
            c = dec.channels[s]
            cin = E if s == 0 else dec.channels[s - 1]
            kpad = (c + 63) // 64 * 64
            sg = g.stages[s]
            k = 2 * dec.num_residual_blocks
            for j in reversed(range(dec.num_residual_blocks)):
                rp = f"decoder.upsample_blocks.{s * per_stage + 1 + j}.block."
                for ci, bi in ((3, 4), (0, 1)):
                    k -= 1
                    buf(0, rp + f"{bi}.weight", lambda p_, sg=sg, k=k: sg.bn_g.__setitem__(k, p_.value), c)
                    buf(0, rp + f"{bi}.bias", lambda p_, sg=sg, k=k: sg.bn_b.__setitem__(k, p_.value), c)
                    buf(0, rp + f"{ci}.weight", lambda p_, sg=sg, k=k: sg.w_conv.__setitem__(k, p_.value), 9, kpad, c)
                    post.append((rp + f"{ci}.weight", lambda t, c=c: t[:, :c, :].permute(2, 1, 0).reshape(c, c, 3, 3).contiguous()))
            pfx = f"decoder.upsample_blocks.{s * per_stage}.upsample."
            buf(0, pfx + "1.weight", lambda p_, sg=sg: setattr(sg, "bn_up_g", p_), c)
            buf(0, pfx + "1.bias", lambda p_, sg=sg: setattr(sg, "bn_up_b", p_), c)
            buf(0, pfx + "0.weight", lambda p_, sg=sg: setattr(sg, "w_up", p_), 4 * c, cin)
            post.append((pfx + "0.weight", lambda t, c=c, cin=cin: t.view(2, 2, c, cin).permute(3, 2, 0, 1).contiguous()))
            zero_names.append((pfx + "0.bias", c))
        # arena: groups in completion order, every slice 16-byte aligned
        req.sort(key=lambda r: r[0])
        off, spans, bounds, pads = 0, [], {}, []
        for group, name, shape, _ in req:
            n = 1
            for d_ in shape:
                n *= d_
            spans.append((off, n))
            lo, hi = bounds.get(group, (off, off))
            bounds[group] = (min(lo, off), off + n)
            if n % 4:
                pads.append((off + n, off + (n + 3) // 4 * 4))
            off += (n + 3) // 4 * 4
        arena = torch.empty(off, dtype=torch.float32, device=dev)
        for lo, hi in pads:        # alignment words between slices belong to no gradient; keep them defined for the collectives
            arena[lo:hi].zero_()
        bufs = {}
        for (group, name, shape, setter), (o, n) in zip(req, spans):
            t = arena[o:o + n].view(shape)
            bufs[name] = t
            setter(C.c_void_p(t.data_ptr()))
        with torch.cuda.device(dev):
            _lib.check(lib.msd_model_backward(self._handle, _lib.ptr(grad_y), B, _lib.ptr(ws), ws.numel(), C.byref(g), _lib.stream_ptr()))
            dp = getattr(self, "_dp", None)
            if dp is not None:
                reduce_gradient_groups(arena, [bounds[k] for k in sorted(bounds)], sorted(bounds), dp[0],
                                       lambda k, stream_ptr: _lib.check(lib.msd_model_grad_group_wait(self._handle, k, stream_ptr)))
        for name, n in zero_names:
            bufs[name] = torch.zeros(n, dtype=torch.float32, device=dev)
        for name, fn in post:
            bufs[name] = fn(bufs[name])
        out = {}
        for name, param in self.named_parameters():
            if name in bufs:
                out[name] = bufs[name].reshape(param.shape)
        return out

    # ------------------------------------------------------------------ reference helper API
    def count_parameters(self) -> Dict[str, int]:
        """Same keys as the reference (satmae_restoration.py:137-165)."""
        e = self.encoder.get_num_params()
        d_total = sum(p.numel() for p in self.decoder.parameters())
        d_train = sum(p.numel() for p in self.decoder.parameters() if p.requires_grad)
        total, trainable = e["total"] + d_total, e["trainable"] + d_train
        return {"encoder_total": e["total"], "encoder_trainable": e["trainable"], "encoder_frozen": e["frozen"],
                "decoder_total": d_total, "decoder_trainable": d_train, "total": total, "trainable": trainable,
                "frozen": total - trainable, "trainable_percent": 100 * trainable / total if total > 0 else 0}

    def _log_architecture(self):
        p = self.count_parameters()
        bands = self.config["data"]["num_bands"]
        logger.info("SatMAE restoration model (libmsd_b200): in/out [%d, %d, %d]; encoder %s params (%s trainable), "
                    "decoder %s params, total %s (%.1f%% trainable)", bands, self.image_size, self.image_size,
                    f"{p['encoder_total']:,}", f"{p['encoder_trainable']:,}", f"{p['decoder_total']:,}",
                    f"{p['total']:,}", p["trainable_percent"])

    def get_optimizer_param_groups(self, lr, weight_decay: float = 0.05):
        """[encoder params @ lr*0.1, everything else @ lr], split by 'encoder' in the name (satmae_restoration.py:195-242)."""
        if isinstance(lr, str):
            lr = float(lr)
        elif isinstance(lr, (list, tuple)):
            lr = float(lr[0])
        enc_params, dec_params = [], []
        for name, param in self.named_parameters():
            if param.requires_grad:
                (enc_params if "encoder" in name else dec_params).append(param)
        logger.info("Created optimizer groups: encoder (lr=%.2e), decoder (lr=%.2e)", lr * 0.1, lr)
        return [{"params": enc_params, "lr": lr * 0.1, "weight_decay": weight_decay},
                {"params": dec_params, "lr": lr, "weight_decay": weight_decay}]

    @torch.no_grad()
    def profile_memory(self, batch_size: int = 8) -> Dict[str, float]:
        """Peak-VRAM estimate around one forward, same keys as the reference (satmae_restoration.py:244-313)."""
        if not torch.cuda.is_available():
            logger.warning("CUDA not available, cannot profile memory")
            return {}
        device = next(self.parameters()).device
        torch.cuda.reset_peak_memory_stats(device)
        torch.cuda.empty_cache()
        model_b = sum(p.nelement() * p.element_size() for p in self.parameters())
        train_b = sum(p.nelement() * p.element_size() for p in self.parameters() if p.requires_grad)
        was_training = self.training
        self.eval()
        dummy = torch.randn(batch_size, self.config["data"]["num_bands"], self.image_size, self.image_size, device=device)
        _ = self.forward(dummy)
        self.train(was_training)
        peak = torch.cuda.max_memory_allocated(device)
        opt_b, grad_b = 2 * train_b, train_b
        act_b = max(0, peak - (model_b + opt_b + grad_b))
        return {"model_weights_gb": model_b / 1e9, "optimizer_gb": opt_b / 1e9, "activations_gb": act_b / 1e9,
                "gradients_gb": grad_b / 1e9, "peak_vram_gb": peak / 1e9, "batch_size": batch_size}
