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
from typing import Dict

import torch
import torch.nn as nn

from .. import _lib
from .decoder import LightweightDecoder
from .encoder import SatMAEEncoder

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
    def _pack(self, device):
        """bf16 / BN-folded copies of the parameters in the layouts include/msd.h documents."""
        enc, dec = self.encoder, self.decoder
        keep = []  # tensors that must outlive the C struct

        def f32(t):
            t = t.detach().float().contiguous()
            keep.append(t)
            return _lib.ptr(t)

        def b16(t):
            t = _bf16(t.detach())
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
        keep.append(blocks)
        w.blocks = C.cast(blocks, C.POINTER(_lib.BlockWeights))
        w.norm_g, w.norm_b = f32(enc.norm.weight), f32(enc.norm.bias)
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
            raise _lib.MsdError("SatMAERestoration.forward: CUDA tensor required (libmsd_b200 has no CPU path)")
        if self.training or (torch.is_grad_enabled() and x.requires_grad):
            raise _lib.MsdError("SatMAERestoration: the training forward/backward (batch-stat BatchNorm, dgrad/wgrad) is "
                                "not built yet; call .eval() and run under torch.no_grad()")
        if next(self.parameters()).device != x.device:
            raise _lib.MsdError("SatMAERestoration: model and input are on different devices")
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
