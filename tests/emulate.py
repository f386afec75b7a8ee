"""CPU emulation of the GEMM 'Problem' the CUDA kernels execute (csrc/gemm.cuh), used by the host-logic tests to
prove that the weight packing, the tap-offset table over halo-padded NHWC rows and the ConvTranspose pixel-shuffle
scatter reproduce F.conv2d / F.conv_transpose2d.  fp32 throughout (no bf16 rounding) -- this checks indexing."""
from __future__ import annotations

import torch


def halo_nhwc(x_nchw: torch.Tensor) -> torch.Tensor:
    """[B,C,H,W] -> [B,H+2,W+2,C] with a zero halo."""
    B, C, H, W = x_nchw.shape
    out = x_nchw.new_zeros(B, H + 2, W + 2, C)
    out[:, 1:-1, 1:-1, :] = x_nchw.permute(0, 2, 3, 1)
    return out


def unhalo_nchw(x: torch.Tensor) -> torch.Tensor:
    return x[:, 1:-1, 1:-1, :].permute(0, 3, 1, 2).contiguous()


def conv3x3_rows(x_halo: torch.Tensor, w_packed: torch.Tensor, bias: torch.Tensor, H: int, W: int, skip=None, act=True):
    """x_halo [B,H+2,W+2,Cin]; w_packed [Cout, 9*Kpad]; returns [B,H+2,W+2,Cout] with a zero halo (EPI_CONV)."""
    B, _, _, cin = x_halo.shape
    cout = w_packed.shape[0]
    kpad = (cin + 63) // 64 * 64
    ntaps = w_packed.shape[1] // kpad          # 10 when the identity skip tap is packed
    rows = x_halo.reshape(-1, cin)
    M = rows.shape[0]
    acc = torch.zeros(M, cout)
    for ky in range(3):
        for kx in range(3):
            off = (ky - 1) * (W + 2) + (kx - 1)
            idx = torch.arange(M) + off
            valid = (idx >= 0) & (idx < M)                       # TMA out-of-bounds rows read as zero
            a = torch.zeros(M, cin)
            a[valid] = rows[idx[valid]]
            wt = w_packed[:, (ky * 3 + kx) * kpad:(ky * 3 + kx) * kpad + cin]
            acc += a @ wt.t()
    acc = acc + bias
    p = torch.arange(M)
    rem = p % ((H + 2) * (W + 2))
    yy, xx = rem // (W + 2), rem % (W + 2)
    interior = (yy >= 1) & (yy <= H) & (xx >= 1) & (xx <= W)
    if skip is not None:
        if ntaps == 10:   # the skip connection is the tenth tap: A = skip rows (shift 0), W = identity block
            acc = acc + skip.reshape(-1, cout)[:, :cin] @ w_packed[:, 9 * kpad:9 * kpad + cin].t()
        else:
            acc = acc + skip.reshape(-1, cout)
    if act:
        acc = torch.nn.functional.gelu(acc)
    acc = torch.where(interior[:, None], acc, torch.zeros_like(acc))
    return acc.reshape(B, H + 2, W + 2, cout)


def convt2x2_rows(x: torch.Tensor, w_packed: torch.Tensor, bias: torch.Tensor, Hin: int, Win: int, in_padded: bool):
    """x [B,Hin(+2),Win(+2),Cin]; w_packed [4*Cout,Kpad] rows (di,dj,o); returns [B,2Hin+2,2Win+2,Cout] (EPI_CONVT)."""
    B, cin = x.shape[0], x.shape[-1]
    cout = w_packed.shape[0] // 4
    rows = x.reshape(-1, cin)
    acc = torch.nn.functional.gelu(rows @ w_packed[:, :cin].t() + bias)
    ph, pw = (Hin + 2, Win + 2) if in_padded else (Hin, Win)
    out = torch.zeros(B * (2 * Hin + 2) * (2 * Win + 2), cout)
    out_plane, out_pitch = (2 * Hin + 2) * (2 * Win + 2), 2 * Win + 2
    m = torch.arange(rows.shape[0])
    img, rem = m // (ph * pw), m % (ph * pw)
    yy, xx = rem // pw, rem % pw
    if in_padded:
        interior = (yy >= 1) & (yy <= Hin) & (xx >= 1) & (xx <= Win)
        yy, xx = yy - 1, xx - 1
    else:
        interior = torch.ones_like(m, dtype=torch.bool)
    for q in range(4):
        opix = img * out_plane + (2 * yy + (q >> 1) + 1) * out_pitch + (2 * xx + (q & 1) + 1)
        out[opix[interior]] = acc[interior][:, q * cout:(q + 1) * cout]
    return out.reshape(B, 2 * Hin + 2, 2 * Win + 2, cout)
