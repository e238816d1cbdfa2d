"""Fused attention kernel (csrc/flash_attn.cu) against a plain fp64 softmax(QK^T)V on the same bf16-rounded operands."""
import math

import pytest
import torch

from helpers import rel_err


def _ref(q, k, v, causal, off):
    """q [BH, Lq, 64], k / v [BH, Lk, 64] (fp64)."""
    s = q @ k.transpose(1, 2) / math.sqrt(64)
    if causal:
        Lq, Lk = q.shape[1], k.shape[1]
        mask = torch.arange(Lk)[None, :] > (torch.arange(Lq)[:, None] + off)
        s = s.masked_fill(mask[None], float("-inf"))
    p = torch.softmax(s, dim=-1)
    return p @ v, torch.logsumexp(s, dim=-1)


@pytest.mark.gpu
@pytest.mark.parametrize("B,H,Lq,Lk,Lmax,causal", [
    (2, 3, 514, 514, 752, True),      # the prefill of the bench: 2 context frames, partial last query / key tile
    (1, 2, 75, 75, 80, True),         # single partial tile
    (3, 1, 128, 128, 128, True),      # exactly one tile
    (2, 2, 130, 390, 392, True),      # chunk appended to an existing cache (causal offset 260)
    (1, 4, 200, 333, 336, False),     # non-causal (cross-attention style), ragged key count
    (64, 12, 257, 257, 264, True),    # more work items than SMs: the persistent loop, barrier phases across items
])
def test_flash_attention_vs_fp64(cuda, B, H, Lq, Lk, Lmax, causal):
    from ivideogpt_b200 import ops
    g = torch.Generator().manual_seed(B * 1000 + Lq)
    q = (torch.randn(B, H, Lq, 64, generator=g) * 1.5).to(torch.bfloat16)
    k = torch.zeros(B, H, Lmax, 64, dtype=torch.bfloat16)
    v = torch.zeros(B, H, Lmax, 64, dtype=torch.bfloat16)
    k[:, :, :Lk] = (torch.randn(B, H, Lk, 64, generator=g) * 1.5).to(torch.bfloat16)
    v[:, :, :Lk] = torch.randn(B, H, Lk, 64, generator=g).to(torch.bfloat16)
    # poison the cache beyond Lk: the kernel must not read it
    k[:, :, Lk:] = 300.0
    v[:, :, Lk:] = -300.0
    vt = v.transpose(2, 3).contiguous()                                   # [B, H, 64, Lmax]
    out = torch.zeros(B * Lq, H * 64, dtype=torch.bfloat16, device=cuda)
    lse = torch.zeros(B * H, Lq, dtype=torch.float32, device=cuda)
    ops.flash_attn(q.to(cuda), k.to(cuda), vt.to(cuda), out, B, H, Lq, Lk, Lmax * 64, 64 * Lmax, Lmax, causal=causal,
                   scale=0.125, lse=lse)
    want, want_lse = _ref(q.double().reshape(B * H, Lq, 64), k[:, :, :Lk].double().reshape(B * H, Lk, 64),
                          v[:, :, :Lk].double().reshape(B * H, Lk, 64), causal, Lk - Lq)
    got = out.float().cpu().view(B, Lq, H, 64).permute(0, 2, 1, 3).reshape(B * H, Lq, 64)
    assert torch.isfinite(got).all()
    e = rel_err(got, want)
    assert e < 6e-3, f"flash attention rel err {e}"                       # P and the output are rounded to bf16 (2^-9 each)
    assert float((lse.cpu().double() - want_lse).abs().max()) < 2e-3
