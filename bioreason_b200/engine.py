"""Kernel-level orchestration of the hot path: NT-v2 encoder forward, projector+scatter, Qwen3 decoder forward.

Python here only sequences C-ABI kernel launches on the current CUDA stream (no tensor math in torch on the
product path beyond index bookkeeping on tiny int tensors).  Reference call sites: dna_llm.py:103-179 (encode,
project, regroup), :208-244 (merge + LLM forward).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional

import torch

from . import ops
from .packing import DecoderW, EncoderW


# ---------------------------------------------------------------------------------------------
# attention-window bookkeeping (the reference's 0/1 masks are always one contiguous run per row)
# ---------------------------------------------------------------------------------------------
def mask_window(attention_mask: torch.Tensor):
    """[B, L] 0/1 mask -> (kv_start[B], kv_end[B]) int32 on the same device, no host sync."""
    m = attention_mask != 0
    L = m.shape[1]
    idx = torch.arange(L, device=m.device)
    start = torch.where(m, idx, L).amin(dim=1)
    end = torch.where(m, idx + 1, 0).amax(dim=1)
    start = torch.minimum(start, end)
    return start.to(torch.int32), end.to(torch.int32)


def forward_positions(B: int, L: int, device) -> torch.Tensor:
    """DNALLMModel.forward passes no position_ids -> HF uses arange over the PADDED row."""
    return torch.arange(L, device=device, dtype=torch.int32).repeat(B)


def generate_positions(attention_mask: torch.Tensor) -> torch.Tensor:
    """HF generate(): position_ids = cumsum(mask) - 1, pads clamped to 1 (generation/utils.py:719-721)."""
    pos = attention_mask.long().cumsum(-1) - 1
    pos = pos.masked_fill(attention_mask == 0, 1)
    return pos.to(torch.int32).reshape(-1)


# ---------------------------------------------------------------------------------------------
# LoRA adapters in kernel layout (packed from the fp32 master parameters each optimizer step)
# ---------------------------------------------------------------------------------------------
@dataclass
class LoraLayerW:
    a_qkv: torch.Tensor      # [3r, d]      rows = A_q | A_k | A_v
    b_qkv: torch.Tensor      # [(Hq+2Hkv)D, 3r]  block diagonal
    a_o: torch.Tensor        # [r, HqD]
    b_o: torch.Tensor        # [d, r]
    a_gu: torch.Tensor       # [2r, d]      rows = A_gate | A_up
    b_gu: torch.Tensor       # [2F, 2r]     row 2j = (B_gate[j], 0), row 2j+1 = (0, B_up[j])
    a_down: torch.Tensor     # [r, F]
    b_down: torch.Tensor     # [d, r]


@dataclass
class LoraW:
    r: int
    scale: float             # alpha / r, applied in the A-GEMM epilogue
    layers: List[LoraLayerW]


@dataclass
class LayerSaved:
    """Activations one decoder layer keeps for the hand-written backward."""
    h_in: torch.Tensor = None
    rstd1: torch.Tensor = None
    xn1: torch.Tensor = None
    q: torch.Tensor = None            # post qk-norm + RoPE (views of one [M, (Hq+Hkv)D] buffer)
    k: torch.Tensor = None
    v: torch.Tensor = None            # view of the QKV GEMM output
    qkv_pre: torch.Tensor = None      # the QKV GEMM output itself: pre-norm q/k (for the qk-norm backward) | v
    attn: torch.Tensor = None
    lse: torch.Tensor = None
    h_mid: torch.Tensor = None
    rstd2: torch.Tensor = None
    xn2: torch.Tensor = None
    gu: torch.Tensor = None
    act: torch.Tensor = None
    t_qkv: torch.Tensor = None
    t_o: torch.Tensor = None
    t_gu: torch.Tensor = None
    t_down: torch.Tensor = None


_ROPE_TABLES = {}


def _rope_table(n_pos: int, D: int, theta: float, device):
    """(cos, sin) pairs of positions [0, n_pos), bf16-rounded like HF's rotary tables; cached per (device, D, theta), grown on demand."""
    key = (str(device), D, float(theta))
    t = _ROPE_TABLES.get(key)
    if t is None or t.shape[0] < n_pos:
        t = _ROPE_TABLES[key] = ops.rope_table(max(n_pos, 4096), D, theta, device)
    return t


def _lin(x, w, *, lora_a=None, lora_b=None, lora_scale=1.0, saved_t=None, **kw):
    """y = x @ w.T (+ (scale * x @ A.T) @ B.T as a second K segment of the same tcgen05 accumulation)."""
    if lora_a is None:
        return ops.gemm(x, w, **kw), None
    t = ops.gemm(x, lora_a, alpha=lora_scale)
    return ops.gemm(x, w, a2=t, b2=lora_b, **kw), t


def decoder_forward(W: DecoderW, h: torch.Tensor, B: int, L: int, positions: torch.Tensor, kv_start, kv_end, *,
                    lora: Optional[LoraW] = None, saved: Optional[List[LayerSaved]] = None,
                    kv_sink: Optional[Callable[[int, torch.Tensor], None]] = None, final_norm: bool = True) -> torch.Tensor:
    """Qwen3 decoder stack over dense rows [B, L] (HF qwen3/modeling_qwen3.py:294-336, 378-430).

    h: merged input embeddings [B*L, d] bf16 (not modified).  Returns the final-normed hidden states [B*L, d].
    """
    cfg = W.cfg
    Hq, Hkv, D = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
    eps = cfg.rms_norm_eps
    theta = cfg.rope_parameters["rope_theta"] if hasattr(cfg, "rope_parameters") else cfg.rope_theta
    qo, ko, vo = 0, Hq * D, (Hq + Hkv) * D
    assert saved is None or kv_sink is None, "the KV sink reads roped K|V from the fused buffer; the training path keeps that buffer pre-norm"
    rope = _rope_table(L, D, theta, h.device)                          # cos/sin of positions 0..L-1, built once per (L, theta)
    for li, Lw in enumerate(W.layers):
        lw = lora.layers[li] if lora is not None else None
        ls = lora.scale if lora is not None else 1.0
        S = LayerSaved() if saved is not None else None
        if S is not None:
            xn, rstd1 = ops.rmsnorm(h, Lw.ln1, eps, want_rstd=True)
            S.h_in, S.rstd1, S.xn1 = h, rstd1, xn
        else:
            xn = ops.rmsnorm(h, Lw.ln1, eps)
        qkv, t = _lin(xn, Lw.w_qkv, lora_a=lw.a_qkv if lw else None, lora_b=lw.b_qkv if lw else None, lora_scale=ls)
        if S is not None:
            # training: the roped q|k go to their own buffer, the GEMM output keeps the pre-norm q|k (qk-norm backward) and V -- no copy
            S.t_qkv = t
            S.qkv_pre = qkv
            qk = torch.empty(h.shape[0], vo, device=h.device, dtype=torch.bfloat16)
            ops.qk_rope_(qkv, Hq, Hkv, D, positions, theta, q_norm_w=Lw.q_norm, k_norm_w=Lw.k_norm, eps=eps, mode=0, out=qk, rope=rope)
            q, k, v = qk[:, qo:ko], qk[:, ko:vo], qkv[:, vo:]
            S.q, S.k, S.v = q, k, v
        else:
            ops.qk_rope_(qkv, Hq, Hkv, D, positions, theta, q_norm_w=Lw.q_norm, k_norm_w=Lw.k_norm, eps=eps, mode=0, rope=rope)
            q, k, v = qkv[:, qo:ko], qkv[:, ko:vo], qkv[:, vo:]
        if kv_sink is not None:
            kv_sink(li, qkv)
        if S is not None:
            attn, lse = ops.attn_fwd(q, k, v, B, L, Hq, Hkv, D, kv_start=kv_start, kv_end=kv_end, causal=True, want_lse=True)
            S.attn, S.lse = attn, lse
        else:
            attn = ops.attn_fwd(q, k, v, B, L, Hq, Hkv, D, kv_start=kv_start, kv_end=kv_end, causal=True)
        h2, t = _lin(attn, Lw.w_o, lora_a=lw.a_o if lw else None, lora_b=lw.b_o if lw else None, lora_scale=ls, residual=h)
        if S is not None:
            S.t_o = t
            xn2, rstd2 = ops.rmsnorm(h2, Lw.ln2, eps, want_rstd=True)
            S.h_mid, S.rstd2, S.xn2 = h2, rstd2, xn2
            gu = torch.empty(h.shape[0], Lw.w_gu.shape[0], device=h.device, dtype=torch.bfloat16)
        else:
            xn2 = ops.rmsnorm(h2, Lw.ln2, eps)
            gu = None
        act, t = _lin(xn2, Lw.w_gu, lora_a=lw.a_gu if lw else None, lora_b=lw.b_gu if lw else None, lora_scale=ls, act=1, aux_out=gu)
        if S is not None:
            S.t_gu, S.gu, S.act = t, gu, act
        h, t = _lin(act, Lw.w_down, lora_a=lw.a_down if lw else None, lora_b=lw.b_down if lw else None, lora_scale=ls, residual=h2)
        if S is not None:
            S.t_down = t
            saved.append(S)
    if not final_norm:
        return h
    return ops.rmsnorm(h, W.final_norm, eps)


def encoder_forward(W: EncoderW, input_ids: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor:
    """NT-v2 / ESM encoder, forward only (the reference wraps it in no_grad, dna_llm.py:121): last hidden state
    after the final LayerNorm (== hidden_states[-1], HF esm/modeling_esm.py:511-512).  Returns [n_seq*S, d] bf16."""
    cfg = W.cfg
    n_seq, S = input_ids.shape
    nh = cfg.num_attention_heads
    D = cfg.hidden_size // nh
    d = cfg.hidden_size
    eps = cfg.layer_norm_eps
    x = ops.embed_gather(input_ids, W.embed, keep=attention_mask)          # embeddings * attention_mask (esm:232-233)
    ks, ke = mask_window(attention_mask)
    pos = torch.arange(S, device=x.device, dtype=torch.int32).repeat(n_seq)
    for Lw in W.layers:
        xn = ops.layernorm(x, Lw.ln1_w, Lw.ln1_b, eps)
        qkv = ops.gemm(xn, Lw.w_qkv, bias=Lw.b_qkv)
        ops.qk_rope_(qkv, nh, nh, D, pos, 10000.0, q_scale=D ** -0.5, mode=1)
        a = ops.attn_fwd(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], n_seq, S, nh, nh, D, kv_start=ks, kv_end=ke, scale=1.0, causal=False)
        x = ops.gemm(a, Lw.w_o, bias=Lw.b_o, residual=x)
        xn = ops.layernorm(x, Lw.ln2_w, Lw.ln2_b, eps)
        act = ops.gemm(xn, Lw.w_gu, bias=Lw.b_gu, act=1)
        x = ops.gemm(act, Lw.w_down, bias=Lw.b_down, residual=x)
    return ops.layernorm(x, W.final_ln_w, W.final_ln_b, eps)


def dna_row_map(input_ids: torch.Tensor, dna_token_id: int, dna_mask: torch.Tensor, batch_idx_map: List[int]):
    """Destination row (in the flattened [B*L] embedding buffer) of every encoder output row, or -1 for DNA pads.

    Restates dna_llm.py:166-177 + :216-229 without the per-sequence host syncs: valid tokens of the sequences, taken
    in (batch item, sequence) order, fill the <|dna_pad|> slots in row-major order.  Returns (row_map int32
    [n_seq*S], n_features, n_slots) with the two counts still on the device.
    """
    n_seq, S = dna_mask.shape
    dev = input_ids.device
    order = sorted(range(n_seq), key=lambda i: batch_idx_map[i])           # stable: regroup by batch item
    order_t = torch.tensor(order, device=dev, dtype=torch.long)
    valid_len = dna_mask.sum(dim=1)                                        # [:valid_length] slicing (dna_llm.py:168-169)
    tok = torch.arange(S, device=dev)
    valid = tok[None, :] < valid_len[:, None]                              # [n_seq, S] (first valid_len tokens)
    valid_sorted = valid[order_t]
    rank_sorted = valid_sorted.reshape(-1).long().cumsum(0) - 1            # feature index of each valid token
    slot_mask = (input_ids == dna_token_id).reshape(-1)
    n_slots = slot_mask.sum()
    n_feat = valid.sum()
    # position of the r-th slot
    slot_rank = slot_mask.long().cumsum(0) - 1
    N = slot_mask.numel()
    pos_of_rank = torch.full((N + 1,), -1, device=dev, dtype=torch.long)
    pos_of_rank.scatter_(0, torch.where(slot_mask, slot_rank, torch.full_like(slot_rank, N)), torch.arange(N, device=dev))
    dest_sorted = torch.where(valid_sorted.reshape(-1), pos_of_rank[rank_sorted.clamp(min=0, max=N)], torch.full_like(rank_sorted, -1))
    dest = torch.empty(n_seq, S, device=dev, dtype=torch.long)
    dest[order_t] = dest_sorted.view(n_seq, S)
    return dest.reshape(-1).to(torch.int32), n_feat, n_slots
