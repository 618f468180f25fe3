"""Policy forward with saved activations + the hand-written backward through the LoRA-adapted Qwen3 decoder,
the fused lm_head log-prob, and the DNA projector.

`policy_logps` is the differentiable equivalent of `_get_per_token_logps(...)[:, P-1:]` (grpo_trainer.py:510-520, :779):
gradients flow to the LoRA A/B masters and to `dna_projection` only -- the base weights are frozen and the encoder
is under no_grad in the reference (dna_llm.py:121).
"""
from __future__ import annotations

from typing import List, Optional

import torch

from . import engine, ops
from .engine import LayerSaved


class PolicyCtx:
    pass


def policy_forward(model, input_ids, attention_mask, dna_tokenized, batch_idx_map, keep_last: int, *, save: bool = True,
                   lora="policy", targets: Optional[torch.Tensor] = None) -> "tuple[torch.Tensor, Optional[PolicyCtx]]":
    """Returns (logps [B, keep_last] fp32, ctx).  lora: "policy" (adapters on), None (base weights = reference policy).
    targets: optional [B, keep_last] class ids scored at the last keep_last positions before the end (default: the realised next
    tokens input_ids[:, L-keep_last:]); entries < 0 are ignored (log-prob 0, no gradient) -- the SFT label mask."""
    W = model._dec
    dev = W.embed.device
    input_ids = input_ids.to(dev)
    attention_mask = attention_mask.to(dev)
    B, L = input_ids.shape
    use_lora = model._lora.w if (lora == "policy" and model._lora is not None) else None
    if lora is None and getattr(model, "_proj_ref", None) is not None:
        # the reference's ref_model is a deep copy taken at init (grpo_trainer.py:314-316): initial projector too
        pw, pb = model._proj_w16, model._proj_b16
        model._proj_w16, model._proj_b16 = model._proj_ref
        try:
            emb, aux = model.merged_embeddings(input_ids, dna_tokenized, batch_idx_map, return_proj_inputs=True)
        finally:
            model._proj_w16, model._proj_b16 = pw, pb
    else:
        emb, aux = model.merged_embeddings(input_ids, dna_tokenized, batch_idx_map, return_proj_inputs=True)
    ks, ke = engine.mask_window(attention_mask)
    pos = engine.forward_positions(B, L, dev)
    saved: Optional[List[LayerSaved]] = [] if save else None
    h = engine.decoder_forward(W, emb, B, L, pos, ks, ke, lora=use_lora, saved=saved, final_norm=False)
    eps = W.cfg.rms_norm_eps
    if save:
        hn, rstd_f = ops.rmsnorm(h, W.final_norm, eps, want_rstd=True)
    else:
        hn, rstd_f = ops.rmsnorm(h, W.final_norm, eps), None
    n = keep_last
    cols = torch.arange(L - 1 - n, L - 1, device=dev)
    rows = (torch.arange(B, device=dev)[:, None] * L + cols[None, :]).reshape(-1).to(torch.int32)
    h_sel = ops.gather_rows(hn, rows)
    tgt = (input_ids[:, L - n:] if targets is None else targets.to(dev)).reshape(-1).to(torch.int32)
    logp, lse = ops.lmhead_logprob(h_sel, W.lm_head, tgt)
    ctx = None
    if save:
        ctx = PolicyCtx()
        ctx.B, ctx.L, ctx.n = B, L, n
        ctx.saved, ctx.h_final, ctx.rstd_f = saved, h, rstd_f
        ctx.rows, ctx.h_sel, ctx.tgt, ctx.lse = rows, h_sel, tgt, lse
        ctx.pos, ctx.ks, ctx.ke, ctx.aux = pos, ks, ke, aux
        ctx.use_lora = use_lora is not None
    return logp.view(B, n), ctx


@torch.no_grad()
def policy_backward(model, ctx: PolicyCtx, dlogp: torch.Tensor, on_layer_done=None):
    """Accumulates d(sum dlogp * logp) into the LoRA flat gradient buffer and the projector's .grad buffers.
    on_layer_done(layer_index): called right after the kernels producing that layer's adapter gradients were enqueued (the
    trainer hangs the overlapped gradient all-reduce of that layer's slice on it)."""
    W = model._dec
    W.build_transposes()
    cfg = W.cfg
    Hq, Hkv, D, d, F = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim, cfg.hidden_size, cfg.intermediate_size
    theta = cfg.rope_parameters["rope_theta"] if hasattr(cfg, "rope_parameters") else cfg.rope_theta
    eps = cfg.rms_norm_eps
    B, L = ctx.B, ctx.L
    M = B * L
    dev = W.embed.device
    lora = model._lora if ctx.use_lora else None
    r = lora.r if lora else 0
    s = lora.scale if lora else 1.0
    qo, ko, vo = 0, Hq * D, (Hq + Hkv) * D

    # ---- lm_head: dlogits tiles recomputed from (h_sel, W) and the saved LSE, then dH = dlogits @ W
    g = dlogp.reshape(-1).float().contiguous()
    dlogits = ops.lmhead_dlogits(ctx.h_sel, W.lm_head, ctx.tgt, ctx.lse, g)
    dh_sel = ops.gemm(dlogits, W.lm_head_T)
    del dlogits
    dhn = torch.zeros(M, d, device=dev, dtype=torch.bfloat16)
    ops.scatter_rows_(dhn, dh_sel, ctx.rows)
    dh = ops.rmsnorm_bwd(ctx.h_final, W.final_norm, ctx.rstd_f, dhn)
    del dhn

    def lin_bwd(dy, w_T, x_in, t_saved, li, names, a_T, b_T, big_cols=None, **kw):
        """dx = dy @ W (+ LoRA path) and LoRA grads.  names: LoRA target names fused in this linear (in packed order)."""
        if lora is None:
            return ops.gemm(dy, w_T, **kw)
        u = ops.gemm(dy, b_T, alpha=s)                                    # [M, r * len(names)] = s * dy @ B
        dx = ops.gemm(dy, w_T, a2=u, b2=a_T, **kw)
        return dx, u

    for li in range(len(W.layers) - 1, -1, -1):
        Lw, S = W.layers[li], ctx.saved[li]
        Tl = lora.wT[li] if lora else None
        # ---------------- MLP: h_out = h_mid + down(act)
        if lora:
            dact, u = lin_bwd(dh, Lw.w_down_T, S.act, S.t_down, li, ("down_proj",), Tl["a_down_T"], Tl["b_down_T"])
            ops.lora_grad_tn(dh, S.t_down, [(lora.grad_view(li, "down_proj", "B"), 0, d, 0, r)])               # dB = dy^T t
            ops.lora_grad_tn(S.act, u, [(lora.grad_view(li, "down_proj", "A"), 0, F, 0, r)], mode=1)           # dA = u^T x
        else:
            dact = ops.gemm(dh, Lw.w_down_T)
        dgu = ops.swiglu_bwd(S.gu, dact)
        del dact
        if lora:
            dxn2, u = lin_bwd(dgu, Lw.w_gu_T, S.xn2, S.t_gu, li, ("gate_proj", "up_proj"), Tl["a_gu_T"], Tl["b_gu_T"])
            # one product dgu^T [2F] x t_gu [2r]: gate rows keep their r columns, up rows theirs (cross blocks are discarded)
            ops.lora_grad_tn(dgu, S.t_gu, [(lora.grad_view(li, "gate_proj", "B"), 0, 2 * F, 0, r),
                                           (lora.grad_view(li, "up_proj", "B"), 0, 2 * F, r, r)], mode=2)
            ops.lora_grad_tn(S.xn2, u[:, :r], [(lora.grad_view(li, "gate_proj", "A"), 0, d, 0, r)], mode=1)
            ops.lora_grad_tn(S.xn2, u[:, r:], [(lora.grad_view(li, "up_proj", "A"), 0, d, 0, r)], mode=1)
        else:
            dxn2 = ops.gemm(dgu, Lw.w_gu_T)
        del dgu
        dh_mid = ops.rmsnorm_bwd(S.h_mid, Lw.ln2, S.rstd2, dxn2, dres=dh)
        del dxn2
        # ---------------- attention: h_mid = h_in + o_proj(attn)
        if lora:
            dattn, u = lin_bwd(dh_mid, Lw.w_o_T, S.attn, S.t_o, li, ("o_proj",), Tl["a_o_T"], Tl["b_o_T"])
            ops.lora_grad_tn(dh_mid, S.t_o, [(lora.grad_view(li, "o_proj", "B"), 0, d, 0, r)])
            ops.lora_grad_tn(S.attn, u, [(lora.grad_view(li, "o_proj", "A"), 0, Hq * D, 0, r)], mode=1)
        else:
            dattn = ops.gemm(dh_mid, Lw.w_o_T)
        dqkv = torch.empty(M, (Hq + 2 * Hkv) * D, device=dev, dtype=torch.bfloat16)
        ops.attn_bwd(S.q, S.k, S.v, S.attn, dattn, S.lse, dqkv[:, qo:ko], dqkv[:, ko:vo], dqkv[:, vo:],
                     B, L, Hq, Hkv, D, kv_start=ctx.ks, kv_end=ctx.ke)
        del dattn
        ops.qk_rope_bwd_(dqkv, S.qkv_pre, Hq, Hkv, D, Lw.q_norm, Lw.k_norm, ctx.pos, theta, eps)
        if lora:
            dxn1, u = lin_bwd(dqkv, Lw.w_qkv_T, S.xn1, S.t_qkv, li, ("q_proj", "k_proj", "v_proj"), Tl["a_qkv_T"], Tl["b_qkv_T"])
            # one product dqkv^T [(Hq+2Hkv)D] x t_qkv [3r]: the q / k / v row blocks keep their own r columns
            ops.lora_grad_tn(dqkv, S.t_qkv, [(lora.grad_view(li, "q_proj", "B"), qo, ko, 0, r), (lora.grad_view(li, "k_proj", "B"), ko, vo, r, r),
                                             (lora.grad_view(li, "v_proj", "B"), vo, vo + Hkv * D, 2 * r, r)])
            for j, name in enumerate(("q_proj", "k_proj", "v_proj")):
                ops.lora_grad_tn(S.xn1, u[:, j * r:(j + 1) * r], [(lora.grad_view(li, name, "A"), 0, d, 0, r)], mode=1)
        else:
            dxn1 = ops.gemm(dqkv, Lw.w_qkv_T)
        del dqkv
        dh = ops.rmsnorm_bwd(S.h_in, Lw.ln1, S.rstd1, dxn1, dres=dh_mid)
        del dxn1, dh_mid
        ctx.saved[li] = None                                              # free this layer's activations
        if on_layer_done is not None and lora is not None:
            on_layer_done(li)

    # ---------------- projector: emb rows that came from DNA features
    if ctx.aux is not None and model.dna_projection.weight.requires_grad:
        enc, row_map = ctx.aux
        dE = ops.gather_rows(dh, row_map)                                  # rows with row_map < 0 (DNA pads) come back as zeros
        dE_T = ops.transpose(dE)                                           # [d_text, n']
        enc_T = ops.transpose(enc)                                         # [d_dna, n']
        gw = ops.gemm(dE_T, enc_T, out_dtype=torch.float32)                # [d_text, d_dna]
        model._proj_grad_w.add_(gw)
        ops.colsum_accumulate_(model._proj_grad_b, dE)
    return dh


def sft_step(model, input_ids, attention_mask, dna_tokenized, batch_idx_map, labels, *, backward: bool = True, grad_scale: float = 1.0):
    """One supervised step (train_dna_qwen.py:179-213 -> HF ForCausalLMLoss, loss/loss_utils.py:28-67): shift, ignore -100, mean CE.
    Returns the loss; with backward=True accumulates d(loss * grad_scale) into the LoRA / projector gradient buffers."""
    dev = model._dec.embed.device
    labels = labels.to(dev)
    B, L = labels.shape
    tgt = torch.where(labels[:, 1:] == -100, torch.full_like(labels[:, 1:], -1), labels[:, 1:])           # position t predicts label t+1
    valid = tgt >= 0
    n = valid.sum().clamp(min=1).float()
    lp, ctx = policy_forward(model, input_ids, attention_mask, dna_tokenized, batch_idx_map, L - 1, save=backward, targets=tgt)
    loss = -(lp * valid).sum() / n
    if backward:
        policy_backward(model, ctx, (-(valid.float()) / n) * grad_scale)
    return loss


class _PolicyLogps(torch.autograd.Function):
    """Autograd bridge so `loss.backward()` (HF Trainer style) reaches the hand-written backward."""

    @staticmethod
    def forward(ctx, model, input_ids, attention_mask, dna_tokenized, batch_idx_map, keep_last, *trainable):
        logp, pctx = policy_forward(model, input_ids, attention_mask, dna_tokenized, batch_idx_map, keep_last, save=True)
        ctx.model, ctx.pctx, ctx.n_train = model, pctx, len(trainable)
        return logp

    @staticmethod
    def backward(ctx, dlogp):
        model = ctx.model
        lora = model._lora
        before = lora.flat_grad.clone() if lora is not None else None
        pw, pb = model._proj_grad_w.clone(), model._proj_grad_b.clone()
        policy_backward(model, ctx.pctx, dlogp)
        grads = []
        if lora is not None:
            delta = lora.flat_grad - before
            lora.flat_grad.copy_(before)                                   # autograd does the accumulation into .grad
            off = 0
            for p in lora.params:
                grads.append(delta[off:off + p.numel()].view_as(p)); off += p.numel()
        gw, gb = model._proj_grad_w - pw, model._proj_grad_b - pb
        model._proj_grad_w.copy_(pw); model._proj_grad_b.copy_(pb)
        grads += [gw, gb]
        return (None, None, None, None, None, None, *grads[:ctx.n_train])


def policy_logps_autograd(model, input_ids, attention_mask, dna_tokenized, batch_idx_map, keep_last):
    trainable = (list(model._lora.params) if model._lora is not None else []) + [model.dna_projection.weight, model.dna_projection.bias]
    return _PolicyLogps.apply(model, input_ids, attention_mask, dna_tokenized, batch_idx_map, keep_last, *trainable)
