"""LoRA adapters for the Qwen3 decoder in kernel layout (peft is not installed in this image).

`inject_lora` mirrors what the reference's callers do with peft (`reason.py:362-394`, `train_dna_qwen.py:136-177`):
every nn.Linear leaf of the text model except `lm_head` gets rank-r adapters (r=32, alpha=64, gaussian init:
A ~ N(0, 1/r), B = 0), base weights frozen.  Module / parameter names follow peft (`base_layer`,
`lora_A.default.weight`, `lora_B.default.weight`) so checkpoints keep their keys.  The fp32 master parameters are
what the optimizer sees; `LoraState.sync()` re-packs them (bf16, fused/blocked like the base weights, plus the
transposes the backward GEMMs read) after every optimizer step.  LoRA dropout is not applied (p = 0); the
reference's 0.05 is a regulariser, not part of the path's arithmetic contract.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List

import torch
import torch.nn as nn

from .engine import LoraLayerW, LoraW
from .packing import gu_views

TARGETS = ("q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj")


class LoraLinear(nn.Module):
    """Container only (its forward is never on the product path): peft-shaped names around a frozen base Linear."""

    def __init__(self, base: nn.Linear, r: int, alpha: float, generator=None):
        super().__init__()
        self.base_layer = base
        base.weight.requires_grad_(False)
        dev = base.weight.device
        a = nn.Linear(base.in_features, r, bias=False, device=dev, dtype=torch.float32)
        b = nn.Linear(r, base.out_features, bias=False, device=dev, dtype=torch.float32)
        with torch.no_grad():
            a.weight.copy_(torch.randn(a.weight.shape, generator=generator, device="cpu").to(dev) / r)   # peft 'gaussian': std 1/r
            b.weight.zero_()
        self.lora_A = nn.ModuleDict({"default": a})
        self.lora_B = nn.ModuleDict({"default": b})
        self.r, self.lora_alpha, self.scaling = r, alpha, alpha / r

    @property
    def weight(self):
        return self.base_layer.weight

    def forward(self, x):                                                  # pragma: no cover - reference semantics only
        return self.base_layer(x) + self.lora_B["default"](self.lora_A["default"](x.float())).to(x.dtype) * self.scaling


@dataclass
class LoraFlat:
    params: List[nn.Parameter]
    flat_grad: torch.Tensor


class LoraState:
    def __init__(self, text_model, dec_w, r: int, alpha: float, seed: int = 0):
        self.r, self.scale = r, alpha / r
        cfg = text_model.config
        self.cfg = cfg
        g = torch.Generator().manual_seed(seed)
        self.modules: List[Dict[str, LoraLinear]] = []
        for layer in text_model.model.layers:
            mods = {}
            for parent, names in ((layer.self_attn, ("q_proj", "k_proj", "v_proj", "o_proj")), (layer.mlp, ("gate_proj", "up_proj", "down_proj"))):
                for n in names:
                    lin = getattr(parent, n)
                    if not isinstance(lin, LoraLinear):
                        lin = LoraLinear(lin, r, alpha, g)
                        setattr(parent, n, lin)
                    mods[n] = lin
            self.modules.append(mods)
        for p in text_model.parameters():
            p.requires_grad_(False)
        self.params: List[nn.Parameter] = []
        for mods in self.modules:
            for n in TARGETS:
                for p in (mods[n].lora_A["default"].weight, mods[n].lora_B["default"].weight):
                    p.requires_grad_(True)
                    self.params.append(p)
        dev = dec_w.embed.device
        # one flat fp32 gradient buffer; every .grad is a contiguous view into it (single all-reduce bucket, C2)
        n = sum(p.numel() for p in self.params)
        self.flat_grad = torch.zeros(n, device=dev, dtype=torch.float32)
        off = 0
        self.grad_views: List[torch.Tensor] = []
        for p in self.params:
            self.grad_views.append(self.flat_grad[off:off + p.numel()].view_as(p))
            off += p.numel()
        self._alloc_packed(dec_w, dev)
        self.sync()

    # ------------------------------------------------------------------
    def _alloc_packed(self, W, dev):
        cfg, r = self.cfg, self.r
        d, F = cfg.hidden_size, cfg.intermediate_size
        Hq, Hkv, D = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
        nqkv = (Hq + 2 * Hkv) * D
        bf = torch.bfloat16
        z = lambda *s: torch.zeros(*s, device=dev, dtype=bf)
        self.w = LoraW(r=r, scale=self.scale, layers=[])
        self.wT: List[Dict[str, torch.Tensor]] = []
        for _ in self.modules:
            self.w.layers.append(LoraLayerW(a_qkv=z(3 * r, d), b_qkv=z(nqkv, 3 * r), a_o=z(r, Hq * D), b_o=z(d, r),
                                            a_gu=z(2 * r, d), b_gu=z(2 * F, 2 * r), a_down=z(r, F), b_down=z(d, r)))
            self.wT.append(dict(a_qkv_T=z(d, 3 * r), b_qkv_T=z(3 * r, nqkv), a_o_T=z(Hq * D, r), b_o_T=z(r, d),
                                a_gu_T=z(d, 2 * r), b_gu_T=z(2 * r, 2 * F), a_down_T=z(F, r), b_down_T=z(r, d)))

    @torch.no_grad()
    def sync(self):
        """fp32 masters -> bf16 kernel layout (+ transposes).  Off-block entries of the block-diagonal B stay zero."""
        cfg, r = self.cfg, self.r
        Hq, Hkv, D = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
        qo, ko, vo = 0, Hq * D, (Hq + Hkv) * D
        for mods, L, T in zip(self.modules, self.w.layers, self.wT):
            A = lambda n: mods[n].lora_A["default"].weight
            B = lambda n: mods[n].lora_B["default"].weight
            L.a_qkv[0:r].copy_(A("q_proj")); L.a_qkv[r:2 * r].copy_(A("k_proj")); L.a_qkv[2 * r:].copy_(A("v_proj"))
            L.b_qkv[qo:ko, 0:r].copy_(B("q_proj")); L.b_qkv[ko:vo, r:2 * r].copy_(B("k_proj")); L.b_qkv[vo:, 2 * r:].copy_(B("v_proj"))
            L.a_o.copy_(A("o_proj")); L.b_o.copy_(B("o_proj"))
            L.a_gu[0:r].copy_(A("gate_proj")); L.a_gu[r:].copy_(A("up_proj"))
            gv, uv = gu_views(L.b_gu)                                      # [F/8, 8, 2r] each
            F = B("gate_proj").shape[0]
            gv[..., 0:r].copy_(B("gate_proj").view(F // 8, 8, r)); uv[..., r:].copy_(B("up_proj").view(F // 8, 8, r))
            L.a_down.copy_(A("down_proj")); L.b_down.copy_(B("down_proj"))
            for k in ("a_qkv", "b_qkv", "a_o", "b_o", "a_gu", "b_gu", "a_down", "b_down"):
                T[k + "_T"].copy_(getattr(L, k).t())

    def zero_grad(self):
        self.flat_grad.zero_()

    def attach_grads(self):
        """Expose the flat buffer as the parameters' .grad (what a torch optimizer / DDP-style all-reduce consumes)."""
        for p, g in zip(self.params, self.grad_views):
            p.grad = g

    def layer_slice(self, layer: int):
        """[lo, hi) of the flat gradient buffer holding every adapter gradient of one decoder layer (parameters are laid out layer by layer)."""
        per = len(TARGETS) * 2
        lo = sum(g.numel() for g in self.grad_views[:layer * per])
        return lo, lo + sum(g.numel() for g in self.grad_views[layer * per:(layer + 1) * per])

    def grad_view(self, layer: int, name: str, which: str) -> torch.Tensor:
        idx = (layer * len(TARGETS) + TARGETS.index(name)) * 2 + (0 if which == "A" else 1)
        return self.grad_views[idx]


@torch.no_grad()
def build_rollout_weights(dec_w, lora: "LoraState | None", out=None):
    """Decode-time weights: W_eff = W + scale * B A per fused weight (the policy that rolls out is base + LoRA; merged with the
    tcgen05 GEMM, base weight as the epilogue residual), with the RMSNorm gains folded into the columns of the matrices that
    consume a normed input (w_qkv <- ln1, w_gu <- ln2, lm_head <- final norm) so the decode step needs no norm launches."""
    from . import ops
    from .packing import DecoderLayerW, DecoderW
    if out is None:
        out = DecoderW(cfg=dec_w.cfg, embed=dec_w.embed, lm_head=torch.empty_like(dec_w.lm_head), final_norm=dec_w.final_norm)
        out.folded = True
        for Lw in dec_w.layers:
            out.layers.append(DecoderLayerW(ln1=Lw.ln1, ln2=Lw.ln2, q_norm=Lw.q_norm, k_norm=Lw.k_norm, w_qkv=torch.empty_like(Lw.w_qkv),
                                            w_o=torch.empty_like(Lw.w_o) if lora is not None else Lw.w_o, w_gu=torch.empty_like(Lw.w_gu),
                                            w_down=torch.empty_like(Lw.w_down) if lora is not None else Lw.w_down))
        out.lm_head.copy_(dec_w.lm_head)
        ops.scale_columns_(out.lm_head, dec_w.final_norm)                  # frozen: folded once
    for i, (Lw, Lo) in enumerate(zip(dec_w.layers, out.layers)):
        if lora is not None:
            if Lo.w_o.data_ptr() == Lw.w_o.data_ptr() or Lo.w_down.data_ptr() == Lw.w_down.data_ptr():
                raise RuntimeError("rollout weights alias the frozen base weights (built before enable_lora); rebuild them with out=None")
            s, Ll, T = lora.scale, lora.w.layers[i], lora.wT[i]
            # [N, K] = B[N, r'] @ (A^T)[K, r']^T ; K-major operands: A_op = B (K = r'), B_op = A^T ([K, r'])
            ops.gemm(Ll.b_qkv, T["a_qkv_T"], alpha=s, residual=Lw.w_qkv, out=Lo.w_qkv)
            ops.gemm(Ll.b_o, T["a_o_T"], alpha=s, residual=Lw.w_o, out=Lo.w_o)
            ops.gemm(Ll.b_gu, T["a_gu_T"], alpha=s, residual=Lw.w_gu, out=Lo.w_gu)
            ops.gemm(Ll.b_down, T["a_down_T"], alpha=s, residual=Lw.w_down, out=Lo.w_down)
        else:
            Lo.w_qkv.copy_(Lw.w_qkv); Lo.w_gu.copy_(Lw.w_gu)
        ops.scale_columns_(Lo.w_qkv, Lw.ln1)
        ops.scale_columns_(Lo.w_gu, Lw.ln2)
    return out
