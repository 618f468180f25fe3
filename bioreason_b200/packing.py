"""Weight layout in HBM for the hot path.

The parameter *containers* stay HF-shaped (`Qwen3ForCausalLM`, `EsmForMaskedLM`: same module tree and
state_dict keys the reference's callers walk) but their storage is re-pointed into fused
buffers laid out for the kernels:

  decoder layer : w_qkv [(Hq+2Hkv)*D, d]   rows = q heads | k heads | v heads       (one QKV GEMM)
                  w_gu  [2F, d]            blocks of 16 rows = 8 gate rows | 8 up rows (SwiGLU fused in the GEMM epilogue,
                                           16-byte-aligned gate/up column groups for the backward kernels)
                  w_o   [d, Hq*D], w_down [d, F]
  encoder layer : w_qkv [3*d, d] (+ b_qkv), w_o, w_gu [2F, d] interleaved (x1_j, x2_j), w_down
Frozen weights additionally get a transposed copy (`*_T`, [in, out]) so that the backward dX GEMMs are also
K-major x K-major (no MN-major descriptors); 180 GB of HBM makes the second copy (≈8 GB for Qwen3-4B) free.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import torch


def gu_views(w_gu: torch.Tensor):
    """(gate, up) views [F/8, 8, ...] of a gate/up-blocked buffer whose leading dim is 2F (blocks of 8 gate | 8 up)."""
    F2 = w_gu.shape[0]
    v = w_gu.view(F2 // 16, 2, 8, *w_gu.shape[1:])
    return v[:, 0], v[:, 1]


def _repoint_gu(gate_p, up_p, w_gu):
    F = gate_p.shape[0]
    assert F % 8 == 0
    gv, uv = gu_views(w_gu)
    with torch.no_grad():
        gv.copy_(gate_p.data.to(w_gu.dtype).view(F // 8, 8, -1))
        uv.copy_(up_p.data.to(w_gu.dtype).view(F // 8, 8, -1))
    # a [F, d] parameter cannot alias the blocked layout as one strided view; the container keeps its own (frozen)
    # storage and `refresh_gu()` re-derives the kernel copy after a load_state_dict
    gate_p.data = gate_p.data.to(device=w_gu.device, dtype=w_gu.dtype)
    up_p.data = up_p.data.to(device=w_gu.device, dtype=w_gu.dtype)


def _repoint(param: torch.nn.Parameter, view: torch.Tensor):
    with torch.no_grad():
        view.copy_(param.data.to(view.dtype))
    param.data = view


@dataclass
class DecoderLayerW:
    ln1: torch.Tensor
    ln2: torch.Tensor
    q_norm: torch.Tensor
    k_norm: torch.Tensor
    w_qkv: torch.Tensor
    w_o: torch.Tensor
    w_gu: torch.Tensor
    w_down: torch.Tensor
    w_qkv_T: Optional[torch.Tensor] = None
    w_o_T: Optional[torch.Tensor] = None
    w_gu_T: Optional[torch.Tensor] = None
    w_down_T: Optional[torch.Tensor] = None


@dataclass
class DecoderW:
    cfg: object
    embed: torch.Tensor            # [V, d] (tied lm_head)
    lm_head: torch.Tensor          # [V, d]
    final_norm: torch.Tensor
    layers: List[DecoderLayerW] = field(default_factory=list)
    lm_head_T: Optional[torch.Tensor] = None   # [d, V] for dH = dlogits @ W

    def build_transposes(self):
        for L in self.layers:
            if L.w_qkv_T is None:
                L.w_qkv_T = L.w_qkv.t().contiguous()
                L.w_o_T = L.w_o.t().contiguous()
                L.w_gu_T = L.w_gu.t().contiguous()
                L.w_down_T = L.w_down.t().contiguous()
        if self.lm_head_T is None:
            self.lm_head_T = self.lm_head.t().contiguous()


def pack_decoder(model, device="cuda") -> DecoderW:
    """Fuse a HF Qwen3ForCausalLM's weights into kernel layout (bf16, on `device`) and re-point its parameters."""
    cfg = model.config
    d, F = cfg.hidden_size, cfg.intermediate_size
    Hq, Hkv, D = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
    bf = torch.bfloat16
    emb_p = model.model.embed_tokens.weight
    embed = torch.empty(emb_p.shape, device=device, dtype=bf)
    _repoint(emb_p, embed)
    if cfg.tie_word_embeddings:
        model.lm_head.weight = model.model.embed_tokens.weight
        lm_head = embed
    else:
        lm_head = torch.empty(model.lm_head.weight.shape, device=device, dtype=bf)
        _repoint(model.lm_head.weight, lm_head)
    fn = torch.empty(d, device=device, dtype=bf)
    _repoint(model.model.norm.weight, fn)
    W = DecoderW(cfg=cfg, embed=embed, lm_head=lm_head, final_norm=fn)
    for layer in model.model.layers:
        at, mlp = layer.self_attn, layer.mlp
        w_qkv = torch.empty((Hq + 2 * Hkv) * D, d, device=device, dtype=bf)
        _repoint(at.q_proj.weight, w_qkv[: Hq * D])
        _repoint(at.k_proj.weight, w_qkv[Hq * D: (Hq + Hkv) * D])
        _repoint(at.v_proj.weight, w_qkv[(Hq + Hkv) * D:])
        w_o = torch.empty(d, Hq * D, device=device, dtype=bf)
        _repoint(at.o_proj.weight, w_o)
        w_gu = torch.empty(2 * F, d, device=device, dtype=bf)
        _repoint_gu(mlp.gate_proj.weight, mlp.up_proj.weight, w_gu)
        w_down = torch.empty(d, F, device=device, dtype=bf)
        _repoint(mlp.down_proj.weight, w_down)
        small = {}
        for name, p in (("ln1", layer.input_layernorm.weight), ("ln2", layer.post_attention_layernorm.weight),
                        ("q_norm", at.q_norm.weight), ("k_norm", at.k_norm.weight)):
            t = torch.empty(p.shape, device=device, dtype=bf)
            _repoint(p, t)
            small[name] = t
        W.layers.append(DecoderLayerW(w_qkv=w_qkv, w_o=w_o, w_gu=w_gu, w_down=w_down, **small))
    # buffers (rotary inv_freq) are not used by the kernels; leave them where they are
    return W


def refresh_decoder_gu(model, W: DecoderW):
    """Re-derive the blocked gate/up kernel copies from the container's gate_proj / up_proj (after loading weights)."""
    with torch.no_grad():
        for layer, Lw in zip(model.model.layers, W.layers):
            gv, uv = gu_views(Lw.w_gu)
            F = layer.mlp.gate_proj.weight.shape[0]
            gv.copy_(layer.mlp.gate_proj.weight.data.view(F // 8, 8, -1))
            uv.copy_(layer.mlp.up_proj.weight.data.view(F // 8, 8, -1))
        for L in W.layers:
            L.w_qkv_T = L.w_o_T = L.w_gu_T = L.w_down_T = None
        W.lm_head_T = None


@dataclass
class EncoderLayerW:
    ln1_w: torch.Tensor
    ln1_b: torch.Tensor
    ln2_w: torch.Tensor
    ln2_b: torch.Tensor
    w_qkv: torch.Tensor
    b_qkv: torch.Tensor
    w_o: torch.Tensor
    b_o: torch.Tensor
    w_gu: torch.Tensor           # gated: [2F, d] interleaved; plain GELU FFN is not on the NT-v2 path
    b_gu: Optional[torch.Tensor]
    w_down: torch.Tensor
    b_down: Optional[torch.Tensor]


@dataclass
class EncoderW:
    cfg: object
    embed: torch.Tensor
    final_ln_w: torch.Tensor
    final_ln_b: torch.Tensor
    layers: List[EncoderLayerW] = field(default_factory=list)


def pack_encoder(model, device="cuda") -> EncoderW:
    """Fuse a (NT-v2 patched) HF EsmForMaskedLM encoder into kernel layout; the MLM head is never packed (unused,
    SURVEY.md §8a A2: the reference computes it for nothing)."""
    cfg = model.config
    d, F = cfg.hidden_size, cfg.intermediate_size
    bf = torch.bfloat16
    if not getattr(cfg, "gated_mlp", False):
        raise NotImplementedError("only the NT-v2 gated-SiLU FFN encoder is on the hot path")
    if getattr(cfg, "position_embedding_type", "absolute") != "rotary" or cfg.emb_layer_norm_before or cfg.token_dropout:
        raise NotImplementedError("encoder kernels implement the NT-v2 configuration (rotary, no emb-LN-before, no token dropout)")

    def mv(p):
        t = torch.empty(p.shape, device=device, dtype=bf)
        _repoint(p, t)
        return t

    esm = model.esm
    W = EncoderW(cfg=cfg, embed=mv(esm.embeddings.word_embeddings.weight),
                 final_ln_w=mv(esm.encoder.emb_layer_norm_after.weight), final_ln_b=mv(esm.encoder.emb_layer_norm_after.bias))
    for layer in esm.encoder.layer:
        sa = layer.attention.self
        w_qkv = torch.empty(3 * d, d, device=device, dtype=bf)
        b_qkv = torch.empty(3 * d, device=device, dtype=bf)
        for i, lin in enumerate((sa.query, sa.key, sa.value)):
            _repoint(lin.weight, w_qkv[i * d:(i + 1) * d])
            _repoint(lin.bias, b_qkv[i * d:(i + 1) * d])
        w_gu = torch.empty(2 * F, d, device=device, dtype=bf)
        inter = layer.intermediate.dense                      # [2F, d]: rows [0,F) = x1 (gate), [F,2F) = x2
        gv, uv = gu_views(w_gu)
        with torch.no_grad():
            gv.copy_(inter.weight.data[:F].to(bf).view(F // 8, 8, d))
            uv.copy_(inter.weight.data[F:].to(bf).view(F // 8, 8, d))
        # the container keeps a [2F, d] parameter; give it its own bf16 storage (frozen, forward-only)
        inter.weight.data = inter.weight.data.to(device=device, dtype=bf)
        b_gu = None
        if inter.bias is not None:
            b_gu = torch.empty(2 * F, device=device, dtype=bf)
            with torch.no_grad():
                bg, bu = gu_views(b_gu)
                bg.copy_(inter.bias.data[:F].to(bf).view(F // 8, 8)); bu.copy_(inter.bias.data[F:].to(bf).view(F // 8, 8))
            inter.bias.data = inter.bias.data.to(device=device, dtype=bf)
        out = layer.output.dense
        W.layers.append(EncoderLayerW(
            ln1_w=mv(layer.attention.LayerNorm.weight), ln1_b=mv(layer.attention.LayerNorm.bias),
            ln2_w=mv(layer.LayerNorm.weight), ln2_b=mv(layer.LayerNorm.bias),
            w_qkv=w_qkv, b_qkv=b_qkv, w_o=mv(layer.attention.output.dense.weight), b_o=mv(layer.attention.output.dense.bias),
            w_gu=w_gu, b_gu=b_gu, w_down=mv(out.weight), b_down=mv(out.bias) if out.bias is not None else None))
    return W
