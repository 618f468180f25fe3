"""Rollout engine: prefix-shared prefill + CUDA-graph decode loop on the paged-KV kernels.

Replaces `DNALLMModel.generate` -> `text_model.generate(inputs_embeds=..., use_cache=True, **kw)` (dna_llm.py:246-305; HF
generation/utils.py:2760-2800).  Semantics kept: completion-only ids; position_ids = cumsum(mask)-1 (pads excluded);
warper order temperature -> top-k -> top-p; finished rows emit pad; output trimmed to the longest unfinished row.
Redundancy removed: identical consecutive prompts (the G samples of a GRPO group) are encoded and
prefilled once and share their prompt KV pages.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from types import SimpleNamespace
from typing import List, Optional

import torch

from . import engine, ops
from .packing import DecoderW

PAGE = 64


@dataclass
class SamplingParams:
    max_new_tokens: int = 20
    do_sample: bool = False
    temperature: float = 1.0
    top_k: int = 50
    top_p: float = 1.0
    eos_token_id: Optional[int] = None
    pad_token_id: Optional[int] = None

    @classmethod
    def from_hf_kwargs(cls, model_cfg, kwargs):
        gc = kwargs.get("generation_config", None)
        def pick(name, default):
            if name in kwargs and kwargs[name] is not None:
                return kwargs[name]
            if gc is not None and getattr(gc, name, None) is not None:
                return getattr(gc, name)
            return default
        eos = pick("eos_token_id", getattr(model_cfg, "eos_token_id", None))
        if isinstance(eos, (list, tuple)):
            uniq = sorted(set(int(e) for e in eos))
            if len(uniq) > 1:
                # HF stops a row on ANY of the ids; the sampler kernel tracks one.  Refuse instead of silently keeping the first.
                raise NotImplementedError(f"generate(): {len(uniq)} distinct eos_token_id values {uniq}; the decode kernels stop on a single id "
                                          "(the reference trainer masks on processing_class.eos_token_id, grpo_trainer.py:605) -- pass that one")
            eos = uniq[0] if uniq else None
        pad = pick("pad_token_id", getattr(model_cfg, "pad_token_id", None))
        if pad is None:
            pad = eos if eos is not None else 0
        return cls(max_new_tokens=int(pick("max_new_tokens", 20)), do_sample=bool(pick("do_sample", False)),
                   temperature=float(pick("temperature", 1.0)), top_k=int(pick("top_k", 50) or 0), top_p=float(pick("top_p", 1.0)),
                   eos_token_id=eos, pad_token_id=pad)


def detect_group_size(input_ids: torch.Tensor, dna_tokenized, batch_idx_map) -> torch.Tensor:
    """eq[i] = row i is identical to row i-1 (text ids and its DNA sequences) -- device tensor, no sync."""
    B = input_ids.shape[0]
    eq = torch.zeros(B, dtype=torch.bool, device=input_ids.device)
    if B > 1:
        eq[1:] = (input_ids[1:] == input_ids[:-1]).all(dim=1)
        if dna_tokenized is not None and batch_idx_map:
            counts = [0] * B
            for b in batch_idx_map:
                counts[b] += 1
            if len(set(counts)) == 1 and list(batch_idx_map) == sorted(batch_idx_map):
                d = dna_tokenized["input_ids"].to(input_ids.device).view(B, -1)
                eq[1:] &= (d[1:] == d[:-1]).all(dim=1)
            else:
                eq[:] = False
    return eq


def group_size_from_flags(eq: List[bool]) -> int:
    B = len(eq)
    for G in range(B, 0, -1):
        if B % G == 0 and all(eq[i] for i in range(B) if i % G != 0):
            return G
    return 1


def stream_gate_plan(layer_grids, lm_head_grid):
    """Wait targets of the optional stream gate (br_stream_gate) for one token step.  layer_grids: per layer {w_qkv, w_o, w_gu, w_down: CTAs
    of that launch}.  Every gated launch adds its grid to one counter when its weights are on chip; a launch waits for the cumulative count
    of everything launched before it -- except the first qkv GEMM of a step (follows the embedding gather) and every o_proj (follows the
    attention): HBM idles before those anyway.  Returns ({(layer, name) | "lm_head": target or None}, arrivals per step)."""
    plan, acc = {}, 0
    for li, grids in enumerate(layer_grids):
        for nm in ("w_qkv", "w_o", "w_gu", "w_down"):
            plan[(li, nm)] = None if (nm == "w_o" or (nm == "w_qkv" and li == 0)) else acc
            acc += int(grids[nm])
    plan["lm_head"] = acc
    return plan, acc + int(lm_head_grid)


def plan_pages(plen: List[int], G: int, C: int):
    """KV page plan of a rollout (host side, pure).  plen[u] = prompt length of unique prompt u; every prompt is sampled G times
    for C new tokens.  Full prompt pages (the first n_shared of every group; one count for all groups) are shared by the G rows
    of a group; the partially filled tail page and the pages of generated tokens are private per row.
    Returns dict(n_shared, max_pages, n_pages, table [U*G][max_pages] (lists), prefill_pages [U] (pages holding the prompt,
    i.e. row 0 of the group), tail_copies [(src_page, dst_page)] to replicate row 0's partially shared tail pages)."""
    U = len(plen)
    n_full = [l // PAGE for l in plen]
    n_shared = min(n_full) if G > 1 else 0
    priv = [math.ceil((l - n_shared * PAGE + C) / PAGE) for l in plen]
    max_pages = n_shared + max(priv)
    nxt = 0
    table = [[0] * max_pages for _ in range(U * G)]
    prefill_pages, tail_copies = [], []
    for u in range(U):
        shared = list(range(nxt, nxt + n_shared)); nxt += n_shared
        for g in range(G):
            r = u * G + g
            mine = list(range(nxt, nxt + priv[u])); nxt += priv[u]
            table[r][:n_shared] = shared
            table[r][n_shared:n_shared + priv[u]] = mine
        n_prompt_pages = math.ceil(plen[u] / PAGE)
        prefill_pages.append(table[u * G][:n_prompt_pages])
        for j in range(n_shared, n_prompt_pages):
            for g in range(1, G):
                tail_copies.append((table[u * G][j], table[u * G + g][j]))
    return dict(n_shared=n_shared, max_pages=max_pages, n_pages=nxt, table=table, prefill_pages=prefill_pages, tail_copies=tail_copies)


class RolloutEngine:
    """Owns the KV page pool, decode scratch and the captured decode-step graph for one model."""

    def __init__(self, model):
        self.model = model
        self._graph = None
        self._graph_key = None
        self._weights = None            # DecoderW used for the rollout (base, or base+LoRA merged)
        self._cached = {}               # rollout shape key -> static buffers + captured decode graph (reused across steps)

    # ------------------------------------------------------------------
    def rollout_weights(self) -> DecoderW:
        """Merged (base + LoRA) weights with the RMSNorm gains folded in -- decode only; the prefill runs the regular
        forward (base weights + LoRA second K segment)."""
        m = self.model
        if getattr(m, "_rollout_dec", None) is None:
            from .lora import build_rollout_weights
            m._rollout_dec = build_rollout_weights(m._dec, m._lora)
        return m._rollout_dec

    @torch.no_grad()
    def generate(self, input_ids, attention_mask, dna_tokenized=None, batch_idx_map=None, *, params: SamplingParams,
                 uniforms: Optional[torch.Tensor] = None, use_graph: bool = True, return_stats: bool = False):
        m = self.model
        W = m._dec                                                          # prefill weights
        Wd = self.rollout_weights()                                         # decode weights (merged + folded)
        cfg = W.cfg
        dev = W.embed.device
        Hq, Hkv, D, d = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim, cfg.hidden_size
        theta = cfg.rope_parameters["rope_theta"] if hasattr(cfg, "rope_parameters") else cfg.rope_theta
        eps = cfg.rms_norm_eps
        input_ids = input_ids.to(dev)
        attention_mask = attention_mask.to(dev)
        B, P = input_ids.shape
        C = params.max_new_tokens

        # ---- one host sync: grouping flags + prompt lengths + the layout check (left-padded rows, one contiguous run of ones ending at
        #      the last column: what `padding_side="left"` produces, grpo_trainer.py:556-565; KV placement and the first-token logits
        #      below assume it, so anything else is refused instead of generating from a pad position)
        eq = detect_group_size(input_ids, dna_tokenized, batch_idx_map)
        lens = attention_mask.sum(dim=1)
        m01 = attention_mask != 0
        layout_ok = (m01[:, -1].all() & (m01[:, 1:] >= m01[:, :-1]).all()).long().reshape(1) if P > 1 else m01[:, -1].all().long().reshape(1)
        host = torch.cat([eq.long(), lens, layout_ok]).tolist()
        if not host[2 * B]:
            raise ValueError("generate(): attention_mask must be left-padded (each row: zeros, then ones up to the last column)")
        G = group_size_from_flags([bool(x) for x in host[:B]])
        U = B // G
        plen = [int(x) for x in host[B:2 * B]][::G]                      # prompt length of each unique row

        # ---- encode + prefill the unique prompts only
        uid = torch.arange(0, B, G, device=dev)
        u_ids, u_mask = input_ids[uid], attention_mask[uid]
        if dna_tokenized is not None and batch_idx_map:
            keep = [i for i, b in enumerate(batch_idx_map) if b % G == 0]
            kt = torch.tensor(keep, device=dev)
            u_dna = {k: v.to(dev)[kt] for k, v in dna_tokenized.items() if k in ("input_ids", "attention_mask")}
            u_map = [batch_idx_map[i] // G for i in keep]
        else:
            u_dna, u_map = None, []
        emb = m.merged_embeddings(u_ids, u_dna, u_map)
        ks, ke = engine.mask_window(u_mask)
        pos = engine.generate_positions(u_mask)

        # ---- page plan: full prompt pages are shared by the group, the tail page + generated tokens are private
        plan = plan_pages(plen, G, C)
        n_shared, max_pages, n_pages = plan["n_shared"], plan["max_pages"], plan["n_pages"]
        table = torch.tensor(plan["table"], dtype=torch.int32)
        prefill_pages = [torch.tensor(p, dtype=torch.int32) for p in plan["prefill_pages"]]
        nl = len(W.layers)
        # Static buffers + the captured decode graph are cached per rollout shape: a training run replays the same graph every
        # step (no per-step capture, no graph-pool / allocator churn -- that churn showed up as multi-second host stalls).
        key = (B, G, tuple(plen), C, n_shared, max_pages, n_pages, params.do_sample, params.temperature, params.top_k, params.top_p,
               params.eos_token_id, params.pad_token_id, id(Wd), bool(use_graph), os.environ.get("BR_DECODE_CHAIN", "0"), os.environ.get("BR_L2PF", "0"), os.environ.get("BR_STREAM_GATE", "0"),
               os.environ.get("BR_ATTN_SS", ""), os.environ.get("BR_ATTN_SP", ""))
        St = self._cached.get(key)
        hit = St is not None
        if not hit:
            if len(self._cached) >= 4:
                self._cached.clear()
            St = SimpleNamespace()
            St.table = table.to(dev)
            # zero-filled once: the unwritten slots of a partially filled page are multiplied by exact-zero probabilities in the P V
            # product, which is only harmless if they hold finite numbers (recycled allocator memory may hold NaN bit patterns)
            St.kc = torch.zeros(nl, n_pages, Hkv, PAGE, D, device=dev, dtype=torch.bfloat16)
            St.vc = torch.zeros_like(St.kc)
            St.pp_dev = [p.to(dev) for p in prefill_pages]
        table, kc, vc, pp_dev = St.table, St.kc, St.vc, St.pp_dev

        def kv_sink(li, qkv):
            for u in range(U):
                first = u * P + (P - plen[u])                             # first real token of the left-padded row
                ops.kv_write_pages(qkv[first:], plen[u], Hq, Hkv, D, pp_dev[u], kc[li], vc[li])

        hidden = engine.decoder_forward(W, emb, U, P, pos, ks, ke, kv_sink=kv_sink, lora=m._lora.w if m._lora is not None else None)
        # replicate each group's partially filled tail page to the other G-1 rows
        if plan["tail_copies"]:
            s_t = torch.tensor([a for a, _ in plan["tail_copies"]], device=dev)
            d_t = torch.tensor([b for _, b in plan["tail_copies"]], device=dev)
            kc[:, d_t] = kc[:, s_t]; vc[:, d_t] = vc[:, s_t]

        # ---- decode state
        R = B
        eos = params.eos_token_id if params.eos_token_id is not None else -1
        pad_fill = params.pad_token_id if params.pad_token_id is not None else 0
        if params.do_sample:
            if uniforms is None:
                uniforms = torch.rand(C, R, device=dev, dtype=torch.float32)
            uniforms = uniforms.to(dev).float().contiguous()
            assert uniforms.shape == (C, R)
        cur0 = torch.tensor([plen[r // G] for r in range(R)], device=dev, dtype=torch.int32)
        if not hit:
            St.tokens = torch.full((R, C), pad_fill, device=dev, dtype=torch.int64)
            St.next_ids = torch.zeros(R, device=dev, dtype=torch.int64)
            St.finished = torch.zeros(R, device=dev, dtype=torch.int32)
            St.step = torch.zeros(1, device=dev, dtype=torch.int32)
            St.cur_len = cur0.clone()
            St.uniforms = uniforms.clone() if params.do_sample else None
            St.scratch = ops.skinny_scratch(max(cfg.vocab_size, 2 * cfg.intermediate_size), dev)
            # two KV tiles per work item where the co-residency cap allows it: both are fetched before the dependency wait, so the tile loop
            # never waits on DRAM (measured: (14, 3) splits 2069 tok/s vs (8, 2) 2028 at 28 shared pages)
            ss_default = min(16, max(8, (n_shared + 1) // 2))
            splits_shared = min(int(os.environ.get("BR_ATTN_SS", ss_default)), n_shared) if n_shared > 0 else 0
            splits_private = int(os.environ.get("BR_ATTN_SP", 3)) if n_shared > 0 else 8
            per_sm = 2 if os.environ.get("BR_DECODE_ATTN_TC5", "0") not in ("0", "") else 3     # CTAs of the fused attention an SM can hold
            cap = per_sm * torch.cuda.get_device_properties(dev).multi_processor_count      # the fused kernel's merger items need co-residency
            n_items = lambda ss, sp: (R // G) * Hkv * ss + R * Hkv * sp
            while n_items(splits_shared, splits_private) > cap and (splits_shared > 1 or splits_private > 1):
                if splits_private > 1 and (splits_private >= splits_shared or splits_shared <= 1):
                    splits_private //= 2
                else:
                    splits_shared = max(1, splits_shared // 2)
            if G * (Hq // Hkv) > 32:
                raise NotImplementedError("fused decode attention handles G * Hq/Hkv <= 32 query vectors per kv head")
            St.splits = (splits_shared, splits_private)
            St.ws = ops.decode_fused_workspace(R, Hq, Hkv, D, splits_shared + splits_private, dev)
            St.attn_out = torch.empty(R, Hq * D, device=dev, dtype=torch.bfloat16)
            St.rope = ops.rope_table(max(plen) + C + 1, D, theta, dev)
            St.h = torch.empty(R, d, device=dev, dtype=torch.bfloat16)
            n_part_ = ((d + 127) // 128) * 4                                    # partial sum-of-squares rows a d-wide GEMM emits
            St.ssq_a = torch.zeros(n_part_, 32, device=dev, dtype=torch.float32)    # sum x^2 of the residual stream entering attention
            St.ssq_b = torch.zeros(n_part_, 32, device=dev, dtype=torch.float32)    # ... entering the MLP (see br_skinny_gemm_ex)
            St.ssq_e = torch.zeros(1, 32, device=dev, dtype=torch.float32)          # ... of the embedding row (first layer)
            St.samp_ws = ops.sample_workspace(R, cfg.vocab_size, dev)
            St.gate = torch.zeros(1, device=dev, dtype=torch.int32)                  # stream-gate arrivals (see br_stream_gate)
            St.graph = None
        else:
            St.gate.zero_()
            St.tokens.fill_(pad_fill); St.finished.zero_(); St.step.zero_(); St.cur_len.copy_(cur0)
            if params.do_sample:
                St.uniforms.copy_(uniforms)
        tokens, next_ids, finished, step, cur_len = St.tokens, St.next_ids, St.finished, St.step, St.cur_len
        uniforms = St.uniforms
        scratch, ws, attn_out, rope, h = St.scratch, St.ws, St.attn_out, St.rope, St.h
        ssq_a, ssq_b, ssq_e, samp_ws = St.ssq_a, St.ssq_b, St.ssq_e, St.samp_ws
        splits_shared, splits_private = St.splits
        n_part = ((d + 127) // 128) * 4

        def sample(logits):
            ops.sample_next(logits, workspace=samp_ws, temperature=params.temperature, top_k=params.top_k, top_p=params.top_p, do_sample=params.do_sample,
                            uniforms=uniforms if params.do_sample else None, step=step, max_steps=C, eos_id=eos,
                            pad_id=params.pad_token_id if params.pad_token_id is not None else 0, finished=finished, tokens=tokens,
                            next_ids=next_ids)

        # ---- first token from the prefill's last position (row u replicated G times)
        last_rows = torch.tensor([u * P + P - 1 for u in range(U) for _ in range(G)], device=dev, dtype=torch.int32)
        h_last = ops.gather_rows(hidden, last_rows)
        logits = ops.skinny_gemm(h_last, W.lm_head, scratch, mode=3)             # prefill output is already final-normed
        sample(logits)
        step += 1                                                         # cur_len stays: the first generated token sits at index plen

        F = cfg.intermediate_size
        if not hit:
            St.b_qkv = torch.empty(R, (Hq + 2 * Hkv) * D, device=dev, dtype=torch.bfloat16)
            St.b_x2 = torch.empty(R, d, device=dev, dtype=torch.bfloat16)
            St.b_act = torch.empty(R, F, device=dev, dtype=torch.bfloat16)
            St.b_logits = torch.empty(R, cfg.vocab_size, device=dev, dtype=torch.float32)
        b_qkv, b_x2, b_act, b_logits = St.b_qkv, St.b_x2, St.b_act, St.b_logits

        use_chain = os.environ.get("BR_DECODE_CHAIN", "0") == "1"

        def decode_step_chain():
            # 2 launches per layer: fused attention, then ONE persistent kernel running o_proj (+res) -> gate/up (folded ln2, SwiGLU)
            # -> down_proj (+res) -> the NEXT layer's qkv projection (folded ln1) / the lm_head, with grid barriers inside.
            # RMSNorm never launches: its statistics ride along in the GEMM epilogues (sum x^2 partials).
            ops.embed_gather_sumsq(next_ids, Wd.embed, h, ssq_e)
            L0 = Wd.layers[0]
            ops.skinny_gemm(h, L0.w_qkv, scratch, out=b_qkv, sumsq_in=ssq_e, sumsq_in_n=1, eps=eps)
            nl_ = len(Wd.layers)
            for li, Lw in enumerate(Wd.layers):
                ops.decode_attn_fused(b_qkv, Lw.q_norm, Lw.k_norm, kc[li], vc[li], table, cur_len, G, Hq, Hkv, D, n_shared, splits_shared,
                                      splits_private, theta, eps, ws, attn_out, rope=rope)
                nxt = (dict(x=h, w=Wd.layers[li + 1].w_qkv, out=b_qkv, mode=0, sumsq_in=ssq_a, sumsq_in_n=n_part) if li + 1 < nl_ else
                       dict(x=h, w=Wd.lm_head, out=b_logits, mode=3, sumsq_in=ssq_a, sumsq_in_n=n_part))
                ops.skinny_chain([dict(x=attn_out, w=Lw.w_o, out=b_x2, mode=1, residual=h, sumsq_out=ssq_b),
                                  dict(x=b_x2, w=Lw.w_gu, out=b_act, mode=2, sumsq_in=ssq_b, sumsq_in_n=n_part),
                                  dict(x=b_act, w=Lw.w_down, out=h, mode=1, residual=b_x2, sumsq_out=ssq_a),
                                  nxt], R, scratch, eps=eps)
            sample(b_logits)
            ops.decode_advance(step, cur_len)

        # L2 staging plan (BR_L2PF=<fraction>, 0 disables): each launch pulls weight tiles of a LATER GEMM into L2 while HBM would idle
        # under this launch's dependency waits / reductions.  Of every chunk a consumer CTA streams, tiles [0, 6) arrive through its own
        # PDL pre-wait ring; a fraction of the rest is staged by the launches before it.
        pf_frac = float(os.environ.get("BR_L2PF", "0"))

        def stage(w, lo_frac, hi_frac):
            if pf_frac <= 0.0:
                return None
            ch = ops.skinny_chunk_units(w)
            span = max(0, ch - 6) * pf_frac
            return (w, 6 + int(span * lo_frac), 6 + int(span * hi_frac))

        # Stream gate (BR_STREAM_GATE=1): a GEMM that becomes resident while its predecessor GEMM is still streaming starts its early weight
        # loads only when the predecessor's weights are on chip (see br_stream_gate).  `St.gate` counts arrivals; the targets are
        # cumulative counts within a token step (the graph is replayed per step, the step counter supplies the epoch).
        use_gate = os.environ.get("BR_STREAM_GATE", "0") not in ("0", "")
        gate_plan, gate_total = {}, 0
        if use_gate:
            gate_plan, gate_total = stream_gate_plan([{nm_: ops.skinny_grid(getattr(Lw_, nm_)) for nm_ in ("w_qkv", "w_o", "w_gu", "w_down")}
                                                      for Lw_ in Wd.layers], ops.skinny_grid(Wd.lm_head))

        def gate(key):
            if not use_gate:
                return None
            return dict(counter=St.gate, epoch=step, epoch_base=1, per_step=gate_total, wait=gate_plan[key], signal=True)

        def decode_step_5():
            # 5 launches per layer: qkv GEMM (folded ln1), fused attention, o_proj (+res, sum x^2), gate/up GEMM (folded ln2, SwiGLU),
            # down_proj (+res, sum x^2); RMSNorm never launches in the decode loop.
            ops.embed_gather_sumsq(next_ids, Wd.embed, h, ssq_e)
            nl_ = len(Wd.layers)
            for li, Lw in enumerate(Wd.layers):
                nxt_w = Wd.layers[li + 1].w_qkv if li + 1 < nl_ else Wd.lm_head
                ops.skinny_gemm(h, Lw.w_qkv, scratch, out=b_qkv, sumsq_in=ssq_e if li == 0 else ssq_a, sumsq_in_n=1 if li == 0 else n_part, eps=eps,
                                prefetch=stage(Lw.w_o, 0.0, 1.0), gate=gate((li, "w_qkv")))
                ops.decode_attn_fused(b_qkv, Lw.q_norm, Lw.k_norm, kc[li], vc[li], table, cur_len, G, Hq, Hkv, D, n_shared, splits_shared,
                                      splits_private, theta, eps, ws, attn_out, rope=rope, prefetch=stage(Lw.w_gu, 0.0, 0.65))
                ops.skinny_gemm(attn_out, Lw.w_o, scratch, mode=1, residual=h, out=b_x2, sumsq_out=ssq_b, prefetch=stage(Lw.w_gu, 0.65, 1.0),
                                gate=gate((li, "w_o")))
                ops.skinny_gemm(b_x2, Lw.w_gu, scratch, mode=2, out=b_act, sumsq_in=ssq_b, sumsq_in_n=n_part, eps=eps, prefetch=stage(Lw.w_down, 0.0, 1.0),
                                gate=gate((li, "w_gu")))
                ops.skinny_gemm(b_act, Lw.w_down, scratch, mode=1, residual=b_x2, out=h, sumsq_out=ssq_a,
                                prefetch=stage(nxt_w, 0.0, 1.0) if li + 1 < nl_ else (stage(nxt_w, 0.0, 0.1) if pf_frac > 0 else None),
                                gate=gate((li, "w_down")))
            ops.skinny_gemm(h, Wd.lm_head, scratch, mode=3, out=b_logits, sumsq_in=ssq_a, sumsq_in_n=n_part, eps=eps, gate=gate("lm_head"))
            sample(b_logits)
            ops.decode_advance(step, cur_len)

        decode_step = decode_step_chain if use_chain else decode_step_5

        n_steps = C - 1
        if not hit:
            St.decode_step = decode_step
            St.per_replay = 0
            if use_graph and n_steps > 2:
                # warm up once on a side stream (allocator + lazy func attributes), then capture one decode step
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                state = [t.clone() for t in (tokens, next_ids, finished, step, cur_len)]
                with torch.cuda.stream(s):
                    decode_step()
                torch.cuda.current_stream().wait_stream(s)
                for t, v in zip((tokens, next_ids, finished, step, cur_len), state):
                    t.copy_(v)                                                # the warm-up step is replayed for real below
                St.gate.zero_()
                St.graph = torch.cuda.CUDAGraph()
                n0 = ops.LAUNCHES[0]
                with torch.cuda.graph(St.graph):
                    decode_step()
                St.per_replay = ops.LAUNCHES[0] - n0
                ops.LAUNCHES[0] = n0
                for t, v in zip((tokens, next_ids, finished, step, cur_len), state):
                    t.copy_(v)
            self._cached[key] = St
        graph, per_replay, decode_step = St.graph, St.per_replay, St.decode_step
        done_steps = 0
        check_every = 16
        while done_steps < n_steps:
            chunk = min(check_every, n_steps - done_steps) if eos >= 0 else n_steps - done_steps
            for _ in range(chunk):
                if graph is not None:
                    graph.replay()
                    ops.LAUNCHES[0] += per_replay
                else:
                    decode_step()
            done_steps += chunk
            if eos >= 0 and done_steps < n_steps and bool(finished.min().item() == 1):
                break
        out = tokens.clone()                                                 # the static buffer is reused by the next rollout
        if eos >= 0:
            # HF stops as soon as every row has finished: trim to the longest row (eos position inclusive)
            is_eos = out == eos
            first = torch.where(is_eos.any(1), is_eos.int().argmax(1) + 1, torch.full((R,), min(done_steps + 1, C), device=dev))
            out = out[:, : int(first.max().item())]
        if return_stats:
            return out, dict(G=G, unique_prompts=U, n_shared_pages=n_shared, pages=n_pages, graph=graph is not None)
        return out
