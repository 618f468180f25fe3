"""CPU restatement of `bioreason/models/dna_llm.py` on top of the installed HF classes.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every function cites the reference lines it
follows; paths are relative to /root/reference.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F
from transformers import EsmForMaskedLM, Qwen3ForCausalLM
from transformers.models.esm import modeling_esm


# --------------------------------------------------------------------------------------
# NT-v2 deltas on the HF ESM skeleton (HF esm/modeling_esm.py:406-427 are GELU + biased).
# Restated from the public NT-v2 model file: dense(hidden -> 2*ffn, bias=add_bias_fc),
# x1, x2 = split; SiLU(x1) * x2; output dense(ffn -> hidden, bias=add_bias_fc).
# --------------------------------------------------------------------------------------
class NTv2Intermediate(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, 2 * config.intermediate_size,
                               bias=getattr(config, "add_bias_fc", False))

    def forward(self, hidden_states):
        hidden_states = self.dense(hidden_states)
        x1, x2 = hidden_states.split(hidden_states.size(-1) // 2, -1)
        return F.silu(x1) * x2


class NTv2Output(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.intermediate_size, config.hidden_size,
                               bias=getattr(config, "add_bias_fc", False))

    def forward(self, hidden_states, input_tensor):
        return self.dense(hidden_states) + input_tensor


def build_dna_model(cfg, seed: int = 1234) -> EsmForMaskedLM:
    """`AutoModelForMaskedLM.from_pretrained(...)` stand-in (dna_llm.py:79-83): seeded random init."""
    torch.manual_seed(seed)
    model = EsmForMaskedLM(cfg)
    if getattr(cfg, "gated_mlp", False):
        g = torch.Generator().manual_seed(seed + 1)
        for layer in model.esm.encoder.layer:
            layer.intermediate = NTv2Intermediate(cfg)
            layer.output = NTv2Output(cfg)
            for lin in (layer.intermediate.dense, layer.output.dense):
                with torch.no_grad():
                    lin.weight.copy_(torch.randn(lin.weight.shape, generator=g) * 0.02)
                    if lin.bias is not None:
                        lin.bias.zero_()
    # make LayerNorm affine / biases non-trivial so a parity test cannot pass by accident
    g = torch.Generator().manual_seed(seed + 2)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("LayerNorm.weight") or n.endswith("layer_norm.weight") or n.endswith("emb_layer_norm_after.weight"):
                p.add_(torch.randn(p.shape, generator=g) * 0.05)
            elif n.endswith(".bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
    return model.eval()


def build_text_model(cfg, seed: int = 1234) -> Qwen3ForCausalLM:
    """`AutoModelForCausalLM.from_pretrained(...)` stand-in (dna_llm.py:64-66): seeded random init."""
    torch.manual_seed(seed)
    model = Qwen3ForCausalLM(cfg)
    g = torch.Generator().manual_seed(seed + 3)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "norm" in n:  # RMSNorm gains (incl. q_norm/k_norm): perturb away from 1
                p.add_(torch.randn(p.shape, generator=g) * 0.05)
    return model.eval()


def round_to_bf16_(module: nn.Module) -> nn.Module:
    """Keep fp32 storage but make every parameter bf16-representable (the "bf16-storage" regime)."""
    with torch.no_grad():
        for p in module.parameters():
            p.copy_(p.to(torch.bfloat16).to(p.dtype))
    return module


# --------------------------------------------------------------------------------------
# DNALLMModel restated
# --------------------------------------------------------------------------------------
class OracleDNALLM(nn.Module):
    """Restates DNALLMModel (dna_llm.py:18-305) minus tokenizers/processor (CPU string work)."""

    def __init__(self, text_model, dna_model, dna_token_id: int, proj_seed: int = 1234):
        super().__init__()
        self.text_model = text_model
        self.dna_model = dna_model
        self.text_config = text_model.config
        self.dna_config = dna_model.config
        self.text_hidden_size = self.text_config.hidden_size          # dna_llm.py:93
        self.dna_hidden_size = self.dna_config.hidden_size            # dna_llm.py:94
        torch.manual_seed(proj_seed)
        self.dna_projection = nn.Linear(self.dna_hidden_size, self.text_hidden_size)  # dna_llm.py:97
        self.dna_token_id = dna_token_id
        self.dna_is_evo2 = False

    # dna_llm.py:103-179
    def process_dna_embeddings(self, dna_tokenized, batch_idx_map, batch_size):
        with torch.no_grad():                                          # :121 (encoder never gets grad)
            outputs = self.dna_model(
                input_ids=dna_tokenized["input_ids"],
                attention_mask=dna_tokenized["attention_mask"],
                output_hidden_states=True,
            )                                                          # :150-154
            hidden_states = outputs.hidden_states[-1]                  # :156
        hidden_states = hidden_states.to(device=self.dna_projection.weight.device,
                                         dtype=self.dna_projection.weight.dtype)   # :159
        projected_states = self.dna_projection(hidden_states)          # :160
        result = [[] for _ in range(batch_size)]
        for seq_idx, batch_idx in enumerate(batch_idx_map):            # :166-170
            valid_length = dna_tokenized["attention_mask"][seq_idx].sum().item()
            result[batch_idx].append(projected_states[seq_idx, :valid_length])
        for i in range(batch_size):                                    # :173-177
            result[i] = torch.cat(result[i], dim=0) if result[i] else torch.zeros((0, self.text_hidden_size))
        return result

    def _merged_embeds(self, input_ids, dna_tokenized, batch_idx_map):
        batch_size = input_ids.shape[0]
        text_inputs_embeds = self.text_model.get_input_embeddings()(input_ids)     # :211
        if dna_tokenized is not None and batch_idx_map:
            batch_dna_embeds = self.process_dna_embeddings(dna_tokenized, batch_idx_map, batch_size)
            mask = input_ids == self.dna_token_id                      # :216
            n_dna_tokens = mask.sum().item()
            dna_embeds_flat = torch.cat(batch_dna_embeds, dim=0)
            n_dna_features = dna_embeds_flat.shape[0]
            if n_dna_features != n_dna_tokens:                         # :222-225
                raise ValueError(
                    f"DNA features and DNA tokens do not match: features {n_dna_features}, tokens: {n_dna_tokens}")
            dna_embeds_flat = dna_embeds_flat.to(dtype=text_inputs_embeds.dtype)
            text_inputs_embeds = text_inputs_embeds.clone()
            text_inputs_embeds[mask] = dna_embeds_flat                 # :229
        return text_inputs_embeds

    # dna_llm.py:181-244
    def forward(self, input_ids=None, attention_mask=None, dna_tokenized=None, batch_idx_map=None,
                labels=None, **kwargs):
        if input_ids is None or attention_mask is None:
            raise ValueError("Either 'inputs' or 'input_ids'/'attention_mask' must be provided")
        embeds = self._merged_embeds(input_ids, dna_tokenized, batch_idx_map)
        return self.text_model(inputs_embeds=embeds, attention_mask=attention_mask, labels=labels, **kwargs)

    # dna_llm.py:246-305
    def generate(self, input_ids=None, attention_mask=None, dna_tokenized=None, batch_idx_map=None,
                 **generation_kwargs):
        if input_ids is None or attention_mask is None:
            raise ValueError("Either 'inputs' or 'input_ids'/'attention_mask' must be provided")
        embeds = self._merged_embeds(input_ids, dna_tokenized, batch_idx_map)
        with torch.no_grad():
            return self.text_model.generate(inputs_embeds=embeds, attention_mask=attention_mask,
                                            use_cache=True, **generation_kwargs)   # :298-304


def build_oracle(text_cfg, dna_cfg, seed: int = 1234, bf16_weights: bool = True) -> OracleDNALLM:
    text = build_text_model(text_cfg, seed)
    dna = build_dna_model(dna_cfg, seed)
    m = OracleDNALLM(text, dna, dna_token_id=text_cfg.dna_token_ids[1], proj_seed=seed + 7)
    if bf16_weights:
        round_to_bf16_(m)
    return m.eval()


# --------------------------------------------------------------------------------------
# Synthetic inputs in the layout DLProcessor produces (processing_dl.py:87-132,185-218)
# --------------------------------------------------------------------------------------
def synth_batch(text_cfg, dna_cfg, *, batch: int, n_seq: int, dna_len, text_len, seed: int = 1234,
                pad_to: Optional[int] = None, same_prompt: bool = False) -> Dict:
    """Token-id tensors shaped like the processor output.

    dna_len / text_len: int, or list[int] per batch item (ragged -> DNA right-padded with pad id 1
    as the ESM tokenizer does, text LEFT-padded, nucleotide_module.py:142).  Each DNA sequence =
    CLS(3) + ids ~ U[4, vocab).  Text = random ids < first special id with, per DNA sequence, a
    <|dna_start|> <|dna_pad|>*n <|dna_end|> block (n = #non-pad DNA tokens, processing_dl.py:185-193).
    """
    g = torch.Generator().manual_seed(seed)
    start_id, pad_id, end_id = text_cfg.dna_token_ids
    first_special = min(start_id, text_cfg.eos_token_id)
    dl = [dna_len] * batch if isinstance(dna_len, int) else list(dna_len)
    tl = [text_len] * batch if isinstance(text_len, int) else list(text_len)
    s_max = max(dl) if n_seq else 0
    dna_ids, dna_mask, idx_map, rows = [], [], [], []
    for b in range(batch):
        if same_prompt and b > 0:
            rows.append(rows[0].clone())
            for s in range(n_seq):
                dna_ids.append(dna_ids[s].clone()); dna_mask.append(dna_mask[s].clone()); idx_map.append(b)
            continue
        n_txt = tl[b]
        txt = torch.randint(0, first_special, (n_txt,), generator=g)
        pieces, cut = [], [round(n_txt * (i + 1) / (n_seq + 1)) for i in range(n_seq)]
        prev = 0
        for s in range(n_seq):
            n = dl[b]
            ids = torch.cat([torch.tensor([dna_cfg.cls_token_id]),
                             torch.randint(4, dna_cfg.vocab_size, (n - 1,), generator=g)])
            ids = torch.cat([ids, torch.full((s_max - n,), dna_cfg.pad_token_id, dtype=torch.long)])
            dna_ids.append(ids); dna_mask.append((ids != dna_cfg.pad_token_id).long()); idx_map.append(b)
            pieces += [txt[prev:cut[s]], torch.tensor([start_id]), torch.full((n,), pad_id), torch.tensor([end_id])]
            prev = cut[s]
        pieces.append(txt[prev:])
        rows.append(torch.cat(pieces).long())
    L = max(r.numel() for r in rows)
    if pad_to is not None:
        L = max(L, pad_to)
    input_ids = torch.full((batch, L), text_cfg.pad_token_id, dtype=torch.long)
    attn = torch.zeros((batch, L), dtype=torch.long)
    for b, r in enumerate(rows):                       # left padding
        input_ids[b, L - r.numel():] = r
        attn[b, L - r.numel():] = 1
    out = dict(input_ids=input_ids, attention_mask=attn, batch_idx_map=idx_map)
    if n_seq:
        out["dna_tokenized"] = dict(input_ids=torch.stack(dna_ids), attention_mask=torch.stack(dna_mask))
    else:
        out["dna_tokenized"] = None
    return out
