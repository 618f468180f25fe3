"""Synthetic token-id batches in the DLProcessor layout (processing_dl.py:87-132,185-218) for bench.py / smoke().

Product-side copy of the generator (the oracle package is test infrastructure and must not be imported by the timed
arm); tests assert both generators emit identical batches."""
from typing import Dict, Optional

import torch


def synth_batch(text_cfg, dna_cfg, *, batch: int, n_seq: int, dna_len, text_len, seed: int = 1234, pad_to: Optional[int] = None,
                same_prompt: bool = False) -> Dict:
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
            ids = torch.cat([torch.tensor([dna_cfg.cls_token_id]), torch.randint(4, dna_cfg.vocab_size, (n - 1,), generator=g)])
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
    for b, r in enumerate(rows):
        input_ids[b, L - r.numel():] = r
        attn[b, L - r.numel():] = 1
    out = dict(input_ids=input_ids, attention_mask=attn, batch_idx_map=idx_map)
    out["dna_tokenized"] = dict(input_ids=torch.stack(dna_ids), attention_mask=torch.stack(dna_mask)) if n_seq else None
    return out
