"""SFT collate / processor fast path: the two per-batch Python loops either side of the hot path, vectorised.

  * `assistant_span_labels`  -- bioreason/dataset/kegg.py:258-323: labels = input_ids inside every `<|im_start|>assistant\\n ... <|im_end|>`
    span, -100 elsewhere.  The reference scans every position of every row with a `torch.all(window == marker)` per position (two tiny
    tensor ops per token: ~30 k op launches for a batch of 8 x 1.8 k tokens); here it is a handful of whole-batch tensor ops (sliding
    window compare, two running maxima), identical output, on whatever device the ids live on (DataLoader workers have no CUDA, so this
    is deliberately plain device-agnostic tensor code, not a kernel of libbioreason_b200).
  * `expand_dna_placeholders` -- bioreason/models/dl/processing_dl.py:185-193: every `<|dna_pad|>` in the prompt text is repeated once
    per non-pad DNA token of its sequence; the reference pays one `.sum().item()` device sync per sequence, here the counts come from
    one reduction and one host transfer.
  * `qwen_dna_collate_fn`    -- kegg.py:223-333 with the two pieces above; same arguments, same batch keys.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch


def _marker_hits(ids: torch.Tensor, marker: Sequence[int]) -> torch.Tensor:
    """hits[b, p] = ids[b, p : p + len(marker)] == marker  (bool [B, L], False where the window would run off the row)."""
    B, L = ids.shape
    m = len(marker)
    hits = torch.zeros(B, L, dtype=torch.bool, device=ids.device)
    if m == 0 or L < m:
        return hits
    mk = torch.as_tensor(list(marker), dtype=ids.dtype, device=ids.device)
    hits[:, : L - m + 1] = (ids.unfold(1, m, 1) == mk).all(-1)
    return hits


def assistant_span_labels(input_ids: torch.Tensor, start_marker_ids: Sequence[int], end_marker_ids: Sequence[int], pad_token_id: int) -> torch.Tensor:
    """kegg.py:258-323.  A token t is labelled iff the latest assistant-start position s <= t exists and no end marker begins at a
    position e with s < e <= t (sections from earlier starts are subsets of that condition); padding is masked last (kegg.py:323)."""
    B, L = input_ids.shape
    dev = input_ids.device
    ms = len(start_marker_ids)
    pos = torch.arange(L, device=dev).expand(B, L)
    start_at = torch.zeros(B, L, dtype=torch.bool, device=dev)            # start_at[b, t]: a start marker ENDS right before t
    hs = _marker_hits(input_ids, start_marker_ids)
    if ms < L:
        start_at[:, ms:] = hs[:, : L - ms]
    end_at = _marker_hits(input_ids, end_marker_ids)                      # end_at[b, t]: an end marker BEGINS at t
    neg = torch.full((B, L), -1, device=dev, dtype=torch.long)
    last_start = torch.where(start_at, pos, neg).cummax(dim=1).values
    last_end = torch.where(end_at, pos, neg).cummax(dim=1).values
    inside = (last_start >= 0) & ~(last_end > last_start)
    labels = torch.where(inside, input_ids, torch.full_like(input_ids, -100))
    labels[input_ids == pad_token_id] = -100
    return labels


def dna_token_counts(dna_input_ids: torch.Tensor, dna_pad_id: int = 1) -> List[int]:
    """Non-pad DNA tokens per sequence: one reduction + one host transfer (processing_dl.py:188 does `.sum().item()` per sequence)."""
    return (dna_input_ids != dna_pad_id).sum(dim=1).tolist()


def expand_dna_placeholders(texts: List[str], counts: Sequence[int], dna_token: str) -> List[str]:
    """processing_dl.py:185-193: the k-th `dna_token` over the whole batch (row-major) becomes counts[k] copies of itself."""
    out, k = [], 0
    for t in texts:
        parts = t.split(dna_token)
        buf = [parts[0]]
        for p in parts[1:]:
            buf.append(dna_token * int(counts[k])); k += 1
            buf.append(p)
        out.append("".join(buf))
    return out


def qwen_dna_collate_fn(examples: List[Dict], processor, max_length_text: int, max_length_dna: int, return_answer_in_batch: bool = False,
                        apply_chat_template=None) -> Dict:
    """kegg.py:223-333 on the vectorised pieces.  `apply_chat_template(example, processor) -> {"prompt": str}` defaults to trl's
    `maybe_apply_chat_template` when trl is importable (the reference imports it at kegg.py:14)."""
    if apply_chat_template is None:
        from trl.data_utils import maybe_apply_chat_template as apply_chat_template          # noqa: N813
    prompts_text = [apply_chat_template(ex, processor)["prompt"] for ex in examples]
    batch = processor(text=prompts_text, batch_dna_sequences=[ex["dna_sequences"] for ex in examples], return_tensors="pt", padding=True,
                      padding_side="left", add_special_tokens=False, max_length_text=max_length_text, max_length_dna=max_length_dna)
    tok = processor.tokenizer
    batch["labels"] = assistant_span_labels(batch["input_ids"], tok.encode("<|im_start|>assistant\n", add_special_tokens=False),
                                            tok.encode("<|im_end|>", add_special_tokens=False), tok.pad_token_id)
    if return_answer_in_batch:
        batch["answer"] = [ex["answer"].strip() for ex in examples]
    return batch
