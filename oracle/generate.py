"""Token-by-token restatement of HF sampling with caller-supplied uniforms -- TEST INFRASTRUCTURE (see oracle/__init__.py).

HF `generate(do_sample=True)` draws with torch.multinomial, whose Philox stream cannot be reproduced by a CUDA kernel.
This loop is HF's own pipeline (the same LogitsWarper classes in the order HF builds them, generation/utils.py:1214-1223;
EOS / pad bookkeeping :2796-2797; position_ids = cumsum(mask)-1 :719-721) with only the draw replaced by an inverse-CDF
over ascending token id using `uniforms[step, row]`.  In greedy mode it must equal `model.generate` exactly
(tests/test_oracle_golden.py::test_manual_generate_matches_hf_greedy).
"""
import torch
from transformers.generation.logits_process import TemperatureLogitsWarper, TopKLogitsWarper, TopPLogitsWarper


@torch.no_grad()
def manual_generate(oracle, batch, *, max_new_tokens, do_sample=False, temperature=1.0, top_k=50, top_p=1.0, uniforms=None,
                    eos_token_id=None, pad_token_id=0, return_margins=False):
    embeds = oracle._merged_embeds(batch["input_ids"], batch.get("dna_tokenized"), batch.get("batch_idx_map"))
    dev = embeds.device                                                     # the checker may run on the GPU (fp32) for the real-vocab tests
    mask = batch["attention_mask"].clone().to(dev)
    B = embeds.shape[0]
    warpers = []
    if do_sample:
        if temperature != 1.0:
            warpers.append(TemperatureLogitsWarper(temperature))
        if top_k:
            warpers.append(TopKLogitsWarper(top_k=top_k, min_tokens_to_keep=1))
        if top_p < 1.0:
            warpers.append(TopPLogitsWarper(top_p=top_p, min_tokens_to_keep=1))
    unfinished = torch.ones(B, dtype=torch.long, device=dev)
    out, margins = [], []
    emb_table = oracle.text_model.get_input_embeddings()
    for step in range(max_new_tokens):
        pos = (mask.long().cumsum(-1) - 1).masked_fill(mask == 0, 1)
        logits = oracle.text_model(inputs_embeds=embeds, attention_mask=mask, position_ids=pos).logits[:, -1, :].float()
        top2 = logits.topk(2, dim=-1).values
        margins.append(top2[:, 0] - top2[:, 1])
        if do_sample:
            scores = logits
            for w in warpers:
                scores = w(None, scores)
            probs = torch.softmax(scores, dim=-1)
            cdf = probs.cumsum(-1)
            u = uniforms[step].to(device=dev, dtype=cdf.dtype)[:, None] * cdf[:, -1:]
            nxt = (cdf > u).int().argmax(-1)
        else:
            nxt = logits.argmax(-1)
        if eos_token_id is not None:
            nxt = nxt * unfinished + pad_token_id * (1 - unfinished)
            unfinished = unfinished & (nxt != eos_token_id).long()
        out.append(nxt)
        embeds = torch.cat([embeds, emb_table(nxt)[:, None, :]], dim=1)
        mask = torch.cat([mask, torch.ones(B, 1, dtype=mask.dtype, device=dev)], dim=1)
        if eos_token_id is not None and unfinished.max() == 0:
            break
    ids = torch.stack(out, dim=1)
    return (ids, torch.stack(margins, dim=1)) if return_margins else ids
