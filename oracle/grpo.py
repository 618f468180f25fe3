"""CPU restatement of the tensor math of `bioreason/trainer/grpo_trainer.py`.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Line numbers refer to
/root/reference/bioreason/trainer/grpo_trainer.py.
"""
from __future__ import annotations

import torch


def per_token_logps(model, input_ids, attention_mask, **mm):
    """:510-520 -- logits[:, :-1] -> per-row log_softmax -> gather(input_ids[:, 1:])."""
    logits = model(input_ids=input_ids, attention_mask=attention_mask, **mm).logits
    logits = logits[:, :-1, :]
    input_ids = input_ids[:, 1:]
    out = []
    for logits_row, ids_row in zip(logits, input_ids):
        log_probs = logits_row.log_softmax(dim=-1)
        out.append(torch.gather(log_probs, dim=1, index=ids_row.unsqueeze(1)).squeeze(1))
    return torch.stack(out)


def completion_mask_from_eos(completion_ids, eos_token_id):
    """:605-609 -- mask everything after the first EOS (EOS itself kept)."""
    is_eos = completion_ids == eos_token_id
    eos_idx = torch.full((is_eos.size(0),), is_eos.size(1), dtype=torch.long)
    eos_idx[is_eos.any(dim=1)] = is_eos.int().argmax(dim=1)[is_eos.any(dim=1)]
    seq = torch.arange(is_eos.size(1)).expand(is_eos.size(0), -1)
    return (seq <= eos_idx.unsqueeze(1)).int()


def group_advantages(rewards_per_func, num_generations):
    """:682-692 -- sum over reward funcs, group mean / UNBIASED std, (r-mu)/(sigma+1e-4)."""
    rewards = rewards_per_func.sum(dim=1)
    mean_g = rewards.view(-1, num_generations).mean(dim=1)
    std_g = rewards.view(-1, num_generations).std(dim=1)
    mean_g = mean_g.repeat_interleave(num_generations, dim=0)
    std_g = std_g.repeat_interleave(num_generations, dim=0)
    return (rewards - mean_g) / (std_g + 1e-4)


def grpo_loss(per_token_logps, old_per_token_logps, ref_per_token_logps, advantages, completion_mask,
              beta=0.04, epsilon_low=0.2, epsilon_high=0.2):
    """:786-812 -- clipped-ratio loss + beta * k3-KL, masked per-row mean then batch mean.

    Returns (loss, mean_kl, clip_ratio).  `old_per_token_logps=None` means mu == 1 (:786).
    """
    if old_per_token_logps is None:
        old_per_token_logps = per_token_logps.detach()
    coef_1 = torch.exp(per_token_logps - old_per_token_logps)
    coef_2 = torch.clamp(coef_1, 1 - epsilon_low, 1 + epsilon_high)
    l1 = coef_1 * advantages.unsqueeze(1)
    l2 = coef_2 * advantages.unsqueeze(1)
    per_token_loss = -torch.min(l1, l2)
    mean_kl = None
    if beta > 0:
        d = ref_per_token_logps - per_token_logps
        per_token_kl = torch.exp(d) - d - 1
        per_token_loss = per_token_loss + beta * per_token_kl
        mean_kl = ((per_token_kl * completion_mask).sum(dim=1) / completion_mask.sum(dim=1)).mean()
    loss = ((per_token_loss * completion_mask).sum(dim=1) / completion_mask.sum(dim=1)).mean()
    is_clipped = (l1 < l2).float()
    clip_ratio = (is_clipped * completion_mask).sum() / completion_mask.sum()
    return loss, mean_kl, clip_ratio


def repeat_random_sampler(n, mini_repeat_count, batch_size=1, repeat_count=1, seed=None):
    """:72-119 RepeatRandomSampler.__iter__ restated as a list."""
    g = torch.Generator()
    if seed is not None:
        g.manual_seed(seed)
    idx = torch.randperm(n, generator=g).tolist()
    chunks = [idx[i:i + batch_size] for i in range(0, len(idx), batch_size)]
    chunks = [c for c in chunks if len(c) == batch_size]
    out = []
    for c in chunks:
        for _ in range(repeat_count):
            for i in c:
                out.extend([i] * mini_repeat_count)
    return out
