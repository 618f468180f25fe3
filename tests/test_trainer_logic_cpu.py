"""Host logic of DNALLMGRPOTrainer.compute_loss on CPU: row-chunked (micro_rows) forward/backward must reproduce the full-batch
loss and gradient of grpo_trainer.py:786-812 exactly (the row-mean of row-means is separable over row chunks).  The CUDA ops
are replaced by the oracle's torch math; only the trainer's own control flow runs."""
import collections
import types

import pytest
import torch

from oracle import grpo as og


def _fake_trainer(beta, mu, micro_rows, ga=1):
    from bioreason_b200.trainer.grpo_trainer import DNALLMGRPOTrainer
    t = object.__new__(DNALLMGRPOTrainer)
    t.args = types.SimpleNamespace(micro_rows=micro_rows, gradient_accumulation_steps=ga)
    t.beta, t.num_iterations, t.epsilon_low, t.epsilon_high = beta, mu, 0.2, 0.2
    from bioreason_b200.trainer.grpo_trainer import TrainerState
    t.state = TrainerState()
    t.global_step, t._step = 0, 0
    t._buffered_inputs = [None] * ga
    t._metrics = collections.defaultdict(list)
    t.timings = collections.defaultdict(float)
    t._ev = []
    t._mark = lambda phase: __import__("contextlib").nullcontext()
    return t


@pytest.mark.parametrize("micro_rows", [None, 1, 3, 4])
@pytest.mark.parametrize("beta,mu", [(0.04, 1), (0.04, 2), (0.0, 2)])
def test_compute_loss_row_chunks_match_full_batch(monkeypatch, micro_rows, beta, mu):
    from bioreason_b200 import ops, training
    from bioreason_b200.trainer import grpo_trainer as gt
    B, P, C = 8, 5, 12
    g = torch.Generator().manual_seed(3)
    lp_full = -torch.rand(B, C, generator=g) * 3
    old = lp_full + torch.randn(B, C, generator=g) * 0.3 if mu > 1 else None
    ref = lp_full + torch.randn(B, C, generator=g) * 0.2 if beta > 0 else None
    adv = torch.randn(B, generator=g)
    cmask = (torch.arange(C)[None, :] < torch.randint(2, C + 1, (B, 1), generator=g)).int()
    row_of = {}                                                        # chunk -> rows, recovered from the ids we pass through

    def fake_policy_forward(model, ids, mask, dna, idx_map, keep_last, save=True, lora="policy", targets=None):
        rows = ids[:, 0].tolist()                                      # row id smuggled in the first prompt token
        ctx = types.SimpleNamespace(rows=rows)
        return lp_full[rows].clone(), ctx

    def fake_loss_raw(lp, old_lp, ref_lp, adv_, mask_, beta_, lo, hi, want_grad=True):
        x = lp.clone().requires_grad_(True)
        loss, kl, clip = og.grpo_loss(x, old_lp, ref_lp, adv_, mask_, beta_, lo, hi)
        loss.backward()
        return torch.stack([loss.detach(), kl.detach() if kl is not None else torch.zeros(()), clip.detach()]), x.grad

    got_grad = torch.zeros(B, C)

    def fake_backward(model, ctx, dlp, on_layer_done=None):
        got_grad[ctx.rows] += dlp

    monkeypatch.setattr(training, "policy_forward", fake_policy_forward)
    monkeypatch.setattr(training, "policy_backward", fake_backward)
    monkeypatch.setattr(ops, "grpo_loss_raw", fake_loss_raw)
    t = _fake_trainer(beta, mu, micro_rows)
    prompt_ids = torch.zeros(B, P, dtype=torch.long); prompt_ids[:, 0] = torch.arange(B)
    inputs = dict(prompt_ids=prompt_ids, prompt_mask=torch.ones(B, P, dtype=torch.long), completion_ids=torch.zeros(B, C, dtype=torch.long),
                  completion_mask=cmask, old_per_token_logps=old, ref_per_token_logps=ref, advantages=adv,
                  multimodal_inputs=dict(dna_tokenized=None, batch_idx_map=[]))
    loss = gt.DNALLMGRPOTrainer.compute_loss(t, None, inputs)
    x = lp_full.clone().requires_grad_(True)
    want, kl, clip = og.grpo_loss(x, old, ref, adv, cmask, beta, 0.2, 0.2)
    want.backward()
    assert abs(loss.item() - want.item()) < 1e-6
    torch.testing.assert_close(got_grad, x.grad, rtol=1e-5, atol=1e-8)
    if beta > 0:
        assert abs(float(t._metrics["kl"][0]) - kl.item()) < 1e-6
    with pytest.raises(ValueError, match="does not support returning outputs"):
        gt.DNALLMGRPOTrainer.compute_loss(t, None, inputs, return_outputs=True)


def test_slice_mm_follows_batch_idx_map():
    from bioreason_b200.trainer.grpo_trainer import _slice_mm
    dna = dict(input_ids=torch.arange(12).view(6, 2), attention_mask=torch.ones(6, 2, dtype=torch.long))
    mm = dict(dna_tokenized=dna, batch_idx_map=[0, 0, 1, 2, 2, 3])
    out = _slice_mm(mm, 1, 3)
    assert out["batch_idx_map"] == [0, 1, 1] and torch.equal(out["dna_tokenized"]["input_ids"], dna["input_ids"][2:5])
    assert _slice_mm(dict(dna_tokenized=None, batch_idx_map=[]), 0, 2) == dict(dna_tokenized=None, batch_idx_map=[])
