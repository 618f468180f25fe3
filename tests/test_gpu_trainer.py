"""GRPO trainer mirror on the GPU: rollout -> scoring -> loss -> backward -> optimizer, against the oracle's math."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _token_reward(completion_ids, **kw):
    return (completion_ids % 7 == 0).float().sum(1) - 0.1 * (completion_ids % 5 == 0).float().sum(1)


def _len_reward(completion_ids, completion_mask, **kw):
    return completion_mask.float().sum(1) / completion_ids.shape[1]


def test_grpo_step_tiny():
    from bioreason_b200.configs import text_config, dna_config
    from bioreason_b200.models import DNALLMModel
    from bioreason_b200.trainer import DNALLMGRPOConfig, DNALLMGRPOTrainer
    from oracle.models import build_oracle, synth_batch
    from oracle import grpo as og
    tc, dc = text_config("tiny"), dna_config("tiny")
    oracle = build_oracle(tc, dc, seed=21)
    batch = synth_batch(tc, dc, batch=4, n_seq=2, dna_len=10, text_len=18, seed=14, same_prompt=True)
    m = DNALLMModel.from_oracle(oracle)
    cfg = DNALLMGRPOConfig(num_generations=4, max_completion_length=6, per_device_train_batch_size=4, learning_rate=1e-3, lora_r=16, lora_alpha=32.0)
    tr = DNALLMGRPOTrainer(m, [_token_reward, _len_reward], cfg)
    with pytest.raises(ValueError, match="evenly"):
        DNALLMGRPOTrainer(m, [_token_reward], DNALLMGRPOConfig(num_generations=3, per_device_train_batch_size=4))
    u = torch.rand(6, 4, generator=torch.Generator().manual_seed(0))
    inp = tr._generate_and_score_completions(batch, m, uniforms=u)
    comp, cmask = inp["completion_ids"].cpu(), inp["completion_mask"].cpu()
    assert torch.equal(cmask, og.completion_mask_from_eos(comp, tc.eos_token_id))
    rpf = torch.stack([_token_reward(comp), _len_reward(comp, cmask)], 1)
    torch.testing.assert_close(inp["advantages"].cpu(), og.group_advantages(rpf, 4), rtol=1e-4, atol=1e-5)
    assert inp["old_per_token_logps"] is None                                  # mu == 1 (grpo_trainer.py:617-626)
    ids = torch.cat([batch["input_ids"], comp], 1)
    mask = torch.cat([batch["attention_mask"], cmask.long()], 1)
    mm = dict(dna_tokenized=batch["dna_tokenized"], batch_idx_map=batch["batch_idx_map"])
    with torch.no_grad():
        ref_o = og.per_token_logps(oracle, ids, mask, **mm)[:, -comp.shape[1]:]
    att = cmask.bool()
    assert (inp["ref_per_token_logps"].cpu() - ref_o)[att].abs().max().item() < 0.03
    # loss value with LoRA B == 0: policy == reference -> kl == 0, loss == -mean(adv) * 1 ... compare with the oracle formula
    loss = tr.compute_loss(m, inp, backward=False)
    lo, klo, clo = og.grpo_loss(ref_o, None, ref_o, inp["advantages"].cpu(), cmask, 0.04, 0.2, 0.2)
    assert abs(loss.item() - lo.item()) < 5e-3
    # a real optimizer step moves the adapters and the projector, and the next rollout uses the merged weights
    p0 = [p.detach().clone() for p in m.trainable_parameters()]
    tr._step = 0
    l1 = tr.training_step(inp)
    moved = sum(int(not torch.equal(a, b.detach())) for a, b in zip(p0, m.trainable_parameters()))
    assert moved > len(p0) // 2 and torch.isfinite(l1)
    assert m._rollout_dec is not None
    l2 = tr.training_step(batch)
    assert torch.isfinite(l2)
    met = tr.log_metrics()
    assert {"completion_length", "reward", "reward_std", "kl", "clip_ratio"} <= set(met)
    assert met["kl"] >= 0


def test_synth_generators_agree():
    from bioreason_b200.configs import text_config, dna_config
    from bioreason_b200.synth import synth_batch as a
    from oracle.models import synth_batch as b
    tc, dc = text_config("tiny"), dna_config("tiny")
    x = a(tc, dc, batch=3, n_seq=2, dna_len=[9, 7, 9], text_len=[20, 14, 11], seed=5)
    y = b(tc, dc, batch=3, n_seq=2, dna_len=[9, 7, 9], text_len=[20, 14, 11], seed=5)
    assert torch.equal(x["input_ids"], y["input_ids"]) and torch.equal(x["dna_tokenized"]["input_ids"], y["dna_tokenized"]["input_ids"])
    assert x["batch_idx_map"] == y["batch_idx_map"]


def test_generate_before_enable_lora_keeps_base_weights_frozen():
    """rollout weights built before the adapters exist alias the frozen base w_o / w_down; the per-step merge must
    never write into them (zero-shot eval -> trainer construction -> optimizer steps)."""
    from bioreason_b200.configs import text_config, dna_config
    from bioreason_b200.models import DNALLMModel
    from bioreason_b200.trainer import DNALLMGRPOConfig, DNALLMGRPOTrainer
    from oracle.models import build_oracle, synth_batch
    tc, dc = text_config("tiny"), dna_config("tiny")
    m = DNALLMModel.from_oracle(build_oracle(tc, dc, seed=5))
    batch = synth_batch(tc, dc, batch=4, n_seq=2, dna_len=10, text_len=18, seed=3, same_prompt=True)
    m.generate(batch["input_ids"], batch["attention_mask"], batch["dna_tokenized"], batch["batch_idx_map"], max_new_tokens=4, do_sample=False)
    assert m._rollout_dec is not None
    base = [(L.w_o.clone(), L.w_down.clone(), L.w_qkv.clone(), L.w_gu.clone()) for L in m._dec.layers]
    cfg = DNALLMGRPOConfig(num_generations=4, max_completion_length=4, per_device_train_batch_size=4, learning_rate=5e-2, lora_r=16, lora_alpha=32.0)
    tr = DNALLMGRPOTrainer(m, [_token_reward], cfg)                       # enable_lora + sync_adapters(rollout=True) inside
    for _ in range(2):
        tr.training_step(batch)
    for L, (wo, wd, wq, wg) in zip(m._dec.layers, base):
        assert torch.equal(L.w_o, wo) and torch.equal(L.w_down, wd) and torch.equal(L.w_qkv, wq) and torch.equal(L.w_gu, wg)
    for L, Lr in zip(m._dec.layers, m._rollout_dec.layers):
        assert Lr.w_o.data_ptr() != L.w_o.data_ptr() and Lr.w_down.data_ptr() != L.w_down.data_ptr()
        assert not torch.equal(Lr.w_o, L.w_o)                             # adapters moved -> merged weights differ from the base
