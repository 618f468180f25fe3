"""Host logic of the reward plumbing (bioreason_b200/trainer/rewards.py) against the reference protocol of
bioreason/trainer/grpo_trainer.py:640-676: decoded completions, conversational wrapping, dataset columns, token-level opt-in."""
import pytest
import torch

from bioreason_b200.trainer import rewards as rw


class FakeTok:
    eos_token_id, pad_token_id = 9, 9

    def batch_decode(self, ids, skip_special_tokens=False):
        out = []
        for row in ids.tolist():
            toks = [t for t in row if not (skip_special_tokens and t == self.eos_token_id)]
            out.append(" ".join(f"w{t}" for t in toks))
        return out


def ref_style_format_reward(completions, **kwargs):                      # reason.py-style: conversational completions
    return [1.0 if c[0]["content"].startswith("w1") else 0.0 for c in completions]


def ref_style_correctness(prompts, completions, answer, **kwargs):        # uses a dataset column, like reason.py's correctness reward
    assert len(prompts) == len(completions) == len(answer)
    return [2.0 if a in c[0]["content"] else -1.0 for c, a in zip(completions, answer)]


def token_len_reward(completion_ids, completion_mask=None, **kw):         # opt-in fast path: device tensors
    return completion_mask.float().sum(1)


def test_protocol_detection():
    assert rw.wants_token_protocol(token_len_reward)
    assert not rw.wants_token_protocol(ref_style_format_reward)
    assert not rw.wants_token_protocol(ref_style_correctness)
    assert not rw.wants_token_protocol(lambda **kw: 0)


def test_reference_protocol_matches_reference_semantics():
    ids = torch.tensor([[1, 2, 9, 9], [3, 4, 5, 9], [1, 7, 7, 7]])
    mask = torch.tensor([[1, 1, 1, 0], [1, 1, 1, 1], [1, 1, 1, 1]])
    examples = [dict(prompt=[{"role": "user", "content": f"q{i}"}], answer=a, dna_sequences=["ACGT"]) for i, a in enumerate(("w2", "w9", "w7"))]
    out = rw.score([ref_style_format_reward, ref_style_correctness, token_len_reward], examples=examples, prompts=None, completion_ids=ids,
                   completion_mask=mask, prompt_ids=torch.zeros(3, 2, dtype=torch.long), processing_class=FakeTok())
    assert out.tolist() == [[1.0, 2.0, 3.0], [0.0, -1.0, 4.0], [1.0, 2.0, 4.0]]
    # the columns every non prompt/completion key becomes (grpo_trainer.py:664-670)
    assert set(rw.reward_columns(examples)) == {"answer", "dna_sequences"}


def test_plain_text_prompts_are_not_wrapped():
    seen = {}

    def f(prompts, completions, **kw):
        seen["c"], seen["p"] = completions, prompts
        return [0.0] * len(completions)
    ids = torch.tensor([[1, 2], [3, 9]])
    rw.score([f], examples=[dict(prompt="a"), dict(prompt="b")], prompts=None, completion_ids=ids, completion_mask=torch.ones(2, 2),
             prompt_ids=ids, processing_class=FakeTok())
    assert seen["c"] == ["w1 w2", "w3"] and seen["p"] == ["a", "b"]          # skip_special_tokens dropped the eos


def test_text_rewards_without_tokenizer_fail_loudly():
    ids = torch.zeros(2, 3, dtype=torch.long)
    with pytest.raises(ValueError, match="batch_decode"):
        rw.score([ref_style_format_reward], examples=None, prompts=None, completion_ids=ids, completion_mask=ids, prompt_ids=ids, processing_class=None)
    # token-level functions never need one
    out = rw.score([token_len_reward], examples=None, prompts=None, completion_ids=ids, completion_mask=torch.ones(2, 3), prompt_ids=ids, processing_class=None)
    assert out[:, 0].tolist() == [3.0, 3.0]


def test_callbacks_fire_like_hf_trainer():
    from bioreason_b200.trainer.grpo_trainer import CallbackHandler, TrainerControl, TrainerState
    import types
    log = []

    class Save:                                                            # shape of reason.py:46-81
        def on_save(self, args, state, control, **kwargs):
            log.append(("save", state.global_step, kwargs.get("model")))
            control.should_save = False
            return control

        def on_step_end(self, args, state, control, **kw):
            log.append(("step", state.global_step))
    tr = types.SimpleNamespace(args=types.SimpleNamespace(output_dir="x"), state=TrainerState(), control=TrainerControl(), model="M",
                               processing_class=None, optimizer=None)
    h = CallbackHandler([Save()], tr)
    tr.state.global_step = 3
    h.fire("on_step_end"); tr.control.should_save = True; h.fire("on_save"); h.fire("on_epoch_end")
    assert log == [("step", 3), ("save", 3, "M")] and tr.control.should_save is False
