"""The drop-in boundary end to end on the GPU (tiny shapes): the reference's own call sequence -- model build, LoRA through the `peft`
entry points, trainer construction with the reference's kwargs, text reward functions, the save callback of reason.py:46-81 -- against
`compat/` (bioreason / peft / trl stand-ins) + checkpoint round trips through every layout checkpoint.py reads."""
import json
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def compat_path():
    p = os.path.join(ROOT, "compat")
    sys.path.insert(0, p)
    for m in [k for k in sys.modules if k.split(".")[0] in ("peft", "trl", "bioreason")]:
        del sys.modules[m]
    yield p
    sys.path.remove(p)
    for m in [k for k in sys.modules if k.split(".")[0] in ("peft", "trl", "bioreason")]:
        del sys.modules[m]


class FakeTok:
    """batch_decode / eos / pad: all the trainer needs from `processing_class` once the batch is tokenised."""
    def __init__(self, eos):
        self.eos_token_id = self.pad_token_id = eos

    def batch_decode(self, ids, skip_special_tokens=False):
        return [" ".join(str(t) for t in row if not (skip_special_tokens and t == self.eos_token_id)) for row in ids.tolist()]


def test_reference_call_sequence_through_compat(compat_path, tmp_path):
    from peft import LoraConfig, get_peft_model, prepare_model_for_kbit_training                  # reason.py:24
    from trl import GRPOConfig, ModelConfig, ScriptArguments, TrlParser, get_peft_config          # reason.py:32
    from bioreason.models.dna_llm import DNALLMModel                                              # reason.py:35
    from bioreason.trainer import DNALLMGRPOConfig, DNALLMGRPOTrainer                             # reason.py:38
    from bioreason_b200.configs import text_config, dna_config
    from oracle.models import synth_batch
    tc, dc = text_config("tiny"), dna_config("tiny")
    # ---- reason.py:407-419
    model = DNALLMModel(text_model_name=tc, dna_model_name=dc, cache_dir=None, max_length_text=512, max_length_dna=2048,
                        text_model_finetune=True, dna_model_finetune=False, debug=False)
    # ---- reason.py:83-113 (_get_target_modules) + :362-394 (_prep_for_training), restated call for call
    for param in model.dna_model.parameters():
        param.requires_grad = False
    target_modules, seen = [], set()
    for name, module in model.text.named_modules():                                               # `.text`: reason.py:89
        if isinstance(module, torch.nn.Linear):
            t = name.split(".")[-1]
            if t != "lm_head" and t not in seen:
                target_modules.append(t); seen.add(t)
    for pattern in ("q_proj", "k_proj", "v_proj", "out_proj", "query", "key", "value"):
        if pattern not in seen:
            target_modules.append(pattern)
    lora_config = LoraConfig(r=16, lora_alpha=32, lora_dropout=0.05, target_modules=target_modules, init_lora_weights="gaussian",
                             bias="none", task_type="CAUSAL_LM")
    model.text_model = prepare_model_for_kbit_training(model.text_model)
    model.text_model = get_peft_model(model.text_model, lora_config)
    for param in model.dna_projection.parameters():
        param.requires_grad = True
    lora_keys = [k for k in model.state_dict() if "lora_" in k]
    assert len(lora_keys) == 2 * 7 * tc.num_hidden_layers and all(".default.weight" in k for k in lora_keys)
    assert not any(p.requires_grad for p in model.dna_model.parameters())

    # ---- text reward functions in the reference protocol (reason.py reward registry shape) + the save callback of reason.py:46-81
    seen_kwargs = {}

    def format_reward(completions, **kwargs):
        seen_kwargs.update(kwargs)
        return [0.5 if len(c[0]["content"]) % 2 == 0 else 0.0 for c in completions]

    def correctness_reward(prompts, completions, answer, **kwargs):
        assert len(prompts) == len(completions) == len(answer)
        return [1.0 if a in c[0]["content"] else -0.25 * (i % 3) for i, (c, a) in enumerate(zip(completions, answer))]

    class SaveWithPyTorchCallback:                                                               # body of reason.py:48-81, same calls
        def on_save(self, args, state, control, **kwargs):
            folder = os.path.join(args.output_dir, f"checkpoint-{state.global_step}")
            os.makedirs(folder, exist_ok=True)
            m = kwargs.get("model")
            m = m.module if hasattr(m, "module") else m
            torch.save(m.state_dict(), os.path.join(folder, "pytorch_model.bin"))
            if hasattr(m, "text_model") and hasattr(m.text_model, "config"):
                m.text_model.config.save_pretrained(folder)
            control.should_save = False
            return control

    (script_args, training_args, model_args) = TrlParser((ScriptArguments, DNALLMGRPOConfig, ModelConfig)).parse_args_and_config(
        ["--output_dir", str(tmp_path), "--num_generations", "4", "--max_completion_length", "6", "--per_device_train_batch_size", "4",
         "--learning_rate", "1e-3", "--save_steps", "1", "--lora_r", "16", "--lora_alpha", "32", "--use_peft", "true"])
    training_args.save_safetensors = False                                                       # reason.py:597
    batch = synth_batch(tc, dc, batch=4, n_seq=2, dna_len=10, text_len=18, seed=14, same_prompt=True)
    examples = [dict(prompt=[{"role": "user", "content": "q"}], answer=str(7 + i), dna_sequences=["ACGT", "GG"]) for i in range(4)]
    trainer = DNALLMGRPOTrainer(model=model, reward_funcs=[format_reward, correctness_reward], args=training_args, dna_module=None,
                                train_dataset=None, eval_dataset=None, peft_config=get_peft_config(model_args),
                                attn_implementation="flash_attention_2", torch_dtype="bfloat16", callbacks=[SaveWithPyTorchCallback()],
                                processing_class=FakeTok(tc.eos_token_id))
    before = {k: v.detach().clone() for k, v in model.state_dict().items() if "lora_B" in k}
    loss = trainer.training_step(dict(batch, examples=examples))
    assert torch.isfinite(loss) and set(seen_kwargs) == {"prompts", "answer", "dna_sequences"}      # prompts= plus every other dataset column
    assert trainer.reward_d2h_bytes == 4 * 6 * 8                                                 # completion ids went to the host once
    ck = os.path.join(str(tmp_path), "checkpoint-1", "pytorch_model.bin")
    assert os.path.exists(ck) and os.path.exists(os.path.join(str(tmp_path), "checkpoint-1", "config.json"))
    moved = sum(int(not torch.equal(before[k], v)) for k, v in model.state_dict().items() if k in before)
    assert moved > 0

    # ---- reload what the callback wrote (reason.py:448-480: raw state dict with LoRA keys -> adapters enabled first)
    fresh = DNALLMModel(text_model_name=tc, dna_model_name=dc, seed=999)                          # different init on purpose
    res = fresh.load_state_dict(torch.load(ck, map_location="cpu"))
    assert not res.unexpected_keys and not res.missing_keys, (res.missing_keys[:3], res.unexpected_keys[:3])
    with torch.no_grad():
        a = model(**batch).logits.float(); b = fresh(**batch).logits.float()
    assert torch.equal(a, b), "a reloaded checkpoint must reproduce the logits bit for bit"
    # Lightning-style nesting of the same tensors
    wrapped = {"state_dict": {"model." + k: v for k, v in torch.load(ck, map_location="cpu").items()}}
    fresh2 = DNALLMModel(text_model_name=tc, dna_model_name=dc, seed=5)
    fresh2.load_state_dict(wrapped)
    assert torch.equal(fresh2(**batch).logits.float(), a)


def test_peft_adapter_dir_and_hf_dir_round_trip(compat_path, tmp_path):
    from peft import PeftModel
    from safetensors.torch import save_file
    from bioreason.models.dna_llm import DNALLMModel
    from bioreason_b200.configs import text_config, dna_config
    from oracle.models import synth_batch
    tc, dc = text_config("tiny"), dna_config("tiny")
    batch = synth_batch(tc, dc, batch=2, n_seq=1, dna_len=9, text_len=20, seed=3)
    src = DNALLMModel(tc, dc, seed=21)
    lora = src.enable_lora(r=16, alpha=32.0, seed=4)
    with torch.no_grad():
        g = torch.Generator().manual_seed(1)
        for p in lora.params[1::2]:
            p.copy_((torch.randn(p.shape, generator=g) * 0.05).to(p.device))
    src.sync_adapters(rollout=False)
    want = src(**batch).logits.float()
    # ---- HF directories for the two base models (dna_llm.py:57-68): config.json + model.safetensors
    tdir, ddir = tmp_path / "text", tmp_path / "dna"; tdir.mkdir(); ddir.mkdir()
    tsd = {k[len("text_model."):].replace(".base_layer.", "."): v.detach().cpu().contiguous() for k, v in src.state_dict().items()
           if k.startswith("text_model.") and "lora_" not in k and k != "text_model.lm_head.weight"}
    save_file(tsd, str(tdir / "model.safetensors"))
    tc.save_pretrained(str(tdir))
    cfgj = json.load(open(tdir / "config.json")); cfgj["dna_token_ids"] = list(tc.dna_token_ids); json.dump(cfgj, open(tdir / "config.json", "w"))
    dsd = {k[len("dna_model."):]: v.detach().cpu().contiguous() for k, v in src.state_dict().items() if k.startswith("dna_model.")}
    save_file({k: v.clone() for k, v in dsd.items()}, str(ddir / "model.safetensors"))
    dc.save_pretrained(str(ddir))
    # ---- a peft adapter directory (what `PeftModel.save_pretrained` writes)
    adir = tmp_path / "adapter"; adir.mkdir()
    asd = {"base_model.model." + k[len("text_model."):].replace(".default.weight", ".weight"): v.detach().cpu().contiguous()
           for k, v in src.state_dict().items() if "lora_" in k}
    save_file(asd, str(adir / "adapter_model.safetensors"))
    json.dump({"r": 16, "lora_alpha": 32, "peft_type": "LORA"}, open(adir / "adapter_config.json", "w"))
    dst = DNALLMModel(str(tdir), str(ddir), seed=77)
    dst.load_weights({k: v for k, v in src.state_dict().items() if k.startswith("dna_projection.")})
    assert dst.text_config.hidden_size == tc.hidden_size and dst._lora is None
    base_logits = dst(**batch).logits.float()
    dst.text_model = PeftModel.from_pretrained(dst.text_model, str(adir), is_trainable=True)     # reason.py:432-436
    assert dst._lora is not None and dst._lora.r == 16
    got = dst(**batch).logits.float()
    assert torch.equal(got, want), "HF dirs + adapter dir must reproduce the source model bit for bit"
    assert not torch.equal(base_logits, want)
    dst.text_model = dst.text_model.merge_and_unload()                                           # reason.py:443-446
    assert dst._lora is None and not any("lora_" in k for k in dst.state_dict())
    merged = dst(**batch).logits.float()
    assert (merged - want).abs().max().item() < 0.05 * want.abs().max().item() + 0.05         # merged bf16 weights vs two-segment accumulation
