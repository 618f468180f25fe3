"""The `peft` / `trl` stand-ins under compat/ against the reference's OWN source: `_get_target_modules` and `_prep_for_training` are
executed from /root/reference/reason.py (extracted with ast, not copied) on a CPU stand-in that records what the kernels' LoRA
entry point would be asked to do.  Skipped where the reference tree is absent (the GPU box)."""
import ast
import os
import sys
import types
import weakref

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REASON = "/root/reference/reason.py"


@pytest.fixture()
def compat_path():
    p = os.path.join(ROOT, "compat")
    sys.path.insert(0, p)
    for m in [k for k in sys.modules if k.split(".")[0] in ("peft", "trl")]:
        del sys.modules[m]
    yield p
    sys.path.remove(p)
    for m in [k for k in sys.modules if k.split(".")[0] in ("peft", "trl")]:
        del sys.modules[m]


class _StandIn(torch.nn.Module):
    """CPU stand-in with the attributes reason.py touches; records the enable_lora() call the shim makes."""
    def __init__(self):
        super().__init__()
        from transformers import Qwen3ForCausalLM
        from bioreason_b200.configs import text_config
        tc = text_config("tiny"); tc.num_hidden_layers = 1
        self.text_model = Qwen3ForCausalLM(tc)
        self.dna_model = torch.nn.Linear(4, 4)
        self.dna_projection = torch.nn.Linear(4, 4)
        self.calls = []
        object.__setattr__(self.text_model, "_b200_owner", weakref.ref(self))

    @property
    def text(self):
        return self.text_model

    def enable_lora(self, r, alpha, seed=0):
        self.calls.append((r, alpha))


@pytest.mark.skipif(not os.path.exists(REASON), reason="reference tree not present")
def test_reference_prep_for_training_runs_against_the_shims(compat_path):
    import peft
    tree = ast.parse(open(REASON).read())
    wanted = {"_get_target_modules", "_prep_for_training"}
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in wanted]
    assert {f.name for f in fns} == wanted
    ns = {"torch": torch, "LoraConfig": peft.LoraConfig, "get_peft_model": peft.get_peft_model,
          "prepare_model_for_kbit_training": peft.prepare_model_for_kbit_training, "DNALLMModel": _StandIn}
    exec(compile(ast.Module(body=fns, type_ignores=[]), REASON, "exec"), ns)
    m = _StandIn()
    args = types.SimpleNamespace(lora_r=32, lora_alpha=64, lora_dropout=0.05)
    cfg = ns["_prep_for_training"](m, args, dna_model_finetune=False)                             # reason.py:362-394, unmodified
    assert m.calls == [(32, 64.0)] and m.lora_dropout == 0.05
    assert cfg.r == 32 and {"q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj"} <= set(cfg.target_modules)
    assert not any(p.requires_grad for p in m.dna_model.parameters()) and all(p.requires_grad for p in m.dna_projection.parameters())
    assert not any(p.requires_grad for p in m.text_model.parameters())                             # prepare_model_for_kbit_training froze the base


def test_shims_refuse_foreign_models_and_partial_targets(compat_path):
    import peft
    with pytest.raises(TypeError, match="DNALLMModel"):
        peft.get_peft_model(torch.nn.Linear(2, 2), peft.LoraConfig(r=4))
    m = _StandIn()
    with pytest.raises(NotImplementedError, match="subset"):
        peft.get_peft_model(m.text_model, peft.LoraConfig(r=4, target_modules=["q_proj", "v_proj"]))


def test_trl_parser_builds_the_dataclass_tuple(compat_path):
    import trl
    from bioreason_b200.trainer import DNALLMGRPOConfig
    s, t, mo = trl.TrlParser((trl.ScriptArguments, DNALLMGRPOConfig, trl.ModelConfig)).parse_args_and_config(
        ["--num_generations", "4", "--learning_rate", "2e-6", "--use_peft", "true", "--lora_r", "8", "--dataset_name", "kegg"])
    assert t.num_generations == 4 and t.learning_rate == 2e-6 and mo.lora_r == 8 and s.dataset_name == "kegg"
    cfg = trl.get_peft_config(mo)
    assert cfg.r == 8 and cfg.bias == "none"
    assert trl.get_peft_config(trl.ModelConfig(use_peft=False)) is None
