"""Import-compatibility stand-in for the `trl` names the reference's scripts import next to `bioreason.trainer`
(reason.py:32: `from trl import GRPOConfig, GRPOTrainer, ModelConfig, ScriptArguments, TrlParser, get_peft_config`).
trl / accelerate are not installed in this image; the trainer on this path is bioreason_b200.trainer.DNALLMGRPOTrainer."""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass, field
from typing import List, Optional

from bioreason_b200.trainer import DNALLMGRPOConfig as GRPOConfig  # noqa: F401
from bioreason_b200.trainer import DNALLMGRPOTrainer as GRPOTrainer  # noqa: F401


@dataclass
class ModelConfig:
    """The trl.ModelConfig fields reason.py reads (model_name_or_path, LoRA settings, attn_implementation, torch_dtype)."""
    model_name_or_path: Optional[str] = None
    torch_dtype: Optional[str] = "bfloat16"
    attn_implementation: Optional[str] = None
    use_peft: bool = True
    lora_r: int = 32
    lora_alpha: int = 64
    lora_dropout: float = 0.05
    lora_target_modules: Optional[List[str]] = None
    lora_modules_to_save: Optional[List[str]] = None
    lora_task_type: str = "CAUSAL_LM"
    load_in_8bit: bool = False
    load_in_4bit: bool = False


@dataclass
class ScriptArguments:
    dataset_name: Optional[str] = None
    dataset_config: Optional[str] = None
    dataset_train_split: str = "train"
    dataset_test_split: str = "test"


def get_peft_config(model_args):
    """trl.get_peft_config: None unless use_peft; the reference passes the result to the trainer, which ignores it once the model
    already carries adapters (grpo_trainer.py:296-303)."""
    if not getattr(model_args, "use_peft", False):
        return None
    from peft import LoraConfig
    return LoraConfig(r=model_args.lora_r, lora_alpha=model_args.lora_alpha, lora_dropout=model_args.lora_dropout,
                      target_modules=model_args.lora_target_modules, task_type=model_args.lora_task_type, bias="none")


class TrlParser:
    """Keyword-only construction of the dataclass tuple (`parse_args_and_config()` with argparse-style `--name value` pairs)."""

    def __init__(self, dataclass_types):
        self.types = list(dataclass_types) if isinstance(dataclass_types, (list, tuple)) else [dataclass_types]

    def parse_args_and_config(self, args=None):
        import sys
        argv = list(sys.argv[1:] if args is None else args)
        kv, i = {}, 0
        while i < len(argv):
            if argv[i].startswith("--"):
                key = argv[i][2:].replace("-", "_")
                if i + 1 < len(argv) and not argv[i + 1].startswith("--"):
                    kv[key] = argv[i + 1]; i += 2
                else:
                    kv[key] = True; i += 1
            else:
                i += 1
        out = []
        for t in self.types:
            fields = {f.name: f for f in dataclasses.fields(t)}
            kw = {}
            for k, v in kv.items():
                if k in fields:
                    ft = fields[k].type
                    kw[k] = _coerce(v, ft)
            out.append(t(**kw))
        return tuple(out)


def _coerce(v, ft):
    s = str(ft)
    if isinstance(v, bool):
        return v
    if "int" in s and "float" not in s:
        try:
            return int(v)
        except ValueError:
            return v
    if "float" in s:
        try:
            return float(v)
        except ValueError:
            return v
    if "bool" in s:
        return str(v).lower() in ("1", "true", "yes")
    if "List" in s or "list" in s:
        return [x for x in str(v).split(",") if x]
    return v
