"""Host logic of checkpoint interop (bioreason_b200/checkpoint.py): the key layouts reason.py:422-537 and the reference's save
callback (reason.py:46-81) produce are mapped onto this model's peft-shaped state_dict keys.  No GPU: tensors are tiny CPU tensors."""
import json
import os

import pytest
import torch

from bioreason_b200 import checkpoint as ck

MODEL_KEYS = [
    "text_model.model.embed_tokens.weight", "text_model.lm_head.weight",
    "text_model.model.layers.0.self_attn.q_proj.base_layer.weight",
    "text_model.model.layers.0.self_attn.q_proj.lora_A.default.weight",
    "text_model.model.layers.0.self_attn.q_proj.lora_B.default.weight",
    "text_model.model.layers.0.mlp.down_proj.base_layer.weight",
    "text_model.model.norm.weight",
    "dna_model.esm.encoder.layer.0.attention.self.query.weight",
    "dna_projection.weight", "dna_projection.bias",
]
T = lambda: torch.zeros(2, 2)


def test_unwrap_layouts_and_wrapper_prefixes():
    raw = {"text_model.model.norm.weight": T(), "dna_projection.bias": T()}
    assert set(ck.unwrap(raw)) == set(raw)
    lightning = {"state_dict": {"model.text_model.model.norm.weight": T(), "model.dna_projection.bias": T()}, "epoch": 3}
    assert set(ck.unwrap(lightning)) == set(raw)
    deepspeed = {"module": {"_forward_module.model.text_model.model.norm.weight": T(), "_forward_module.model.dna_projection.bias": T()}}
    assert set(ck.unwrap(deepspeed)) == set(raw)
    # an HF text checkpoint keeps its own `model.` root (it is not the Lightning wrapper prefix)
    assert set(ck.unwrap({"model.layers.0.mlp.down_proj.weight": T()})) == {"model.layers.0.mlp.down_proj.weight"}
    with pytest.raises(ValueError, match="Unsupported checkpoint format"):
        ck.unwrap([1, 2, 3])


def test_normalize_keys_peft_nesting_and_adapter_names():
    sd = {
        "text_model.base_model.model.model.layers.0.self_attn.q_proj.lora_A.default.weight": T(),    # saved from a PeftModel (reason.py:494-498)
        "text_model.model.layers.0.self_attn.q_proj.lora_B.weight": T(),                             # adapter file: no adapter name
        "text_model.model.layers.0.self_attn.q_proj.weight": T(),                                    # plain HF base weight -> base_layer
        "text_model.base_model.model.model.layers.0.mlp.down_proj.base_layer.weight": T(),
        "dna_model.esm.encoder.layer.0.attention.self.query.weight": T(),
        "dna_projection.weight": T(),
        "something.else": T(),
    }
    mapped, unexpected = ck.normalize_keys(sd, MODEL_KEYS)
    assert unexpected == ["something.else"]
    assert set(mapped) == {
        "text_model.model.layers.0.self_attn.q_proj.lora_A.default.weight", "text_model.model.layers.0.self_attn.q_proj.lora_B.default.weight",
        "text_model.model.layers.0.self_attn.q_proj.base_layer.weight", "text_model.model.layers.0.mlp.down_proj.base_layer.weight",
        "dna_model.esm.encoder.layer.0.attention.self.query.weight", "dna_projection.weight"}
    # a model WITHOUT adapters loading a checkpoint saved WITH them: base_layer.weight -> weight
    plain, _ = ck.normalize_keys({"text_model.model.layers.0.mlp.down_proj.base_layer.weight": T()}, ["text_model.model.layers.0.mlp.down_proj.weight"])
    assert set(plain) == {"text_model.model.layers.0.mlp.down_proj.weight"}


def test_peft_adapter_dir_and_hf_dir_readers(tmp_path):
    from safetensors.torch import save_file
    ad = tmp_path / "adapter"; ad.mkdir()
    save_file({"base_model.model.model.layers.0.self_attn.q_proj.lora_A.weight": torch.ones(4, 8),
               "base_model.model.model.layers.0.self_attn.q_proj.lora_B.weight": torch.zeros(8, 4)}, str(ad / "adapter_model.safetensors"))
    json.dump({"r": 4, "lora_alpha": 8, "target_modules": ["q_proj"]}, open(ad / "adapter_config.json", "w"))
    sd, cfg = ck.read_peft_adapter_dir(str(ad))
    assert cfg["r"] == 4 and set(sd) == {"text_model.model.layers.0.self_attn.q_proj.lora_A.weight", "text_model.model.layers.0.self_attn.q_proj.lora_B.weight"}
    mapped, unexpected = ck.normalize_keys(sd, MODEL_KEYS)
    assert not unexpected and "text_model.model.layers.0.self_attn.q_proj.lora_A.default.weight" in mapped
    # sharded HF directory
    hf = tmp_path / "hf"; hf.mkdir()
    save_file({"model.norm.weight": torch.ones(3)}, str(hf / "model-00001-of-00002.safetensors"))
    save_file({"model.embed_tokens.weight": torch.ones(5, 3)}, str(hf / "model-00002-of-00002.safetensors"))
    json.dump({"weight_map": {"model.norm.weight": "model-00001-of-00002.safetensors", "model.embed_tokens.weight": "model-00002-of-00002.safetensors"}},
              open(hf / "model.safetensors.index.json", "w"))
    assert set(ck.read_hf_dir(str(hf))) == {"model.norm.weight", "model.embed_tokens.weight"}
    torch.save({"state_dict": {"model.dna_projection.bias": torch.ones(2)}}, str(tmp_path / "pytorch_model.bin"))
    assert set(ck._read_file(str(tmp_path / "pytorch_model.bin"))) == {"dna_projection.bias"}
    with pytest.raises(FileNotFoundError):
        ck.read_hf_dir(str(tmp_path / "adapter"))
