"""Checkpoint interop for the hot path: the formats the reference's scripts read and write.

  * `reason.py:448-537`   -- `torch.load` of a PyTorch file that is a raw state dict, or {"state_dict": ...} (Lightning, keys prefixed
                             `model.`), or {"module": ...} (DeepSpeed, keys prefixed `_forward_module.`), with or without LoRA keys, with
                             peft's `text_model.base_model.model.` nesting or without it;
  * `reason.py:422-446`   -- a peft adapter DIRECTORY (`adapter_model.safetensors` / `.bin`, keys `base_model.model.<path>.lora_A.weight`);
  * `dna_llm.py:57-68`    -- HF model DIRECTORIES for the text and DNA models (`config.json` + `model.safetensors` / sharded index /
                             `pytorch_model.bin`);
  * `reason.py:46-81`     -- what `SaveWithPyTorchCallback` writes: `torch.save(model.state_dict(), "pytorch_model.bin")`.

Everything here is host-side key plumbing; tensors are copied into the kernel-layout buffers by `DNALLMModel.load_weights`.
"""
from __future__ import annotations

import json
import os
import re
from typing import Dict, Iterable, Tuple

import torch

def _strip(k: str) -> str:
    """Wrapper prefixes: DeepSpeed engine / DDP / the Lightning module attribute (`model.`; reason.py:454 spells it `=model.`)."""
    changed = True
    while changed:
        changed = False
        for pre in ("_forward_module.", "module."):
            if k.startswith(pre):
                k, changed = k[len(pre):], True
        for pre in ("=model.", "model."):
            if k.startswith(pre) and k[len(pre):].split(".")[0] in ("text_model", "dna_model", "dna_projection"):
                k, changed = k[len(pre):], True
    return k


def unwrap(checkpoint) -> Dict[str, torch.Tensor]:
    """reason.py:458-468: {"state_dict": ...} | {"module": ...} | the state dict itself."""
    if isinstance(checkpoint, dict) and "state_dict" in checkpoint and isinstance(checkpoint["state_dict"], dict):
        checkpoint = checkpoint["state_dict"]
    elif isinstance(checkpoint, dict) and "module" in checkpoint and isinstance(checkpoint["module"], dict):
        checkpoint = checkpoint["module"]
    if not (isinstance(checkpoint, dict) and all(isinstance(k, str) for k in checkpoint.keys())):
        raise ValueError("Unsupported checkpoint format")                      # reason.py:469-470
    return {_strip(k): v for k, v in checkpoint.items() if isinstance(v, torch.Tensor)}


def normalize_keys(sd: Dict[str, torch.Tensor], model_keys: Iterable[str]) -> Tuple[Dict[str, torch.Tensor], list]:
    """Map checkpoint keys onto this model's state_dict keys (peft nesting on either side, adapter name on either side).
    Returns (mapped, unexpected_keys)."""
    model_keys = set(model_keys)
    out, unexpected = {}, []
    for k, v in sd.items():
        cands = [k]
        # peft's PeftModel nesting inside text_model (reason.py:494-505) -- present or absent on either side
        cands.append(k.replace("text_model.base_model.model.", "text_model."))
        if k.startswith("text_model."):
            cands.append("text_model.base_model.model." + k[len("text_model."):])
        # a frozen base weight under a LoRA wrapper is `<lin>.base_layer.weight` here, `<lin>.weight` in a plain HF checkpoint
        more = []
        for c in cands:
            if c.endswith(".weight") and ".lora_" not in c and ".base_layer." not in c:
                more.append(c[:-len(".weight")] + ".base_layer.weight")
            if ".base_layer." in c:
                more.append(c.replace(".base_layer.", "."))
            # adapter files drop the adapter name: lora_A.weight <-> lora_A.default.weight
            m = re.search(r"\.lora_([AB])\.weight$", c)
            if m:
                more.append(c[:m.start()] + f".lora_{m.group(1)}.default.weight")
        cands += more
        # tied embeddings: lm_head.weight may be absent on one side
        hit = next((c for c in cands if c in model_keys), None)
        if hit is None:
            unexpected.append(k)
        else:
            out[hit] = v
    return out, unexpected


def _read_file(path: str) -> Dict[str, torch.Tensor]:
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(path)
    return unwrap(torch.load(path, map_location="cpu", weights_only=True))


def read_hf_dir(path: str) -> Dict[str, torch.Tensor]:
    """All tensors of a HF model directory: model.safetensors, a sharded *.index.json, or pytorch_model.bin."""
    for index in ("model.safetensors.index.json", "pytorch_model.bin.index.json"):
        ip = os.path.join(path, index)
        if os.path.exists(ip):
            shards = sorted(set(json.load(open(ip))["weight_map"].values()))
            sd = {}
            for s in shards:
                sd.update(_read_file(os.path.join(path, s)))
            return sd
    for name in ("model.safetensors", "pytorch_model.bin"):
        fp = os.path.join(path, name)
        if os.path.exists(fp):
            return _read_file(fp)
    raise FileNotFoundError(f"no model.safetensors / pytorch_model.bin (or sharded index) under {path}")


def read_peft_adapter_dir(path: str) -> Tuple[Dict[str, torch.Tensor], dict]:
    """A peft adapter directory (reason.py:428-446): (tensors keyed like this model's text_model.* LoRA parameters, adapter_config)."""
    cfg = {}
    cp = os.path.join(path, "adapter_config.json")
    if os.path.exists(cp):
        cfg = json.load(open(cp))
    for name in ("adapter_model.safetensors", "adapter_model.bin"):
        fp = os.path.join(path, name)
        if os.path.exists(fp):
            raw = _read_file(fp)
            # peft saves `base_model.model.<module path>.lora_A.weight` relative to the wrapped text model
            return {"text_model." + (k[len("base_model.model."):] if k.startswith("base_model.model.") else k): v for k, v in raw.items()}, cfg
    raise FileNotFoundError(f"no adapter_model.safetensors / adapter_model.bin under {path}")


def load_into(model, source, *, strict: bool = False):
    """Load a checkpoint (path to a file / HF dir / peft adapter dir, or an in-memory dict in any of the layouts above) into a
    DNALLMModel.  LoRA keys enable the adapters first (rank and alpha from adapter_config.json or the tensor shapes), as
    reason.py:478-480 does.  Returns torch's (missing_keys, unexpected_keys) pair."""
    adapter_cfg = {}
    if isinstance(source, (str, os.PathLike)):
        source = os.fspath(source)
        if os.path.isdir(source):
            if os.path.exists(os.path.join(source, "adapter_config.json")) or os.path.exists(os.path.join(source, "adapter_model.safetensors")):
                sd, adapter_cfg = read_peft_adapter_dir(source)
            else:
                sd = {"text_model." + k: v for k, v in read_hf_dir(source).items()}
        else:
            sd = _read_file(source)
    else:
        sd = unwrap(source)
    lora_keys = [k for k in sd if ".lora_A" in k]
    if lora_keys and model._lora is None:
        r = int(adapter_cfg.get("r", sd[lora_keys[0]].shape[0]))
        model.enable_lora(r=r, alpha=float(adapter_cfg.get("lora_alpha", 2 * r)))
    mapped, unexpected = normalize_keys(sd, model.state_dict().keys())
    missing = model.load_weights(mapped)
    if strict and (missing or unexpected):
        raise RuntimeError(f"load_into(strict=True): missing {missing[:5]}, unexpected {unexpected[:5]}")
    return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)
