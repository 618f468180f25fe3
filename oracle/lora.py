"""Torch restatement of peft's LoRA Linear (TEST INFRASTRUCTURE): y = base(x) + (alpha/r) * B(A(x)), dropout 0.
Used to differentiate the oracle model exactly the way the reference trains it (reason.py:362-394: r=32, alpha=64,
all nn.Linear of the text model except lm_head; base weights frozen)."""
import torch
import torch.nn as nn

TARGETS = ("q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj")


class OracleLoraLinear(nn.Module):
    def __init__(self, base: nn.Linear, r: int, alpha: float):
        super().__init__()
        self.base_layer = base
        base.weight.requires_grad_(False)
        self.lora_A = nn.ModuleDict({"default": nn.Linear(base.in_features, r, bias=False)})
        self.lora_B = nn.ModuleDict({"default": nn.Linear(r, base.out_features, bias=False)})
        self.scaling = alpha / r

    def forward(self, x):
        return self.base_layer(x) + self.lora_B["default"](self.lora_A["default"](x)) * self.scaling


def inject(text_model, r=32, alpha=64.0):
    for p in text_model.parameters():
        p.requires_grad_(False)
    for layer in text_model.model.layers:
        for parent, names in ((layer.self_attn, TARGETS[:4]), (layer.mlp, TARGETS[4:])):
            for n in names:
                setattr(parent, n, OracleLoraLinear(getattr(parent, n), r, alpha))
    return text_model
