"""Model-shape registry for the hot path (public configs restated).

No weights or tokenizer files exist offline, so every model on this path is built from a
config object + seeded random init.  Shapes follow the public HF configs of
Qwen/Qwen3-4B, Qwen/Qwen3-1.7B ("Qwen3-1B" in the reference README,
reference `train_dna_qwen.py:1016`) and
InstaDeepAI/nucleotide-transformer-v2-500m-multi-species (`train_dna_qwen.py:1017`).
"""
from __future__ import annotations

from transformers import EsmConfig, Qwen3Config

# ids the reference obtains from `add_special_tokens` (dna_llm.py:72-74) on the Qwen3
# tokenizer (151 669 base ids): <|dna_start|>, <|dna_pad|>, <|dna_end|>
DNA_START_ID, DNA_PAD_ID, DNA_END_ID = 151669, 151670, 151671
QWEN_EOS_ID = 151645  # <|im_end|>; pad_token = eos_token (dna_llm.py:70)

_TEXT = {
    "qwen3-4b": dict(hidden_size=2560, num_hidden_layers=36, num_attention_heads=32,
                     num_key_value_heads=8, head_dim=128, intermediate_size=9728,
                     vocab_size=151936, tie_word_embeddings=True),
    "qwen3-1.7b": dict(hidden_size=2048, num_hidden_layers=28, num_attention_heads=16,
                       num_key_value_heads=8, head_dim=128, intermediate_size=6144,
                       vocab_size=151936, tie_word_embeddings=True),
    # shrunken shapes for fast parity tests; head_dim stays 128 (the kernels' tile shape)
    "tiny": dict(hidden_size=256, num_hidden_layers=2, num_attention_heads=4,
                 num_key_value_heads=2, head_dim=128, intermediate_size=512,
                 vocab_size=1024, tie_word_embeddings=True),
    "small": dict(hidden_size=512, num_hidden_layers=4, num_attention_heads=8,
                  num_key_value_heads=2, head_dim=128, intermediate_size=1536,
                  vocab_size=4096, tie_word_embeddings=True),
}

_DNA = {
    "nt-v2-500m": dict(hidden_size=1024, num_hidden_layers=29, num_attention_heads=16,
                       intermediate_size=4096, vocab_size=4107),
    "tiny": dict(hidden_size=128, num_hidden_layers=2, num_attention_heads=2,
                 intermediate_size=256, vocab_size=64),
    "small": dict(hidden_size=256, num_hidden_layers=3, num_attention_heads=4,
                  intermediate_size=512, vocab_size=256),
}


def text_config(name: str) -> Qwen3Config:
    kw = dict(_TEXT[name])
    cfg = Qwen3Config(rope_theta=1e6, rms_norm_eps=1e-6, attention_bias=False,
                      max_position_embeddings=40960, attention_dropout=0.0,
                      use_sliding_window=False, **kw)
    v = cfg.vocab_size
    if name in ("tiny", "small"):
        cfg.dna_token_ids = (v - 3, v - 2, v - 1)
        cfg.eos_token_id = v - 4
    else:
        cfg.dna_token_ids = (DNA_START_ID, DNA_PAD_ID, DNA_END_ID)
        cfg.eos_token_id = QWEN_EOS_ID
    cfg.pad_token_id = cfg.eos_token_id  # dna_llm.py:70
    cfg.bos_token_id = None
    cfg._attn_implementation = "sdpa"
    return cfg


def dna_config(name: str) -> EsmConfig:
    """NT-v2 = ESM skeleton + rotary + gated-SiLU FFN without FFN biases."""
    kw = dict(_DNA[name])
    cfg = EsmConfig(position_embedding_type="rotary", layer_norm_eps=1e-12,
                    pad_token_id=1, mask_token_id=2, token_dropout=False,
                    emb_layer_norm_before=False, hidden_dropout_prob=0.0,
                    attention_probs_dropout_prob=0.0, max_position_embeddings=2050,
                    is_decoder=False, add_cross_attention=False, **kw)
    cfg.cls_token_id = 3
    cfg.add_bias_fc = False      # NT-v2: no bias on the two FFN linears
    cfg.gated_mlp = True         # NT-v2: dense 1024 -> 2*4096, SiLU(x1)*x2
    cfg._attn_implementation = "sdpa"
    return cfg
