"""`DNALLMGRPOConfig` -- field names and defaults of bioreason/trainer/grpo_config.py:22-364 that the hot path reads.

The reference subclasses HF `TrainingArguments` (which needs `accelerate`, absent here); this is a plain dataclass with
the same names so scripts that build the config by keyword keep working.  vLLM fields (`grpo_config.py:231-281`) are
accepted and ignored exactly as the reference trainer ignores them.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Union


@dataclass
class DNALLMGRPOConfig:
    output_dir: str = "grpo_out"
    # data / generation (grpo_config.py:146-228)
    model_init_kwargs: Optional[dict] = None
    remove_unused_columns: Optional[bool] = False
    max_prompt_length: Optional[int] = 512
    num_generations: Optional[int] = 8
    max_completion_length: Optional[int] = 800
    ds3_gather_for_generation: bool = True
    temperature: float = 0.6      # NOTE: the reference trainer hard-codes T=0.6 / top_p=0.95 / top_k=20 (grpo_trainer.py:384-391)
    top_p: float = 0.95
    top_k: Optional[int] = 20
    min_p: Optional[float] = None
    repetition_penalty: float = 1.0
    cache_implementation: Optional[str] = None
    # vLLM (never read by the reference trainer)
    use_vllm: Optional[bool] = False
    vllm_device: Optional[str] = "auto"
    vllm_gpu_memory_utilization: float = 0.9
    vllm_dtype: Optional[str] = "auto"
    vllm_max_model_len: Optional[int] = None
    vllm_enable_prefix_caching: Optional[bool] = True
    vllm_guided_decoding_regex: Optional[str] = None
    # optimisation (grpo_config.py:284-340)
    learning_rate: float = 1e-6
    beta: float = 0.04
    num_iterations: int = 1
    epsilon: float = 0.2
    epsilon_high: Optional[float] = None
    reward_weights: Optional[list] = None
    sync_ref_model: bool = False
    ref_model_mixup_alpha: float = 0.6
    ref_model_sync_steps: int = 512
    log_completions: bool = True
    report_to: Union[None, str, list] = "none"
    logging_first_step: bool = False
    logging_steps: float = 2
    # the TrainingArguments fields the trainer touches
    per_device_train_batch_size: int = 8
    per_device_eval_batch_size: int = 8
    gradient_accumulation_steps: int = 1
    max_steps: int = -1
    num_train_epochs: float = 1.0
    weight_decay: float = 0.0
    adam_beta1: float = 0.9
    adam_beta2: float = 0.999
    adam_epsilon: float = 1e-8
    max_grad_norm: float = 1.0
    seed: int = 42
    bf16: bool = True
    gradient_checkpointing: bool = False
    eval_strategy: str = "no"
    save_steps: int = 0                   # > 0: fire the callbacks' on_save every save_steps optimizer steps (HF save_strategy="steps")
    save_safetensors: bool = False        # reason.py:597 sets it; saving itself is the callbacks' job (reason.py:46-81)
    lora_dropout: float = 0.05            # accepted; the kernels apply no dropout (out of scope)
    # B200 build additions
    lora_r: int = 32
    lora_alpha: float = 64.0
    micro_rows: Optional[int] = None      # rows per forward/backward chunk (None = all rows at once)
    suppress_eos: bool = False            # fixed-length rollouts (bench config c)
