from .grpo_config import DNALLMGRPOConfig  # noqa: F401
from .grpo_trainer import DNALLMGRPOTrainer, RepeatRandomSampler  # noqa: F401

__all__ = ["DNALLMGRPOConfig", "DNALLMGRPOTrainer", "RepeatRandomSampler"]
