from bioreason_b200.trainer import DNALLMGRPOConfig, DNALLMGRPOTrainer  # noqa: F401  (bioreason/trainer/__init__.py:1-7)

__all__ = ["DNALLMGRPOConfig", "DNALLMGRPOTrainer"]
