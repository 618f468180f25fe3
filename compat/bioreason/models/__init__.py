from bioreason_b200.models.dna_llm import DNALLMModel  # noqa: F401

__all__ = ["DNALLMModel"]
