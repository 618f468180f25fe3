from bioreason_b200.models.dna_llm import DNALLMModel  # noqa: F401  (bioreason/models/dna_llm.py:18)
