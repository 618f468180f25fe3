from .dna_llm import DNALLMModel  # noqa: F401
