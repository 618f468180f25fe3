"""The drop-in surface: our DNALLMModel / DNALLMGRPOTrainer / DNALLMGRPOConfig accept what the reference's callers pass
(signatures recorded from /root/reference by tests/golden/make_api_golden.py)."""
import inspect
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
API = json.load(open(os.path.join(HERE, "golden", "reference_api.json")))


def _params(fn):
    return inspect.signature(fn).parameters


def test_model_signatures_cover_reference():
    from bioreason_b200.models.dna_llm import DNALLMModel
    for meth in ("__init__", "forward", "generate", "process_dna_embeddings"):
        ref = API["DNALLMModel"][meth]
        ours = _params(getattr(DNALLMModel, meth))
        for i, a in enumerate(ref["args"]):
            assert a in ours, f"DNALLMModel.{meth} is missing reference argument {a!r}"
        # same positional order for the arguments callers pass positionally
        assert [p for p in ours][: len(ref["args"])] == ref["args"], (meth, list(ours)[: len(ref["args"])], ref["args"])
        if ref["kwargs"]:
            assert any(p.kind is inspect.Parameter.VAR_KEYWORD for p in ours.values()), f"{meth} must accept **{ref['kwargs']}"
    # reason.py:418 passes debug=False although the reference signature lacks it: ours must absorb it
    assert any(p.kind is inspect.Parameter.VAR_KEYWORD for p in _params(DNALLMModel.__init__).values())
    # reference defaults that callers rely on
    ours = _params(DNALLMModel.__init__)
    assert ours["max_length_dna"].default == 2048 and ours["max_length_text"].default == 512
    assert ours["dna_is_evo2"].default is False and ours["dna_embedding_layer"].default is None


def test_trainer_signatures_cover_reference():
    from bioreason_b200.trainer import DNALLMGRPOTrainer, RepeatRandomSampler
    ref = API["DNALLMGRPOTrainer"]["__init__"]
    ours = _params(DNALLMGRPOTrainer.__init__)
    for a in ref["args"]:
        assert a in ours, f"DNALLMGRPOTrainer.__init__ is missing reference argument {a!r}"
    for meth in ("compute_loss", "_get_per_token_logps", "_generate_and_score_completions", "_get_train_sampler"):
        assert hasattr(DNALLMGRPOTrainer, meth)
        for a in API["DNALLMGRPOTrainer"][meth]["args"]:
            assert a in _params(getattr(DNALLMGRPOTrainer, meth)), (meth, a)
    assert list(_params(RepeatRandomSampler.__init__))[:6] == API["RepeatRandomSampler"]["__init__"]["args"]


def test_config_fields_and_defaults():
    from bioreason_b200.trainer import DNALLMGRPOConfig
    import dataclasses
    ours = {f.name: f for f in dataclasses.fields(DNALLMGRPOConfig)}
    ref = API["DNALLMGRPOConfig"]
    missing = [k for k in ref if k not in ours]
    assert not missing, f"config fields missing: {missing}"
    c = DNALLMGRPOConfig()
    for k, d in ref.items():
        if d is None or k == "report_to":            # report_to defaults to "wandb" in the reference; wandb is optional here
            continue
        assert repr(getattr(c, k)) == d or str(getattr(c, k)) == d.strip("'\""), (k, getattr(c, k), d)


def test_compat_package_reexports():
    import importlib, sys
    root = os.path.dirname(HERE)
    sys.path.insert(0, os.path.join(root, "compat"))
    try:
        for name in [m for m in list(sys.modules) if m == "bioreason" or m.startswith("bioreason.")]:
            del sys.modules[name]
        m = importlib.import_module("bioreason.models.dna_llm")
        t = importlib.import_module("bioreason.trainer")
        from bioreason_b200.models.dna_llm import DNALLMModel
        from bioreason_b200.trainer import DNALLMGRPOTrainer
        assert m.DNALLMModel is DNALLMModel and t.DNALLMGRPOTrainer is DNALLMGRPOTrainer
    finally:
        sys.path.remove(os.path.join(root, "compat"))
        for name in [m for m in list(sys.modules) if m == "bioreason" or m.startswith("bioreason.")]:
            del sys.modules[name]
