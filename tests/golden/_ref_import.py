"""Import the UNMODIFIED reference (/root/reference) in this container.

Test infrastructure only (used by make_golden.py to generate fixtures).  The reference
does not import as-is here: `transformers.processing_utils.CommonKwargs`
is gone in transformers 5.x and `trl` / `accelerate` / `peft` are not installed.  We
install inert stand-ins for those *imports* (never for the arithmetic under test) and
then import the reference's own modules, so that the reference's own code for
DNALLMModel.forward/generate, _get_per_token_logps, compute_loss, the advantage block
and RepeatRandomSampler executes unmodified.
"""
import sys, types, importlib

REF = "/root/reference"


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_shims():
    import transformers, transformers.processing_utils as pu
    if not hasattr(pu, "CommonKwargs"):
        from typing import TypedDict
        class CommonKwargs(TypedDict, total=False):
            pass
        pu.CommonKwargs = CommonKwargs
    def _nop(*a, **k):
        raise RuntimeError("shimmed symbol called; not part of the golden path")
    if "trl" not in sys.modules:
        _mod("trl", SFTTrainer=type("SFTTrainer", (), {}))
        _mod("trl.data_utils", apply_chat_template=_nop, is_conversational=lambda x: False,
             maybe_apply_chat_template=_nop)
        _mod("trl.models", create_reference_model=_nop, prepare_deepspeed=_nop,
             unwrap_model_for_generation=_nop)
        _mod("trl.trainer")
        from transformers import TrainingArguments
        _mod("trl.trainer.grpo_config", GRPOConfig=TrainingArguments)
        _mod("trl.trainer.utils", generate_model_card=_nop, get_comet_experiment_url=_nop)
    try:
        import accelerate  # noqa
    except Exception:
        _mod("accelerate")
        _mod("accelerate.utils", is_peft_model=lambda m: False, set_seed=_nop, gather_object=_nop)
    for name in ("AriaForConditionalGeneration", "AriaProcessor", "Qwen2VLForConditionalGeneration",
                 "Qwen2_5_VLForConditionalGeneration"):
        try:
            getattr(transformers, name)
        except Exception:
            setattr(transformers, name, type(name, (), {}))
    if REF not in sys.path:
        sys.path.insert(0, REF)


def load_reference():
    install_shims()
    dna_llm = importlib.import_module("bioreason.models.dna_llm")
    trainer = importlib.import_module("bioreason.trainer.grpo_trainer")
    return dna_llm, trainer
