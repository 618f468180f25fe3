"""Record the reference's public signatures on the hot path (run in the build container; output committed)."""
import ast, json, os
REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_api.json")


def sigs(path, cls, methods):
    tree = ast.parse(open(os.path.join(REF, path)).read())
    out = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for f in node.body:
                if isinstance(f, ast.FunctionDef) and f.name in methods:
                    a = f.args
                    names = [x.arg for x in a.args]
                    defaults = [ast.unparse(d) for d in a.defaults]
                    out[f.name] = dict(args=names, defaults=defaults, kwargs=a.kwarg.arg if a.kwarg else None)
    return out


def fields(path, cls):
    tree = ast.parse(open(os.path.join(REF, path)).read())
    out = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for f in node.body:
                if isinstance(f, ast.AnnAssign) and isinstance(f.target, ast.Name):
                    d = None
                    if isinstance(f.value, ast.Call):
                        for kw in f.value.keywords:
                            if kw.arg == "default":
                                d = ast.unparse(kw.value)
                    elif f.value is not None:
                        d = ast.unparse(f.value)
                    out[f.target.id] = d
    return out


api = {
    "DNALLMModel": sigs("bioreason/models/dna_llm.py", "DNALLMModel", {"__init__", "forward", "generate", "process_dna_embeddings"}),
    "DNALLMGRPOTrainer": sigs("bioreason/trainer/grpo_trainer.py", "DNALLMGRPOTrainer",
                              {"__init__", "compute_loss", "_get_per_token_logps", "_generate_and_score_completions", "_get_train_sampler"}),
    "RepeatRandomSampler": sigs("bioreason/trainer/grpo_trainer.py", "RepeatRandomSampler", {"__init__"}),
    "DNALLMGRPOConfig": fields("bioreason/trainer/grpo_config.py", "DNALLMGRPOConfig"),
}
json.dump(api, open(OUT, "w"), indent=1, sort_keys=True)
print("wrote", OUT)
