"""Determinism probe of the config (c) rollout (real widths, depth 2): eager vs graph, repeated, under the A/B environment switches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bioreason_b200.configs import text_config, dna_config
from bioreason_b200.models import DNALLMModel
from oracle.models import build_oracle, synth_batch
tc, dc = text_config("qwen3-4b"), dna_config("nt-v2-500m")
tc.num_hidden_layers = 2; dc.num_hidden_layers = 2
if hasattr(tc, "layer_types"): tc.layer_types = tc.layer_types[:2]
oracle = build_oracle(tc, dc, seed=41)
m = DNALLMModel.from_oracle(oracle)
G, n = 8, 10
batch = synth_batch(tc, dc, batch=G, n_seq=2, dna_len=668, text_len=512, seed=12, same_prompt=True)
outs = []
for rep in range(int(os.environ.get('REPS', 3))):
    for ug in (False, True):
        ids = m.generate(**batch, max_new_tokens=n, do_sample=False, use_graph=ug).cpu()
        outs.append((ug, ids))
        same_rows = all(torch.equal(ids[0], ids[r]) for r in range(G))
        if os.environ.get("QUIET") and same_rows and torch.equal(ids, outs[0][1]): continue
        print(f"rep {rep} graph={ug}: row0 {ids[0].tolist()} rows identical {same_rows}" + ("" if same_rows else f" | differing rows {[r for r in range(G) if not torch.equal(ids[0], ids[r])]}"))
ref = outs[0][1]
print("all runs equal:", all(torch.equal(ref, o) for _, o in outs))
