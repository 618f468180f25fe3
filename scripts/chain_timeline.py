"""Per-phase timeline of one persistent chain launch (globaltimer stamps) at Qwen3-4B decode shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bioreason_b200 import ops
from bioreason_b200._lib import lib, ffi
dev, bf = "cuda", torch.bfloat16
d, F, HqD, NQ, R = 2560, 9728, 4096, 6144, 8
mk = lambda *s: (torch.randn(*s, device=dev) * 0.02).to(bf)
sets = [dict(o=mk(d, HqD), gu=mk(2 * F, d), down=mk(d, F), qkv=mk(NQ, d)) for _ in range(3)]
attn = mk(R, HqD); h = mk(R, d); x2 = torch.empty(R, d, device=dev, dtype=bf); act = torch.empty(R, F, device=dev, dtype=bf)
qkv = torch.empty(R, NQ, device=dev, dtype=bf)
n_part = (d // 128) * 4
ssa = torch.ones(n_part, 32, device=dev); ssb = torch.ones(n_part, 32, device=dev)
scratch = ops.skinny_scratch(2 * F, dev)
nsm = torch.cuda.get_device_properties(0).multi_processor_count
dbg = torch.zeros(nsm, 32, dtype=torch.int64, device=dev)
def run(w):
    ops.skinny_chain([dict(x=attn, w=w["o"], out=x2, mode=1, residual=h, sumsq_out=ssb), dict(x=x2, w=w["gu"], out=act, mode=2, sumsq_in=ssb, sumsq_in_n=n_part),
                      dict(x=act, w=w["down"], out=h, mode=1, residual=x2, sumsq_out=ssa), dict(x=h, w=w["qkv"], out=qkv, mode=0, sumsq_in=ssa, sumsq_in_n=n_part)], R, scratch, eps=1e-6)
for w in sets: run(w)
torch.cuda.synchronize()
lib().br_skinny_chain_debug(ffi.cast("long long*", dbg.data_ptr()))
run(sets[0]); torch.cuda.synchronize()
lib().br_skinny_chain_debug(ffi.NULL)
t = dbg.double().cpu(); t0 = t[:, 0].min(); rel = (t - t0) / 1e3
print("start %.1f/%.1f  depwait %.1f/%.1f" % (rel[:, 0].mean(), rel[:, 0].max(), rel[:, 1].mean(), rel[:, 1].max()))
for pi, name in enumerate(("o", "gate_up", "down", "qkv")):
    b = 2 + pi * 6
    cols = [rel[:, b + k] for k in range(5)]
    live = t[:, b + 1] > 0
    f = lambda c, m=live: (c[m].mean().item(), c[m].max().item()) if m.any() else (float("nan"), float("nan"))
    print(f"{name:8s} begin %.1f/%.1f  first-acc %.1f/%.1f  tiles-done %.1f/%.1f  bar-arrive %.1f/%.1f  bar-pass %.1f/%.1f" %
          (*f(cols[0], t[:, b] > 0), *f(cols[1]), *f(cols[2], t[:, b + 2] > 0), *f(cols[3], t[:, b + 3] > 0), *f(cols[4], t[:, b + 4] > 0)))
