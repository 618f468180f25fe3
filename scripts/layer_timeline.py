"""%globaltimer timeline of one decode layer in the production 5-launch layout (qkv GEMM -> fused attention -> o_proj -> gate/up -> down),
all kernels PDL-chained, at Qwen3-4B shapes with R = 8 rows.  Prints, per kernel, microseconds since the layer's first CTA start."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bioreason_b200 import ops
from bioreason_b200._lib import lib, ffi
dev, bf = "cuda", torch.bfloat16
d, F, Hq, Hkv, D, R, G, P, gen = 2560, 9728, 32, 8, 128, 8, 8, 1848, 256
NL = 4
mk = lambda *s: (torch.randn(*s, device=dev) * 0.02).to(bf)
ws = [dict(qkv=mk((Hq + 2 * Hkv) * D, d), o=mk(d, Hq * D), gu=mk(2 * F, d), down=mk(d, F)) for _ in range(NL)]
scratch = ops.skinny_scratch(2 * F, dev)
PAGE = 64; T = P + gen; n_shared = P // PAGE; priv = math.ceil((T + 1 - n_shared * PAGE) / PAGE); max_pages = n_shared + priv
table = torch.zeros(R, max_pages, dtype=torch.int32); nxt = n_shared
for r in range(R):
    table[r, :n_shared] = torch.arange(n_shared, dtype=torch.int32); table[r, n_shared:] = torch.arange(nxt, nxt + priv, dtype=torch.int32); nxt += priv
table = table.to(dev)
kc = torch.randn(nxt, Hkv, PAGE, D, device=dev).to(bf); vc = torch.randn_like(kc)
qn = torch.ones(D, device=dev).to(bf); kn = torch.ones(D, device=dev).to(bf)
cur = torch.full((R,), T, dtype=torch.int32, device=dev)
rope = ops.rope_table(T + 8, D, 1e6, dev)
n_part = (d // 128) * 4
ssa = torch.ones(n_part, 32, device=dev); ssb = torch.ones(n_part, 32, device=dev)
SS, SP = int(os.environ.get("BR_SS", 8)), int(os.environ.get("BR_SP", 2))
wsf = ops.decode_fused_workspace(R, Hq, Hkv, D, SS + SP, dev)
attn_out = torch.empty(R, Hq * D, device=dev, dtype=bf)
x0 = mk(R, d)
def chain():
    x = x0
    for w in ws:
        q = ops.skinny_gemm(x, w["qkv"], scratch, sumsq_in=ssa, sumsq_in_n=n_part, eps=1e-6)
        ops.decode_attn_fused(q, qn, kn, kc, vc, table, cur, G, Hq, Hkv, D, n_shared, SS, SP, 1e6, 1e-6, wsf, attn_out, rope=rope)
        x2 = ops.skinny_gemm(attn_out, w["o"], scratch, mode=1, residual=x, sumsq_out=ssb)
        a = ops.skinny_gemm(x2, w["gu"], scratch, mode=2, sumsq_in=ssb, sumsq_in_n=n_part, eps=1e-6)
        x = ops.skinny_gemm(a, w["down"], scratch, mode=1, residual=x2, sumsq_out=ssa)
for _ in range(3):
    chain()
torch.cuda.synchronize()
sk = torch.zeros(NL * 4, 160, 8, dtype=torch.int64, device=dev)
items = (R // G) * Hkv * SS + R * Hkv * SP
at = torch.zeros(items, 16, dtype=torch.int64, device=dev)
lib().br_skinny_debug(ffi.cast("long long*", sk.data_ptr()))
lib().br_decode_attn_fused_debug(ffi.cast("long long*", at.data_ptr()))
chain(); torch.cuda.synchronize()
lib().br_skinny_debug(ffi.NULL); lib().br_decode_attn_fused_debug(ffi.NULL)
sk = sk.double().cpu(); at = at.double().cpu()
li = NL - 1
t0 = sk[li * 4][:, 0][sk[li * 4][:, 0] > 0].min()
names = ["entry", "dep-pass", "ready", "first-acc", "published", "red-loads", "red-epi", "done"]
def stat(col):
    v = col[col > 0]
    return "   -   " if v.numel() == 0 else f"{(v.mean() - t0) / 1e3:5.1f}/{(v.max() - t0) / 1e3:5.1f}"
for k, nm in enumerate(("qkv", "o", "gate_up", "down")):
    blk = sk[li * 4 + k]
    n_cta = int((blk[:, 0] > 0).sum())
    print(f"{nm:8s} ctas={n_cta:3d} " + "  ".join(f"{n}={stat(blk[:, i])}" for i, n in enumerate(names)) + f"  first-start={(blk[:, 0][blk[:, 0] > 0].min() - t0) / 1e3:5.1f}")
an = ["start", "dep_wait", "kv_appended", "tile0", "tiles", "partials", "counter", "end", "q_loaded", "q_roped"]
for lab, sl in (("attn-shared", slice(0, (R // G) * Hkv * SS)), ("attn-private", slice((R // G) * Hkv * SS, items))):
    print(f"{lab:12s} " + "  ".join(f"{n}={stat(at[sl, i])}" for i, n in enumerate(an)))
nxt0 = sk[0][:, 0]
print("(mean/max in us since the first qkv CTA of the layer started; the attention stamps belong to the last layer's launch)")
