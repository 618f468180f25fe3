"""Per-kernel timing of one decode step's launches at config (c) shapes, each captured in a CUDA graph (no host launch
overhead) and timed with CUDA events.  Prints us/launch and achieved GB/s for the weight-streaming GEMMs."""
import math, sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bioreason_b200 import ops
from bioreason_b200.configs import text_config

torch.manual_seed(0)
tc = text_config(sys.argv[1] if len(sys.argv) > 1 else "qwen3-4b")
d, F, V = tc.hidden_size, tc.intermediate_size, tc.vocab_size
Hq, Hkv, D = tc.num_attention_heads, tc.num_key_value_heads, tc.head_dim
R, G, P, gen = 8, 8, 1848, 256
dev = "cuda"
NL = 6                                   # distinct weight sets so every launch streams from HBM, not L2
bf = torch.bfloat16


def timed_graph(fn, reps=20, inner=1):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(inner):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * inner)


scratch = ops.skinny_scratch(max(V, 2 * F), dev)
res = {}
shapes = {"qkv": ((Hq + 2 * Hkv) * D, d, 0), "o": (d, Hq * D, 1), "gate_up": (2 * F, d, 2), "down": (d, F, 1), "lm_head": (V, d, 3)}
for name, (N, K, mode) in shapes.items():
    n_sets = NL if name != "lm_head" else 2
    ws = [torch.randn(N, K, device=dev).to(bf) * 0.02 for _ in range(n_sets)]
    x = torch.randn(R, K, device=dev).to(bf)
    r = torch.randn(R, N, device=dev).to(bf)
    ssq = torch.ones(32, device=dev)
    def fn():
        for w in ws:
            ops.skinny_gemm(x, w, scratch, mode=mode, residual=r if mode == 1 else None, sumsq_in=ssq if mode in (0, 2, 3) else None, eps=1e-6)
    us = timed_graph(fn) / n_sets
    gbs = N * K * 2 / (us * 1e-6) / 1e9
    res[name] = (us, gbs)
    print(f"skinny {name:8s} N={N:6d} K={K:5d}: {us:8.2f} us  {gbs:7.1f} GB/s")
    del ws

# fused attention at ctx = P + gen
PAGE = 64
T = P + gen
n_shared = P // PAGE
priv = math.ceil((T + 1 - n_shared * PAGE) / PAGE)
max_pages = n_shared + priv
n_pages = n_shared + R * priv
table = torch.zeros(R, max_pages, dtype=torch.int32)
nxt = n_shared
for r_ in range(R):
    table[r_, :n_shared] = torch.arange(n_shared, dtype=torch.int32)
    table[r_, n_shared:] = torch.arange(nxt, nxt + priv, dtype=torch.int32); nxt += priv
table = table.to(dev)
kc = torch.randn(n_pages, Hkv, PAGE, D, device=dev).to(bf); vc = torch.randn_like(kc)
qkv = torch.randn(R, (Hq + 2 * Hkv) * D, device=dev).to(bf)
qn = torch.ones(D, device=dev).to(bf); kn = torch.ones(D, device=dev).to(bf)
cur = torch.full((R,), T, dtype=torch.int32, device=dev)
rope = ops.rope_table(T + 8, D, 1e6, dev)
for ss, sp in ((8, 2), (14, 3), (16, 2), (28, 3), (28, 2), (4, 1)):
    wsf = ops.decode_fused_workspace(R, Hq, Hkv, D, ss + sp, dev)
    out = torch.empty(R, Hq * D, device=dev, dtype=bf)
    us = timed_graph(lambda: ops.decode_attn_fused(qkv, qn, kn, kc, vc, table, cur, G, Hq, Hkv, D, n_shared, ss, sp, 1e6, 1e-6, wsf, out, rope=rope), inner=4)
    print(f"decode_attn_fused ctx={T} splits=({ss},{sp}): {us:8.2f} us")
    res[f"attn_fused_{ss}_{sp}"] = (us, 0)

# phase timestamps of one fused-attention launch
from bioreason_b200._lib import lib as _lib, ffi as _ffi
items = 64 + 128
dbg = torch.zeros(items, 16, dtype=torch.int64, device=dev)
_lib().br_decode_attn_fused_debug(_ffi.cast("long long*", dbg.data_ptr()))
wsf = ops.decode_fused_workspace(R, Hq, Hkv, D, 10, dev); out = torch.empty(R, Hq * D, device=dev, dtype=bf)
for _ in range(3):
    ops.decode_attn_fused(qkv, qn, kn, kc, vc, table, cur, G, Hq, Hkv, D, n_shared, 8, 2, 1e6, 1e-6, wsf, out, rope=rope)
torch.cuda.synchronize()
_lib().br_decode_attn_fused_debug(_ffi.NULL)
t = dbg.double().cpu(); t0 = t[:, 0].min()
rel = (t - t0) / 1e3
names = ["start", "dep_wait", "q_prep", "tile0", "tiles", "partials", "counter", "end", "q_loaded", "q_roped"]
for lab, sl in (("shared", slice(0, 64)), ("private", slice(64, 192))):
    print(f"fused attention phases ({lab}, us since first CTA start): " + "  ".join(f"{n}={rel[sl, i].mean():.1f}/{rel[sl, i].max():.1f}" for i, n in enumerate(names)))

# sampler
logits = torch.randn(R, V, device=dev)
tokens = torch.zeros(R, 4, dtype=torch.int64, device=dev); nx = torch.zeros(R, dtype=torch.int64, device=dev)
fin = torch.zeros(R, dtype=torch.int32, device=dev); step = torch.zeros(1, dtype=torch.int32, device=dev); uni = torch.rand(4, R, device=dev)
us = timed_graph(lambda: ops.sample_next(logits, temperature=0.6, top_k=20, top_p=0.95, do_sample=True, uniforms=uni, step=step, max_steps=4,
                                          finished=fin, tokens=tokens, next_ids=nx), inner=2)
print(f"sampler 1-stage: {us:8.2f} us")
if hasattr(ops, "sample_workspace"):
    sw = ops.sample_workspace(R, V, dev)
    us = timed_graph(lambda: ops.sample_next(logits, temperature=0.6, top_k=20, top_p=0.95, do_sample=True, uniforms=uni, step=step, max_steps=4,
                                              finished=fin, tokens=tokens, next_ids=nx, workspace=sw), inner=2)
    print(f"sampler 2-stage: {us:8.2f} us")
emb = torch.randn(V, d, device=dev).to(bf); h = torch.empty(R, d, device=dev, dtype=bf); ssq = torch.zeros(32, device=dev)
us = timed_graph(lambda: ops.embed_gather_sumsq(nx, emb, h, ssq), inner=4)
print(f"embed_gather_sumsq: {us:8.2f} us")
us = timed_graph(lambda: ops.decode_advance(step, cur), inner=4)
print(f"decode_advance: {us:8.2f} us")
tot = 36 * (res["qkv"][0] + res["o"][0] + res["gate_up"][0] + res["down"][0] + res["attn_fused_8_2"][0]) + res["lm_head"][0]
print(f"estimated token step (36 layers): {tot / 1e3:.3f} ms")

# ---- a real layer chain (qkv -> fused attention -> o -> gate/up -> down) x NL in one graph: what a token step actually costs
ws_l = [dict(qkv=torch.randn((Hq + 2 * Hkv) * D, d, device=dev).to(bf) * 0.02, o=torch.randn(d, Hq * D, device=dev).to(bf) * 0.02,
             gu=torch.randn(2 * F, d, device=dev).to(bf) * 0.02, down=torch.randn(d, F, device=dev).to(bf) * 0.02) for _ in range(NL)]
x0 = torch.randn(R, d, device=dev).to(bf)
cur.fill_(T); step.zero_()          # decode_advance above moved them
n_part = ((d + 127) // 128) * 4
ssa = torch.ones(n_part, 32, device=dev); ssb = torch.ones(n_part, 32, device=dev)
wsf = ops.decode_fused_workspace(R, Hq, Hkv, D, 10, dev)
attn_out = torch.empty(R, Hq * D, device=dev, dtype=bf)
SS_, SP_ = int(os.environ.get("BR_ATTN_SS", 8)), int(os.environ.get("BR_ATTN_SP", 2))
wsf = ops.decode_fused_workspace(R, Hq, Hkv, D, SS_ + SP_, dev)
for frac in (0.0, 0.5, 1.0):
    def stage(w, lo, hi):
        if frac <= 0:
            return None
        span = max(0, ops.skinny_chunk_units(w) - 6) * frac
        return (w, 6 + int(span * lo), 6 + int(span * hi))
    def chain():
        x = x0
        for i, w in enumerate(ws_l):
            wn = ws_l[(i + 1) % NL]["qkv"]
            q = ops.skinny_gemm(x, w["qkv"], scratch, sumsq_in=ssa, sumsq_in_n=n_part, eps=1e-6, prefetch=stage(w["o"], 0, 1))
            ops.decode_attn_fused(q, qn, kn, kc, vc, table, cur, G, Hq, Hkv, D, n_shared, SS_, SP_, 1e6, 1e-6, wsf, attn_out, rope=rope, prefetch=stage(w["gu"], 0, 0.65))
            x2 = ops.skinny_gemm(attn_out, w["o"], scratch, mode=1, residual=x, sumsq_out=ssb, prefetch=stage(w["gu"], 0.65, 1))
            a = ops.skinny_gemm(x2, w["gu"], scratch, mode=2, sumsq_in=ssb, sumsq_in_n=n_part, eps=1e-6, prefetch=stage(w["down"], 0, 1))
            x = ops.skinny_gemm(a, w["down"], scratch, mode=1, residual=x2, sumsq_out=ssa, prefetch=stage(wn, 0, 1))
    us = timed_graph(chain) / NL
    print(f"layer chain (5 launches, splits {SS_},{SP_}, L2 staging {frac:.1f}): {us:8.2f} us per layer  -> {36 * us / 1e3:.3f} ms per token (36 layers)   PDL={'off' if os.environ.get('BR_NO_PDL') else 'on'}")

# ---- the production layout: fused attention + ONE persistent chain kernel (o -> gate/up -> down -> next qkv) per layer
b_qkv = torch.empty(R, (Hq + 2 * Hkv) * D, device=dev, dtype=bf); b_x2 = torch.empty(R, d, device=dev, dtype=bf)
b_act = torch.empty(R, F, device=dev, dtype=bf); hbuf = x0.clone()
wsf2 = ops.decode_fused_workspace(R, Hq, Hkv, D, 10, dev)
def chain2():
    for i, w in enumerate(ws_l):
        ops.decode_attn_fused(b_qkv, qn, kn, kc, vc, table, cur, G, Hq, Hkv, D, n_shared, 8, 2, 1e6, 1e-6, wsf2, attn_out, rope=rope)
        ops.skinny_chain([dict(x=attn_out, w=w["o"], out=b_x2, mode=1, residual=hbuf, sumsq_out=ssb),
                          dict(x=b_x2, w=w["gu"], out=b_act, mode=2, sumsq_in=ssb, sumsq_in_n=n_part),
                          dict(x=b_act, w=w["down"], out=hbuf, mode=1, residual=b_x2, sumsq_out=ssa),
                          dict(x=hbuf, w=ws_l[(i + 1) % NL]["qkv"], out=b_qkv, mode=0, sumsq_in=ssa, sumsq_in_n=n_part)], R, scratch, eps=1e-6)
us = timed_graph(chain2) / NL
print(f"layer = fused attention + persistent 4-phase GEMM chain: {us:8.2f} us per layer -> {36 * us / 1e3:.3f} ms per token; "
      f"weight bytes {sum(v.numel() for v in ws_l[0].values()) * 2 / 1e6:.1f} MB -> {sum(v.numel() for v in ws_l[0].values()) * 2 / (us * 1e-6) / 1e9:.0f} GB/s incl. attention")
def chain_only():
    for i, w in enumerate(ws_l):
        ops.skinny_chain([dict(x=attn_out, w=w["o"], out=b_x2, mode=1, residual=hbuf, sumsq_out=ssb),
                          dict(x=b_x2, w=w["gu"], out=b_act, mode=2, sumsq_in=ssb, sumsq_in_n=n_part),
                          dict(x=b_act, w=w["down"], out=hbuf, mode=1, residual=b_x2, sumsq_out=ssa),
                          dict(x=hbuf, w=ws_l[(i + 1) % NL]["qkv"], out=b_qkv, mode=0, sumsq_in=ssa, sumsq_in_n=n_part)], R, scratch, eps=1e-6)
us = timed_graph(chain_only) / NL
print(f"persistent 4-phase GEMM chain alone: {us:8.2f} us per layer  ({sum(v.numel() for v in ws_l[0].values()) * 2 / (us * 1e-6) / 1e9:.0f} GB/s)")
