"""attention forward / backward only, config (c) shape, one profiled launch each (ncu --profile-from-start off)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bioreason_b200 import ops
B, L, Hq, Hkv, D = 8, 2364, 32, 8, 128
M = B * L
torch.manual_seed(0)
qkv = (torch.randn(M, (Hq + 2 * Hkv) * D, device="cuda") * 0.5).bfloat16()
dout = (torch.randn(M, Hq * D, device="cuda") * 0.02).bfloat16()
dqkv = torch.empty_like(qkv)
q, k, v = qkv[:, :Hq * D], qkv[:, Hq * D:(Hq + Hkv) * D], qkv[:, (Hq + Hkv) * D:]
def run():
    o, lse = ops.attn_fwd(q, k, v, B, L, Hq, Hkv, D, causal=True, want_lse=True)
    ops.attn_bwd(q, k, v, o, dout, lse, dqkv[:, :Hq * D], dqkv[:, Hq * D:(Hq + Hkv) * D], dqkv[:, (Hq + Hkv) * D:], B, L, Hq, Hkv, D)
    return o
run(); torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
e[0].record(); o, lse = ops.attn_fwd(q, k, v, B, L, Hq, Hkv, D, causal=True, want_lse=True); e[1].record()
ops.attn_bwd(q, k, v, o, dout, lse, dqkv[:, :Hq * D], dqkv[:, Hq * D:(Hq + Hkv) * D], dqkv[:, (Hq + Hkv) * D:], B, L, Hq, Hkv, D); e[2].record()
torch.cuda.synchronize()
fl = 4 * Hq * D * L * L / 2 * B
print(f"attn fwd {e[0].elapsed_time(e[1]):.3f} ms = {fl / e[0].elapsed_time(e[1]) / 1e9:.0f} TF/s; bwd {e[1].elapsed_time(e[2]):.3f} ms = {3.5 * fl / e[1].elapsed_time(e[2]) / 1e9:.0f} TF/s executed ({2.5 * fl / e[1].elapsed_time(e[2]) / 1e9:.0f} TF/s algorithmic)")
torch.cuda.profiler.start(); run(); torch.cuda.synchronize(); torch.cuda.profiler.stop()
