"""A handful of representative launches for `ncu --set full`: the big decoder GEMMs (tensor-bound) and the decode weight
streaming GEMMs (HBM-bound) at config (c) shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bioreason_b200 import ops
dev, bf = "cuda", torch.bfloat16
M, d, F, NQ, V = 8 * 2360, 2560, 9728, 6144, 151936
mk = lambda *s: (torch.randn(*s, device=dev) * 0.02).to(bf)
x = mk(M, d); w_qkv = mk(NQ, d); w_gu = mk(2 * F, d); act = mk(M, F); w_down = mk(d, F); res = mk(M, d)
emb = mk(V, d); h_sel = mk(4096, d); tgt = torch.randint(0, V, (4096,), device=dev)
for _ in range(2):                                       # launches 0-7: warm-up + the ones ncu keeps (use -s 4 -c 4)
    ops.gemm(x, w_qkv)                                   # [18880 x 2560] x [6144 x 2560]^T
    ops.gemm(x, w_gu, act=1)                             # gate/up with the fused SwiGLU epilogue
    ops.gemm(act, w_down, residual=res)                  # down_proj + residual
    ops.lmhead_logprob(h_sel, emb, tgt)                  # fused lm_head + LSE (logits never in HBM)
torch.cuda.synchronize()
R = 8
xs = mk(R, d); scratch = ops.skinny_scratch(V, dev); ssq = torch.ones(80, 32, device=dev)
for _ in range(2):                                       # skinny launches: lm_head, gate/up, qkv
    ops.skinny_gemm(xs, emb, scratch, mode=3, sumsq_in=ssq, sumsq_in_n=80, eps=1e-6)
    ops.skinny_gemm(xs, w_gu, scratch, mode=2, sumsq_in=ssq, sumsq_in_n=80, eps=1e-6)
    ops.skinny_gemm(xs, w_qkv, scratch, sumsq_in=ssq, sumsq_in_n=80, eps=1e-6)
torch.cuda.synchronize()
print("done")
