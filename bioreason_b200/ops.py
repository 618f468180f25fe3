"""Thin Python wrappers: PyTorch tensors in, PyTorch tensors out, math in libbioreason_b200 (C ABI).

PyTorch here is plumbing only: it owns device memory and streams.  Every op raises if its tensors
are not CUDA tensors -- there is no CPU path.
"""
from __future__ import annotations

import os
from typing import Optional

import torch

from ._lib import COUNTER as LAUNCHES_RAW, check, ffi, lib, ptr

BF16, F32 = 0, 1
LAUNCHES = LAUNCHES_RAW   # kernels launched through the C ABI (bench.py's gpu_launches); graph replays add their captured count


def _stream():
    return ffi.cast("void*", torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("bioreason_b200 ops run on CUDA tensors only (no CPU fallback)")


# ------------------------------------------------------------------ GRPO
def grpo_advantages(rewards_per_func: torch.Tensor, num_generations: int, return_stats: bool = False):
    """grpo_trainer.py:682-692."""
    _need_cuda(rewards_per_func)
    r = rewards_per_func.float().contiguous()
    rows, nf = r.shape
    adv = torch.empty(rows, device=r.device, dtype=torch.float32)
    gm = torch.empty(rows // num_generations, device=r.device, dtype=torch.float32)
    gs = torch.empty_like(gm)
    check(lib().br_grpo_advantages(ptr(r, "float*"), rows, nf, num_generations, ptr(adv, "float*"),
                                   ptr(gm, "float*"), ptr(gs, "float*"), _stream()), "grpo_advantages")
    return (adv, gm, gs) if return_stats else adv


def grpo_loss_raw(lp, old_lp, ref_lp, adv, mask, beta, eps_low, eps_high, want_grad=True):
    _need_cuda(lp, adv, mask)
    B, C = lp.shape
    lp = lp.float().contiguous()
    old_lp = None if old_lp is None else old_lp.float().contiguous()
    ref_lp = None if ref_lp is None else ref_lp.float().contiguous()
    adv = adv.float().contiguous()
    mask = mask.to(torch.int32).contiguous()
    out3 = torch.empty(3, device=lp.device, dtype=torch.float32)
    dlp = torch.empty_like(lp) if want_grad else None
    check(lib().br_grpo_loss_fwd_bwd(ptr(lp, "float*"), ptr(old_lp, "float*"), ptr(ref_lp, "float*"), ptr(adv, "float*"),
                                     ptr(mask, "int32_t*"), B, C, float(beta), float(eps_low), float(eps_high),
                                     ptr(out3, "float*"), ptr(dlp, "float*"), _stream()), "grpo_loss")
    return out3, dlp


class _GRPOLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, lp, old_lp, ref_lp, adv, mask, beta, eps_low, eps_high):
        out3, dlp = grpo_loss_raw(lp.detach(), old_lp, ref_lp, adv, mask, beta, eps_low, eps_high)
        ctx.save_for_backward(dlp)
        ctx.mark_non_differentiable(out3)
        return out3[0].clone(), out3

    @staticmethod
    def backward(ctx, g, _g3):
        (dlp,) = ctx.saved_tensors
        return dlp * g, None, None, None, None, None, None, None


def grpo_loss(lp, old_lp, ref_lp, adv, mask, beta=0.04, eps_low=0.2, eps_high=0.2):
    """grpo_trainer.py:786-812 -> (loss, out3=[loss, mean_kl, clip_ratio]); differentiable w.r.t. lp."""
    return _GRPOLoss.apply(lp, old_lp, ref_lp, adv, mask, beta, eps_low, eps_high)


def eos_mask(completion_ids: torch.Tensor, eos_id: int) -> torch.Tensor:
    """grpo_trainer.py:605-609."""
    _need_cuda(completion_ids)
    ids = completion_ids.to(torch.int64).contiguous()
    B, C = ids.shape
    m = torch.empty(B, C, device=ids.device, dtype=torch.int32)
    check(lib().br_eos_mask(ptr(ids, "int64_t*"), B, C, int(eos_id), ptr(m, "int32_t*"), _stream()), "eos_mask")
    return m


# ------------------------------------------------------------------ GEMM
def _row_major_2d(t):
    assert t.dim() == 2 and t.stride(1) == 1, "need a 2-D tensor with contiguous last dim"
    return t.stride(0)


def gemm(a: torch.Tensor, b: torch.Tensor, *, bias=None, residual=None, alpha: float = 1.0, act: int = 0,
         out: Optional[torch.Tensor] = None, out_dtype=torch.bfloat16, row_map=None, aux_out=None,
         a2=None, b2=None) -> torch.Tensor:
    """out[M,N] = epilogue(a[M,K] @ b[N,K].T (+ a2[M,K2] @ b2[N,K2].T)) on tcgen05 (bf16 in, fp32 accumulate)."""
    _need_cuda(a, b)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    M, K = a.shape
    N, Kb = b.shape
    assert K == Kb
    n_out = N // 2 if act == 1 else N
    if out is None:
        out = torch.empty(M, n_out, device=a.device, dtype=out_dtype)
    e = ffi.new("br_gemm_epilogue*")
    e.alpha = alpha
    e.act = act
    e.out_dtype = F32 if out.dtype == torch.float32 else BF16
    keep = [a, b, out]
    if bias is not None:
        e.bias = ptr(bias); e.bias_dtype = F32 if bias.dtype == torch.float32 else BF16; keep.append(bias)
    if residual is not None:
        assert residual.dtype == torch.bfloat16
        e.residual = ptr(residual); e.ldr = _row_major_2d(residual); keep.append(residual)
    if row_map is not None:
        assert row_map.dtype == torch.int32
        e.row_map = ptr(row_map, "int32_t*"); keep.append(row_map)
    if aux_out is not None:
        e.aux_out = ptr(aux_out); e.ld_aux = _row_major_2d(aux_out); keep.append(aux_out)
    if a2 is not None:
        assert a2.dtype == torch.bfloat16 and b2.dtype == torch.bfloat16 and a2.shape[0] == M and b2.shape[0] == N
        e.A2 = ptr(a2); e.lda2 = _row_major_2d(a2); e.B2 = ptr(b2); e.ldb2 = _row_major_2d(b2); e.K2 = a2.shape[1]
        keep += [a2, b2]
    check(lib().br_gemm_bf16(ptr(a), _row_major_2d(a), ptr(b), _row_major_2d(b), ptr(out), _row_major_2d(out),
                             M, N, K, e, _stream()), "gemm_bf16")
    return out


def lmhead_logprob(h: torch.Tensor, w: torch.Tensor, target: torch.Tensor, scale: float = 1.0):
    """Fused lm_head + log_softmax + gather (grpo_trainer.py:511-520): returns (logp[M], lse[M]) fp32."""
    _need_cuda(h, w, target)
    M, K = h.shape
    V = w.shape[0]
    tgt = target.to(torch.int32).contiguous()
    ws = torch.empty(lib().br_lmhead_workspace_bytes(M, V), device=h.device, dtype=torch.uint8)
    logp = torch.empty(M, device=h.device, dtype=torch.float32)
    lse = torch.empty(M, device=h.device, dtype=torch.float32)
    check(lib().br_lmhead_logprob_fwd(ptr(h), _row_major_2d(h), ptr(w), _row_major_2d(w), ptr(tgt, "int32_t*"), M, V, K,
                                      float(scale), ptr(logp, "float*"), ptr(lse, "float*"), ptr(ws), _stream()),
          "lmhead_logprob_fwd")
    return logp, lse


def lmhead_dlogits(h, w, target, lse, gscale, scale: float = 1.0):
    """bf16 [M, V] tile-recomputed gradient of sum_m gscale[m] * logp[m] w.r.t. the logits."""
    _need_cuda(h, w, target)
    M, K = h.shape
    V = w.shape[0]
    tgt = target.to(torch.int32).contiguous()
    d = torch.empty(M, V, device=h.device, dtype=torch.bfloat16)
    check(lib().br_lmhead_dlogits(ptr(h), _row_major_2d(h), ptr(w), _row_major_2d(w), ptr(tgt, "int32_t*"),
                                  ptr(lse, "float*"), ptr(gscale.float().contiguous(), "float*"), M, V, K, float(scale),
                                  ptr(d), V, _stream()), "lmhead_dlogits")
    return d


# ------------------------------------------------------------------ row kernels
def rmsnorm(x, w, eps: float, out=None, want_rstd: bool = False):
    _need_cuda(x, w)
    M, d = x.shape
    if out is None:
        out = torch.empty(M, d, device=x.device, dtype=torch.bfloat16)
    rstd = torch.empty(M, device=x.device, dtype=torch.float32) if want_rstd else None
    check(lib().br_rmsnorm(ptr(x), _row_major_2d(x), ptr(w), ptr(out), _row_major_2d(out), ptr(rstd, "float*"), M, d, float(eps),
                           _stream()), "rmsnorm")
    return (out, rstd) if want_rstd else out


def layernorm(x, w, b, eps: float, out=None):
    _need_cuda(x, w, b)
    M, d = x.shape
    if out is None:
        out = torch.empty(M, d, device=x.device, dtype=torch.bfloat16)
    check(lib().br_layernorm(ptr(x), _row_major_2d(x), ptr(w), ptr(b), ptr(out), _row_major_2d(out), M, d, float(eps), _stream()),
          "layernorm")
    return out


def qk_rope_(qkv, n_q, n_k, head_dim, positions, theta, *, q_norm_w=None, k_norm_w=None, eps=1e-6, q_scale=1.0, mode=0, out=None, rope=None):
    """On the fused QKV buffer [M, >= (n_q+n_k)*head_dim]: in place, or (out=) into a separate [M, >= (n_q+n_k)*head_dim] buffer so the
    pre-norm values stay available for the backward.  rope: optional cos/sin table from rope_table() (decoder rows, mode 0)."""
    _need_cuda(qkv, positions)
    assert positions.dtype == torch.int32 and positions.numel() == qkv.shape[0]
    check(lib().br_qk_rope_ex(ptr(qkv), _row_major_2d(qkv), ptr(out), _row_major_2d(out) if out is not None else 0, qkv.shape[0], n_q, n_k, head_dim,
                              ptr(q_norm_w), ptr(k_norm_w), ptr(positions, "int32_t*"), float(theta), float(eps), float(q_scale), mode,
                              ptr(rope, "float*"), rope.shape[0] if rope is not None else 0, _stream()), "qk_rope")
    return qkv if out is None else out


def embed_gather(ids, table, keep=None, out=None):
    _need_cuda(ids, table)
    ids = ids.reshape(-1).to(torch.int64).contiguous()
    M, d = ids.numel(), table.shape[1]
    if out is None:
        out = torch.empty(M, d, device=table.device, dtype=torch.bfloat16)
    if keep is not None:
        keep = keep.reshape(-1).to(torch.int32).contiguous()
    check(lib().br_embed_gather(ptr(ids, "int64_t*"), ptr(table), _row_major_2d(table), table.shape[0], ptr(out),
                                _row_major_2d(out), M, d, ptr(keep, "int32_t*"), _stream()), "embed_gather")
    return out


def scatter_rows_(dst, src, row_map):
    _need_cuda(dst, src, row_map)
    assert row_map.dtype == torch.int32
    check(lib().br_scatter_rows(ptr(src), _row_major_2d(src), ptr(row_map, "int32_t*"), ptr(dst), _row_major_2d(dst),
                                src.shape[0], src.shape[1], _stream()), "scatter_rows")
    return dst


def gather_rows(src, idx, out=None):
    _need_cuda(src, idx)
    assert idx.dtype == torch.int32
    if out is None:
        out = torch.empty(idx.numel(), src.shape[1], device=src.device, dtype=torch.bfloat16)
    check(lib().br_gather_rows(ptr(src), _row_major_2d(src), ptr(idx, "int32_t*"), ptr(out), _row_major_2d(out),
                               idx.numel(), src.shape[1], _stream()), "gather_rows")
    return out


# ------------------------------------------------------------------ attention
def attn_fwd(q, k, v, B, L, n_q, n_kv, head_dim, *, kv_start=None, kv_end=None, scale=None, causal=True, want_lse=False, out=None):
    """q/k/v: 2-D views [B*L, heads*head_dim] (may alias one fused QKV buffer)."""
    _need_cuda(q, k, v)
    if out is None:
        out = torch.empty(B * L, n_q * head_dim, device=q.device, dtype=torch.bfloat16)
    lse = torch.empty(B, n_q, L, device=q.device, dtype=torch.float32) if want_lse else None
    if scale is None:
        scale = head_dim ** -0.5
    check(lib().br_attn_fwd(ptr(q), _row_major_2d(q), ptr(k), _row_major_2d(k), ptr(v), _row_major_2d(v), ptr(out), _row_major_2d(out),
                            ptr(lse, "float*"), B, L, n_q, n_kv, head_dim, ptr(kv_start, "int32_t*"), ptr(kv_end, "int32_t*"),
                            float(scale), 1 if causal else 0, _stream()), "attn_fwd")
    return (out, lse) if want_lse else out


# ------------------------------------------------------------------ decode
def skinny_scratch(max_n: int, device) -> torch.Tensor:
    return torch.zeros(lib().br_skinny_scratch_bytes(max_n), device=device, dtype=torch.uint8)


def _l2_prefetch(spec):
    """(W_later, unit_lo, unit_hi) -> br_l2_prefetch* (or NULL): stage units [lo, hi) of every stream-K chunk of a later decode GEMM into L2."""
    if spec is None:
        return ffi.NULL, None
    w, lo, hi = spec
    if hi <= lo:
        return ffi.NULL, None
    pf = ffi.new("br_l2_prefetch*")
    pf.W = ptr(w); pf.ldw = _row_major_2d(w); pf.N = w.shape[0]; pf.K = w.shape[1]; pf.unit_lo = int(lo); pf.unit_hi = int(hi)
    return pf, w


def _gate(spec):
    """dict(counter, epoch, epoch_base, per_step, wait (cumulative arrivals or None), signal (bool)) -> br_stream_gate* (or NULL)."""
    if spec is None:
        return ffi.NULL
    g = ffi.new("br_stream_gate*")
    g.counter = ptr(spec["counter"], "int32_t*"); g.epoch = ptr(spec["epoch"], "int32_t*")
    g.epoch_base = int(spec.get("epoch_base", 0)); g.per_step = int(spec["per_step"])
    g.wait_prefix = -1 if spec.get("wait") is None else int(spec["wait"])
    g.signal = 1 if spec.get("signal", True) else 0
    return g


def skinny_grid(w) -> int:
    """CTAs skinny_gemm launches for weight w [N, K]."""
    return lib().br_skinny_grid(w.shape[0], w.shape[1])


def skinny_chunk_units(w) -> int:
    """16 KB weight tiles one CTA of skinny_gemm streams for weight w [N, K] (the stream-K chunk; mirrors br_skinny_gemm_ex)."""
    n_sms = torch.cuda.get_device_properties(w.device).multi_processor_count
    units = ((w.shape[0] + 127) // 128) * ((w.shape[1] + 63) // 64)
    grid = min(units, n_sms)
    return (units + grid - 1) // grid


def skinny_gemm(x, w, scratch, *, mode=0, residual=None, out=None, sumsq_in=None, sumsq_in_n=1, sumsq_out=None, eps=0.0, prefetch=None, gate=None):
    """out[R, N] = x[R, K] @ w[N, K].T for R <= 32 decode rows (optionally with the folded-RMSNorm statistics).
    prefetch=(W_later, unit_lo, unit_hi): also stage tiles of a later GEMM of the chain into L2 (see br_l2_prefetch)."""
    _need_cuda(x, w)
    R, K = x.shape
    N = w.shape[0]
    if out is None:
        if mode == 3:
            out = torch.empty(R, N, device=x.device, dtype=torch.float32)
        else:
            out = torch.empty(R, N // 2 if mode == 2 else N, device=x.device, dtype=torch.bfloat16)
    pf, _keep = _l2_prefetch(prefetch)
    check(lib().br_skinny_gemm_gated(ptr(x), _row_major_2d(x), ptr(w), _row_major_2d(w), ptr(out), _row_major_2d(out), R, N, K, mode,
                                  ptr(residual), _row_major_2d(residual) if residual is not None else 0, ptr(scratch),
                                  ptr(sumsq_in, "float*"), int(sumsq_in_n) if sumsq_in is not None else 0, ptr(sumsq_out, "float*"),
                                  float(eps), pf, _gate(gate), _stream()),
          "skinny_gemm")
    return out


def skinny_chain(phases, R, scratch, eps=0.0):
    """phases: list of dicts(x, w, out, mode=0, residual=None, sumsq_in=None, sumsq_in_n=1, sumsq_out=None) -- up to 4
    dependent decode GEMMs in one persistent launch (grid barrier between phases, weights prefetched across it)."""
    n = len(phases)
    arr = ffi.new("br_skinny_phase[]", n)
    keep = []
    for i, ph in enumerate(phases):
        x, w, out = ph["x"], ph["w"], ph["out"]
        a = arr[i]
        a.X = ptr(x); a.ldx = _row_major_2d(x); a.W = ptr(w); a.ldw = _row_major_2d(w); a.out = ptr(out); a.ldo = _row_major_2d(out)
        a.N = w.shape[0]; a.K = w.shape[1]; a.mode = ph.get("mode", 0)
        res = ph.get("residual")
        a.residual = ptr(res); a.ldr = _row_major_2d(res) if res is not None else 0
        si = ph.get("sumsq_in")
        a.sumsq_in = ptr(si, "float*"); a.sumsq_in_n = int(ph.get("sumsq_in_n", 1)) if si is not None else 0
        a.sumsq_out = ptr(ph.get("sumsq_out"), "float*")
        keep += [x, w, out, res, si]
    check(lib().br_skinny_chain(arr, n, R, float(eps), ptr(scratch), _stream()), "skinny_chain")


def embed_gather_sumsq(ids, table, out, sumsq):
    ids = ids.reshape(-1)
    assert ids.dtype == torch.int64
    check(lib().br_embed_gather_sumsq(ptr(ids, "int64_t*"), ptr(table), _row_major_2d(table), table.shape[0], ptr(out), _row_major_2d(out),
                                      ids.numel(), table.shape[1], ptr(sumsq, "float*"), _stream()), "embed_gather_sumsq")
    return out


def scale_columns_(w, scale):
    check(lib().br_scale_columns(ptr(w), _row_major_2d(w), w.shape[0], w.shape[1], ptr(scale), _stream()), "scale_columns")
    return w


def decode_rope_append(qkv, n_q, n_kv, head_dim, q_norm_w, k_norm_w, cur_len, page_table, kcache, vcache, theta, eps):
    check(lib().br_decode_rope_append(ptr(qkv), _row_major_2d(qkv), qkv.shape[0], n_q, n_kv, head_dim, ptr(q_norm_w), ptr(k_norm_w),
                                      ptr(cur_len, "int32_t*"), ptr(page_table, "int32_t*"), page_table.shape[1], ptr(kcache),
                                      ptr(vcache), float(theta), float(eps), _stream()), "decode_rope_append")


def kv_write_pages(qkv_from_first_token, n_tok, n_q, n_kv, head_dim, pages, kcache, vcache):
    check(lib().br_kv_write_pages(ptr(qkv_from_first_token), _row_major_2d(qkv_from_first_token), n_tok, n_q, n_kv, head_dim,
                                  ptr(pages, "int32_t*"), ptr(kcache), ptr(vcache), _stream()), "kv_write_pages")


def decode_attn_workspace(R, n_q, head_dim, n_slots, device):
    return torch.empty(lib().br_decode_attn_workspace_bytes(R, n_q, head_dim, n_slots), device=device, dtype=torch.uint8)


def decode_attn(qkv, kcache, vcache, page_table, cur_len, G, n_q, n_kv, head_dim, n_shared_pages, splits_shared, splits_private,
                workspace, out, scale=None):
    R = qkv.shape[0]
    if scale is None:
        scale = head_dim ** -0.5
    check(lib().br_decode_attn(ptr(qkv), _row_major_2d(qkv), ptr(kcache), ptr(vcache), ptr(page_table, "int32_t*"), page_table.shape[1],
                               ptr(cur_len, "int32_t*"), R, G, n_q, n_kv, head_dim, n_shared_pages, splits_shared, splits_private,
                               float(scale), ptr(workspace), ptr(out), _row_major_2d(out), _stream()), "decode_attn")
    return out


def decode_fused_workspace(R, n_q, n_kv, head_dim, n_slots, device):
    return torch.zeros(lib().br_decode_fused_workspace_bytes(R, n_q, n_kv, head_dim, n_slots), device=device, dtype=torch.uint8)


def rope_table(n_pos, head_dim, theta, device):
    t = torch.empty(n_pos, head_dim // 2, 2, device=device, dtype=torch.float32)
    check(lib().br_rope_table(ptr(t, "float*"), n_pos, head_dim, float(theta), _stream()), "rope_table")
    return t


def decode_attn_fused(qkv_raw, q_norm_w, k_norm_w, kcache, vcache, page_table, cur_len, G, n_q, n_kv, head_dim, n_shared_pages,
                      splits_shared, splits_private, theta, eps, workspace, out, scale=None, rope=None, prefetch=None):
    R = qkv_raw.shape[0]
    if scale is None:
        scale = head_dim ** -0.5
    pf, _keep = _l2_prefetch(prefetch)
    check(lib().br_decode_attn_fused_pf(ptr(qkv_raw), _row_major_2d(qkv_raw), ptr(q_norm_w), ptr(k_norm_w), ptr(kcache), ptr(vcache),
                                        ptr(page_table, "int32_t*"), page_table.shape[1], ptr(cur_len, "int32_t*"), R, G, n_q, n_kv,
                                        head_dim, n_shared_pages, splits_shared, splits_private, float(scale), float(theta), float(eps),
                                        ptr(rope, "float*"), rope.shape[0] if rope is not None else 0,
                                        ptr(workspace), ptr(out), _row_major_2d(out), pf, _stream()), "decode_attn_fused")
    return out


def sample_workspace(R, V, device):
    return torch.empty(lib().br_sample_workspace_bytes(R, V), device=device, dtype=torch.uint8)


def sample_next(logits, *, temperature=1.0, top_k=20, top_p=1.0, do_sample=True, uniforms=None, step=None, max_steps=1,
                eos_id=-1, pad_id=0, finished=None, tokens=None, next_ids=None, workspace=None):
    R, V = logits.shape
    assert logits.dtype == torch.float32
    if workspace is not None and (not do_sample or top_k <= 32) and not os.environ.get("BR_SAMPLER_1STAGE"):
        check(lib().br_sample_next_2stage(ptr(logits, "float*"), _row_major_2d(logits), R, V, float(temperature), int(top_k), float(top_p),
                                          1 if do_sample else 0, ptr(uniforms, "float*"), ptr(step, "int32_t*"), int(max_steps), int(eos_id),
                                          int(pad_id), ptr(finished, "int32_t*"), ptr(tokens, "int64_t*"), ptr(next_ids, "int64_t*"),
                                          ptr(workspace), _stream()), "sample_next_2stage")
        return
    check(lib().br_sample_next(ptr(logits, "float*"), _row_major_2d(logits), R, V, float(temperature), int(top_k), float(top_p),
                               1 if do_sample else 0, ptr(uniforms, "float*"), ptr(step, "int32_t*"), int(max_steps), int(eos_id),
                               int(pad_id), ptr(finished, "int32_t*"), ptr(tokens, "int64_t*"), ptr(next_ids, "int64_t*"), _stream()),
          "sample_next")


def decode_advance(step, cur_len):
    check(lib().br_decode_advance(ptr(step, "int32_t*"), ptr(cur_len, "int32_t*"), cur_len.numel(), _stream()), "decode_advance")


# ------------------------------------------------------------------ backward
def attn_bwd(q, k, v, o, dout, lse, dq, dk, dv, B, L, n_q, n_kv, head_dim, *, kv_start=None, kv_end=None, scale=None):
    if scale is None:
        scale = head_dim ** -0.5
    ws = torch.empty(lib().br_attn_bwd_workspace_bytes(B, L, n_q, head_dim), device=q.device, dtype=torch.uint8)
    check(lib().br_attn_bwd(ptr(q), _row_major_2d(q), ptr(k), _row_major_2d(k), ptr(v), _row_major_2d(v), ptr(o), _row_major_2d(o),
                            ptr(dout), _row_major_2d(dout), ptr(lse, "float*"), ptr(dq), _row_major_2d(dq), ptr(dk), _row_major_2d(dk),
                            ptr(dv), _row_major_2d(dv), B, L, n_q, n_kv, head_dim, ptr(kv_start, "int32_t*"), ptr(kv_end, "int32_t*"),
                            float(scale), ptr(ws), _stream()), "attn_bwd")


def rmsnorm_bwd(x, w, rstd, dy, dres=None, out=None):
    M, d = x.shape
    if out is None:
        out = torch.empty(M, d, device=x.device, dtype=torch.bfloat16)
    check(lib().br_rmsnorm_bwd(ptr(x), _row_major_2d(x), ptr(w), ptr(rstd, "float*"), ptr(dy), _row_major_2d(dy), ptr(dres),
                               _row_major_2d(dres) if dres is not None else 0, ptr(out), _row_major_2d(out), M, d, _stream()), "rmsnorm_bwd")
    return out


def swiglu_bwd(gu, dact, out=None):
    M, F2 = gu.shape
    if out is None:
        out = torch.empty(M, F2, device=gu.device, dtype=torch.bfloat16)
    check(lib().br_swiglu_bwd(ptr(gu), _row_major_2d(gu), ptr(dact), _row_major_2d(dact), ptr(out), _row_major_2d(out), M, F2 // 2,
                              _stream()), "swiglu_bwd")
    return out


def qk_rope_bwd_(dqkv, qk_pre, n_q, n_k, head_dim, q_norm_w, k_norm_w, positions, theta, eps):
    check(lib().br_qk_rope_bwd(ptr(dqkv), _row_major_2d(dqkv), ptr(qk_pre), _row_major_2d(qk_pre), dqkv.shape[0], n_q, n_k, head_dim,
                               ptr(q_norm_w), ptr(k_norm_w), ptr(positions, "int32_t*"), float(theta), float(eps), _stream()), "qk_rope_bwd")
    return dqkv


_LORA_WS = {}


def lora_grad_tn(big, small, segs, *, mode=0):
    """Deterministic tcgen05 TN GEMM: product[P, N] = big[M, P]^T @ small[M, N]; the blocks named by `segs` are ADDED into fp32 views.
    segs: list of (dst fp32 2-D view with contiguous rows, row_lo, row_hi, col_lo, n_cols); mode 1: one segment, dst[n, p] (transposed);
    mode 2: gate/up-blocked product rows (segment 0 = gate rows, 1 = up rows)."""
    _need_cuda(big, small)
    assert big.dtype == torch.bfloat16 and small.dtype == torch.bfloat16 and big.shape[0] == small.shape[0]
    M, P = big.shape
    N = small.shape[1]
    dev = big.device
    ws = _LORA_WS.get(dev)
    if ws is None:
        ws = _LORA_WS[dev] = torch.zeros(lib().br_lora_grad_workspace_bytes(), device=dev, dtype=torch.uint8)
    arr = ffi.new("br_lora_grad_seg[]", len(segs))
    for i, (dst, row_lo, row_hi, col_lo, n_cols) in enumerate(segs):
        assert dst.dtype == torch.float32 and dst.dim() == 2 and dst.stride(1) == 1
        arr[i].dst = ptr(dst, "float*"); arr[i].ld = dst.stride(0)
        arr[i].row_lo, arr[i].row_hi, arr[i].col_lo, arr[i].n_cols = int(row_lo), int(row_hi), int(col_lo), int(n_cols)
    check(lib().br_lora_grad_tn(ptr(big), _row_major_2d(big), ptr(small), _row_major_2d(small), M, P, N, int(mode), arr, len(segs), ptr(ws),
                                _stream()), "lora_grad_tn")


def transpose(x, out=None, pad_cols_to: int = 8):
    """bf16 [M, N] -> [N, M'] with M' = M rounded up to `pad_cols_to` (zero padded) so it can feed the TMA GEMM."""
    M, N = x.shape
    Mp = (M + pad_cols_to - 1) // pad_cols_to * pad_cols_to
    if out is None:
        out = torch.zeros(N, Mp, device=x.device, dtype=torch.bfloat16) if Mp != M else torch.empty(N, Mp, device=x.device, dtype=torch.bfloat16)
    check(lib().br_transpose_bf16(ptr(x), _row_major_2d(x), ptr(out), _row_major_2d(out), M, N, _stream()), "transpose")
    return out


def colsum_accumulate_(out, x):
    M, N = x.shape
    check(lib().br_colsum_accumulate(ptr(x), _row_major_2d(x), ptr(out, "float*"), M, N, _stream()), "colsum")
    return out
