/* libbioreason_b200 -- C ABI of the B200-native BioReason hot path.
 *
 * The reference (bowang-lab/BioReason) has no FFI: its hot path is Python calling HuggingFace/PyTorch
 * (SURVEY.md §8b).  This header is the boundary the build introduces *below* the reference's Python
 * surface (`DNALLMModel`, `DNALLMGRPOTrainer`); each entry point names the reference call site it
 * replaces.  Conventions:
 *   - plain pointers + sizes, no torch types; every pointer is a CUDA device pointer unless noted;
 *   - every call enqueues on `stream` (a cudaStream_t) and returns without synchronising;
 *   - return 0 on success, <0 on error; `br_last_error()` gives the (thread-local) message;
 *   - the caller (PyTorch) owns all buffers (including every workspace / scratch buffer named below); the library keeps no
 *     pointer past the call.  There is no communicator handle: the two collectives of the path (reward all-gather, flat
 *     gradient all-reduce) are issued by the host through torch.distributed / NCCL, not through this ABI;
 *   - tensors are (pointer, leading dimension in elements) pairs, row-major; there is no tensor struct;
 *   - bf16 = __nv_bfloat16 storage, fp32 accumulation everywhere.
 * This file is parsed by cffi (ABI mode): keep it plain C, no macros beyond the constants below.
 */
#ifndef BIOREASON_B200_H
#define BIOREASON_B200_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BR_BF16 0
#define BR_F32 1

int br_version(void);
/* copies the calling thread's last error message into buf (NUL-terminated); returns its length */
int br_last_error(char* buf, size_t n);
/* 1 if the visible device is sm_100 (B200); the library refuses to run elsewhere */
int br_device_ok(void);

/* ---------------------------------------------------------------------------------------------
 * GRPO advantages + loss  (replaces bioreason/trainer/grpo_trainer.py:682-692 and :786-812)
 * ------------------------------------------------------------------------------------------- */
/* rewards_per_func [rows, n_funcs] f32 -> advantages [rows] f32; groups are G consecutive rows. */
int br_grpo_advantages(const float* rewards_per_func, int rows, int n_funcs, int G, float* advantages,
                       float* group_mean, float* group_std, void* stream);
/* lp/old_lp/ref_lp [B, C] f32 (old_lp NULL => mu == 1; ref_lp NULL => beta == 0), adv [B], mask [B, C] int32.
 * out3 = {loss, mean_kl, clip_ratio}; dlp [B, C] = d loss / d lp (may be NULL). One launch. */
int br_grpo_loss_fwd_bwd(const float* lp, const float* old_lp, const float* ref_lp, const float* adv,
                         const int32_t* mask, int B, int C, float beta, float eps_low, float eps_high,
                         float* out3, float* dlp, void* stream);
/* completion_mask[b, t] = t <= first_eos(b) (grpo_trainer.py:605-609); ids int64 [B, C] -> mask int32 */
int br_eos_mask(const int64_t* completion_ids, int B, int C, int64_t eos_id, int32_t* mask, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Dense contractions on tcgen05 (replaces every nn.Linear / lm_head reached through
 * dna_llm.py:150-160,237-242; SURVEY.md §2.3 K1,K2,K5,K6,K7,K12)
 * ------------------------------------------------------------------------------------------- */
typedef struct br_gemm_epilogue {
    const void* bias;        /* [N] or NULL */
    int32_t bias_dtype;      /* BR_BF16 / BR_F32 */
    const void* residual;    /* bf16 [M, ldr] added after bias (indexed by OUTPUT row) or NULL */
    int64_t ldr;
    float alpha;             /* scales the accumulator first */
    int32_t act;             /* 0 none; 1: gated SiLU, columns in blocks of 16 = 8 gate | 8 up: out[:, 8b+i] = silu(acc[:, 16b+i]) * acc[:, 16b+8+i] */
    int32_t out_dtype;       /* BR_BF16 / BR_F32 */
    const int32_t* row_map;  /* optional [M]: output row of input row m (<0: dropped) -- projector scatter */
    void* aux_out;           /* act==1: optional bf16 [M, ld_aux] copy of the pre-activation accumulator */
    int64_t ld_aux;
    const void* A2;          /* optional second K segment accumulated into the same tile: */
    int64_t lda2;            /*   D += A2[M, K2] . B2[N, K2]^T   (LoRA delta) */
    const void* B2;
    int64_t ldb2;
    int32_t K2;
} br_gemm_epilogue;

/* D[M, N] = epilogue(A[M, K] . B[N, K]^T); A, B bf16 row-major (K contiguous); ld* in elements. */
int br_gemm_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, void* D, int64_t ldd,
                 int M, int N, int K, const br_gemm_epilogue* epi, void* stream);

/* Fused lm_head + log-softmax + gather (replaces grpo_trainer.py:511-520 and HF loss_utils CE):
 * logp[m] = scale*H[m].W[target[m]] - logsumexp_v(scale*H[m].W[v]); logits never reach HBM.
 * target[m] < 0 => logp 0 (ignored row).  lse [M] is kept for the backward. */
int64_t br_lmhead_workspace_bytes(int M, int V);
int br_lmhead_logprob_fwd(const void* H, int64_t ldh, const void* W, int64_t ldw, const int32_t* target,
                          int M, int V, int K, float scale, float* logp, float* lse, void* workspace, void* stream);
/* dlogits[m, v] = gscale[m] * (onehot(target[m])[v] - softmax(H[m].W)[v]) as bf16 [M, ldd] (recomputed tiles) */
int br_lmhead_dlogits(const void* H, int64_t ldh, const void* W, int64_t ldw, const int32_t* target,
                      const float* lse, const float* gscale, int M, int V, int K, float scale,
                      void* dlogits, int64_t ldd, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Row kernels (HBM-bound): norms, rotary, gathers  (replace HF Qwen3RMSNorm qwen3/modeling_qwen3.py:50-64,
 * torch LayerNorm in esm/modeling_esm.py:386-400,476-479,511, apply_rotary_pos_emb qwen3:120-150 / esm:45-55,
 * embed_tokens + masked scatter dna_llm.py:211-229)
 * ------------------------------------------------------------------------------------------- */
/* y = w * bf16(x * rsqrt(mean(x^2)+eps)); bf16 [M, d]; rstd [M] f32 optional (kept for the backward) */
int br_rmsnorm(const void* x, int64_t ldx, const void* w, void* y, int64_t ldy, float* rstd, int M, int d, float eps, void* stream);
int br_layernorm(const void* x, int64_t ldx, const void* w, const void* b, void* y, int64_t ldy, int M, int d, float eps, void* stream);
/* In place on the fused QKV activation [M, ld]: heads 0..n_q-1 are queries, the next n_k are keys.
 * mode 0 (Qwen3): per-head RMSNorm with q_norm_w / k_norm_w [head_dim] (NULL = skip) then rotate-half RoPE;
 * mode 1 (ESM/NT-v2): queries scaled by q_scale, then RoPE.  positions [M] int32 (explicit: the reference uses
 * arange over the padded row in forward() and cumsum(mask)-1 in generate()). */
int br_qk_rope(void* qkv, int64_t ld, int M, int n_q_heads, int n_k_heads, int head_dim, const void* q_norm_w,
               const void* k_norm_w, const int32_t* positions, float theta, float eps, float q_scale, int mode, void* stream);
/* same; out != NULL writes the roped q|k heads to out[M, >= (n_q+n_k)*head_dim] instead of in place (the pre-norm values stay in qkv for
 * the backward: no copy); rope_table (br_rope_table, [n_pos, head_dim/2] (cos, sin) pairs) replaces the per-element powf/sincosf (mode 0) */
int br_qk_rope_ex(void* qkv, int64_t ld, void* out, int64_t ldo, int M, int n_q_heads, int n_k_heads, int head_dim, const void* q_norm_w,
                  const void* k_norm_w, const int32_t* positions, float theta, float eps, float q_scale, int mode, const float* rope_table,
                  int rope_n_pos, void* stream);
/* out[m] = table[ids[m]] (zeros if keep && !keep[m], or id out of range); ids int64 */
int br_embed_gather(const int64_t* ids, const void* table, int64_t ldt, int64_t vocab, void* out, int64_t ldo, int M, int d,
                    const int32_t* keep, void* stream);
int br_scatter_rows(const void* src, int64_t lds, const int32_t* row_map, void* dst, int64_t ldd, int M, int d, void* stream);
/* out[m] = src[idx[m]] (idx < 0 -> zero row) */
int br_gather_rows(const void* src, int64_t lds, const int32_t* idx, void* dst, int64_t ldd, int M, int d, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Attention (replaces the SDPA call under Qwen3Attention / EsmSelfAttention)
 * q/k/v/o bf16 token-major [B*L, ld] with head h at column h*head_dim; row b attends keys in
 * [kv_start[b], kv_end[b]) (NULL = whole row) and, if causal, j <= i.  lse [B, Hq, L] f32 optional.
 * ------------------------------------------------------------------------------------------- */
int br_attn_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o, int64_t ldo,
                float* lse, int B, int L, int n_q_heads, int n_kv_heads, int head_dim, const int32_t* kv_start,
                const int32_t* kv_end, float scale, int causal, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Rollout / decode (replaces the HF generate() token loop, DynamicCache and logits warpers reached from
 * dna_llm.py:298-304; HF generation/utils.py:2760-2800, 1214-1223)
 * KV cache per layer: K and V are [n_pages, Hkv, 64, head_dim] bf16; page_table int32 [R, max_pages];
 * cur_len int32 [R] = tokens already cached for the row (the position of the token being decoded).
 * ------------------------------------------------------------------------------------------- */
int64_t br_skinny_scratch_bytes(int max_N);
/* out[R, N] = X[R, K] . W[N, K]^T for R <= 32 (HBM-bound weight streaming). mode 0: bf16; 1: bf16(out) + residual;
 * 2: SwiGLU over (8 gate | 8 up) row blocks -> [R, N/2]; 3: fp32.  scratch: zero-initialised once (arrival counters self-reset). */
int br_skinny_gemm(const void* X, int64_t ldx, const void* W, int64_t ldw, void* out, int64_t ldo, int R, int N, int K, int mode,
                   const void* residual, int64_t ldr, void* scratch, void* stream);
/* Same with a folded RMSNorm for the decode step (norm weight pre-multiplied into W's columns by br_scale_columns):
 * sumsq_in [sumsq_in_n, 32]: partial sums of x^2 per row; out rows are scaled by rsqrt(sum_i sumsq_in[i, r] / K + eps);
 * sumsq_out [ceil(N/128)*4, 32]: partial sums of the bf16-rounded outputs squared (modes 0/1), one partial row per 32
 * features (so the consumer passes sumsq_in_n = ceil(N/128)*4).  No floating-point atomics: the rollout is reproducible. */
int br_skinny_gemm_ex(const void* X, int64_t ldx, const void* W, int64_t ldw, void* out, int64_t ldo, int R, int N, int K, int mode,
                      const void* residual, int64_t ldr, void* scratch, const float* sumsq_in, int sumsq_in_n, float* sumsq_out,
                      float eps, void* stream);
/* L2 staging for the decode loop.  The decode step is a chain of small dependent kernels; while one of them waits on its predecessor or
 * reduces partial tiles, HBM idles.  A launch can therefore pull weight tiles of a LATER GEMM of the chain into the 126 MB L2:
 * `W [N, K]` is that GEMM's weight, and of every chunk its CTAs will stream (the stream-K decomposition of br_skinny_gemm_ex) the 16 KB
 * tiles [unit_lo, unit_hi) are prefetched.  Weights are constant during a rollout, so this is safe at any point of the chain. */
typedef struct br_l2_prefetch { const void* W; int64_t ldw; int32_t N, K; int32_t unit_lo, unit_hi; } br_l2_prefetch;
int br_skinny_gemm_pf(const void* X, int64_t ldx, const void* W, int64_t ldw, void* out, int64_t ldo, int R, int N, int K, int mode,
                      const void* residual, int64_t ldr, void* scratch, const float* sumsq_in, int sumsq_in_n, float* sumsq_out,
                      float eps, const br_l2_prefetch* prefetch, void* stream);
/* Stream gate for the decode chain.  A decode GEMM that becomes resident early fills its weight ring BEFORE its dependency resolves; if
 * the previous GEMM is still streaming, those loads only take bandwidth away from it (its exchange tail starts later by the same amount).
 * With a gate the early loads start when the previous GEMM's weights have all ARRIVED, i.e. they run under its exchange tail, when HBM
 * would idle.  `counter` (int32, zero at the start of a rollout) counts "my weights are on chip" arrivals, one per CTA of every gated
 * launch; a launch starts prefetching once counter >= (*epoch - epoch_base) * per_step + wait_prefix (wait_prefix < 0: at once).  The
 * gate is a timing hint only: the wait is bounded, and a wrong specification costs time, never correctness. */
typedef struct br_stream_gate { int32_t* counter; const int32_t* epoch; int32_t epoch_base, per_step, wait_prefix, signal; } br_stream_gate;
int br_skinny_gemm_gated(const void* X, int64_t ldx, const void* W, int64_t ldw, void* out, int64_t ldo, int R, int N, int K, int mode,
                         const void* residual, int64_t ldr, void* scratch, const float* sumsq_in, int sumsq_in_n, float* sumsq_out,
                         float eps, const br_l2_prefetch* prefetch, const br_stream_gate* gate, void* stream);
/* CTAs br_skinny_gemm launches for a weight [N, K] */
int br_skinny_grid(int N, int K);
/* Up to 4 dependent decode GEMMs in ONE persistent launch (e.g. o_proj -> gate/up -> down_proj -> next layer's qkv):
 * phases are separated by a grid-wide barrier inside the kernel and the weight producer prefetches across it, so the
 * HBM stream does not stall at layer boundaries.  Same per-phase semantics as br_skinny_gemm_ex. */
typedef struct br_skinny_phase {
    const void* X; int64_t ldx;         /* [R, K] bf16 input (for phases > 0: the output of an earlier phase) */
    const void* W; int64_t ldw;         /* [N, K] bf16 weight */
    void* out; int64_t ldo;
    int32_t N, K, mode;
    const void* residual; int64_t ldr;
    const float* sumsq_in; int32_t sumsq_in_n;
    float* sumsq_out;
} br_skinny_phase;
int br_skinny_chain(const br_skinny_phase* phases, int n_phases, int R, float eps, void* scratch, void* stream);
/* profiling aid: [n_sms, 32] int64 %globaltimer stamps (per CTA: start, dep-wait, then per phase: begin, first accumulator,
 * tiles done, barrier arrive, barrier pass) written by the next br_skinny_chain launches; NULL disables */
int br_skinny_chain_debug(long long* buf);
/* profiling aid for br_skinny_gemm(_ex): [n_launches, 160, 8] int64 %globaltimer stamps per CTA of the following launches
 * (kernel entry, dependency wait passed, prologue done, first accumulator, partial published, reduction loads, reducer epilogue, done) */
int br_skinny_debug(long long* buf);
int br_embed_gather_sumsq(const int64_t* ids, const void* table, int64_t ldt, int64_t vocab, void* out, int64_t ldo, int M, int d,
                          float* sumsq, void* stream);
/* W[n, k] *= scale[k] in place (bf16) */
int br_scale_columns(void* W, int64_t ld, int64_t N, int K, const void* scale, void* stream);
/* per-head q/k RMSNorm + RoPE at position cur_len[r]; K and V of the new token go into the row's page, Q stays in qkv */
int br_decode_rope_append(void* qkv, int64_t ld, int R, int n_q_heads, int n_kv_heads, int head_dim, const void* q_norm_w,
                          const void* k_norm_w, const int32_t* cur_len, const int32_t* page_table, int max_pages,
                          void* kcache, void* vcache, float theta, float eps, void* stream);
/* prefill: copy roped K / V of tokens [0, n_tok) of one prompt row (qkv points at its first real token) into pages[] */
int br_kv_write_pages(const void* qkv, int64_t ld, int n_tok, int n_q_heads, int n_kv_heads, int head_dim, const int32_t* pages,
                      void* kcache, void* vcache, void* stream);
int64_t br_decode_attn_workspace_bytes(int R, int n_q_heads, int head_dim, int n_slots);
/* one decode-attention step; rows are R/G groups whose first n_shared_pages table entries are identical (prefix sharing:
 * the shared pass reads each prompt K/V tile once per group). n_slots = splits_shared (if used) + splits_private. */
int br_decode_attn(const void* qkv, int64_t ld, const void* kcache, const void* vcache, const int32_t* page_table, int max_pages,
                   const int32_t* cur_len, int R, int G, int n_q_heads, int n_kv_heads, int head_dim, int n_shared_pages,
                   int splits_shared, int splits_private, float scale, void* workspace, void* out, int64_t ldo, void* stream);
/* temperature -> top-k -> top-p -> inverse-CDF draw with uniforms[step*R + r] (or argmax when !do_sample); finished rows
 * emit pad_id; writes tokens[r, step] (int64 [R, max_steps]) and next_ids[r]; eos_id < 0 disables EOS. */
int br_sample_next(const float* logits, int64_t ld, int R, int V, float temperature, int top_k, float top_p, int do_sample,
                   const float* uniforms, const int32_t* step, int max_steps, int64_t eos_id, int64_t pad_id, int32_t* finished,
                   int64_t* tokens, int64_t* next_ids, void* stream);
/* Same semantics in two stages for large vocabularies: stage 1 (V/4096 CTAs per row) reduces each row to <= 64 candidates
 * per 4096-logit chunk, stage 2 samples from the candidates (top_k <= 32). */
int64_t br_sample_workspace_bytes(int R, int V);
int br_sample_next_2stage(const float* logits, int64_t ld, int R, int V, float temperature, int top_k, float top_p, int do_sample,
                          const float* uniforms, const int32_t* step, int max_steps, int64_t eos_id, int64_t pad_id,
                          int32_t* finished, int64_t* tokens, int64_t* next_ids, void* workspace, void* stream);
int br_decode_advance(int32_t* step, int32_t* cur_len, int R, void* stream);

/* Fused decode attention (one launch per layer per step): per-head q/k RMSNorm + RoPE at cur_len[r], K/V append to the
 * row's page, prefix-shared + private paged attention, split merge.  qkv_raw is the un-normalised fused projection of
 * the new tokens.  workspace: br_decode_fused_workspace_bytes, zero-initialised once (arrival counters are self-resetting). */
int64_t br_decode_fused_workspace_bytes(int R, int n_q_heads, int n_kv_heads, int head_dim, int n_slots);
int br_decode_attn_fused(const void* qkv_raw, int64_t ld, const void* q_norm_w, const void* k_norm_w, void* kcache, void* vcache,
                         const int32_t* page_table, int max_pages, const int32_t* cur_len, int R, int G, int n_q_heads,
                         int n_kv_heads, int head_dim, int n_shared_pages, int splits_shared, int splits_private, float scale,
                         float theta, float eps, const float* rope_table, int rope_n_pos, void* workspace, void* out, int64_t ldo,
                         void* stream);
/* same, plus L2 staging of a later GEMM's weight tiles from this (HBM-light) launch; see br_l2_prefetch */
int br_decode_attn_fused_pf(const void* qkv_raw, int64_t ld, const void* q_norm_w, const void* k_norm_w, void* kcache, void* vcache,
                            const int32_t* page_table, int max_pages, const int32_t* cur_len, int R, int G, int n_q_heads, int n_kv_heads,
                            int head_dim, int n_shared_pages, int splits_shared, int splits_private, float scale, float theta, float eps,
                            const float* rope_table, int rope_n_pos, void* workspace, void* out, int64_t ldo, const br_l2_prefetch* prefetch,
                            void* stream);
/* rope_table [n_pos, head_dim/2, 2] f32 = (cos, sin) rounded to bf16 precision (HF builds its tables in the model dtype);
 * optional input of br_decode_attn_fused: removes powf/sincosf from the decode loop. */
/* profiling aid: per-item phase timestamps ([items, 16] int64, %globaltimer ns) for the next fused-attention launches */
int br_decode_attn_fused_debug(long long* buf);
int br_rope_table(float* out, int n_pos, int head_dim, float theta, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Backward (autograd counterparts; frozen base weights + LoRA adapters, reason.py:362-394)
 * ------------------------------------------------------------------------------------------- */
int64_t br_attn_bwd_workspace_bytes(int B, int L, int n_q_heads, int head_dim);
/* dq/dk/dv (bf16, strided -- typically the three column blocks of one fused dqkv buffer) from dout; causal, head_dim 128 */
int br_attn_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* o, int64_t ldo,
                const void* dout, int64_t lddo, const float* lse, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv,
                int64_t lddv, int B, int L, int n_q_heads, int n_kv_heads, int head_dim, const int32_t* kv_start,
                const int32_t* kv_end, float scale, void* workspace, void* stream);
/* dx = d(RMSNorm)/dx . dy (+ dres): x, dy, dres, dx bf16 [M, d]; rstd from the forward */
int br_rmsnorm_bwd(const void* x, int64_t ldx, const void* w, const float* rstd, const void* dy, int64_t lddy, const void* dres,
                   int64_t lddr, void* dx, int64_t lddx, int M, int d, void* stream);
/* gu, dgu [M, 2F] in the blocked (8 gate | 8 up) layout; dact [M, F] */
int br_swiglu_bwd(const void* gu, int64_t ldgu, const void* dact, int64_t ldda, void* dgu, int64_t lddgu, int M, int F, void* stream);
/* in place on the q and k head columns of dqkv: inverse RoPE then per-head RMSNorm backward (qk_pre = pre-norm q|k) */
int br_qk_rope_bwd(void* dqkv, int64_t ldd, const void* qk_pre, int64_t ldp, int M, int n_q_heads, int n_k_heads, int head_dim,
                   const void* q_norm_w, const void* k_norm_w, const int32_t* positions, float theta, float eps, void* stream);
/* LoRA weight gradients on tcgen05 (deterministic):  product[P, N] = big[M, P]^T . small[M, N] over the M tokens (both token-major,
 * bf16, read as MN-major tensor-core operands), then  dst (+)= the blocks the segments name.
 *   mode 0: segment i adds product rows [row_lo, row_hi), columns [col_lo, col_lo + n_cols) into dst[(row - row_lo) * ld + col - col_lo]
 *           (dB of one adapter, or the q / k / v blocks of the fused qkv product);
 *   mode 1: one segment, transposed: dst[n * ld + p] += product[p, n]  (dA = u^T x written as [r, in]);
 *   mode 2: gate/up-blocked rows (16 = 8 gate | 8 up): segment 0 takes the gate rows, segment 1 the up rows -> dst row (p / 16) * 8 + p % 8.
 * workspace: br_lora_grad_workspace_bytes(), zero-initialised once.  Replaces torch autograd through peft's LoRA Linear (reason.py:362-394). */
typedef struct br_lora_grad_seg { float* dst; int64_t ld; int32_t row_lo, row_hi, col_lo, n_cols; } br_lora_grad_seg;
int64_t br_lora_grad_workspace_bytes(void);
int br_lora_grad_tn(const void* big, int64_t ldb, const void* small, int64_t lds, int M, int P, int N, int mode,
                    const br_lora_grad_seg* segs, int n_seg, void* workspace, void* stream);
int br_transpose_bf16(const void* in, int64_t ldi, void* out, int64_t ldo, int M, int N, void* stream);
int br_colsum_accumulate(const void* in, int64_t ldi, float* out, int M, int N, void* stream);

#ifdef __cplusplus
}
#endif
#endif
