// Paged-KV decode for the GRPO rollout (replaces the HF generate() token loop + DynamicCache torch.cat growth,
// HF generation/utils.py:2760-2800).
//
// KV cache layout (per layer): K and V are [n_pages, Hkv, 64, D] bf16 -- the 64 keys of one (page, kv head) are one
// contiguous 16 KB tile, i.e. exactly the shared-memory tile of the attention kernel.  A row's context is a page
// table; the G samples of a prompt group point at the SAME prompt pages (prefix sharing), so
//   * the shared-prefix pass treats the whole group as one problem: G x (Hq/Hkv) query vectors per kv head form the
//     M dimension of the QK^T / PV mma tiles, and every prompt K/V tile is read once per group instead of G times;
//   * the private pass covers each row's own pages (tail of the prompt + generated tokens);
//   * a combine pass merges the per-split partial (O, LSE) pairs.
// Everything that changes from step to step (row lengths) is read from device memory, so one captured CUDA graph
// replays for every token.
#include "br_common.cuh"
#include "../../include/bioreason_b200.h"
#include "attn_common.cuh"
using namespace attn;

namespace {

__device__ __forceinline__ float rbf(float x) { return __bfloat162float(__float2bfloat16(x)); }

// ------------------------------------------------------------------------------------------------
// q/k norm + RoPE at position cur_len[row] for the new token of every row; K (roped) and V go straight into the
// row's page; Q stays in the qkv buffer.  One warp per (row, head) over q, k and v heads.
// ------------------------------------------------------------------------------------------------
template <int D>
__global__ void rope_append_kernel(bf16* __restrict__ qkv, long long ld, int R, int Hq, int Hkv, const bf16* __restrict__ qw,
                                   const bf16* __restrict__ kw, const int* __restrict__ cur_len, const int* __restrict__ page_table,
                                   int max_pages, bf16* __restrict__ kcache, bf16* __restrict__ vcache, float theta, float eps) {
    constexpr int E = D / 64;
    const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int heads = Hq + 2 * Hkv;
    if (wid >= R * heads) return;
    const int r = wid / heads, h = wid % heads;
    bf16* p = qkv + (long long)r * ld + (long long)h * D;
    const int pos = cur_len[r];
    const int page = page_table[(long long)r * max_pages + (pos >> 6)];
    const int slot = pos & 63;
    if (h >= Hq + Hkv) {                                               // V: plain copy into the page
        bf16* dst = vcache + (((long long)page * Hkv + (h - Hq - Hkv)) * 64 + slot) * D;
        for (int i = lane; i < D / 8; i += 32) reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(p)[i];
        return;
    }
    float lo[E], hi[E];
#pragma unroll
    for (int e = 0; e < E; ++e) { lo[e] = __bfloat162float(p[lane * E + e]); hi[e] = __bfloat162float(p[D / 2 + lane * E + e]); }
    const bf16* w = (h < Hq) ? qw : kw;
    float ss = 0.f;
#pragma unroll
    for (int e = 0; e < E; ++e) ss += lo[e] * lo[e] + hi[e] * hi[e];
    const float rstd = rsqrtf(br::warp_sum(ss) / (float)D + eps);
    bf16* dst = (h < Hq) ? p : kcache + (((long long)page * Hkv + (h - Hq)) * 64 + slot) * D;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int j = lane * E + e;
        const float a = rbf(__bfloat162float(w[j]) * rbf(lo[e] * rstd));
        const float b = rbf(__bfloat162float(w[D / 2 + j]) * rbf(hi[e] * rstd));
        const float inv_freq = 1.0f / powf(theta, (float)(2 * j) / (float)D);
        float sn, cs;
        sincosf((float)pos * inv_freq, &sn, &cs);
        sn = rbf(sn); cs = rbf(cs);
        dst[j] = __float2bfloat16(rbf(a * cs) + rbf(-b * sn));
        dst[D / 2 + j] = __float2bfloat16(rbf(b * cs) + rbf(a * sn));
    }
}

// prefill: copy the (already roped) K and V of tokens [0, n_tok) of one prompt row into its pages
template <int D>
__global__ void kv_write_pages_kernel(const bf16* __restrict__ qkv, long long ld, int n_tok, int Hq, int Hkv, const int* __restrict__ pages,
                                      bf16* __restrict__ kcache, bf16* __restrict__ vcache) {
    const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (wid >= n_tok * 2 * Hkv) return;
    const int tok = wid / (2 * Hkv), hh = wid % (2 * Hkv);
    const bool is_v = hh >= Hkv;
    const int kvh = is_v ? hh - Hkv : hh;
    const bf16* src = qkv + (long long)tok * ld + (long long)(Hq + hh) * D;
    const int page = pages[tok >> 6], slot = tok & 63;
    bf16* dst = (is_v ? vcache : kcache) + (((long long)page * Hkv + kvh) * 64 + slot) * D;
    for (int i = lane; i < D / 8; i += 32) reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
}

// ------------------------------------------------------------------------------------------------
// decode attention partials
// ------------------------------------------------------------------------------------------------
struct DecParams {
    const bf16* qkv; long long ld;       // [R, ld] (q heads first)
    const bf16 *kcache, *vcache;
    const int* page_table; int max_pages;
    const int* cur_len;                  // [R] tokens in cache BEFORE this step's append (new token sits at index cur_len)
    int R, Hq, Hkv, GQ;
    int rows_per_unit;                   // G for the shared pass, 1 for the private pass
    int n_shared_pages;                  // shared pass: pages [0, n_shared); private pass: pages [n_shared, ...)
    int shared_pass;
    int n_splits, slot_base, n_slots;    // partial slot = slot_base + split
    float* part_o;                       // [R, Hq, n_slots, D] fp32
    float* part_lse;                     // [R, Hq, n_slots]
    float scale_log2;
};

template <int D, int NW>
__global__ void __launch_bounds__(32 * NW) decode_attn_kernel(const DecParams p) {
    constexpr int BN = 64, TILE = 64 * D * 2, NT = 32 * NW, QROWS = 16 * NW;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sQ = smem;                          // QROWS x D
    uint8_t* sK = smem + QROWS * D * 2;          // 2 stages
    uint8_t* sV = sK + 2 * TILE;

    const int split = blockIdx.x, kvh = blockIdx.y, unit = blockIdx.z;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int row_base = unit * p.rows_per_unit;

    // page range of this CTA
    int kv_len = 0;                               // keys valid (private pass): max over the unit's rows (= its single row)
    int pg_lo, pg_hi;                             // [lo, hi) page indices in the table, stepping by n_splits
    if (p.shared_pass) { pg_lo = split; pg_hi = p.n_shared_pages; }
    else {
        kv_len = p.cur_len[row_base] + 1;
        pg_lo = p.n_shared_pages + split;
        pg_hi = (kv_len + 63) >> 6;
    }
    const int* table = p.page_table + (long long)row_base * p.max_pages;

    // Q tile: slot s -> (row_base + s / GQ, head kvh*GQ + s % GQ); rows beyond the unit are zero
    {
        constexpr int CH = D / 8;
        for (int c = tid; c < QROWS * CH; c += NT) {
            const int s = c / CH, ch = c % CH;
            const int rr = s / p.GQ, hh = kvh * p.GQ + s % p.GQ;
            const bool ok = rr < p.rows_per_unit && (row_base + rr) < p.R;
            const bf16* src = p.qkv + (long long)(ok ? row_base + rr : 0) * p.ld + (long long)hh * D + ch * 8;
            cp_async16(tile_ptr<D>(sQ, s, ch), src, ok);
        }
    }
    const long long page_stride = (long long)p.Hkv * 64 * D;
    auto tile_src = [&](const bf16* cache, int pg) { return cache + (long long)table[pg] * page_stride + (long long)kvh * 64 * D; };
    if (pg_lo < pg_hi) {
        load_tile<D, NT>(sK, tile_src(p.kcache, pg_lo), D, 0, 64, tid);
        load_tile<D, NT>(sV, tile_src(p.vcache, pg_lo), D, 0, 64, tid);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    uint32_t qf[D / 16][4];
#pragma unroll
    for (int kk = 0; kk < D / 16; ++kk)
        ldsm_x4(qf[kk], tile_ptr<D>(sQ, warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, kk * 2 + (lane >> 4)));

    float o[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

    int it = 0;
    for (int pg = pg_lo; pg < pg_hi; pg += p.n_splits, ++it) {
        const int st = it & 1;
        uint8_t* cK = sK + st * TILE;
        uint8_t* cV = sV + st * TILE;
        if (pg + p.n_splits < pg_hi) {
            load_tile<D, NT>(sK + (st ^ 1) * TILE, tile_src(p.kcache, pg + p.n_splits), D, 0, 64, tid);
            load_tile<D, NT>(sV + (st ^ 1) * TILE, tile_src(p.vcache, pg + p.n_splits), D, 0, 64, tid);
        }
        cp_async_commit();

        float s[BN / 8][4];
#pragma unroll
        for (int i = 0; i < BN / 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
#pragma unroll
            for (int np = 0; np < BN / 16; ++np) {
                uint32_t kf[4];
                ldsm_x4(kf, tile_ptr<D>(cK, np * 16 + (lane & 7) + (lane >> 4) * 8, kk * 2 + ((lane >> 3) & 1)));
                mma16816(s[2 * np], qf[kk], kf[0], kf[1]);
                mma16816(s[2 * np + 1], qf[kk], kf[2], kf[3]);
            }
        }
        const int nbase = pg * BN;
        const bool need_mask = !p.shared_pass && (nbase + BN > kv_len);
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < BN / 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float v = s[nt][e] * p.scale_log2;
                if (need_mask) {
                    const int j = nbase + nt * 8 + 2 * t + (e & 1);
                    v = (j < kv_len) ? v : -INFINITY;
                }
                s[nt][e] = v;
            }
            mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
            mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
        const float ms0 = (mn0 == -INFINITY) ? 0.f : mn0, ms1 = (mn1 == -INFINITY) ? 0.f : mn1;
        const float a0 = exp2f(m0 - ms0), a1 = exp2f(m1 - ms1);
        m0 = mn0; m1 = mn1;
        float rs0 = 0.f, rs1 = 0.f;
        uint32_t pf[BN / 16][4];
#pragma unroll
        for (int nt = 0; nt < BN / 8; ++nt) {
            const float p0 = exp2f(s[nt][0] - ms0), p1 = exp2f(s[nt][1] - ms0);
            const float p2 = exp2f(s[nt][2] - ms1), p3 = exp2f(s[nt][3] - ms1);
            rs0 += p0 + p1; rs1 += p2 + p3;
            pf[nt >> 1][(nt & 1) * 2 + 0] = br::pack_bf16(p0, p1);
            pf[nt >> 1][(nt & 1) * 2 + 1] = br::pack_bf16(p2, p3);
        }
        l0 = l0 * a0 + rs0; l1 = l1 * a1 + rs1;
#pragma unroll
        for (int i = 0; i < D / 8; ++i) { o[i][0] *= a0; o[i][1] *= a0; o[i][2] *= a1; o[i][3] *= a1; }
#pragma unroll
        for (int kk = 0; kk < BN / 16; ++kk) {
#pragma unroll
            for (int dp = 0; dp < D / 16; ++dp) {
                uint32_t vf[4];
                ldsm_x4_t(vf, tile_ptr<D>(cV, kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, dp * 2 + (lane >> 4)));
                mma16816(o[2 * dp], pf[kk], vf[0], vf[1]);
                mma16816(o[2 * dp + 1], pf[kk], vf[2], vf[3]);
            }
        }
        cp_async_wait<0>();
        __syncthreads();
    }

    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float LN2 = 0.6931471805599453f;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int s_idx = warp * 16 + g + half * 8;
        const int rr = s_idx / p.GQ, hh = kvh * p.GQ + s_idx % p.GQ;
        if (rr >= p.rows_per_unit || row_base + rr >= p.R) continue;
        const float l = half ? l1 : l0, m = half ? m1 : m0;
        const float inv = l > 0.f ? 1.f / l : 0.f;
        const long long base = ((long long)(row_base + rr) * p.Hq + hh) * p.n_slots + p.slot_base + split;
        float* po = p.part_o + base * D;
#pragma unroll
        for (int dt = 0; dt < D / 8; ++dt) {
            const float x = half ? o[dt][2] : o[dt][0], y = half ? o[dt][3] : o[dt][1];
            *reinterpret_cast<float2*>(po + dt * 8 + 2 * t) = make_float2(x * inv, y * inv);
        }
        if (t == 0) p.part_lse[base] = l > 0.f ? m * LN2 + logf(l) : -INFINITY;
    }
}

// out[r, h, :] = sum_s w_s * part_o[r, h, s, :],  w_s = exp(lse_s - max) / sum   -> bf16 [R, Hq*D]
template <int D>
__global__ void decode_combine_kernel(const float* __restrict__ part_o, const float* __restrict__ part_lse, int n_slots, bf16* __restrict__ out,
                                      long long ldo, int R, int Hq) {
    const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (wid >= R * Hq) return;
    const float* lse = part_lse + (long long)wid * n_slots;
    float mx = -INFINITY;
    for (int s = 0; s < n_slots; ++s) mx = fmaxf(mx, lse[s]);
    float acc[D / 32];
#pragma unroll
    for (int i = 0; i < D / 32; ++i) acc[i] = 0.f;
    float den = 0.f;
    for (int s = 0; s < n_slots; ++s) {
        const float l = lse[s];
        if (l == -INFINITY) continue;
        const float w = __expf(l - mx);
        den += w;
        const float* po = part_o + ((long long)wid * n_slots + s) * D;
#pragma unroll
        for (int i = 0; i < D / 32; ++i) acc[i] += w * po[lane + 32 * i];
    }
    const float inv = den > 0.f ? 1.f / den : 0.f;
    const int r = wid / Hq, h = wid % Hq;
#pragma unroll
    for (int i = 0; i < D / 32; ++i) out[(long long)r * ldo + (long long)h * D + lane + 32 * i] = __float2bfloat16(acc[i] * inv);
}

template <int NW>
int launch_dec(const DecParams& p, int n_units, cudaStream_t st) {
    constexpr int D = 128;
    constexpr int SMEM = 16 * NW * D * 2 + 4 * 64 * D * 2;
    auto kern = decode_attn_kernel<D, NW>;
    static bool done = false;
    if (!done) { BR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM)); done = true; }
    dim3 grid(p.n_splits, p.Hkv, n_units);
    kern<<<grid, 32 * NW, SMEM, st>>>(p);
    BR_CHECK_LAUNCH();
    return BR_OK;
}

}  // namespace

extern "C" {

int br_decode_rope_append(void* qkv, int64_t ld, int R, int n_q_heads, int n_kv_heads, int head_dim, const void* q_norm_w, const void* k_norm_w,
                          const int32_t* cur_len, const int32_t* page_table, int max_pages, void* kcache, void* vcache, float theta, float eps,
                          void* stream) {
    BR_CHECK_ARG(head_dim == 128, "decode path is built for head_dim 128 (Qwen3)");
    BR_CHECK_ARG(q_norm_w && k_norm_w, "decode_rope_append: q/k norm weights required (Qwen3)");
    const int warps = R * (n_q_heads + 2 * n_kv_heads), wpb = 8;
    rope_append_kernel<128><<<(warps + wpb - 1) / wpb, wpb * 32, 0, (cudaStream_t)stream>>>(
        (bf16*)qkv, ld, R, n_q_heads, n_kv_heads, (const bf16*)q_norm_w, (const bf16*)k_norm_w, cur_len, page_table, max_pages, (bf16*)kcache,
        (bf16*)vcache, theta, eps);
    BR_CHECK_LAUNCH();
    return BR_OK;
}

int br_kv_write_pages(const void* qkv, int64_t ld, int n_tok, int n_q_heads, int n_kv_heads, int head_dim, const int32_t* pages, void* kcache,
                      void* vcache, void* stream) {
    BR_CHECK_ARG(head_dim == 128 && n_tok > 0, "kv_write_pages: head_dim 128, n_tok > 0");
    const int warps = n_tok * 2 * n_kv_heads, wpb = 8;
    kv_write_pages_kernel<128><<<(warps + wpb - 1) / wpb, wpb * 32, 0, (cudaStream_t)stream>>>((const bf16*)qkv, ld, n_tok, n_q_heads, n_kv_heads,
                                                                                             pages, (bf16*)kcache, (bf16*)vcache);
    BR_CHECK_LAUNCH();
    return BR_OK;
}

int64_t br_decode_attn_workspace_bytes(int R, int n_q_heads, int head_dim, int n_slots) {
    return (int64_t)R * n_q_heads * n_slots * (head_dim + 1) * sizeof(float);
}

/* One decode-attention step for R rows organised as R/G groups whose first n_shared_pages pages are common. */
int br_decode_attn(const void* qkv, int64_t ld, const void* kcache, const void* vcache, const int32_t* page_table, int max_pages,
                   const int32_t* cur_len, int R, int G, int n_q_heads, int n_kv_heads, int head_dim, int n_shared_pages, int splits_shared,
                   int splits_private, float scale, void* workspace, void* out, int64_t ldo, void* stream) {
    BR_CHECK_ARG(head_dim == 128, "decode_attn: head_dim 128 only");
    BR_CHECK_ARG(R > 0 && G > 0 && R % G == 0, "decode_attn: R=%d must be a multiple of G=%d", R, G);
    const int GQ = n_q_heads / n_kv_heads;
    BR_CHECK_ARG(GQ <= 16 && 16 % GQ == 0, "decode_attn: Hq/Hkv must divide 16");
    BR_CHECK_ARG(splits_private >= 1 && splits_shared >= 0, "decode_attn: bad split counts");
    cudaStream_t st = (cudaStream_t)stream;
    const int use_shared = (n_shared_pages > 0 && splits_shared > 0) ? 1 : 0;
    const int n_slots = (use_shared ? splits_shared : 0) + splits_private;
    DecParams p;
    p.qkv = (const bf16*)qkv; p.ld = ld; p.kcache = (const bf16*)kcache; p.vcache = (const bf16*)vcache;
    p.page_table = page_table; p.max_pages = max_pages; p.cur_len = cur_len; p.R = R; p.Hq = n_q_heads; p.Hkv = n_kv_heads; p.GQ = GQ;
    p.n_slots = n_slots; p.part_o = (float*)workspace; p.part_lse = p.part_o + (int64_t)R * n_q_heads * n_slots * head_dim;
    p.scale_log2 = scale * 1.4426950408889634f;
    int rc;
    if (use_shared) {
        const int qv = G * GQ;
        BR_CHECK_ARG(qv <= 64, "decode_attn: G * Hq/Hkv = %d query vectors per kv head exceed 64", qv);
        p.rows_per_unit = G; p.n_shared_pages = n_shared_pages; p.shared_pass = 1; p.n_splits = splits_shared; p.slot_base = 0;
        if (qv <= 16) rc = launch_dec<1>(p, R / G, st);
        else if (qv <= 32) rc = launch_dec<2>(p, R / G, st);
        else rc = launch_dec<4>(p, R / G, st);
        if (rc) return rc;
    }
    p.rows_per_unit = 1; p.n_shared_pages = use_shared ? n_shared_pages : 0; p.shared_pass = 0; p.n_splits = splits_private;
    p.slot_base = use_shared ? splits_shared : 0;
    if ((rc = launch_dec<1>(p, R, st))) return rc;
    const int warps = R * n_q_heads, wpb = 4;
    decode_combine_kernel<128><<<(warps + wpb - 1) / wpb, wpb * 32, 0, st>>>(p.part_o, p.part_lse, n_slots, (bf16*)out, ldo, R, n_q_heads);
    BR_CHECK_LAUNCH();
    return BR_OK;
}

}  // extern "C"
