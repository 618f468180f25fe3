// Flash attention forward (causal GQA decoder rows and bidirectional encoder rows), online softmax, no score matrix in HBM.
// Replaces the SDPA call HF reaches from Qwen3Attention.forward (qwen3/modeling_qwen3.py:255-263) and EsmSelfAttention.forward
// (esm/modeling_esm.py:349-359); SURVEY.md §2.3 K1/K5.
//
// Rows are dense [B, L] token-major; each row b attends keys j in [kv_start[b], kv_end[b]) (left pads / post-EOS tail are
// outside the window -- the reference's 0/1 attention_mask is always one contiguous run), plus j <= i when causal.
// CTA = 4 warps x 16 query rows, 64-key tiles double-buffered with cp.async into XOR-swizzled shared memory, bf16
// mma.sync.m16n8k16 with fp32 accumulation (tensor-core legacy path; the tcgen05 version of this kernel is the next step).
#include "br_common.cuh"
#include "../../include/bioreason_b200.h"

namespace {

struct AttnParams {
    const bf16 *q, *k, *v;
    bf16* o;
    float* lse;            // [B, Hq, L] or null
    long long ldq, ldk, ldv, ldo;
    int B, L, Hq, Hkv;
    const int *kv_start, *kv_end;
    float scale_log2;      // softmax scale * log2(e)
};

}  // namespace
#include "attn_common.cuh"
using namespace attn;
namespace {

template <int D, bool CAUSAL>
__global__ void __launch_bounds__(128) attn_fwd_kernel(const AttnParams p) {
    constexpr int BM = 64, BN = 64, CH = D / 8, TILE = 64 * D * 2;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sQ = smem;
    uint8_t* sK = smem + TILE;          // 2 stages
    uint8_t* sV = smem + 3 * TILE;      // 2 stages

    const int qb = gridDim.x - 1 - blockIdx.x;     // heavy (late) causal blocks first
    const int h = blockIdx.y, b = blockIdx.z;
    const int hk = h / (p.Hq / p.Hkv);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int q0 = qb * BM;
    const int ks = p.kv_start ? p.kv_start[b] : 0;
    const int ke = p.kv_end ? p.kv_end[b] : p.L;

    const long long tok0 = (long long)b * p.L;
    const bf16* Q = p.q + tok0 * p.ldq + (long long)h * D;
    const bf16* K = p.k + tok0 * p.ldk + (long long)hk * D;
    const bf16* V = p.v + tok0 * p.ldv + (long long)hk * D;

    int jb_lo = ks / BN;
    int last_key = ke - 1;
    if (CAUSAL) last_key = min(last_key, q0 + BM - 1);
    int jb_hi = last_key >= 0 ? last_key / BN : -1;     // inclusive
    if (ke <= ks) jb_hi = jb_lo - 1;

    load_tile<D>(sQ, Q, p.ldq, q0, p.L, tid);
    if (jb_lo <= jb_hi) {
        load_tile<D>(sK, K, p.ldk, jb_lo * BN, p.L, tid);
        load_tile<D>(sV, V, p.ldv, jb_lo * BN, p.L, tid);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    // Q fragments stay in registers for the whole kernel
    uint32_t qf[D / 16][4];
#pragma unroll
    for (int kk = 0; kk < D / 16; ++kk)
        ldsm_x4(qf[kk], tile_ptr<D>(sQ, warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, kk * 2 + (lane >> 4)));

    float o[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    const int row0 = q0 + warp * 16 + g, row1 = row0 + 8;

    for (int jb = jb_lo; jb <= jb_hi; ++jb) {
        const int st = (jb - jb_lo) & 1;
        uint8_t* cK = sK + st * TILE;
        uint8_t* cV = sV + st * TILE;
        if (jb + 1 <= jb_hi) {                         // prefetch next K/V tile into the other stage
            load_tile<D>(sK + (st ^ 1) * TILE, K, p.ldk, (jb + 1) * BN, p.L, tid);
            load_tile<D>(sV + (st ^ 1) * TILE, V, p.ldv, (jb + 1) * BN, p.L, tid);
        }
        cp_async_commit();

        // S = Q K^T  (16 x 64 per warp)
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
        // mask + online softmax (log2 domain)
        const int nbase = jb * BN;
        const bool need_mask = (nbase < ks) || (nbase + BN > ke) || (CAUSAL && nbase + BN - 1 > q0 + warp * 16);
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < BN / 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float v = s[nt][e] * p.scale_log2;
                if (need_mask) {
                    const int j = nbase + nt * 8 + 2 * t + (e & 1);
                    const int i = (e < 2) ? row0 : row1;
                    const bool ok = (j >= ks) && (j < ke) && (!CAUSAL || j <= i);
                    v = ok ? v : -INFINITY;
                }
                s[nt][e] = v;
            }
            mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
            mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
        const float ms0 = (mn0 == -INFINITY) ? 0.f : mn0, ms1 = (mn1 == -INFINITY) ? 0.f : mn1;   // fully-masked rows stay finite
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
        // O += P V
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

    // finalize: quad-reduce the row sums, normalise, stage through (now free) sQ rows of this warp, 16-byte stores
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = l0 > 0.f ? 1.f / l0 : 0.f, inv1 = l1 > 0.f ? 1.f / l1 : 0.f;
    __syncwarp();
#pragma unroll
    for (int dt = 0; dt < D / 8; ++dt) {
        const int r_lo = warp * 16 + g, r_hi = r_lo + 8;
        *reinterpret_cast<uint32_t*>(tile_ptr<D>(sQ, r_lo, dt) + t * 4) = br::pack_bf16(o[dt][0] * inv0, o[dt][1] * inv0);
        *reinterpret_cast<uint32_t*>(tile_ptr<D>(sQ, r_hi, dt) + t * 4) = br::pack_bf16(o[dt][2] * inv1, o[dt][3] * inv1);
    }
    __syncwarp();
    bf16* O = p.o + tok0 * p.ldo + (long long)h * D;
#pragma unroll
    for (int i = 0; i < (16 * CH) / 32; ++i) {
        const int c = lane + i * 32;
        const int r = warp * 16 + c / CH, ch = c % CH;
        if (q0 + r < p.L)
            *reinterpret_cast<uint4*>(O + (long long)(q0 + r) * p.ldo + ch * 8) = *reinterpret_cast<const uint4*>(tile_ptr<D>(sQ, r, ch));
    }
    if (p.lse && t == 0) {
        const float LN2 = 0.6931471805599453f;
        float* lp = p.lse + ((long long)b * p.Hq + h) * p.L;
        if (row0 < p.L) lp[row0] = l0 > 0.f ? m0 * LN2 + logf(l0) : INFINITY;
        if (row1 < p.L) lp[row1] = l1 > 0.f ? m1 * LN2 + logf(l1) : INFINITY;
    }
}

template <int D, bool CAUSAL>
int launch_fwd(const AttnParams& p, cudaStream_t st) {
    constexpr int SMEM = 5 * 64 * D * 2;
    auto kern = attn_fwd_kernel<D, CAUSAL>;
    static bool done = false;
    if (!done) { BR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM)); done = true; }
    dim3 grid((p.L + 63) / 64, p.Hq, p.B);
    kern<<<grid, 128, SMEM, st>>>(p);
    BR_CHECK_LAUNCH();
    return BR_OK;
}

}  // namespace

int br_attn_fwd_tc5_impl(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o, int64_t ldo, float* lse,
                         int B, int L, int n_q_heads, int n_kv_heads, int head_dim, const int32_t* kv_start, const int32_t* kv_end,
                         float scale, int causal, cudaStream_t st);

extern "C" int br_attn_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o, int64_t ldo, float* lse,
                           int B, int L, int n_q_heads, int n_kv_heads, int head_dim, const int32_t* kv_start, const int32_t* kv_end,
                           float scale, int causal, void* stream) {
    BR_CHECK_ARG(B > 0 && L > 0 && n_q_heads % n_kv_heads == 0, "attn_fwd: bad shape B=%d L=%d Hq=%d Hkv=%d", B, L, n_q_heads, n_kv_heads);
    BR_CHECK_ARG(head_dim == 128 || head_dim == 64, "attn_fwd: head_dim must be 64 or 128");
    BR_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, "attn_fwd: strides must be multiples of 8 elements");
    static const bool legacy = getenv("BR_ATTN_LEGACY") != nullptr;       // debugging switch: the previous mma.sync kernel
    if (!legacy)
        return br_attn_fwd_tc5_impl(q, ldq, k, ldk, v, ldv, o, ldo, lse, B, L, n_q_heads, n_kv_heads, head_dim, kv_start, kv_end, scale, causal,
                                    (cudaStream_t)stream);
    AttnParams p;
    p.q = (const bf16*)q; p.k = (const bf16*)k; p.v = (const bf16*)v; p.o = (bf16*)o; p.lse = lse;
    p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.ldo = ldo; p.B = B; p.L = L; p.Hq = n_q_heads; p.Hkv = n_kv_heads;
    p.kv_start = kv_start; p.kv_end = kv_end; p.scale_log2 = scale * 1.4426950408889634f;
    cudaStream_t st = (cudaStream_t)stream;
    if (head_dim == 128) return causal ? launch_fwd<128, true>(p, st) : launch_fwd<128, false>(p, st);
    return causal ? launch_fwd<64, true>(p, st) : launch_fwd<64, false>(p, st);
}
