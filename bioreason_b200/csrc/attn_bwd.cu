// Flash attention backward (causal GQA, window-masked rows) -- recomputes the probabilities from Q, K and the saved
// log-sum-exp, never materialising scores in HBM.  Autograd counterpart of br_attn_fwd.
//
// CTA = one 64-key block of one kv head; its 4 warps own 16 keys each and keep dK / dV for them in registers while
// looping over the Hq/Hkv query heads of the group and over every 64-query block that can see the keys, so dK and dV
// are written once without atomics.  Per tile (all bf16 mma.sync.m16n8k16, fp32 accumulate):
//     S^T = K Q^T            P^T = exp(S^T - lse_q)
//     dP^T = V dO^T          dS^T = scale * P^T o (dP^T - delta_q)
//     dV += P^T dO           dK += dS^T Q
//     dQ  += dS K            (dS^T staged in shared memory, re-read transposed; fp32 atomics into a dQ accumulator)
#include "br_common.cuh"
#include "../../include/bioreason_b200.h"
#include "attn_common.cuh"
using namespace attn;

namespace {

struct BwdParams {
    const bf16 *q, *k, *v, *o, *dout;
    long long ldq, ldk, ldv, ldo, lddo;
    const float* lse;       // [B, Hq, L]
    float* delta;           // [B, Hq, L]
    float* dq_acc;          // [B*L, Hq*D] fp32, zeroed
    bf16 *dk, *dv;          // strided outputs (fused dqkv buffer)
    long long lddk, lddv;
    int B, L, Hq, Hkv;
    const int *kv_start, *kv_end;
    float scale, scale_log2;
};

// delta[b, h, i] = sum_d dO[i, h, d] * O[i, h, d]
template <int D>
__global__ void delta_kernel(const bf16* __restrict__ o, long long ldo, const bf16* __restrict__ dout, long long lddo, float* __restrict__ delta,
                             int B, int L, int Hq) {
    const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (wid >= (long long)B * L * Hq) return;
    const long long tok = wid / Hq; const int h = (int)(wid % Hq);
    const bf16* op = o + tok * ldo + (long long)h * D;
    const bf16* dp = dout + tok * lddo + (long long)h * D;
    float s = 0.f;
    for (int i = lane * 2; i < D; i += 64) {
        float2 a = br::unpack_bf16(*reinterpret_cast<const uint32_t*>(op + i)), b = br::unpack_bf16(*reinterpret_cast<const uint32_t*>(dp + i));
        s += a.x * b.x + a.y * b.y;
    }
    s = br::warp_sum(s);
    if (lane == 0) { const int b_ = (int)(tok / L), i_ = (int)(tok % L); delta[((long long)b_ * Hq + h) * L + i_] = s; }
}

template <int D>
__global__ void __launch_bounds__(128) attn_bwd_kernel(const BwdParams p) {
    constexpr int BM = 64, BN = 64, TILE = 64 * D * 2;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sK = smem;
    uint8_t* sV = smem + TILE;
    uint8_t* sQ = smem + 2 * TILE;        // 2 stages
    uint8_t* sDO = smem + 4 * TILE;       // 2 stages
    uint8_t* sdS = smem + 6 * TILE;       // 64 x 64 bf16 = 8 KB
    float* sLse = reinterpret_cast<float*>(sdS + 64 * 64 * 2);   // 2 x 64
    float* sDelta = sLse + 128;                                  // 2 x 64

    const int jb = blockIdx.x, hk = blockIdx.y, b = blockIdx.z;
    const int GQ = p.Hq / p.Hkv;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int ks = p.kv_start ? p.kv_start[b] : 0;
    const int ke = p.kv_end ? p.kv_end[b] : p.L;
    const int key0 = jb * BN;
    const long long tok0 = (long long)b * p.L;
    const bf16* K = p.k + tok0 * p.ldk + (long long)hk * D;
    const bf16* V = p.v + tok0 * p.ldv + (long long)hk * D;

    float dk[D / 8][4], dv[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) { dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f; dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f; }

    const bool block_live = (key0 < ke) && (key0 + BN > ks) && (key0 < p.L);
    const int ib_lo = key0 / BM;                                   // causal: first query block that can see key0
    const int n_ib = (p.L + BM - 1) / BM;
    const int iters = block_live ? GQ * (n_ib - ib_lo) : 0;

    load_tile<D>(sK, K, p.ldk, key0, p.L, tid);
    load_tile<D>(sV, V, p.ldv, key0, p.L, tid);
    auto issue = [&](int it, int st) {
        const int h = hk * GQ + it / (n_ib - ib_lo);
        const int ib = ib_lo + it % (n_ib - ib_lo);
        const bf16* Q = p.q + tok0 * p.ldq + (long long)h * D;
        const bf16* DO = p.dout + tok0 * p.lddo + (long long)h * D;
        load_tile<D>(sQ + st * TILE, Q, p.ldq, ib * BM, p.L, tid);
        load_tile<D>(sDO + st * TILE, DO, p.lddo, ib * BM, p.L, tid);
        if (tid < 64) {
            const int i = ib * BM + tid;
            const long long off = ((long long)b * p.Hq + h) * p.L + i;
            sLse[st * 64 + tid] = (i < p.L) ? p.lse[off] : INFINITY;
            sDelta[st * 64 + tid] = (i < p.L) ? p.delta[off] : 0.f;
        }
    };
    if (iters > 0) issue(0, 0);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    const float LOG2E = 1.4426950408889634f;
    for (int it = 0; it < iters; ++it) {
        const int st = it & 1;
        if (it + 1 < iters) issue(it + 1, st ^ 1);
        cp_async_commit();
        const int h = hk * GQ + it / (n_ib - ib_lo);
        const int ib = ib_lo + it % (n_ib - ib_lo);
        const int q0 = ib * BM;
        uint8_t* cQ = sQ + st * TILE;
        uint8_t* cDO = sDO + st * TILE;
        const float* cL = sLse + st * 64;
        const float* cDl = sDelta + st * 64;

        // S^T (16 keys x 64 queries per warp) and dP^T
        float s[BM / 8][4], dp[BM / 8][4];
#pragma unroll
        for (int i = 0; i < BM / 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f; }
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
            uint32_t ka[4], va[4];
            ldsm_x4(ka, tile_ptr<D>(sK, warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, kk * 2 + (lane >> 4)));
            ldsm_x4(va, tile_ptr<D>(sV, warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, kk * 2 + (lane >> 4)));
#pragma unroll
            for (int np = 0; np < BM / 16; ++np) {
                uint32_t qb[4], ob[4];
                ldsm_x4(qb, tile_ptr<D>(cQ, np * 16 + (lane & 7) + (lane >> 4) * 8, kk * 2 + ((lane >> 3) & 1)));
                mma16816(s[2 * np], ka, qb[0], qb[1]);
                mma16816(s[2 * np + 1], ka, qb[2], qb[3]);
                ldsm_x4(ob, tile_ptr<D>(cDO, np * 16 + (lane & 7) + (lane >> 4) * 8, kk * 2 + ((lane >> 3) & 1)));
                mma16816(dp[2 * np], va, ob[0], ob[1]);
                mma16816(dp[2 * np + 1], va, ob[2], ob[3]);
            }
        }
        // P^T and dS^T  (rows = keys key0 + warp*16 + g (+8); cols = queries q0 + nt*8 + 2t (+1))
        uint32_t pf[BM / 16][4], dsf[BM / 16][4];
#pragma unroll
        for (int nt = 0; nt < BM / 8; ++nt) {
            float pv[4], dsv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = key0 + warp * 16 + g + ((e >> 1) ? 8 : 0);
                const int ql = nt * 8 + 2 * t + (e & 1);
                const int i = q0 + ql;
                const bool ok = (j >= ks) && (j < ke) && (j <= i) && (i < p.L);
                const float pr = ok ? exp2f(s[nt][e] * p.scale_log2 - cL[ql] * LOG2E) : 0.f;
                pv[e] = pr;
                dsv[e] = pr * (dp[nt][e] - cDl[ql]) * p.scale;
            }
            pf[nt >> 1][(nt & 1) * 2 + 0] = br::pack_bf16(pv[0], pv[1]);
            pf[nt >> 1][(nt & 1) * 2 + 1] = br::pack_bf16(pv[2], pv[3]);
            const uint32_t d01 = br::pack_bf16(dsv[0], dsv[1]), d23 = br::pack_bf16(dsv[2], dsv[3]);
            dsf[nt >> 1][(nt & 1) * 2 + 0] = d01;
            dsf[nt >> 1][(nt & 1) * 2 + 1] = d23;
            *reinterpret_cast<uint32_t*>(tile_ptr<64>(sdS, warp * 16 + g, nt) + t * 4) = d01;
            *reinterpret_cast<uint32_t*>(tile_ptr<64>(sdS, warp * 16 + g + 8, nt) + t * 4) = d23;
        }
        // dV += P^T dO ; dK += dS^T Q   (k = queries, B operands transposed from the [query][d] tiles)
#pragma unroll
        for (int kk = 0; kk < BM / 16; ++kk) {
#pragma unroll
            for (int dpair = 0; dpair < D / 16; ++dpair) {
                uint32_t f[4];
                ldsm_x4_t(f, tile_ptr<D>(cDO, kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, dpair * 2 + (lane >> 4)));
                mma16816(dv[2 * dpair], pf[kk], f[0], f[1]);
                mma16816(dv[2 * dpair + 1], pf[kk], f[2], f[3]);
                ldsm_x4_t(f, tile_ptr<D>(cQ, kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, dpair * 2 + (lane >> 4)));
                mma16816(dk[2 * dpair], dsf[kk], f[0], f[1]);
                mma16816(dk[2 * dpair + 1], dsf[kk], f[2], f[3]);
            }
        }
        __syncthreads();                                            // sdS complete
        // dQ (16 query rows per warp) += dS[q, keys] K[keys, d]
        {
            float dq[D / 8][4];
#pragma unroll
            for (int i = 0; i < D / 8; ++i) { dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f; }
#pragma unroll
            for (int kk = 0; kk < BN / 16; ++kk) {
                uint32_t a[4];
                ldsm_x4_t(a, tile_ptr<64>(sdS, kk * 16 + (lane & 7) + (lane >> 4) * 8, warp * 2 + ((lane >> 3) & 1)));
#pragma unroll
                for (int dpair = 0; dpair < D / 16; ++dpair) {
                    uint32_t f[4];
                    ldsm_x4_t(f, tile_ptr<D>(sK, kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, dpair * 2 + (lane >> 4)));
                    mma16816(dq[2 * dpair], a, f[0], f[1]);
                    mma16816(dq[2 * dpair + 1], a, f[2], f[3]);
                }
            }
            const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
            float* base = p.dq_acc + (tok0) * (long long)p.Hq * D + (long long)h * D;
#pragma unroll
            for (int dt = 0; dt < D / 8; ++dt) {
                const int c = dt * 8 + 2 * t;
                if (r0 < p.L) { atomicAdd(base + (long long)r0 * p.Hq * D + c, dq[dt][0]); atomicAdd(base + (long long)r0 * p.Hq * D + c + 1, dq[dt][1]); }
                if (r1 < p.L) { atomicAdd(base + (long long)r1 * p.Hq * D + c, dq[dt][2]); atomicAdd(base + (long long)r1 * p.Hq * D + c + 1, dq[dt][3]); }
            }
        }
        cp_async_wait<0>();
        __syncthreads();                                            // next Q/dO stage landed; sdS free again
    }

    // write dK, dV (bf16) for this warp's 16 keys
    bf16* DK = p.dk + tok0 * p.lddk + (long long)hk * D;
    bf16* DV = p.dv + tok0 * p.lddv + (long long)hk * D;
    const int j0 = key0 + warp * 16 + g, j1 = j0 + 8;
#pragma unroll
    for (int dt = 0; dt < D / 8; ++dt) {
        const int c = dt * 8 + 2 * t;
        if (j0 < p.L) {
            *reinterpret_cast<uint32_t*>(DK + (long long)j0 * p.lddk + c) = br::pack_bf16(dk[dt][0], dk[dt][1]);
            *reinterpret_cast<uint32_t*>(DV + (long long)j0 * p.lddv + c) = br::pack_bf16(dv[dt][0], dv[dt][1]);
        }
        if (j1 < p.L) {
            *reinterpret_cast<uint32_t*>(DK + (long long)j1 * p.lddk + c) = br::pack_bf16(dk[dt][2], dk[dt][3]);
            *reinterpret_cast<uint32_t*>(DV + (long long)j1 * p.lddv + c) = br::pack_bf16(dv[dt][2], dv[dt][3]);
        }
    }
}

// dq bf16[tok, h*D + d] (stride lddq) = dq_acc fp32 [tok, Hq*D]
__global__ void dq_convert_kernel(const float* __restrict__ acc, bf16* __restrict__ dq, long long lddq, long long rows, int width) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // one thread per 4 elements
    const long long per_row = width / 4;
    if (i >= rows * per_row) return;
    const long long r = i / per_row; const int c = (int)(i % per_row) * 4;
    const float4 v = *reinterpret_cast<const float4*>(acc + r * width + c);
    uint2 o = make_uint2(br::pack_bf16(v.x, v.y), br::pack_bf16(v.z, v.w));
    *reinterpret_cast<uint2*>(dq + r * lddq + c) = o;
}

}  // namespace

int br_attn_bwd_tc5_impl(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* o, int64_t ldo,
                         const void* dout, int64_t lddo, const float* lse, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv,
                         int B, int L, int n_q_heads, int n_kv_heads, const int32_t* kv_start, const int32_t* kv_end, float scale,
                         void* workspace, cudaStream_t st);

extern "C" {

int64_t br_attn_bwd_workspace_bytes(int B, int L, int n_q_heads, int head_dim) {
    return (int64_t)B * L * n_q_heads * (head_dim + 1) * sizeof(float);
}

int br_attn_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* o, int64_t ldo, const void* dout,
                int64_t lddo, const float* lse, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, int B, int L, int n_q_heads,
                int n_kv_heads, int head_dim, const int32_t* kv_start, const int32_t* kv_end, float scale, void* workspace, void* stream) {
    BR_CHECK_ARG(head_dim == 128, "attn_bwd: head_dim 128 (the decoder) only; the encoder is forward-only (dna_llm.py:121)");
    BR_CHECK_ARG(B > 0 && L > 0 && n_q_heads % n_kv_heads == 0, "attn_bwd: bad shape");
    constexpr int D = 128;
    cudaStream_t st = (cudaStream_t)stream;
    static const bool legacy = getenv("BR_ATTN_LEGACY") != nullptr;       // debugging switch: the previous mma.sync kernel
    if (!legacy)
        return br_attn_bwd_tc5_impl(q, ldq, k, ldk, v, ldv, o, ldo, dout, lddo, lse, dq, lddq, dk, lddk, dv, lddv, B, L, n_q_heads, n_kv_heads,
                                    kv_start, kv_end, scale, workspace, st);
    BwdParams p;
    p.q = (const bf16*)q; p.k = (const bf16*)k; p.v = (const bf16*)v; p.o = (const bf16*)o; p.dout = (const bf16*)dout;
    p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.ldo = ldo; p.lddo = lddo; p.lse = lse;
    p.dq_acc = (float*)workspace; p.delta = p.dq_acc + (int64_t)B * L * n_q_heads * D;
    p.dk = (bf16*)dk; p.dv = (bf16*)dv; p.lddk = lddk; p.lddv = lddv; p.B = B; p.L = L; p.Hq = n_q_heads; p.Hkv = n_kv_heads;
    p.kv_start = kv_start; p.kv_end = kv_end; p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f;
    BR_CHECK_CUDA(cudaMemsetAsync(p.dq_acc, 0, (size_t)B * L * n_q_heads * D * sizeof(float), st));
    {
        const long long warps = (long long)B * L * n_q_heads; const int wpb = 8;
        delta_kernel<D><<<(unsigned)((warps + wpb - 1) / wpb), wpb * 32, 0, st>>>(p.o, ldo, p.dout, lddo, p.delta, B, L, n_q_heads);
        BR_CHECK_LAUNCH();
    }
    constexpr int SMEM = 6 * 64 * D * 2 + 64 * 64 * 2 + 4 * 64 * 4;
    static bool done = false;
    if (!done) { BR_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM)); done = true; }
    dim3 grid((L + 63) / 64, n_kv_heads, B);
    attn_bwd_kernel<D><<<grid, 128, SMEM, st>>>(p);
    BR_CHECK_LAUNCH();
    const long long n4 = (long long)B * L * n_q_heads * D / 4;
    dq_convert_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(p.dq_acc, (bf16*)dq, lddq, (long long)B * L, n_q_heads * D);
    BR_CHECK_LAUNCH();
    return BR_OK;
}

}  // extern "C"
