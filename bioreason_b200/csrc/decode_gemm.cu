// Skinny GEMM for autoregressive decode:  out[R, N] = X[R, K] . W[N, K]^T  with R <= 32 live rows (the G samples of a
// prompt group).  Each decode step streams every weight byte once, so this kernel is HBM-bound (8 GB
// per step for Qwen3-4B): the design goal is bytes in flight, not FLOPs.
//   * W rows are the M operand of mma.sync.m16n8k16 (16 output features per warp tile), the R rows of X are the N
//     operand (8 per mma) -- i.e. the swap-AB form, so no tensor-core lanes are wasted on padding rows;
//   * every thread issues 16-byte loads; the four lanes of a quad cover 64 contiguous bytes of a weight row and the
//     four warps of a CTA interleave 64-byte segments (256 contiguous bytes per row per round).  The k index inside
//     a 32-wide block is permuted identically for W and X, which leaves the dot product unchanged and lets the
//     16-byte register chunks feed the mma fragments directly (no shared-memory staging of weights);
//   * split-K across CTAs with fp32 atomics into a scratch tile + "last CTA does the epilogue and re-zeros" so that
//     small-N layers (o_proj, down_proj: 160 row tiles) still put >1000 CTAs in flight.
// Epilogues: bf16 store, +residual, SwiGLU on (8 gate | 8 up) row blocks, fp32 logits.
#include "br_common.cuh"
#include "../../include/bioreason_b200.h"

namespace {

__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

struct SkinnyParams {
    const bf16* W; long long ldw;
    const bf16* X; long long ldx;
    int R, N, K, ksplit;
    int mode;                 // 0 bf16, 1 bf16 + residual, 2 SwiGLU pairs, 3 fp32
    void* out; long long ldo;
    const bf16* res; long long ldr;
    float* scratch;           // [32, N] fp32, zero between launches (split-K only)
    int* counters;            // [N/16], zero between launches
};

__device__ __forceinline__ float rbf(float x) { return __bfloat162float(__float2bfloat16(x)); }

template <int NB>
__device__ __forceinline__ void epilogue(const SkinnyParams& p, const float* tile /* [8*NB][16] in smem or registers-view */, int n0) {
    const int R = p.R;
    if (p.mode == 2) {
        for (int idx = threadIdx.x; idx < 8 * NB * 8; idx += blockDim.x) {
            const int r = idx >> 3, j = idx & 7;
            if (r >= R) continue;
            const float g = rbf(tile[r * 16 + j]), u = rbf(tile[r * 16 + 8 + j]);   // rows 0-7 gate, 8-15 up
            const float sg = rbf(g / (1.f + __expf(-g)));
            reinterpret_cast<bf16*>(p.out)[(long long)r * p.ldo + (n0 >> 1) + j] = __float2bfloat16(sg * u);
        }
        return;
    }
    for (int idx = threadIdx.x; idx < 8 * NB * 16; idx += blockDim.x) {
        const int r = idx >> 4, n = idx & 15;
        if (r >= R) continue;
        float v = tile[r * 16 + n];
        if (p.mode == 3) { reinterpret_cast<float*>(p.out)[(long long)r * p.ldo + n0 + n] = v; continue; }
        if (p.mode == 1) v = rbf(v) + __bfloat162float(p.res[(long long)r * p.ldr + n0 + n]);
        reinterpret_cast<bf16*>(p.out)[(long long)r * p.ldo + n0 + n] = __float2bfloat16(v);
    }
}

template <int NB>
__global__ void __launch_bounds__(128) skinny_gemm_kernel(const SkinnyParams p) {
    __shared__ float red[4][8 * NB][16];
    __shared__ int s_last;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int n0 = blockIdx.x * 16;
    const int nblk = p.K >> 5;                                   // 32-wide k blocks
    const int per = (nblk + p.ksplit - 1) / p.ksplit;
    const int kb_lo = blockIdx.y * per, kb_hi = min(nblk, kb_lo + per);

    float c[NB][4];
#pragma unroll
    for (int i = 0; i < NB; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;

    const bf16* w0 = p.W + (long long)(n0 + g) * p.ldw + t * 8;
    const bf16* w1 = w0 + 8 * p.ldw;
    const bf16* xr[NB];
    bool xok[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) { xok[i] = (i * 8 + g) < p.R; xr[i] = p.X + (long long)(xok[i] ? i * 8 + g : 0) * p.ldx + t * 8; }

    constexpr int UN = 4;
    int kb = kb_lo + warp;
    for (; kb + 4 * (UN - 1) < kb_hi; kb += 4 * UN) {
        uint4 a[UN], b[UN], x[UN][NB];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const long long ko = (long long)(kb + 4 * u) << 5;
            a[u] = ldg_stream(w0 + ko);
            b[u] = ldg_stream(w1 + ko);
#pragma unroll
            for (int i = 0; i < NB; ++i) x[u][i] = xok[i] ? *reinterpret_cast<const uint4*>(xr[i] + ko) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                mma16816(c[i], a[u].x, b[u].x, a[u].y, b[u].y, x[u][i].x, x[u][i].y);
                mma16816(c[i], a[u].z, b[u].z, a[u].w, b[u].w, x[u][i].z, x[u][i].w);
            }
        }
    }
    for (; kb < kb_hi; kb += 4) {
        const long long ko = (long long)kb << 5;
        const uint4 a = ldg_stream(w0 + ko), b = ldg_stream(w1 + ko);
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const uint4 x = xok[i] ? *reinterpret_cast<const uint4*>(xr[i] + ko) : make_uint4(0, 0, 0, 0);
            mma16816(c[i], a.x, b.x, a.y, b.y, x.x, x.y);
            mma16816(c[i], a.z, b.z, a.w, b.w, x.z, x.w);
        }
    }
    // c[i][0..1] = (feature n0+g, rows 8i+2t, +1); c[i][2..3] = (feature n0+g+8, same rows)
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        red[warp][i * 8 + 2 * t][g] = c[i][0];
        red[warp][i * 8 + 2 * t + 1][g] = c[i][1];
        red[warp][i * 8 + 2 * t][g + 8] = c[i][2];
        red[warp][i * 8 + 2 * t + 1][g + 8] = c[i][3];
    }
    __syncthreads();
    float* tile = &red[0][0][0];
    for (int idx = threadIdx.x; idx < 8 * NB * 16; idx += blockDim.x)
        tile[idx] = red[0][0][idx] + (&red[1][0][0])[idx] + (&red[2][0][0])[idx] + (&red[3][0][0])[idx];
    __syncthreads();
    if (p.ksplit > 1) {
        for (int idx = threadIdx.x; idx < 8 * NB * 16; idx += blockDim.x) {
            const int r = idx >> 4, n = idx & 15;
            if (r < p.R) atomicAdd(p.scratch + (long long)r * p.N + n0 + n, tile[idx]);
        }
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) s_last = (atomicAdd(p.counters + blockIdx.x, 1) == p.ksplit - 1);
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        for (int idx = threadIdx.x; idx < 8 * NB * 16; idx += blockDim.x) {
            const int r = idx >> 4, n = idx & 15;
            float* sp = p.scratch + (long long)r * p.N + n0 + n;
            tile[idx] = (r < p.R) ? __ldcg(sp) : 0.f;
            if (r < p.R) __stcg(sp, 0.f);                        // leave the scratch clean for the next launch
        }
        if (threadIdx.x == 0) p.counters[blockIdx.x] = 0;
        __syncthreads();
    }
    epilogue<NB>(p, tile, n0);
}

}  // namespace

extern "C" {

int64_t br_skinny_scratch_bytes(int max_N) { return (int64_t)32 * max_N * sizeof(float) + (int64_t)(max_N / 16 + 1) * sizeof(int); }

int br_skinny_gemm(const void* X, int64_t ldx, const void* W, int64_t ldw, void* out, int64_t ldo, int R, int N, int K, int mode,
                   const void* residual, int64_t ldr, void* scratch, void* stream) {
    BR_CHECK_ARG(R >= 1 && R <= 32, "skinny_gemm: R=%d must be in [1, 32]", R);
    BR_CHECK_ARG(N % 16 == 0 && K % 32 == 0 && ldx % 8 == 0 && ldw % 8 == 0, "skinny_gemm: N %% 16, K %% 32, ld %% 8 (N=%d K=%d)", N, K);
    BR_CHECK_ARG(mode >= 0 && mode <= 3 && !(mode == 1 && !residual), "skinny_gemm: bad mode %d", mode);
    BR_CHECK_ARG(scratch != nullptr, "skinny_gemm: scratch (br_skinny_scratch_bytes, zero-initialised once) is required");
    SkinnyParams p;
    p.W = (const bf16*)W; p.ldw = ldw; p.X = (const bf16*)X; p.ldx = ldx; p.R = R; p.N = N; p.K = K; p.mode = mode;
    p.out = out; p.ldo = ldo; p.res = (const bf16*)residual; p.ldr = ldr;
    p.scratch = (float*)scratch; p.counters = (int*)((float*)scratch + (int64_t)32 * N);
    // enough CTAs for ~8 per SM, but keep >= 8 k-blocks (256 k) per warp
    const int tiles = N / 16, nblk = K / 32;
    int ks = 1;
    const int target = 8 * br_num_sms();
    while (tiles * ks * 2 <= target && nblk / (ks * 2) >= 4 * 8) ks *= 2;
    p.ksplit = ks;
    dim3 grid(tiles, ks);
    cudaStream_t st = (cudaStream_t)stream;
    const int NB = (R + 7) / 8;
    if (NB == 1) skinny_gemm_kernel<1><<<grid, 128, 0, st>>>(p);
    else if (NB == 2) skinny_gemm_kernel<2><<<grid, 128, 0, st>>>(p);
    else skinny_gemm_kernel<4><<<grid, 128, 0, st>>>(p);
    BR_CHECK_LAUNCH();
    return BR_OK;
}

}  // extern "C"
