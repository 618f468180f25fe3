// Row / elementwise kernels of the hand-written decoder backward: RMSNorm backward (+ residual
// gradient add), SwiGLU backward on the blocked gate/up layout, q/k-norm + RoPE backward, the LoRA "X^T Y" gradient
// reduction, bf16 transpose and column sums (projector gradients).  Base weights are frozen (LoRA), so no weight
// gradients exist for the norms or the base linears.
#include "br_common.cuh"
#include "../../include/bioreason_b200.h"


namespace {

__device__ __forceinline__ float rbf(float x) { return __bfloat162float(__float2bfloat16(x)); }

// dx = rstd * (w o dy) - x * rstd^3 * mean(x o w o dy)  (+ dres)
template <int VEC_ITERS>
__global__ void rmsnorm_bwd_kernel(const bf16* __restrict__ x, long long ldx, const bf16* __restrict__ w, const float* __restrict__ rstd,
                                   const bf16* __restrict__ dy, long long lddy, const bf16* __restrict__ dres, long long lddr,
                                   bf16* __restrict__ dx, long long lddx, int M, int d) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const int nvec = d >> 3;
    const uint4* xp = reinterpret_cast<const uint4*>(x + (long long)row * ldx);
    const uint4* gp = reinterpret_cast<const uint4*>(dy + (long long)row * lddy);
    const uint4* wp = reinterpret_cast<const uint4*>(w);
    float xv[VEC_ITERS][8], dn[VEC_ITERS][8];
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < VEC_ITERS; ++i) {
        const int idx = lane + i * 32;
        if (idx < nvec) {
            const uint4 a = xp[idx], g = gp[idx], ww = __ldg(wp + idx);
            const uint32_t as[4] = {a.x, a.y, a.z, a.w}, gs[4] = {g.x, g.y, g.z, g.w}, ws[4] = {ww.x, ww.y, ww.z, ww.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 xf = br::unpack_bf16(as[j]), gf = br::unpack_bf16(gs[j]), wf = br::unpack_bf16(ws[j]);
                xv[i][2 * j] = xf.x; xv[i][2 * j + 1] = xf.y;
                dn[i][2 * j] = gf.x * wf.x; dn[i][2 * j + 1] = gf.y * wf.y;
                dot += xf.x * dn[i][2 * j] + xf.y * dn[i][2 * j + 1];
            }
        }
    }
    dot = br::warp_sum(dot);
    const float r = rstd[row];
    const float c = dot * r * r * r / (float)d;
    const uint4* rp = dres ? reinterpret_cast<const uint4*>(dres + (long long)row * lddr) : nullptr;
    uint4* op = reinterpret_cast<uint4*>(dx + (long long)row * lddx);
#pragma unroll
    for (int i = 0; i < VEC_ITERS; ++i) {
        const int idx = lane + i * 32;
        if (idx < nvec) {
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = r * dn[i][j] - xv[i][j] * c;
            if (rp) {
                const uint4 rr = rp[idx];
                const uint32_t rs[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) { const float2 f = br::unpack_bf16(rs[j]); o[2 * j] += f.x; o[2 * j + 1] += f.y; }
            }
            op[idx] = make_uint4(br::pack_bf16(o[0], o[1]), br::pack_bf16(o[2], o[3]), br::pack_bf16(o[4], o[5]), br::pack_bf16(o[6], o[7]));
        }
    }
}

// gu / dgu: blocks of 16 columns = 8 gate | 8 up;  dact [M, F]
__global__ void swiglu_bwd_kernel(const bf16* __restrict__ gu, long long ldgu, const bf16* __restrict__ dact, long long ldda,
                                  bf16* __restrict__ dgu, long long lddgu, long long M, int F) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // one thread per 8 outputs
    const int per_row = F >> 3;
    if (i >= M * per_row) return;
    const long long m = i / per_row; const int blk = (int)(i % per_row);
    const uint4 gv = *reinterpret_cast<const uint4*>(gu + m * ldgu + blk * 16);
    const uint4 uv = *reinterpret_cast<const uint4*>(gu + m * ldgu + blk * 16 + 8);
    const uint4 dv = *reinterpret_cast<const uint4*>(dact + m * ldda + blk * 8);
    const uint32_t gs[4] = {gv.x, gv.y, gv.z, gv.w}, us[4] = {uv.x, uv.y, uv.z, uv.w}, ds[4] = {dv.x, dv.y, dv.z, dv.w};
    uint32_t og[4], ou[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 g = br::unpack_bf16(gs[j]), u = br::unpack_bf16(us[j]), d = br::unpack_bf16(ds[j]);
        const float s0 = 1.f / (1.f + __expf(-g.x)), s1 = 1.f / (1.f + __expf(-g.y));
        og[j] = br::pack_bf16(d.x * u.x * s0 * (1.f + g.x * (1.f - s0)), d.y * u.y * s1 * (1.f + g.y * (1.f - s1)));
        ou[j] = br::pack_bf16(d.x * g.x * s0, d.y * g.y * s1);
    }
    *reinterpret_cast<uint4*>(dgu + m * lddgu + blk * 16) = make_uint4(og[0], og[1], og[2], og[3]);
    *reinterpret_cast<uint4*>(dgu + m * lddgu + blk * 16 + 8) = make_uint4(ou[0], ou[1], ou[2], ou[3]);
}

// in place on dqkv (q and k heads): inverse rotation, then the per-head RMSNorm backward using the saved pre-norm values
template <int D>
__global__ void qk_rope_bwd_kernel(bf16* __restrict__ dqkv, long long ldd, const bf16* __restrict__ pre, long long ldp, int M, int n_q, int n_k,
                                   const bf16* __restrict__ qw, const bf16* __restrict__ kw, const int* __restrict__ pos, float theta, float eps) {
    constexpr int E = D / 64;
    const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int heads = n_q + n_k;
    if (wid >= (long long)M * heads) return;
    const long long m = wid / heads; const int h = (int)(wid % heads);
    bf16* gp = dqkv + m * ldd + (long long)h * D;
    const bf16* xp = pre + m * ldp + (long long)h * D;
    const bf16* w = (h < n_q) ? qw : kw;
    const float position = (float)pos[m];
    float da[E], db[E], xa[E], xb[E], dna[E], dnb[E];
    float ss = 0.f, dot = 0.f;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int j = lane * E + e;
        const float inv_freq = 1.0f / powf(theta, (float)(2 * j) / (float)D);
        float sn, cs;
        sincosf(position * inv_freq, &sn, &cs);
        sn = rbf(sn); cs = rbf(cs);
        const float glo = __bfloat162float(gp[j]), ghi = __bfloat162float(gp[D / 2 + j]);
        da[e] = glo * cs + ghi * sn;                     // R(-theta) dy
        db[e] = -glo * sn + ghi * cs;
        xa[e] = __bfloat162float(xp[j]); xb[e] = __bfloat162float(xp[D / 2 + j]);
        ss += xa[e] * xa[e] + xb[e] * xb[e];
        dna[e] = da[e] * __bfloat162float(w[j]); dnb[e] = db[e] * __bfloat162float(w[D / 2 + j]);
        dot += dna[e] * xa[e] + dnb[e] * xb[e];
    }
    ss = br::warp_sum(ss); dot = br::warp_sum(dot);
    const float r = rsqrtf(ss / (float)D + eps);
    const float c = dot * r * r * r / (float)D;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int j = lane * E + e;
        gp[j] = __float2bfloat16(r * dna[e] - xa[e] * c);
        gp[D / 2 + j] = __float2bfloat16(r * dnb[e] - xb[e] * c);
    }
}

// out[N, M] = in[M, N]^T (bf16), rows of `out` beyond... (plain tiled transpose; out row stride ldo >= M)
__global__ void transpose_kernel(const bf16* __restrict__ in, long long ldi, bf16* __restrict__ out, long long ldo, int M, int N) {
    __shared__ bf16 tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int r = by + j, c = bx + threadIdx.x;
        tile[j][threadIdx.x] = (r < M && c < N) ? in[(long long)r * ldi + c] : __float2bfloat16(0.f);
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int r = bx + j, c = by + threadIdx.x;      // out[r = n][c = m]
        if (r < N && c < M) out[(long long)r * ldo + c] = tile[threadIdx.x][j];
    }
}

// out[n] += sum_m in[m, n]  (fp32 out).  CTA = 32 columns x 8 row groups; each row group sums its rows in order, the 8 partials
// are added in a fixed order through shared memory: no atomics, bit-reproducible.
__global__ void __launch_bounds__(256) colsum_kernel(const bf16* __restrict__ in, long long ldi, float* __restrict__ out, int M, int N) {
    __shared__ float part[8][32];
    const int c = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int n = blockIdx.x * 32 + c;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (n < N) {
        int m = g;
        for (; m + 24 < M; m += 32) {                      // 4 independent loads in flight per thread
            s0 += __bfloat162float(in[(long long)m * ldi + n]); s1 += __bfloat162float(in[(long long)(m + 8) * ldi + n]);
            s2 += __bfloat162float(in[(long long)(m + 16) * ldi + n]); s3 += __bfloat162float(in[(long long)(m + 24) * ldi + n]);
        }
        for (; m < M; m += 8) s0 += __bfloat162float(in[(long long)m * ldi + n]);
    }
    part[g][c] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (g == 0 && n < N) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += part[i][c];
        out[n] += t;
    }
}

}  // namespace

extern "C" {

int br_rmsnorm_bwd(const void* x, int64_t ldx, const void* w, const float* rstd, const void* dy, int64_t lddy, const void* dres, int64_t lddr,
                   void* dx, int64_t lddx, int M, int d, void* stream) {
    BR_CHECK_ARG(M > 0 && d % 8 == 0 && d <= 32 * 8 * 16, "rmsnorm_bwd: d=%d must be a multiple of 8, <= 4096", d);
    const int iters = (d / 8 + 31) / 32;
    const int wpb = 8;
    dim3 grid((M + wpb - 1) / wpb);
    cudaStream_t st = (cudaStream_t)stream;
#define RB(V) rmsnorm_bwd_kernel<V><<<grid, wpb * 32, 0, st>>>((const bf16*)x, ldx, (const bf16*)w, rstd, (const bf16*)dy, lddy, (const bf16*)dres, lddr, (bf16*)dx, lddx, M, d)
    if (iters <= 1) RB(1); else if (iters <= 4) RB(4); else if (iters <= 8) RB(8); else RB(16);
#undef RB
    BR_CHECK_LAUNCH();
    return BR_OK;
}

int br_swiglu_bwd(const void* gu, int64_t ldgu, const void* dact, int64_t ldda, void* dgu, int64_t lddgu, int M, int F, void* stream) {
    BR_CHECK_ARG(M > 0 && F % 8 == 0, "swiglu_bwd: F %% 8");
    const long long n = (long long)M * (F / 8);
    swiglu_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)gu, ldgu, (const bf16*)dact, ldda, (bf16*)dgu, lddgu, M, F);
    BR_CHECK_LAUNCH();
    return BR_OK;
}

int br_qk_rope_bwd(void* dqkv, int64_t ldd, const void* qk_pre, int64_t ldp, int M, int n_q_heads, int n_k_heads, int head_dim, const void* q_norm_w,
                   const void* k_norm_w, const int32_t* positions, float theta, float eps, void* stream) {
    BR_CHECK_ARG(head_dim == 128 && q_norm_w && k_norm_w, "qk_rope_bwd: Qwen3 layout (head_dim 128, q/k norms) only");
    const long long warps = (long long)M * (n_q_heads + n_k_heads); const int wpb = 8;
    qk_rope_bwd_kernel<128><<<(unsigned)((warps + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
        (bf16*)dqkv, ldd, (const bf16*)qk_pre, ldp, M, n_q_heads, n_k_heads, (const bf16*)q_norm_w, (const bf16*)k_norm_w, positions, theta, eps);
    BR_CHECK_LAUNCH();
    return BR_OK;
}

int br_transpose_bf16(const void* in, int64_t ldi, void* out, int64_t ldo, int M, int N, void* stream) {
    BR_CHECK_ARG(M > 0 && N > 0, "transpose: empty");
    dim3 grid((N + 31) / 32, (M + 31) / 32), block(32, 8);
    transpose_kernel<<<grid, block, 0, (cudaStream_t)stream>>>((const bf16*)in, ldi, (bf16*)out, ldo, M, N);
    BR_CHECK_LAUNCH();
    return BR_OK;
}

int br_colsum_accumulate(const void* in, int64_t ldi, float* out, int M, int N, void* stream) {
    BR_CHECK_ARG(M > 0 && N > 0, "colsum: empty");
    colsum_kernel<<<(N + 31) / 32, 256, 0, (cudaStream_t)stream>>>((const bf16*)in, ldi, out, M, N);
    BR_CHECK_LAUNCH();
    return BR_OK;
}

}  // extern "C"
