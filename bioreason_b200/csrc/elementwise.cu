// Memory-bound row kernels of the hot path: RMSNorm / LayerNorm, per-head q/k RMSNorm + RoPE, embedding gather,
// masked row scatter.  All are one-warp-per-row (or per-head) with 16-byte vector loads; bf16 storage, fp32 math.
// Rounding points follow HF so that the bf16 regime matches the reference's (HF qwen3/modeling_qwen3.py:50-64,
// :120-150; esm/modeling_esm.py:45-55,331-345).
#include "br_common.cuh"
#include "../../include/bioreason_b200.h"

namespace {

__device__ __forceinline__ float rbf(float x) { return __bfloat162float(__float2bfloat16(x)); }

// ---------------------------------------------------------------- RMSNorm (HF Qwen3RMSNorm)
// y = w * bf16(x * rsqrt(mean(x^2) + eps));   optional residual-free; stores rstd for the backward
template <int VEC_ITERS>
__global__ void rmsnorm_kernel(const bf16* __restrict__ x, long long ldx, const bf16* __restrict__ w, bf16* __restrict__ y, long long ldy,
                               float* __restrict__ rstd_out, int M, int d, float eps) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const uint4* xp = reinterpret_cast<const uint4*>(x + (long long)row * ldx);
    const int nvec = d >> 3;
    uint4 v[VEC_ITERS];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < VEC_ITERS; ++i) {
        int idx = lane + i * 32;
        if (idx < nvec) {
            v[i] = xp[idx];
            float2 a = br::unpack_bf16(v[i].x), b = br::unpack_bf16(v[i].y), c = br::unpack_bf16(v[i].z), e = br::unpack_bf16(v[i].w);
            ss += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y + c.x * c.x + c.y * c.y + e.x * e.x + e.y * e.y;
        }
    }
    ss = br::warp_sum(ss);
    const float rstd = rsqrtf(ss / (float)d + eps);
    if (rstd_out && lane == 0) rstd_out[row] = rstd;
    const uint4* wp = reinterpret_cast<const uint4*>(w);
    uint4* yp = reinterpret_cast<uint4*>(y + (long long)row * ldy);
#pragma unroll
    for (int i = 0; i < VEC_ITERS; ++i) {
        int idx = lane + i * 32;
        if (idx < nvec) {
            uint4 ww = __ldg(wp + idx);
            uint32_t xs[4] = {v[i].x, v[i].y, v[i].z, v[i].w}, ws[4] = {ww.x, ww.y, ww.z, ww.w}, o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float2 xv = br::unpack_bf16(xs[j]), wv = br::unpack_bf16(ws[j]);
                o[j] = br::pack_bf16(wv.x * rbf(xv.x * rstd), wv.y * rbf(xv.y * rstd));
            }
            yp[idx] = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
}

// ---------------------------------------------------------------- LayerNorm (torch.nn.LayerNorm, fp32 statistics)
template <int VEC_ITERS>
__global__ void layernorm_kernel(const bf16* __restrict__ x, long long ldx, const bf16* __restrict__ w, const bf16* __restrict__ b,
                                 bf16* __restrict__ y, long long ldy, int M, int d, float eps) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const uint4* xp = reinterpret_cast<const uint4*>(x + (long long)row * ldx);
    const int nvec = d >> 3;
    float f[VEC_ITERS][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VEC_ITERS; ++i) {
        int idx = lane + i * 32;
        if (idx < nvec) {
            uint4 v = xp[idx];
            float2 a = br::unpack_bf16(v.x), bb = br::unpack_bf16(v.y), c = br::unpack_bf16(v.z), e = br::unpack_bf16(v.w);
            f[i][0] = a.x; f[i][1] = a.y; f[i][2] = bb.x; f[i][3] = bb.y; f[i][4] = c.x; f[i][5] = c.y; f[i][6] = e.x; f[i][7] = e.y;
#pragma unroll
            for (int j = 0; j < 8; ++j) s += f[i][j];
        }
    }
    const float mean = br::warp_sum(s) / (float)d;
    float vs = 0.f;
#pragma unroll
    for (int i = 0; i < VEC_ITERS; ++i) {
        int idx = lane + i * 32;
        if (idx < nvec) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { float t = f[i][j] - mean; vs += t * t; }
        }
    }
    const float rstd = rsqrtf(br::warp_sum(vs) / (float)d + eps);
    const uint4* wp = reinterpret_cast<const uint4*>(w);
    const uint4* bp = reinterpret_cast<const uint4*>(b);
    uint4* yp = reinterpret_cast<uint4*>(y + (long long)row * ldy);
#pragma unroll
    for (int i = 0; i < VEC_ITERS; ++i) {
        int idx = lane + i * 32;
        if (idx < nvec) {
            uint4 ww = __ldg(wp + idx), bv = __ldg(bp + idx);
            uint32_t ws[4] = {ww.x, ww.y, ww.z, ww.w}, bs[4] = {bv.x, bv.y, bv.z, bv.w}, o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float2 wv = br::unpack_bf16(ws[j]), bb = br::unpack_bf16(bs[j]);
                o[j] = br::pack_bf16((f[i][2 * j] - mean) * rstd * wv.x + bb.x, (f[i][2 * j + 1] - mean) * rstd * wv.y + bb.y);
            }
            yp[idx] = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
}

// ---------------------------------------------------------------- per-head q/k norm + rotary, in place on the fused QKV buffer
// One warp per (token, head); head_dim D = 32 lanes * (D/32) elements, lane owns the pair (j, j + D/2) pattern:
// lane l holds elements [l*E, l*E+E) of the first half and the same of the second half, E = D/64.
// mode 0 (Qwen3): x = w * bf16(x * rstd) (if w), then bf16-rounded rotate-half RoPE with bf16 cos/sin (HF rounding points)
// mode 1 (ESM)  : x = x * qscale (q heads only), then fp32 RoPE with one final rounding
template <int D>
__global__ void qk_rope_kernel(bf16* qkv, long long ld, int M, int n_q, int n_k, const bf16* __restrict__ qw,
                               const bf16* __restrict__ kw, const int* __restrict__ pos, float theta, float eps, float qscale, int mode,
                               bf16* out, long long ldo, const float2* __restrict__ rope, int rope_n_pos) {
    constexpr int E = D / 64;
    const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int heads = n_q + n_k;
    if (wid >= M * heads) return;
    const int m = wid / heads, h = wid % heads;
    const bf16* p = qkv + (long long)m * ld + (long long)h * D;
    bf16* o = out ? out + (long long)m * ldo + (long long)h * D : qkv + (long long)m * ld + (long long)h * D;
    float lo[E], hi[E];
#pragma unroll
    for (int e = 0; e < E; ++e) { lo[e] = __bfloat162float(p[lane * E + e]); hi[e] = __bfloat162float(p[D / 2 + lane * E + e]); }
    const float position = (float)pos[m];
    if (mode == 0) {
        const bf16* w = (h < n_q) ? qw : kw;
        if (w) {
            float ss = 0.f;
#pragma unroll
            for (int e = 0; e < E; ++e) ss += lo[e] * lo[e] + hi[e] * hi[e];
            const float rstd = rsqrtf(br::warp_sum(ss) / (float)D + eps);
#pragma unroll
            for (int e = 0; e < E; ++e) {
                lo[e] = rbf(__bfloat162float(w[lane * E + e]) * rbf(lo[e] * rstd));
                hi[e] = rbf(__bfloat162float(w[D / 2 + lane * E + e]) * rbf(hi[e] * rstd));
            }
        }
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int j = lane * E + e;
            float sn, cs;
            if (rope && pos[m] < rope_n_pos) {                       // table built once by br_rope_table (same arithmetic, bf16-rounded)
                const float2 t = __ldg(rope + (long long)pos[m] * (D / 2) + j);
                cs = t.x; sn = t.y;
            } else {
                const float inv_freq = 1.0f / powf(theta, (float)(2 * j) / (float)D);
                sincosf(position * inv_freq, &sn, &cs);
                sn = rbf(sn); cs = rbf(cs);
            }
            const float a = lo[e], b = hi[e];
            o[j] = __float2bfloat16(rbf(a * cs) + rbf(-b * sn));
            o[D / 2 + j] = __float2bfloat16(rbf(b * cs) + rbf(a * sn));
        }
    } else {
        const float sc = (h < n_q) ? qscale : 1.f;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int j = lane * E + e;
            const float inv_freq = 1.0f / powf(theta, (float)(2 * j) / (float)D);
            float sn, cs;
            sincosf(position * inv_freq, &sn, &cs);
            const float a = rbf(lo[e] * sc), b = rbf(hi[e] * sc);
            o[j] = __float2bfloat16(a * cs - b * sn);
            o[D / 2 + j] = __float2bfloat16(b * cs + a * sn);
        }
    }
}

// ---------------------------------------------------------------- embedding gather: out[m] = table[ids[m]] * (mult ? mult[m] : 1)
__global__ void embed_gather_kernel(const long long* __restrict__ ids, const bf16* __restrict__ table, long long ldt, bf16* __restrict__ out,
                                    long long ldo, int M, int d, const int* __restrict__ keep, long long vocab, float* __restrict__ sumsq) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    br::launch_dependents();
    br::grid_dep_wait();
    if (row >= M) return;
    long long id = __ldcg(ids + row);
    const bool zero = (keep && keep[row] == 0) || id < 0 || id >= vocab;
    const uint4* src = reinterpret_cast<const uint4*>(table + (zero ? 0 : id) * ldt);
    uint4* dst = reinterpret_cast<uint4*>(out + (long long)row * ldo);
    float ss = 0.f;
    for (int i = lane; i < (d >> 3); i += 32) {
        const uint4 v = zero ? make_uint4(0, 0, 0, 0) : __ldg(src + i);
        dst[i] = v;
        if (sumsq) {
            const float2 a = br::unpack_bf16(v.x), b = br::unpack_bf16(v.y), c = br::unpack_bf16(v.z), e = br::unpack_bf16(v.w);
            ss += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y + c.x * c.x + c.y * c.y + e.x * e.x + e.y * e.y;
        }
    }
    if (sumsq) { ss = br::warp_sum(ss); if (lane == 0) sumsq[row] = ss; }
}

// W[n, k] *= s[k]   (fold a norm weight into the columns of a frozen / merged rollout weight)
__global__ void scale_columns_kernel(bf16* __restrict__ W, long long ld, long long N, int K, const bf16* __restrict__ s) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // one thread per 8 elements
    const int per_row = K >> 3;
    if (i >= N * per_row) return;
    const long long n = i / per_row; const int c = (int)(i % per_row) * 8;
    uint4* p = reinterpret_cast<uint4*>(W + n * ld + c);
    const uint4 v = *p, sv = __ldg(reinterpret_cast<const uint4*>(s + c));
    const uint32_t vs[4] = {v.x, v.y, v.z, v.w}, ss[4] = {sv.x, sv.y, sv.z, sv.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 a = br::unpack_bf16(vs[j]), b = br::unpack_bf16(ss[j]); o[j] = br::pack_bf16(a.x * b.x, a.y * b.y); }
    *p = make_uint4(o[0], o[1], o[2], o[3]);
}

// rows of src copied to dst[row_map[m]] (row_map < 0 skipped)
__global__ void scatter_rows_kernel(const bf16* __restrict__ src, long long lds, const int* __restrict__ row_map, bf16* __restrict__ dst,
                                    long long ldd, int M, int d) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const int o = row_map[row];
    if (o < 0) return;
    const uint4* s = reinterpret_cast<const uint4*>(src + (long long)row * lds);
    uint4* t = reinterpret_cast<uint4*>(dst + (long long)o * ldd);
    for (int i = lane; i < (d >> 3); i += 32) t[i] = s[i];
}

// out[m] = src[idx[m]]  (idx < 0 -> zeros)
__global__ void gather_rows_kernel(const bf16* __restrict__ src, long long lds, const int* __restrict__ idx, bf16* __restrict__ dst,
                                   long long ldd, int M, int d) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const int ix = idx[row];
    const uint4* s = reinterpret_cast<const uint4*>(src + (long long)(ix < 0 ? 0 : ix) * lds);
    uint4* t = reinterpret_cast<uint4*>(dst + (long long)row * ldd);
    for (int i = lane; i < (d >> 3); i += 32) t[i] = ix < 0 ? make_uint4(0, 0, 0, 0) : __ldg(s + i);
}

}  // namespace

extern "C" {

#define ROW_LAUNCH(kern, M, ...)                                                  \
    do {                                                                          \
        const int wpb = 8;                                                        \
        kern<<<((M) + wpb - 1) / wpb, wpb * 32, 0, (cudaStream_t)stream>>>(__VA_ARGS__); \
        BR_CHECK_LAUNCH();                                                        \
    } while (0)

int br_rmsnorm(const void* x, int64_t ldx, const void* w, void* y, int64_t ldy, float* rstd, int M, int d, float eps, void* stream) {
    BR_CHECK_ARG(M > 0 && d % 8 == 0 && d <= 32 * 8 * 40 && ldx % 8 == 0 && ldy % 8 == 0, "rmsnorm: d=%d must be a multiple of 8, <= 10240", d);
    const int iters = (d / 8 + 31) / 32;
    const bf16 *xp = (const bf16*)x, *wp = (const bf16*)w; bf16* yp = (bf16*)y;
    if (iters <= 1) ROW_LAUNCH(rmsnorm_kernel<1>, M, xp, ldx, wp, yp, ldy, rstd, M, d, eps);
    else if (iters <= 4) ROW_LAUNCH(rmsnorm_kernel<4>, M, xp, ldx, wp, yp, ldy, rstd, M, d, eps);
    else if (iters <= 8) ROW_LAUNCH(rmsnorm_kernel<8>, M, xp, ldx, wp, yp, ldy, rstd, M, d, eps);
    else if (iters <= 16) ROW_LAUNCH(rmsnorm_kernel<16>, M, xp, ldx, wp, yp, ldy, rstd, M, d, eps);
    else ROW_LAUNCH(rmsnorm_kernel<40>, M, xp, ldx, wp, yp, ldy, rstd, M, d, eps);
    return BR_OK;
}

int br_layernorm(const void* x, int64_t ldx, const void* w, const void* b, void* y, int64_t ldy, int M, int d, float eps, void* stream) {
    BR_CHECK_ARG(M > 0 && d % 8 == 0 && d <= 32 * 8 * 8 && ldx % 8 == 0 && ldy % 8 == 0, "layernorm: d=%d must be a multiple of 8, <= 2048", d);
    const int iters = (d / 8 + 31) / 32;
    const bf16 *xp = (const bf16*)x, *wp = (const bf16*)w, *bp = (const bf16*)b; bf16* yp = (bf16*)y;
    if (iters <= 1) ROW_LAUNCH(layernorm_kernel<1>, M, xp, ldx, wp, bp, yp, ldy, M, d, eps);
    else if (iters <= 4) ROW_LAUNCH(layernorm_kernel<4>, M, xp, ldx, wp, bp, yp, ldy, M, d, eps);
    else ROW_LAUNCH(layernorm_kernel<8>, M, xp, ldx, wp, bp, yp, ldy, M, d, eps);
    return BR_OK;
}

int br_qk_rope_ex(void* qkv, int64_t ld, void* out, int64_t ldo, int M, int n_q_heads, int n_k_heads, int head_dim, const void* q_norm_w,
                  const void* k_norm_w, const int32_t* positions, float theta, float eps, float q_scale, int mode, const float* rope_table,
                  int rope_n_pos, void* stream) {
    BR_CHECK_ARG(M > 0 && (head_dim == 128 || head_dim == 64), "qk_rope: head_dim must be 64 or 128 (got %d)", head_dim);
    BR_CHECK_ARG(!(rope_table && mode != 0), "qk_rope: the cos/sin table holds the bf16-rounded Qwen3 values (mode 0) only");
    BR_CHECK_ARG(!out || ldo % 8 == 0, "qk_rope: ldo %% 8");
    const long long warps = (long long)M * (n_q_heads + n_k_heads);
    const int wpb = 8;
    const unsigned grid = (unsigned)((warps + wpb - 1) / wpb);
    if (head_dim == 128)
        qk_rope_kernel<128><<<grid, wpb * 32, 0, (cudaStream_t)stream>>>((bf16*)qkv, ld, M, n_q_heads, n_k_heads, (const bf16*)q_norm_w,
                                                                         (const bf16*)k_norm_w, positions, theta, eps, q_scale, mode,
                                                                         (bf16*)out, ldo, (const float2*)rope_table, rope_n_pos);
    else
        qk_rope_kernel<64><<<grid, wpb * 32, 0, (cudaStream_t)stream>>>((bf16*)qkv, ld, M, n_q_heads, n_k_heads, (const bf16*)q_norm_w,
                                                                        (const bf16*)k_norm_w, positions, theta, eps, q_scale, mode,
                                                                        (bf16*)out, ldo, (const float2*)rope_table, rope_n_pos);
    BR_CHECK_LAUNCH();
    return BR_OK;
}

int br_qk_rope(void* qkv, int64_t ld, int M, int n_q_heads, int n_k_heads, int head_dim, const void* q_norm_w, const void* k_norm_w,
               const int32_t* positions, float theta, float eps, float q_scale, int mode, void* stream) {
    return br_qk_rope_ex(qkv, ld, nullptr, 0, M, n_q_heads, n_k_heads, head_dim, q_norm_w, k_norm_w, positions, theta, eps, q_scale, mode, nullptr, 0, stream);
}

int br_embed_gather(const int64_t* ids, const void* table, int64_t ldt, int64_t vocab, void* out, int64_t ldo, int M, int d, const int32_t* keep,
                    void* stream) {
    BR_CHECK_ARG(M > 0 && d % 8 == 0 && ldt % 8 == 0 && ldo % 8 == 0, "embed_gather: d %% 8");
    ROW_LAUNCH(embed_gather_kernel, M, (const long long*)ids, (const bf16*)table, ldt, (bf16*)out, ldo, M, d, keep, (long long)vocab, (float*)nullptr);
    return BR_OK;
}

int br_embed_gather_sumsq(const int64_t* ids, const void* table, int64_t ldt, int64_t vocab, void* out, int64_t ldo, int M, int d, float* sumsq,
                          void* stream) {
    BR_CHECK_ARG(M > 0 && d % 8 == 0 && ldt % 8 == 0 && ldo % 8 == 0, "embed_gather_sumsq: d %% 8");
    BR_CHECK_CUDA(br_launch_pdl(embed_gather_kernel, dim3((M + 7) / 8), dim3(256), 0, (cudaStream_t)stream, (const long long*)ids, (const bf16*)table,
                                (long long)ldt, (bf16*)out, (long long)ldo, M, d, (const int*)nullptr, (long long)vocab, sumsq));
    return BR_OK;
}

int br_scale_columns(void* W, int64_t ld, int64_t N, int K, const void* scale, void* stream) {
    BR_CHECK_ARG(N > 0 && K % 8 == 0 && ld % 8 == 0, "scale_columns: K %% 8");
    const long long n = N * (K / 8);
    scale_columns_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((bf16*)W, ld, N, K, (const bf16*)scale);
    BR_CHECK_LAUNCH();
    return BR_OK;
}

int br_scatter_rows(const void* src, int64_t lds, const int32_t* row_map, void* dst, int64_t ldd, int M, int d, void* stream) {
    BR_CHECK_ARG(M > 0 && d % 8 == 0 && lds % 8 == 0 && ldd % 8 == 0, "scatter_rows: d %% 8");
    ROW_LAUNCH(scatter_rows_kernel, M, (const bf16*)src, lds, row_map, (bf16*)dst, ldd, M, d);
    return BR_OK;
}

int br_gather_rows(const void* src, int64_t lds, const int32_t* idx, void* dst, int64_t ldd, int M, int d, void* stream) {
    BR_CHECK_ARG(M > 0 && d % 8 == 0 && lds % 8 == 0 && ldd % 8 == 0, "gather_rows: d %% 8");
    ROW_LAUNCH(gather_rows_kernel, M, (const bf16*)src, lds, idx, (bf16*)dst, ldd, M, d);
    return BR_OK;
}

}  // extern "C"
