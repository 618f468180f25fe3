// tcgen05 / TMEM / TMA GEMM for sm_100a:  D[M,N] = alpha * (A[M,K] . B[N,K]^T  (+ A2[M,K2] . B2[N,K2]^T)) (+bias)(+residual)
//
// Replaces every nn.Linear / lm_head matmul the reference reaches through HF (SURVEY.md §2.3 K1,K2,K5,K6,K12):
// both operands are K-major (nn.Linear weight layout [out, in]), bf16 in, fp32 accumulate in TMEM.
//
// Structure (one persistent CTA per SM, 192 threads):
//   warp 0      TMA producer   : cp.async.bulk.tensor 2-D, 128B-swizzled 128x64 (A) and BNx64 (B) tiles, NSTAGE ring
//   warp 1      MMA issuer     : one lane issues tcgen05.mma.cta_group::1.kind::f16 (UMMA 128 x BN x 16), accumulators
//                                double-buffered in TMEM (2 x BN fp32 columns) so the epilogue of tile i overlaps the
//                                main loop of tile i+1; tcgen05.commit releases smem stages / publishes accumulators
//   warps 2..5  epilogue       : tcgen05.ld 32x32b.x32 (thread <-> accumulator row), fused epilogue, vectorised stores
// Epilogue modes: plain (+bias, +residual, gated-SiLU on interleaved column pairs, fp32/bf16 out, row scatter),
// online log-sum-exp partials + target-logit gather (lm_head; logits never reach HBM), and softmax-gradient tiles.
#include "br_common.cuh"
#include "../../include/bioreason_b200.h"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int NTHREADS = 192;

enum { MODE_STD = 0, MODE_LSE = 1, MODE_DLOGITS = 2 };

struct GemmParams {
    int M, N, K, K2;
    int ldd;                 // elements
    void* D;
    const void* bias; int bias_f32;
    const bf16* residual; long long ldr;
    float alpha;
    int act;                 // 1: gated SiLU over column blocks of 16 = 8 gate | 8 up -> 8 outputs
    int out_f32;
    const int* row_map;
    bf16* aux; long long ld_aux;   // act==1: raw (pre-activation) accumulator pairs, bf16 [M, N]
    // MODE_LSE / MODE_DLOGITS
    const int* target;       // [M] class index per row (or <0)
    float* pmax; float* psum; float* tgt_logit;   // [M, n_tiles_n], [M, n_tiles_n], [M]
    const float* lse; const float* gscale;        // [M]
    int n_tiles_m, n_tiles_n;
    int group_m;             // m-blocks per raster group (see tile_coords)
};

template <int BN>
struct SmemLayout {
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int NSTAGE = (BN == 256) ? 4 : 6;
    static constexpr int TILE_BYTES = NSTAGE * STAGE_BYTES;
    static constexpr int TOTAL = TILE_BYTES + 256 + 1024;   // + barriers + alignment slack
};

__device__ __forceinline__ void tile_coords(int tile, int ntm, int ntn, int GROUP_M, int& mb, int& nb) {
    // grouped rasterisation: GROUP_M m-blocks x all n-blocks per group.  The group's A panel (GROUP_M x 128 x K, sized by the host to
    // ~40 MB) stays L2-resident while every B panel streams past it once, so B is re-read from DRAM once per GROUP (with the fixed
    // GROUP_M = 16 of round 1 the [18880 x 2560] x [19456 x 2560]^T gate/up GEMM re-read B 9 times: 1.08 GB of DRAM reads for 196 MB
    // of operands).
    int per_group = GROUP_M * ntn;
    int g = tile / per_group;
    int first_m = g * GROUP_M;
    int gsize = min(GROUP_M, ntm - first_m);
    int r = tile - g * per_group;
    mb = first_m + (r % gsize);
    nb = r / gsize;
}

template <int BN, int MODE>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc5_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                const GemmParams p) {
    using L = SmemLayout<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::TILE_BYTES);
    uint64_t* empty_bar = full_bar + L::NSTAGE;
    uint64_t* tfull_bar = empty_bar + L::NSTAGE;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_tiles = p.n_tiles_m * p.n_tiles_n;
    const int kb1 = (p.K + BK - 1) / BK;
    const int kb2 = (p.K2 + BK - 1) / BK;
    const int num_kb = kb1 + kb2;

    if (warp == 0 && lane == 0) {
        br::tma_prefetch_desc(&tmA);
        br::tma_prefetch_desc(&tmB);
        if (kb2) { br::tma_prefetch_desc(&tmA2); br::tma_prefetch_desc(&tmB2); }
        for (int s = 0; s < L::NSTAGE; ++s) { br::mbar_init(&full_bar[s], 1); br::mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { br::mbar_init(&tfull_bar[s], 1); br::mbar_init(&tempty_bar[s], 4); }
        br::mbar_fence_init();
    }
    if (warp == 1) {
        br::tmem_alloc(tmem_slot, 2 * BN);
        br::tmem_relinquish();
    }
    br::tc_fence_before();
    __syncthreads();
    br::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                int mb, nb; tile_coords(tile, p.n_tiles_m, p.n_tiles_n, p.group_m, mb, nb);
                for (int kb = 0; kb < num_kb; ++kb) {
                    br::mbar_wait(&empty_bar[s], ph ^ 1);
                    uint8_t* sa = smem + s * L::STAGE_BYTES;
                    uint8_t* sb = sa + L::A_BYTES;
                    br::mbar_expect_tx(&full_bar[s], L::STAGE_BYTES);
                    if (kb < kb1) {
                        br::tma_load_2d(sa, &tmA, &full_bar[s], kb * BK, mb * BM);
                        br::tma_load_2d(sb, &tmB, &full_bar[s], kb * BK, nb * BN);
                    } else {
                        br::tma_load_2d(sa, &tmA2, &full_bar[s], (kb - kb1) * BK, mb * BM);
                        br::tma_load_2d(sb, &tmB2, &full_bar[s], (kb - kb1) * BK, nb * BN);
                    }
                    if (++s == L::NSTAGE) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = br::make_idesc_bf16(BM, BN);
            int s = 0; uint32_t ph = 0; int as = 0; uint32_t aph = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                br::mbar_wait(&tempty_bar[as], aph ^ 1);
                br::tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    br::mbar_wait(&full_bar[s], ph);
                    br::tc_fence_after();
                    const uint32_t sa = br::smem_u32(smem + s * L::STAGE_BYTES);
                    const uint64_t adesc = br::make_sw128_kmajor_desc(sa);
                    const uint64_t bdesc = br::make_sw128_kmajor_desc(sa + L::A_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        // advance 16 bf16 = 32 B along K inside the 128 B swizzle span: +2 in the (addr >> 4) field
                        br::tc_mma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                    }
                    br::tc_commit(&empty_bar[s]);          // smem stage reusable once these MMAs retire
                    if (++s == L::NSTAGE) { s = 0; ph ^= 1; }
                }
                br::tc_commit(&tfull_bar[as]);             // accumulator ready for the epilogue warps
                if (++as == 2) { as = 0; aph ^= 1; }
            }
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int lane_grp = warp & 3;                     // TMEM lane group this warp may access
        int as = 0; uint32_t aph = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            int mb, nb; tile_coords(tile, p.n_tiles_m, p.n_tiles_n, p.group_m, mb, nb);
            br::mbar_wait(&tfull_bar[as], aph);
            br::tc_fence_after();
            const int row = mb * BM + lane_grp * 32 + lane;
            const bool row_ok = row < p.M;
            const uint32_t taddr = tmem_base + as * BN + ((uint32_t)(lane_grp * 32) << 16);
            const int n0 = nb * BN;

            if constexpr (MODE == MODE_STD) {
                long long orow = row;
                if (p.row_map && row_ok) orow = p.row_map[row];
                const bool store_ok = row_ok && orow >= 0;
#pragma unroll 1
                for (int c = 0; c < BN; c += 32) {
                    if (n0 + c >= p.N) break;              // warp-uniform
                    uint32_t r[32];
                    __syncwarp();
                    br::tmem_ld_32x32(taddr + c, r);
                    br::tmem_ld_wait();
                    if (!store_ok) continue;
                    float v[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * p.alpha;
                    const int col = n0 + c;
                    const int ncols = min(32, p.N - col);  // multiple of 8 (N % 8 == 0)
                    if (p.bias) {
                        if (p.bias_f32) {
                            const float* b = reinterpret_cast<const float*>(p.bias) + col;
#pragma unroll
                            for (int i = 0; i < 32; ++i) if (i < ncols) v[i] += __ldg(b + i);
                        } else {
                            const bf16* b = reinterpret_cast<const bf16*>(p.bias) + col;
#pragma unroll
                            for (int i = 0; i < 32; ++i) if (i < ncols) v[i] += __bfloat162float(b[i]);
                        }
                    }
                    if (p.act == 1) {
                        if (p.aux) {
                            uint4* a = reinterpret_cast<uint4*>(p.aux + orow * p.ld_aux + col);
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                if (q * 8 < ncols)
                                    a[q] = make_uint4(br::pack_bf16(v[q * 8 + 0], v[q * 8 + 1]), br::pack_bf16(v[q * 8 + 2], v[q * 8 + 3]),
                                                      br::pack_bf16(v[q * 8 + 4], v[q * 8 + 5]), br::pack_bf16(v[q * 8 + 6], v[q * 8 + 7]));
                        }
                        float o[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            // columns come in blocks of 16 = 8 gate | 8 up (packing.py); HF computes act_fn(gate) in bf16
                            // then multiplies: round at the same places
                            const int gi = (i >> 3) * 16 + (i & 7);
                            float g = __bfloat162float(__float2bfloat16(v[gi])), u = __bfloat162float(__float2bfloat16(v[gi + 8]));
                            float sg = __bfloat162float(__float2bfloat16(g / (1.f + __expf(-g))));
                            o[i] = sg * u;
                        }
                        bf16* d = reinterpret_cast<bf16*>(p.D) + orow * (long long)p.ldd + col / 2;
                        uint4* d4 = reinterpret_cast<uint4*>(d);
#pragma unroll
                        for (int q = 0; q < 2; ++q)
                            if (q * 16 < ncols)
                                d4[q] = make_uint4(br::pack_bf16(o[q * 8 + 0], o[q * 8 + 1]), br::pack_bf16(o[q * 8 + 2], o[q * 8 + 3]),
                                                   br::pack_bf16(o[q * 8 + 4], o[q * 8 + 5]), br::pack_bf16(o[q * 8 + 6], o[q * 8 + 7]));
                        continue;
                    }
                    if (p.residual) {
                        const uint4* rp = reinterpret_cast<const uint4*>(p.residual + orow * p.ldr + col);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            if (q * 8 < ncols) {
                                uint4 rr = __ldg(rp + q);
                                float2 a = br::unpack_bf16(rr.x), b = br::unpack_bf16(rr.y), c2 = br::unpack_bf16(rr.z), d2 = br::unpack_bf16(rr.w);
                                // nn.Linear output is rounded to bf16 before the residual add in HF
                                v[q * 8 + 0] = __bfloat162float(__float2bfloat16(v[q * 8 + 0])) + a.x;
                                v[q * 8 + 1] = __bfloat162float(__float2bfloat16(v[q * 8 + 1])) + a.y;
                                v[q * 8 + 2] = __bfloat162float(__float2bfloat16(v[q * 8 + 2])) + b.x;
                                v[q * 8 + 3] = __bfloat162float(__float2bfloat16(v[q * 8 + 3])) + b.y;
                                v[q * 8 + 4] = __bfloat162float(__float2bfloat16(v[q * 8 + 4])) + c2.x;
                                v[q * 8 + 5] = __bfloat162float(__float2bfloat16(v[q * 8 + 5])) + c2.y;
                                v[q * 8 + 6] = __bfloat162float(__float2bfloat16(v[q * 8 + 6])) + d2.x;
                                v[q * 8 + 7] = __bfloat162float(__float2bfloat16(v[q * 8 + 7])) + d2.y;
                            }
                        }
                    }
                    if (p.out_f32) {
                        float4* d4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.D) + orow * (long long)p.ldd + col);
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            if (q * 4 < ncols) d4[q] = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
                    } else {
                        uint4* d4 = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.D) + orow * (long long)p.ldd + col);
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (q * 8 < ncols)
                                d4[q] = make_uint4(br::pack_bf16(v[q * 8 + 0], v[q * 8 + 1]), br::pack_bf16(v[q * 8 + 2], v[q * 8 + 3]),
                                                   br::pack_bf16(v[q * 8 + 4], v[q * 8 + 5]), br::pack_bf16(v[q * 8 + 6], v[q * 8 + 7]));
                    }
                }
            } else if constexpr (MODE == MODE_LSE) {
                // per-row online max / sum-exp over this tile's columns + target-logit pick
                float mx = -INFINITY, sm = 0.f;
                const int tgt = row_ok ? p.target[row] : -1;
#pragma unroll 1
                for (int c = 0; c < BN; c += 32) {
                    if (n0 + c >= p.N) break;
                    uint32_t r[32];
                    __syncwarp();
                    br::tmem_ld_32x32(taddr + c, r);
                    br::tmem_ld_wait();
                    const int col = n0 + c;
                    const int ncols = min(32, p.N - col);
                    float cm = -INFINITY;
#pragma unroll
                    for (int i = 0; i < 32; ++i) if (i < ncols) cm = fmaxf(cm, __uint_as_float(r[i]) * p.alpha);
                    const float nm = fmaxf(mx, cm);
                    float acc = 0.f;
#pragma unroll
                    for (int i = 0; i < 32; ++i) if (i < ncols) acc += __expf(__uint_as_float(r[i]) * p.alpha - nm);
                    sm = sm * __expf(mx - nm) + acc;
                    mx = nm;
                    if (tgt >= col && tgt < col + ncols) {
                        float t = 0.f;
#pragma unroll
                        for (int i = 0; i < 32; ++i) if (col + i == tgt) t = __uint_as_float(r[i]) * p.alpha;
                        p.tgt_logit[row] = t;
                    }
                }
                if (row_ok) {
                    p.pmax[(long long)row * p.n_tiles_n + nb] = mx;
                    p.psum[(long long)row * p.n_tiles_n + nb] = sm;
                }
            } else {
                // dlogits[m, n] = gscale[m] * (onehot(target[m])[n] - exp(logit - lse[m]))   (bf16 out)
                const int tgt = row_ok ? p.target[row] : -1;
                const float lse = row_ok ? p.lse[row] : 0.f;
                const float gs = row_ok ? p.gscale[row] : 0.f;
#pragma unroll 1
                for (int c = 0; c < BN; c += 32) {
                    if (n0 + c >= p.N) break;
                    uint32_t r[32];
                    __syncwarp();
                    br::tmem_ld_32x32(taddr + c, r);
                    br::tmem_ld_wait();
                    if (!row_ok) continue;
                    const int col = n0 + c;
                    const int ncols = min(32, p.N - col);
                    float v[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        float pr = __expf(__uint_as_float(r[i]) * p.alpha - lse);
                        v[i] = gs * (((col + i) == tgt ? 1.f : 0.f) - pr);
                    }
                    uint4* d4 = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.D) + (long long)row * p.ldd + col);
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (q * 8 < ncols)
                            d4[q] = make_uint4(br::pack_bf16(v[q * 8 + 0], v[q * 8 + 1]), br::pack_bf16(v[q * 8 + 2], v[q * 8 + 3]),
                                               br::pack_bf16(v[q * 8 + 4], v[q * 8 + 5]), br::pack_bf16(v[q * 8 + 6], v[q * 8 + 7]));
                }
            }
            br::tc_fence_before();
            __syncwarp();
            if (lane == 0) br::mbar_arrive(&tempty_bar[as]);
            if (++as == 2) { as = 0; aph ^= 1; }
        }
    }

    br::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        br::tc_fence_after();
        br::tmem_dealloc(tmem_base, 2 * BN);
    }
}

// lse[m] = log sum_t psum[m,t] * exp(pmax[m,t] - gmax) + gmax ; logp[m] = tgt_logit[m] - lse[m]
__global__ void lse_combine_kernel(const float* __restrict__ pmax, const float* __restrict__ psum, const float* __restrict__ tgt_logit,
                                   const int* __restrict__ target, int M, int nt, float* __restrict__ lse, float* __restrict__ logp) {
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    int lane = threadIdx.x & 31;
    float mx = -INFINITY;
    for (int t = lane; t < nt; t += 32) mx = fmaxf(mx, pmax[(long long)row * nt + t]);
    mx = br::warp_max(mx);
    float s = 0.f;
    for (int t = lane; t < nt; t += 32) s += psum[(long long)row * nt + t] * __expf(pmax[(long long)row * nt + t] - mx);
    s = br::warp_sum(s);
    if (lane == 0) {
        float l = logf(s) + mx;
        if (lse) lse[row] = l;
        if (logp) logp[row] = (target[row] >= 0) ? tgt_logit[row] - l : 0.f;
    }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(f);
    }
    return fn;
}

template <int BN, int MODE>
int launch(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& a2, const CUtensorMap& b2, const GemmParams& p, cudaStream_t st) {
    using L = SmemLayout<BN>;
    auto kern = gemm_tc5_kernel<BN, MODE>;
    static bool attr_set = false;
    if (!attr_set) {
        BR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
        attr_set = true;
    }
    int tiles = p.n_tiles_m * p.n_tiles_n;
    int grid = tiles < br_num_sms() ? tiles : br_num_sms();
    kern<<<grid, NTHREADS, L::TOTAL, st>>>(a, b, a2, b2, p);
    BR_CHECK_LAUNCH();
    return BR_OK;
}

int run_gemm(int mode, const void* A, int64_t lda, const void* B, int64_t ldb, int M, int N, int K, const void* A2, int64_t lda2,
             const void* B2, int64_t ldb2, int K2, GemmParams& p, cudaStream_t st) {
    BR_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
    BR_CHECK_ARG(N % 8 == 0 && K % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0, "gemm: N, K, lda, ldb must be multiples of 8");
    BR_CHECK_ARG(((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0), "gemm: operands must be 16-byte aligned");
    const int BN = (N <= 128 || (long long)((M + 127) / 128) * ((N + 255) / 256) < br_num_sms()) ? 128 : 256;
    p.M = M; p.N = N; p.K = K; p.K2 = (A2 && B2) ? K2 : 0;
    p.n_tiles_m = (M + BM - 1) / BM;
    p.n_tiles_n = (N + BN - 1) / BN;
    {   // A panel of one raster group ~ 40 MB of the 126 MB L2 (the concurrently streaming B panels and the outputs need the rest)
        const long long a_block = (long long)BM * (K + p.K2) * 2;
        long long g = (40ll << 20) / (a_block > 0 ? a_block : 1);
        if (g < 8) g = 8;
        if (g > p.n_tiles_m) g = p.n_tiles_m;
        p.group_m = (int)g;
    }
    CUtensorMap ta, tb, ta2, tb2;
    int rc;
    if ((rc = br_make_tmap_2d_bf16(&ta, A, M, K, lda, BM))) return rc;
    if ((rc = br_make_tmap_2d_bf16(&tb, B, N, K, ldb, BN))) return rc;
    if (p.K2) {
        BR_CHECK_ARG(K2 % 8 == 0 && lda2 % 8 == 0 && ldb2 % 8 == 0, "gemm: K2, lda2, ldb2 must be multiples of 8");
        if ((rc = br_make_tmap_2d_bf16(&ta2, A2, M, K2, lda2, BM))) return rc;
        if ((rc = br_make_tmap_2d_bf16(&tb2, B2, N, K2, ldb2, BN))) return rc;
    } else { ta2 = ta; tb2 = tb; }
#define BR_LAUNCH(bn, md) return launch<bn, md>(ta, tb, ta2, tb2, p, st)
    if (BN == 128) {
        if (mode == MODE_STD) BR_LAUNCH(128, MODE_STD);
        if (mode == MODE_LSE) BR_LAUNCH(128, MODE_LSE);
        BR_LAUNCH(128, MODE_DLOGITS);
    } else {
        if (mode == MODE_STD) BR_LAUNCH(256, MODE_STD);
        if (mode == MODE_LSE) BR_LAUNCH(256, MODE_LSE);
        BR_LAUNCH(256, MODE_DLOGITS);
    }
#undef BR_LAUNCH
}

}  // namespace

int br_make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems, uint32_t box_rows) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) { br_set_error("cuTensorMapEncodeTiled not available from the driver"); return BR_ERR_CUDA; }
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstr[1] = {ld_elems * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        br_set_error("cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu ld=%llu box_rows=%u base=%p", (int)r, (unsigned long long)rows,
                     (unsigned long long)cols, (unsigned long long)ld_elems, box_rows, base);
        return BR_ERR_CUDA;
    }
    return BR_OK;
}

extern "C" {

int br_gemm_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, void* D, int64_t ldd, int M, int N, int K,
                 const br_gemm_epilogue* e, void* stream) {
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.D = D; p.ldd = (int)ldd; p.alpha = 1.f;
    const void *A2 = nullptr, *B2 = nullptr; int64_t lda2 = 0, ldb2 = 0; int K2 = 0;
    if (e) {
        p.bias = e->bias; p.bias_f32 = e->bias_dtype == BR_F32;
        p.residual = reinterpret_cast<const bf16*>(e->residual); p.ldr = e->ldr;
        p.alpha = e->alpha; p.act = e->act; p.out_f32 = e->out_dtype == BR_F32;
        p.row_map = e->row_map; p.aux = reinterpret_cast<bf16*>(e->aux_out); p.ld_aux = e->ld_aux;
        A2 = e->A2; B2 = e->B2; lda2 = e->lda2; ldb2 = e->ldb2; K2 = e->K2;
        BR_CHECK_ARG(!(p.act == 1 && (p.out_f32 || p.residual)), "gemm: gated-SiLU epilogue writes bf16 without residual");
        BR_CHECK_ARG(!(p.act == 1 && N % 16 != 0), "gemm: gated-SiLU epilogue needs N %% 16 == 0");
    }
    BR_CHECK_ARG(ldd % 8 == 0 && (uintptr_t)D % 16 == 0, "gemm: D must be 16-byte aligned with ldd %% 8 == 0");
    return run_gemm(MODE_STD, A, lda, B, ldb, M, N, K, A2, lda2, B2, ldb2, K2, p, (cudaStream_t)stream);
}

int64_t br_lmhead_workspace_bytes(int M, int V) {
    int nt = (V + 127) / 128;   // worst case (BN = 128)
    return (int64_t)M * nt * 2 * sizeof(float) + (int64_t)M * sizeof(float);
}

int br_lmhead_logprob_fwd(const void* H, int64_t ldh, const void* W, int64_t ldw, const int32_t* target, int M, int V, int K, float scale,
                          float* logp, float* lse, void* workspace, void* stream) {
    GemmParams p;
    memset(&p, 0, sizeof(p));
    int nt_max = (V + 127) / 128;
    p.alpha = scale; p.target = target;
    p.pmax = reinterpret_cast<float*>(workspace);
    p.psum = p.pmax + (int64_t)M * nt_max;
    p.tgt_logit = p.psum + (int64_t)M * nt_max;
    cudaStream_t st = (cudaStream_t)stream;
    BR_CHECK_CUDA(cudaMemsetAsync(p.tgt_logit, 0, (size_t)M * sizeof(float), st));
    int rc = run_gemm(MODE_LSE, H, ldh, W, ldw, M, V, K, nullptr, 0, nullptr, 0, 0, p, st);
    if (rc) return rc;
    const int wpb = 8;
    lse_combine_kernel<<<(M + wpb - 1) / wpb, wpb * 32, 0, st>>>(p.pmax, p.psum, p.tgt_logit, target, M, p.n_tiles_n, lse, logp);
    BR_CHECK_LAUNCH();
    return BR_OK;
}

int br_lmhead_dlogits(const void* H, int64_t ldh, const void* W, int64_t ldw, const int32_t* target, const float* lse, const float* gscale,
                      int M, int V, int K, float scale, void* dlogits, int64_t ldd, void* stream) {
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.alpha = scale; p.target = target; p.lse = lse; p.gscale = gscale; p.D = dlogits; p.ldd = (int)ldd;
    BR_CHECK_ARG(ldd % 8 == 0, "lmhead_dlogits: ldd %% 8");
    return run_gemm(MODE_DLOGITS, H, ldh, W, ldw, M, V, K, nullptr, 0, nullptr, 0, 0, p, (cudaStream_t)stream);
}

}  // extern "C"
