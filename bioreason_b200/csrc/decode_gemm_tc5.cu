// Decode-time weight-streaming GEMM on tcgen05 (swap-AB, stream-K, persistent):  out[R, N] = X[R, K] . W[N, K]^T, R <= 32.
//
// Every decode step reads every weight byte once (8 GB / step for Qwen3-4B), so the kernel is HBM-bound and is
// built around bytes in flight, not FLOPs:
//   * swap-AB: the weight matrix is the M operand (128 output features per UMMA, M=128), the R live rows of X are the N operand
//     (N = 16 or 32; TMA zero-fills the rows beyond R), accumulator [128 features x N] fp32 in TMEM;
//   * one persistent CTA per SM; a TMA producer warp keeps a 6-stage ring of 128x64 weight tiles (16 KB each, 128B-swizzled)
//     in flight (108 KB of shared memory, so this kernel and its PDL successor co-reside on an SM; the successor fills its ring
//     BEFORE it waits for this kernel -- weights are constant during a rollout); an MMA warp issues tcgen05.mma, 4 epilogue
//     warps drain TMEM.  (Pulling more of the chunk into L2 ahead of time was measured SLOWER: more bytes in flight only add
//     queueing delay to the small latency-critical messages -- partial tiles, counters, activations.);
//   * stream-K: the (feature tile, k block) units of the whole layer are cut into equal contiguous chunks, one per CTA, so
//     small-N layers (o_proj / down_proj: 20 feature tiles) still load all 148 SMs evenly.  A tile finished by several CTAs
//     is reduced deterministically: every contributor writes its fp32 partial tile to its own scratch slot, the last arriver
//     (arrival counter) sums the slots in ascending CTA order and applies the epilogue -- no floating-point atomics.
// Epilogues: bf16 store, +residual, SwiGLU over (8 gate | 8 up) feature blocks, fp32 logits.
// PDL: the kernel is launched with programmatic stream serialization.  Kernels chained this way are NOT separated by the
// usual launch-boundary L1 invalidation, so every load of data another kernel of the chain rewrites (residual, statistics,
// scratch) goes through L2 (ld.global.cg / TMA), never through L1.
#include "br_common.cuh"
#include "../../include/bioreason_b200.h"

namespace {

constexpr int BM = 128, BK = 64, NTHREADS = 192;

struct SkParams {
    int R, N, K;
    int tiles_n, KB, units, chunk;     // units = tiles_n * KB, chunk = units per CTA
    int mode;                           // 0 bf16, 1 bf16 + residual, 2 SwiGLU blocks, 3 fp32
    void* out; long long ldo;
    const bf16* res; long long ldr;
    float* scratch;                     // [grid, 2, BNX, 128] fp32 partial tiles (slot 0: CTA's first tile, 1: its last tile)
    int* counters;                      // [tiles_n], zero between launches (self-resetting)
    // folded RMSNorm (decode): out[r, :] *= rsqrt(sum_i sumsq_in[i, r] / K + eps) (the norm weight is pre-multiplied into W's
    // columns); sumsq_out[(tile*4 + warp), r] = sum over that warp's 32 features of out[r, f]^2 (bf16-rounded) -- partials are
    // written, never accumulated with atomics, and summed in a fixed order by the consumer: the rollout is reproducible.
    const float* sumsq_in; int sumsq_in_n; float* sumsq_out; float eps;
    long long* dbg; int dbg_slot;       // optional %globaltimer stamps [slot][cta][8] (profiling aid)
    br::L2Prefetch pf; int pf_on;       // L2 staging of a later GEMM's weights (see br_common.cuh)
    int* gate_counter; const int* gate_epoch; int gate_base, gate_per_step, gate_wait, gate_signal;   // stream gate (see br_stream_gate)
    int w_evict_first;                  // weight tiles are read once per token step: mark them evict-first in L2 so the small
                                        // latency-critical buffers (activations, partial tiles, statistics, tables) stay resident
};

__device__ __forceinline__ float rbf(float x) { return __bfloat162float(__float2bfloat16(x)); }

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

// PARK: tensor memory as a second-level weight buffer.  The CTA uses 2 * BNX of the 256 TMEM columns it may take (two CTAs share an SM);
// the rest holds NPARK weight tiles in the A-operand layout (tcgen05.cp shared -> tensor memory, consumed by the TS form of the MMA), all
// of them filled BEFORE the dependency on the previous kernel resolves -- on top of the shared-memory ring.
template <int BNX, bool PARK>
struct SL {
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BNX * BK * 2;
    static constexpr int STAGE = A_BYTES + B_BYTES;
#ifndef BR_SK_NSTAGE
#define BR_SK_NSTAGE 6
#endif
    static constexpr int NSTAGE = PARK ? 5 : BR_SK_NSTAGE;   // 6 stages = 108 KB: two CTAs (this kernel + its PDL successor) fit one SM
    static constexpr int TMEM_COLS = PARK ? 256 : (2 * BNX < 32 ? 32 : 2 * BNX);
    static constexpr int NPARK = PARK ? (256 - 2 * BNX) / (BK / 2) : 0;      // a 128 x 64 bf16 tile = 32 columns
    static constexpr int TILE_BYTES = NSTAGE * STAGE;
    static constexpr int XP_BYTES = NPARK * B_BYTES;          // activation tiles of the parked weight tiles (loaded after the dependency wait)
    static constexpr int TOTAL = TILE_BYTES + XP_BYTES + 1024 + 1024;   // + barriers / flags / per-row rstd + alignment slack
};

// residual values of feature f for all live rows, issued as independent L2 loads (one round trip instead of R dependent ones);
// called BEFORE the accumulator wait so the latency hides under the weight stream
template <int BNX>
__device__ __forceinline__ void load_residual(const SkParams& p, int f, float (&res)[BNX]) {
    const bool on = p.mode == 1 && f < p.N;
#pragma unroll
    for (int r = 0; r < BNX; ++r) {
        res[r] = 0.f;
        if (on && r < p.R) res[r] = __bfloat162float(__ushort_as_bfloat16(__ldcg(reinterpret_cast<const unsigned short*>(p.res) + (long long)r * p.ldr + f)));
    }
}

// per-feature epilogue: v[r] = sum for row r of feature f
template <int BNX>
__device__ __forceinline__ void apply_epilogue(const SkParams& p, int f, int lane, const float (&v)[BNX], const float (&res)[BNX], const float* s_rs, int part_row) {
    const bool f_ok = f < p.N;
    float rs[BNX];
#pragma unroll
    for (int r = 0; r < BNX; ++r) rs[r] = s_rs[r];
    if (p.mode == 2) {
        // lanes 0-7 / 16-23 hold gate features, 8-15 / 24-31 the matching up features (blocks of 16 features)
#pragma unroll
        for (int r = 0; r < BNX; ++r) {
            const float other = __shfl_down_sync(0xffffffffu, v[r], 8);
            if (r < p.R && f_ok && (lane & 8) == 0) {
                const float g = rbf(v[r] * rs[r]), u = rbf(other * rs[r]);
                const float sg = rbf(g / (1.f + __expf(-g)));
                reinterpret_cast<bf16*>(p.out)[(long long)r * p.ldo + (f >> 4) * 8 + (f & 7)] = __float2bfloat16(sg * u);
            }
        }
        return;
    }
    float sq[BNX];
#pragma unroll
    for (int r = 0; r < BNX; ++r) {
        sq[r] = 0.f;
        if (r < p.R && f_ok) {
            float x = v[r] * rs[r];
            if (p.mode == 3) reinterpret_cast<float*>(p.out)[(long long)r * p.ldo + f] = x;
            else {
                if (p.mode == 1) x = rbf(x) + res[r];
                const bf16 xb = __float2bfloat16(x);
                reinterpret_cast<bf16*>(p.out)[(long long)r * p.ldo + f] = xb;
                sq[r] = __bfloat162float(xb) * __bfloat162float(xb);
            }
        }
    }
    if (p.sumsq_out) {
#pragma unroll
        for (int r = 0; r < BNX; ++r) {
            if (r >= p.R) break;                                   // warp-uniform
            const float t = br::warp_sum(sq[r]);
            if (lane == 0) p.sumsq_out[(long long)part_row * 32 + r] = t;
        }
    }
}


// Per-row rstd of the folded RMSNorm from the producer's partial sums of squares, summed in a FIXED order (reproducible) but
// with the L2 loads spread over all 128 epilogue threads and issued in batches (a serial loop of ~80 dependent L2 round
// trips here used to cost ~30 us per GEMM).  s_part: [4][32] floats of shared scratch.  Ends with the epilogue-group barrier.
__device__ __forceinline__ void compute_row_rstd(const SkParams& p, int et, float* s_rs, float* s_part) {
    const int r = et & 31, q = et >> 5;                       // row, quarter of the partial list
    float acc = 0.f;
    if (p.sumsq_in && r < p.R) {
        const int n = p.sumsq_in_n;
        const int per = (n + 3) >> 2, lo = q * per, hi = min(n, lo + per);
        for (int i = lo; i < hi; i += 32) {                       // 32 independent L2 loads in flight: one round trip for d <= 4096
            float t[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) t[j] = (i + j < hi) ? __ldcg(p.sumsq_in + (long long)(i + j) * 32 + r) : 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) acc += t[j];               // fixed order: reproducible
        }
    }
    s_part[q * 32 + r] = acc;
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (et < 32) {
        float rsv = 1.f;
        if (p.sumsq_in && et < p.R) rsv = rsqrtf((((s_part[et] + s_part[32 + et]) + s_part[64 + et]) + s_part[96 + et]) / (float)p.K + p.eps);
        s_rs[et] = rsv;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
}

__device__ __forceinline__ long long gtime_sk() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define SKSTAMP(k) do { if (p.dbg && (threadIdx.x == 64)) p.dbg[((long long)p.dbg_slot * 160 + blockIdx.x) * 8 + (k)] = gtime_sk(); } while (0)

__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
}

// BNX: UMMA N (rows of X the tensor core sees, zero-filled beyond R); RM: rows the epilogue code is generated for (R <= RM <= BNX).
// The epilogue runs once per CTA per launch -- straight-line, instruction-fetch-bound code -- so the common R <= 8 decode batch gets its
// own half-size instantiation.
template <int BNX, int RM, bool PARK>
__global__ void __launch_bounds__(NTHREADS, 1)
skinny_tc5_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmP,
                  const SkParams p) {
    using L = SL<BNX, PARK>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* xp = smem + L::TILE_BYTES;                        // PARK: [NPARK] activation tiles
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::TILE_BYTES + L::XP_BYTES);
    uint64_t* empty_bar = full_bar + L::NSTAGE;
    uint64_t* tfull_bar = empty_bar + L::NSTAGE;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint64_t* pfull_bar = tempty_bar + 2;                      // PARK: weight tile of a stage landed (parking phase)
    uint64_t* pempty_bar = pfull_bar + L::NSTAGE;              // PARK: stage copied to tensor memory
    uint64_t* xp_bar = pempty_bar + L::NSTAGE;                 // PARK: activation tiles of the parked weight tiles landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xp_bar + 1);
    int* s_flag = reinterpret_cast<int*>(tmem_slot + 1);
    float* s_rs = reinterpret_cast<float*>(s_flag + 1);       // [32] per-row rstd of the folded RMSNorm

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int u_lo = blockIdx.x * p.chunk;
    const int u_hi = min(p.units, u_lo + p.chunk);

    br::launch_dependents();
    SKSTAMP(0);
    if (warp == 0 && lane == 0) {
        br::tma_prefetch_desc(&tmW);
        br::tma_prefetch_desc(&tmX);
        for (int s = 0; s < L::NSTAGE; ++s) { br::mbar_init(&full_bar[s], 1); br::mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { br::mbar_init(&tfull_bar[s], 1); br::mbar_init(&tempty_bar[s], 4); }
        if constexpr (PARK) {
            for (int s = 0; s < L::NSTAGE; ++s) { br::mbar_init(&pfull_bar[s], 1); br::mbar_init(&pempty_bar[s], 1); }
            br::mbar_init(xp_bar, 1);
        }
        br::mbar_fence_init();
    }
    if (warp == 1) {
        br::tmem_alloc(tmem_slot, L::TMEM_COLS);
        br::tmem_relinquish();
    }
    br::tc_fence_before();
    __syncthreads();
    br::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // the first n_park units of the chunk are parked in tensor memory (they are consumed first: the accumulation order is unchanged),
    // the ring holds the units after them
    const int n_units = u_hi - u_lo;
    const int n_park = PARK ? min(L::NPARK, max(0, n_units - L::NSTAGE)) : 0;
    const uint32_t tmem_park = tmem_base + 2 * BNX;

    if (warp == 0) {
        if (lane == 0) {
            // The weights are constant during the rollout: fill the whole ring with weight tiles BEFORE waiting for the
            // previous kernel (PDL), so the HBM stream of this layer overlaps the tail of the previous kernel.
            const int n_pre = min(L::NSTAGE, n_units - n_park);
            const uint64_t pol = br::make_policy_evict_first();
            if (p.gate_counter && p.gate_wait >= 0) {             // start the early loads under the previous GEMM's exchange tail, not under its stream
                const int target = (__ldcg(p.gate_epoch) - p.gate_base) * p.gate_per_step + p.gate_wait;
                for (int it = 0; it < 32; ++it) {                 // bounded: the gate is a timing hint
                    int seen;
                    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(p.gate_counter) : "memory");
                    if (seen >= target) break;
                }
            }
            auto load_w = [&](void* dst, uint64_t* bar, int c0, int c1) {
                if (p.w_evict_first) br::tma_load_2d_hint(dst, &tmW, bar, c0, c1, pol);
                else br::tma_load_2d(dst, &tmW, bar, c0, c1);
            };
            if constexpr (PARK) {
                for (int i = 0; i < n_park; ++i) {                                   // through the ring into tensor memory (MMA thread copies)
                    const int u = u_lo + i, tile = u / p.KB, kb = u - tile * p.KB;
                    const int st = i % L::NSTAGE, use = i / L::NSTAGE;
                    if (use > 0) br::mbar_wait(&pempty_bar[st], (use - 1) & 1);
                    br::mbar_expect_tx(&pfull_bar[st], L::A_BYTES);
                    load_w(smem + st * L::STAGE, &pfull_bar[st], kb * BK, tile * BM);
                }
            }
            for (int i = 0; i < n_pre; ++i) {
                const int u = u_lo + n_park + i, tile = u / p.KB, kb = u - tile * p.KB;
                if constexpr (PARK) {                                                // the stage may still hold a tile on its way to tensor memory
                    const int uses = (n_park - i + L::NSTAGE - 1) / L::NSTAGE;       // parked tiles that went through stage i
                    if (i < n_park && uses > 0) br::mbar_wait(&pempty_bar[i], (uses - 1) & 1);
                }
                br::mbar_expect_tx(&full_bar[i], L::STAGE);
                load_w(smem + i * L::STAGE, &full_bar[i], kb * BK, tile * BM);
            }
            if (p.pf_on) br::l2_prefetch_issue(&tmP, p.pf, blockIdx.x, gridDim.x);     // a LATER GEMM's tiles -> L2 (HBM is otherwise idle here)
            br::grid_dep_wait();
            if constexpr (PARK) {
                if (n_park > 0) {
                    br::mbar_expect_tx(xp_bar, n_park * L::B_BYTES);
                    for (int i = 0; i < n_park; ++i) {
                        const int u = u_lo + i, tile = u / p.KB, kb = u - tile * p.KB;
                        br::tma_load_2d(xp + i * L::B_BYTES, &tmX, xp_bar, kb * BK, 0);
                    }
                }
            }
            for (int i = 0; i < n_pre; ++i) {
                const int u = u_lo + n_park + i, tile = u / p.KB, kb = u - tile * p.KB;
                br::tma_load_2d(smem + i * L::STAGE + L::A_BYTES, &tmX, &full_bar[i], kb * BK, 0);
            }
            int s = n_pre % L::NSTAGE; uint32_t ph = (n_pre == L::NSTAGE) ? 1u : 0u;
            for (int u = u_lo + n_park + n_pre; u < u_hi; ++u) {
                const int tile = u / p.KB, kb = u - tile * p.KB;
                br::mbar_wait(&empty_bar[s], ph ^ 1);
                uint8_t* sa = smem + s * L::STAGE;
                br::mbar_expect_tx(&full_bar[s], L::STAGE);
                load_w(sa, &full_bar[s], kb * BK, tile * BM);
                br::tma_load_2d(sa + L::A_BYTES, &tmX, &full_bar[s], kb * BK, 0);
                if (++s == L::NSTAGE) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = br::make_idesc_bf16(BM, BNX);
            if constexpr (PARK) {
                for (int i = 0; i < n_park; ++i) {                                   // parking phase (before the dependency resolves)
                    const int st = i % L::NSTAGE, use = i / L::NSTAGE;
                    br::mbar_wait(&pfull_bar[st], use & 1);
                    br::tc_fence_after();
                    const uint64_t adesc = br::make_sw128_kmajor_desc(br::smem_u32(smem + st * L::STAGE));
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) br::tc_cp_128x256b(tmem_park + i * (BK / 2) + k * 8, adesc + 2 * k);
                    br::tc_commit(&pempty_bar[st]);
                }
            }
            int s = 0; uint32_t ph = 0; int as = 0; uint32_t aph = 0;
            int u = u_lo;
            bool xp_ready = false;
            while (u < u_hi) {
                const int tile = u / p.KB;
                const int seg_end = min(u_hi, (tile + 1) * p.KB);
                br::mbar_wait(&tempty_bar[as], aph ^ 1);
                br::tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * BNX;
                for (int i = 0; u < seg_end; ++u, ++i) {
                    if constexpr (PARK) {
                        if (u - u_lo < n_park) {                                     // weight tile in tensor memory, activation tile in xp
                            if (!xp_ready) { br::mbar_wait(xp_bar, 0); br::tc_fence_after(); xp_ready = true; }
                            const int j = u - u_lo;
                            const uint64_t bdesc = br::make_sw128_kmajor_desc(br::smem_u32(xp + j * L::B_BYTES));
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k)
                                br::tc_mma_bf16_ts(tmem_d, tmem_park + j * (BK / 2) + k * 8, bdesc + 2 * k, idesc, (i | k) != 0);
                            continue;
                        }
                    }
                    br::mbar_wait(&full_bar[s], ph);
                    if (u == u_hi - 1 && p.gate_counter && p.gate_signal)          // every weight tile of this CTA is on chip
                        asm volatile("red.relaxed.gpu.global.add.s32 [%0], 1;" ::"l"(p.gate_counter) : "memory");
                    br::tc_fence_after();
                    const uint32_t sa = br::smem_u32(smem + s * L::STAGE);
                    const uint64_t adesc = br::make_sw128_kmajor_desc(sa);
                    const uint64_t bdesc = br::make_sw128_kmajor_desc(sa + L::A_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) br::tc_mma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (i | k) != 0);
                    br::tc_commit(&empty_bar[s]);
                    if (++s == L::NSTAGE) { s = 0; ph ^= 1; }
                }
                br::tc_commit(&tfull_bar[as]);
                if (++as == 2) { as = 0; aph ^= 1; }
            }
        }
    } else {
        const int lane_grp = warp & 3;
        const int et = threadIdx.x - 64;                      // 0..127 within the epilogue group
        SKSTAMP(2);                                           // TMEM allocated, barriers initialised
        br::grid_dep_wait();                                  // everything below touches data shared with earlier kernels
        SKSTAMP(1);
        compute_row_rstd(p, et, s_rs, s_rs + 32);
        int as = 0; uint32_t aph = 0;
        int u = u_lo;
        while (u < u_hi) {
            const int tile = u / p.KB;
            const int seg_end = min(u_hi, (tile + 1) * p.KB);
            const bool whole = (u == tile * p.KB) && (seg_end == (tile + 1) * p.KB);
            const int f = tile * BM + lane_grp * 32 + lane;
            const int part_row = tile * 4 + lane_grp;
            float res[RM];
            load_residual<RM>(p, f, res);                        // in flight while the accumulator is still being produced
            br::mbar_wait(&tfull_bar[as], aph);
            if (u == u_lo) SKSTAMP(3);
            br::tc_fence_after();
            const uint32_t taddr = tmem_base + as * BNX + ((uint32_t)(lane_grp * 32) << 16);
            float v[RM];
            if constexpr (RM == 8) {
                uint32_t r[8];
                __syncwarp();
                tmem_ld_32x8(taddr, r);
                br::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
            } else {
#pragma unroll
                for (int c = 0; c < RM; c += 16) {
                    uint32_t r[16];
                    __syncwarp();
                    tmem_ld_32x16(taddr + c, r);
                    br::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[c + i] = __uint_as_float(r[i]);
                }
            }
            br::tc_fence_before();
            __syncwarp();
            if (lane == 0) br::mbar_arrive(&tempty_bar[as]);   // accumulator drained into registers
            if (++as == 2) { as = 0; aph ^= 1; }
            if (whole) {
                apply_epilogue<RM>(p, f, lane, v, res, s_rs, part_row);
            } else {
                // Deterministic stream-K exchange.  A tile that spans several CTAs is finished by the FIRST of them (lowest index):
                // for that CTA the tile is the last segment of its chunk, so it has nothing else left to do, while every other
                // contributor meets the tile at the START of its chunk and publishes early.  Contributors store their fp32 partial
                // tile to their scratch slot and signal with one release-reduction per warp (no CTA barrier, no fence, no returning
                // atomic); the reducer acquires the counter, gathers all partials in ONE batch of independent L2 loads and adds them
                // to its own registers in ascending CTA order -- no floating-point atomics, bit-reproducible.
                const int first_c = (tile * p.KB) / p.chunk, last_c = ((tile + 1) * p.KB - 1) / p.chunk;
                if ((int)blockIdx.x != first_c) {
                    float* mine = p.scratch + ((long long)blockIdx.x * 2 * BNX) * BM + lane_grp * 32 + lane;   // slot 0: the CTA's first tile
#pragma unroll
                    for (int r = 0; r < RM; ++r)
                        if (r < p.R) __stcg(mine + r * BM, v[r]);
                    __syncwarp();
                    if (lane == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.counters + tile) : "memory");
                    if (seg_end == u_hi) SKSTAMP(4);
                } else {
                    if (et == 0) {
                        const unsigned want = 4u * (unsigned)(last_c - first_c);
                        unsigned seen;
                        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.counters + tile) : "memory"); } while (seen < want);
                        p.counters[tile] = 0;                                    // nobody touches it again before the next launch
                    }
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    SKSTAMP(4);
                    for (int c0 = first_c + 1; c0 <= last_c; c0 += 8) {          // 8 contributors x 8 rows of loads in flight
#pragma unroll
                        for (int r0 = 0; r0 < RM; r0 += 8) {
                            if (r0 >= p.R) break;                                // warp-uniform
                            float t[8][8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float* src = p.scratch + ((long long)(c0 + j) * 2 * BNX) * BM + lane_grp * 32 + lane;
#pragma unroll
                                for (int r = 0; r < 8; ++r) t[j][r] = (c0 + j <= last_c && r0 + r < p.R) ? __ldcg(src + (r0 + r) * BM) : 0.f;
                            }
#pragma unroll
                            for (int j = 0; j < 8; ++j)
#pragma unroll
                                for (int r = 0; r < 8; ++r) v[r0 + r] += t[j][r];            // ascending CTA order: deterministic
                        }
                    }
                    SKSTAMP(5);
                    apply_epilogue<RM>(p, f, lane, v, res, s_rs, part_row);
                    SKSTAMP(6);
                }
            }
            u = seg_end;
        }
        SKSTAMP(7);
    }
    br::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        br::tc_fence_after();
        br::tmem_dealloc(tmem_base, L::TMEM_COLS);
    }
}

template <int BNX, int RM, bool PARK>
int launch(const CUtensorMap& tw, const CUtensorMap& tx, const CUtensorMap& tp, const SkParams& p, int grid, cudaStream_t st) {
    using L = SL<BNX, PARK>;
    auto kern = skinny_tc5_kernel<BNX, RM, PARK>;
    static bool done = false;
    if (!done) {
        BR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
        done = true;
    }
    BR_CHECK_CUDA(br_launch_pdl(kern, dim3(grid), dim3(NTHREADS), (size_t)L::TOTAL, st, tw, tx, tp, p));
    return BR_OK;
}


// ================================================================================================================
// Multi-phase persistent variant: up to 4 dependent GEMMs (o_proj -> gate/up -> down_proj -> next layer's qkv, or
// ... -> lm_head) in ONE launch.  Phases are separated by a grid-wide barrier instead of a kernel boundary; the weight
// producer is not gated by the barrier, so while the CTAs synchronise (and while the last tiles of a phase are reduced)
// the ring already fills with the next phase's weights -- the HBM stream does not stop at phase boundaries.  A second
// producer thread loads the activation tiles and is the only one that waits for "phase inputs ready".
// ================================================================================================================
constexpr int CHAIN_MAX = 4;
constexpr int CHAIN_THREADS = 224;           // warp 0: W producer, 1: MMA, 2-5: epilogue, 6: X producer

struct ChainPhase { CUtensorMap tmW; CUtensorMap tmX; SkParams p; };
struct ChainParams { ChainPhase ph[CHAIN_MAX]; int n_phases; int* gbar; long long* dbg; };
__device__ __forceinline__ long long gtime() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define CSTAMP(k) do { if (cp.dbg && et == 0) cp.dbg[(long long)blockIdx.x * 32 + (k)] = gtime(); } while (0)

template <int BNX>
__global__ void __launch_bounds__(CHAIN_THREADS, 1) skinny_chain_kernel(const __grid_constant__ ChainParams cp) {
    using L = SL<BNX, false>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::TILE_BYTES);
    uint64_t* empty_bar = full_bar + L::NSTAGE;
    uint64_t* tfull_bar = empty_bar + L::NSTAGE;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    int* s_flag = reinterpret_cast<int*>(tmem_slot + 1);
    volatile int* s_ready = reinterpret_cast<volatile int*>(s_flag + 1);      // number of grid barriers this CTA has passed
    float* s_rs = reinterpret_cast<float*>(s_flag + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nph = cp.n_phases;

    br::launch_dependents();
    if (warp == 0 && lane == 0) {
        for (int i = 0; i < nph; ++i) { br::tma_prefetch_desc(&cp.ph[i].tmW); br::tma_prefetch_desc(&cp.ph[i].tmX); }
        for (int s = 0; s < L::NSTAGE; ++s) { br::mbar_init(&full_bar[s], 1); br::mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { br::mbar_init(&tfull_bar[s], 1); br::mbar_init(&tempty_bar[s], 4); }
        br::mbar_fence_init();
        *s_ready = 0;
    }
    if (warp == 1) {
        br::tmem_alloc(tmem_slot, 2 * BNX < 32 ? 32 : 2 * BNX);
        br::tmem_relinquish();
    }
    br::tc_fence_before();
    __syncthreads();
    br::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---------------- weight producer: never waits for other kernels or phases (weights are constant) ----------------
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int pi = 0; pi < nph; ++pi) {
                const SkParams& p = cp.ph[pi].p;
                const int u_lo = blockIdx.x * p.chunk, u_hi = min(p.units, u_lo + p.chunk);
                for (int u = u_lo; u < u_hi; ++u) {
                    const int tile = u / p.KB, kb = u - tile * p.KB;
                    br::mbar_wait(&empty_bar[s], ph ^ 1);
                    br::mbar_expect_tx(&full_bar[s], L::STAGE);
                    br::tma_load_2d(smem + s * L::STAGE, &cp.ph[pi].tmW, &full_bar[s], kb * BK, tile * BM);
                    if (++s == L::NSTAGE) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 6) {
        // ---------------- activation producer: gated by "inputs of phase pi are complete" ----------------
        if (lane == 0) {
            br::grid_dep_wait();
            int s = 0; uint32_t ph = 0;
            for (int pi = 0; pi < nph; ++pi) {
                const SkParams& p = cp.ph[pi].p;
                const int u_lo = blockIdx.x * p.chunk, u_hi = min(p.units, u_lo + p.chunk);
                if (u_lo < u_hi) while (*s_ready < pi) __nanosleep(32);
                for (int u = u_lo; u < u_hi; ++u) {
                    const int tile = u / p.KB, kb = u - tile * p.KB;
                    br::mbar_wait(&empty_bar[s], ph ^ 1);
                    br::tma_load_2d(smem + s * L::STAGE + L::A_BYTES, &cp.ph[pi].tmX, &full_bar[s], kb * BK, 0);
                    if (++s == L::NSTAGE) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = br::make_idesc_bf16(BM, BNX);
            int s = 0; uint32_t ph = 0; int as = 0; uint32_t aph = 0;
            for (int pi = 0; pi < nph; ++pi) {
                const SkParams& p = cp.ph[pi].p;
                const int u_lo = blockIdx.x * p.chunk, u_hi = min(p.units, u_lo + p.chunk);
                int u = u_lo;
                while (u < u_hi) {
                    const int tile = u / p.KB;
                    const int seg_end = min(u_hi, (tile + 1) * p.KB);
                    br::mbar_wait(&tempty_bar[as], aph ^ 1);
                    br::tc_fence_after();
                    const uint32_t tmem_d = tmem_base + as * BNX;
                    for (int i = 0; u < seg_end; ++u, ++i) {
                        br::mbar_wait(&full_bar[s], ph);
                        br::tc_fence_after();
                        const uint32_t sa = br::smem_u32(smem + s * L::STAGE);
                        const uint64_t adesc = br::make_sw128_kmajor_desc(sa);
                        const uint64_t bdesc = br::make_sw128_kmajor_desc(sa + L::A_BYTES);
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) br::tc_mma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (i | k) != 0);
                        br::tc_commit(&empty_bar[s]);
                        if (++s == L::NSTAGE) { s = 0; ph ^= 1; }
                    }
                    br::tc_commit(&tfull_bar[as]);
                    if (++as == 2) { as = 0; aph ^= 1; }
                }
            }
        }
    } else {
        // ---------------- epilogue warps 2..5 ----------------
        const int lane_grp = warp & 3;
        const int et = threadIdx.x - 64;
        CSTAMP(0);
        br::grid_dep_wait();
        CSTAMP(1);
        int as = 0; uint32_t aph = 0;
        for (int pi = 0; pi < nph; ++pi) {
            const SkParams& p = cp.ph[pi].p;
            const int u_lo = blockIdx.x * p.chunk, u_hi = min(p.units, u_lo + p.chunk);
            CSTAMP(2 + pi * 6);
            compute_row_rstd(p, et, s_rs, s_rs + 32);      // inputs complete: barrier pi-1 passed
            int u = u_lo;
            while (u < u_hi) {
                const int tile = u / p.KB;
                const int seg_end = min(u_hi, (tile + 1) * p.KB);
                const bool whole = (u == tile * p.KB) && (seg_end == (tile + 1) * p.KB);
                br::mbar_wait(&tfull_bar[as], aph);
                if (u == u_lo) CSTAMP(3 + pi * 6);
                br::tc_fence_after();
                const uint32_t taddr = tmem_base + as * BNX + ((uint32_t)(lane_grp * 32) << 16);
                float v[BNX];
#pragma unroll
                for (int c = 0; c < BNX; c += 16) {
                    uint32_t r[16];
                    __syncwarp();
                    tmem_ld_32x16(taddr + c, r);
                    br::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[c + i] = __uint_as_float(r[i]);
                }
                br::tc_fence_before();
                __syncwarp();
                if (lane == 0) br::mbar_arrive(&tempty_bar[as]);
                if (++as == 2) { as = 0; aph ^= 1; }
                const int f = tile * BM + lane_grp * 32 + lane;
                const int part_row = tile * 4 + lane_grp;
                float res[BNX];
                load_residual<BNX>(p, f, res);
                if (whole) {
                    apply_epilogue<BNX>(p, f, lane, v, res, s_rs, part_row);
                } else {
                    const int first_c = (tile * p.KB) / p.chunk, last_c = ((tile + 1) * p.KB - 1) / p.chunk;
                    const int my_slot = (tile == u_lo / p.KB) ? 0 : 1;
                    float* mine = p.scratch + (((long long)blockIdx.x * 2 + my_slot) * BNX) * BM + lane_grp * 32 + lane;
#pragma unroll
                    for (int r = 0; r < BNX; ++r)
                        if (r < p.R) __stcg(mine + r * BM, v[r]);
                    __threadfence();
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    if (et == 0) *s_flag = (atomicAdd(p.counters + tile, 1) == last_c - first_c);
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    if (*s_flag) {
                        __threadfence();
#pragma unroll
                        for (int r = 0; r < BNX; ++r) v[r] = 0.f;
                        for (int c0 = first_c; c0 <= last_c; c0 += 4) {
                            float t[4][BNX];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int c = c0 + j;
                                const int slot = (tile == (c * p.chunk) / p.KB) ? 0 : 1;
                                const float* src = p.scratch + (((long long)c * 2 + slot) * BNX) * BM + lane_grp * 32 + lane;
#pragma unroll
                                for (int r = 0; r < BNX; ++r) t[j][r] = (c <= last_c && r < p.R) ? __ldcg(src + r * BM) : 0.f;
                            }
#pragma unroll
                            for (int j = 0; j < 4; ++j)
#pragma unroll
                                for (int r = 0; r < BNX; ++r) v[r] += t[j][r];
                        }
                        if (et == 0) p.counters[tile] = 0;
                        apply_epilogue<BNX>(p, f, lane, v, res, s_rs, part_row);
                    }
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                }
                u = seg_end;
            }
            CSTAMP(4 + pi * 6);
            if (pi + 1 < nph) {
                // grid barrier: every CTA's outputs of phase pi are globally visible before anyone loads them as phase pi+1 inputs
                __threadfence();
                asm volatile("bar.sync 1, 128;" ::: "memory");
                CSTAMP(5 + pi * 6);
                if (et == 0) {
                    atomicAdd(cp.gbar, 1);
                    const int target = (pi + 1) * (int)gridDim.x;
                    int seen;
                    do { asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(cp.gbar) : "memory"); } while (seen < target);
                    *s_ready = pi + 1;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                CSTAMP(6 + pi * 6);
            }
        }
        // the last CTA to finish re-zeros the barrier words for the next launch (everyone else has left the barrier code)
        if (et == 0 && nph > 1) {
            if (atomicAdd(cp.gbar + 1, 1) == (int)gridDim.x - 1) { cp.gbar[0] = 0; cp.gbar[1] = 0; __threadfence(); }
        }
    }
    br::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        br::tc_fence_after();
        br::tmem_dealloc(tmem_base, 2 * BNX < 32 ? 32 : 2 * BNX);
    }
}

template <int BNX>
int launch_chain(const ChainParams& cp, int grid, cudaStream_t st) {
    using L = SL<BNX, false>;
    auto kern = skinny_chain_kernel<BNX>;
    static bool done = false;
    if (!done) { BR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL)); done = true; }
    BR_CHECK_CUDA(br_launch_pdl(kern, dim3(grid), dim3(CHAIN_THREADS), (size_t)L::TOTAL, st, cp));
    return BR_OK;
}

}  // namespace

static long long* g_sk_dbg = nullptr;
static int g_sk_dbg_slot = 0;

int br_make_l2_prefetch(const br_l2_prefetch* spec, CUtensorMap* tmap, int* KB, int* units, int* chunk, int* n_chunks, int* a, int* b) {
    BR_CHECK_ARG(spec->N % 16 == 0 && spec->K % 8 == 0 && spec->ldw % 8 == 0 && spec->unit_lo >= 0, "l2_prefetch: bad weight shape");
    const int tiles_n = (spec->N + BM - 1) / BM;
    *KB = (spec->K + BK - 1) / BK; *units = tiles_n * *KB;
    int grid = *units < br_num_sms() ? *units : br_num_sms();
    *chunk = (*units + grid - 1) / grid;
    *n_chunks = (*units + *chunk - 1) / *chunk;
    *a = spec->unit_lo; *b = spec->unit_hi;
    return br_make_tmap_2d_bf16(tmap, spec->W, spec->N, spec->K, spec->ldw, BM);
}

extern "C" {

int64_t br_skinny_scratch_bytes(int max_N) {
    // partial tiles [n_sms, 2, 32, 128] fp32 | grid-barrier word (16 ints) | one arrival counter per 128-feature tile
    return (int64_t)br_num_sms() * 2 * 32 * BM * sizeof(float) + 16 * sizeof(int) + (int64_t)(max_N / BM + 2) * sizeof(int);
}

int br_skinny_gemm(const void* X, int64_t ldx, const void* W, int64_t ldw, void* out, int64_t ldo, int R, int N, int K, int mode,
                   const void* residual, int64_t ldr, void* scratch, void* stream) {
    return br_skinny_gemm_ex(X, ldx, W, ldw, out, ldo, R, N, K, mode, residual, ldr, scratch, nullptr, 0, nullptr, 0.f, stream);
}

int br_skinny_gemm_ex(const void* X, int64_t ldx, const void* W, int64_t ldw, void* out, int64_t ldo, int R, int N, int K, int mode,
                      const void* residual, int64_t ldr, void* scratch, const float* sumsq_in, int sumsq_in_n, float* sumsq_out, float eps,
                      void* stream) {
    return br_skinny_gemm_pf(X, ldx, W, ldw, out, ldo, R, N, K, mode, residual, ldr, scratch, sumsq_in, sumsq_in_n, sumsq_out, eps, nullptr, stream);
}

int br_skinny_grid(int N, int K) {
    const int units = ((N + BM - 1) / BM) * ((K + BK - 1) / BK);
    int grid = units < br_num_sms() ? units : br_num_sms();
    const int chunk = (units + grid - 1) / grid;
    return (units + chunk - 1) / chunk;
}

int br_skinny_gemm_pf(const void* X, int64_t ldx, const void* W, int64_t ldw, void* out, int64_t ldo, int R, int N, int K, int mode,
                      const void* residual, int64_t ldr, void* scratch, const float* sumsq_in, int sumsq_in_n, float* sumsq_out, float eps,
                      const br_l2_prefetch* prefetch, void* stream) {
    return br_skinny_gemm_gated(X, ldx, W, ldw, out, ldo, R, N, K, mode, residual, ldr, scratch, sumsq_in, sumsq_in_n, sumsq_out, eps, prefetch, nullptr, stream);
}

int br_skinny_gemm_gated(const void* X, int64_t ldx, const void* W, int64_t ldw, void* out, int64_t ldo, int R, int N, int K, int mode,
                         const void* residual, int64_t ldr, void* scratch, const float* sumsq_in, int sumsq_in_n, float* sumsq_out, float eps,
                         const br_l2_prefetch* prefetch, const br_stream_gate* gate, void* stream) {
    BR_CHECK_ARG(R >= 1 && R <= 32, "skinny_gemm: R=%d must be in [1, 32]", R);
    BR_CHECK_ARG(N % 16 == 0 && K % 8 == 0 && ldx % 8 == 0 && ldw % 8 == 0, "skinny_gemm: N %% 16, K %% 8, ld %% 8 (N=%d K=%d)", N, K);
    BR_CHECK_ARG(mode >= 0 && mode <= 3 && !(mode == 1 && !residual), "skinny_gemm: bad mode %d", mode);
    BR_CHECK_ARG(scratch != nullptr, "skinny_gemm: scratch (br_skinny_scratch_bytes, zero-initialised once) is required");
    SkParams p;
    p.R = R; p.N = N; p.K = K; p.mode = mode; p.out = out; p.ldo = ldo; p.res = (const bf16*)residual; p.ldr = ldr;
    p.scratch = (float*)scratch; p.counters = (int*)((float*)scratch + (int64_t)br_num_sms() * 2 * 32 * BM) + 16;
    p.sumsq_in = sumsq_in; p.sumsq_in_n = sumsq_in_n; p.sumsq_out = sumsq_out; p.eps = eps;
    BR_CHECK_ARG(!(sumsq_out && mode >= 2), "skinny_gemm: sumsq_out only with bf16 outputs (mode 0/1)");
    p.dbg = g_sk_dbg; p.dbg_slot = g_sk_dbg ? g_sk_dbg_slot++ : 0;
    p.gate_counter = nullptr; p.gate_epoch = nullptr; p.gate_base = p.gate_per_step = p.gate_signal = 0; p.gate_wait = -1;
    if (gate && gate->counter) {
        BR_CHECK_ARG(gate->epoch != nullptr && gate->per_step >= 0, "skinny_gemm: stream gate needs the epoch counter");
        p.gate_counter = gate->counter; p.gate_epoch = gate->epoch; p.gate_base = gate->epoch_base; p.gate_per_step = gate->per_step;
        p.gate_wait = gate->wait_prefix; p.gate_signal = gate->signal;
    }
    { const char* e = getenv("BR_SKINNY_EVICT_FIRST"); p.w_evict_first = e ? atoi(e) : 1; }
    p.tiles_n = (N + BM - 1) / BM; p.KB = (K + BK - 1) / BK; p.units = p.tiles_n * p.KB;
    int grid = p.units < br_num_sms() ? p.units : br_num_sms();
    p.chunk = (p.units + grid - 1) / grid;
    grid = (p.units + p.chunk - 1) / p.chunk;
    const int BNX = R <= 16 ? 16 : 32;
    CUtensorMap tw, tx;
    int rc;
    if ((rc = br_make_tmap_2d_bf16(&tw, W, N, K, ldw, BM))) return rc;
    if ((rc = br_make_tmap_2d_bf16(&tx, X, R, K, ldx, BNX))) return rc;
    CUtensorMap tp = tw;
    p.pf_on = 0;
    if (prefetch && prefetch->W && prefetch->unit_hi > prefetch->unit_lo) {
        if ((rc = br_make_l2_prefetch(prefetch, &tp, &p.pf.KB, &p.pf.units, &p.pf.chunk, &p.pf.n_chunks, &p.pf.a, &p.pf.b))) return rc;
        p.pf_on = 1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    // BR_SKINNY_PARK=1: also buffer weight tiles in tensor memory before the dependency resolves (A/B switch until it is the measured default)
    static const bool park = getenv("BR_SKINNY_PARK") && atoi(getenv("BR_SKINNY_PARK")) != 0;
    if (park) {
        if (R <= 8) return launch<16, 8, true>(tw, tx, tp, p, grid, st);
        return BNX == 16 ? launch<16, 16, true>(tw, tx, tp, p, grid, st) : launch<32, 32, true>(tw, tx, tp, p, grid, st);
    }
    if (R <= 8) return launch<16, 8, false>(tw, tx, tp, p, grid, st);
    return BNX == 16 ? launch<16, 16, false>(tw, tx, tp, p, grid, st) : launch<32, 32, false>(tw, tx, tp, p, grid, st);
}


/* profiling aid: [n_launches, 160, 8] int64 %globaltimer stamps of the next br_skinny_gemm launches (NULL disables):
 * 0 kernel entry, 1 dependency wait passed, 2 prologue done (TMEM allocated, barriers initialised), 3 first accumulator, 4 last partial published, 5 reduction loads done,
 * 6 reducer epilogue done, 7 CTA done */
int br_skinny_debug(long long* buf) { g_sk_dbg = buf; g_sk_dbg_slot = 0; return BR_OK; }

static long long* g_chain_dbg = nullptr;
/* profiling aid: [n_sms, 32] int64 %globaltimer stamps of the next chain launches (NULL disables) */
int br_skinny_chain_debug(long long* buf) { g_chain_dbg = buf; return BR_OK; }

int br_skinny_chain(const br_skinny_phase* phases, int n_phases, int R, float eps, void* scratch, void* stream) {
    BR_CHECK_ARG(n_phases >= 1 && n_phases <= CHAIN_MAX && R >= 1 && R <= 32 && scratch, "skinny_chain: 1..%d phases, R in [1, 32]", CHAIN_MAX);
    ChainParams cp;
    memset(&cp, 0, sizeof(cp));
    cp.n_phases = n_phases;
    cp.dbg = g_chain_dbg;
    const int BNX = R <= 16 ? 16 : 32;
    const int grid = br_num_sms();                         // every phase uses the full grid: the barrier counts gridDim.x arrivals
    float* part = (float*)scratch;
    cp.gbar = (int*)(part + (int64_t)br_num_sms() * 2 * 32 * BM);
    for (int i = 0; i < n_phases; ++i) {
        const br_skinny_phase& h = phases[i];
        BR_CHECK_ARG(h.N % 16 == 0 && h.K % 8 == 0 && h.ldx % 8 == 0 && h.ldw % 8 == 0, "skinny_chain[%d]: N %% 16, K %% 8, ld %% 8", i);
        BR_CHECK_ARG(h.mode >= 0 && h.mode <= 3 && !(h.mode == 1 && !h.residual) && !(h.sumsq_out && h.mode >= 2), "skinny_chain[%d]: bad mode", i);
        SkParams& p = cp.ph[i].p;
        p.R = R; p.N = h.N; p.K = h.K; p.mode = h.mode; p.out = h.out; p.ldo = h.ldo; p.res = (const bf16*)h.residual; p.ldr = h.ldr;
        p.scratch = part; p.counters = cp.gbar + 16;
        p.sumsq_in = h.sumsq_in; p.sumsq_in_n = h.sumsq_in_n; p.sumsq_out = h.sumsq_out; p.eps = eps; p.dbg = nullptr; p.dbg_slot = 0; p.w_evict_first = 0;
        p.tiles_n = (h.N + BM - 1) / BM; p.KB = (h.K + BK - 1) / BK; p.units = p.tiles_n * p.KB;
        p.chunk = (p.units + grid - 1) / grid;
        int rc;
        if ((rc = br_make_tmap_2d_bf16(&cp.ph[i].tmW, h.W, h.N, h.K, h.ldw, BM))) return rc;
        if ((rc = br_make_tmap_2d_bf16(&cp.ph[i].tmX, h.X, R, h.K, h.ldx, BNX))) return rc;
    }
    cudaStream_t st = (cudaStream_t)stream;      // barrier words gbar[0..1] are zero between launches (self-resetting)
    return BNX == 16 ? launch_chain<16>(cp, grid, st) : launch_chain<32>(cp, grid, st);
}

}  // extern "C"
