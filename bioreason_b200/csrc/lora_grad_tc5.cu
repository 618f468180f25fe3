// LoRA weight gradients on tcgen05:  out (+)= big[M, P]^T . small[M, N]   (contraction over the M tokens; fp32 out, N <= 128).
//
// Autograd of the adapters the reference trains with peft (reason.py:362-394): dB = dy^T t and dA = u^T x.
// Both operands are read exactly as the forward/backward left them -- token-major [M, features] -- as MN-MAJOR tcgen05 operands
// (the TMA box [64 tokens x 64 features] is one swizzle atom column; no transposed copies).  One launch covers a whole fused linear:
// the full [P, N] product of e.g. dqkv^T (6144 features) with t_qkv (3r columns) is formed in TMEM and the epilogue writes only the
// block each adapter owns (q rows x its r columns, ...), so 14 launches per decoder layer become 8.
// Split-K over the token dimension fills the 148 SMs for narrow outputs; partial tiles are exchanged through a workspace and summed
// by the split-0 CTA in ascending split order (release/acquire counter, no floating-point atomics): gradients are bit-reproducible.
#include "br_common.cuh"
#include "../../include/bioreason_b200.h"

namespace {

constexpr int BM = 128, BKT = 64, NTHREADS = 192, NSTAGE = 4;
constexpr int A_BYTES = 2 * 64 * 128;          // two [64 tokens x 64 features] blocks
constexpr int B_BLK = 64 * 128;

struct Seg { float* dst; long long ld; int row_lo, row_hi, col_lo, n_cols; };   // rows [row_lo,row_hi) of the product, columns [col_lo, col_lo+n_cols)

struct TnParams {
    int M, P, N, Npad, NBB;                    // NBB = 64-column blocks of `small`
    int tiles, splits, kb_total;
    int mode;                                  // 0: segments (row-major dst[(p - row_lo) * ld + n]); 1: transposed dst[n * ld + p];
                                               // 2: gate/up interleave: product row p = block of 16 = 8 gate | 8 up -> dst rows (p/16)*8 + p%8
    Seg seg[3]; int n_seg;
    float* ws; int* counters;
};

template <int NPAD>
__global__ void __launch_bounds__(NTHREADS, 1)
tn_gemm_tc5_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TnParams p) {
    constexpr int NBB = (NPAD + 63) / 64;
    constexpr int STAGE = A_BYTES + NBB * B_BLK;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + NSTAGE * STAGE);
    uint64_t* empty_bar = full_bar + NSTAGE;
    uint64_t* acc_bar = empty_bar + NSTAGE;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x / p.splits, split = blockIdx.x % p.splits;
    const int kb_per = (p.kb_total + p.splits - 1) / p.splits;
    const int kb_lo = split * kb_per, kb_hi = min(p.kb_total, kb_lo + kb_per);
    const int n_kb = max(0, kb_hi - kb_lo);

    if (warp == 0 && lane == 0) {
        br::tma_prefetch_desc(&tmA); br::tma_prefetch_desc(&tmB);
        for (int s = 0; s < NSTAGE; ++s) { br::mbar_init(&full_bar[s], 1); br::mbar_init(&empty_bar[s], 1); }
        br::mbar_init(acc_bar, 1);
        br::mbar_fence_init();
    }
    if (warp == 1) { br::tmem_alloc(tmem_slot, NPAD <= 32 ? 32 : (NPAD <= 64 ? 64 : 128)); br::tmem_relinquish(); }
    br::tc_fence_before();
    __syncthreads();
    br::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int kb = kb_lo; kb < kb_hi; ++kb) {
                br::mbar_wait(&empty_bar[s], ph ^ 1);
                uint8_t* sa = smem + s * STAGE;
                br::mbar_expect_tx(&full_bar[s], STAGE);
                br::tma_load_2d(sa, &tmA, &full_bar[s], tile * BM, kb * BKT);
                br::tma_load_2d(sa + 64 * 128, &tmA, &full_bar[s], tile * BM + 64, kb * BKT);
#pragma unroll
                for (int nb = 0; nb < NBB; ++nb) br::tma_load_2d(sa + A_BYTES + nb * B_BLK, &tmB, &full_bar[s], nb * 64, kb * BKT);
                if (++s == NSTAGE) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && n_kb > 0) {
            constexpr uint32_t idesc = br::make_idesc_bf16_major(BM, NPAD, 1, 1);      // both operands MN-major
            int s = 0; uint32_t ph = 0;
            for (int i = 0; i < n_kb; ++i) {
                br::mbar_wait(&full_bar[s], ph);
                br::tc_fence_after();
                const uint32_t sa = br::smem_u32(smem + s * STAGE);
#pragma unroll
                for (int kk = 0; kk < BKT / 16; ++kk) {
                    // 16 tokens = 2 groups of 8 rows (SBO 1024 B); 64-feature blocks are 8192 B apart (LBO)
                    const uint64_t ad = br::make_sw128_mnmajor_desc(sa + kk * 2048, 64 * 128, 1024);
                    const uint64_t bd = br::make_sw128_mnmajor_desc(sa + A_BYTES + kk * 2048, B_BLK, 1024);
                    br::tc_mma_bf16(tmem_base, ad, bd, idesc, (i | kk) != 0);
                }
                br::tc_commit(&empty_bar[s]);
                if (++s == NSTAGE) { s = 0; ph ^= 1; }
            }
            br::tc_commit(acc_bar);
        }
    } else {
        const int lane_grp = warp & 3;
        const int row = lane_grp * 32 + lane;
        const int prow = tile * BM + row;                              // row of the product = feature index of `big`
        const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16);
        float v[NPAD];
        if (n_kb > 0) {
            br::mbar_wait(acc_bar, 0);
            br::tc_fence_after();
#pragma unroll
            for (int c = 0; c < NPAD; c += 16) {
                uint32_t r[16];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                      "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                    : "r"(taddr + c) : "memory");
                br::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) v[c + i] = __uint_as_float(r[i]);
            }
        } else {
#pragma unroll
            for (int c = 0; c < NPAD; ++c) v[c] = 0.f;
        }
        if (split != 0) {
            float* mine = p.ws + ((long long)blockIdx.x * NPAD) * BM + row;
#pragma unroll
            for (int c = 0; c < NPAD; ++c) __stcg(mine + c * BM, v[c]);
            __syncwarp();
            if (lane == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.counters + tile) : "memory");
        } else {
            if (p.splits > 1) {
                if (threadIdx.x == 64) {
                    const unsigned want = 4u * (unsigned)(p.splits - 1);
                    unsigned seen;
                    do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.counters + tile) : "memory"); } while (seen < want);
                    p.counters[tile] = 0;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                for (int s2 = 1; s2 < p.splits; ++s2) {                       // ascending split order: deterministic
                    const float* src = p.ws + ((long long)(tile * p.splits + s2) * NPAD) * BM + row;
#pragma unroll
                    for (int c = 0; c < NPAD; ++c) v[c] += __ldcg(src + c * BM);
                }
            }
            if (prow < p.P) {
                if (p.mode == 1) {
                    float* d = p.seg[0].dst + prow;
#pragma unroll
                    for (int c = 0; c < NPAD; ++c)
                        if (c < p.N) d[(long long)c * p.seg[0].ld] += v[c];
                } else {
#pragma unroll
                    for (int sgi = 0; sgi < 3; ++sgi) {
                        if (sgi >= p.n_seg) break;
                        const Seg& sg = p.seg[sgi];
                        int drow;
                        if (p.mode == 2) {                                    // gate/up interleave: seg 0 = gate rows, seg 1 = up rows
                            if (((prow >> 3) & 1) != sgi) continue;
                            drow = (prow >> 4) * 8 + (prow & 7);
                        } else {
                            if (prow < sg.row_lo || prow >= sg.row_hi) continue;
                            drow = prow - sg.row_lo;
                        }
                        float* d = sg.dst + (long long)drow * sg.ld;
#pragma unroll
                        for (int c = 0; c < NPAD; ++c)
                            if (c >= sg.col_lo && c < sg.col_lo + sg.n_cols) d[c - sg.col_lo] += v[c];
                    }
                }
            }
        }
    }
    br::tc_fence_before();
    __syncthreads();
    if (warp == 1) { br::tc_fence_after(); br::tmem_dealloc(tmem_base, NPAD <= 32 ? 32 : (NPAD <= 64 ? 64 : 128)); }
}

template <int NPAD>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const TnParams& p, cudaStream_t st) {
    constexpr int NBB = (NPAD + 63) / 64;
    constexpr int SMEM = NSTAGE * (A_BYTES + NBB * B_BLK) + 256 + 1024;
    auto kern = tn_gemm_tc5_kernel<NPAD>;
    static bool done = false;
    if (!done) { BR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM)); done = true; }
    kern<<<p.tiles * p.splits, NTHREADS, SMEM, st>>>(ta, tb, p);
    BR_CHECK_LAUNCH();
    return BR_OK;
}

}  // namespace

extern "C" {

int64_t br_lora_grad_workspace_bytes(void) {
    // partial tiles [n_sms][128 cols][128 rows] fp32 + one counter per 128-feature tile (zero-initialised once; self-resetting)
    return (int64_t)br_num_sms() * 128 * BM * sizeof(float) + 4096 * sizeof(int);
}

int br_lora_grad_tn(const void* big, int64_t ldb, const void* small, int64_t lds, int M, int P, int N, int mode,
                    const br_lora_grad_seg* segs, int n_seg, void* workspace, void* stream) {
    BR_CHECK_ARG(M > 0 && P > 0 && N >= 8 && N <= 128 && N % 8 == 0, "lora_grad_tn: M=%d P=%d N=%d (N %% 8, <= 128)", M, P, N);
    BR_CHECK_ARG(P % 8 == 0 && ldb % 8 == 0 && lds % 8 == 0 && ((uintptr_t)big % 16 == 0) && ((uintptr_t)small % 16 == 0), "lora_grad_tn: alignment");
    BR_CHECK_ARG(mode >= 0 && mode <= 2 && n_seg >= 1 && n_seg <= 3 && segs && workspace, "lora_grad_tn: bad mode / segments");
    BR_CHECK_ARG((P + BM - 1) / BM <= 4096, "lora_grad_tn: P too large");
    TnParams p;
    memset(&p, 0, sizeof(p));
    p.M = M; p.P = P; p.N = N; p.mode = mode; p.n_seg = n_seg;
    for (int i = 0; i < n_seg; ++i) {
        p.seg[i].dst = segs[i].dst; p.seg[i].ld = segs[i].ld; p.seg[i].row_lo = segs[i].row_lo; p.seg[i].row_hi = segs[i].row_hi;
        p.seg[i].col_lo = segs[i].col_lo; p.seg[i].n_cols = segs[i].n_cols;
        BR_CHECK_ARG(segs[i].dst && segs[i].col_lo >= 0 && segs[i].col_lo + segs[i].n_cols <= N, "lora_grad_tn: segment %d columns outside [0, N)", i);
    }
    p.Npad = N <= 32 ? 32 : (N <= 64 ? 64 : (N <= 96 ? 96 : 128));
    p.tiles = (P + BM - 1) / BM;
    p.kb_total = (M + BKT - 1) / BKT;
    int splits = br_num_sms() / p.tiles;                        // every CTA must be co-resident (the reducer spins on its peers)
    if (splits < 1) splits = 1;
    if (splits > 16) splits = 16;
    if (splits > p.kb_total) splits = p.kb_total;
    if (p.tiles > br_num_sms()) splits = 1;                     // more tiles than SMs: whole-K tiles, no exchange
    p.splits = splits;
    p.ws = (float*)workspace;
    p.counters = (int*)(p.ws + (int64_t)br_num_sms() * 128 * BM);
    CUtensorMap ta, tb;
    int rc;
    // token-major matrices, box = [64 tokens x 64 features]: tokens beyond M and features beyond the row are zero-filled by TMA
    if ((rc = br_make_tmap_2d_bf16(&ta, big, (uint64_t)M, (uint64_t)P, (uint64_t)ldb, BKT))) return rc;
    if ((rc = br_make_tmap_2d_bf16(&tb, small, (uint64_t)M, (uint64_t)N, (uint64_t)lds, BKT))) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    switch (p.Npad) {
        case 32: return launch<32>(ta, tb, p, st);
        case 64: return launch<64>(ta, tb, p, st);
        case 96: return launch<96>(ta, tb, p, st);
        default: return launch<128>(ta, tb, p, st);
    }
}

}  // extern "C"
