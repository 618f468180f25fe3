// Shared device helpers for the mma.sync attention kernels (prefill/training flash attention, paged decode attention).
#pragma once
#include "br_common.cuh"

namespace attn {

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
    const uint32_t s = br::smem_u32(smem);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(br::smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(br::smem_u32(p)));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int D>
__device__ __forceinline__ uint8_t* tile_ptr(uint8_t* base, int row, int chunk) {
    constexpr int CH = D / 8;
    return base + ((row * CH + (chunk ^ (row & 7))) << 4);
}

// cooperative 64 x D tile load (rows row0.. of a [*, ld] matrix, clamped to nrows_valid -> zero fill)
template <int D, int NT = 128, int ROWS = 64>
__device__ __forceinline__ void load_tile(uint8_t* s, const bf16* g, long long ld, int row0, int row_limit, int tid) {
    constexpr int CH = D / 8;
#pragma unroll
    for (int i = 0; i < (ROWS * CH + NT - 1) / NT; ++i) {
        const int c = tid + i * NT;
        if ((ROWS * CH) % NT != 0 && c >= ROWS * CH) break;
        const int r = c / CH, ch = c % CH;
        const bool ok = (row0 + r) < row_limit;
        const bf16* src = g + (long long)(ok ? row0 + r : 0) * ld + ch * 8;
        cp_async16(tile_ptr<D>(s, r, ch), src, ok);
    }
}


}  // namespace attn
