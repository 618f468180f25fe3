// Fused next-token sampler for the rollout: temperature -> top-k -> top-p -> draw, plus EOS / pad bookkeeping, one CTA
// per row, no host round trip.  Replaces the HF logits warpers + softmax + torch.multinomial / argmax and the
// unfinished_sequences bookkeeping (HF generation/utils.py:1214-1223, 2762-2797; grpo_trainer.py:384-391).
//
// top-k threshold: exact radix select over the order-preserving uint32 image of the fp32 logits (3 histogram passes
// of 11/11/10 bits, 151 936 logits stay L2-resident), ties at the threshold are all kept like HF's
// `scores < topk(scores)[..., -1]`.  top-p follows TopPLogitsWarper (ascending cumulative sum, drop while
// cum <= 1 - p, always keep the best).  The draw is inverse-CDF over the kept tokens in ascending token id with a
// caller-supplied uniform per (step, row): torch.multinomial's Philox consumption cannot be reproduced outside
// torch, so parity is defined on supplied uniforms (SURVEY.md §7 "Sampling parity").
#include "br_common.cuh"
#include "../../include/bioreason_b200.h"

namespace {

constexpr int MAXC = 1024;

__device__ __forceinline__ uint32_t fkey(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// find bin b (descending scan) such that count(bins > b) < k <= count(bins >= b); returns b and updates k_rem
__device__ int find_bin(const int* hist, int nbins, int& k_rem, int* s_tmp) {
    // executed by warp 0; nbins multiple of 32
    const int lane = threadIdx.x & 31;
    const int per = nbins / 32;
    const int hi = nbins - 1 - lane * per;                      // lane 0 owns the top chunk
    int sum = 0;
    for (int i = 0; i < per; ++i) sum += hist[hi - i];
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    const int excl = incl - sum;
    const bool mine = (excl < k_rem) && (incl >= k_rem);
    if (mine) {
        int acc = excl, b = hi;
        for (int i = 0; i < per; ++i) {
            b = hi - i;
            if (acc + hist[b] >= k_rem) break;
            acc += hist[b];
        }
        s_tmp[0] = b; s_tmp[1] = k_rem - acc;
    }
    __syncwarp();
    int b = s_tmp[0];
    k_rem = s_tmp[1];
    return b;
}

// Fast selection of a SUPERSET of the k largest values (the common case; exactness comes from the consumer, which ranks the superset):
// one shared-memory histogram over the distance to the maximum, bin = floor((max - v) * 32) -- monotone in v, so "all elements in bins
// <= b*" (b* = the bin holding the k-th largest) contains the k largest and every tie of the k-th.  The exact 3-pass radix select over
// the raw key bits (below) puts ~all logits of a chunk into a handful of exponent bins in its first pass (serialised shared-memory
// atomics) and needs three histogram rounds; this needs one, with the counts spread over ~100 bins.  Values further than 64 below the
// maximum are not counted (they can only matter when k exceeds everything closer, which falls back to the radix select).
constexpr int FBINS = 2048;
constexpr float FBIN_SCALE = 32.f;
__device__ __forceinline__ int fast_bin(float mx, float v) {
    const float d = (mx - v) * FBIN_SCALE;                       // NaN (mx = v = -inf) -> 0, +inf -> saturates
    return d >= (float)(FBINS - 1) ? FBINS - 1 : __float2int_rd(d);
}
// warp 0: smallest bin b with count(bins <= b) >= k -> out[0] = b (or -1), out[1] = count(bins <= b)
__device__ void find_bin_asc(const int* hist, int k, int* out) {
    const int lane = threadIdx.x & 31;
    int acc = 0, found = -1, cnt = 0;
    for (int base = 0; base < FBINS - 32; base += 32) {          // the last bin (saturated values) is never counted
        const int c = hist[base + lane];
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (acc + total >= k) {
            const unsigned m = __ballot_sync(0xffffffffu, acc + incl >= k);
            const int l = __ffs(m) - 1;
            found = base + l; cnt = acc + __shfl_sync(0xffffffffu, incl, l);
            break;
        }
        acc += total;
    }
    if (lane == 0) { out[0] = found; out[1] = cnt; }
}
__device__ __forceinline__ float block_max(float v, float* s_red) {   // all threads get the maximum; s_red: 32 floats
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    float m = lane < nw ? s_red[lane] : -INFINITY;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    __syncthreads();
    return m;
}

// Stage 1 (many CTAs): each CTA owns a 4096-logit chunk of one row, keeps it in registers, finds the chunk's exact
// top_k-th value by radix select on shared-memory histograms and emits every element >= that value (value, token id) --
// a superset of the row's global top-k.  Stage 2 (sampler_kernel, one CTA per row) then works on <= chunks*CAND_CAP
// candidates instead of 151 936 logits: the sampler drops from ~200 us to ~20 us per decode step.
constexpr int CHUNK = 4096, CAND_CAP = 64;

__global__ void __launch_bounds__(256) sampler_partial_kernel(const float* __restrict__ logits, long long ld, int V, int top_k,
                                                              float* __restrict__ cand_val, int* __restrict__ cand_idx, int n_chunks) {
    __shared__ int hist[2048];
    __shared__ int s_tmp[4];
    __shared__ int s_count;
    const int chunk = blockIdx.x, row = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    br::launch_dependents();
    br::grid_dep_wait();
    const float* x = logits + (long long)row * ld;
    const int base = chunk * CHUNK;
    float v[16]; uint32_t key[16];
    int n_valid = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int idx = base + i * 256 + tid;
        const bool ok = idx < V;
        v[i] = ok ? __ldcg(x + idx) : -INFINITY;          // L2 loads throughout: PDL-chained kernels keep no coherent L1
        key[i] = ok ? fkey(v[i]) : 0u;
        n_valid += ok;
    }
    const int n_here = min(CHUNK, V - base);
    const int k = min(top_k, n_here);
    float* cv = cand_val + ((long long)row * n_chunks + chunk) * CAND_CAP;
    int* ci = cand_idx + ((long long)row * n_chunks + chunk) * CAND_CAP;
    // ---- fast path: one histogram over the distance to the chunk maximum (see fast_bin)
    {
        __shared__ float s_red[32];
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < 16; ++i) mx = fmaxf(mx, v[i]);
        mx = block_max(mx, s_red);
        for (int i = tid; i < FBINS; i += 256) hist[i] = 0;
        if (tid == 0) s_count = 0;
        __syncthreads();
        int bin[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            bin[i] = (base + i * 256 + tid < V) ? fast_bin(mx, v[i]) : FBINS - 1;
            if (bin[i] < FBINS - 1) atomicAdd(&hist[bin[i]], 1);
        }
        __syncthreads();
        if (warp == 0) find_bin_asc(hist, k, s_tmp);
        __syncthreads();
        const int bstar = s_tmp[0], cnt = s_tmp[1];
        if (mx > -INFINITY && bstar >= 0 && cnt <= CAND_CAP) {                  // uniform across the CTA
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (bin[i] <= bstar) {
                    const int s = atomicAdd(&s_count, 1);
                    cv[s] = v[i]; ci[s] = base + i * 256 + tid;
                }
            }
            for (int s = cnt + tid; s < CAND_CAP; s += 256) { cv[s] = -INFINITY; ci[s] = 0x7fffffff; }
            return;
        }
        __syncthreads();
    }
    // ---- exact path (degenerate chunks: > CAND_CAP values within 1/32 of the k-th, or nothing finite)
    uint32_t prefix = 0; int k_rem = k;
    for (int pass = 0; pass < 3; ++pass) {
        const int shift = pass == 0 ? 21 : (pass == 1 ? 10 : 0);
        const int nb = pass == 2 ? 1024 : 2048;
        for (int i = tid; i < 2048; i += 256) hist[i] = 0;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int idx = base + i * 256 + tid;
            if (idx < V) {
                bool in;
                if (pass == 0) in = true; else if (pass == 1) in = (key[i] >> 21) == prefix; else in = (key[i] >> 10) == prefix;
                if (in) atomicAdd(&hist[(key[i] >> shift) & (nb - 1)], 1);
            }
        }
        __syncthreads();
        if (warp == 0) {
            int kr = k_rem;
            int b = find_bin(hist, nb, kr, s_tmp);
            if (lane == 0) { s_tmp[2] = b; s_tmp[3] = kr; }
        }
        __syncthreads();
        const int b = s_tmp[2];
        k_rem = s_tmp[3];
        prefix = pass == 0 ? (uint32_t)b : (pass == 1 ? ((prefix << 11) | (uint32_t)b) : ((prefix << 10) | (uint32_t)b));
        __syncthreads();
    }
    // values strictly above the k-th (fewer than k) always fit; ties OF the k-th fill the remaining slots in token-id order, so the
    // emitted set does not depend on the order in which threads reach an atomic
    if (tid == 0) s_count = 0;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int idx = base + i * 256 + tid;
        if (idx < V && key[i] > prefix) {
            const int s = atomicAdd(&s_count, 1);
            if (s < CAND_CAP) { cv[s] = v[i]; ci[s] = idx; }
        }
    }
    __syncthreads();
    int filled = min(s_count, CAND_CAP);
    __shared__ int s_w[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        if (filled >= CAND_CAP) break;                                 // uniform across the CTA
        const int idx = base + i * 256 + tid;
        const bool tie = idx < V && key[i] == prefix;
        const unsigned m = __ballot_sync(0xffffffffu, tie);
        if (lane == 0) s_w[warp] = __popc(m);
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) { const int c = s_w[w]; total += c; if (w < warp) before += c; }
        const int slot = filled + before + __popc(m & ((1u << lane) - 1u));
        if (tie && slot < CAND_CAP) { cv[slot] = v[i]; ci[slot] = idx; }
        filled = min(CAND_CAP, filled + total);
        __syncthreads();
    }
    for (int s = filled + tid; s < CAND_CAP; s += 256) { cv[s] = -INFINITY; ci[s] = 0x7fffffff; }
}

// Stage 2 / single-stage sampler.  cand_idx == nullptr: x is the full logits row (index = position).
__global__ void __launch_bounds__(1024) sampler_kernel(const float* __restrict__ logits, long long ld, int V, const int* __restrict__ cand_idx_all, float temperature, int top_k,
                                                       float top_p, int do_sample, const float* __restrict__ uniforms,
                                                       const int* __restrict__ step_ptr, int R, int max_steps, long long eos_id,
                                                       long long pad_id, int* __restrict__ finished, long long* __restrict__ tokens,
                                                       long long* __restrict__ next_ids) {
    __shared__ int hist[2048];
    __shared__ int s_tmp[4];
    __shared__ float c_val[MAXC];
    __shared__ int c_idx[MAXC];
    __shared__ float o_val[MAXC];
    __shared__ int o_idx[MAXC];
    __shared__ int s_count;
    __shared__ float r_val[32];
    __shared__ int r_idx[32];

    const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    br::launch_dependents();
    br::grid_dep_wait();
    const float* x = logits + (long long)row * ld;
    const int* xi = cand_idx_all ? cand_idx_all + (long long)row * ld : nullptr;
    auto IDX = [&](int i) { return xi ? __ldcg(xi + i) : i; };
    const int step = step_ptr ? __ldcg(step_ptr) : 0;
    long long choice;

    if (!do_sample) {
        float bv = -INFINITY; int bi = 0x7fffffff;
        for (int i = tid; i < V; i += blockDim.x) { float v = __ldcg(x + i); const int id = IDX(i); if (v > bv || (v == bv && id < bi)) { bv = v; bi = id; } }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, bv, o); int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { r_val[warp] = bv; r_idx[warp] = bi; }
        __syncthreads();
        if (warp == 0) {
            bv = lane < (blockDim.x >> 5) ? r_val[lane] : -INFINITY; bi = lane < (blockDim.x >> 5) ? r_idx[lane] : 0x7fffffff;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                float ov = __shfl_xor_sync(0xffffffffu, bv, o); int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (lane == 0) s_tmp[2] = bi;
        }
        __syncthreads();
        choice = s_tmp[2];
    } else {
      bool fast_done = false;
      if (xi != nullptr && V <= 4 * (int)blockDim.x) {
        // ---- fast path over the stage-1 candidates (<= 4 per thread): superset by distance-to-maximum bins, exact ranking below
        float v[4]; int id[4]; int bin[4];
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = tid + i * blockDim.x;
            const bool ok = idx < V;
            v[i] = ok ? __ldcg(x + idx) : -INFINITY; id[i] = ok ? __ldcg(xi + idx) : 0x7fffffff;
            mx = fmaxf(mx, v[i]);
        }
        mx = block_max(mx, r_val);
        for (int i = tid; i < FBINS; i += blockDim.x) hist[i] = 0;
        if (tid == 0) s_count = 0;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            bin[i] = v[i] > -INFINITY ? fast_bin(mx, v[i]) : FBINS - 1;
            if (bin[i] < FBINS - 1) atomicAdd(&hist[bin[i]], 1);
        }
        __syncthreads();
        if (warp == 0) find_bin_asc(hist, top_k, s_tmp);
        __syncthreads();
        const int bstar = s_tmp[0], cnt = s_tmp[1];
        __syncthreads();
        if (mx > -INFINITY && bstar >= 0 && cnt <= MAXC) {                       // uniform across the CTA
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (bin[i] <= bstar) { const int s = atomicAdd(&s_count, 1); c_val[s] = v[i]; c_idx[s] = id[i]; }
            }
            fast_done = true;
        }
        __syncthreads();
      }
      if (!fast_done) {
        // ---- exact k-th largest key by 3-pass radix select
        uint32_t prefix = 0; int k_rem = top_k;
        for (int pass = 0; pass < 3; ++pass) {
            const int shift = pass == 0 ? 21 : (pass == 1 ? 10 : 0);
            const int nb = pass == 2 ? 1024 : 2048;
            for (int i = tid; i < 2048; i += blockDim.x) hist[i] = 0;
            __syncthreads();
            for (int i = tid; i < V; i += blockDim.x) {
                const uint32_t k = fkey(__ldcg(x + i));
                bool in;
                if (pass == 0) in = true; else if (pass == 1) in = (k >> 21) == prefix; else in = (k >> 10) == prefix;
                if (in) atomicAdd(&hist[(k >> shift) & (nb - 1)], 1);
            }
            __syncthreads();
            if (warp == 0) {
                int kr = k_rem;
                int b = find_bin(hist, nb, kr, s_tmp);
                if (lane == 0) { s_tmp[2] = b; s_tmp[3] = kr; }
            }
            __syncthreads();
            const int b = s_tmp[2];
            k_rem = s_tmp[3];
            prefix = pass == 0 ? (uint32_t)b : (pass == 1 ? ((prefix << 11) | (uint32_t)b) : ((prefix << 10) | (uint32_t)b));
            __syncthreads();
        }
        const uint32_t thr = prefix;
        if (tid == 0) s_count = 0;
        __syncthreads();
        for (int i = tid; i < V; i += blockDim.x) {
            const float v = __ldcg(x + i);
            if (fkey(v) >= thr && v > -INFINITY) {
                const int s = atomicAdd(&s_count, 1);
                if (s < MAXC) { c_val[s] = v; c_idx[s] = IDX(i); }
            }
        }
        __syncthreads();
      }
        const int c_all = min(s_count, MAXC);
        // ---- rank sort: descending value, ties by ascending index
        if (tid < c_all) {
            const float v = c_val[tid]; const int id = c_idx[tid];
            int rank = 0;
            for (int j = 0; j < c_all; ++j) rank += (c_val[j] > v) || (c_val[j] == v && c_idx[j] < id);
            o_val[rank] = v; o_idx[rank] = id;
        }
        __syncthreads();
        if (tid == 0) {
            // top-k with HF's tie rule (`scores < topk(scores)[..., -1]` are dropped: every value equal to the k-th stays); the
            // collected set may be a superset (fast path) or exactly that set (radix path)
            int c = min(top_k, c_all);
            if (c > 0) { const float kth = o_val[c - 1]; while (c < c_all && o_val[c] == kth) ++c; }
            // softmax over the kept-by-top-k set at temperature T (fp32), descending order
            const float inv_t = 1.f / temperature;
            const float mx = o_val[0] * inv_t;
            float tot = 0.f;
            for (int j = 0; j < c; ++j) { c_val[j] = __expf(o_val[j] * inv_t - mx); tot += c_val[j]; }
            // top-p: ascending cumulative sum; drop while cum <= 1 - p; the best token always stays
            int keep = c;
            if (top_p < 1.f) {
                float cum = 0.f;
                const float lim = 1.f - top_p;
                for (int j = c - 1; j >= 1; --j) {
                    cum += c_val[j] / tot;
                    if (cum <= lim) keep = j; else break;
                }
            }
            float ktot = 0.f;
            for (int j = 0; j < keep; ++j) ktot += c_val[j];
            // inverse CDF in ascending token id
            const float u = uniforms[(long long)step * R + row];
            const float target = u * ktot;
            // selection by repeatedly taking the smallest remaining id (keep is ~20)
            float acc = 0.f; int chosen = o_idx[0]; int last_id = -1;
            for (int n = 0; n < keep; ++n) {
                int best = -1;
                for (int j = 0; j < keep; ++j) if (o_idx[j] > last_id && (best < 0 || o_idx[j] < o_idx[best])) best = j;
                acc += c_val[best]; last_id = o_idx[best]; chosen = o_idx[best];
                if (acc > target) break;
            }
            s_tmp[2] = chosen;
        }
        __syncthreads();
        choice = s_tmp[2];
    }
    if (tid == 0) {
        const int fin = finished ? __ldcg(finished + row) : 0;
        long long tok = fin ? pad_id : choice;                         // finished rows emit pad (HF :2796-2797)
        if (tokens && step < max_steps) tokens[(long long)row * max_steps + step] = tok;
        if (next_ids) next_ids[row] = tok;
        if (finished && !fin && eos_id >= 0 && tok == eos_id) finished[row] = 1;
    }
}

__global__ void advance_kernel(int* step, int* cur_len, int R) {
    const int i = threadIdx.x;
    // NO early launch_dependents() here: this kernel is the token-step boundary.  Later kernels of the chain (the fused decode
    // attention) read cur_len BEFORE their own dependency wait; keeping the implicit trigger at completion guarantees that nothing
    // of step N+1 starts before cur_len of step N is final (and, transitively, before every kernel of step N has completed).
    br::grid_dep_wait();
    if (i < R) atomicAdd(cur_len + i, 1);
    if (i == 0 && step) atomicAdd(step, 1);
}

}  // namespace

extern "C" {

int br_sample_next(const float* logits, int64_t ld, int R, int V, float temperature, int top_k, float top_p, int do_sample,
                   const float* uniforms, const int32_t* step, int max_steps, int64_t eos_id, int64_t pad_id, int32_t* finished,
                   int64_t* tokens, int64_t* next_ids, void* stream) {
    BR_CHECK_ARG(R > 0 && V > 0, "sample_next: empty");
    if (do_sample) {
        BR_CHECK_ARG(temperature > 0.f && top_k >= 1 && top_k <= MAXC && top_p > 0.f && uniforms, "sample_next: need T > 0, 1 <= top_k <= %d, top_p > 0 and a uniforms buffer", MAXC);
    }
    sampler_kernel<<<R, 1024, 0, (cudaStream_t)stream>>>(logits, ld, V, nullptr, temperature, top_k, top_p, do_sample, uniforms, step, R, max_steps,
                                                        (long long)eos_id, (long long)pad_id, finished, (long long*)tokens, (long long*)next_ids);
    BR_CHECK_LAUNCH();
    return BR_OK;
}

int64_t br_sample_workspace_bytes(int R, int V) {
    const int n_chunks = (V + CHUNK - 1) / CHUNK;
    return (int64_t)R * n_chunks * CAND_CAP * (sizeof(float) + sizeof(int));
}

/* two-stage variant for large vocabularies (same semantics as br_sample_next) */
int br_sample_next_2stage(const float* logits, int64_t ld, int R, int V, float temperature, int top_k, float top_p, int do_sample,
                          const float* uniforms, const int32_t* step, int max_steps, int64_t eos_id, int64_t pad_id, int32_t* finished,
                          int64_t* tokens, int64_t* next_ids, void* workspace, void* stream) {
    BR_CHECK_ARG(R > 0 && V > 0 && workspace, "sample_next_2stage: empty / no workspace");
    const int k = do_sample ? top_k : 1;
    BR_CHECK_ARG(k >= 1 && k <= CAND_CAP / 2, "sample_next_2stage: top_k must be in [1, %d]", CAND_CAP / 2);
    if (do_sample) BR_CHECK_ARG(temperature > 0.f && top_p > 0.f && uniforms, "sample_next_2stage: need T > 0, top_p > 0 and a uniforms buffer");
    const int n_chunks = (V + CHUNK - 1) / CHUNK;
    float* cv = (float*)workspace;
    int* ci = (int*)(cv + (int64_t)R * n_chunks * CAND_CAP);
    cudaStream_t st = (cudaStream_t)stream;
    BR_CHECK_CUDA(br_launch_pdl(sampler_partial_kernel, dim3(n_chunks, R), dim3(256), 0, st, logits, (long long)ld, V, k, cv, ci, n_chunks));
    const int n_cand = n_chunks * CAND_CAP;
    BR_CHECK_CUDA(br_launch_pdl(sampler_kernel, dim3(R), dim3(1024), 0, st, (const float*)cv, (long long)n_cand, n_cand, (const int*)ci, temperature,
                                top_k, top_p, do_sample, uniforms, step, R, max_steps, (long long)eos_id, (long long)pad_id, finished,
                                (long long*)tokens, (long long*)next_ids));
    return BR_OK;
}

int br_decode_advance(int32_t* step, int32_t* cur_len, int R, void* stream) {
    BR_CHECK_ARG(R > 0 && R <= 1024, "decode_advance: R in [1, 1024]");
    BR_CHECK_CUDA(br_launch_pdl(advance_kernel, dim3(1), dim3(1024), 0, (cudaStream_t)stream, step, cur_len, R));
    return BR_OK;
}

}  // extern "C"
