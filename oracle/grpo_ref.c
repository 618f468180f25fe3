/* Plain-C restatement of the GRPO advantage + loss arithmetic -- TEST INFRASTRUCTURE ONLY.
 *
 * Follows /root/reference/bioreason/trainer/grpo_trainer.py:
 *   :682-692  group-relative advantages (sum over reward funcs, group mean, UNBIASED std, +1e-4)
 *   :786-812  clipped-ratio loss, k3 KL to the reference policy, masked row mean -> batch mean,
 *             `mean_kl` and `clip_ratio` metrics.
 * Also the analytic d(loss)/d(logp) that torch autograd would produce for that expression
 * (old = logp.detach() when mu == 1, :786), which is what the CUDA kernel emits.
 * Built by oracle/Makefile into oracle/_build/libgrpo_ref.so; loaded with ctypes by tests only.
 */
#include <math.h>
#include <stddef.h>

void oracle_group_advantages(const float* rewards_per_func, int rows, int n_funcs, int G, float* adv) {
    for (int g0 = 0; g0 < rows; g0 += G) {
        float r[1024];
        float mean = 0.f;
        for (int i = 0; i < G; ++i) {
            float s = 0.f;
            for (int f = 0; f < n_funcs; ++f) s += rewards_per_func[(size_t)(g0 + i) * n_funcs + f];
            r[i] = s; mean += s;
        }
        mean /= (float)G;
        float var = 0.f;
        for (int i = 0; i < G; ++i) var += (r[i] - mean) * (r[i] - mean);
        float sd = sqrtf(var / (float)(G - 1));            /* torch.std default: unbiased */
        for (int i = 0; i < G; ++i) adv[g0 + i] = (r[i] - mean) / (sd + 1e-4f);
    }
}

/* old_lp / ref_lp may be NULL (mu == 1 / beta == 0). out3 = {loss, mean_kl, clip_ratio}. */
void oracle_grpo_loss(const float* lp, const float* old_lp, const float* ref_lp, const float* adv,
                      const int* mask, int B, int C, float beta, float eps_lo, float eps_hi,
                      float* out3, float* dlp) {
    double loss = 0.0, kl_acc = 0.0, clip_num = 0.0, mask_tot = 0.0;
    for (int b = 0; b < B; ++b) {
        double cnt = 0.0;
        for (int t = 0; t < C; ++t) cnt += mask[(size_t)b * C + t];
        double row_l = 0.0, row_kl = 0.0;
        for (int t = 0; t < C; ++t) {
            size_t i = (size_t)b * C + t;
            float o = old_lp ? old_lp[i] : lp[i];
            float c1 = expf(lp[i] - o);
            float c2 = fminf(fmaxf(c1, 1.f - eps_lo), 1.f + eps_hi);
            float l1 = c1 * adv[b], l2 = c2 * adv[b];
            float l = -fminf(l1, l2);
            /* d(-min(l1,l2))/dlp: l1 is chosen when l1 <= l2 (torch.min ties -> equal grads split;
               with c1 == c2 the clamp passes gradient, so the total is c1*adv either way). */
            float g;
            if (l1 < l2) g = -c1 * adv[b];
            else if (l1 > l2) g = (c1 > 1.f - eps_lo && c1 < 1.f + eps_hi) ? -c1 * adv[b] : 0.f;
            else g = (c1 >= 1.f - eps_lo && c1 <= 1.f + eps_hi) ? -c1 * adv[b] : -0.5f * c1 * adv[b];
            float kl = 0.f;
            if (beta > 0.f && ref_lp) {
                float d = ref_lp[i] - lp[i];
                kl = expf(d) - d - 1.f;
                l += beta * kl;
                g += beta * (1.f - expf(d));
            }
            float m = (float)mask[i];
            row_l += (double)(l * m); row_kl += (double)(kl * m);
            clip_num += (double)((l1 < l2 ? 1.f : 0.f) * m);
            if (dlp) dlp[i] = (cnt > 0.0) ? (float)(g * m / (cnt * B)) : 0.f;
        }
        mask_tot += cnt;
        loss += row_l / cnt; kl_acc += row_kl / cnt;
    }
    out3[0] = (float)(loss / B);
    out3[1] = (float)(kl_acc / B);
    out3[2] = (float)(clip_num / mask_tot);
}
