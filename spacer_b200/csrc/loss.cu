// Fused GRPO loss tail (forward + backward in one pass), K16+K17 of SURVEY.md 2.3:
//   per-token log-prob  lp = logit[target] - logsumexp   (from the lm_head GEMM's LMHEAD epilogue partials;
//                       the [.,V] logits are never materialised)        SG_RLVR_trainer.py:353-366
//   completion mask     positions <= first EOS                           :489-494
//   KL estimator        x = clamp(ref - lp, -10, 10); kl = e^x - x - 1   :551-552
//   loss                -(exp(lp - sg(lp)) * A - beta * kl), per-row masked mean, mean over rows   :640-643
//   backward            coef = dLoss/dlp per token (feeds the DLOGITS epilogue of the recompute GEMM)
// One CTA per completion row; a finalize kernel averages rows deterministically.
#include "common.cuh"
#include "spacer_b200.h"

namespace {

constexpr int LT = 256;

__global__ void __launch_bounds__(LT)
grpo_row_kernel(const float2* __restrict__ lse_part, int n_tiles, const float* __restrict__ tgt_logit,
                const int* __restrict__ comp_ids, int C, int eos_id, const float* __restrict__ ref_lp,
                const float* __restrict__ adv, float beta, int G, float* __restrict__ lp_out,
                float* __restrict__ lse_out, float* __restrict__ coef_out, int* __restrict__ mask_out,
                float* __restrict__ row_loss, float* __restrict__ row_kl, int* __restrict__ row_len) {
  __shared__ float red[32];
  __shared__ int s_eos;
  const int g = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // first EOS
  if (tid == 0) s_eos = C;
  __syncthreads();
  int first = C;
  for (int t = tid; t < C; t += LT)
    if (comp_ids[(long long)g * C + t] == eos_id) { first = t; break; }
  if (first < C) atomicMin(&s_eos, first);
  __syncthreads();
  const int len = min(s_eos + 1, C);   // EOS itself is inside the mask
  // log-probs: one warp per token reduces the per-tile (max, sumexp) partials
  for (int t = warp; t < C; t += LT / 32) {
    const long long r = (long long)g * C + t;
    const float2* pp = lse_part + r * n_tiles;
    float mx = -INFINITY;
    for (int i = lane; i < n_tiles; i += 32) mx = fmaxf(mx, pp[i].x);
    mx = warp_max(mx);
    float s = 0.f;
    for (int i = lane; i < n_tiles; i += 32) {
      const float2 v = pp[i];
      if (v.x > -INFINITY) s += v.y * __expf(v.x - mx);
    }
    s = warp_sum(s);
    if (lane == 0) {
      const float lse = mx + logf(s);
      lse_out[r] = lse;
      lp_out[r] = tgt_logit[r] - lse;
    }
  }
  __syncthreads();
  const float A = adv[g];
  float sl = 0.f, sk = 0.f;
  for (int t = tid; t < C; t += LT) {
    const long long r = (long long)g * C + t;
    const float lp = lp_out[r];
    const float d = ref_lp ? ref_lp[r] - lp : 0.f;
    const float x = fminf(fmaxf(d, -10.f), 10.f);
    const float ex = __expf(x);
    const float kl = ex - x - 1.f;
    const bool in = t < len;
    if (in) {
      sl += -(A - beta * kl);
      sk += kl;
    }
    // d/dlp: ratio term -> -A; kl term (inside the clamp) -> beta * (1 - e^x)
    const float dkl = (d > -10.f && d < 10.f) ? (1.f - ex) : 0.f;
    coef_out[r] = in ? (-A + beta * dkl) / ((float)len * (float)G) : 0.f;
    if (mask_out) mask_out[r] = in ? 1 : 0;
  }
  sl = block_sum(sl, red);
  sk = block_sum(sk, red);
  if (tid == 0) {
    row_loss[g] = sl / (float)len;
    row_kl[g] = sk / (float)len;
    row_len[g] = len;
  }
}

__global__ void grpo_finalize_kernel(const float* __restrict__ row_loss, const float* __restrict__ row_kl, int G,
                                     float* __restrict__ out /*[2]: loss, mean_kl*/) {
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int g = 0; g < G; ++g) { a += row_loss[g]; b += row_kl[g]; }
    out[0] = a / G;
    out[1] = b / G;
  }
}

// lp only (reference-policy scoring pass): lp = tgt - lse from LMHEAD partials
__global__ void logprob_kernel(const float2* __restrict__ lse_part, int n_tiles, const float* __restrict__ tgt_logit,
                               float* __restrict__ lp_out, long long rows) {
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const float2* pp = lse_part + r * n_tiles;
  float mx = -INFINITY;
  for (int i = lane; i < n_tiles; i += 32) mx = fmaxf(mx, pp[i].x);
  mx = warp_max(mx);
  float s = 0.f;
  for (int i = lane; i < n_tiles; i += 32) {
    const float2 v = pp[i];
    if (v.x > -INFINITY) s += v.y * __expf(v.x - mx);
  }
  s = warp_sum(s);
  if (lane == 0) lp_out[r] = tgt_logit[r] - (mx + logf(s));
}

}  // namespace

extern "C" int sb_grpo_loss(const float* lse_part, int n_tiles, const float* tgt_logit, const int* comp_ids, int G,
                            int C, int eos_id, const float* ref_lp, const float* adv, float beta, float* lp_out,
                            float* lse_out, float* coef_out, int* mask_out, float* row_loss, float* row_kl,
                            int* row_len, float* out2, sb_stream_t stream) {
  SB_REQUIRE(lse_part && tgt_logit && comp_ids && adv && lp_out && lse_out && coef_out && row_loss && row_kl &&
                 row_len && out2, "sb_grpo_loss: null pointer");
  SB_REQUIRE(G > 0 && C > 0 && n_tiles > 0, "sb_grpo_loss: bad sizes");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  grpo_row_kernel<<<G, LT, 0, st>>>(reinterpret_cast<const float2*>(lse_part), n_tiles, tgt_logit, comp_ids, C,
                                    eos_id, ref_lp, adv, beta, G, lp_out, lse_out, coef_out, mask_out, row_loss,
                                    row_kl, row_len);
  if (sb_check_launch("sb_grpo_loss")) return 1;
  grpo_finalize_kernel<<<1, 32, 0, st>>>(row_loss, row_kl, G, out2);
  return sb_check_launch("sb_grpo_loss(finalize)");
}

extern "C" int sb_logprob_from_partials(const float* lse_part, int n_tiles, const float* tgt_logit, float* lp_out,
                                        long long rows, sb_stream_t stream) {
  SB_REQUIRE(lse_part && tgt_logit && lp_out && rows > 0 && n_tiles > 0, "sb_logprob_from_partials: bad arguments");
  logprob_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float2*>(lse_part), n_tiles, tgt_logit, lp_out, rows);
  return sb_check_launch("sb_logprob_from_partials");
}
