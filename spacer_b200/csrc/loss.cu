// Fused GRPO loss tail (forward + backward in one pass), K16+K17 of SURVEY.md 2.3:
//   per-token log-prob  lp = logit[target] - logsumexp   (from the lm_head GEMM's LMHEAD epilogue partials;
//                       the [.,V] logits are never materialised)        SG_RLVR_trainer.py:353-366
//   completion mask     positions <= first EOS                           :489-494
//   KL estimator        x = clamp(ref - lp, -10, 10); kl = e^x - x - 1   :551-552
//   loss                -(exp(lp - sg(lp)) * A - beta * kl), per-row masked mean, mean over rows   :640-643
//   backward            coef = dLoss/dlp per token (feeds the DLOGITS epilogue of the recompute GEMM)
// Two launches: a streaming fold of the partials (the bandwidth part), then the loss kernel whose last-arriving CTAs
// fold the per-CTA sums in a fixed order (deterministic).
#include "common.cuh"
#include "spacer_b200.h"

namespace {


// fold one token's per-tile (max, sumexp) partials: every lane keeps an online (max, sum) over float4 = 2 tiles
SB_DEVICE float fold_lse(const float2* __restrict__ pp, int n_tiles, int lane) {
  float mx = -INFINITY, sm = 0.f;
  if ((reinterpret_cast<uintptr_t>(pp) & 15) != 0) {   // odd n_tiles: rows are only 8-byte aligned
    for (int i = lane; i < n_tiles; i += 32) {
      const float2 v = pp[i];
      if (v.x > -INFINITY) {
        const float nm = fmaxf(mx, v.x);
        sm = sm * __expf(mx - nm) + v.y * __expf(v.x - nm);
        mx = nm;
      }
    }
    const float gm0 = warp_max(mx);
    return gm0 + logf(warp_sum(mx > -INFINITY ? sm * __expf(mx - gm0) : 0.f));
  }
  const float4* p4 = reinterpret_cast<const float4*>(pp);
  const int n4 = n_tiles >> 1;
  constexpr int NB = 10;   // float4 loads in flight per lane (V = 152064 -> 594 tiles -> 297 float4 -> 10 per lane)
  for (int base = 0; base < n4; base += NB * 32) {
    float4 v[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) {             // issue every load of the batch before touching any of them
      const int i = base + j * 32 + lane;
      v[j] = i < n4 ? __ldcs(p4 + i) : make_float4(-INFINITY, 0.f, -INFINITY, 0.f);   // streamed once: evict-first
    }
    float bm = -INFINITY;
#pragma unroll
    for (int j = 0; j < NB; ++j) bm = fmaxf(bm, fmaxf(v[j].x, v[j].z));
    if (bm > -INFINITY) {
      const float nm = fmaxf(mx, bm);
      float acc = sm * __expf(mx - nm);
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        acc += (v[j].x > -INFINITY ? v[j].y * __expf(v[j].x - nm) : 0.f) +
               (v[j].z > -INFINITY ? v[j].w * __expf(v[j].z - nm) : 0.f);
      }
      sm = acc;
      mx = nm;
    }
  }
  if ((n_tiles & 1) && lane == 0) {
    const float2 v = pp[n_tiles - 1];
    if (v.x > -INFINITY) {
      const float nm = fmaxf(mx, v.x);
      sm = sm * __expf(mx - nm) + v.y * __expf(v.x - nm);
      mx = nm;
    }
  }
  const float gm = warp_max(mx);
  const float part = mx > -INFINITY ? sm * __expf(mx - gm) : 0.f;
  return gm + logf(warp_sum(part));
}

// Streaming fold of the lm_head partials: lp[r] = tgt_logit[r] - logsumexp over the n_tiles (max, sumexp) pairs.
// This is the HBM-bound part of the loss (n_tiles x 8 B per token, 4.75 KB at V = 152064): one warp per token, every
// lane issues its 10 x 16 B loads (evict-first) before consuming any, no block-level synchronisation at all.
// (A cp.async.bulk/mbarrier ring per warp was measured at 1.3 TB/s on 4.75 KB rows -- 2.4x slower -- and dropped.)
__global__ void logprob_kernel(const float2* __restrict__ lse_part, int n_tiles, const float* __restrict__ tgt_logit,
                               float* __restrict__ lp_out, float* __restrict__ lse_out, long long rows) {
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const float lse = fold_lse(lse_part + r * n_tiles, n_tiles, lane);
  if (lane == 0) {
    lp_out[r] = tgt_logit[r] - lse;
    if (lse_out) lse_out[r] = lse;
  }
}

int launch_fold(const float* lse_part, int n_tiles, const float* tgt_logit, float* lp_out, float* lse_out,
                long long rows, cudaStream_t st, const char* what) {
  logprob_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(reinterpret_cast<const float2*>(lse_part), n_tiles, tgt_logit,
                                                           lp_out, lse_out, rows);
  return sb_check_launch(what);
}

// The loss itself, one launch.  grid (ceil(C / 256), G): CTA (c, g) owns 256 completion tokens of row g: completion
// mask, KL, loss terms and dLoss/dlogprob from the log-probs; per-CTA partial sums go to `ws`; the last CTA of each
// row (ticket counters in ws, self-resetting) adds the row's sums in a fixed order, the last row folds the scalar
// loss -> deterministic.
constexpr int LT = 256;

__global__ void __launch_bounds__(LT)
grpo_loss_kernel(const int* __restrict__ comp_ids, int C, int eos_id, const float* __restrict__ ref_lp,
                 const float* __restrict__ adv, float beta, int G, const float* __restrict__ lp_in,
                 float* __restrict__ coef_out, int* __restrict__ mask_out,
                 float* __restrict__ row_loss, float* __restrict__ row_kl, int* __restrict__ row_len,
                 float* __restrict__ out2, float* __restrict__ ws) {
  __shared__ float red[32];
  __shared__ int s_eos;
  __shared__ bool s_last;
  const int g = blockIdx.y, chunk = blockIdx.x, n_chunks = gridDim.x;
  const int tid = threadIdx.x;
  // first EOS of the row (every CTA of the row scans the C ids: 4 bytes per token, L2-resident)
  if (tid == 0) s_eos = C;
  __syncthreads();
  int first = C;
  for (int t = tid; t < C; t += LT)
    if (comp_ids[(long long)g * C + t] == eos_id) { first = t; break; }
  if (first < C) atomicMin(&s_eos, first);
  __syncthreads();
  const int len = min(s_eos + 1, C);   // EOS itself is inside the mask
  const int t = chunk * LT + tid;
  float sl = 0.f, sk = 0.f;
  if (t < C) {
    const long long r = (long long)g * C + t;
    const float A = adv[g];
    const float lp = lp_in[r];
    const float d = ref_lp ? ref_lp[r] - lp : 0.f;
    const float x = fminf(fmaxf(d, -10.f), 10.f);
    const float ex = __expf(x);
    const float kl = ex - x - 1.f;
    const bool in = t < len;
    if (in) {
      sl = -(A - beta * kl);
      sk = kl;
    }
    // d/dlp: ratio term -> -A; kl term (inside the clamp) -> beta * (1 - e^x)
    const float dkl = (d > -10.f && d < 10.f) ? (1.f - ex) : 0.f;
    coef_out[r] = in ? (-A + beta * dkl) / ((float)len * (float)G) : 0.f;
    if (mask_out) mask_out[r] = in ? 1 : 0;
  }
  sl = block_sum(sl, red);
  sk = block_sum(sk, red);
  // two-level completion tickets (a single counter would serialise thousands of same-address atomics):
  //   ws[0]: rows finished, ws[4 + g]: chunks of row g finished, then the per-CTA sums [G][n_chunks][2]
  unsigned int* tickets = reinterpret_cast<unsigned int*>(ws);
  float* psum = ws + 4 + ((G + 3) & ~3);
  if (tid == 0) {
    psum[((long long)g * n_chunks + chunk) * 2] = sl;
    psum[((long long)g * n_chunks + chunk) * 2 + 1] = sk;
    __threadfence();
    const unsigned int done = atomicAdd(tickets + 4 + g, 1u);
    s_last = (done == (unsigned int)n_chunks - 1u);
  }
  __syncthreads();
  if (!s_last) return;
  // last CTA of row g: fixed-order sum of the row's chunks
  __threadfence();
  float rl = 0.f, rk = 0.f;
  for (int c = tid; c < n_chunks; c += LT) {
    rl += __ldcg(psum + ((long long)g * n_chunks + c) * 2);
    rk += __ldcg(psum + ((long long)g * n_chunks + c) * 2 + 1);
  }
  rl = block_sum(rl, red);
  rk = block_sum(rk, red);
  if (tid == 0) {
    row_loss[g] = rl / (float)len;
    row_kl[g] = rk / (float)len;
    row_len[g] = len;
    tickets[4 + g] = 0u;
    __threadfence();
    const unsigned int rows_done = atomicAdd(tickets, 1u);
    s_last = (rows_done == (unsigned int)G - 1u);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (tid == 0) {   // last row: scalar loss and mean KL in row order
    float a = 0.f, b = 0.f;
    for (int gg = 0; gg < G; ++gg) { a += __ldcg(row_loss + gg); b += __ldcg(row_kl + gg); }
    out2[0] = a / G;
    out2[1] = b / G;
    tickets[0] = 0u;                   // ready for the next launch
  }
}

}  // namespace

extern "C" int sb_grpo_loss_workspace(int G, int C, long long* floats_out) {
  SB_REQUIRE(floats_out && G > 0 && C > 0, "sb_grpo_loss_workspace: bad arguments");
  *floats_out = 4 + ((G + 3) & ~3) + 2LL * G * ((C + LT - 1) / LT);
  return 0;
}

extern "C" int sb_grpo_loss(const float* lse_part, int n_tiles, const float* tgt_logit, const int* comp_ids, int G,
                            int C, int eos_id, const float* ref_lp, const float* adv, float beta, float* lp_out,
                            float* lse_out, float* coef_out, int* mask_out, float* row_loss, float* row_kl,
                            int* row_len, float* out2, float* workspace, sb_stream_t stream) {
  SB_REQUIRE(lse_part && tgt_logit && comp_ids && adv && lp_out && lse_out && coef_out && row_loss && row_kl &&
                 row_len && out2 && workspace, "sb_grpo_loss: null pointer");
  SB_REQUIRE(G > 0 && C > 0 && n_tiles > 0, "sb_grpo_loss: bad sizes");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (launch_fold(lse_part, n_tiles, tgt_logit, lp_out, lse_out, (long long)G * C, st, "sb_grpo_loss(fold)")) return 1;
  dim3 grid((C + LT - 1) / LT, G);
  grpo_loss_kernel<<<grid, LT, 0, st>>>(comp_ids, C, eos_id, ref_lp, adv, beta, G, lp_out, coef_out, mask_out, row_loss,
                                        row_kl, row_len, out2, workspace);
  return sb_check_launch("sb_grpo_loss");
}

extern "C" int sb_logprob_from_partials(const float* lse_part, int n_tiles, const float* tgt_logit, float* lp_out,
                                        long long rows, sb_stream_t stream) {
  SB_REQUIRE(lse_part && tgt_logit && lp_out && rows > 0 && n_tiles > 0, "sb_logprob_from_partials: bad arguments");
  if (launch_fold(lse_part, n_tiles, tgt_logit, lp_out, nullptr, rows, reinterpret_cast<cudaStream_t>(stream),
                  "sb_logprob_from_partials")) return 1;
  return sb_check_launch("sb_logprob_from_partials");
}
