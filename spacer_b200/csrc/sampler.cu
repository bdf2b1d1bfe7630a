// Fused top-p sampler for rollout: one thread-block CLUSTER of 8 CTAs per decode row; each CTA keeps its
// 1/8 slice of the row's logits in shared memory, partial reductions travel through distributed shared
// memory, phases are separated by cluster barriers.  No sort: the top-p cut is found by a two-level radix
// select (8 + 8 bits) over the order-preserving 16-bit keys of the bf16-rounded logits.
//
// Semantics follow the reference's generate() configuration (SG_RLVR_trainer.py:277-284: do_sample,
// temperature 1, top_p 0.95) as implemented by transformers:
//   logits -> fp32 (generation/utils.py:2763), TopPLogitsWarper (logits_process.py:521-533: ascending sort,
//   drop the prefix whose cumulative probability <= 1 - top_p, keep >= 1 token; ties are dropped in ascending
//   index order like a stable sort), softmax, multinomial (utils.py:2789-2791), finished rows emit pad
//   (utils.py:2797), EOS marks a row finished.  The random stream is Philox4x32-10 (not torch's generator),
//   so parity with the reference is distributional, not bitwise.
#include "common.cuh"
#include "spacer_b200.h"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace {

constexpr int CL = 8;            // CTAs per cluster (= per row)
constexpr int ST = 512;          // threads per CTA
constexpr int NW = ST / 32;

SB_DEVICE uint32_t key16(float v) {
  // order-preserving map of a bf16-representable float to 16 bits (ascending)
  uint32_t u = __float_as_uint(v) >> 16;
  return (u & 0x8000u) ? (~u & 0xFFFFu) : (u | 0x8000u);
}

SB_DEVICE void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                             uint32_t* out) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct SampleParams {
  const float* logits; long long ld; int V;
  float top_p;
  int top_k;                // 0 = off (TopKLogitsWarper, logits_process.py: scores < k-th largest removed, ties kept)
  float inv_temp;           // 1 / temperature (TemperatureLogitsWarper)
  float rep_pen;            // repetition penalty (RepetitionPenaltyLogitsProcessor); 1 = off
  const uint32_t* seen_r;   // bitmap [R][seen_ld] of the tokens already in prompt + completion (repetition penalty)
  uint32_t* seen_w;         // same buffer, updated with the sampled token
  long long seen_ld;
  unsigned long long seed;
  const long long* seed_dev;  // optional: overrides `seed` (lets one captured CUDA graph serve every rollout)
  const int* step_ptr;
  int* finished;            // [R] in/out
  int* out_tokens;          // [R]
  int* out_ids; long long out_ld;   // optional [R][out_ld], column = step
  float* out_logprob;       // optional [R]: log-prob of the sampled token under the filtered distribution
  int eos[4]; int n_eos;    // every id in the model's eos list finishes a row (generation/utils.py stopping criteria)
  int pad_id, suppress_eos;
  int slice;                // elements per CTA
};

SB_DEVICE bool is_eos(const SampleParams& p, int tok) {
  bool e = false;
#pragma unroll
  for (int k = 0; k < 4; ++k) e |= (k < p.n_eos && tok == p.eos[k]);
  return e;
}

// logit of token gi of `row` as the samplers see it: the model's bf16 logit, then (HF's processor order) repetition
// penalty and temperature in fp32.  Transformed values are rounded to bf16 again (the radix select works on 16-bit keys;
// HF keeps fp32 there: a deliberate deviation), untransformed ones are exactly the model's.
SB_DEVICE float load_logit(const SampleParams& p, const float* lg, int row, int gi) {
  float v = bf16_round(lg[gi]);
  bool changed = false;
  if (p.seen_r != nullptr && p.rep_pen != 1.f && ((p.seen_r[(long long)row * p.seen_ld + (gi >> 5)] >> (gi & 31)) & 1u)) {
    v = v < 0.f ? v * p.rep_pen : v / p.rep_pen;
    changed = true;
  }
  if (p.inv_temp != 1.f) { v *= p.inv_temp; changed = true; }
  if (changed) v = bf16_round(v);
  if (p.suppress_eos && is_eos(p, gi)) v = -INFINITY;
  return v;
}

SB_DEVICE void emit_token(const SampleParams& p, int row, int step, int tok) {
  const int fin = p.finished ? p.finished[row] : 0;
  if (fin) tok = p.pad_id;
  else {
    if (is_eos(p, tok) && p.finished) p.finished[row] = 1;
    if (p.seen_w) atomicOr(&p.seen_w[(long long)row * p.seen_ld + (tok >> 5)], 1u << (tok & 31));
  }
  p.out_tokens[row] = tok;
  if (p.out_ids) p.out_ids[(long long)row * p.out_ld + step] = tok;
}

struct Shared {
  float part_f[4];          // [0] max, [1] Z, [2] kept mass
  float hist_m[256];
  int hist_c[256];
  float scan_f[ST];
  int scan_i[ST];
  float g_hist[256];
  int g_cnt[256];
  float wtot_f[NW + 1];
  int wtot_i[NW + 1];
  int b1, b2, sel, last;
  int kb, kabove;
  float below;
};

// block-wide exclusive scans (ST threads); `total` is returned to every thread
SB_DEVICE float excl_scan_f(float v, float* wtot, float& total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();
  if (lane == 31) wtot[w] = inc;
  __syncthreads();
  float off = 0.f, tot = 0.f;
  for (int k = 0; k < NW; ++k) { if (k < w) off += wtot[k]; tot += wtot[k]; }
  total = tot;
  return off + inc - v;
}
SB_DEVICE int excl_scan_i(int v, int* wtot, int& total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();
  if (lane == 31) wtot[w] = inc;
  __syncthreads();
  int off = 0, tot = 0;
  for (int k = 0; k < NW; ++k) { if (k < w) off += wtot[k]; tot += wtot[k]; }
  total = tot;
  return off + inc - v;
}

// lowest bin b such that start + sum_{k<=b} g[k] > thr; below = start + sum_{k<b} g[k]   (all threads call)
SB_DEVICE void find_bin(Shared* sh, float start, float thr, int* out_b, float* out_below) {
  const int tid = threadIdx.x;
  float tot;
  const float g = tid < 256 ? sh->g_hist[tid] : 0.f;
  const float ex = excl_scan_f(g, sh->wtot_f, tot);
  if (tid == 0) { *out_b = 256; }
  __syncthreads();
  if (tid < 256 && start + ex + g > thr) atomicMin(out_b, tid);
  __syncthreads();
  int b = *out_b;
  if (b == 256) b = 255;
  __syncthreads();
  if (tid == b) { *out_below = start + ex; *out_b = b; }
  __syncthreads();
}

// top-k: highest bin b (of sh->g_cnt) such that count(bins > b) < k <= count(bins >= b); above = count(bins > b).
// Returns the total count; when total < k nothing is found (*out_b = -1).   (all threads call)
SB_DEVICE int find_bin_top(Shared* sh, int k, int* out_b, int* out_above) {
  const int tid = threadIdx.x;
  const int bin = 255 - tid;
  const int c = tid < 256 ? sh->g_cnt[bin] : 0;
  int tot;
  const int above = excl_scan_i(c, sh->wtot_i, tot);
  if (tid == 0) { *out_b = -1; *out_above = 0; }
  __syncthreads();
  if (tid < 256 && above < k && above + c >= k) { *out_b = bin; *out_above = above; }
  __syncthreads();
  return tot;
}

SB_DEVICE bool kept_token(uint32_t k, uint32_t kstar, int& tie_rank, int n_drop) {
  if (k > kstar) return true;
  if (k < kstar) return false;
  const bool keep = tie_rank >= n_drop;
  ++tie_rank;
  return keep;
}

// Everything after the slice is in shared memory: cluster max, exact top-k by count, nucleus cut by mass, draw.  CLUSTER:
// the 8 CTAs of a row work on their slices and exchange histograms through distributed shared memory; !CLUSTER: one CTA
// alone on a short list (the compacted top-k candidates, `gidx` = their token ids), same code, same semantics.
template <bool CLUSTER>
SB_DEVICE void sample_core(const SampleParams& p, Shared* sh, float* wh_m, int* wh_c, float* sl, int n, int i0,
                           const int* gidx, int row, float mx, float* red) {
  cg::cluster_group cluster = cg::this_cluster();
  constexpr int NR = CLUSTER ? CL : 1;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int rank = CLUSTER ? (int)cluster.block_rank() : 0;
  auto csync = [&]() { if constexpr (CLUSTER) cluster.sync(); else __syncthreads(); };
  auto peer = [&](int c) -> Shared* { if constexpr (CLUSTER) return cluster.map_shared_rank(sh, c); else return sh; };
  mx = block_max(mx, red);
  if (tid == 0) sh->part_f[0] = mx;
  csync();
  float m = -INFINITY;
  for (int c = 0; c < NR; ++c) m = fmaxf(m, peer(c)->part_f[0]);

  // ---- phase A2 (top_k > 0): two-level radix select of the k-th largest key by COUNT; everything strictly below it
  // is removed (-inf) before the nucleus cut, ties at the k-th value stay (TopKLogitsWarper)
  if (p.top_k > 0 && p.top_k < p.V) {
    uint32_t kth_key = 0;
    int want = p.top_k, hi_bin = -1;
    for (int level = 0; level < 2; ++level) {
      for (int i = tid; i < NW * 256; i += ST) wh_c[i] = 0;
      __syncthreads();
      for (int i = tid; i < n; i += ST) {
        const float v = sl[i];
        if (v == -INFINITY) continue;
        const uint32_t k = key16(v);
        if (level == 0) atomicAdd(&wh_c[warp * 256 + (k >> 8)], 1);
        else if ((int)(k >> 8) == hi_bin) atomicAdd(&wh_c[warp * 256 + (k & 255)], 1);
      }
      __syncthreads();
      if (tid < 256) {
        int c = 0;
        for (int w = 0; w < NW; ++w) c += wh_c[w * 256 + tid];
        sh->hist_c[tid] = c;
      }
      csync();
      if (tid < 256) {
        int c = 0;
        for (int q = 0; q < NR; ++q) c += peer(q)->hist_c[tid];
        sh->g_cnt[tid] = c;
      }
      __syncthreads();
      find_bin_top(sh, want, &sh->kb, &sh->kabove);
      const int b = sh->kb, above = sh->kabove;
      csync();   // remote reads of hist_c are done before the next level overwrites it
      if (b < 0) { kth_key = 0; break; }          // fewer than k candidates: keep all (CTA- and cluster-uniform)
      if (level == 0) { hi_bin = b; want -= above; }
      else kth_key = ((uint32_t)hi_bin << 8) | (uint32_t)b;
    }
    for (int i = tid; i < n; i += ST) {
      const float v = sl[i];
      if (v != -INFINITY && key16(v) < kth_key) sl[i] = -INFINITY;
    }
    __syncthreads();
  }

  // ---- phase B: Z and level-1 histogram (high 8 key bits) of probability mass
  for (int i = tid; i < NW * 256; i += ST) { wh_m[i] = 0.f; wh_c[i] = 0; }
  __syncthreads();
  float z = 0.f;
  for (int i = tid; i < n; i += ST) {
    const float v = sl[i];
    if (v == -INFINITY) continue;
    const float e = __expf(v - m);
    z += e;
    atomicAdd(&wh_m[warp * 256 + (key16(v) >> 8)], e);
  }
  z = block_sum(z, red);
  __syncthreads();
  if (tid < 256) {
    float s = 0.f;
    for (int w = 0; w < NW; ++w) s += wh_m[w * 256 + tid];
    sh->hist_m[tid] = s;
  }
  if (tid == 0) sh->part_f[1] = z;
  csync();
  float Z = 0.f;
  for (int c = 0; c < NR; ++c) Z += peer(c)->part_f[1];
  const float thr = (1.f - p.top_p) * Z;   // unnormalised mass that may be removed (cum <= thr)
  if (tid < 256) {
    float s = 0.f;
    for (int c = 0; c < NR; ++c) s += peer(c)->hist_m[tid];
    sh->g_hist[tid] = s;
  }
  __syncthreads();
  find_bin(sh, 0.f, thr, &sh->b1, &sh->below);
  const int b1 = sh->b1;
  const float below1 = sh->below;
  csync();  // all remote reads of hist_m are done before it is overwritten

  // ---- phase C: level-2 histogram (low 8 key bits) inside bin b1: mass and counts
  for (int i = tid; i < NW * 256; i += ST) { wh_m[i] = 0.f; wh_c[i] = 0; }
  __syncthreads();
  for (int i = tid; i < n; i += ST) {
    const float v = sl[i];
    if (v == -INFINITY) continue;
    const uint32_t k = key16(v);
    if ((int)(k >> 8) == b1) {
      atomicAdd(&wh_m[warp * 256 + (k & 255)], __expf(v - m));
      atomicAdd(&wh_c[warp * 256 + (k & 255)], 1);
    }
  }
  __syncthreads();
  if (tid < 256) {
    float s = 0.f;
    int c = 0;
    for (int w = 0; w < NW; ++w) { s += wh_m[w * 256 + tid]; c += wh_c[w * 256 + tid]; }
    sh->hist_m[tid] = s;
    sh->hist_c[tid] = c;
  }
  csync();
  if (tid < 256) {
    float s = 0.f;
    for (int c = 0; c < NR; ++c) s += peer(c)->hist_m[tid];
    sh->g_hist[tid] = s;
  }
  __syncthreads();
  find_bin(sh, below1, thr, &sh->b2, &sh->below);
  const int b2 = sh->b2;
  const float below = sh->below;
  const uint32_t kstar = ((uint32_t)b1 << 8) | (uint32_t)b2;
  // ties at the threshold key share one probability; `n_drop` of them (lowest indices first) are removed
  int tie_total = 0, tie_before = 0;
  for (int c = 0; c < NR; ++c) {
    const int tc = peer(c)->hist_c[b2];
    tie_total += tc;
    if (c < rank) tie_before += tc;
  }
  const uint32_t kb = (kstar & 0x8000u) ? (kstar & 0x7FFFu) : (~kstar & 0xFFFFu);
  const float tie_e = __expf(__uint_as_float(kb << 16) - m);
  int n_drop = tie_e > 0.f ? (int)floorf((thr - below) / tie_e) : 0;
  n_drop = max(0, min(n_drop, tie_total - 1));

  // ---- phase D: kept mass (each thread owns a contiguous index chunk so tie ranks follow index order)
  const int per = (n + ST - 1) / ST;
  const int a0 = min(tid * per, n), a1 = min(a0 + per, n);
  int my_ties = 0;
  for (int i = a0; i < a1; ++i) my_ties += (sl[i] != -INFINITY && key16(sl[i]) == kstar) ? 1 : 0;
  int ties_cta;
  const int tie_rank0 = tie_before + excl_scan_i(my_ties, sh->wtot_i, ties_cta);
  float kept = 0.f;
  {
    int tr = tie_rank0;
    for (int i = a0; i < a1; ++i) {
      const float v = sl[i];
      if (v == -INFINITY) continue;
      if (kept_token(key16(v), kstar, tr, n_drop)) kept += __expf(v - m);
    }
  }
  float my_kept;
  const float pre = excl_scan_f(kept, sh->wtot_f, my_kept);
  if (tid == 0) { sh->part_f[2] = my_kept; sh->sel = -1; sh->last = -1; }
  csync();
  float kept_total = 0.f, kept_before = 0.f;
  bool later_mass = false;
  for (int c = 0; c < NR; ++c) {
    const float kc = peer(c)->part_f[2];
    if (c < rank) kept_before += kc;
    if (c > rank && kc > 0.f) later_mass = true;
    kept_total += kc;
  }

  // ---- phase E: draw u, locate the token in index order
  const int step = *p.step_ptr;
  uint32_t rnd[4];
  const unsigned long long seed = p.seed_dev ? (unsigned long long)*p.seed_dev : p.seed;
  philox4x32_10((uint32_t)step, (uint32_t)row, 0x5BACE200u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), rnd);
  const float u = ((rnd[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float target = u * kept_total;
  // the CTA whose kept-mass interval holds target owns the draw; the last CTA with mass absorbs round-off
  const bool mine = my_kept > 0.f && target >= kept_before && (target < kept_before + my_kept || !later_mass);
  if (mine) {   // CTA-uniform
    const float local = target - kept_before;
    if (kept > 0.f) {
      atomicMax(&sh->last, tid);
      if (local >= pre && local < pre + kept) atomicMax(&sh->sel, tid);
    }
    __syncthreads();
    const int sel = sh->sel >= 0 ? sh->sel : sh->last;
    if (tid == sel) {
      float run = pre;
      int chosen = -1, last_kept = -1;
      float chosen_e = 0.f, last_e = 0.f;
      int tr = tie_rank0;
      for (int i = a0; i < a1 && chosen < 0; ++i) {
        const float v = sl[i];
        if (v == -INFINITY) continue;
        if (!kept_token(key16(v), kstar, tr, n_drop)) continue;
        const float e = __expf(v - m);
        last_kept = i; last_e = e;
        if (local < run + e) { chosen = i; chosen_e = e; }
        run += e;
      }
      if (chosen < 0) { chosen = last_kept; chosen_e = last_e; }
      emit_token(p, row, step, gidx ? gidx[chosen] : i0 + chosen);
      if (p.out_logprob) p.out_logprob[row] = logf(chosen_e / kept_total);
    }
  }
}

constexpr int KCAP = 128;       // top-k fast path: candidates one CTA may contribute (local top-k incl. ties)
constexpr int KMAX = 64;        // ... for top_k up to this

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(ST)
sample_kernel(const SampleParams p) {
  pdl_launch_dependents();
  const int tr = (blockIdx.x == 0 && threadIdx.x == 0) ? sb_trace_begin(SB_TR_SAMPLE) : -1;
  pdl_wait();
  sb_trace_mark(tr, 1);
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) uint8_t smem_raw[];
  Shared* sh = reinterpret_cast<Shared*>(smem_raw);
  float* wh_m = reinterpret_cast<float*>(smem_raw + ((sizeof(Shared) + 15) & ~15));  // [NW][256] per-warp mass
  int* wh_c = reinterpret_cast<int*>(wh_m + NW * 256);                               // [NW][256] per-warp counts
  float* sl = reinterpret_cast<float*>(wh_c + NW * 256);                             // this CTA's logits slice
  __shared__ float red[32];
  __shared__ float c_val[CL * KCAP];      // rank 0: the row's compacted top-k candidates, ascending token id
  __shared__ int c_idx[CL * KCAP];
  __shared__ int c_cnt;                   // candidates of this CTA (<= KCAP, or KCAP + 1 = overflow)

  const int tid = threadIdx.x, warp = tid >> 5;
  const int rank = cluster.block_rank();
  const int row = blockIdx.x / CL;
  const int i0 = rank * p.slice;
  const int n = max(0, min(p.slice, p.V - i0));
  const float* lg = p.logits + (long long)row * p.ld;

  // ---- phase A: load slice (bf16-rounded)
  float mx = -INFINITY;
  for (int i = tid; i < n; i += ST) {
    const float v = load_logit(p, lg, row, i0 + i);
    sl[i] = v;
    mx = fmaxf(mx, v);
  }

  // ---- top-k fast path (0 < top_k <= KMAX, the rollout's configuration): the row's top-k set is contained in the union
  // of the slices' LOCAL top-k sets (ties at a local k-th value included), so every CTA selects its own <= KCAP
  // candidates without talking to the others, the candidates are gathered in rank 0's shared memory in ascending token
  // order (two cluster barriers), and rank 0 alone finishes on that short list.  The generic path below needs ten cluster
  // barriers and seven passes over the slices; it still runs when a slice has more than KCAP candidates (ties).
  if (p.top_k > 0 && p.top_k <= KMAX) {
    __syncthreads();
    // local k-th largest key by count (two-level radix select on this CTA's histogram only)
    uint32_t kth_key = 0;
    int want = p.top_k, hi_bin = -1;
    for (int level = 0; level < 2; ++level) {
      for (int i = tid; i < NW * 256; i += ST) wh_c[i] = 0;
      __syncthreads();
      for (int i = tid; i < n; i += ST) {
        const float v = sl[i];
        if (v == -INFINITY) continue;
        const uint32_t k = key16(v);
        if (level == 0) atomicAdd(&wh_c[warp * 256 + (k >> 8)], 1);
        else if ((int)(k >> 8) == hi_bin) atomicAdd(&wh_c[warp * 256 + (k & 255)], 1);
      }
      __syncthreads();
      if (tid < 256) {
        int c = 0;
        for (int w = 0; w < NW; ++w) c += wh_c[w * 256 + tid];
        sh->g_cnt[tid] = c;
      }
      __syncthreads();
      find_bin_top(sh, want, &sh->kb, &sh->kabove);
      const int b = sh->kb, above = sh->kabove;
      __syncthreads();
      if (b < 0) { kth_key = 0; break; }          // fewer than k finite entries in this slice: all of them are candidates
      if (level == 0) { hi_bin = b; want -= above; }
      else kth_key = ((uint32_t)hi_bin << 8) | (uint32_t)b;
    }
    // compaction in index order: each thread owns a contiguous chunk
    const int per = (n + ST - 1) / ST;
    const int a0 = min(tid * per, n), a1 = min(a0 + per, n);
    int mine = 0;
    for (int i = a0; i < a1; ++i) mine += (sl[i] != -INFINITY && key16(sl[i]) >= kth_key) ? 1 : 0;
    int total;
    const int pos0 = excl_scan_i(mine, sh->wtot_i, total);
    if (tid == 0) c_cnt = total <= KCAP ? total : KCAP + 1;
    cluster.sync();
    int before = 0;
    bool overflow = false;
    for (int c = 0; c < CL; ++c) {
      const int cc = *cluster.map_shared_rank(&c_cnt, c);
      overflow |= cc > KCAP;
      if (c < rank) before += cc;
    }
    if (!overflow) {     // cluster-uniform
      float* dv = cluster.map_shared_rank(c_val, 0);
      int* di = cluster.map_shared_rank(c_idx, 0);
      int pos = before + pos0;
      for (int i = a0; i < a1; ++i) {
        const float v = sl[i];
        if (v != -INFINITY && key16(v) >= kth_key) { dv[pos] = v; di[pos] = i0 + i; ++pos; }
      }
      int n_all = 0;
      for (int c = 0; c < CL; ++c) n_all += *cluster.map_shared_rank(&c_cnt, c);
      cluster.sync();    // candidates are in rank 0's shared memory; nobody reads remote memory after this point
      if (rank != 0) return;
      float m2 = -INFINITY;
      for (int i = tid; i < n_all; i += ST) m2 = fmaxf(m2, c_val[i]);
      sample_core<false>(p, sh, wh_m, wh_c, c_val, n_all, 0, c_idx, row, m2, red);
      sb_trace_mark(tr, 2);
      return;
    }
    __syncthreads();
  }

  sample_core<true>(p, sh, wh_m, wh_c, sl, n, i0, nullptr, row, mx, red);
  cluster.sync();  // keep every CTA's shared memory alive until all remote reads are done
  sb_trace_mark(tr, 2);
}

// Greedy decoding (SpaceR-Eval: `model.generate(..., temperature=0.01)` on top of the checkpoint's top_k = 1 generation
// config, data_utils/vsibench.py:174; HF `do_sample=False`): token = argmax of the bf16-rounded logits, lowest index on
// ties (torch.argmax); EOS / pad bookkeeping identical to sample_kernel.  One CTA per row.
__global__ void __launch_bounds__(1024)
greedy_kernel(const SampleParams p) {
  pdl_launch_dependents();
  const int tr = (blockIdx.x == 0 && threadIdx.x == 0) ? sb_trace_begin(SB_TR_SAMPLE) : -1;
  pdl_wait();
  sb_trace_mark(tr, 1);
  __shared__ float s_v[32];
  __shared__ int s_i[32];
  const int row = blockIdx.x;
  const float* lg = p.logits + (long long)row * p.ld;
  float best = -INFINITY;
  int best_i = 0x7fffffff;
  for (int i = threadIdx.x; i < p.V; i += blockDim.x) {
    const float v = load_logit(p, lg, row, i);
    if (v > best) { best = v; best_i = i; }       // ascending i per thread: the first maximum is kept
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { s_v[w] = best; s_i[w] = best_i; }
  __syncthreads();
  if (w == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    best = lane < nw ? s_v[lane] : -INFINITY;
    best_i = lane < nw ? s_i[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
      if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
    }
    if (lane == 0) emit_token(p, row, *p.step_ptr, best_i);
  }
  sb_trace_mark(tr, 2);
}

// bitmap[row][tok] = 1 for every token of ids[0..n) and every row (the prompt is shared by the rows of a rollout)
__global__ void bitmap_set_kernel(const int* __restrict__ ids, int n, uint32_t* __restrict__ seen, long long seen_ld,
                                  int rows, int V) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int tok = ids[i];
  if (tok < 0 || tok >= V) return;
  for (int r = 0; r < rows; ++r) atomicOr(&seen[(long long)r * seen_ld + (tok >> 5)], 1u << (tok & 31));
}

__global__ void step_advance_kernel(int* step_ptr) {
  pdl_launch_dependents();
  const int tr = sb_trace_begin(SB_TR_ADVANCE);
  pdl_wait();
  sb_trace_mark(tr, 1);
  *step_ptr += 1;
  sb_trace_mark(tr, 2);
}

}  // namespace

SB_DEFINE_TRACE_SETTER(sb_trace_set_sampler)

static int fill_params(const sb_sample_args* a, SampleParams& p, const char* who) {
  SB_REQUIRE(a && a->logits && a->step_ptr && a->out_tokens && a->R > 0 && a->V > 0, "%s: bad arguments", who);
  SB_REQUIRE(a->n_eos >= 0 && a->n_eos <= 4, "%s: at most 4 eos ids, got %d", who, a->n_eos);
  SB_REQUIRE(a->temperature > 0.f, "%s: temperature must be > 0, got %f", who, a->temperature);
  SB_REQUIRE(a->repetition_penalty > 0.f, "%s: repetition_penalty must be > 0, got %f", who, a->repetition_penalty);
  SB_REQUIRE(a->repetition_penalty == 1.f || a->seen != nullptr, "%s: repetition_penalty needs the token bitmap", who);
  SB_REQUIRE(a->seen == nullptr || a->seen_ld * 32 >= a->V, "%s: token bitmap narrower than the vocabulary", who);
  p.logits = a->logits; p.ld = a->ld; p.V = a->V; p.top_p = a->top_p; p.top_k = a->top_k;
  p.inv_temp = 1.f / a->temperature; p.rep_pen = a->repetition_penalty;
  p.seen_r = a->seen; p.seen_w = a->seen; p.seen_ld = a->seen_ld;
  p.seed = a->seed; p.seed_dev = a->seed_dev; p.step_ptr = a->step_ptr;
  p.finished = a->finished; p.out_tokens = a->out_tokens; p.out_ids = a->out_ids; p.out_ld = a->out_ld;
  p.out_logprob = a->out_logprob;
  for (int k = 0; k < 4; ++k) p.eos[k] = k < a->n_eos ? a->eos_ids[k] : -1;
  p.n_eos = a->n_eos; p.pad_id = a->pad_id; p.suppress_eos = a->suppress_eos;
  p.slice = ((a->V + CL - 1) / CL + 3) & ~3;
  return 0;
}

extern "C" int sb_sample(const sb_sample_args* a, sb_stream_t stream) {
  SampleParams p;
  if (fill_params(a, p, "sb_sample")) return 1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (a->mode == 1) {
    SB_CUDA(sb_launch(greedy_kernel, dim3(a->R), dim3(1024), 0, st, sb_pdl_enabled(), p));
    return sb_check_launch("sb_sample(greedy)");
  }
  SB_REQUIRE(a->mode == 0, "sb_sample: mode must be 0 (sample) or 1 (greedy)");
  SB_REQUIRE(a->top_p > 0.f && a->top_p <= 1.f, "sb_sample: top_p must be in (0,1], got %f", a->top_p);
  SB_REQUIRE(a->top_k >= 0, "sb_sample: top_k must be >= 0");
  const size_t smem = ((sizeof(Shared) + 15) & ~(size_t)15) + (size_t)NW * 256 * 8 + (size_t)p.slice * 4;
  SB_REQUIRE(smem <= 220 * 1024, "sb_sample: vocabulary %d too large for the 8-CTA cluster sampler", a->V);
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    SB_CUDA(cudaFuncSetAttribute(sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  SB_CUDA(sb_launch(sample_kernel, dim3(a->R * CL), dim3(ST), smem, st, sb_pdl_enabled(), p));
  return sb_check_launch("sb_sample");
}

extern "C" int sb_token_bitmap_set(const int* ids, int n, unsigned int* seen, long long seen_ld, int rows, int V,
                                   sb_stream_t stream) {
  SB_REQUIRE(ids && seen && n >= 0 && rows > 0 && seen_ld * 32 >= V, "sb_token_bitmap_set: bad arguments");
  if (n == 0) return 0;
  SB_CUDA(sb_launch(bitmap_set_kernel, dim3((n + 255) / 256), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), false,
                    ids, n, seen, seen_ld, rows, V));
  return sb_check_launch("sb_token_bitmap_set");
}

// the round-1 entry points: one eos id, no top-k / temperature / repetition penalty
extern "C" int sb_sample_greedy(const float* logits, long long ld, int R, int V, const int* step_ptr, int* finished,
                                int* out_tokens, int* out_ids, long long out_ld, int eos_id, int pad_id, int suppress_eos,
                                sb_stream_t stream) {
  sb_sample_args a = {};
  a.logits = logits; a.ld = ld; a.R = R; a.V = V; a.mode = 1; a.top_p = 1.f; a.temperature = 1.f;
  a.repetition_penalty = 1.f; a.step_ptr = step_ptr; a.finished = finished; a.out_tokens = out_tokens;
  a.out_ids = out_ids; a.out_ld = out_ld; a.eos_ids[0] = eos_id; a.n_eos = 1; a.pad_id = pad_id;
  a.suppress_eos = suppress_eos;
  return sb_sample(&a, stream);
}

extern "C" int sb_sample_top_p(const float* logits, long long ld, int R, int V, float top_p, unsigned long long seed,
                               const int* step_ptr, int* finished, int* out_tokens, int* out_ids, long long out_ld,
                               float* out_logprob, int eos_id, int pad_id, int suppress_eos, const long long* seed_dev,
                               sb_stream_t stream) {
  sb_sample_args a = {};
  a.logits = logits; a.ld = ld; a.R = R; a.V = V; a.mode = 0; a.top_p = top_p; a.temperature = 1.f;
  a.repetition_penalty = 1.f; a.seed = seed; a.seed_dev = seed_dev; a.step_ptr = step_ptr; a.finished = finished;
  a.out_tokens = out_tokens; a.out_ids = out_ids; a.out_ld = out_ld; a.out_logprob = out_logprob;
  a.eos_ids[0] = eos_id; a.n_eos = 1; a.pad_id = pad_id; a.suppress_eos = suppress_eos;
  return sb_sample(&a, stream);
}

extern "C" int sb_step_advance(int* step_ptr, sb_stream_t stream) {
  SB_REQUIRE(step_ptr, "sb_step_advance: null pointer");
  SB_CUDA(sb_launch(step_advance_kernel, dim3(1), dim3(1), 0, reinterpret_cast<cudaStream_t>(stream), sb_pdl_enabled(), step_ptr));
  return sb_check_launch("sb_step_advance");
}
