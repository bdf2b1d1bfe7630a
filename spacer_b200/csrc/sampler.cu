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
  unsigned long long seed;
  const long long* seed_dev;  // optional: overrides `seed` (lets one captured CUDA graph serve every rollout)
  const int* step_ptr;
  int* finished;            // [R] in/out
  int* out_tokens;          // [R]
  int* out_ids; long long out_ld;   // optional [R][out_ld], column = step
  float* out_logprob;       // optional [R]: log-prob of the sampled token under the filtered distribution
  int eos_id, pad_id, suppress_eos;
  int slice;                // elements per CTA
};

struct Shared {
  float part_f[4];          // [0] max, [1] Z, [2] kept mass
  float hist_m[256];
  int hist_c[256];
  float scan_f[ST];
  int scan_i[ST];
  float g_hist[256];
  float wtot_f[NW + 1];
  int wtot_i[NW + 1];
  int b1, b2, sel, last;
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

SB_DEVICE bool kept_token(uint32_t k, uint32_t kstar, int& tie_rank, int n_drop) {
  if (k > kstar) return true;
  if (k < kstar) return false;
  const bool keep = tie_rank >= n_drop;
  ++tie_rank;
  return keep;
}

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

  const int tid = threadIdx.x, warp = tid >> 5;
  const int rank = cluster.block_rank();
  const int row = blockIdx.x / CL;
  const int i0 = rank * p.slice;
  const int n = max(0, min(p.slice, p.V - i0));
  const float* lg = p.logits + (long long)row * p.ld;

  // ---- phase A: load slice (bf16-rounded), cluster max
  float mx = -INFINITY;
  for (int i = tid; i < n; i += ST) {
    float v = bf16_round(lg[i0 + i]);
    if (p.suppress_eos && i0 + i == p.eos_id) v = -INFINITY;
    sl[i] = v;
    mx = fmaxf(mx, v);
  }
  mx = block_max(mx, red);
  if (tid == 0) sh->part_f[0] = mx;
  cluster.sync();
  float m = -INFINITY;
  for (int c = 0; c < CL; ++c) m = fmaxf(m, cluster.map_shared_rank(sh, c)->part_f[0]);

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
  cluster.sync();
  float Z = 0.f;
  for (int c = 0; c < CL; ++c) Z += cluster.map_shared_rank(sh, c)->part_f[1];
  const float thr = (1.f - p.top_p) * Z;   // unnormalised mass that may be removed (cum <= thr)
  if (tid < 256) {
    float s = 0.f;
    for (int c = 0; c < CL; ++c) s += cluster.map_shared_rank(sh, c)->hist_m[tid];
    sh->g_hist[tid] = s;
  }
  __syncthreads();
  find_bin(sh, 0.f, thr, &sh->b1, &sh->below);
  const int b1 = sh->b1;
  const float below1 = sh->below;
  cluster.sync();  // all remote reads of hist_m are done before it is overwritten

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
  cluster.sync();
  if (tid < 256) {
    float s = 0.f;
    for (int c = 0; c < CL; ++c) s += cluster.map_shared_rank(sh, c)->hist_m[tid];
    sh->g_hist[tid] = s;
  }
  __syncthreads();
  find_bin(sh, below1, thr, &sh->b2, &sh->below);
  const int b2 = sh->b2;
  const float below = sh->below;
  const uint32_t kstar = ((uint32_t)b1 << 8) | (uint32_t)b2;
  // ties at the threshold key share one probability; `n_drop` of them (lowest indices first) are removed
  int tie_total = 0, tie_before = 0;
  for (int c = 0; c < CL; ++c) {
    const int tc = cluster.map_shared_rank(sh, c)->hist_c[b2];
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
  cluster.sync();
  float kept_total = 0.f, kept_before = 0.f;
  bool later_mass = false;
  for (int c = 0; c < CL; ++c) {
    const float kc = cluster.map_shared_rank(sh, c)->part_f[2];
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
      int tok = i0 + chosen;
      const int fin = p.finished ? p.finished[row] : 0;
      if (fin) tok = p.pad_id;
      else if (tok == p.eos_id && p.finished) p.finished[row] = 1;
      p.out_tokens[row] = tok;
      if (p.out_ids) p.out_ids[(long long)row * p.out_ld + step] = tok;
      if (p.out_logprob) p.out_logprob[row] = logf(chosen_e / kept_total);
    }
  }
  cluster.sync();  // keep every CTA's shared memory alive until all remote reads are done
  sb_trace_mark(tr, 2);
}

// Greedy decoding (SpaceR-Eval: `model.generate(..., temperature=0.01)` on top of the checkpoint's top_k = 1 generation
// config, data_utils/vsibench.py:174; HF `do_sample=False`): token = argmax of the bf16-rounded logits, lowest index on
// ties (torch.argmax); EOS / pad bookkeeping identical to sample_kernel.  One CTA per row.
__global__ void __launch_bounds__(1024)
greedy_kernel(const float* __restrict__ logits, long long ld, int V, const int* __restrict__ step_ptr,
              int* __restrict__ finished, int* __restrict__ out_tokens, int* __restrict__ out_ids, long long out_ld,
              int eos_id, int pad_id, int suppress_eos) {
  pdl_launch_dependents();
  const int tr = (blockIdx.x == 0 && threadIdx.x == 0) ? sb_trace_begin(SB_TR_SAMPLE) : -1;
  pdl_wait();
  sb_trace_mark(tr, 1);
  __shared__ float s_v[32];
  __shared__ int s_i[32];
  const int row = blockIdx.x;
  const float* lg = logits + (long long)row * ld;
  float best = -INFINITY;
  int best_i = 0x7fffffff;
  for (int i = threadIdx.x; i < V; i += blockDim.x) {
    float v = bf16_round(lg[i]);
    if (suppress_eos && i == eos_id) v = -INFINITY;
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
    if (lane == 0) {
      const int step = *step_ptr;
      int tok = best_i;
      const int fin = finished ? finished[row] : 0;
      if (fin) tok = pad_id;
      else if (tok == eos_id && finished) finished[row] = 1;
      out_tokens[row] = tok;
      if (out_ids) out_ids[(long long)row * out_ld + step] = tok;
    }
  }
  sb_trace_mark(tr, 2);
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

extern "C" int sb_sample_greedy(const float* logits, long long ld, int R, int V, const int* step_ptr, int* finished,
                                int* out_tokens, int* out_ids, long long out_ld, int eos_id, int pad_id, int suppress_eos,
                                sb_stream_t stream) {
  SB_REQUIRE(logits && step_ptr && out_tokens && R > 0 && V > 0, "sb_sample_greedy: bad arguments");
  SB_CUDA(sb_launch(greedy_kernel, dim3(R), dim3(1024), 0, reinterpret_cast<cudaStream_t>(stream), sb_pdl_enabled(), logits,
                    ld, V, step_ptr, finished, out_tokens, out_ids, out_ld, eos_id, pad_id, suppress_eos));
  return sb_check_launch("sb_sample_greedy");
}

extern "C" int sb_sample_top_p(const float* logits, long long ld, int R, int V, float top_p, unsigned long long seed,
                               const int* step_ptr, int* finished, int* out_tokens, int* out_ids, long long out_ld,
                               float* out_logprob, int eos_id, int pad_id, int suppress_eos, const long long* seed_dev,
                               sb_stream_t stream) {
  SB_REQUIRE(logits && step_ptr && out_tokens && R > 0 && V > 0, "sb_sample_top_p: bad arguments");
  SB_REQUIRE(top_p > 0.f && top_p <= 1.f, "sb_sample_top_p: top_p must be in (0,1], got %f", top_p);
  SampleParams p;
  p.logits = logits; p.ld = ld; p.V = V; p.top_p = top_p; p.seed = seed; p.seed_dev = seed_dev; p.step_ptr = step_ptr;
  p.finished = finished; p.out_tokens = out_tokens; p.out_ids = out_ids; p.out_ld = out_ld;
  p.out_logprob = out_logprob; p.eos_id = eos_id; p.pad_id = pad_id; p.suppress_eos = suppress_eos;
  p.slice = ((V + CL - 1) / CL + 3) & ~3;
  const size_t smem = ((sizeof(Shared) + 15) & ~(size_t)15) + (size_t)NW * 256 * 8 + (size_t)p.slice * 4;
  SB_REQUIRE(smem <= 220 * 1024, "sb_sample_top_p: vocabulary %d too large for the 8-CTA cluster sampler", V);
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    SB_CUDA(cudaFuncSetAttribute(sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  SB_CUDA(sb_launch(sample_kernel, dim3(R * CL), dim3(ST), smem, reinterpret_cast<cudaStream_t>(stream), sb_pdl_enabled(), p));
  return sb_check_launch("sb_sample_top_p");
}

extern "C" int sb_step_advance(int* step_ptr, sb_stream_t stream) {
  SB_REQUIRE(step_ptr, "sb_step_advance: null pointer");
  SB_CUDA(sb_launch(step_advance_kernel, dim3(1), dim3(1), 0, reinterpret_cast<cudaStream_t>(stream), sb_pdl_enabled(), step_ptr));
  return sb_check_launch("sb_step_advance");
}
