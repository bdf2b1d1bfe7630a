// Decode-step attention (q_len = 1) for R <= 32 rows that share one or two prompts -- "prefix-shared flash decoding".
// replaces the attention of the q_len=1 iterations of GenerationMixin._sample (generation/utils.py:2743-2806) through
// Qwen2VLAttention (MQ2:507-594) over DynamicCache (cache_utils.py:102-120), where the reference keeps G copies of the
// prompt KV and reads all of them every step.
//
// The context of row r is [shared prompt of r's group | r's own completion so far].  Work is cut into independent
// items, one CTA each, all in ONE launch:
//   prefix items (group, q-block, kv head, key split): the <= 64 query vectors (rows of the group x the q heads of the
//                kv head) against a slice of the SHARED prompt K/V -> the prompt cache is read once per group per
//                step, not once per row, and the products run on tensor cores (mma.sync m16n8k16, M = rows x heads);
//   own items    (row, kv head, key split): the row's q heads against a slice of its own completion cache.
// Every item writes a normalised partial (O, log2-sum-exp) for its (row, head) pairs; a combine kernel merges them.
// The step is read from device memory so that one captured CUDA graph replays for every step.
#include "common.cuh"
#include "spacer_b200.h"

namespace {

constexpr int HD = 128;
constexpr int LD = HD + 8;       // padded smem row (elements): conflict-free ldmatrix
constexpr int TQ = 64;           // query vectors per item (4 warps x 16)
constexpr int TK = 64;           // keys per pipeline stage
constexpr int THREADS = 128;
constexpr int TILE = 64 * LD;    // elements of one [64][LD] tile

SB_DEVICE void cp_async16(uint32_t dst, const void* src, bool pred) {
  const int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
SB_DEVICE void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
SB_DEVICE void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
SB_DEVICE void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
SB_DEVICE void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
SB_DEVICE void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct Plan {
  int rep;            // q heads per kv head
  int n_rows[2];      // rows per group
  int n_qb[2];        // 64-query blocks per group
  int p_chunk;        // prompt keys per prefix split (multiple of TK)
  int n_psplit, n_csplit, NS;
  int n_prefix_items, n_items;
};

Plan make_plan(int R, int rows_group0, int P, int c_max, int nh, int nkv, int sms) {
  Plan pl;
  pl.rep = nh / nkv;
  pl.n_rows[0] = rows_group0 < R ? rows_group0 : R;
  pl.n_rows[1] = R - pl.n_rows[0];
  for (int g = 0; g < 2; ++g) pl.n_qb[g] = (pl.n_rows[g] * pl.rep + TQ - 1) / TQ;
  const int base = (pl.n_qb[0] + pl.n_qb[1]) * nkv;
  int target = sms / (base > 0 ? base : 1);
  if (target < 1) target = 1;
  int chunk = (P + target - 1) / target;
  chunk = (chunk + TK - 1) / TK * TK;
  if (chunk < 2 * TK) chunk = 2 * TK;
  pl.p_chunk = chunk;
  pl.n_psplit = P > 0 ? (P + chunk - 1) / chunk : 0;
  int cs = (c_max + 127) / 128;
  cs = cs < 1 ? 1 : (cs > 8 ? 8 : cs);
  // two CTAs fit per SM (87 KB of shared memory each): keep the whole launch in ONE wave when possible -- a 40-CTA
  // second wave showed up as a 3 us gap before the combine kernel (profiles/r01_decode_trace_unfused.txt)
  const int room = 2 * sms - base * (P > 0 ? (P + chunk - 1) / chunk : 0);
  const int per_split = R * nkv;
  if (per_split > 0 && room >= per_split && cs > room / per_split) cs = room / per_split;
  pl.n_csplit = cs;
  pl.NS = pl.n_psplit + pl.n_csplit;
  pl.n_prefix_items = base * pl.n_psplit;
  pl.n_items = pl.n_prefix_items + R * nkv * pl.n_csplit;
  return pl;
}

struct DecAttnParams {
  const bf16* q;                       // [R][nh*HD]
  const bf16* kp[2]; const bf16* vp[2];  // shared prompt caches [P][nkv*HD] per group
  const bf16* kc; const bf16* vc;      // completion caches [R][c_max][nkv*HD]
  long long cache_stride_r;
  const int* step_ptr;
  int R, P, nh, nkv;
  float scale_log2;
  Plan pl;
  float* o_part;                       // [R][nh][NS][HD]
  float* lse_part;                     // [R][nh][NS]   (log2 domain; -inf = empty)
};

__global__ void __launch_bounds__(THREADS)
dec_attn_kernel(const DecAttnParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* sQ = reinterpret_cast<bf16*>(smem_raw);
  bf16* sK = sQ + TILE;        // 2 stages
  bf16* sV = sK + 2 * TILE;    // 2 stages
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const Plan& pl = p.pl;
  const int rep = pl.rep;

  pdl_launch_dependents();
  const int tr = (blockIdx.x == 0 && tid == 0) ? sb_trace_begin(SB_TR_ATTN) : -1;
  // ---- decode the item
  int item = blockIdx.x;
  const bf16 *kbase, *vbase;
  int j_lo, j_hi, slot, kvh, row0, n_q;   // query vector m -> (row0 + m / rep, head kvh*rep + m % rep), m < n_q
  int m0 = 0;                             // first query vector of this q-block
  if (item < pl.n_prefix_items) {
    const int per_group0 = pl.n_qb[0] * p.nkv * pl.n_psplit;
    int grp = 0;
    if (item >= per_group0) { grp = 1; item -= per_group0; }
    const int s = item % pl.n_psplit;
    kvh = (item / pl.n_psplit) % p.nkv;
    const int qb = item / (pl.n_psplit * p.nkv);
    row0 = grp == 0 ? 0 : pl.n_rows[0];
    m0 = qb * TQ;
    n_q = min(TQ, pl.n_rows[grp] * rep - m0);
    kbase = p.kp[grp] + (long long)kvh * HD;
    vbase = p.vp[grp] + (long long)kvh * HD;
    j_lo = s * pl.p_chunk;
    j_hi = min(j_lo + pl.p_chunk, p.P);
    slot = s;
  } else {
    pdl_wait();                              // the step counter and the completion cache come from earlier kernels
    item -= pl.n_prefix_items;
    const int s = item % pl.n_csplit;
    kvh = (item / pl.n_csplit) % p.nkv;
    row0 = item / (pl.n_csplit * p.nkv);
    n_q = rep;
    kbase = p.kc + row0 * p.cache_stride_r + (long long)kvh * HD;
    vbase = p.vc + row0 * p.cache_stride_r + (long long)kvh * HD;
    const int n_ctx = *p.step_ptr + 1;
    int per = (n_ctx + pl.n_csplit - 1) / pl.n_csplit;
    per = (per + TK - 1) / TK * TK;
    j_lo = s * per;
    j_hi = min(j_lo + per, n_ctx);
    slot = pl.n_psplit + s;
  }
  const bool is_prefix = blockIdx.x < pl.n_prefix_items;
  const long long kv_ld = (long long)p.nkv * HD;
  const int n_tiles = j_hi > j_lo ? (j_hi - j_lo + TK - 1) / TK : 0;

  auto load_kv = [&](int t, int buf) {
    const int k0 = j_lo + t * TK;
    for (int i = tid; i < TK * (HD / 8); i += THREADS) {
      const int r = i / (HD / 8), c = i % (HD / 8);
      const bool ok = (k0 + r) < j_hi;
      const long long off = (long long)(ok ? (k0 + r) : j_lo) * kv_ld + c * 8;
      cp_async16(smem_u32(sK + buf * TILE + r * LD + c * 8), kbase + off, ok);
      cp_async16(smem_u32(sV + buf * TILE + r * LD + c * 8), vbase + off, ok);
    }
  };

  // ---- first K/V tile + Q tile (gathered query vectors).  The prompt cache is constant during decode, so prefix
  // items request their first K/V tile before waiting for the kernels that produce q.
  if (is_prefix) {
    if (n_tiles > 0) load_kv(0, 0);
    pdl_wait();
  }
  sb_trace_mark(tr, 1);
  for (int i = tid; i < TQ * (HD / 8); i += THREADS) {
    const int m = i / (HD / 8), c = i % (HD / 8);
    const bool ok = m < n_q;
    const int mm = ok ? (m0 + m) : m0;
    const bf16* src = p.q + (long long)(row0 + mm / rep) * p.nh * HD + (long long)(kvh * rep + mm % rep) * HD + c * 8;
    cp_async16(smem_u32(sQ + m * LD + c * 8), src, ok);
  }
  if (!is_prefix && n_tiles > 0) load_kv(0, 0);
  cp_async_commit();

  const bool active = warp * 16 < n_q;   // warp-uniform: this warp owns at least one real query vector
  uint32_t qf[HD / 16][4];
  float o_acc[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) { o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f; }
  float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;

  for (int t = 0; t < n_tiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < n_tiles) load_kv(t + 1, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (active) {
      if (t == 0) {
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks)
          ldsm_x4(smem_u32(sQ + (warp * 16 + (lane & 15)) * LD + ks * 16 + (lane >> 4) * 8), qf[ks][0], qf[ks][1],
                  qf[ks][2], qf[ks][3]);
      }
      const bf16* cK = sK + buf * TILE;
      const bf16* cV = sV + buf * TILE;
      float s[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < HD / 16; ++ks) {
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4(smem_u32(cK + (np * 16 + (lane & 7) + (lane >> 4) * 8) * LD + ks * 16 + ((lane >> 3) & 1) * 8), b0, b1,
                  b2, b3);
          mma16816(s[np * 2], qf[ks], b0, b1);
          mma16816(s[np * 2 + 1], qf[ks], b2, b3);
        }
      }
      const int k0 = j_lo + t * TK;
      const bool full = k0 + TK <= j_hi;
      float mx_lo = -INFINITY, mx_hi = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float v = s[nt][e] * p.scale_log2;
          if (!full && (k0 + nt * 8 + t4 * 2 + (e & 1)) >= j_hi) v = -INFINITY;
          s[nt][e] = v;
          if (e < 2) mx_lo = fmaxf(mx_lo, v); else mx_hi = fmaxf(mx_hi, v);
        }
      }
      mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
      mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
      mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
      mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
      const float nm_lo = fmaxf(m_lo, mx_lo), nm_hi = fmaxf(m_hi, mx_hi);   // finite: every tile has >= 1 valid key
      const float cr_lo = exp2f(m_lo - nm_lo), cr_hi = exp2f(m_hi - nm_hi);
      m_lo = nm_lo; m_hi = nm_hi;
      float rs_lo = 0.f, rs_hi = 0.f;
      uint32_t pf[4][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float p0 = exp2f(s[nt][0] - nm_lo), p1 = exp2f(s[nt][1] - nm_lo);
        const float p2 = exp2f(s[nt][2] - nm_hi), p3 = exp2f(s[nt][3] - nm_hi);
        rs_lo += p0 + p1; rs_hi += p2 + p3;
        pf[nt >> 1][(nt & 1) * 2] = pack_bf16(p0, p1);
        pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(p2, p3);
      }
      l_lo = l_lo * cr_lo + rs_lo;
      l_hi = l_hi * cr_hi + rs_hi;
#pragma unroll
      for (int i = 0; i < HD / 8; ++i) {
        o_acc[i][0] *= cr_lo; o_acc[i][1] *= cr_lo; o_acc[i][2] *= cr_hi; o_acc[i][3] *= cr_hi;
      }
#pragma unroll
      for (int kk = 0; kk < TK / 16; ++kk) {
#pragma unroll
        for (int dp = 0; dp < HD / 16; ++dp) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4_t(smem_u32(cV + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + dp * 16 + (lane >> 4) * 8), b0, b1,
                    b2, b3);
          mma16816(o_acc[dp * 2], pf[kk], b0, b1);
          mma16816(o_acc[dp * 2 + 1], pf[kk], b2, b3);
        }
      }
    }
    __syncthreads();
  }
  cp_async_wait<0>();
  if (!active) return;

  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int m = warp * 16 + g + half * 8;
    if (m >= n_q) continue;
    const int mm = m0 + m;
    const int row = row0 + mm / rep, head = kvh * rep + mm % rep;
    const float l = half ? l_hi : l_lo, mxv = half ? m_hi : m_lo;
    const float inv = l > 0.f ? 1.f / l : 0.f;
    const long long pbase = ((long long)row * p.nh + head) * pl.NS + slot;
    float* op = p.o_part + pbase * HD;
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
      const float2 v = half ? make_float2(o_acc[i][2] * inv, o_acc[i][3] * inv)
                            : make_float2(o_acc[i][0] * inv, o_acc[i][1] * inv);
      *reinterpret_cast<float2*>(op + i * 8 + t4 * 2) = v;
    }
    if (t4 == 0) p.lse_part[pbase] = l > 0.f ? mxv + log2f(l) : -INFINITY;
  }
  sb_trace_mark(tr, 2);
}

// out[row][head*HD + d] = sum_s w_s O_s / sum_s w_s, w_s = 2^(lse_s - max lse); one warp per (row, head)
__global__ void __launch_bounds__(128)
dec_attn_combine_kernel(const float* __restrict__ o_part, const float* __restrict__ lse_part, int NS, int n_pairs,
                        bf16* __restrict__ out) {
  pdl_launch_dependents();
  const int tr = (blockIdx.x == 0 && threadIdx.x == 0) ? sb_trace_begin(SB_TR_COMBINE) : -1;
  pdl_wait();
  sb_trace_mark(tr, 1);
  const int pair = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (pair >= n_pairs) return;
  const int lane = threadIdx.x & 31;
  const float* lp = lse_part + (long long)pair * NS;
  const float* op = o_part + (long long)pair * NS * HD + lane * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float wsum = 0.f;
  if (NS <= 32) {
    // one round trip to L2: the log-sum-exps and ALL partial outputs are requested before anything is consumed
    const float ls = lane < NS ? lp[lane] : -INFINITY;
    float4 v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < NS) v[j] = *reinterpret_cast<const float4*>(op + (long long)j * HD);
    const float mx = warp_max(ls);
    const float w = ls > -INFINITY ? exp2f(ls - mx) : 0.f;   // mx == -inf only if every partial is empty (w = 0)
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float wj = __shfl_sync(0xffffffffu, w, j);
      if (j < NS && wj != 0.f) {   // empty partials may hold stale bits
        acc.x += wj * v[j].x; acc.y += wj * v[j].y; acc.z += wj * v[j].z; acc.w += wj * v[j].w;
        wsum += wj;
      }
    }
  } else {
    float mx = -INFINITY;
    for (int s = lane; s < NS; s += 32) mx = fmaxf(mx, lp[s]);
    mx = warp_max(mx);
    if (mx > -INFINITY) {
      // batches of 8 splits: all loads of a batch are issued before any is consumed (the partials sit in L2)
      for (int s0 = 0; s0 < NS; s0 += 8) {
        float4 v[8];
        float w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int s = min(s0 + j, NS - 1);
          v[j] = *reinterpret_cast<const float4*>(op + (long long)s * HD);
          const float ls = lp[s];
          w[j] = (s0 + j < NS && ls > -INFINITY) ? exp2f(ls - mx) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (w[j] != 0.f) {   // empty partials may hold stale bits
            acc.x += w[j] * v[j].x; acc.y += w[j] * v[j].y; acc.z += w[j] * v[j].z; acc.w += w[j] * v[j].w;
            wsum += w[j];
          }
        }
      }
    }
  }
  const float inv = wsum > 0.f ? 1.f / wsum : 0.f;
  uint2 u;
  u.x = pack_bf16(acc.x * inv, acc.y * inv);
  u.y = pack_bf16(acc.z * inv, acc.w * inv);
  *reinterpret_cast<uint2*>(out + (long long)pair * HD + lane * 4) = u;
  sb_trace_mark(tr, 2);
}

int g_sms = 0;
int sm_count() {
  if (g_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_sms <= 0) g_sms = 148;
  }
  return g_sms;
}

constexpr int DEC_ATTN_SMEM = 5 * TILE * 2;

}  // namespace

SB_DEFINE_TRACE_SETTER(sb_trace_set_dec_attn)

extern "C" int sb_dec_attn_workspace(int R, int rows_group0, int P, int c_max, int n_heads, int n_kv_heads,
                                     long long* floats_out) {
  SB_REQUIRE(floats_out && R > 0 && P >= 0 && c_max > 0 && n_heads > 0 && n_kv_heads > 0 && n_heads % n_kv_heads == 0,
             "sb_dec_attn_workspace: bad arguments");
  const Plan pl = make_plan(R, rows_group0, P, c_max, n_heads, n_kv_heads, sm_count());
  *floats_out = (long long)R * n_heads * pl.NS * (HD + 1);
  return 0;
}

extern "C" int sb_dec_attn(const void* q, const void* kp0, const void* vp0, const void* kp1, const void* vp1,
                           int rows_group0, int P, const void* k_cache, const void* v_cache, long long cache_stride_r,
                           int c_max, const int* step_ptr, int n_heads, int n_kv_heads, int head_dim, float scale,
                           float* workspace, long long workspace_floats, void* out, int R, sb_stream_t stream) {
  SB_REQUIRE(q && k_cache && v_cache && step_ptr && workspace && out, "sb_dec_attn: null pointer");
  SB_REQUIRE(P == 0 || (kp0 && vp0), "sb_dec_attn: prompt cache missing");
  SB_REQUIRE(head_dim == HD, "sb_dec_attn: head_dim must be 128, got %d", head_dim);
  SB_REQUIRE(R > 0 && R <= 32 && c_max > 0 && n_heads % n_kv_heads == 0, "sb_dec_attn: bad sizes");
  SB_REQUIRE(n_heads / n_kv_heads <= 16, "sb_dec_attn: at most 16 q heads per kv head");
  SB_REQUIRE(rows_group0 >= R || (kp1 && vp1), "sb_dec_attn: second prompt cache missing for rows >= rows_group0");
  DecAttnParams p;
  p.q = (const bf16*)q;
  p.kp[0] = (const bf16*)kp0; p.vp[0] = (const bf16*)vp0;
  p.kp[1] = (const bf16*)(kp1 ? kp1 : kp0); p.vp[1] = (const bf16*)(vp1 ? vp1 : vp0);
  p.kc = (const bf16*)k_cache; p.vc = (const bf16*)v_cache; p.cache_stride_r = cache_stride_r;
  p.step_ptr = step_ptr; p.R = R; p.P = P; p.nh = n_heads; p.nkv = n_kv_heads;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.pl = make_plan(R, rows_group0, P, c_max, n_heads, n_kv_heads, sm_count());
  const long long need = (long long)R * n_heads * p.pl.NS * (HD + 1);
  SB_REQUIRE(workspace_floats >= need, "sb_dec_attn: workspace too small (%lld floats, need %lld; see sb_dec_attn_workspace)",
             workspace_floats, need);
  p.o_part = workspace;
  p.lse_part = workspace + (long long)R * n_heads * p.pl.NS * HD;
  static bool done = false;
  if (!done) {
    SB_CUDA(cudaFuncSetAttribute(dec_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DEC_ATTN_SMEM));
    done = true;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  SB_CUDA(sb_launch(dec_attn_kernel, dim3(p.pl.n_items), dim3(THREADS), (size_t)DEC_ATTN_SMEM, st, sb_pdl_enabled(), p));
  if (sb_check_launch("sb_dec_attn")) return 1;
  const int n_pairs = R * n_heads;
  SB_CUDA(sb_launch(dec_attn_combine_kernel, dim3((n_pairs + 3) / 4), dim3(128), 0, st, sb_pdl_enabled(),
                    (const float*)p.o_part, (const float*)p.lse_part, p.pl.NS, n_pairs, (bf16*)out));
  return sb_check_launch("sb_dec_attn(combine)");
}
