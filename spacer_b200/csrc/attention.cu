// Attention for the SG-RLVR hot path: one forward and one backward kernel cover
//   * the ViT's block-diagonal (per temporal slab) non-causal attention, head_dim 80   (MQ2:415-454, cu_seqlens :772-780)
//   * Qwen2 causal GQA attention over (prompt + completion), head_dim 128              (MQ2:575-590)
//   * the prefix-shared training layout [prompt | completion_0 | ... | completion_{G-1}] in which every
//     completion token sees the whole shared prompt plus its own completion causally.
// Visibility is described per query token by meta[t] = (prefix_len, seg_start, kv_end):
//     key j is visible to query t  <=>  j < prefix_len  ||  (seg_start <= j < kv_end).
// Round-1 implementation: flash-attention-2 style online softmax on mma.sync.m16n8k16 (bf16, fp32 accum),
// cp.async double-buffered K/V tiles.  (The tcgen05/TMEM version is the next step; the GEMMs already are.)
#include "common.cuh"
#include "spacer_b200.h"

// attention_bwd_tc.cu
int sb_attn_bwd_dkdv_tc(const sb_attn_args* a, const int* qtb, void* dk_out, void* dv_out, long long ld_dk,
                        long long ld_dv, cudaStream_t st);
int sb_attn_bwd_dq_tc(const sb_attn_args* a, cudaStream_t st);

namespace {

constexpr int ATT_THREADS = 128;  // 4 warps x 16 query rows
constexpr int BQ = 64;
constexpr int BKV = 64;

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

struct AttnParams {
  const bf16* q; const bf16* k; const bf16* v;   // row-major, head h at column h*HD (kv head for k,v)
  long long ldq, ldk, ldv;
  bf16* o; long long ldo;
  float* lse;                 // [n_heads][T]
  const int4* meta;           // [T] (prefix_len, seg_start, kv_end, _)
  int T, Tk, n_heads, n_kv_heads;
  float scale;
  // backward
  const bf16* d_o; long long lddo;
  const float* delta;         // [n_heads][T]
  bf16* dk; bf16* dv; long long lddk, lddv;
};

template <int HD>
struct Smem {
  static constexpr int LD = HD + 8;  // padded row (elements): conflict-free ldmatrix
  static constexpr int TILE = 64 * LD;
};

// load a [64 x HD] bf16 tile (rows r0.., zero-filled past `rows`) into padded smem
template <int HD>
SB_DEVICE void load_tile(uint32_t sdst, const bf16* g, long long ld, int r0, int rows, int tid) {
  constexpr int CH = HD / 8;  // 16B chunks per row
  for (int i = tid; i < 64 * CH; i += ATT_THREADS) {
    const int r = i / CH, c = i % CH;
    const bool ok = (r0 + r) < rows;
    const bf16* src = g + (long long)(ok ? (r0 + r) : 0) * ld + c * 8;
    cp_async16(sdst + (uint32_t)(r * Smem<HD>::LD + c * 8) * 2, src, ok);
  }
}

struct TileBounds { int pmin, pmax, smin, smax, emin, emax; };

SB_DEVICE bool kv_tile_relevant(const TileBounds& b, int j0) {
  return (j0 < b.pmax) || (j0 + BKV > b.smin && j0 < b.emax);
}
SB_DEVICE bool kv_tile_full(const TileBounds& b, int j0) {
  return (j0 + BKV <= b.pmin) || (j0 >= b.smax && j0 + BKV <= b.emin);
}
SB_DEVICE int next_kv_tile(const TileBounds& b, int jt, int n_tiles) {
  while (jt < n_tiles && !kv_tile_relevant(b, jt * BKV)) ++jt;
  return jt;
}

SB_DEVICE TileBounds tile_bounds(const int4* meta, int q0, int T, int* sh /*3*64 ints*/, int tid) {
  if (tid < 64) {
    int4 m = make_int4(0, 0, 0, 0);
    if (q0 + tid < T) m = meta[q0 + tid];
    sh[tid] = m.x; sh[64 + tid] = m.y; sh[128 + tid] = m.z;
  }
  __syncthreads();
  TileBounds b;
  b.pmin = 1 << 30; b.pmax = 0; b.smin = 1 << 30; b.smax = 0; b.emin = 1 << 30; b.emax = 0;
  for (int i = 0; i < 64; ++i) {
    if (q0 + i >= T) break;
    const int p = sh[i], s = sh[64 + i], e = sh[128 + i];
    b.pmin = min(b.pmin, p); b.pmax = max(b.pmax, p);
    if (e > s) {  // row has an own-segment range
      b.smin = min(b.smin, s); b.emax = max(b.emax, e);
    }
    b.smax = max(b.smax, s); b.emin = min(b.emin, e);
  }
  return b;
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(ATT_THREADS)
attn_fwd_kernel(const AttnParams p) {
  using S = Smem<HD>;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* sQ = reinterpret_cast<bf16*>(smem_raw);
  bf16* sK = sQ + S::TILE;          // 2 buffers
  bf16* sV = sK + 2 * S::TILE;      // 2 buffers
  int* sMeta = reinterpret_cast<int*>(sV + 2 * S::TILE);  // 192 ints

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int q0 = blockIdx.x * BQ;
  const int head = blockIdx.y;
  const int kvh = head / (p.n_heads / p.n_kv_heads);
  const bf16* qg = p.q + (long long)head * HD;
  const bf16* kg = p.k + (long long)kvh * HD;
  const bf16* vg = p.v + (long long)kvh * HD;

  const TileBounds tb = tile_bounds(p.meta, q0, p.T, sMeta, tid);
  const int r_lo = warp * 16 + g, r_hi = r_lo + 8;
  const int pre_lo = sMeta[r_lo], st_lo = sMeta[64 + r_lo], en_lo = sMeta[128 + r_lo];
  const int pre_hi = sMeta[r_hi], st_hi = sMeta[64 + r_hi], en_hi = sMeta[128 + r_hi];

  const int n_tiles = (min(max(tb.pmax, tb.emax), p.Tk) + BKV - 1) / BKV;

  load_tile<HD>(smem_u32(sQ), qg, p.ldq, q0, p.T, tid);
  int jt = next_kv_tile(tb, 0, n_tiles);
  if (jt < n_tiles) {
    load_tile<HD>(smem_u32(sK), kg, p.ldk, jt * BKV, p.Tk, tid);
    load_tile<HD>(smem_u32(sV), vg, p.ldv, jt * BKV, p.Tk, tid);
  }
  cp_async_commit();

  constexpr int KS = HD / 16;   // k-steps over the head dim
  constexpr int NO = HD / 8;    // output n-tiles
  uint32_t qf[KS][4];
  float o_acc[NO][4];
#pragma unroll
  for (int i = 0; i < NO; ++i) { o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f; }
  float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;
  const float sc = p.scale * 1.4426950408889634f;

  int buf = 0;
  bool q_loaded = false;
  while (jt < n_tiles) {
    const int jn = next_kv_tile(tb, jt + 1, n_tiles);
    if (jn < n_tiles) {
      load_tile<HD>(smem_u32(sK + (buf ^ 1) * S::TILE), kg, p.ldk, jn * BKV, p.Tk, tid);
      load_tile<HD>(smem_u32(sV + (buf ^ 1) * S::TILE), vg, p.ldv, jn * BKV, p.Tk, tid);
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (!q_loaded) {
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const uint32_t a = smem_u32(sQ + (warp * 16 + (lane & 15)) * S::LD + ks * 16 + (lane >> 4) * 8);
        ldsm_x4(a, qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
      }
      q_loaded = true;
    }
    const bf16* cK = sK + buf * S::TILE;
    const bf16* cV = sV + buf * S::TILE;
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {  // pairs of key n-tiles
        uint32_t b0, b1, b2, b3;
        const uint32_t a = smem_u32(cK + (np * 16 + (lane & 7) + (lane >> 4) * 8) * S::LD + ks * 16 + ((lane >> 3) & 1) * 8);
        ldsm_x4(a, b0, b1, b2, b3);
        mma16816(s[np * 2], qf[ks], b0, b1);
        mma16816(s[np * 2 + 1], qf[ks], b2, b3);
      }
    }
    const int j0 = jt * BKV;
    const bool full = kv_tile_full(tb, j0) && (j0 + BKV <= p.Tk);
    float mx_lo = -INFINITY, mx_hi = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float v = s[nt][e] * sc;
        if (!full) {
          const int j = j0 + nt * 8 + t4 * 2 + (e & 1);
          const bool hi = e >= 2;
          const int pre = hi ? pre_hi : pre_lo, st = hi ? st_hi : st_lo, en = hi ? en_hi : en_lo;
          const bool vis = (j < p.Tk) && ((j < pre) || (j >= st && j < en));
          if (!vis) v = -INFINITY;
        }
        s[nt][e] = v;
        if (e < 2) mx_lo = fmaxf(mx_lo, v); else mx_hi = fmaxf(mx_hi, v);
      }
    }
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
    const float nm_lo = fmaxf(m_lo, mx_lo), nm_hi = fmaxf(m_hi, mx_hi);
    const float ms_lo = nm_lo == -INFINITY ? 0.f : nm_lo, ms_hi = nm_hi == -INFINITY ? 0.f : nm_hi;
    const float cr_lo = exp2f(m_lo - ms_lo), cr_hi = exp2f(m_hi - ms_hi);
    m_lo = nm_lo; m_hi = nm_hi;
    float rs_lo = 0.f, rs_hi = 0.f;
    uint32_t pf[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = exp2f(s[nt][0] - ms_lo), p1 = exp2f(s[nt][1] - ms_lo);
      const float p2 = exp2f(s[nt][2] - ms_hi), p3 = exp2f(s[nt][3] - ms_hi);
      rs_lo += p0 + p1; rs_hi += p2 + p3;
      pf[nt >> 1][(nt & 1) * 2] = pack_bf16(p0, p1);
      pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(p2, p3);
    }
    l_lo = l_lo * cr_lo + rs_lo;
    l_hi = l_hi * cr_hi + rs_hi;
#pragma unroll
    for (int i = 0; i < NO; ++i) {
      o_acc[i][0] *= cr_lo; o_acc[i][1] *= cr_lo; o_acc[i][2] *= cr_hi; o_acc[i][3] *= cr_hi;
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {       // 16 keys per step
#pragma unroll
      for (int dp = 0; dp < NO / 2; ++dp) { // pairs of output n-tiles
        uint32_t b0, b1, b2, b3;
        const uint32_t a = smem_u32(cV + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * S::LD + dp * 16 + (lane >> 4) * 8);
        ldsm_x4_t(a, b0, b1, b2, b3);
        mma16816(o_acc[dp * 2], pf[kk], b0, b1);
        mma16816(o_acc[dp * 2 + 1], pf[kk], b2, b3);
      }
    }
    __syncthreads();
    buf ^= 1;
    jt = jn;
  }
  cp_async_wait<0>();

  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
  const float inv_lo = l_lo > 0.f ? 1.f / l_lo : 0.f, inv_hi = l_hi > 0.f ? 1.f / l_hi : 0.f;
  const int row_lo = q0 + r_lo, row_hi = q0 + r_hi;
#pragma unroll
  for (int i = 0; i < NO; ++i) {
    const int col = head * HD + i * 8 + t4 * 2;
    if (row_lo < p.T)
      *reinterpret_cast<uint32_t*>(p.o + (long long)row_lo * p.ldo + col) = pack_bf16(o_acc[i][0] * inv_lo, o_acc[i][1] * inv_lo);
    if (row_hi < p.T)
      *reinterpret_cast<uint32_t*>(p.o + (long long)row_hi * p.ldo + col) = pack_bf16(o_acc[i][2] * inv_hi, o_acc[i][3] * inv_hi);
  }
  if (p.lse && t4 == 0) {
    if (row_lo < p.T) p.lse[(long long)head * p.T + row_lo] = l_lo > 0.f ? (m_lo + log2f(l_lo)) * 0.6931471805599453f : -INFINITY;
    if (row_hi < p.T) p.lse[(long long)head * p.T + row_hi] = l_hi > 0.f ? (m_hi + log2f(l_hi)) * 0.6931471805599453f : -INFINITY;
  }
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
// delta[h][t] = sum_d dO[t][h][d] * O[t][h][d]
template <int HD>
__global__ void attn_delta_kernel(const bf16* __restrict__ o, long long ldo, const bf16* __restrict__ d_o,
                                  long long lddo, float* __restrict__ delta, int T, int n_heads) {
  const int warps_per_cta = blockDim.x >> 5;
  const long long item = (long long)blockIdx.x * warps_per_cta + (threadIdx.x >> 5);
  if (item >= (long long)T * n_heads) return;
  const int h = item % n_heads;
  const int t = item / n_heads;
  const int lane = threadIdx.x & 31;
  float s = 0.f;
  for (int d = lane * 2; d < HD; d += 64) {
    const float2 a = unpack_bf16(*reinterpret_cast<const uint32_t*>(o + (long long)t * ldo + h * HD + d));
    const float2 b = unpack_bf16(*reinterpret_cast<const uint32_t*>(d_o + (long long)t * lddo + h * HD + d));
    s += a.x * b.x + a.y * b.y;
  }
  s = warp_sum(s);
  if (lane == 0) delta[(long long)h * T + t] = s;
}

// Backward in two kernels (no atomics, deterministic):
//   attn_bwd_dkdv_kernel : one CTA per (64-key tile, q head); loops over the 64-query tiles that can see the keys.
//                          Each warp owns 16 keys:  S^T = K Q^T,  P^T = exp(S^T - lse),  dV += P^T dO,
//                          dP^T = V dO^T,  dS^T = P^T o (dP^T - delta),  dK += dS^T Q.
//                          With GQA the per-q-head dK/dV go to an expanded [T][n_heads*HD] scratch and are summed
//                          over the group by attn_gqa_reduce_kernel.
//   attn_bwd_dq_kernel   : one CTA per (64-query tile, q head); loops over the visible key tiles like the forward:
//                          S = Q K^T, P, dP = dO V^T, dS, dQ += dS K  (S and dP are recomputed: +2 GEMMs instead of
//                          64x128 fp32 atomics per tile pair).
// Visibility of a (query tile, key tile) pair is decided from per-query-tile bounds computed once per call.
struct QTileBounds { int pmin, pmax, smin, smax, emin, emax, pad0, pad1; };

__global__ void attn_qtile_bounds_kernel(const int4* __restrict__ meta, int T, QTileBounds* __restrict__ out) {
  const int qt = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int q0 = qt * BQ;
  if (q0 >= T) return;
  const int lane = threadIdx.x & 31;
  int pmin = 1 << 30, pmax = 0, smin = 1 << 30, smax = 0, emin = 1 << 30, emax = 0;
  for (int i = lane; i < BQ && q0 + i < T; i += 32) {
    const int4 m = meta[q0 + i];
    pmin = min(pmin, m.x); pmax = max(pmax, m.x);
    if (m.z > m.y) { smin = min(smin, m.y); emax = max(emax, m.z); }
    smax = max(smax, m.y); emin = min(emin, m.z);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    pmin = min(pmin, __shfl_xor_sync(0xffffffffu, pmin, o)); pmax = max(pmax, __shfl_xor_sync(0xffffffffu, pmax, o));
    smin = min(smin, __shfl_xor_sync(0xffffffffu, smin, o)); smax = max(smax, __shfl_xor_sync(0xffffffffu, smax, o));
    emin = min(emin, __shfl_xor_sync(0xffffffffu, emin, o)); emax = max(emax, __shfl_xor_sync(0xffffffffu, emax, o));
  }
  if (lane == 0) {
    QTileBounds b;
    b.pmin = pmin; b.pmax = pmax; b.smin = smin; b.smax = smax; b.emin = emin; b.emax = emax; b.pad0 = b.pad1 = 0;
    out[qt] = b;
  }
}

template <int HD>
__global__ void __launch_bounds__(ATT_THREADS)
attn_bwd_dkdv_kernel(const AttnParams p, const QTileBounds* __restrict__ qtb, bf16* __restrict__ dk_out,
                     bf16* __restrict__ dv_out, long long ld_dk, long long ld_dv) {
  using S = Smem<HD>;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* sK = reinterpret_cast<bf16*>(smem_raw);          // [64][LD]
  bf16* sV = sK + S::TILE;                               // [64][LD]
  bf16* sQ = sV + S::TILE;                               // 2 x [64][LD]
  bf16* sdO = sQ + 2 * S::TILE;                          // 2 x [64][LD]
  float* sLse = reinterpret_cast<float*>(sdO + 2 * S::TILE);  // 2 x 64 (log2 domain)
  float* sDelta = sLse + 2 * BQ;                          // 2 x 64
  int* sMeta = reinterpret_cast<int*>(sDelta + 2 * BQ);   // 2 x 3 x 64

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int j0 = blockIdx.x * BKV;
  const int hh = blockIdx.y;
  const int kvh = hh / (p.n_heads / p.n_kv_heads);
  const float sc2 = p.scale * 1.4426950408889634f;

  load_tile<HD>(smem_u32(sK), p.k + (long long)kvh * HD, p.ldk, j0, p.Tk, tid);
  load_tile<HD>(smem_u32(sV), p.v + (long long)kvh * HD, p.ldv, j0, p.Tk, tid);
  cp_async_commit();

  constexpr int KS = HD / 16;
  constexpr int ND = HD / 8;
  float dk_acc[ND][4], dv_acc[ND][4];
#pragma unroll
  for (int i = 0; i < ND; ++i) {
    dk_acc[i][0] = dk_acc[i][1] = dk_acc[i][2] = dk_acc[i][3] = 0.f;
    dv_acc[i][0] = dv_acc[i][1] = dv_acc[i][2] = dv_acc[i][3] = 0.f;
  }
  const int n_qt = (p.T + BQ - 1) / BQ;

  auto issue = [&](int qt, int b) {
    const int q0 = qt * BQ;
    load_tile<HD>(smem_u32(sQ + b * S::TILE), p.q + (long long)hh * HD, p.ldq, q0, p.T, tid);
    load_tile<HD>(smem_u32(sdO + b * S::TILE), p.d_o + (long long)hh * HD, p.lddo, q0, p.T, tid);
    if (tid < BQ) {
      const bool ok = (q0 + tid) < p.T;
      sLse[b * BQ + tid] = ok ? p.lse[(long long)hh * p.T + q0 + tid] * 1.4426950408889634f : 0.f;
      sDelta[b * BQ + tid] = ok ? p.delta[(long long)hh * p.T + q0 + tid] : 0.f;
      int4 m = make_int4(0, 0, 0, 0);
      if (ok) m = p.meta[q0 + tid];
      sMeta[b * 192 + tid] = m.x; sMeta[b * 192 + 64 + tid] = m.y; sMeta[b * 192 + 128 + tid] = m.z;
    }
  };
  auto next_qt = [&](int qt) -> int {
    while (qt < n_qt) {
      const QTileBounds b = qtb[qt];
      if ((j0 < b.pmax) || (j0 + BKV > b.smin && j0 < b.emax)) return qt;
      ++qt;
    }
    return n_qt;
  };

  int qt = next_qt(0);
  int buf = 0;
  if (qt < n_qt) issue(qt, 0);
  cp_async_commit();

  while (qt < n_qt) {
    const int qn = next_qt(qt + 1);
    if (qn < n_qt) issue(qn, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();

    const int q0 = qt * BQ;
    const bf16* cQ = sQ + buf * S::TILE;
    const bf16* cdO = sdO + buf * S::TILE;
    const float* cL = sLse + buf * BQ;
    const float* cD = sDelta + buf * BQ;
    const int* cM = sMeta + buf * 192;
    const QTileBounds tb = qtb[qt];
    const bool full = ((j0 + BKV <= tb.pmin) || (j0 >= tb.smax && j0 + BKV <= tb.emin)) && (j0 + BKV <= p.Tk) &&
                      (q0 + BQ <= p.T);

#pragma unroll 1
    for (int hf = 0; hf < 2; ++hf) {   // two halves of 32 queries: keeps S^T / dP^T at 16 registers each
      const int qh = hf * 32;
      float st[4][4], dp[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f;
        dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
      }
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t ka[4], va[4];
        const int arow = warp * 16 + (lane & 15), acol = ks * 16 + (lane >> 4) * 8;
        ldsm_x4(smem_u32(sK + arow * S::LD + acol), ka[0], ka[1], ka[2], ka[3]);
        ldsm_x4(smem_u32(sV + arow * S::LD + acol), va[0], va[1], va[2], va[3]);
#pragma unroll
        for (int np = 0; np < 2; ++np) {
          uint32_t b0, b1, b2, b3;
          const int brow = qh + np * 16 + (lane & 7) + (lane >> 4) * 8, bcol = ks * 16 + ((lane >> 3) & 1) * 8;
          ldsm_x4(smem_u32(cQ + brow * S::LD + bcol), b0, b1, b2, b3);
          mma16816(st[np * 2], ka, b0, b1);
          mma16816(st[np * 2 + 1], ka, b2, b3);
          ldsm_x4(smem_u32(cdO + brow * S::LD + bcol), b0, b1, b2, b3);
          mma16816(dp[np * 2], va, b0, b1);
          mma16816(dp[np * 2 + 1], va, b2, b3);
        }
      }
      uint32_t pf[2][4], dsf[2][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        float pv[4], dsv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int ql = qh + nt * 8 + t4 * 2 + (e & 1);
          bool vis = true;
          if (!full) {
            const int key = j0 + warp * 16 + g + (e >= 2 ? 8 : 0);
            const int pre = cM[ql], s0 = cM[64 + ql], en = cM[128 + ql];
            vis = (q0 + ql < p.T) && (key < p.Tk) && ((key < pre) || (key >= s0 && key < en));
          }
          const float pr = vis ? exp2f(st[nt][e] * sc2 - cL[ql]) : 0.f;
          pv[e] = pr;
          dsv[e] = pr * (dp[nt][e] - cD[ql]);
        }
        pf[nt >> 1][(nt & 1) * 2] = pack_bf16(pv[0], pv[1]);
        pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(pv[2], pv[3]);
        dsf[nt >> 1][(nt & 1) * 2] = pack_bf16(dsv[0], dsv[1]);
        dsf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(dsv[2], dsv[3]);
      }
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
        for (int dpair = 0; dpair < ND / 2; ++dpair) {
          uint32_t b0, b1, b2, b3;
          const int brow = qh + kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, bcol = dpair * 16 + (lane >> 4) * 8;
          ldsm_x4_t(smem_u32(cdO + brow * S::LD + bcol), b0, b1, b2, b3);
          mma16816(dv_acc[dpair * 2], pf[kk], b0, b1);
          mma16816(dv_acc[dpair * 2 + 1], pf[kk], b2, b3);
          ldsm_x4_t(smem_u32(cQ + brow * S::LD + bcol), b0, b1, b2, b3);
          mma16816(dk_acc[dpair * 2], dsf[kk], b0, b1);
          mma16816(dk_acc[dpair * 2 + 1], dsf[kk], b2, b3);
        }
      }
    }
    __syncthreads();
    buf ^= 1;
    qt = qn;
  }
  cp_async_wait<0>();

  const int k_lo = j0 + warp * 16 + g, k_hi = k_lo + 8;
#pragma unroll
  for (int i = 0; i < ND; ++i) {
    const int col = hh * HD + i * 8 + t4 * 2;   // q-head column (== kv-head column when there is no GQA)
    if (k_lo < p.Tk) {
      *reinterpret_cast<uint32_t*>(dk_out + (long long)k_lo * ld_dk + col) = pack_bf16(dk_acc[i][0] * p.scale, dk_acc[i][1] * p.scale);
      *reinterpret_cast<uint32_t*>(dv_out + (long long)k_lo * ld_dv + col) = pack_bf16(dv_acc[i][0], dv_acc[i][1]);
    }
    if (k_hi < p.Tk) {
      *reinterpret_cast<uint32_t*>(dk_out + (long long)k_hi * ld_dk + col) = pack_bf16(dk_acc[i][2] * p.scale, dk_acc[i][3] * p.scale);
      *reinterpret_cast<uint32_t*>(dv_out + (long long)k_hi * ld_dv + col) = pack_bf16(dv_acc[i][2], dv_acc[i][3]);
    }
  }
}

// dk[t][kvh*HD + d] = sum over the q heads of the group of the expanded per-head gradients (fp32 sum)
template <int HD>
__global__ void attn_gqa_reduce_kernel(const bf16* __restrict__ xk, const bf16* __restrict__ xv, long long ldx, int rep,
                                       bf16* __restrict__ dk, bf16* __restrict__ dv, long long lddk, long long lddv,
                                       int Tk, int nkv) {
  constexpr int CH = HD / 8;
  const long long total = (long long)Tk * nkv * CH;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = idx % CH;
    const int kvh = (idx / CH) % nkv;
    const long long t = idx / ((long long)CH * nkv);
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      const bf16* src = (which ? xv : xk) + t * ldx + (long long)kvh * rep * HD + c * 8;
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int h = 0; h < rep; ++h) {
        const uint4 u = *reinterpret_cast<const uint4*>(src + (long long)h * HD);
        const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), cc = unpack_bf16(u.z), d = unpack_bf16(u.w);
        acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y; acc[4] += cc.x; acc[5] += cc.y; acc[6] += d.x; acc[7] += d.y;
      }
      uint4 o;
      o.x = pack_bf16(acc[0], acc[1]); o.y = pack_bf16(acc[2], acc[3]);
      o.z = pack_bf16(acc[4], acc[5]); o.w = pack_bf16(acc[6], acc[7]);
      bf16* dst = (which ? dv + t * lddv : dk + t * lddk) + (long long)kvh * HD + c * 8;
      *reinterpret_cast<uint4*>(dst) = o;
    }
  }
}

template <int HD>
__global__ void __launch_bounds__(ATT_THREADS)
attn_bwd_dq_kernel(const AttnParams p, bf16* __restrict__ dq_out, long long ld_dq) {
  using S = Smem<HD>;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* sQ = reinterpret_cast<bf16*>(smem_raw);
  bf16* sdO = sQ + S::TILE;
  bf16* sK = sdO + S::TILE;          // 2 buffers
  bf16* sV = sK + 2 * S::TILE;       // 2 buffers
  int* sMeta = reinterpret_cast<int*>(sV + 2 * S::TILE);  // 192 ints

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int q0 = blockIdx.x * BQ;
  const int head = blockIdx.y;
  const int kvh = head / (p.n_heads / p.n_kv_heads);
  const bf16* kg = p.k + (long long)kvh * HD;
  const bf16* vg = p.v + (long long)kvh * HD;

  const TileBounds tb = tile_bounds(p.meta, q0, p.T, sMeta, tid);
  const int r_lo = warp * 16 + g, r_hi = r_lo + 8;
  const int pre_lo = sMeta[r_lo], st_lo = sMeta[64 + r_lo], en_lo = sMeta[128 + r_lo];
  const int pre_hi = sMeta[r_hi], st_hi = sMeta[64 + r_hi], en_hi = sMeta[128 + r_hi];
  const int row_lo = q0 + r_lo, row_hi = q0 + r_hi;
  const float L2E = 1.4426950408889634f;
  const float lse_lo = row_lo < p.T ? p.lse[(long long)head * p.T + row_lo] * L2E : 0.f;
  const float lse_hi = row_hi < p.T ? p.lse[(long long)head * p.T + row_hi] * L2E : 0.f;
  const float dl_lo = row_lo < p.T ? p.delta[(long long)head * p.T + row_lo] : 0.f;
  const float dl_hi = row_hi < p.T ? p.delta[(long long)head * p.T + row_hi] : 0.f;
  const int n_tiles = (min(max(tb.pmax, tb.emax), p.Tk) + BKV - 1) / BKV;

  load_tile<HD>(smem_u32(sQ), p.q + (long long)head * HD, p.ldq, q0, p.T, tid);
  load_tile<HD>(smem_u32(sdO), p.d_o + (long long)head * HD, p.lddo, q0, p.T, tid);
  int jt = next_kv_tile(tb, 0, n_tiles);
  if (jt < n_tiles) {
    load_tile<HD>(smem_u32(sK), kg, p.ldk, jt * BKV, p.Tk, tid);
    load_tile<HD>(smem_u32(sV), vg, p.ldv, jt * BKV, p.Tk, tid);
  }
  cp_async_commit();

  constexpr int KS = HD / 16;
  constexpr int NO = HD / 8;
  uint32_t qf[KS][4];
  float dq_acc[NO][4];
#pragma unroll
  for (int i = 0; i < NO; ++i) { dq_acc[i][0] = dq_acc[i][1] = dq_acc[i][2] = dq_acc[i][3] = 0.f; }
  const float sc2 = p.scale * L2E;

  int buf = 0;
  bool q_loaded = false;
  while (jt < n_tiles) {
    const int jn = next_kv_tile(tb, jt + 1, n_tiles);
    if (jn < n_tiles) {
      load_tile<HD>(smem_u32(sK + (buf ^ 1) * S::TILE), kg, p.ldk, jn * BKV, p.Tk, tid);
      load_tile<HD>(smem_u32(sV + (buf ^ 1) * S::TILE), vg, p.ldv, jn * BKV, p.Tk, tid);
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (!q_loaded) {
#pragma unroll
      for (int ks = 0; ks < KS; ++ks)
        ldsm_x4(smem_u32(sQ + (warp * 16 + (lane & 15)) * S::LD + ks * 16 + (lane >> 4) * 8), qf[ks][0], qf[ks][1],
                qf[ks][2], qf[ks][3]);
      q_loaded = true;
    }
    const bf16* cK = sK + buf * S::TILE;
    const bf16* cV = sV + buf * S::TILE;
    const int j0 = jt * BKV;
    const bool full = kv_tile_full(tb, j0) && (j0 + BKV <= p.Tk);
#pragma unroll 1
    for (int hf = 0; hf < 2; ++hf) {   // two halves of 32 keys
      const int kh = hf * 32;
      float s[4][4], dp[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
        dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
      }
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t da[4];
        ldsm_x4(smem_u32(sdO + (warp * 16 + (lane & 15)) * S::LD + ks * 16 + (lane >> 4) * 8), da[0], da[1], da[2], da[3]);
#pragma unroll
        for (int np = 0; np < 2; ++np) {
          uint32_t b0, b1, b2, b3;
          const int brow = kh + np * 16 + (lane & 7) + (lane >> 4) * 8, bcol = ks * 16 + ((lane >> 3) & 1) * 8;
          ldsm_x4(smem_u32(cK + brow * S::LD + bcol), b0, b1, b2, b3);
          mma16816(s[np * 2], qf[ks], b0, b1);
          mma16816(s[np * 2 + 1], qf[ks], b2, b3);
          ldsm_x4(smem_u32(cV + brow * S::LD + bcol), b0, b1, b2, b3);
          mma16816(dp[np * 2], da, b0, b1);
          mma16816(dp[np * 2 + 1], da, b2, b3);
        }
      }
      uint32_t dsf[2][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        float dsv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const bool hi = e >= 2;
          bool vis = true;
          if (!full) {
            const int j = j0 + kh + nt * 8 + t4 * 2 + (e & 1);
            const int pre = hi ? pre_hi : pre_lo, st = hi ? st_hi : st_lo, en = hi ? en_hi : en_lo;
            vis = (j < p.Tk) && ((j < pre) || (j >= st && j < en));
          }
          const float pr = vis ? exp2f(s[nt][e] * sc2 - (hi ? lse_hi : lse_lo)) : 0.f;
          dsv[e] = pr * (dp[nt][e] - (hi ? dl_hi : dl_lo));
        }
        dsf[nt >> 1][(nt & 1) * 2] = pack_bf16(dsv[0], dsv[1]);
        dsf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(dsv[2], dsv[3]);
      }
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
        for (int dpr = 0; dpr < NO / 2; ++dpr) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4_t(smem_u32(cK + (kh + kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * S::LD + dpr * 16 + (lane >> 4) * 8),
                    b0, b1, b2, b3);
          mma16816(dq_acc[dpr * 2], dsf[kk], b0, b1);
          mma16816(dq_acc[dpr * 2 + 1], dsf[kk], b2, b3);
        }
      }
    }
    __syncthreads();
    buf ^= 1;
    jt = jn;
  }
  cp_async_wait<0>();
#pragma unroll
  for (int i = 0; i < NO; ++i) {
    const int col = head * HD + i * 8 + t4 * 2;
    if (row_lo < p.T)
      *reinterpret_cast<uint32_t*>(dq_out + (long long)row_lo * ld_dq + col) = pack_bf16(dq_acc[i][0] * p.scale, dq_acc[i][1] * p.scale);
    if (row_hi < p.T)
      *reinterpret_cast<uint32_t*>(dq_out + (long long)row_hi * ld_dq + col) = pack_bf16(dq_acc[i][2] * p.scale, dq_acc[i][3] * p.scale);
  }
}

__global__ void f32_to_bf16_strided_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int T, int W,
                                           long long ldd) {
  const int nv = W / 4;
  const long long total = (long long)T * nv;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = idx % nv;
    const long long t = idx / nv;
    const float4 v = *reinterpret_cast<const float4*>(src + t * W + c * 4);
    uint2 u;
    u.x = pack_bf16(v.x, v.y);
    u.y = pack_bf16(v.z, v.w);
    *reinterpret_cast<uint2*>(dst + t * ldd + c * 4) = u;
  }
}

template <int HD>
size_t fwd_smem() { return (size_t)5 * Smem<HD>::TILE * 2 + 192 * 4; }
template <int HD>
size_t bwd_dkdv_smem() { return (size_t)6 * Smem<HD>::TILE * 2 + (4 * BQ) * 4 + 2 * 192 * 4; }
template <int HD>
size_t bwd_dq_smem() { return (size_t)6 * Smem<HD>::TILE * 2 + 192 * 4; }

template <int HD>
int launch_fwd(const AttnParams& p, cudaStream_t st) {
  static bool done = false;
  if (!done) {
    SB_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd_smem<HD>()));
    done = true;
  }
  dim3 grid((p.T + BQ - 1) / BQ, p.n_heads);
  attn_fwd_kernel<HD><<<grid, ATT_THREADS, fwd_smem<HD>(), st>>>(p);
  return sb_check_launch("sb_attn_fwd");
}

int g_attn_bwd_impl = 0;   // 0 = tcgen05 backward (default); bit 0: dQ on mma.sync, bit 1: dK/dV on mma.sync

template <int HD>
int launch_bwd(const AttnParams& p, const sb_attn_args* a, bf16* dq, long long lddq, bf16* gqa_ws, int* tile_ws,
               cudaStream_t st) {
  static bool done = false;
  if (!done) {
    SB_CUDA(cudaFuncSetAttribute(attn_bwd_dkdv_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_dkdv_smem<HD>()));
    SB_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_dq_smem<HD>()));
    done = true;
  }
  const long long items = (long long)p.T * p.n_heads;
  attn_delta_kernel<HD><<<(unsigned)((items + 7) / 8), 256, 0, st>>>(p.o, p.ldo, p.d_o, p.lddo, const_cast<float*>(p.delta), p.T, p.n_heads);
  if (sb_check_launch("sb_attn_bwd(delta)")) return 1;
  const int n_qt = (p.T + BQ - 1) / BQ;
  QTileBounds* qtb = reinterpret_cast<QTileBounds*>(tile_ws);
  attn_qtile_bounds_kernel<<<(n_qt + 3) / 4, 128, 0, st>>>(p.meta, p.T, qtb);
  if (sb_check_launch("sb_attn_bwd(bounds)")) return 1;
  const int rep = p.n_heads / p.n_kv_heads;
  const long long ldx = (long long)p.n_heads * HD;
  bf16* xk = rep > 1 ? gqa_ws : p.dk;
  bf16* xv = rep > 1 ? gqa_ws + (long long)p.Tk * ldx : p.dv;
  dim3 grid_kv((p.Tk + BKV - 1) / BKV, p.n_heads);
  // head offset folded into the output pointers: per-q-head columns in the expanded scratch, kv-head columns otherwise
  // (rep == 1: q head == kv head)
  if (g_attn_bwd_impl & 2) {
    attn_bwd_dkdv_kernel<HD><<<grid_kv, ATT_THREADS, bwd_dkdv_smem<HD>(), st>>>(p, qtb, xk, xv, rep > 1 ? ldx : p.lddk,
                                                                                rep > 1 ? ldx : p.lddv);
    if (sb_check_launch("sb_attn_bwd(dkdv)")) return 1;
  } else {
    if (sb_attn_bwd_dkdv_tc(a, reinterpret_cast<const int*>(qtb), xk, xv, rep > 1 ? ldx : p.lddk, rep > 1 ? ldx : p.lddv, st))
      return 1;
  }
  if (rep > 1) {
    const long long total = (long long)p.Tk * p.n_kv_heads * (HD / 8);
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    attn_gqa_reduce_kernel<HD><<<blocks, 256, 0, st>>>(xk, xv, ldx, rep, p.dk, p.dv, p.lddk, p.lddv, p.Tk, p.n_kv_heads);
    if (sb_check_launch("sb_attn_bwd(gqa reduce)")) return 1;
  }
  if (!(g_attn_bwd_impl & 1)) return sb_attn_bwd_dq_tc(a, st);
  dim3 grid_q((p.T + BQ - 1) / BQ, p.n_heads);
  attn_bwd_dq_kernel<HD><<<grid_q, ATT_THREADS, bwd_dq_smem<HD>(), st>>>(p, dq, lddq);
  return sb_check_launch("sb_attn_bwd(dq)");
}

}  // namespace

int sb_attn_fwd_tc(const sb_attn_args* a, cudaStream_t st);   // attention_tc.cu
namespace {
int g_attn_impl = 0;   // 0 = tcgen05 forward (default), 1 = mma.sync forward
}
extern "C" int sb_set_attn_impl(int impl) {
  SB_REQUIRE(impl == 0 || impl == 1, "sb_set_attn_impl: 0 = tcgen05 (default), 1 = mma.sync");
  g_attn_impl = impl;
  return 0;
}

extern "C" int sb_set_attn_bwd_impl(int impl) {
  SB_REQUIRE(impl >= 0 && impl <= 3, "sb_set_attn_bwd_impl: 0 = tcgen05 (default), bit 0 = dQ on mma.sync, bit 1 = dK/dV on mma.sync");
  g_attn_bwd_impl = impl;
  return 0;
}

extern "C" int sb_attn_fwd(const sb_attn_args* a, sb_stream_t stream) {
  SB_REQUIRE(a && a->q && a->k && a->v && a->o && a->meta, "sb_attn_fwd: null pointer");
  SB_REQUIRE(a->T > 0 && a->n_heads > 0 && a->n_kv_heads > 0 && a->n_heads % a->n_kv_heads == 0, "sb_attn_fwd: bad sizes");
  SB_REQUIRE(a->ldq % 8 == 0 && a->ldk % 8 == 0 && a->ldv % 8 == 0 && a->ldo % 2 == 0, "sb_attn_fwd: strides must be multiples of 8");
  if (g_attn_impl == 0) return sb_attn_fwd_tc(a, reinterpret_cast<cudaStream_t>(stream));
  AttnParams p{};
  p.q = (const bf16*)a->q; p.k = (const bf16*)a->k; p.v = (const bf16*)a->v;
  p.ldq = a->ldq; p.ldk = a->ldk; p.ldv = a->ldv;
  p.o = (bf16*)a->o; p.ldo = a->ldo; p.lse = a->lse; p.meta = (const int4*)a->meta;
  p.T = a->T; p.Tk = a->Tk > 0 ? a->Tk : a->T; p.n_heads = a->n_heads; p.n_kv_heads = a->n_kv_heads;
  p.scale = a->scale;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (a->head_dim == 128) return launch_fwd<128>(p, st);
  if (a->head_dim == 80) return launch_fwd<80>(p, st);
  sb_set_error("sb_attn_fwd: head_dim %d not supported (80 or 128)", a->head_dim);
  return 1;
}

extern "C" int sb_attn_bwd_workspace(int T, int Tk, int n_heads, int n_kv_heads, int head_dim, long long* gqa_ws_elems,
                                     long long* tile_ws_ints) {
  SB_REQUIRE(gqa_ws_elems && tile_ws_ints && T > 0 && n_heads > 0 && n_kv_heads > 0, "sb_attn_bwd_workspace: bad arguments");
  if (Tk <= 0) Tk = T;
  *gqa_ws_elems = n_heads != n_kv_heads ? 2LL * Tk * n_heads * head_dim : 0;
  *tile_ws_ints = (long long)((T + BQ - 1) / BQ) * 8;
  return 0;
}

extern "C" int sb_attn_bwd(const sb_attn_args* a, sb_stream_t stream) {
  SB_REQUIRE(a && a->q && a->k && a->v && a->o && a->meta && a->lse && a->d_o && a->delta && a->dq && a->dk && a->dv &&
                 a->tile_ws, "sb_attn_bwd: null pointer");
  SB_REQUIRE(a->T > 0 && a->n_heads % a->n_kv_heads == 0, "sb_attn_bwd: bad sizes");
  SB_REQUIRE(a->n_heads == a->n_kv_heads || a->gqa_ws, "sb_attn_bwd: GQA needs the gqa_ws scratch (sb_attn_bwd_workspace)");
  AttnParams p{};
  p.q = (const bf16*)a->q; p.k = (const bf16*)a->k; p.v = (const bf16*)a->v;
  p.ldq = a->ldq; p.ldk = a->ldk; p.ldv = a->ldv;
  p.o = (bf16*)a->o; p.ldo = a->ldo; p.lse = a->lse; p.meta = (const int4*)a->meta;
  p.T = a->T; p.Tk = a->Tk > 0 ? a->Tk : a->T; p.n_heads = a->n_heads; p.n_kv_heads = a->n_kv_heads;
  p.scale = a->scale;
  p.d_o = (const bf16*)a->d_o; p.lddo = a->lddo; p.delta = a->delta;
  p.dk = (bf16*)a->dk; p.dv = (bf16*)a->dv; p.lddk = a->lddk; p.lddv = a->lddv;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (a->head_dim == 128) return launch_bwd<128>(p, a, (bf16*)a->dq, a->lddq, (bf16*)a->gqa_ws, a->tile_ws, st);
  if (a->head_dim == 80) return launch_bwd<80>(p, a, (bf16*)a->dq, a->lddq, (bf16*)a->gqa_ws, a->tile_ws, st);
  sb_set_error("sb_attn_bwd: head_dim %d not supported (80 or 128)", a->head_dim);
  return 1;
}

extern "C" int sb_f32_to_bf16_2d(const float* src, void* dst, int T, int W, long long ldd, sb_stream_t stream) {
  SB_REQUIRE(src && dst && T > 0 && W > 0 && W % 4 == 0 && ldd % 4 == 0, "sb_f32_to_bf16_2d: bad arguments");
  long long items = (long long)T * W / 4;
  int blocks = (int)((items + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  f32_to_bf16_strided_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, (bf16*)dst, T, W, ldd);
  return sb_check_launch("sb_f32_to_bf16_2d");
}
