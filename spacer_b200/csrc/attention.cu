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
  float* dq_acc;              // fp32 [T][n_heads*HD]
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

// One CTA owns a 64-key tile of one kv head; it loops over the q heads of the group and over every
// 32-query tile that can see the keys.  Each warp owns 16 keys: S^T = K Q^T, dV += P^T dO,
// dP^T = V dO^T, dS^T = P^T o (dP^T - delta), dK += dS^T Q; dQ += dS K goes through smem + fp32 atomics.
constexpr int BQB = 32;  // query rows per backward iteration

template <int HD>
__global__ void __launch_bounds__(ATT_THREADS)
attn_bwd_kernel(const AttnParams p) {
  using S = Smem<HD>;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* sK = reinterpret_cast<bf16*>(smem_raw);          // [64][LD]
  bf16* sV = sK + S::TILE;                               // [64][LD]
  bf16* sQ = sV + S::TILE;                               // 2 x [32][LD]
  bf16* sdO = sQ + 2 * BQB * S::LD;                      // 2 x [32][LD]
  bf16* sdS = sdO + 2 * BQB * S::LD;                     // [32 q][64+8 keys]
  float* sLse = reinterpret_cast<float*>(sdS + BQB * 72);  // 2 x 32
  float* sDelta = sLse + 2 * BQB;                         // 2 x 32
  int* sMeta = reinterpret_cast<int*>(sDelta + 2 * BQB);  // 2 x 3 x 32

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int j0 = blockIdx.x * BKV;
  const int kvh = blockIdx.y;
  const int rep = p.n_heads / p.n_kv_heads;
  const float sc = p.scale;

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

  const int n_qt = (p.T + BQB - 1) / BQB;
  const int total = n_qt * rep;   // iteration space: (q tile, head in group); q tile outer

  auto issue = [&](int it, int b) {
    const int qt = it / rep, hh = kvh * rep + it % rep;
    const int q0 = qt * BQB;
    constexpr int CH = HD / 8;
    for (int i = tid; i < BQB * CH; i += ATT_THREADS) {
      const int r = i / CH, c = i % CH;
      const bool ok = (q0 + r) < p.T;
      const long long row = ok ? (q0 + r) : 0;
      cp_async16(smem_u32(sQ + b * BQB * S::LD + r * S::LD + c * 8), p.q + row * p.ldq + (long long)hh * HD + c * 8, ok);
      cp_async16(smem_u32(sdO + b * BQB * S::LD + r * S::LD + c * 8), p.d_o + row * p.lddo + (long long)hh * HD + c * 8, ok);
    }
    if (tid < BQB) {
      const bool ok = (q0 + tid) < p.T;
      sLse[b * BQB + tid] = ok ? p.lse[(long long)hh * p.T + q0 + tid] : 0.f;
      sDelta[b * BQB + tid] = ok ? p.delta[(long long)hh * p.T + q0 + tid] : 0.f;
      int4 m = make_int4(0, 0, 0, 0);
      if (ok) m = p.meta[q0 + tid];
      sMeta[b * 96 + tid] = m.x; sMeta[b * 96 + 32 + tid] = m.y; sMeta[b * 96 + 64 + tid] = m.z;
    }
  };
  // a q tile can see this key tile iff some row has prefix_len > j0 or an own range intersecting it.
  // Cheap conservative host-free test using the tile's first/last rows is not enough for general metas,
  // so the relevance test reads the meta of all 32 rows (L2-resident, tiny).
  auto relevant = [&](int qt) -> bool {
    const int q0 = qt * BQB;
    bool r = false;
    for (int i = 0; i < BQB && q0 + i < p.T; ++i) {
      const int4 m = p.meta[q0 + i];
      r |= (j0 < m.x) || (j0 + BKV > m.y && j0 < m.z && m.z > m.y);
    }
    return r;
  };
  auto next_it = [&](int it) -> int {
    while (it < total) {
      if (relevant(it / rep)) return it;
      it = (it / rep + 1) * rep;  // skip the whole q tile
    }
    return total;
  };

  int it = next_it(0);
  int buf = 0;
  if (it < total) issue(it, 0);
  cp_async_commit();

  while (it < total) {
    const int itn = ((it + 1) % rep == 0) ? next_it(it + 1) : it + 1;
    if (itn < total) issue(itn, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();

    const int qt = it / rep, hh = kvh * rep + it % rep;
    const int q0 = qt * BQB;
    const bf16* cQ = sQ + buf * BQB * S::LD;
    const bf16* cdO = sdO + buf * BQB * S::LD;
    const float* cL = sLse + buf * BQB;
    const float* cD = sDelta + buf * BQB;
    const int* cM = sMeta + buf * 96;

    // S^T [16 keys x 32 q] = K_w Q^T ; dP^T = V_w dO^T
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
        const int brow = np * 16 + (lane & 7) + (lane >> 4) * 8, bcol = ks * 16 + ((lane >> 3) & 1) * 8;
        ldsm_x4(smem_u32(cQ + brow * S::LD + bcol), b0, b1, b2, b3);
        mma16816(st[np * 2], ka, b0, b1);
        mma16816(st[np * 2 + 1], ka, b2, b3);
        ldsm_x4(smem_u32(cdO + brow * S::LD + bcol), b0, b1, b2, b3);
        mma16816(dp[np * 2], va, b0, b1);
        mma16816(dp[np * 2 + 1], va, b2, b3);
      }
    }
    // P^T, dS^T
    uint32_t pf[2][4], dsf[2][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      float pv[4], dsv[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int ql = nt * 8 + t4 * 2 + (e & 1);       // query within tile
        const int key = j0 + warp * 16 + g + (e >= 2 ? 8 : 0);
        const int pre = cM[ql], s0 = cM[32 + ql], en = cM[64 + ql];
        const bool vis = (q0 + ql < p.T) && (key < p.Tk) && ((key < pre) || (key >= s0 && key < en));
        const float pr = vis ? __expf(st[nt][e] * sc - cL[ql]) : 0.f;
        pv[e] = pr;
        dsv[e] = pr * (dp[nt][e] - cD[ql]);
      }
      // accumulator (rows = keys g/g+8, cols = q) -> A fragment of a 16(keys) x 16(q) block
      pf[nt >> 1][(nt & 1) * 2] = pack_bf16(pv[0], pv[1]);
      pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(pv[2], pv[3]);
      dsf[nt >> 1][(nt & 1) * 2] = pack_bf16(dsv[0], dsv[1]);
      dsf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(dsv[2], dsv[3]);
      // dS (not transposed) to smem for the dQ product: sdS[q][key]
      const int kl = warp * 16 + g;
      sdS[(nt * 8 + t4 * 2) * 72 + kl] = __float2bfloat16_rn(dsv[0]);
      sdS[(nt * 8 + t4 * 2 + 1) * 72 + kl] = __float2bfloat16_rn(dsv[1]);
      sdS[(nt * 8 + t4 * 2) * 72 + kl + 8] = __float2bfloat16_rn(dsv[2]);
      sdS[(nt * 8 + t4 * 2 + 1) * 72 + kl + 8] = __float2bfloat16_rn(dsv[3]);
    }
    // dV += P^T dO ; dK += dS^T Q     (k = 32 queries = 2 k-steps; B from [q][d] smem via .trans)
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
      for (int dpair = 0; dpair < ND / 2; ++dpair) {
        uint32_t b0, b1, b2, b3;
        const int brow = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, bcol = dpair * 16 + (lane >> 4) * 8;
        ldsm_x4_t(smem_u32(cdO + brow * S::LD + bcol), b0, b1, b2, b3);
        mma16816(dv_acc[dpair * 2], pf[kk], b0, b1);
        mma16816(dv_acc[dpair * 2 + 1], pf[kk], b2, b3);
        ldsm_x4_t(smem_u32(cQ + brow * S::LD + bcol), b0, b1, b2, b3);
        mma16816(dk_acc[dpair * 2], dsf[kk], b0, b1);
        mma16816(dk_acc[dpair * 2 + 1], dsf[kk], b2, b3);
      }
    }
    __syncthreads();  // sdS complete
    // dQ [32 q x HD] += dS [32 x 64 keys] K [64 x HD]; warp w: q rows (w&1)*16.., head-dim half (w>>1)
    {
      const int qr = (warp & 1) * 16;
      constexpr int NH = ND / 2;            // n-tiles per head-dim half
      const int dbase = (warp >> 1) * (HD / 2);
      float dq[NH][4];
#pragma unroll
      for (int i = 0; i < NH; ++i) { dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f; }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t a[4];
        ldsm_x4(smem_u32(sdS + (qr + (lane & 15)) * 72 + kk * 16 + (lane >> 4) * 8), a[0], a[1], a[2], a[3]);
#pragma unroll
        for (int i = 0; i < NH; i += 2) {
          uint32_t b0, b1, b2, b3;
          const int brow = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, bcol = dbase + i * 8 + (lane >> 4) * 8;
          if (i + 1 < NH) {
            ldsm_x4_t(smem_u32(sK + brow * S::LD + bcol), b0, b1, b2, b3);
            mma16816(dq[i], a, b0, b1);
            mma16816(dq[i + 1], a, b2, b3);
          } else {
            // odd tail (HD=80: 5 n-tiles per half): x4 load would run past the half; use the first pair only
            ldsm_x4_t(smem_u32(sK + brow * S::LD + (bcol - (lane >> 4) * 8)), b0, b1, b2, b3);
            mma16816(dq[i], a, b0, b1);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < NH; ++i) {
        const int col = hh * HD + dbase + i * 8 + t4 * 2;
        const int r_lo = q0 + qr + g, r_hi = r_lo + 8;
        if (r_lo < p.T) {
          atomicAdd(p.dq_acc + (long long)r_lo * p.n_heads * HD + col, dq[i][0] * sc);
          atomicAdd(p.dq_acc + (long long)r_lo * p.n_heads * HD + col + 1, dq[i][1] * sc);
        }
        if (r_hi < p.T) {
          atomicAdd(p.dq_acc + (long long)r_hi * p.n_heads * HD + col, dq[i][2] * sc);
          atomicAdd(p.dq_acc + (long long)r_hi * p.n_heads * HD + col + 1, dq[i][3] * sc);
        }
      }
    }
    __syncthreads();
    buf ^= 1;
    it = itn;
  }
  cp_async_wait<0>();

  const int k_lo = j0 + warp * 16 + g, k_hi = k_lo + 8;
#pragma unroll
  for (int i = 0; i < ND; ++i) {
    const int col = kvh * HD + i * 8 + t4 * 2;
    if (k_lo < p.Tk) {
      *reinterpret_cast<uint32_t*>(p.dk + (long long)k_lo * p.lddk + col) = pack_bf16(dk_acc[i][0] * sc, dk_acc[i][1] * sc);
      *reinterpret_cast<uint32_t*>(p.dv + (long long)k_lo * p.lddv + col) = pack_bf16(dv_acc[i][0], dv_acc[i][1]);
    }
    if (k_hi < p.Tk) {
      *reinterpret_cast<uint32_t*>(p.dk + (long long)k_hi * p.lddk + col) = pack_bf16(dk_acc[i][2] * sc, dk_acc[i][3] * sc);
      *reinterpret_cast<uint32_t*>(p.dv + (long long)k_hi * p.lddv + col) = pack_bf16(dv_acc[i][2], dv_acc[i][3]);
    }
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
size_t bwd_smem() {
  return (size_t)(2 * Smem<HD>::TILE + 4 * BQB * Smem<HD>::LD + BQB * 72) * 2 + (4 * BQB) * 4 + 2 * 96 * 4;
}

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

template <int HD>
int launch_bwd(const AttnParams& p, cudaStream_t st) {
  static bool done = false;
  if (!done) {
    SB_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_smem<HD>()));
    done = true;
  }
  const long long items = (long long)p.T * p.n_heads;
  attn_delta_kernel<HD><<<(unsigned)((items + 7) / 8), 256, 0, st>>>(p.o, p.ldo, p.d_o, p.lddo, const_cast<float*>(p.delta), p.T, p.n_heads);
  if (sb_check_launch("sb_attn_bwd(delta)")) return 1;
  dim3 grid((p.Tk + BKV - 1) / BKV, p.n_kv_heads);
  attn_bwd_kernel<HD><<<grid, ATT_THREADS, bwd_smem<HD>(), st>>>(p);
  return sb_check_launch("sb_attn_bwd");
}

}  // namespace

extern "C" int sb_attn_fwd(const sb_attn_args* a, sb_stream_t stream) {
  SB_REQUIRE(a && a->q && a->k && a->v && a->o && a->meta, "sb_attn_fwd: null pointer");
  SB_REQUIRE(a->T > 0 && a->n_heads > 0 && a->n_kv_heads > 0 && a->n_heads % a->n_kv_heads == 0, "sb_attn_fwd: bad sizes");
  SB_REQUIRE(a->ldq % 8 == 0 && a->ldk % 8 == 0 && a->ldv % 8 == 0 && a->ldo % 2 == 0, "sb_attn_fwd: strides must be multiples of 8");
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

extern "C" int sb_attn_bwd(const sb_attn_args* a, sb_stream_t stream) {
  SB_REQUIRE(a && a->q && a->k && a->v && a->o && a->meta && a->lse && a->d_o && a->delta && a->dq_acc && a->dk && a->dv,
             "sb_attn_bwd: null pointer");
  SB_REQUIRE(a->T > 0 && a->n_heads % a->n_kv_heads == 0, "sb_attn_bwd: bad sizes");
  AttnParams p{};
  p.q = (const bf16*)a->q; p.k = (const bf16*)a->k; p.v = (const bf16*)a->v;
  p.ldq = a->ldq; p.ldk = a->ldk; p.ldv = a->ldv;
  p.o = (bf16*)a->o; p.ldo = a->ldo; p.lse = a->lse; p.meta = (const int4*)a->meta;
  p.T = a->T; p.Tk = a->Tk > 0 ? a->Tk : a->T; p.n_heads = a->n_heads; p.n_kv_heads = a->n_kv_heads;
  p.scale = a->scale;
  p.d_o = (const bf16*)a->d_o; p.lddo = a->lddo; p.delta = a->delta; p.dq_acc = a->dq_acc;
  p.dk = (bf16*)a->dk; p.dv = (bf16*)a->dv; p.lddk = a->lddk; p.lddv = a->lddv;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (a->head_dim == 128) return launch_bwd<128>(p, st);
  if (a->head_dim == 80) return launch_bwd<80>(p, st);
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
