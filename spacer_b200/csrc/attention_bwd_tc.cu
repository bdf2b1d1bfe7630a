// Flash-attention BACKWARD on the 5th-generation tensor cores (tcgen05 + TMEM + TMA); same masks, inputs and outputs as
// the mma.sync kernels in attention.cu (which remain selectable through sb_set_attn_bwd_impl for cross-checks).
// replaces the backward of flash_attn_varlen_func / sdpa under MQ2:415-454 (ViT, head_dim 80) and MQ2:575-590 (Qwen2
// causal GQA, head_dim 128) that autograd runs inside accelerator.backward.
//
// Deterministic, no atomics: two kernels, each recomputing S and dP for the tile pairs it visits.
//   attn_bwd_dq_tc_kernel    one CTA per (128-query tile, q head); TMEM lane = query row.  Per 64-key tile:
//                            S = Q K^T, dP = dO V^T (fp32 in TMEM, double-buffered) -> P = exp2(S c - lse),
//                            dS = P o (dP - delta) c' as bf16 into swizzled shared memory -> dQ += dS K.
//   attn_bwd_dkdv_tc_kernel  one CTA per (128-key tile, q head); TMEM lane = key row.  Per 64-query tile that can see
//                            the keys: S^T = K Q^T, dP^T = V dO^T -> P^T, dS^T as bf16 operands -> dV += P^T dO,
//                            dK += dS^T Q.  With GQA the per-q-head dK/dV land in an expanded scratch that
//                            attn_gqa_reduce_kernel (attention.cu) sums over the group.
// Warp roles: warp 0 TMA producer, warp 1 single-thread MMA issuer, warps 2..9 element-wise work and epilogue -- two
// warps per TMEM lane quarter, each taking 32 of the 64 columns of a tile (the element-wise pass, not the tensor
// pipe, bounds a tile: exp2 + mask + two bf16 operand rows per score).  One CTA per SM (512 TMEM columns, 144-163 KB smem);
// the MMAs of tile i+1 (S, dP) are issued before the element-wise pass of tile i is consumed.
#include "common.cuh"
#include "spacer_b200.h"
#include <cuda.h>
#include <cudaTypedefs.h>

namespace {

constexpr int THREADS = 320;          // TMA warp, MMA warp, 8 element-wise warps (two per TMEM lane quarter)
constexpr int EW_THREADS = 256;
constexpr int TB = 64;            // inner tile (keys in the dQ kernel, queries in the dK/dV kernel)
constexpr int SLAB = 128 * 128;   // bytes of one [128 rows x 64 bf16] 128B-swizzled slab
constexpr int HSLAB = 64 * 128;   // bytes of one [64 rows x 64 bf16] slab
constexpr float L2E = 1.4426950408889634f;

SB_DEVICE float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
SB_DEVICE void epi_bar2() { asm volatile("bar.sync 2, 256;" ::: "memory"); }

struct BwdParams {
  const float* lse;      // [n_heads][T] natural log
  const float* delta;    // [n_heads][T]
  const int4* meta;
  int T, Tk, n_heads, n_kv_heads;
  float scale;
  bf16* out0; long long ld0;   // dq                      | dk (or expanded scratch)
  bf16* out1; long long ld1;   // unused                  | dv (or expanded scratch)
  const int* qtb;              // dK/dV kernel: per-64-query-tile bounds {pmin,pmax,smin,smax,emin,emax,0,0}
};

struct Bounds { int pmin, pmax, smin, smax, emin, emax; };

// ---------------------------------------------------------------------------------------------------------------
// dQ
// ---------------------------------------------------------------------------------------------------------------
SB_DEVICE bool ktile_relevant(const Bounds& b, int j0) { return (j0 < b.pmax) || (j0 + TB > b.smin && j0 < b.emax); }
SB_DEVICE bool ktile_full(const Bounds& b, int j0) { return (j0 + TB <= b.pmin) || (j0 >= b.smax && j0 + TB <= b.emin); }
SB_DEVICE int next_ktile(const Bounds& b, int jt, int n_tiles) {
  while (jt < n_tiles && !ktile_relevant(b, jt * TB)) ++jt;
  return jt;
}

template <int HD>
struct LayQ {
  static constexpr int NSLAB = (HD + 63) / 64;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_DO = OFF_Q + NSLAB * SLAB;
  static constexpr int OFF_K = OFF_DO + NSLAB * SLAB;            // 2 stages
  static constexpr int OFF_V = OFF_K + 2 * NSLAB * HSLAB;        // 2 stages
  static constexpr int OFF_DS = OFF_V + 2 * NSLAB * HSLAB;       // [128 q x 64 keys] K-major
  static constexpr int OFF_BAR = OFF_DS + SLAB;
  static constexpr int SMEM = OFF_BAR + 256;
  static constexpr int S_COL = 0, DP_COL = 128, ACC_COL = 256;   // S0,S1 | dP0,dP1 | dQ
};

template <int HD>
__global__ void __launch_bounds__(THREADS, 1)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmdO,
                      const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                      const BwdParams p) {
  using L = LayQ<HD>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  const uint32_t sbase = smem_u32(smem);
  if ((sbase & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  const uint32_t q_full = smem_u32(bars + 0);
  const uint32_t kv_full0 = smem_u32(bars + 1);    // [2]
  const uint32_t kv_empty0 = smem_u32(bars + 3);   // [2]
  const uint32_t s_full0 = smem_u32(bars + 5);     // [2]  S and dP of a buffer complete
  const uint32_t ds_full = smem_u32(bars + 7);
  const uint32_t acc_done = smem_u32(bars + 8);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  int* sMeta = reinterpret_cast<int*>(smem + L::OFF_DS);   // staged in the (not yet used) dS tile

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128;
  const int head = blockIdx.y;
  const int kvh = head / (p.n_heads / p.n_kv_heads);

  if (threadIdx.x < 128) {
    int4 m = make_int4(0, 0, 0, 0);
    if (q0 + (int)threadIdx.x < p.T) m = p.meta[q0 + threadIdx.x];
    sMeta[threadIdx.x] = m.x; sMeta[128 + threadIdx.x] = m.y; sMeta[256 + threadIdx.x] = m.z;
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmdO); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(kv_full0 + 8 * s, 1); mbar_init(kv_empty0 + 8 * s, 1); mbar_init(s_full0 + 8 * s, 1); }
    mbar_init(ds_full, EW_THREADS);
    mbar_init(acc_done, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  Bounds tb;
  tb.pmin = 1 << 30; tb.pmax = 0; tb.smin = 1 << 30; tb.smax = 0; tb.emin = 1 << 30; tb.emax = 0;
  for (int i = 0; i < 128 && q0 + i < p.T; ++i) {
    const int pr = sMeta[i], s = sMeta[128 + i], e = sMeta[256 + i];
    tb.pmin = min(tb.pmin, pr); tb.pmax = max(tb.pmax, pr);
    if (e > s) { tb.smin = min(tb.smin, s); tb.emax = max(tb.emax, e); }
    tb.smax = max(tb.smax, s); tb.emin = min(tb.emin, e);
  }
  const int n_tiles = (min(max(tb.pmax, tb.emax), p.Tk) + TB - 1) / TB;
  int pre = 0, seg = 0, kve = 0;
  if (warp >= 2) {
    const int rr = (warp & 3) * 32 + lane;
    pre = sMeta[rr]; seg = sMeta[128 + rr]; kve = sMeta[256 + rr];
  }
  __syncthreads();   // sMeta lives in the dS tile

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, 2 * L::NSLAB * SLAB);
#pragma unroll
      for (int c = 0; c < L::NSLAB; ++c) {
        tma_load_2d(sbase + L::OFF_Q + c * SLAB, &tmQ, q_full, head * HD + c * 64, q0);
        tma_load_2d(sbase + L::OFF_DO + c * SLAB, &tmdO, q_full, head * HD + c * 64, q0);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int jt = next_ktile(tb, 0, n_tiles); jt < n_tiles; jt = next_ktile(tb, jt + 1, n_tiles)) {
        mbar_wait(kv_empty0 + 8 * stage, phase ^ 1);
        const uint32_t fb = kv_full0 + 8 * stage;
        mbar_expect_tx(fb, 2 * L::NSLAB * HSLAB);
#pragma unroll
        for (int c = 0; c < L::NSLAB; ++c) {
          tma_load_2d(sbase + L::OFF_K + (stage * L::NSLAB + c) * HSLAB, &tmK, fb, kvh * HD + c * 64, jt * TB);
          tma_load_2d(sbase + L::OFF_V + (stage * L::NSLAB + c) * HSLAB, &tmV, fb, kvh * HD + c * 64, jt * TB);
        }
        if (++stage == 2) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, TB, false, false);
      constexpr uint32_t idesc_acc = umma_idesc_bf16(128, HD, false, true);
      auto issue_s = [&](int stage, int buf) {
        const uint32_t sk = sbase + L::OFF_K + stage * L::NSLAB * HSLAB;
        const uint32_t sv = sbase + L::OFF_V + stage * L::NSLAB * HSLAB;
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const uint64_t a = umma_desc_sw128(sbase + L::OFF_Q + (kk / 4) * SLAB, 0, 1024) + (uint64_t)((kk % 4) * 2);
          const uint64_t b = umma_desc_sw128(sk + (kk / 4) * HSLAB, 0, 1024) + (uint64_t)((kk % 4) * 2);
          tc_mma_bf16(tmem_base + L::S_COL + buf * TB, a, b, idesc_s, kk > 0 ? 1u : 0u);
        }
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const uint64_t a = umma_desc_sw128(sbase + L::OFF_DO + (kk / 4) * SLAB, 0, 1024) + (uint64_t)((kk % 4) * 2);
          const uint64_t b = umma_desc_sw128(sv + (kk / 4) * HSLAB, 0, 1024) + (uint64_t)((kk % 4) * 2);
          tc_mma_bf16(tmem_base + L::DP_COL + buf * TB, a, b, idesc_s, kk > 0 ? 1u : 0u);
        }
        tc_commit(s_full0 + 8 * buf);
      };
      mbar_wait(q_full, 0);
      tc_fence_after();
      int jt = next_ktile(tb, 0, n_tiles);
      if (jt < n_tiles) {
        mbar_wait(kv_full0, 0);
        tc_fence_after();
        issue_s(0, 0);
      }
      for (int it = 0; jt < n_tiles; ++it) {
        const int jn = next_ktile(tb, jt + 1, n_tiles);
        const int stage = it & 1;
        if (jn < n_tiles) {
          const int nt = it + 1;
          mbar_wait(kv_full0 + 8 * (nt & 1), (nt >> 1) & 1);
          tc_fence_after();
          issue_s(nt & 1, nt & 1);
        }
        mbar_wait(ds_full, it & 1);
        tc_fence_after();
        const uint32_t sds = sbase + L::OFF_DS;
        const uint32_t sk = sbase + L::OFF_K + stage * L::NSLAB * HSLAB;
#pragma unroll
        for (int kk = 0; kk < TB / 16; ++kk) {
          const uint64_t a = umma_desc_sw128(sds, 0, 1024) + (uint64_t)(kk * 2);
          const uint64_t b = umma_desc_sw128(sk, HSLAB, 1024) + (uint64_t)(kk * (2048 >> 4));
          tc_mma_bf16(tmem_base + L::ACC_COL, a, b, idesc_acc, (it > 0 || kk > 0) ? 1u : 0u);
        }
        tc_commit(kv_empty0 + 8 * stage);
        tc_commit(acc_done);
        jt = jn;
      }
    }
  } else {
    const int quarter = warp & 3;
    const int hf = (warp - 2) >> 2;          // which 32 of the tile's 64 key columns this warp handles
    const int r = quarter * 32 + lane;
    const int row = q0 + r;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    float lse2 = INFINITY, dl = 0.f;
    if (row < p.T) {
      const float l = p.lse[(long long)head * p.T + row];
      lse2 = (l == -INFINITY) ? INFINITY : l * L2E;
      dl = p.delta[(long long)head * p.T + row];
    }
    const float sc2 = p.scale * L2E;
    uint8_t* rowp = smem + L::OFF_DS + r * 128;
    int it = 0;
    for (int jt = next_ktile(tb, 0, n_tiles); jt < n_tiles; jt = next_ktile(tb, jt + 1, n_tiles), ++it) {
      const int buf = it & 1;
      mbar_wait(s_full0 + 8 * buf, (it >> 1) & 1);
      tc_fence_after();
      const int j0 = jt * TB;
      const bool full = ktile_full(tb, j0) && (j0 + TB <= p.Tk) && (q0 + 128 <= p.T);
      uint32_t sv_[32], dp_[32];
      tmem_ld_32x32(lane_addr + L::S_COL + buf * TB + hf * 32, sv_);
      tmem_ld_32x32(lane_addr + L::DP_COL + buf * TB + hf * 32, dp_);
      tmem_ld_wait();
      if (it > 0) {   // the previous dQ MMA has finished reading the dS tile
        mbar_wait(acc_done, (it - 1) & 1);
        tc_fence_after();
      }
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          float ds[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int c = c8 * 8 + j + e;
            bool vis = true;
            if (!full) {
              const int key = j0 + hf * 32 + c;
              vis = (row < p.T) && (key < p.Tk) && ((key < pre) || (key >= seg && key < kve));
            }
            const float pr = vis ? fast_exp2(fmaf(__uint_as_float(sv_[c]), sc2, -lse2)) : 0.f;
            ds[e] = pr * (__uint_as_float(dp_[c]) - dl) * p.scale;
          }
          pk[j >> 1] = pack_bf16(ds[0], ds[1]);
        }
        *reinterpret_cast<uint4*>(rowp + (((hf * 4 + c8) ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(ds_full);
    }
    if (it > 0) {
      mbar_wait(acc_done, (it - 1) & 1);
      tc_fence_after();
    }
    bf16* orow = p.out0 + (long long)row * p.ld0 + (long long)head * HD;
    constexpr int NCH = HD / 16;
    const int c_lo = hf == 0 ? 0 : (NCH + 1) / 2, c_hi = hf == 0 ? (NCH + 1) / 2 : NCH;
#pragma unroll 1
    for (int c = c_lo; c < c_hi; ++c) {
      uint32_t v[16];
      if (it > 0) {
        tmem_ld_32x16(lane_addr + L::ACC_COL + c * 16, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0u;
      }
      if (row < p.T) {
#pragma unroll
        for (int j = 0; j < 16; j += 8) {
          uint4 u;
          u.x = pack_bf16(__uint_as_float(v[j]), __uint_as_float(v[j + 1]));
          u.y = pack_bf16(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
          u.z = pack_bf16(__uint_as_float(v[j + 4]), __uint_as_float(v[j + 5]));
          u.w = pack_bf16(__uint_as_float(v[j + 6]), __uint_as_float(v[j + 7]));
          *reinterpret_cast<uint4*>(orow + c * 16 + j) = u;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// dK, dV
// ---------------------------------------------------------------------------------------------------------------
SB_DEVICE Bounds load_qtb(const int* qtb, int qt) {
  const int4 a = *reinterpret_cast<const int4*>(qtb + qt * 8);
  const int2 b = *reinterpret_cast<const int2*>(qtb + qt * 8 + 4);
  Bounds r;
  r.pmin = a.x; r.pmax = a.y; r.smin = a.z; r.smax = a.w; r.emin = b.x; r.emax = b.y;
  return r;
}
// first 64-query tile >= qt that sees at least one key of this CTA's key tile (flags: shared memory, see below)
SB_DEVICE int next_qtile(const uint8_t* flags, int qt, int n_qt) {
  while (qt < n_qt && !(flags[qt] & 1)) ++qt;
  return qt;
}

template <int HD>
struct LayK {
  static constexpr int NSLAB = (HD + 63) / 64;
  static constexpr int OFF_K = 0;
  static constexpr int OFF_V = OFF_K + NSLAB * SLAB;
  static constexpr int OFF_Q = OFF_V + NSLAB * SLAB;             // 2 stages of [64 q x HD]
  static constexpr int OFF_DO = OFF_Q + 2 * NSLAB * HSLAB;       // 2 stages
  static constexpr int OFF_PT = OFF_DO + 2 * NSLAB * HSLAB;      // [128 keys x 64 q] K-major
  static constexpr int OFF_DST = OFF_PT + SLAB;
  static constexpr int OFF_ROW = OFF_DST + SLAB;                 // 2 x {lse2[64], delta[64], pre[64], seg[64], kve[64]}
  static constexpr int OFF_FLAG = OFF_ROW + 2 * 5 * 64 * 4;      // per 64-query tile: bit 0 relevant, bit 1 fully visible
  static constexpr int MAX_QT = 4096;
  static constexpr int OFF_BAR = OFF_FLAG + MAX_QT;
  static constexpr int SMEM = OFF_BAR + 256;
  static constexpr int S_COL = 0, DP_COL = 128, DV_COL = 256, DK_COL = 384;
};

template <int HD>
__global__ void __launch_bounds__(THREADS, 1)
attn_bwd_dkdv_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmdO,
                        const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                        const BwdParams p) {
  using L = LayK<HD>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  const uint32_t sbase = smem_u32(smem);
  if ((sbase & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  const uint32_t kv_full = smem_u32(bars + 0);
  const uint32_t q_full0 = smem_u32(bars + 1);     // [2]
  const uint32_t q_empty0 = smem_u32(bars + 3);    // [2]
  const uint32_t s_full0 = smem_u32(bars + 5);     // [2]
  const uint32_t pds_full = smem_u32(bars + 7);
  const uint32_t acc_done = smem_u32(bars + 8);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  float* sRow = reinterpret_cast<float*>(smem + L::OFF_ROW);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int j0 = blockIdx.x * 128;
  const int head = blockIdx.y;
  const int kvh = head / (p.n_heads / p.n_kv_heads);
  const int n_qt = (p.T + TB - 1) / TB;
  uint8_t* flags = smem + L::OFF_FLAG;

  // which query tiles see this key tile: evaluated once, in parallel (a sequential scan of the bounds in global memory
  // costs one dependent load per tile per warp role)
  for (int qt = threadIdx.x; qt < n_qt; qt += THREADS) {
    const Bounds b = load_qtb(p.qtb, qt);
    const bool rel = (j0 < b.pmax) || (j0 + 128 > b.smin && j0 < b.emax);
    const bool full = ((j0 + 128 <= b.pmin) || (j0 >= b.smax && j0 + 128 <= b.emin)) && (j0 + 128 <= p.Tk) &&
                      (qt * TB + TB <= p.T);
    flags[qt] = (rel ? 1 : 0) | (full ? 2 : 0);
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmdO); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(q_full0 + 8 * s, 1); mbar_init(q_empty0 + 8 * s, 1); mbar_init(s_full0 + 8 * s, 1); }
    mbar_init(pds_full, EW_THREADS);
    mbar_init(acc_done, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(kv_full, 2 * L::NSLAB * SLAB);
#pragma unroll
      for (int c = 0; c < L::NSLAB; ++c) {
        tma_load_2d(sbase + L::OFF_K + c * SLAB, &tmK, kv_full, kvh * HD + c * 64, j0);
        tma_load_2d(sbase + L::OFF_V + c * SLAB, &tmV, kv_full, kvh * HD + c * 64, j0);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int qt = next_qtile(flags, 0, n_qt); qt < n_qt; qt = next_qtile(flags, qt + 1, n_qt)) {
        mbar_wait(q_empty0 + 8 * stage, phase ^ 1);
        const uint32_t fb = q_full0 + 8 * stage;
        mbar_expect_tx(fb, 2 * L::NSLAB * HSLAB);
#pragma unroll
        for (int c = 0; c < L::NSLAB; ++c) {
          tma_load_2d(sbase + L::OFF_Q + (stage * L::NSLAB + c) * HSLAB, &tmQ, fb, head * HD + c * 64, qt * TB);
          tma_load_2d(sbase + L::OFF_DO + (stage * L::NSLAB + c) * HSLAB, &tmdO, fb, head * HD + c * 64, qt * TB);
        }
        if (++stage == 2) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, TB, false, false);
      constexpr uint32_t idesc_acc = umma_idesc_bf16(128, HD, false, true);
      auto issue_s = [&](int stage, int buf) {
        const uint32_t sq = sbase + L::OFF_Q + stage * L::NSLAB * HSLAB;
        const uint32_t sdo = sbase + L::OFF_DO + stage * L::NSLAB * HSLAB;
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const uint64_t a = umma_desc_sw128(sbase + L::OFF_K + (kk / 4) * SLAB, 0, 1024) + (uint64_t)((kk % 4) * 2);
          const uint64_t b = umma_desc_sw128(sq + (kk / 4) * HSLAB, 0, 1024) + (uint64_t)((kk % 4) * 2);
          tc_mma_bf16(tmem_base + L::S_COL + buf * TB, a, b, idesc_s, kk > 0 ? 1u : 0u);
        }
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const uint64_t a = umma_desc_sw128(sbase + L::OFF_V + (kk / 4) * SLAB, 0, 1024) + (uint64_t)((kk % 4) * 2);
          const uint64_t b = umma_desc_sw128(sdo + (kk / 4) * HSLAB, 0, 1024) + (uint64_t)((kk % 4) * 2);
          tc_mma_bf16(tmem_base + L::DP_COL + buf * TB, a, b, idesc_s, kk > 0 ? 1u : 0u);
        }
        tc_commit(s_full0 + 8 * buf);
      };
      mbar_wait(kv_full, 0);
      tc_fence_after();
      int qt = next_qtile(flags, 0, n_qt);
      if (qt < n_qt) {
        mbar_wait(q_full0, 0);
        tc_fence_after();
        issue_s(0, 0);
      }
      for (int it = 0; qt < n_qt; ++it) {
        const int qn = next_qtile(flags, qt + 1, n_qt);
        const int stage = it & 1;
        if (qn < n_qt) {
          const int nt = it + 1;
          mbar_wait(q_full0 + 8 * (nt & 1), (nt >> 1) & 1);
          tc_fence_after();
          issue_s(nt & 1, nt & 1);
        }
        mbar_wait(pds_full, it & 1);
        tc_fence_after();
        const uint32_t sq = sbase + L::OFF_Q + stage * L::NSLAB * HSLAB;
        const uint32_t sdo = sbase + L::OFF_DO + stage * L::NSLAB * HSLAB;
#pragma unroll
        for (int kk = 0; kk < TB / 16; ++kk) {
          const uint64_t a = umma_desc_sw128(sbase + L::OFF_PT, 0, 1024) + (uint64_t)(kk * 2);
          const uint64_t b = umma_desc_sw128(sdo, HSLAB, 1024) + (uint64_t)(kk * (2048 >> 4));
          tc_mma_bf16(tmem_base + L::DV_COL, a, b, idesc_acc, (it > 0 || kk > 0) ? 1u : 0u);
        }
#pragma unroll
        for (int kk = 0; kk < TB / 16; ++kk) {
          const uint64_t a = umma_desc_sw128(sbase + L::OFF_DST, 0, 1024) + (uint64_t)(kk * 2);
          const uint64_t b = umma_desc_sw128(sq, HSLAB, 1024) + (uint64_t)(kk * (2048 >> 4));
          tc_mma_bf16(tmem_base + L::DK_COL, a, b, idesc_acc, (it > 0 || kk > 0) ? 1u : 0u);
        }
        tc_commit(q_empty0 + 8 * stage);
        tc_commit(acc_done);
        qt = qn;
      }
    }
  } else {
    const int quarter = warp & 3;
    const int hf = (warp - 2) >> 2;           // which 32 of the tile's 64 query columns this warp handles
    const int r = quarter * 32 + lane;        // key row of this thread
    const int ew = (warp - 2) * 32 + lane;    // 0..255
    const int key = j0 + r;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const float sc2 = p.scale * L2E;
    uint8_t* rowP = smem + L::OFF_PT + r * 128;
    uint8_t* rowD = smem + L::OFF_DST + r * 128;
    int it = 0;
    for (int qt = next_qtile(flags, 0, n_qt); qt < n_qt; qt = next_qtile(flags, qt + 1, n_qt), ++it) {
      const int buf = it & 1;
      const int q0 = qt * TB;
      // per-query row parameters of this tile -> shared memory (double-buffered by tile parity)
      float* cL = sRow + buf * 5 * 64;
      float* cD = cL + 64;
      int* cM = reinterpret_cast<int*>(cD + 64);
      if (ew < 64) {
        const int q = q0 + ew;
        float l2 = INFINITY, dl = 0.f;
        int4 m = make_int4(0, 0, 0, 0);
        if (q < p.T) {
          const float l = p.lse[(long long)head * p.T + q];
          l2 = (l == -INFINITY) ? INFINITY : l * L2E;
          dl = p.delta[(long long)head * p.T + q];
          m = p.meta[q];
        }
        cL[ew] = -l2; cD[ew] = dl; cM[ew] = m.x; cM[64 + ew] = m.y; cM[128 + ew] = m.z;
      }
      epi_bar2();
      const bool full = (flags[qt] & 2) != 0;
      mbar_wait(s_full0 + 8 * buf, (it >> 1) & 1);
      tc_fence_after();
      uint32_t sv_[32], dp_[32];
      tmem_ld_32x32(lane_addr + L::S_COL + buf * TB + hf * 32, sv_);
      tmem_ld_32x32(lane_addr + L::DP_COL + buf * TB + hf * 32, dp_);
      tmem_ld_wait();
      if (it > 0) {   // the previous dV/dK MMAs have finished reading the P^T / dS^T tiles
        mbar_wait(acc_done, (it - 1) & 1);
        tc_fence_after();
      }
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        const int qb = hf * 32 + c8 * 8;
        const float4 la = *reinterpret_cast<const float4*>(cL + qb), lb = *reinterpret_cast<const float4*>(cL + qb + 4);
        const float4 da = *reinterpret_cast<const float4*>(cD + qb), db = *reinterpret_cast<const float4*>(cD + qb + 4);
        const float nl[8] = {la.x, la.y, la.z, la.w, lb.x, lb.y, lb.z, lb.w};
        const float dd[8] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w};
        uint32_t pk[4], dk_[4];
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          float pv[2], ds[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int c = c8 * 8 + j + e;
            bool vis = true;
            if (!full) {
              const int ql = qb + j + e;
              const int pre = cM[ql], seg = cM[64 + ql], kve = cM[128 + ql];
              vis = (q0 + ql < p.T) && (key < p.Tk) && ((key < pre) || (key >= seg && key < kve));
            }
            const float pr = vis ? fast_exp2(fmaf(__uint_as_float(sv_[c]), sc2, nl[j + e])) : 0.f;
            pv[e] = pr;
            ds[e] = pr * (__uint_as_float(dp_[c]) - dd[j + e]) * p.scale;
          }
          pk[j >> 1] = pack_bf16(pv[0], pv[1]);
          dk_[j >> 1] = pack_bf16(ds[0], ds[1]);
        }
        const int ch = ((hf * 4 + c8) ^ (r & 7)) << 4;
        *reinterpret_cast<uint4*>(rowP + ch) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(rowD + ch) = make_uint4(dk_[0], dk_[1], dk_[2], dk_[3]);
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(pds_full);
    }
    if (it > 0) {
      mbar_wait(acc_done, (it - 1) & 1);
      tc_fence_after();
    }
    {   // warps of column half 0 write dK, the others dV
      const int which = hf;
      bf16* orow = (which ? p.out1 + (long long)key * p.ld1 : p.out0 + (long long)key * p.ld0) + (long long)head * HD;
      const uint32_t col0 = which ? L::DV_COL : L::DK_COL;
#pragma unroll 1
      for (int c = 0; c < HD / 16; ++c) {
        uint32_t v[16];
        if (it > 0) {
          tmem_ld_32x16(lane_addr + col0 + c * 16, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0u;
        }
        if (key < p.Tk) {
#pragma unroll
          for (int j = 0; j < 16; j += 8) {
            uint4 u;
            u.x = pack_bf16(__uint_as_float(v[j]), __uint_as_float(v[j + 1]));
            u.y = pack_bf16(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            u.z = pack_bf16(__uint_as_float(v[j + 4]), __uint_as_float(v[j + 5]));
            u.w = pack_bf16(__uint_as_float(v[j + 6]), __uint_as_float(v[j + 7]));
            *reinterpret_cast<uint4*>(orow + c * 16 + j) = u;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 g_encode_b = nullptr;

int get_encode_b() {
  if (g_encode_b) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) {
    sb_set_error("cuTensorMapEncodeTiled not available from the driver");
    return 1;
  }
  g_encode_b = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  return 0;
}

// 2D bf16 map over [rows, cols] with row stride ld elements; box = [box_rows x 64 columns], 128B swizzle, OOB = 0
int make_map_b(CUtensorMap* m, const void* ptr, uint64_t cols, uint64_t rows, uint64_t ld, uint32_t box_rows) {
  if (get_encode_b()) return 1;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode_b(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    sb_set_error("sb_attn_bwd(tcgen05): cuTensorMapEncodeTiled failed (%d): ptr=%p cols=%llu rows=%llu ld=%llu", (int)r, ptr,
                 (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)ld);
    return 1;
  }
  return 0;
}

BwdParams base_params(const sb_attn_args* a) {
  BwdParams p{};
  p.lse = a->lse; p.delta = a->delta; p.meta = (const int4*)a->meta;
  p.T = a->T; p.Tk = a->Tk > 0 ? a->Tk : a->T; p.n_heads = a->n_heads; p.n_kv_heads = a->n_kv_heads;
  p.scale = a->scale;
  return p;
}

template <int HD>
int launch_dq(const sb_attn_args* a, cudaStream_t st) {
  using L = LayQ<HD>;
  static bool done = false;
  if (!done) {
    SB_CUDA(cudaFuncSetAttribute(attn_bwd_dq_tc_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::SMEM));
    done = true;
  }
  BwdParams p = base_params(a);
  p.out0 = (bf16*)a->dq; p.ld0 = a->lddq;
  CUtensorMap tq, tdo, tk, tv;
  if (make_map_b(&tq, a->q, (uint64_t)a->n_heads * HD, a->T, a->ldq, 128)) return 1;
  if (make_map_b(&tdo, a->d_o, (uint64_t)a->n_heads * HD, a->T, a->lddo, 128)) return 1;
  if (make_map_b(&tk, a->k, (uint64_t)a->n_kv_heads * HD, p.Tk, a->ldk, 64)) return 1;
  if (make_map_b(&tv, a->v, (uint64_t)a->n_kv_heads * HD, p.Tk, a->ldv, 64)) return 1;
  dim3 grid((a->T + 127) / 128, a->n_heads);
  attn_bwd_dq_tc_kernel<HD><<<grid, THREADS, L::SMEM, st>>>(tq, tdo, tk, tv, p);
  return sb_check_launch("sb_attn_bwd(dq, tcgen05)");
}

template <int HD>
int launch_dkdv(const sb_attn_args* a, const int* qtb, void* dk_out, void* dv_out, long long ld_dk, long long ld_dv,
                cudaStream_t st) {
  using L = LayK<HD>;
  static bool done = false;
  if (!done) {
    SB_CUDA(cudaFuncSetAttribute(attn_bwd_dkdv_tc_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::SMEM));
    done = true;
  }
  SB_REQUIRE((a->T + TB - 1) / TB <= L::MAX_QT, "sb_attn_bwd(tcgen05): at most %d query tiles of %d (T = %d)", L::MAX_QT, TB, a->T);
  BwdParams p = base_params(a);
  p.out0 = (bf16*)dk_out; p.ld0 = ld_dk;
  p.out1 = (bf16*)dv_out; p.ld1 = ld_dv;
  p.qtb = qtb;
  CUtensorMap tq, tdo, tk, tv;
  if (make_map_b(&tq, a->q, (uint64_t)a->n_heads * HD, a->T, a->ldq, 64)) return 1;
  if (make_map_b(&tdo, a->d_o, (uint64_t)a->n_heads * HD, a->T, a->lddo, 64)) return 1;
  if (make_map_b(&tk, a->k, (uint64_t)a->n_kv_heads * HD, p.Tk, a->ldk, 128)) return 1;
  if (make_map_b(&tv, a->v, (uint64_t)a->n_kv_heads * HD, p.Tk, a->ldv, 128)) return 1;
  dim3 grid((p.Tk + 127) / 128, a->n_heads);
  attn_bwd_dkdv_tc_kernel<HD><<<grid, THREADS, L::SMEM, st>>>(tq, tdo, tk, tv, p);
  return sb_check_launch("sb_attn_bwd(dkdv, tcgen05)");
}

}  // namespace

// called from sb_attn_bwd (attention.cu).  qtb: bounds of the 64-query tiles (attn_qtile_bounds_kernel).
// dk_out / dv_out: where q head h's dK/dV go, head h at column h * head_dim (the expanded scratch under GQA).
int sb_attn_bwd_dkdv_tc(const sb_attn_args* a, const int* qtb, void* dk_out, void* dv_out, long long ld_dk,
                        long long ld_dv, cudaStream_t st) {
  SB_REQUIRE(a->lddo % 8 == 0 && ld_dk % 8 == 0 && ld_dv % 8 == 0, "sb_attn_bwd(tcgen05): strides must be multiples of 8");
  SB_REQUIRE((reinterpret_cast<uintptr_t>(a->d_o) & 15) == 0 && (reinterpret_cast<uintptr_t>(dk_out) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(dv_out) & 15) == 0, "sb_attn_bwd(tcgen05): pointers must be 16-byte aligned");
  if (a->head_dim == 128) return launch_dkdv<128>(a, qtb, dk_out, dv_out, ld_dk, ld_dv, st);
  if (a->head_dim == 80) return launch_dkdv<80>(a, qtb, dk_out, dv_out, ld_dk, ld_dv, st);
  sb_set_error("sb_attn_bwd: head_dim %d not supported (80 or 128)", a->head_dim);
  return 1;
}

int sb_attn_bwd_dq_tc(const sb_attn_args* a, cudaStream_t st) {
  SB_REQUIRE(a->lddo % 8 == 0 && a->lddq % 8 == 0, "sb_attn_bwd(tcgen05): strides must be multiples of 8");
  SB_REQUIRE((reinterpret_cast<uintptr_t>(a->d_o) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->dq) & 15) == 0,
             "sb_attn_bwd(tcgen05): pointers must be 16-byte aligned");
  if (a->head_dim == 128) return launch_dq<128>(a, st);
  if (a->head_dim == 80) return launch_dq<80>(a, st);
  sb_set_error("sb_attn_bwd: head_dim %d not supported (80 or 128)", a->head_dim);
  return 1;
}
