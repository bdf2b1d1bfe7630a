// Flash-attention forward on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), same interface and masks as
// attention.cu's mma.sync kernel:  ViT block-diagonal attention (head_dim 80, MQ2:415-454), Qwen2 causal GQA attention
// (head_dim 128, MQ2:575-590) and the prefix-shared training layout.  Visibility per query token:
//     key j visible  <=>  j < prefix_len  ||  seg_start <= j < kv_end.
//
// One CTA per (128-query tile, q head), 192 threads, TWO CTAs per SM (112 KB of shared memory and 256 TMEM columns
// each) so that one CTA's softmax overlaps the other's MMAs and every scheduler has two softmax warps to interleave:
//   warp 0      TMA producer: Q tile once, then K and V tiles of 64 keys through a 2-stage mbarrier ring
//   warp 1      single-thread tcgen05.mma issuer:  S_b = Q K^T (fp32 in TMEM, two S buffers),  O += P V
//   warps 2..5  softmax: thread r owns query row r (TMEM lane r): one tcgen05.ld of the 64 scores -> scale/mask ->
//               running max with lazy rescale (O in TMEM is rescaled only when a row max grows by more than 2^8) ->
//               P = exp2(S - m) as bf16 into 128B-swizzled shared memory (the A operand of the PV MMA) -> row sums;
//               epilogue O / l -> global.
// S for tile i+1 is issued before the softmax of tile i is consumed, so QK^T overlaps the softmax.
// TMEM columns: S0 [0,64), S1 [64,128), O [128, 128 + HD).
#include "common.cuh"
#include "spacer_b200.h"
#include <cuda.h>
#include <cudaTypedefs.h>

namespace {

constexpr int TQ = 128;          // queries per CTA
constexpr int TK = 64;           // keys per pipeline stage
constexpr int THREADS = 192;
constexpr int KV_STAGES = 2;
constexpr int SLAB = 128 * 128;  // bytes of one [128 rows x 64 bf16] 128B-swizzled K-major slab

SB_DEVICE void tmem_st_32x32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
SB_DEVICE void tmem_st_32x16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
SB_DEVICE void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
SB_DEVICE float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct TcParams {
  bf16* o; long long ldo;
  float* lse;             // [n_heads][T] natural log, optional
  const int4* meta;
  int T, Tk, n_heads, n_kv_heads;
  float scale_log2;
};

struct Bounds { int pmin, pmax, smin, smax, emin, emax; };

SB_DEVICE bool tile_relevant(const Bounds& b, int j0) { return (j0 < b.pmax) || (j0 + TK > b.smin && j0 < b.emax); }
SB_DEVICE bool tile_full(const Bounds& b, int j0) { return (j0 + TK <= b.pmin) || (j0 >= b.smax && j0 + TK <= b.emin); }
SB_DEVICE int next_tile(const Bounds& b, int jt, int n_tiles) {
  while (jt < n_tiles && !tile_relevant(b, jt * TK)) ++jt;
  return jt;
}

template <int HD>
struct Lay {
  static constexpr int NSLAB = (HD + 63) / 64;              // 64-wide head-dim chunks (HD 80 -> 2, second half-used)
  static constexpr int Q_BYTES = NSLAB * SLAB;
  static constexpr int K_BYTES = NSLAB * 8192;              // NSLAB slabs of [64 keys x 64 d], used K-major (k = d)
  static constexpr int V_BYTES = NSLAB * 8192;              // same tile shape, used MN-major (n = d, k = keys)
  static constexpr int P_BYTES = SLAB;                      // [128 q x 64 keys] K-major
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + Q_BYTES;
  static constexpr int OFF_V = OFF_K + KV_STAGES * K_BYTES;
  static constexpr int OFF_P = OFF_V + KV_STAGES * V_BYTES;
  static constexpr int OFF_BAR = OFF_P + P_BYTES;
  static constexpr int SMEM = OFF_BAR + 256;   // 112.25 KB: two CTAs per SM (the per-row mask metadata is staged in
                                               // the P tile before the first P is written)
  static constexpr int TMEM_COLS = 256;
  static constexpr int O_COL = 128;
};

template <int HD>
__global__ void __launch_bounds__(THREADS, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const TcParams p) {
  using L = Lay<HD>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  const uint32_t sbase = smem_u32(smem);
  if ((sbase & 1023u) != 0) __trap();     // the swizzled tiles need 1024-byte alignment
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  const uint32_t q_full = smem_u32(bars + 0);
  const uint32_t kv_full0 = smem_u32(bars + 1);      // [2]
  const uint32_t kv_empty0 = smem_u32(bars + 3);     // [2]
  const uint32_t s_full0 = smem_u32(bars + 5);       // [2]
  const uint32_t p_full = smem_u32(bars + 7);
  const uint32_t pv_done = smem_u32(bars + 8);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  int* sMeta = reinterpret_cast<int*>(smem + L::OFF_P);   // 3 x 128 ints staged in the (not yet used) P tile

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * TQ;
  const int head = blockIdx.y;
  const int kvh = head / (p.n_heads / p.n_kv_heads);

  // visibility bounds of this query tile
  if (threadIdx.x < TQ) {
    int4 m = make_int4(0, 0, 0, 0);
    if (q0 + (int)threadIdx.x < p.T) m = p.meta[q0 + threadIdx.x];
    sMeta[threadIdx.x] = m.x; sMeta[TQ + threadIdx.x] = m.y; sMeta[2 * TQ + threadIdx.x] = m.z;
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < KV_STAGES; ++s) { mbar_init(kv_full0 + 8 * s, 1); mbar_init(kv_empty0 + 8 * s, 1); }
    mbar_init(s_full0, 1); mbar_init(s_full0 + 8, 1);
    mbar_init(p_full, TQ);
    mbar_init(pv_done, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), L::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  Bounds tb;
  tb.pmin = 1 << 30; tb.pmax = 0; tb.smin = 1 << 30; tb.smax = 0; tb.emin = 1 << 30; tb.emax = 0;
  for (int i = 0; i < TQ && q0 + i < p.T; ++i) {
    const int pr = sMeta[i], s = sMeta[TQ + i], e = sMeta[2 * TQ + i];
    tb.pmin = min(tb.pmin, pr); tb.pmax = max(tb.pmax, pr);
    if (e > s) { tb.smin = min(tb.smin, s); tb.emax = max(tb.emax, e); }
    tb.smax = max(tb.smax, s); tb.emin = min(tb.emin, e);
  }
  const int n_tiles = (min(max(tb.pmax, tb.emax), p.Tk) + TK - 1) / TK;
  int pre = 0, seg = 0, kve = 0;
  if (warp >= 2) {   // softmax thread of row r = (warp & 3) * 32 + lane
    const int rr = (warp & 3) * 32 + lane;
    pre = sMeta[rr]; seg = sMeta[TQ + rr]; kve = sMeta[2 * TQ + rr];
  }
  __syncthreads();   // sMeta lives in the P tile: everybody is done with it before the first P is written

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      mbar_expect_tx(q_full, L::Q_BYTES);
#pragma unroll
      for (int c = 0; c < L::NSLAB; ++c) tma_load_2d(sbase + L::OFF_Q + c * SLAB, &tmQ, q_full, head * HD + c * 64, q0);
      int stage = 0;
      uint32_t phase = 0;
      for (int jt = next_tile(tb, 0, n_tiles); jt < n_tiles; jt = next_tile(tb, jt + 1, n_tiles)) {
        mbar_wait(kv_empty0 + 8 * stage, phase ^ 1);
        const uint32_t fb = kv_full0 + 8 * stage;
        mbar_expect_tx(fb, L::K_BYTES + L::V_BYTES);
        const uint32_t sk = sbase + L::OFF_K + stage * L::K_BYTES;
        const uint32_t sv = sbase + L::OFF_V + stage * L::V_BYTES;
#pragma unroll
        for (int c = 0; c < L::NSLAB; ++c) {
          tma_load_2d(sk + c * 8192, &tmK, fb, kvh * HD + c * 64, jt * TK);
          tma_load_2d(sv + c * 8192, &tmV, fb, kvh * HD + c * 64, jt * TK);
        }
        if (++stage == KV_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(TQ, TK, false, false);
      constexpr uint32_t idesc_o = umma_idesc_bf16(TQ, HD, false, true);
      auto issue_s = [&](int stage, int sbuf) {
        const uint32_t sq = sbase + L::OFF_Q;
        const uint32_t sk = sbase + L::OFF_K + stage * L::K_BYTES;
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const uint64_t adesc = umma_desc_sw128(sq + (kk / 4) * SLAB, 0, 1024) + (uint64_t)((kk % 4) * 2);
          const uint64_t bdesc = umma_desc_sw128(sk + (kk / 4) * 8192, 0, 1024) + (uint64_t)((kk % 4) * 2);
          tc_mma_bf16(tmem_base + sbuf * TK, adesc, bdesc, idesc_s, kk > 0 ? 1u : 0u);
        }
        tc_commit(s_full0 + 8 * sbuf);
      };
      mbar_wait(q_full, 0);
      tc_fence_after();
      int jt = next_tile(tb, 0, n_tiles);
      if (jt < n_tiles) {
        mbar_wait(kv_full0, 0);
        tc_fence_after();
        issue_s(0, 0);
      }
      for (int it = 0; jt < n_tiles; ++it) {
        const int jn = next_tile(tb, jt + 1, n_tiles);
        const int stage = it & 1;
        if (jn < n_tiles) {
          // S of the next tile overlaps the softmax of this one.  Its S buffer and K/V stage were released by tile
          // it - 1 (p_full waited / PV committed in the previous iteration).
          const int nt = it + 1;
          mbar_wait(kv_full0 + 8 * (nt & 1), (nt >> 1) & 1);
          tc_fence_after();
          issue_s(nt & 1, nt & 1);
        }
        mbar_wait(p_full, it & 1);
        tc_fence_after();
        const uint32_t sp = sbase + L::OFF_P;
        const uint32_t sv = sbase + L::OFF_V + stage * L::V_BYTES;
#pragma unroll
        for (int kk = 0; kk < TK / 16; ++kk) {
          const uint64_t adesc = umma_desc_sw128(sp, 0, 1024) + (uint64_t)(kk * 2);
          const uint64_t bdesc = umma_desc_sw128(sv, 8192, 1024) + (uint64_t)(kk * (2048 >> 4));
          tc_mma_bf16(tmem_base + L::O_COL, adesc, bdesc, idesc_o, (it > 0 || kk > 0) ? 1u : 0u);
        }
        tc_commit(kv_empty0 + 8 * stage);
        tc_commit(pv_done);
        jt = jn;
      }
    }
  } else {
    // ------------------------------ softmax + epilogue (4 warps, thread = query row) ------------------------------
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int row = q0 + r;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    float m_ref = -INFINITY, l = 0.f;
    int it = 0;
    uint8_t* sP = smem + L::OFF_P;
    for (int jt = next_tile(tb, 0, n_tiles); jt < n_tiles; jt = next_tile(tb, jt + 1, n_tiles), ++it) {
      const int sbuf = it & 1;
      mbar_wait(s_full0 + 8 * sbuf, (it >> 1) & 1);
      tc_fence_after();
      const int j0 = jt * TK;
      const bool full = tile_full(tb, j0) && (j0 + TK <= p.Tk) && (q0 + TQ <= p.T);
      const uint32_t s_addr = lane_addr + sbuf * TK;
      // the 64 scores of this row, read from TMEM once
      uint32_t sv_[TK];
      tmem_ld_32x32(s_addr, sv_);
      tmem_ld_32x32(s_addr + 32, sv_ + 32);
      tmem_ld_wait();
      float mx = -INFINITY;
      if (full) {
#pragma unroll
        for (int j = 0; j < TK; ++j) mx = fmaxf(mx, __uint_as_float(sv_[j]));
      } else {
#pragma unroll
        for (int j = 0; j < TK; ++j) {
          const int key = j0 + j;
          const bool vis = (row < p.T) && (key < p.Tk) && ((key < pre) || (key >= seg && key < kve));
          if (!vis) sv_[j] = 0xff800000u;   // -inf
          mx = fmaxf(mx, __uint_as_float(sv_[j]));
        }
      }
      mx *= p.scale_log2;                   // scale > 0: max commutes with the scaling
      // lazy rescale: move the reference max only when it grows by more than 2^8 (P stays <= 256)
      float factor = 1.f;
      if (mx > m_ref + 8.f) {
        factor = (m_ref == -INFINITY) ? 0.f : fast_exp2(m_ref - mx);
        m_ref = mx;
      }
      // the previous PV must be complete before O is rescaled or P is overwritten
      if (it > 0) {
        mbar_wait(pv_done, (it - 1) & 1);
        tc_fence_after();
        if (__any_sync(0xffffffffu, factor != 1.f)) {
          constexpr int NCH = HD / 32;
#pragma unroll 1
          for (int c = 0; c < NCH; ++c) {
            uint32_t v[32];
            tmem_ld_32x32(lane_addr + L::O_COL + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) * factor);
            tmem_st_32x32(lane_addr + L::O_COL + c * 32, v);
          }
          if constexpr (HD % 32 != 0) {
            uint32_t v[16];
            tmem_ld_32x16(lane_addr + L::O_COL + NCH * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) * factor);
            tmem_st_32x16(lane_addr + L::O_COL + NCH * 32, v);
          }
          tmem_st_wait();
        }
      }
      l *= factor;
      // P = exp2(S * scale - m_ref) -> bf16 -> swizzled smem (A operand of the PV MMA); row sum in fp32
      const float neg_m = (m_ref == -INFINITY) ? 0.f : -m_ref;   // fully masked so far: every score is -inf -> P = 0
      float rs0 = 0.f, rs1 = 0.f;
      uint8_t* rowp = sP + r * 128;
#pragma unroll
      for (int c8 = 0; c8 < TK / 8; ++c8) {   // 8 keys = one 16-byte chunk
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          const float p0 = fast_exp2(fmaf(__uint_as_float(sv_[c8 * 8 + j]), p.scale_log2, neg_m));
          const float p1 = fast_exp2(fmaf(__uint_as_float(sv_[c8 * 8 + j + 1]), p.scale_log2, neg_m));
          rs0 += p0; rs1 += p1;
          pk[j >> 1] = pack_bf16(p0, p1);
        }
        *reinterpret_cast<uint4*>(rowp + ((c8 ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      const float rs = rs0 + rs1;
      l += rs;
      fence_proxy_async();      // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc_fence_before();
      mbar_arrive(p_full);
    }
    // epilogue
    if (it > 0) {
      mbar_wait(pv_done, (it - 1) & 1);
      tc_fence_after();
    }
    const float inv = l > 0.f ? 1.f / l : 0.f;
    bf16* orow = p.o + (long long)row * p.ldo + (long long)head * HD;
    constexpr int NCH = HD / 32;
    // tcgen05.ld is warp-collective: every lane loads, only rows < T store
#pragma unroll 1
    for (int c = 0; c < NCH; ++c) {
      uint32_t v[32];
      if (it > 0) {
        tmem_ld_32x32(lane_addr + L::O_COL + c * 32, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
      if (row < p.T) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint4 u;
          u.x = pack_bf16(__uint_as_float(v[j]) * inv, __uint_as_float(v[j + 1]) * inv);
          u.y = pack_bf16(__uint_as_float(v[j + 2]) * inv, __uint_as_float(v[j + 3]) * inv);
          u.z = pack_bf16(__uint_as_float(v[j + 4]) * inv, __uint_as_float(v[j + 5]) * inv);
          u.w = pack_bf16(__uint_as_float(v[j + 6]) * inv, __uint_as_float(v[j + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + c * 32 + j) = u;
        }
      }
    }
    if constexpr (HD % 32 != 0) {
      uint32_t v[16];
      if (it > 0) {
        tmem_ld_32x16(lane_addr + L::O_COL + NCH * 32, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0u;
      }
      if (row < p.T) {
#pragma unroll
        for (int j = 0; j < 16; j += 8) {
          uint4 u;
          u.x = pack_bf16(__uint_as_float(v[j]) * inv, __uint_as_float(v[j + 1]) * inv);
          u.y = pack_bf16(__uint_as_float(v[j + 2]) * inv, __uint_as_float(v[j + 3]) * inv);
          u.z = pack_bf16(__uint_as_float(v[j + 4]) * inv, __uint_as_float(v[j + 5]) * inv);
          u.w = pack_bf16(__uint_as_float(v[j + 6]) * inv, __uint_as_float(v[j + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + NCH * 32 + j) = u;
        }
      }
    }
    if (p.lse && row < p.T)
      p.lse[(long long)head * p.T + row] = l > 0.f ? (m_ref + log2f(l)) * 0.6931471805599453f : -INFINITY;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, L::TMEM_COLS);
}

PFN_cuTensorMapEncodeTiled_v12000 g_encode_tc = nullptr;

int get_encode_tc() {
  if (g_encode_tc) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) {
    sb_set_error("cuTensorMapEncodeTiled not available from the driver");
    return 1;
  }
  g_encode_tc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  return 0;
}

// 2D bf16 map over [rows, cols] with row stride ld elements; box = [box_rows x 64 columns], 128B swizzle
int make_map(CUtensorMap* m, const void* ptr, uint64_t cols, uint64_t rows, uint64_t ld, uint32_t box_rows) {
  if (get_encode_tc()) return 1;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode_tc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    sb_set_error("sb_attn_fwd(tcgen05): cuTensorMapEncodeTiled failed (%d): ptr=%p cols=%llu rows=%llu ld=%llu", (int)r, ptr,
                 (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)ld);
    return 1;
  }
  return 0;
}

template <int HD>
int launch_tc(const sb_attn_args* a, cudaStream_t st) {
  using L = Lay<HD>;
  static bool done = false;
  if (!done) {
    SB_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::SMEM));
    done = true;
  }
  const int Tk = a->Tk > 0 ? a->Tk : a->T;
  CUtensorMap tq, tk, tv;
  if (make_map(&tq, a->q, (uint64_t)a->n_heads * HD, a->T, a->ldq, 128)) return 1;
  if (make_map(&tk, a->k, (uint64_t)a->n_kv_heads * HD, Tk, a->ldk, 64)) return 1;
  if (make_map(&tv, a->v, (uint64_t)a->n_kv_heads * HD, Tk, a->ldv, 64)) return 1;
  TcParams p;
  p.o = (bf16*)a->o; p.ldo = a->ldo; p.lse = a->lse; p.meta = (const int4*)a->meta;
  p.T = a->T; p.Tk = Tk; p.n_heads = a->n_heads; p.n_kv_heads = a->n_kv_heads;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  dim3 grid((a->T + TQ - 1) / TQ, a->n_heads);
  attn_fwd_tc_kernel<HD><<<grid, THREADS, L::SMEM, st>>>(tq, tk, tv, p);
  return sb_check_launch("sb_attn_fwd(tcgen05)");
}

}  // namespace

// called from sb_attn_fwd (attention.cu) when the tcgen05 path is selected
int sb_attn_fwd_tc(const sb_attn_args* a, cudaStream_t st) {
  SB_REQUIRE((reinterpret_cast<uintptr_t>(a->q) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->k) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(a->v) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->o) & 15) == 0,
             "sb_attn_fwd(tcgen05): q/k/v/o must be 16-byte aligned");
  SB_REQUIRE(a->ldo % 8 == 0, "sb_attn_fwd(tcgen05): ldo must be a multiple of 8");
  if (a->head_dim == 128) return launch_tc<128>(a, st);
  if (a->head_dim == 80) return launch_tc<80>(a, st);
  sb_set_error("sb_attn_fwd: head_dim %d not supported (80 or 128)", a->head_dim);
  return 1;
}
