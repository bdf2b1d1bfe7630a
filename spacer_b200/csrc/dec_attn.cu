// Decode-step attention (q_len = 1) for R <= 32 rows that share one or two prompts -- "prefix-shared flash decoding".
// replaces the attention of the q_len=1 iterations of GenerationMixin._sample (generation/utils.py:2743-2806) through
// Qwen2VLAttention (MQ2:507-594) over DynamicCache (cache_utils.py:102-120), where the reference keeps G copies of the
// prompt KV and reads all of them every step.
//
// The context of row r is [shared prompt of r's group | r's own completion so far].  Work is cut into independent
// items, one CTA each, all in ONE launch:
//   prefix items (group, q-block, kv head, key split): the <= 128 query vectors (rows of the group x the q heads of the
//                kv head) against a slice of the SHARED prompt K/V -> the prompt cache is read once per group per
//                step, not once per row, and the products run on tcgen05 tensor cores (M = rows x heads);
//   own items    (row, kv head, key split): the row's q heads against a slice of its own completion cache.
// Every item writes a normalised partial (O, log2-sum-exp) for its (row, head) pairs; a combine kernel merges them.
// The step is read from device memory so that one captured CUDA graph replays for every step.
// (Rounds 1-2 ran these items on mma.sync m16n8k16 -- the last legacy-tensor-path kernel of the decode step; the tcgen05
// version measured the same step time and replaced it.)
#include "common.cuh"
#include "spacer_b200.h"
#include <cuda.h>
#include <cudaTypedefs.h>

namespace {

constexpr int HD = 128;
SB_DEVICE void cp_async16(uint32_t dst, const void* src, bool pred) {
  const int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
SB_DEVICE void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
SB_DEVICE void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct Plan {
  int rep;            // q heads per kv head
  int n_rows[2];      // rows per group
  int rows_qb;        // rows per query block (each row brings its `rep` q heads)
  int n_qb[2];        // query blocks per group
  int p_chunk;        // prompt keys per prefix split (multiple of TK)
  int n_psplit, n_csplit, NS;
  int n_prefix_items, n_items;
};

Plan make_plan(int R, int rows_group0, int P, int c_max, int nh, int nkv, int sms) {
  constexpr int tq = 128, TK = 64;     // query vectors per item (TMEM lanes), keys per tile
  Plan pl;
  pl.rep = nh / nkv;
  pl.n_rows[0] = rows_group0 < R ? rows_group0 : R;
  pl.n_rows[1] = R - pl.n_rows[0];
  // a q-block holds whole rows (all rep heads of a row): rows_per_block = tq / rep
  pl.rows_qb = tq / pl.rep;
  for (int g = 0; g < 2; ++g) pl.n_qb[g] = (pl.n_rows[g] + pl.rows_qb - 1) / pl.rows_qb;
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

// ---------------------------------------------------------------------------------------------------------------------
// The same items on the 5th-generation tensor cores (tcgen05 + TMEM + TMA) -- the engine of attention_tc.cu's forward
// kernel with the decode step's work split.  With mma.sync the kernel was bound by the legacy tensor path: 256 m16n8k16
// per warp for a 128-key item, ~6 us of the layer's dependency chain.
//   warp 0      TMA producer: K and V tiles of 64 keys through a 2-stage mbarrier ring.  Prefix items (constant prompt
//               cache) and all but the current token's tile of own items are requested BEFORE the dependency wait.
//   warp 1      single-thread tcgen05.mma issuer: S_b = Q K^T (two S buffers in TMEM), O += P V
//   warps 2..5  gather the <= 128 query vectors of the item into the 128B-swizzled Q slabs (after the wait: q comes from
//               the qkv_post kernel), then softmax: thread r owns query m = r (TMEM lane r), lazy rescale of O in TMEM,
//               P as bf16 into swizzled shared memory; epilogue: normalised partial (O, log2-sum-exp) for rows < n_q.
// Two CTAs per SM (112 KB shared memory, 256 TMEM columns each), like the mma.sync version.
// ---------------------------------------------------------------------------------------------------------------------
namespace tc {

constexpr int TQ = 128;          // query vectors per item (TMEM lanes)
constexpr int TK = 64;
constexpr int THREADS = 192;
constexpr int KV_STAGES = 2;
constexpr int SLAB = 128 * 128;  // [128 rows x 64 bf16] 128B-swizzled K-major slab
constexpr int Q_BYTES = 2 * SLAB;
constexpr int K_BYTES = 2 * 8192;   // two slabs of [64 keys x 64 d], K-major (k = d)
constexpr int V_BYTES = 2 * 8192;   // same tiles used MN-major (n = d, k = keys)
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + Q_BYTES;
constexpr int OFF_V = OFF_K + KV_STAGES * K_BYTES;
constexpr int OFF_P = OFF_V + KV_STAGES * V_BYTES;
constexpr int OFF_BAR = OFF_P + SLAB;
constexpr int SMEM = OFF_BAR + 256;
constexpr int TMEM_COLS = 256;
constexpr int O_COL = 128;

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
SB_DEVICE void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
SB_DEVICE float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// tensor maps: prompt K/V of both groups ([P][nkv*HD]) and the completion caches seen as [R*c_max][nkv*HD]
struct Maps { CUtensorMap kp[2], vp[2], kc, vc; };

__global__ void __launch_bounds__(THREADS, 2)
dec_attn_tc_kernel(const __grid_constant__ Maps maps, const DecAttnParams p, int c_max) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  if ((sbase & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  const uint32_t q_full = smem_u32(bars + 0);
  const uint32_t kv_full0 = smem_u32(bars + 1);      // [2]
  const uint32_t kv_empty0 = smem_u32(bars + 3);     // [2]
  const uint32_t s_full0 = smem_u32(bars + 5);       // [2]
  const uint32_t p_full = smem_u32(bars + 7);
  const uint32_t pv_done = smem_u32(bars + 8);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const Plan& pl = p.pl;
  const int rep = pl.rep;

  pdl_launch_dependents();
  const int tr = (blockIdx.x == 0 && threadIdx.x == 128) ? sb_trace_begin(SB_TR_ATTN) : -1;   // warp 4 = TMEM quarter 0 (rows 0-31)

  // ---- decode the item (every thread; own items read the step counter before the wait, see dec_attn_kernel)
  int item = blockIdx.x;
  const bool is_prefix = item < pl.n_prefix_items;
  int j_lo, j_hi, slot, kvh, row0, n_q, m0 = 0, grp = 0, j_new = -1;
  long long kv_row0 = 0;     // first row of this item's cache in the tensor map
  if (is_prefix) {
    const int per_group0 = pl.n_qb[0] * p.nkv * pl.n_psplit;
    if (item >= per_group0) { grp = 1; item -= per_group0; }
    const int s = item % pl.n_psplit;
    kvh = (item / pl.n_psplit) % p.nkv;
    const int qb = item / (pl.n_psplit * p.nkv);
    row0 = grp == 0 ? 0 : pl.n_rows[0];
    m0 = qb * pl.rows_qb * rep;
    n_q = min(pl.rows_qb, pl.n_rows[grp] - qb * pl.rows_qb) * rep;
    j_lo = s * pl.p_chunk;
    j_hi = min(j_lo + pl.p_chunk, p.P);
    slot = s;
  } else {
    item -= pl.n_prefix_items;
    const int s = item % pl.n_csplit;
    kvh = (item / pl.n_csplit) % p.nkv;
    row0 = item / (pl.n_csplit * p.nkv);
    n_q = rep;
    const int n_ctx = *p.step_ptr + 1;
    int per = (n_ctx + pl.n_csplit - 1) / pl.n_csplit;
    per = (per + TK - 1) / TK * TK;
    j_lo = s * per;
    j_hi = min(j_lo + per, n_ctx);
    slot = pl.n_psplit + s;
    j_new = n_ctx - 1;
    kv_row0 = (long long)row0 * c_max;
  }
  const int n_tiles = j_hi > j_lo ? (j_hi - j_lo + TK - 1) / TK : 0;

  if (warp == 0 && lane == 0) {
    mbar_init(q_full, TQ);
    for (int s = 0; s < KV_STAGES; ++s) { mbar_init(kv_full0 + 8 * s, 1); mbar_init(kv_empty0 + 8 * s, 1); }
    mbar_init(s_full0, 1); mbar_init(s_full0 + 8, 1);
    mbar_init(p_full, TQ);
    mbar_init(pv_done, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      const CUtensorMap* mk = is_prefix ? &maps.kp[grp] : &maps.kc;
      const CUtensorMap* mv = is_prefix ? &maps.vp[grp] : &maps.vc;
      tma_prefetch_desc(mk); tma_prefetch_desc(mv);
      int stage = 0;
      uint32_t phase = 0;
      bool waited = false;
      for (int t = 0; t < n_tiles; ++t) {
        const int k0 = j_lo + t * TK;
        if (!waited && j_new >= k0 && j_new < k0 + TK) {   // the tile with the current token's row: written by qkv_post
          pdl_wait();
          waited = true;
        }
        mbar_wait(kv_empty0 + 8 * stage, phase ^ 1);
        const uint32_t fb = kv_full0 + 8 * stage;
        mbar_expect_tx(fb, K_BYTES + V_BYTES);
        const uint32_t sk = sbase + OFF_K + stage * K_BYTES;
        const uint32_t sv = sbase + OFF_V + stage * V_BYTES;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          tma_load_2d(sk + c * 8192, mk, fb, kvh * HD + c * 64, (int)(kv_row0 + k0));
          tma_load_2d(sv + c * 8192, mv, fb, kvh * HD + c * 64, (int)(kv_row0 + k0));
        }
        if (++stage == KV_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0 && n_tiles > 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(TQ, TK, false, false);
      constexpr uint32_t idesc_o = umma_idesc_bf16(TQ, HD, false, true);
      auto issue_s = [&](int stage, int sbuf) {
        const uint32_t sq = sbase + OFF_Q;
        const uint32_t sk = sbase + OFF_K + stage * K_BYTES;
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
      mbar_wait(kv_full0, 0);
      tc_fence_after();
      issue_s(0, 0);
      for (int it = 0; it < n_tiles; ++it) {
        const int stage = it & 1;
        if (it + 1 < n_tiles) {   // S of the next tile overlaps the softmax of this one
          const int nt = it + 1;
          mbar_wait(kv_full0 + 8 * (nt & 1), (nt >> 1) & 1);
          tc_fence_after();
          issue_s(nt & 1, nt & 1);
        }
        mbar_wait(p_full, it & 1);
        tc_fence_after();
        const uint32_t sp = sbase + OFF_P;
        const uint32_t sv = sbase + OFF_V + stage * V_BYTES;
#pragma unroll
        for (int kk = 0; kk < TK / 16; ++kk) {
          const uint64_t adesc = umma_desc_sw128(sp, 0, 1024) + (uint64_t)(kk * 2);
          const uint64_t bdesc = umma_desc_sw128(sv, 8192, 1024) + (uint64_t)(kk * (2048 >> 4));
          tc_mma_bf16(tmem_base + O_COL, adesc, bdesc, idesc_o, (it > 0 || kk > 0) ? 1u : 0u);
        }
        tc_commit(kv_empty0 + 8 * stage);
        tc_commit(pv_done);
      }
    }
  } else {
    // ------------------------------ q gather, softmax, epilogue (4 warps, thread = query vector) ----------------------
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    pdl_wait();
    sb_trace_mark(tr, 1);
    // Q slabs: vector m, 16-byte chunk c (8 dims) -> slab c / 8, row m, chunk (c % 8) ^ (m % 8)   (128B swizzle).
    // Only the n_q real vectors are gathered (thread = one chunk column of 8 rows m, m + 8, ...): the other rows of the
    // 128-row MMA tile keep whatever the shared memory held -- a row of S, P and O depends on the same row of Q only,
    // and rows >= n_q are never stored.
    {
      const int c = r & 15;
      const uint32_t dst0 = sbase + OFF_Q + (c >> 3) * SLAB;
      for (int m = r >> 4; m < n_q; m += 8) {
        const int mm = m0 + m;
        const int qr = mm / rep, qh = mm - qr * rep;
        const bf16* src = p.q + (long long)(row0 + qr) * p.nh * HD + (long long)(kvh * rep + qh) * HD + c * 8;
        cp_async16(dst0 + m * 128 + (((c & 7) ^ (m & 7)) << 4), src, true);
      }
    }
    cp_async_commit();
    cp_async_wait<0>();
    fence_proxy_async();
    mbar_arrive(q_full);

    float m_ref = -INFINITY, l = 0.f;
    uint8_t* sP = smem + OFF_P;
    // a warp whose 32 rows are all beyond n_q has no softmax to do (own items use 7 of the 128 rows): it only keeps the
    // barrier phases in step, which leaves the exp2 units to the warps with real rows
    const bool warp_real = quarter * 32 < n_q;
    for (int it = 0; it < n_tiles; ++it) {
      const int sbuf = it & 1;
      mbar_wait(s_full0 + 8 * sbuf, (it >> 1) & 1);
      tc_fence_after();
      if (!warp_real) {
        if (it > 0) mbar_wait(pv_done, (it - 1) & 1);
        mbar_arrive(p_full);
        continue;
      }
      const int j0 = j_lo + it * TK;
      const uint32_t s_addr = lane_addr + sbuf * TK;
      uint32_t sv_[TK];
      tmem_ld_32x32(s_addr, sv_);
      tmem_ld_32x32(s_addr + 32, sv_ + 32);
      tmem_ld_wait();
      float mx = -INFINITY;
      if (j0 + TK <= j_hi) {
#pragma unroll
        for (int j = 0; j < TK; ++j) mx = fmaxf(mx, __uint_as_float(sv_[j]));
      } else {
#pragma unroll
        for (int j = 0; j < TK; ++j) {
          if (j0 + j >= j_hi) sv_[j] = 0xff800000u;   // -inf: beyond this item's key range
          mx = fmaxf(mx, __uint_as_float(sv_[j]));
        }
      }
      mx *= p.scale_log2;
      float factor = 1.f;
      if (mx > m_ref + 8.f) {     // lazy rescale: the reference max moves only when it grows by more than 2^8
        factor = (m_ref == -INFINITY) ? 0.f : fast_exp2(m_ref - mx);
        m_ref = mx;
      }
      if (it > 0) {
        mbar_wait(pv_done, (it - 1) & 1);
        tc_fence_after();
        // (lanes beyond n_q hold whatever their never-written Q row produced: they must not trigger the rescale)
        if (__any_sync(0xffffffffu, r < n_q && factor != 1.f)) {
#pragma unroll 1
          for (int c = 0; c < HD / 32; ++c) {
            uint32_t v[32];
            tmem_ld_32x32(lane_addr + O_COL + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) * factor);
            tmem_st_32x32(lane_addr + O_COL + c * 32, v);
          }
          tmem_st_wait();
        }
      }
      l *= factor;
      const float neg_m = (m_ref == -INFINITY) ? 0.f : -m_ref;
      float rs0 = 0.f, rs1 = 0.f;
      uint8_t* rowp = sP + r * 128;
#pragma unroll
      for (int c8 = 0; c8 < TK / 8; ++c8) {
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
      l += rs0 + rs1;
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(p_full);
    }
    // epilogue: normalised partial for the real query vectors of this item
    if (n_tiles > 0) {
      mbar_wait(pv_done, (n_tiles - 1) & 1);
      tc_fence_after();
    }
    if (!warp_real) goto done;
    {
    const bool real = r < n_q;
    const int mm = m0 + (real ? r : 0);
    const int row = row0 + mm / rep, head = kvh * rep + mm % rep;
    const long long pbase = ((long long)row * p.nh + head) * pl.NS + slot;
    const float inv = l > 0.f ? 1.f / l : 0.f;
    float* op = p.o_part + pbase * HD;
#pragma unroll 1
    for (int c = 0; c < HD / 32; ++c) {   // tcgen05.ld is warp-collective: every lane loads, only real rows store
      uint32_t v[32];
      if (n_tiles > 0) {
        tmem_ld_32x32(lane_addr + O_COL + c * 32, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
      if (real) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(op + c * 32 + j) =
              make_float4(__uint_as_float(v[j]) * inv, __uint_as_float(v[j + 1]) * inv, __uint_as_float(v[j + 2]) * inv,
                          __uint_as_float(v[j + 3]) * inv);
      }
    }
    if (real) p.lse_part[pbase] = l > 0.f ? m_ref + log2f(l) : -INFINITY;
    }
  done:
    sb_trace_mark(tr, 2);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

// 2D bf16 map over [rows][cols] (row stride = cols), box [64 rows x 64 columns], 128B swizzle
int make_map(CUtensorMap* m, const void* ptr, uint64_t cols, uint64_t rows) {
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) {
      sb_set_error("sb_dec_attn: cuTensorMapEncodeTiled not available from the driver");
      return 1;
    }
    g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {64, 64};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    sb_set_error("sb_dec_attn: cuTensorMapEncodeTiled failed (%d): ptr=%p cols=%llu rows=%llu", (int)r, ptr,
                 (unsigned long long)cols, (unsigned long long)rows);
    return 1;
  }
  return 0;
}

}  // namespace tc

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
  SB_REQUIRE(n_heads / n_kv_heads <= 128, "sb_dec_attn: at most 128 q heads per kv head");
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
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  static bool done_tc = false;
  if (!done_tc) {
    SB_CUDA(cudaFuncSetAttribute(tc::dec_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM));
    done_tc = true;
  }
  SB_REQUIRE(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k_cache) | reinterpret_cast<uintptr_t>(v_cache) |
               reinterpret_cast<uintptr_t>(kp0) | reinterpret_cast<uintptr_t>(vp0) | reinterpret_cast<uintptr_t>(kp1) |
               reinterpret_cast<uintptr_t>(vp1)) & 15) == 0, "sb_dec_attn: q and the caches must be 16-byte aligned");
  SB_REQUIRE(cache_stride_r == (long long)c_max * n_kv_heads * HD,
             "sb_dec_attn: the completion cache must be contiguous [R][c_max][n_kv_heads*head_dim]");
  tc::Maps maps;
  const uint64_t cols = (uint64_t)n_kv_heads * HD;
  if (tc::make_map(&maps.kc, p.kc, cols, (uint64_t)R * c_max) || tc::make_map(&maps.vc, p.vc, cols, (uint64_t)R * c_max)) return 1;
  for (int g = 0; g < 2; ++g) {
    if (P > 0) {
      if (tc::make_map(&maps.kp[g], p.kp[g], cols, (uint64_t)P) || tc::make_map(&maps.vp[g], p.vp[g], cols, (uint64_t)P)) return 1;
    } else {
      maps.kp[g] = maps.kc; maps.vp[g] = maps.vc;   // no prefix items: never dereferenced
    }
  }
  SB_CUDA(sb_launch(tc::dec_attn_tc_kernel, dim3(p.pl.n_items), dim3(tc::THREADS), (size_t)tc::SMEM, st, sb_pdl_enabled(), maps, p,
                    c_max));
  if (sb_check_launch("sb_dec_attn")) return 1;
  const int n_pairs = R * n_heads;
  SB_CUDA(sb_launch(dec_attn_combine_kernel, dim3((n_pairs + 3) / 4), dim3(128), 0, st, sb_pdl_enabled(),
                    (const float*)p.o_part, (const float*)p.lse_part, p.pl.NS, n_pairs, (bf16*)out));
  return sb_check_launch("sb_dec_attn(combine)");
}
