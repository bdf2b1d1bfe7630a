// tcgen05 / TMEM / TMA GEMM for sm_100a: the one dense-contraction kernel of the SG-RLVR hot path.
//
//   D[M,N] = epilogue( A[M,K] * B[N,K]^T )        bf16 operands, fp32 accumulation in TMEM
//
// It replaces every cuBLAS call the reference reaches through transformers:
//   ViT patch-embed / qkv / proj / fc1 / fc2 / merger   (modeling_qwen2_vl.py:304-337,401-405)
//   Qwen2 q/k/v/o projections and the SwiGLU MLP        (modeling_qwen2_vl.py:502-504,559-593)
//   lm_head                                             (modeling_qwen2_vl.py:1437-1438)
// and their backward GEMMs (dX = dY*W uses an MN-major B operand, dW = dY^T*X uses MN-major A and B),
// and -- with A := weight rows, B := the <=16 decode rows ("swap-AB") plus split-K -- the
// weight-streaming GEMVs of autoregressive decode.
//
// Structure (one CTA per SM, persistent over output tiles, 192 threads):
//   warp 0      : TMA producer  (cp.async.bulk.tensor 2D, 128B swizzle, NSTAGES-deep mbarrier ring)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (128 x BN x 16, cta_group::1)
//   warps 2..5  : epilogue (tcgen05.ld 32x32b -> registers -> fused epilogue -> global), double-buffered
//                 TMEM accumulators so the epilogue of tile i overlaps the main loop of tile i+1.
#include "common.cuh"
#include "spacer_b200.h"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cstdlib>
#include <cstring>

namespace {

#ifndef SB_GEMV_STAGES
#define SB_GEMV_STAGES 10
#endif
constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int GEMM_THREADS = 192;
constexpr int A_STAGE_BYTES = BM * BK * 2;

// Decode GEMVs run SB_GEMV_PIPES independent producer/MMA-issuer thread pairs per CTA.  One pair tops out at ~48 GB/s of
// weights whatever the number of CTAs: its per-stage work is SERIAL (mbarrier try_wait ~90 cycles + expect_tx + two TMA
// issues on one thread; try_wait + fence + four tcgen05.mma + commit on the other) and costs ~0.34 us per 16 KB stage.
// Two pairs on alternating K blocks, each with its own half of the stage ring and its own TMEM accumulator (summed by
// the epilogue), reach 62 GB/s per CTA at 112 CTAs (HBM-bound, 6.9 TB/s) and 82-88 GB/s at <= 74 CTAs
// (tools/labs/ingest_dual_lab.cu, profiles/r02_ingest_dual_lab.md) -- which is what lets the cluster-fused epilogues, that
// could only occupy 108-112 SMs, stream at full HBM speed (they were measured again, stayed slower than the plain
// chain and were removed: profiles/r02_ingest_dual_lab.md).
// Four pipelines on a 12-stage ring for the split-K GEMVs (SB_GEMV_PIPES_F32T = 4, SB_GEMV_STAGES_F32T = 12) were measured
// too: 3.036 ms per decode step against 3.01-3.05 ms with two -- at 144-148 CTAs the stream is HBM-bound either way -- so
// both kinds of GEMV run two.
#ifndef SB_GEMV_PIPES
#define SB_GEMV_PIPES 2
#endif
#ifndef SB_GEMV_PIPES_F32T
#define SB_GEMV_PIPES_F32T 2
#endif
#ifndef SB_GEMV_STAGES_F32T
#define SB_GEMV_STAGES_F32T SB_GEMV_STAGES
#endif
constexpr bool is_dec_epi(int epi) { return epi == SB_EPI_F32T || epi == SB_EPI_F32T_SWIGLU; }
constexpr int gemm_pipes(int epi) { return epi == SB_EPI_F32T ? SB_GEMV_PIPES_F32T : (epi == SB_EPI_F32T_SWIGLU ? SB_GEMV_PIPES : 1); }
// warps: 0 producer, 1 MMA issuer (+ TMEM allocator), 2-5 epilogue, then one (producer, MMA) warp pair per extra pipeline
constexpr int gemm_threads(int epi) { return GEMM_THREADS + 64 * (gemm_pipes(epi) - 1); }

struct GemmParams {
  int M, N, K;
  int m_tiles, n_tiles, k_splits, k_iters, k_iters_per_split;
  void* D;
  long long ldd;
  const bf16* bias;
  const bf16* residual;
  long long ldr;
  bf16* aux;
  long long ldaux;
  const int* targets;
  float2* lse_part;
  float* tgt_logit;
  const float* lse;
  const float* coef;
  int l2_hints;                  // decode GEMVs: the weight stream is tagged evict_first (sb_set_dec_l2_hints)
  int a_3d, b_3d;                // MN-major operand described by ONE 3-D tensor map (all 64-wide chunks of a stage in one TMA)
};

// per-CTA timeline of the decode GEMVs (sb_trace_enable_gemv_ctas): CTA b appends {M ^ (K << 32), t_ready, t_end} to its
// own list buf[b * (1 + 3 * cap)] (element 0 = count).  Shows the spread of the CTAs' finishing times (the kernel-level
// trace only sees block 0).
static __device__ unsigned long long* sb_gemv_cta_trace = nullptr;
static __device__ int sb_gemv_cta_cap = 0;

template <int BN, int EPI>
struct Cfg {
  static constexpr int NP = gemm_pipes(EPI);
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  // gate/up exchange of the fused decode SwiGLU epilogue (the only epilogue with shared-memory scratch)
  static constexpr int XBUF_BYTES = EPI == SB_EPI_F32T_SWIGLU ? 128 * 33 * 4 + 256 : 0;
  static constexpr int FIXED_BYTES = 1024 /*align*/ + 256 /*barriers*/ + XBUF_BYTES;
  // the narrow decode tiles (BN <= 32) take the stages that fit next to the fixed part (227 KB per CTA), up to
  // SB_GEMV_STAGES(_F32T), rounded to a whole number of stages per pipeline; the wide tiles are MMA-bound at 4-8 stages
  static constexpr int NSTAGES_RAW = BN <= 32 ? (227 * 1024 - FIXED_BYTES) / STAGE_BYTES : (196 * 1024) / STAGE_BYTES;
  static constexpr int NSTAGES_CAP = BN <= 32 ? (EPI == SB_EPI_F32T ? SB_GEMV_STAGES_F32T : SB_GEMV_STAGES) : 8;
  static constexpr int NSTAGES_FIT = NSTAGES_RAW > NSTAGES_CAP ? NSTAGES_CAP : NSTAGES_RAW;
  static constexpr int NSTAGES = (NSTAGES_FIT / NP) * NP;
  // 2 accumulators (epilogue overlap) x NP pipelines x BN columns, a power of two >= 32
  static constexpr int TMEM_COLS = (2 * NP * BN <= 32) ? 32 : (2 * NP * BN <= 64) ? 64 : (2 * NP * BN <= 128) ? 128 :
                                   (2 * NP * BN <= 256) ? 256 : 512;
  static constexpr int SMEM_BYTES = NSTAGES * STAGE_BYTES + FIXED_BYTES;
  static_assert(NSTAGES >= NP && SMEM_BYTES <= 227 * 1024, "stage ring does not fit");
};

SB_DEVICE float quick_gelu(float x) { return x / (1.f + __expf(-1.702f * x)); }
SB_DEVICE float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
SB_DEVICE float silu(float x) { return x / (1.f + __expf(-x)); }

// store 8 consecutive bf16 (16 bytes)
SB_DEVICE void st8(bf16* p, const float* v) {
  uint4 u;
  u.x = pack_bf16(v[0], v[1]);
  u.y = pack_bf16(v[2], v[3]);
  u.z = pack_bf16(v[4], v[5]);
  u.w = pack_bf16(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
SB_DEVICE void ld8(const bf16* p, float* v) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}

SB_DEVICE void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }   // the 128 epilogue threads

// this thread's BN accumulator columns of a decode tile, summed over the pipelines that hold a share of it (fixed order)
template <int BN, int NP>
SB_DEVICE void dec_load_acc(uint32_t taddr, int n_pl, uint32_t* r) {
  if constexpr (BN == 16) tmem_ld_32x16(taddr, r);
  else tmem_ld_32x32(taddr, r);
  tmem_ld_wait();
#pragma unroll
  for (int pl = 1; pl < NP; ++pl) {
    if (pl < n_pl) {
      uint32_t r2[BN];
      if constexpr (BN == 16) tmem_ld_32x16(taddr + pl * BN, r2);
      else tmem_ld_32x32(taddr + pl * BN, r2);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < BN; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
    }
  }
}

template <bool A_MN, bool B_MN, int BN, int EPI>
__global__ void __launch_bounds__(gemm_threads(EPI), 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const GemmParams p) {
  using C = Cfg<BN, EPI>;
  constexpr int NSTAGES = C::NSTAGES;
  constexpr int NP = C::NP;                      // producer / MMA-issuer pairs (pipelines) in this CTA
  static_assert(NSTAGES % NP == 0, "the stage ring is split evenly between the pipelines");
  constexpr int NSUB = NSTAGES / NP;             // stages per pipeline: pipeline pl owns stages pl, pl + NP, ...
  // the swap-AB weight-streaming GEMVs of the decode step (PDL launch, early weight ring, evict_first weight stream)
  constexpr bool kDec = is_dec_epi(EPI);
  constexpr int ACC_COLS = BN;                              // TMEM columns of one accumulator
  constexpr int TMEM_COLS = C::TMEM_COLS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NSTAGES * C::STAGE_BYTES);
  const uint32_t full0 = smem_u32(bars);
  const uint32_t empty0 = smem_u32(bars + NSTAGES);
  const uint32_t tfull0 = smem_u32(bars + 2 * NSTAGES);
  const uint32_t tempty0 = smem_u32(bars + 2 * NSTAGES + 2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGES + 4);
  float* xbuf = reinterpret_cast<float*>(smem + NSTAGES * C::STAGE_BYTES + 256);
  unsigned long long* xtime = reinterpret_cast<unsigned long long*>(tmem_slot + 2);   // t_ready of the per-CTA trace
  const uint32_t smem_base = smem_u32(smem);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < NSTAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull0 + 8 * a, NP);        // every pipeline's issuer reports its share of the tile
      mbar_init(tempty0 + 8 * a, 4);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_tiles = p.m_tiles * p.n_tiles * p.k_splits;
  // tile t = (ks * n_tiles + nt) * m_tiles + mt.  Persistent: CTA b takes t = b, b + grid, ...
  const int t_begin = blockIdx.x, t_step = gridDim.x;

  pdl_launch_dependents();
  // pipeline of this warp when it is a producer / an MMA issuer (-1 otherwise)
  const int prod_pl = warp == 0 ? 0 : ((NP > 1 && warp >= 6 && ((warp - 6) & 1) == 0) ? 1 + (warp - 6) / 2 : -1);
  const int mma_pl = warp == 1 ? 0 : ((NP > 1 && warp >= 6 && ((warp - 6) & 1) == 1) ? 1 + (warp - 6) / 2 : -1);
  if (prod_pl >= 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      const int pl = prod_pl;
      // Decode GEMVs (F32T, launched with PDL): the A operand is a weight matrix nobody writes during decode, so the
      // first ring of A tiles is requested BEFORE waiting for the preceding kernels -- the weight stream starts while
      // the small kernel that produces the activations (B operand) is still running.
      int pre = 0;
      int tr = -1;
      uint64_t pol_stream = 0;
      if constexpr (kDec) {
        if (p.l2_hints) pol_stream = l2_policy_evict_first();
        if (blockIdx.x == 0 && pl == 0) { tr = sb_trace_begin(SB_TR_GEMV); *reinterpret_cast<volatile int*>(tmem_slot + 1) = tr; }
        for (int t = t_begin; t < total_tiles && pre < NSUB; t += t_step) {
          const int mt = t % p.m_tiles;
          const int ks = t / (p.m_tiles * p.n_tiles);
          const int kb0 = ks * p.k_iters_per_split;
          const int kb1 = min(kb0 + p.k_iters_per_split, p.k_iters);
          for (int kb = kb0 + pl; kb < kb1 && pre < NSUB; kb += NP, ++pre) {
            const int st = pl + NP * pre;
            const uint32_t fb = full0 + 8 * st;
            mbar_expect_tx(fb, C::STAGE_BYTES);
            if (p.l2_hints) tma_load_2d_hint(smem_base + st * C::STAGE_BYTES, &tmA, fb, kb * BK, mt * BM, pol_stream);
            else tma_load_2d(smem_base + st * C::STAGE_BYTES, &tmA, fb, kb * BK, mt * BM);
          }
        }
      }
      pdl_wait();
      sb_trace_mark(tr, 1);
      if constexpr (kDec) {
        if (pl == 0 && sb_gemv_cta_trace != nullptr) *reinterpret_cast<volatile unsigned long long*>(xtime) = sb_gtime();
      }
      int si = 0;                 // position in this pipeline's sub-ring; stage = pl + NP * si
      uint32_t phase = 0;
      int issued = 0;
      for (int t = t_begin; t < total_tiles; t += t_step) {
        const int mt = t % p.m_tiles;
        const int nt = (t / p.m_tiles) % p.n_tiles;
        const int ks = t / (p.m_tiles * p.n_tiles);
        const int kb0 = ks * p.k_iters_per_split;
        const int kb1 = min(kb0 + p.k_iters_per_split, p.k_iters);
        for (int kb = kb0 + pl; kb < kb1; kb += NP, ++issued) {
          const int stage = pl + NP * si;
          const uint32_t fb = full0 + 8 * stage;
          const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
          const uint32_t sb = sa + A_STAGE_BYTES;
          if (issued >= pre) {
            mbar_wait(empty0 + 8 * stage, phase ^ 1);
            mbar_expect_tx(fb, C::STAGE_BYTES);
            if (!A_MN) {
              if (kDec && p.l2_hints) tma_load_2d_hint(sa, &tmA, fb, kb * BK, mt * BM, pol_stream);
              else tma_load_2d(sa, &tmA, fb, kb * BK, mt * BM);
            } else if (p.a_3d) {     // [BM/64] chunk tiles of [BK][64] in one instruction
              tma_load_3d(sa, &tmA, fb, 0, kb * BK, mt * (BM / 64));
            } else {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j)
                tma_load_2d(sa + j * 8192, &tmA, fb, mt * BM + j * 64, kb * BK);
            }
          }
          if (!B_MN) {
            tma_load_2d(sb, &tmB, fb, kb * BK, nt * BN);
          } else if (p.b_3d) {
            tma_load_3d(sb, &tmB, fb, 0, kb * BK, nt * (BN / 64));
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(sb + j * 8192, &tmB, fb, nt * BN + j * 64, kb * BK);
          }
          if (++si == NSUB) { si = 0; phase ^= 1; }
        }
      }
    }
  } else if (mma_pl >= 0) {
    // ------------------------------ MMA issuer ------------------------------
    pdl_wait();
    if (lane == 0) {
      const int pl = mma_pl;
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN, B_MN);
      // per-UMMA_K(16) start-address advance inside a stage, in 16-byte units
      constexpr uint32_t a_adv = A_MN ? (2048 >> 4) : (32 >> 4);
      constexpr uint32_t b_adv = B_MN ? (2048 >> 4) : (32 >> 4);
      int si = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = t_begin; t < total_tiles; t += t_step) {
        const int ks = t / (p.m_tiles * p.n_tiles);
        const int kb0 = ks * p.k_iters_per_split;
        const int kb1 = min(kb0 + p.k_iters_per_split, p.k_iters);
        mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (acc * NP + pl) * ACC_COLS;   // this pipeline's own accumulator
        for (int kb = kb0 + pl; kb < kb1; kb += NP) {
          const int stage = pl + NP * si;
          mbar_wait(full0 + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
          const uint32_t sb = sa + A_STAGE_BYTES;
          const uint64_t adesc = umma_desc_sw128(sa, A_MN ? 8192 : 0, 1024);
          const uint64_t bdesc = umma_desc_sw128(sb, B_MN ? 8192 : 0, 1024);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            tc_mma_bf16(d_tmem, adesc + (uint64_t)(k * a_adv), bdesc + (uint64_t)(k * b_adv), idesc,
                        (kb > kb0 + pl || k > 0) ? 1u : 0u);
          }
          tc_commit(empty0 + 8 * stage);  // frees the smem slot when these MMAs retire
          if (++si == NSUB) { si = 0; phase ^= 1; }
        }
        // accumulator share complete -> epilogue (a pipeline without a K block in this tile just reports in: the epilogue
        // does not read its accumulator)
        if (kb0 + pl < kb1) tc_commit(tfull0 + 8 * acc);
        else mbar_arrive(tfull0 + 8 * acc);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 2 && warp < 6) {
    // ------------------------------ epilogue (4 warps) ------------------------------
    pdl_wait();
    const int q = warp & 3;  // TMEM lane quarter this warp may touch
    const int row_in_tile = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = t_begin; t < total_tiles; t += t_step) {
      const int mt = t % p.m_tiles;
      const int nt = (t / p.m_tiles) % p.n_tiles;
      const int ks = t / (p.m_tiles * p.n_tiles);
      mbar_wait(tfull0 + 8 * acc, acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * NP * ACC_COLS;
      // pipelines that hold a share of this tile's sum (pipeline pl took K blocks kb0 + pl, kb0 + pl + NP, ...)
      const int n_kb = min(p.k_iters_per_split, p.k_iters - ks * p.k_iters_per_split);
      const int n_pl = n_kb < NP ? n_kb : NP;
      const int row = mt * BM + row_in_tile;
      const bool row_ok = row < p.M;
      const int col_base = nt * BN;

      if constexpr (EPI == SB_EPI_F32T) {
        // split-K / swap-AB partial: out[ks][col][row] fp32 (row = weight row, col = decode row)
        static_assert(BN == 16 || BN == 32, "F32T epilogue is for the narrow decode tiles");
        uint32_t r[BN];
        dec_load_acc<BN, NP>(taddr, n_pl, r);
        float* out = reinterpret_cast<float*>(p.D);
        if (row_ok) {
#pragma unroll
          for (int j = 0; j < BN; ++j) {
            const int col = col_base + j;
            const float v = __uint_as_float(r[j]);
            if (col < p.N) out[((long long)ks * p.N + col) * p.ldd + row] = v;
          }
        }
      } else if constexpr (EPI == SB_EPI_F32T_SWIGLU) {
        // decode gate|up GEMV (swap-AB, no split-K) with SwiGLU fused: the tile's 128 weight rows are [64 gate | 64 up]
        // of the same 64 activation columns, i.e. gate sits in TMEM lanes 0-63 (epilogue warps 0,1) and up in lanes
        // 64-127 (warps 2,3).  Warps 2,3 hand their values over through shared memory; warps 0,1 write
        // act[r][c] = bf16( bf16(silu(bf16 g)) * bf16 u ) for the <= BN decode rows r.
        static_assert(BN == 16 || BN == 32, "decode tile");
        uint32_t r[BN];
        dec_load_acc<BN, NP>(taddr, n_pl, r);
        if (q >= 2) {
#pragma unroll
          for (int j = 0; j < BN; ++j) xbuf[(row_in_tile - 64) * 33 + j] = __uint_as_float(r[j]);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (q < 2) {
          bf16* Dp = reinterpret_cast<bf16*>(p.D);
          const int c = mt * 64 + row_in_tile;   // activation column
#pragma unroll
          for (int j = 0; j < BN; ++j) {
            const int drow = col_base + j;       // decode row
            if (drow < p.N) {
              const float g = bf16_round(__uint_as_float(r[j]));
              const float u = bf16_round(xbuf[row_in_tile * 33 + j]);
              Dp[(long long)drow * p.ldd + c] = __float2bfloat16_rn(bf16_round(silu(g)) * u);
            }
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");   // xbuf is free for the next tile
      } else if constexpr (EPI == SB_EPI_SWIGLU) {
        // weight rows are interleaved [64 gate | 64 up] per 128 output columns of the fused gate/up
        // matrix; D gets silu(gate)*up (N/2 columns), aux (optional) the raw [gate|up] tile.
        static_assert(BN % 128 == 0, "swiglu tile");
        bf16* Dp = reinterpret_cast<bf16*>(p.D);
#pragma unroll 1
        for (int g = 0; g < BN / 128; ++g) {
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {
            uint32_t rg[32], ru[32];
            tmem_ld_32x32(taddr + g * 128 + h * 32, rg);
            tmem_ld_32x32(taddr + g * 128 + 64 + h * 32, ru);
            tmem_ld_wait();
            const int gcol = col_base + g * 128 + h * 32;  // raw column of the gate chunk
            if (row_ok && gcol < p.N) {
              float vg[32], vu[32], o[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                vg[j] = __uint_as_float(rg[j]);
                vu[j] = __uint_as_float(ru[j]);
              }
              if (p.bias) {   // interleaved like the rows: bias[gcol..] = gate bias, bias[gcol + 64..] = up bias
#pragma unroll
                for (int j8 = 0; j8 < 32; j8 += 8) {
                  float bg[8], bu[8];
                  ld8(p.bias + gcol + j8, bg);
                  ld8(p.bias + gcol + 64 + j8, bu);
#pragma unroll
                  for (int j = 0; j < 8; ++j) { vg[j8 + j] += bg[j]; vu[j8 + j] += bu[j]; }
                }
              }
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                vg[j] = bf16_round(vg[j]);
                vu[j] = bf16_round(vu[j]);
                o[j] = bf16_round(silu(vg[j])) * vu[j];
              }
              if (p.aux) {
                bf16* ar = p.aux + (long long)row * p.ldaux;
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                  st8(ar + gcol + j, vg + j);
                  st8(ar + gcol + 64 + j, vu + j);
                }
              }
              bf16* dr = Dp + (long long)row * p.ldd + (col_base / 2 + g * 64 + h * 32);
#pragma unroll
              for (int j = 0; j < 32; j += 8) st8(dr + j, o + j);
            }
          }
        }
      } else if constexpr (EPI == SB_EPI_LMHEAD) {
        // per-row online logsumexp over this N tile of bf16-rounded logits + target-logit gather
        float mx = -INFINITY, sm = 0.f;
        const int tg = row_ok ? p.targets[row] : -1;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_ld_wait();
          const int c0 = col_base + c * 32;
          float v[32];
          float cm = -INFINITY;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v[j] = (c0 + j < p.N) ? bf16_round(__uint_as_float(r[j])) : -INFINITY;
            cm = fmaxf(cm, v[j]);
          }
          if (cm > -INFINITY) {
            const float nm = fmaxf(mx, cm);
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) s += __expf(v[j] - nm);
            sm = sm * __expf(mx - nm) + s;
            mx = nm;
          }
          if (tg >= c0 && tg < c0 + 32) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c0 + j == tg) p.tgt_logit[row] = v[j];
          }
        }
        if (row_ok) p.lse_part[(long long)row * p.n_tiles + nt] = make_float2(mx, sm);
      } else {
        bf16* Dp = reinterpret_cast<bf16*>(p.D);
        float lse = 0.f, coef = 0.f;
        int tg = -1;
        if constexpr (EPI == SB_EPI_DLOGITS) {
          if (row_ok) { lse = p.lse[row]; coef = p.coef[row]; tg = p.targets[row]; }
        }
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_ld_wait();
          const int c0 = col_base + c * 32;
          if (row_ok) {
#pragma unroll
            for (int j8 = 0; j8 < 32; j8 += 8) {
              const int col = c0 + j8;
              if (col < p.N) {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j8 + j]);
                if constexpr (EPI == SB_EPI_DLOGITS) {
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    const float pj = __expf(bf16_round(v[j]) - lse);
                    v[j] = coef * (((col + j) == tg ? 1.f : 0.f) - pj);
                  }
                } else {
                  if (p.bias) {
                    float b[8];
                    ld8(p.bias + col, b);
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] += b[j];
                  }
                  if constexpr (EPI == SB_EPI_QUICKGELU || EPI == SB_EPI_GELU) {
                    if (p.aux) st8(p.aux + (long long)row * p.ldaux + col, v);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                      const float z = bf16_round(v[j]);   // the activation sees the bf16 pre-activation, stored or not
                      v[j] = (EPI == SB_EPI_QUICKGELU) ? quick_gelu(z) : gelu_erf(z);
                    }
                  }
                  if (p.residual) {
                    float b[8];
                    ld8(p.residual + (long long)row * p.ldr + col, b);
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = bf16_round(v[j]) + b[j];
                  }
                }
                st8(Dp + (long long)row * p.ldd + col, v);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (kDec) {
    if (blockIdx.x == 0 && threadIdx.x == 0) sb_trace_mark(*reinterpret_cast<volatile int*>(tmem_slot + 1), 2);
    if (threadIdx.x == 0 && sb_gemv_cta_trace != nullptr) {
      unsigned long long* mine = sb_gemv_cta_trace + (long long)blockIdx.x * (1 + 3 * sb_gemv_cta_cap);
      const unsigned long long k = mine[0];
      if (k < (unsigned long long)sb_gemv_cta_cap) {
        mine[1 + 3 * k] = (unsigned long long)p.M ^ ((unsigned long long)p.K << 32);
        mine[2 + 3 * k] = *reinterpret_cast<volatile unsigned long long*>(xtime);
        mine[3 + 3 * k] = sb_gtime();
        mine[0] = k + 1;
      }
    }
  }
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

int get_encode() {
  if (g_encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) {
    sb_set_error("cuTensorMapEncodeTiled not available from the driver (%s)",
                 e != cudaSuccess ? cudaGetErrorString(e) : "symbol not found");
    return 1;
  }
  g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  return 0;
}

// 2D bf16 tensor map over a row-major [outer, inner] matrix with row stride ld (elements)
int make_tmap(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld,
              uint32_t box_inner, uint32_t box_outer) {
  if (get_encode()) return 1;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides,
                        box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    sb_set_error("cuTensorMapEncodeTiled failed (%d): ptr=%p inner=%llu outer=%llu ld=%llu box=%ux%u",
                 (int)r, ptr, (unsigned long long)inner, (unsigned long long)outer,
                 (unsigned long long)ld, box_inner, box_outer);
    return 1;
  }
  return 0;
}

// 3-D bf16 tensor map over an MN-major operand stored [k_rows][ld] with `mn` (a multiple of 64) contiguous elements per
// row: dims (64 within a chunk, k, chunk) so that ONE box [64][box_k][box_chunks] lands in shared memory as box_chunks
// consecutive 128B-swizzled [box_k][64] tiles -- the layout the UMMA MN-major descriptor walks (LBO = 8192).  The 2-D form
// needs one TMA per 64-wide chunk (6 issues per stage for the weight-gradient GEMMs, which made their single producer
// thread the bottleneck: 73 % tensor pipe against 82-86 % for the K-major forward GEMMs).  Returns 1 (no error set) when
// the driver rejects the descriptor; the caller then falls back to the 2-D form.
int make_tmap_mn3(CUtensorMap* m, const void* ptr, uint64_t mn, uint64_t k_rows, uint64_t ld, uint32_t box_k,
                  uint32_t box_chunks) {
  if (get_encode()) return 1;
  cuuint64_t dims[3] = {64, k_rows, mn / 64};
  cuuint64_t strides[2] = {ld * 2, 128};
  cuuint32_t box[3] = {64, box_k, box_chunks};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 1;
}

// decode GEMVs: L2 eviction-priority hints (sb_set_dec_l2_hints; SB_NO_L2_HINTS in the environment turns them off)
int g_l2_hints = getenv("SB_NO_L2_HINTS") == nullptr ? 1 : 0;

int g_num_sms = 0;
int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

template <bool A_MN, bool B_MN, int BN, int EPI>
int launch(const sb_gemm_args* a, cudaStream_t stream) {
  using C = Cfg<BN, EPI>;
  auto kfn = gemm_kernel<A_MN, B_MN, BN, EPI>;
  static bool attr_done = false;
  if (!attr_done) {
    SB_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_done = true;
  }
  CUtensorMap tmA, tmB;
  GemmParams p;
  p.a_3d = p.b_3d = 0;
  static const bool mn3 = getenv("SB_GEMM_NO_TMA3D") == nullptr;
  if (!A_MN) {
    if (make_tmap(&tmA, a->A, a->K, a->M, a->lda, BK, BM)) return 1;
  } else {
    if (mn3 && a->M % 64 == 0 && make_tmap_mn3(&tmA, a->A, a->M, a->K, a->lda, BK, BM / 64) == 0) p.a_3d = 1;
    else if (make_tmap(&tmA, a->A, a->M, a->K, a->lda, 64, BK)) return 1;
  }
  if (!B_MN) {
    if (make_tmap(&tmB, a->B, a->K, a->N, a->ldb, BK, BN)) return 1;
  } else {
    if (mn3 && BN >= 64 && a->N % 64 == 0 && make_tmap_mn3(&tmB, a->B, a->N, a->K, a->ldb, BK, BN / 64) == 0) p.b_3d = 1;
    else if (make_tmap(&tmB, a->B, a->N, a->K, a->ldb, 64, BK)) return 1;
  }
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.m_tiles = (a->M + BM - 1) / BM;
  p.n_tiles = (a->N + BN - 1) / BN;
  p.k_iters = (a->K + BK - 1) / BK;
  p.k_splits = a->k_splits > 0 ? a->k_splits : 1;
  if (p.k_splits > p.k_iters) p.k_splits = p.k_iters;
  p.k_iters_per_split = (p.k_iters + p.k_splits - 1) / p.k_splits;
  p.k_splits = (p.k_iters + p.k_iters_per_split - 1) / p.k_iters_per_split;  // no empty splits
  p.D = a->D; p.ldd = a->ldd;
  p.bias = reinterpret_cast<const bf16*>(a->bias);
  p.residual = reinterpret_cast<const bf16*>(a->residual); p.ldr = a->ldr;
  p.aux = reinterpret_cast<bf16*>(a->aux); p.ldaux = a->ldaux;
  p.targets = a->targets;
  p.lse_part = reinterpret_cast<float2*>(a->lse_part);
  p.tgt_logit = a->tgt_logit;
  p.lse = a->lse; p.coef = a->coef;
  p.l2_hints = g_l2_hints;
  const int total = p.m_tiles * p.n_tiles * p.k_splits;
  const int grid = total < num_sms() ? total : num_sms();
  const bool pdl = (EPI == SB_EPI_F32T || EPI == SB_EPI_F32T_SWIGLU) && sb_pdl_enabled();
  cudaError_t le = sb_launch(kfn, dim3(grid), dim3(gemm_threads(EPI)), (size_t)C::SMEM_BYTES, stream, pdl, tmA, tmB, p);
  if (le != cudaSuccess) {
    sb_set_error("sb_gemm: launch failed: %s", cudaGetErrorString(le));
    return 1;
  }
  return sb_check_launch("sb_gemm");
}

}  // namespace

SB_DEFINE_TRACE_SETTER(sb_trace_set_gemm)

extern "C" int sb_trace_enable_gemv_ctas(unsigned long long* buf_dev, int capacity_per_cta) {
  SB_REQUIRE((buf_dev == nullptr) == (capacity_per_cta == 0), "sb_trace_enable_gemv_ctas: pass (NULL, 0) to disable");
  SB_CUDA(cudaMemcpyToSymbol(sb_gemv_cta_trace, &buf_dev, sizeof(buf_dev)));
  SB_CUDA(cudaMemcpyToSymbol(sb_gemv_cta_cap, &capacity_per_cta, sizeof(capacity_per_cta)));
  return 0;
}

extern "C" int sb_set_dec_l2_hints(int enable) {
  g_l2_hints = enable ? 1 : 0;
  return 0;
}

extern "C" int sb_gemm_effective_splits(int K, int k_splits) {
  int k_iters = (K + BK - 1) / BK;
  if (k_splits < 1) k_splits = 1;
  if (k_splits > k_iters) k_splits = k_iters;
  int per = (k_iters + k_splits - 1) / k_splits;
  return (k_iters + per - 1) / per;
}

extern "C" int sb_gemm(const sb_gemm_args* a, sb_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(a != nullptr, "sb_gemm: null args");
  SB_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, "sb_gemm: empty problem M=%d N=%d K=%d", a->M, a->N, a->K);
  SB_REQUIRE(a->A && a->B && a->D, "sb_gemm: null operand pointer");
  SB_REQUIRE((reinterpret_cast<uintptr_t>(a->A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->B) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(a->D) & 15) == 0,
             "sb_gemm: operands must be 16-byte aligned");
  SB_REQUIRE(a->lda % 8 == 0 && a->ldb % 8 == 0, "sb_gemm: lda/ldb must be multiples of 8 elements");
  const int e = a->epilogue;
  const bool amn = a->a_mn != 0, bmn = a->b_mn != 0;
  if (e == SB_EPI_F32T) {
    SB_REQUIRE(!amn && !bmn, "sb_gemm: F32T epilogue needs K-major operands");
    SB_REQUIRE(a->N <= 32, "sb_gemm: F32T epilogue is for N<=32 (decode rows), got %d", a->N);
    SB_REQUIRE(a->ldd >= a->M, "sb_gemm: F32T ldd (%lld) < M (%d)", a->ldd, a->M);
    if (a->N <= 16) return launch<false, false, 16, SB_EPI_F32T>(a, stream);
    return launch<false, false, 32, SB_EPI_F32T>(a, stream);
  }
  if (e == SB_EPI_F32T_SWIGLU) {
    SB_REQUIRE(!amn && !bmn, "sb_gemm: F32T_SWIGLU epilogue needs K-major operands");
    SB_REQUIRE(a->N <= 32, "sb_gemm: F32T_SWIGLU epilogue is for N<=32 (decode rows), got %d", a->N);
    SB_REQUIRE(a->M % 128 == 0, "sb_gemm: F32T_SWIGLU needs M %% 128 == 0 (interleaved gate/up rows), got %d", a->M);
    SB_REQUIRE(a->k_splits <= 1, "sb_gemm: F32T_SWIGLU cannot be split along K");
    SB_REQUIRE(a->ldd >= a->M / 2, "sb_gemm: F32T_SWIGLU ldd (%lld) < M/2 (%d)", a->ldd, a->M / 2);
    if (a->N <= 16) return launch<false, false, 16, SB_EPI_F32T_SWIGLU>(a, stream);
    return launch<false, false, 32, SB_EPI_F32T_SWIGLU>(a, stream);
  }
  SB_REQUIRE(a->N % 8 == 0 && a->ldd % 8 == 0, "sb_gemm: N and ldd must be multiples of 8");
  SB_REQUIRE(a->k_splits <= 1, "sb_gemm: split-K only with the F32T epilogue");
  if (e == SB_EPI_LMHEAD) {
    SB_REQUIRE(!amn && !bmn, "sb_gemm: LMHEAD epilogue needs K-major operands");
    SB_REQUIRE(a->targets && a->lse_part && a->tgt_logit, "sb_gemm: LMHEAD needs targets/lse_part/tgt_logit");
    return launch<false, false, 256, SB_EPI_LMHEAD>(a, stream);
  }
  if (e == SB_EPI_DLOGITS) {
    SB_REQUIRE(!amn && !bmn, "sb_gemm: DLOGITS epilogue needs K-major operands");
    SB_REQUIRE(a->targets && a->lse && a->coef, "sb_gemm: DLOGITS needs targets/lse/coef");
    return launch<false, false, 256, SB_EPI_DLOGITS>(a, stream);
  }
  if (e == SB_EPI_SWIGLU) {
    SB_REQUIRE(!amn && !bmn, "sb_gemm: SWIGLU epilogue needs K-major operands");
    SB_REQUIRE(a->N % 128 == 0, "sb_gemm: SWIGLU needs N %% 128 == 0 (interleaved gate/up), got %d", a->N);
    return launch<false, false, 256, SB_EPI_SWIGLU>(a, stream);
  }
  if (e == SB_EPI_QUICKGELU) {
    SB_REQUIRE(!amn && !bmn, "sb_gemm: activation epilogues need K-major operands");
    return launch<false, false, 256, SB_EPI_QUICKGELU>(a, stream);
  }
  if (e == SB_EPI_GELU) {
    SB_REQUIRE(!amn && !bmn, "sb_gemm: activation epilogues need K-major operands");
    return launch<false, false, 256, SB_EPI_GELU>(a, stream);
  }
  SB_REQUIRE(e == SB_EPI_STORE, "sb_gemm: unknown epilogue %d", e);
  if (amn) SB_REQUIRE(a->M % 8 == 0, "sb_gemm: MN-major A needs M %% 8 == 0");
  if (!amn && !bmn) {
    if (a->N <= 64 || a->bn == 64) return launch<false, false, 64, SB_EPI_STORE>(a, stream);
    if (a->N <= 128 || a->bn == 128) return launch<false, false, 128, SB_EPI_STORE>(a, stream);
    return launch<false, false, 256, SB_EPI_STORE>(a, stream);
  }
  if (!amn && bmn) {
    if (a->N <= 128 || a->bn == 128) return launch<false, true, 128, SB_EPI_STORE>(a, stream);
    return launch<false, true, 256, SB_EPI_STORE>(a, stream);
  }
  if (amn && bmn) {
    if (a->N <= 128 || a->bn == 128) return launch<true, true, 128, SB_EPI_STORE>(a, stream);
    return launch<true, true, 256, SB_EPI_STORE>(a, stream);
  }
  if (a->N <= 128 || a->bn == 128) return launch<true, false, 128, SB_EPI_STORE>(a, stream);
  return launch<true, false, 256, SB_EPI_STORE>(a, stream);
}
