// Does the decode GEMV's per-CTA ingest ceiling (41-48 GB/s, profiles/r01_ingest_labs.md) come from the SERIAL issue work
// of its single producer thread and single MMA thread (mbarrier try_wait ~90 cyc + expect_tx + 2 TMA issues; try_wait +
// fence + 4 tcgen05.mma + commit per 16 KB stage)?  This lab runs P independent producer/consumer pipelines inside one CTA
// (each with its own sub-ring of stages and its own TMEM accumulator; the swap-AB stage of sb_gemm's F32T path: 16 KB
// weight box + 2 KB activation box, 4 x tcgen05.mma M=128 N=16 K=16) and reports GB/s of weights per CTA.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I spacer_b200/csrc -I include \
//        -o tools/labs/ingest_dual_lab tools/labs/ingest_dual_lab.cu -lcuda
#include "common.cuh"
#include <cudaTypedefs.h>
#include <cstdio>
#include <cstdlib>

void sb_set_error(const char*, ...) {}
int sb_check_launch(const char*) { return 0; }
bool sb_pdl_enabled() { return false; }

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int W_BYTES = 16384;     // weight tile per stage
constexpr int X_BYTES = 2048;      // activation tile per stage (16 rows x 64 k)
constexpr int STAGE = W_BYTES + X_BYTES;

// P pipelines, NST stages in total (NST % P == 0).  mma = 0: the consumer frees the stage at once (load path only).
template <int P>
__global__ void __launch_bounds__(64 * P, 1)
lab_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX, int NST, int mma, int n_tiles,
           int k_tiles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NST * STAGE);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + NST), sbase = smem_u32(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NST);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    mbar_fence_init();
  }
  if (warp == P) tmem_alloc(smem_u32(tmem_slot), 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int per = NST / P;                       // stages per pipeline
  if (warp < P && lane == 0) {                   // producer of pipeline `warp`
    const int pl = warp;
    int si = 0; uint32_t phase = 0; int i = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++i) {
      if (i % P != pl) continue;
      const int stage = pl + P * si;
      const int mt = t / k_tiles, kt = t % k_tiles;
      mbar_wait(empty0 + 8 * stage, phase ^ 1);
      mbar_expect_tx(full0 + 8 * stage, W_BYTES + X_BYTES);
      tma_load_2d(sbase + stage * STAGE, &tmW, full0 + 8 * stage, kt * 64, mt * 128);
      tma_load_2d(sbase + stage * STAGE + W_BYTES, &tmX, full0 + 8 * stage, kt * 64, 0);
      if (++si == per) { si = 0; phase ^= 1; }
    }
  } else if (warp >= P && warp < 2 * P && lane == 0) {   // consumer of pipeline warp - P
    const int pl = warp - P;
    constexpr uint32_t idesc = umma_idesc_bf16(128, 16, false, false);
    const uint32_t d_tmem = tmem_base + 16 * pl;
    int si = 0; uint32_t phase = 0; bool first = true; int i = 0, n = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++i) {
      if (i % P != pl) continue;
      const int stage = pl + P * si;
      mbar_wait(full0 + 8 * stage, phase);
      tc_fence_after();
      const uint32_t sw = sbase + stage * STAGE, sx = sw + W_BYTES;
      if (!mma) {
        mbar_arrive(empty0 + 8 * stage);
      } else {
        const uint64_t wdesc = umma_desc_sw128(sw, 0, 1024), xdesc = umma_desc_sw128(sx, 0, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma_bf16(d_tmem, wdesc + (uint64_t)(k * 2), xdesc + (uint64_t)(k * 2), idesc, (first && k == 0) ? 0u : 1u);
        first = false;
        tc_commit(empty0 + 8 * stage);
      }
      ++n;
      if (++si == per) { si = 0; phase ^= 1; }
    }
    if (mma && n > 0) {       // let the last MMAs retire before TMEM is released
      const int last_si = (n - 1) % per;
      mbar_wait(empty0 + 8 * (pl + P * last_si), ((n - 1) / per) & 1);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == P) tmem_dealloc(tmem_base, 128);
}

template <int P>
void run(const CUtensorMap& tmW, const CUtensorMap& tmX, int n_tiles, int k_tiles, cudaEvent_t e0, cudaEvent_t e1) {
  const int grids[] = {148, 112, 74, 56, 37};
  for (int NST : {8, 12})
    for (int mma = 1; mma >= 0; --mma) {
      if (NST % P) continue;
      const size_t smem = (size_t)NST * STAGE + 256;
      CK(cudaFuncSetAttribute(lab_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      for (int g : grids) {
        lab_kernel<P><<<g, 64 * P, smem>>>(tmW, tmX, NST, mma, n_tiles, k_tiles);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int i = 0; i < 3; ++i) lab_kernel<P><<<g, 64 * P, smem>>>(tmW, tmX, NST, mma, n_tiles, k_tiles);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 3;
        const double gbs = (double)n_tiles * W_BYTES / (ms * 1e-3) / 1e9;
        printf("{\"pipelines\": %d, \"stages\": %d, \"mma\": %d, \"ctas\": %d, \"ms\": %.3f, \"weight_gbs\": %.0f, \"gbs_per_cta\": %.1f}\n",
               P, NST, mma, g, ms, gbs, gbs / g);
        fflush(stdout);
      }
    }
}

int main() {
  const int M = 37888, K = 3584, n_copies = 4;
  const size_t bytes = (size_t)M * K * 2;
  uint8_t* buf; CK(cudaMalloc(&buf, bytes * n_copies)); CK(cudaMemset(buf, 0, bytes * n_copies));
  uint8_t* xb; CK(cudaMalloc(&xb, (size_t)16 * K * 2)); CK(cudaMemset(xb, 0, (size_t)16 * K * 2));
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  auto enc = (PFN_cuTensorMapEncodeTiled_v12000)fn;
  auto mk = [&](void* p, uint64_t inner, uint64_t outer, uint32_t box_rows) {
    CUtensorMap m; cuuint64_t dims[2] = {inner, outer}; cuuint64_t str[1] = {inner * 2}; cuuint32_t box[2] = {64, box_rows}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
    return m;
  };
  const int k_tiles = K / 64, rows_total = M * n_copies, n_tiles = (rows_total / 128) * k_tiles;
  CUtensorMap tmW = mk(buf, K, rows_total, 128), tmX = mk(xb, K, 16, 16);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  run<1>(tmW, tmX, n_tiles, k_tiles, e0, e1);
  run<2>(tmW, tmX, n_tiles, k_tiles, e0, e1);
  run<4>(tmW, tmX, n_tiles, k_tiles, e0, e1);
  return 0;
}
