// Per-CTA ingest ceiling of the decode GEMV's main loop WITH its tcgen05 consumer (sm_100a), two operand orientations:
//   mode 0: "swap-AB" (what sb_gemm's F32T path does today): A = weight tile [128 rows x 64 k] (16 KB per stage),
//           B = the 16 activation rows [16 x 64] (2 KB), 4 x tcgen05.mma M=128 N=16 K=16 per stage
//   mode 1: activations as A: A = [64 rows (16 valid) x 64 k], B = weight tile [128 x 64], 4 x tcgen05.mma M=64 N=128 K=16
//   mode 2: no MMA (the consumer frees the stage at once) -- the load path alone
//   mode 3: swap-AB MMAs as in mode 0, but the activation tile is resident (loaded once): ONE TMA per stage
//   mode 4: activations-as-A MMAs as in mode 1 with the resident activation tile
// Same TMA boxes and 8-stage ring in all modes; the weights are streamed from a 1.09 GB buffer (> L2).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I spacer_b200/csrc -I include \
//        -o tools/labs/ingest_mma_lab tools/labs/ingest_mma_lab.cu -lcuda
#include "common.cuh"
#include <cudaTypedefs.h>
#include <cstdio>
#include <cstdlib>

void sb_set_error(const char*, ...) {}
int sb_check_launch(const char*) { return 0; }
bool sb_pdl_enabled() { return false; }

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int NST = 8;
constexpr int W_BYTES = 16384;     // weight tile per stage
constexpr int X_BYTES = 8192;      // activation region per stage (64 rows reserved, 16 loaded)
constexpr int STAGE = W_BYTES + X_BYTES;

__global__ void __launch_bounds__(96, 1)
lab_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX, int mode, int n_tiles, int k_tiles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NST * STAGE);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + NST), sbase = smem_u32(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NST);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0 && lane == 0) {          // producer
    int stage = 0; uint32_t phase = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int mt = t / k_tiles, kt = t % k_tiles;
      mbar_wait(empty0 + 8 * stage, phase ^ 1);
      const bool load_x = mode < 3 || t == (int)blockIdx.x;      // modes 3/4: the activation tile is loaded once
      mbar_expect_tx(full0 + 8 * stage, W_BYTES + (load_x ? 2048 : 0));
      tma_load_2d(sbase + stage * STAGE, &tmW, full0 + 8 * stage, kt * 64, mt * 128);
      if (load_x) tma_load_2d(sbase + stage * STAGE + W_BYTES, &tmX, full0 + 8 * stage, kt * 64, 0);
      if (++stage == NST) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1 && lane == 0) {   // consumer
    constexpr uint32_t idesc_swap = umma_idesc_bf16(128, 16, false, false);
    constexpr uint32_t idesc_actA = umma_idesc_bf16(64, 128, false, false);
    int stage = 0; uint32_t phase = 0; bool first = true;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      mbar_wait(full0 + 8 * stage, phase);
      tc_fence_after();
      const uint32_t sw = sbase + stage * STAGE, sx = (mode >= 3 ? sbase : sw) + W_BYTES;   // resident tile: stage 0's
      if (mode == 2) {
        mbar_arrive(empty0 + 8 * stage);
      } else {
        const uint64_t wdesc = umma_desc_sw128(sw, 0, 1024), xdesc = umma_desc_sw128(sx, 0, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (mode == 0 || mode == 3) tc_mma_bf16(tmem_base, wdesc + (uint64_t)(k * 2), xdesc + (uint64_t)(k * 2), idesc_swap, (first && k == 0) ? 0u : 1u);
          else tc_mma_bf16(tmem_base, xdesc + (uint64_t)(k * 2), wdesc + (uint64_t)(k * 2), idesc_actA, (first && k == 0) ? 0u : 1u);
        }
        first = false;
        tc_commit(empty0 + 8 * stage);
      }
      if (++stage == NST) { stage = 0; phase ^= 1; }
    }
  }
  // let outstanding MMAs finish: the consumer's last commits arrive on `empty` barriers; wait for the final one
  __syncthreads();
  if (warp == 1 && lane == 0 && mode != 2) {
    // total stages issued by this CTA
    int n = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) ++n;
    if (n > 0) {
      const int last = (n - 1) % NST;
      mbar_wait(empty0 + 8 * last, ((n - 1) / NST) & 1);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 128);
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
  const size_t smem = (size_t)NST * STAGE + 256;
  CK(cudaFuncSetAttribute(lab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int grids[] = {148, 112, 74, 37};
  for (int mode = 0; mode < 5; ++mode)
    for (int g : grids) {
      lab_kernel<<<g, 96, smem>>>(tmW, tmX, mode, n_tiles, k_tiles);
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      for (int i = 0; i < 3; ++i) lab_kernel<<<g, 96, smem>>>(tmW, tmX, mode, n_tiles, k_tiles);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 3;
      const double gbs = (double)n_tiles * W_BYTES / (ms * 1e-3) / 1e9;
      printf("{\"mode\": %d, \"ctas\": %d, \"ms\": %.3f, \"weight_gbs\": %.0f, \"gbs_per_cta\": %.1f}\n", mode, g, ms, gbs, gbs / g);
      fflush(stdout);
    }
  return 0;
}
