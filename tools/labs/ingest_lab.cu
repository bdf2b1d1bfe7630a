// Per-SM HBM/L2 ingest microbenchmark (sm_100a): how many GB/s can ONE CTA pull with TMA, as a function of the box shape
// and of how many CTAs share the chip?  Motivation (profiles/r01_decode_fused_epilogues.md): the decode GEMVs never
// exceed ~48 GB/s per CTA, which is why every fusion that costs CTAs loses.  No MMA here: a consumer thread releases
// each stage as soon as it lands, so the number is the ingest ceiling of the load path alone.
//   mode 0: 2D tensor map over a row-major [M][K] bf16 matrix, box 128 rows x 64 cols, 128B swizzle (the GEMV's A load)
//   mode 1: same box over a matrix stored tile-contiguously ([M/128][K/64] tiles of 16 KB: rows of a box are adjacent)
//   mode 2: 1D bulk copies (cp.async.bulk) of 16 KB contiguous chunks
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/labs/ingest_lab tools/labs/ingest_lab.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void tma_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

constexpr int STAGE = 16384;

// each CTA streams tiles t = blockIdx.x, blockIdx.x + grid, ... of n_tiles 16 KB tiles
template <int NST>
__global__ void __launch_bounds__(64, 1)
ingest_kernel(const __grid_constant__ CUtensorMap tm, const uint8_t* base, int mode, int n_tiles, int k_tiles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NST * STAGE);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + NST), sbase = smem_u32(smem);
  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {          // producer
    int stage = 0; uint32_t phase = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      mbar_wait(empty0 + 8 * stage, phase ^ 1);
      mbar_expect_tx(full0 + 8 * stage, STAGE);
      if (mode == 2) bulk_1d(sbase + stage * STAGE, base + (size_t)t * STAGE, STAGE, full0 + 8 * stage);
      else {
        // like the GEMV: a CTA walks along K inside one 128-row block before moving on
        const int mt = t / k_tiles, kt = t % k_tiles;
        if (mode == 0) tma_2d(sbase + stage * STAGE, &tm, full0 + 8 * stage, kt * 64, mt * 128);
        else tma_2d(sbase + stage * STAGE, &tm, full0 + 8 * stage, 0, t * 128);
      }
      if (++stage == NST) { stage = 0; phase ^= 1; }
    }
  } else if (threadIdx.x == 32) {  // consumer: frees the stage at once
    int stage = 0; uint32_t phase = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      mbar_wait(full0 + 8 * stage, phase);
      mbar_arrive(empty0 + 8 * stage);
      if (++stage == NST) { stage = 0; phase ^= 1; }
    }
  }
}

int main() {
  const int M = 37888, K = 3584;                 // the gate|up matrix of one layer: 271.6 MB of bf16
  const size_t bytes = (size_t)M * K * 2;
  const int n_copies = 4;                         // > L2: stream 1.09 GB per launch
  uint8_t* buf; CK(cudaMalloc(&buf, bytes * n_copies)); CK(cudaMemset(buf, 1, bytes * n_copies));
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  auto enc = (PFN_cuTensorMapEncodeTiled_v12000)fn;
  auto mk = [&](uint64_t inner, uint64_t outer, uint64_t ld_elems) {
    CUtensorMap m; cuuint64_t dims[2] = {inner, outer}; cuuint64_t str[1] = {ld_elems * 2}; cuuint32_t box[2] = {64, 128}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
    return m;
  };
  const int k_tiles = K / 64, rows_total = M * n_copies;
  const int n_tiles = (rows_total / 128) * k_tiles;
  CUtensorMap tm_row = mk(K, rows_total, K);                      // row-major [rows][K]
  CUtensorMap tm_tile = mk(64, (uint64_t)n_tiles * 128, 64);      // tile-contiguous: [tiles*128][64]
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto run = [&](auto kern, int nst, int mode, int grid) {
    const size_t smem = (size_t)nst * STAGE + 256;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const CUtensorMap& tm = mode == 1 ? tm_tile : tm_row;
    kern<<<grid, 64, smem>>>(tm, buf, mode, n_tiles, k_tiles);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 3; ++i) kern<<<grid, 64, smem>>>(tm, buf, mode, n_tiles, k_tiles);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 3;
    const double gbs = (double)n_tiles * STAGE / (ms * 1e-3) / 1e9;
    printf("{\"mode\": %d, \"stages\": %d, \"ctas\": %d, \"ms\": %.3f, \"gbs\": %.0f, \"gbs_per_cta\": %.1f}\n", mode, nst, grid, ms, gbs, gbs / grid);
  };
  const int grids[] = {148, 112, 74, 37, 16};
  for (int mode = 0; mode < 3; ++mode)
    for (int g : grids) {
      run(ingest_kernel<8>, 8, mode, g);
      run(ingest_kernel<12>, 12, mode, g);
    }
  return 0;
}
