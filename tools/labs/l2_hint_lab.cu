// Can the HBM-idle window of the decode step's attention phase be used to pull part of the NEXT big weight matrix (gate|up,
// 271 MB) into L2 so that the gate|up GEMV then runs above the HBM rate?  Round 1/2 tried `cp.async.bulk.prefetch.L2` of the
// contiguous HEAD of the matrix and saw nothing.  Two suspects, both tested here:
//   (1) layout: a contiguous head belongs to the first tiles of a few CTAs only, and a CTA that reads an L2-resident run
//       followed by an HBM run never overlaps the two sources.  Alternative: prefetch every n-th K block of every tile
//       (tensor-map box prefetch), so each CTA's stream mixes L2 hits into a saturated HBM stream.
//   (2) eviction: between the prefetch and its use ~70 MB of other traffic (o weights, KV, the ring prefill) passes through
//       L2 and the prefetched lines are the oldest.  Alternative: L2 cache hints -- evict_first on the streams,
//       evict_last on the prefetch.
// Sequence per trial: flush (stream 800 MB) -> prefetch kernel -> interference stream (X MB, TMA) -> timed consumer (the
// 2-pipeline swap-AB GEMV stage loop of sb_gemm's F32T path over the 271 MB matrix, CTA b owns tiles b and b + 148).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I spacer_b200/csrc -I include \
//        -o tools/labs/l2_hint_lab tools/labs/l2_hint_lab.cu -lcuda
#include "common.cuh"
#include <cudaTypedefs.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

void sb_set_error(const char*, ...) {}
int sb_check_launch(const char*) { return 0; }
bool sb_pdl_enabled() { return false; }

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int W_BYTES = 16384, X_BYTES = 2048, STAGE = W_BYTES + X_BYTES;
constexpr int P = 2, NST = 10;

__device__ __forceinline__ uint64_t make_policy(int kind) {
  uint64_t pol = 0;
  if (kind == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  else if (kind == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  else if (kind == 3) asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* m, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_2d_hint(const CUtensorMap* m, int c0, int c1, uint64_t pol) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile.L2::cache_hint [%0, {%1, %2}], %3;" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(c0), "r"(c1), "l"(pol) : "memory");
}

// consumer: CTA b streams m-tiles mt0 + b, mt0 + b + gridDim, ... (< mt0 + n_mt) of the matrix, K blocks in order,
// alternating between the two pipelines.  stamps[b] = {first TMA issue, end}
__global__ void __launch_bounds__(64 * P, 1)
consumer(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX, int mt0, int n_mt, int k_tiles,
         int policy, unsigned long long* stamps, int ks = 1) {
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
  constexpr int per = NST / P;
  if (warp < P && lane == 0) {
    const int pl = warp;
    const uint64_t pol = make_policy(policy);
    if (pl == 0 && stamps) stamps[2 * blockIdx.x] = sb_gtime();
    int si = 0; uint32_t phase = 0;
    for (int u = blockIdx.x; u < n_mt * ks; u += gridDim.x) {
      const int t = u / ks, kb0 = (u % ks) * (k_tiles / ks), kb1 = kb0 + k_tiles / ks;
      for (int kb = kb0 + pl; kb < kb1; kb += P) {
        const int stage = pl + P * si;
        mbar_wait(empty0 + 8 * stage, phase ^ 1);
        mbar_expect_tx(full0 + 8 * stage, W_BYTES + X_BYTES);
        if (policy) tma_load_2d_hint(sbase + stage * STAGE, &tmW, full0 + 8 * stage, kb * 64, (mt0 + t) * 128, pol);
        else tma_load_2d(sbase + stage * STAGE, &tmW, full0 + 8 * stage, kb * 64, (mt0 + t) * 128);
        tma_load_2d(sbase + stage * STAGE + W_BYTES, &tmX, full0 + 8 * stage, kb * 64, 0);
        if (++si == per) { si = 0; phase ^= 1; }
      }
    }
  } else if (warp >= P && warp < 2 * P && lane == 0) {
    const int pl = warp - P;
    constexpr uint32_t idesc = umma_idesc_bf16(128, 16, false, false);
    const uint32_t d_tmem = tmem_base + 16 * pl;
    int si = 0; uint32_t phase = 0; bool first = true; int n = 0;
    for (int u = blockIdx.x; u < n_mt * ks; u += gridDim.x) {
      const int kb0 = (u % ks) * (k_tiles / ks), kb1 = kb0 + k_tiles / ks;
      for (int kb = kb0 + pl; kb < kb1; kb += P) {
        const int stage = pl + P * si;
        mbar_wait(full0 + 8 * stage, phase);
        tc_fence_after();
        const uint32_t sw = sbase + stage * STAGE, sx = sw + W_BYTES;
        const uint64_t wdesc = umma_desc_sw128(sw, 0, 1024), xdesc = umma_desc_sw128(sx, 0, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma_bf16(d_tmem, wdesc + (uint64_t)(k * 2), xdesc + (uint64_t)(k * 2), idesc, (first && k == 0) ? 0u : 1u);
        first = false;
        tc_commit(empty0 + 8 * stage);
        ++n;
        if (++si == per) { si = 0; phase ^= 1; }
      }
    }
    if (n > 0) {
      const int last_si = (n - 1) % per;
      mbar_wait(empty0 + 8 * (pl + P * last_si), ((n - 1) / per) & 1);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0 && stamps) stamps[2 * blockIdx.x + 1] = sb_gtime();
  if (warp == P) tmem_dealloc(tmem_base, 128);
}

// prefetch K blocks of the tiles CTA b will own.  mode 0: K blocks [k_lo, k_hi) of the first `tiles_per_cta` tiles;
// mode 1: every K block with kb % stride == stride - 1 and kb >= k_lo of the first `tiles_per_cta` tiles;
// mode 2: contiguous head: all K blocks of m-tiles [0, head_tiles) (the old layout), spread over the grid
__global__ void prefetcher(const __grid_constant__ CUtensorMap tmW, int mt0, int n_mt, int k_tiles, int mode, int k_lo, int k_hi,
                           int stride, int tiles_per_cta, int head_tiles, int policy, const uint8_t* base = nullptr) {
  const uint64_t pol = make_policy(policy);
  if (mode == 3) {   // rows [0, k_hi) of every tile this CTA owns: one contiguous run of k_hi * 7168 bytes per tile, 7168-byte pieces
    int j = 0;
    const long long row_bytes = (long long)k_tiles * 128;
    for (int t = blockIdx.x; t < n_mt && j < tiles_per_cta; t += gridDim.x, ++j) {
      const uint8_t* tp = base + (long long)(mt0 + t) * 128 * row_bytes;
      for (int r = threadIdx.x; r < k_hi; r += blockDim.x) {
        if (policy) asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(tp + r * row_bytes), "r"((uint32_t)row_bytes), "l"(pol) : "memory");
        else asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(tp + r * row_bytes), "r"((uint32_t)row_bytes) : "memory");
      }
    }
    return;
  }
  if (mode == 2) {
    const int total = head_tiles * k_tiles;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
      const int t = i / k_tiles, kb = i % k_tiles;
      if (policy) tma_prefetch_2d_hint(&tmW, kb * 64, (mt0 + t) * 128, pol); else tma_prefetch_2d(&tmW, kb * 64, (mt0 + t) * 128);
    }
    return;
  }
  int j = 0;
  for (int t = blockIdx.x; t < n_mt && j < tiles_per_cta; t += gridDim.x, ++j) {
    for (int kb = threadIdx.x; kb < k_tiles; kb += blockDim.x) {
      const bool sel = mode == 0 ? (kb >= k_lo && kb < k_hi) : (kb >= k_lo && (kb % stride) == stride - 1);
      if (!sel) continue;
      if (policy) tma_prefetch_2d_hint(&tmW, kb * 64, (mt0 + t) * 128, pol); else tma_prefetch_2d(&tmW, kb * 64, (mt0 + t) * 128);
    }
  }
}

__global__ void spin_us(int us) {
  const unsigned long long t0 = sb_gtime();
  while (sb_gtime() - t0 < (unsigned long long)us * 1000ull) {}
}

struct Trial { const char* name; int mode, k_lo, k_hi, stride, tiles_per_cta, head_tiles, pf_policy, stream_policy, interf_tiles, interf_policy, spin; };

int main() {
  const int M = 37888, K = 3584, n_copies = 4, k_tiles = K / 64, mt_per_copy = M / 128;
  const size_t bytes = (size_t)M * K * 2;
  uint8_t* buf; CK(cudaMalloc(&buf, bytes * n_copies)); CK(cudaMemset(buf, 0, bytes * n_copies));
  const int other_tiles = 592;    // 543 MB of "other" traffic to draw the interference from
  uint8_t* other; CK(cudaMalloc(&other, (size_t)other_tiles * 128 * K * 2)); CK(cudaMemset(other, 0, (size_t)other_tiles * 128 * K * 2));
  uint8_t* xb; CK(cudaMalloc(&xb, (size_t)16 * K * 2)); CK(cudaMemset(xb, 0, (size_t)16 * K * 2));
  unsigned long long* stamps; CK(cudaMalloc(&stamps, 148 * 2 * 8));
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
  CUtensorMap tmW = mk(buf, K, (uint64_t)M * n_copies, 128), tmO = mk(other, K, (uint64_t)other_tiles * 128, 128), tmX = mk(xb, K, 16, 16);
  const size_t smem = (size_t)NST * STAGE + 256;
  CK(cudaFuncSetAttribute(consumer, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

  // interference of ~70 MB = 76 tiles; 0 = none
  const Trial trials[] = {
      //                         mode k_lo k_hi stride(ks) tpc tiles pfpol strpol interf ipol spin
      // --- how fast does an L2-resident matrix stream (mode -2: `tiles` m-tiles x `stride` K splits, read 4 times)
      {"hot_37x4",                 -2, 0, 0, 4, 0, 37,  0, 0, 0, 0, 0},
      {"hot_74x2",                 -2, 0, 0, 2, 0, 74,  0, 0, 0, 0, 0},
      {"hot_111x1",                -2, 0, 0, 1, 0, 111, 0, 0, 0, 0, 0},
      {"hot_148x1",                -2, 0, 0, 1, 0, 148, 0, 0, 0, 0, 0},
      // --- nothing prefetched
      {"baseline",                 -1, 0, 0, 1, 0, 0,   0, 0, 76, 0, 0},
      {"baseline_evict_first",     -1, 0, 0, 1, 0, 0,   0, 1, 76, 1, 0},
      // --- contiguous head of the matrix (the layout rounds 1-2 tried), every 4th K block of every tile
      {"head53tiles",               2, 0, 0, 1, 0, 53,  0, 0, 76, 0, 0},
      {"head53tiles_nointerf",      2, 0, 0, 1, 0, 53,  0, 0, 0, 0, 12},
      {"head53tiles_pol",           2, 0, 0, 1, 0, 53,  2, 1, 76, 1, 0},
      {"inter4_both_nointerf",      1, 10, 0, 4, 2, 0,  0, 0, 0, 0, 12},
      {"inter4_both",               1, 10, 0, 4, 2, 0,  0, 0, 76, 0, 0},
      {"inter4_both_pol",           1, 10, 0, 4, 2, 0,  2, 1, 76, 1, 0},
      {"inter8_both_pol",           1, 10, 0, 8, 2, 0,  2, 1, 76, 1, 0},
      // --- the first rows of every 128-row tile
      {"rows32_both_nointerf",      3, 0, 32, 1, 2, 0,  0, 0, 0, 0, 14},
      {"rows32_both",               3, 0, 32, 1, 2, 0,  0, 0, 76, 0, 0},
      {"rows32_both_pol",           3, 0, 32, 1, 2, 0,  2, 1, 76, 1, 0},
      {"rows24_both_nointerf",      3, 0, 24, 1, 2, 0,  0, 0, 0, 0, 14},
      {"rows24_both",               3, 0, 24, 1, 2, 0,  0, 0, 76, 0, 0},
      {"rows24_both_pol",           3, 0, 24, 1, 2, 0,  2, 1, 76, 1, 0},
      {"rows16_both_nointerf",      3, 0, 16, 1, 2, 0,  0, 0, 0, 0, 14},
      {"rows16_both",               3, 0, 16, 1, 2, 0,  0, 0, 76, 0, 0},
      {"rows16_both_pol",           3, 0, 16, 1, 2, 0,  2, 1, 76, 1, 0},
      {"rows48_both_nointerf",      3, 0, 48, 1, 2, 0,  0, 0, 0, 0, 20},
      {"rows64_first_nointerf",     3, 0, 64, 1, 1, 0,  0, 0, 0, 0, 14},
      {"rows64_first_pol",          3, 0, 64, 1, 1, 0,  2, 1, 76, 1, 0},
      {"rows32_both_pol_int38",     3, 0, 32, 1, 2, 0,  2, 1, 38, 1, 0},
  };
  int copy = 0, ocur = 0;
  for (const Trial& tr : trials) {
    std::vector<double> dur, med;
    double pf_mb = 0;
    for (int rep = 0; rep < 7; ++rep) {
      copy = (copy + 1) % n_copies;
      const int mt0 = copy * mt_per_copy;
      if (tr.mode == -2) {
        // hot: a 37 MB matrix (40 tiles... use 148 tiles x 14 K blocks is not expressible; take 37 tiles = 34 MB) read twice
        for (int w = 0; w < 3; ++w) consumer<<<tr.head_tiles * tr.stride, 64 * P, smem>>>(tmW, tmX, mt0, tr.head_tiles, k_tiles, tr.stream_policy, nullptr, tr.stride);
        consumer<<<tr.head_tiles * tr.stride, 64 * P, smem>>>(tmW, tmX, mt0, tr.head_tiles, k_tiles, tr.stream_policy, stamps, tr.stride);
      } else {
        // flush: the three other copies (813 MB)
        for (int c = 1; c < n_copies; ++c)
          consumer<<<148, 64 * P, smem>>>(tmW, tmX, ((copy + c) % n_copies) * mt_per_copy, mt_per_copy, k_tiles, 0, nullptr);
        if (tr.mode >= 0)
          prefetcher<<<148, 64>>>(tmW, mt0, mt_per_copy, k_tiles, tr.mode, tr.k_lo, tr.k_hi, tr.stride, tr.tiles_per_cta, tr.head_tiles, tr.pf_policy, buf);
        if (tr.spin) spin_us<<<1, 1>>>(tr.spin);
        if (tr.interf_tiles) {
          if (ocur + tr.interf_tiles > other_tiles) ocur = 0;
          consumer<<<148, 64 * P, smem>>>(tmO, tmX, ocur, tr.interf_tiles, k_tiles, tr.interf_policy, nullptr);
          ocur += tr.interf_tiles;
        }
        consumer<<<148, 64 * P, smem>>>(tmW, tmX, mt0, mt_per_copy, k_tiles, tr.stream_policy, stamps);
      }
      CK(cudaDeviceSynchronize());
      unsigned long long h[296];
      CK(cudaMemcpy(h, stamps, sizeof(h), cudaMemcpyDeviceToHost));
      const int n = tr.mode == -2 ? tr.head_tiles * tr.stride : 148;
      unsigned long long s0 = ~0ull, e1 = 0;
      std::vector<unsigned long long> ends;
      for (int b = 0; b < n; ++b) { s0 = std::min(s0, h[2 * b]); e1 = std::max(e1, h[2 * b + 1]); ends.push_back(h[2 * b + 1]); }
      std::sort(ends.begin(), ends.end());
      if (rep >= 2) { dur.push_back((e1 - s0) / 1e3); med.push_back((ends[n / 2] - s0) / 1e3); }
    }
    if (tr.mode == 2) pf_mb = tr.head_tiles * k_tiles * 16384 / 1e6;
    else if (tr.mode == 3) pf_mb = 148.0 * tr.tiles_per_cta * tr.k_hi * k_tiles * 128 / 1e6;
    else if (tr.mode == 0) pf_mb = 148.0 * tr.tiles_per_cta * (tr.k_hi - tr.k_lo) * 16384 / 1e6;
    else if (tr.mode == 1) { int c = 0; for (int kb = tr.k_lo; kb < k_tiles; ++kb) c += (kb % tr.stride) == tr.stride - 1; pf_mb = 148.0 * tr.tiles_per_cta * c * 16384 / 1e6; }
    std::sort(dur.begin(), dur.end()); std::sort(med.begin(), med.end());
    const double mb = tr.mode == -2 ? tr.head_tiles * k_tiles * 16384 / 1e6 : bytes / 1e6;
    printf("{\"trial\": \"%s\", \"prefetch_mb\": %.1f, \"interference_mb\": %.1f, \"consumer_us_median\": %.2f, \"consumer_us_min\": %.2f, "
           "\"median_cta_end_us\": %.2f, \"tb_s\": %.2f}\n", tr.name, pf_mb, tr.interf_tiles * k_tiles * 16384 / 1e6, dur[dur.size() / 2], dur[0],
           med[med.size() / 2], mb / dur[dur.size() / 2]);
    fflush(stdout);
  }
  return 0;
}
