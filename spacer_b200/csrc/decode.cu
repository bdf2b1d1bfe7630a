// Small kernels of the autoregressive decode step (q_len = 1, R <= 32 rows that share one or two prompts).
// The weight-streaming GEMVs are sb_gemm (swap-AB, split-K, fp32 partials [S][R][N]); the kernels here
// consume those partials: residual + RMSNorm, qkv bias + M-RoPE + KV append, SwiGLU.  (Attention: dec_attn.cu.)
// replaces the q_len=1 iterations of GenerationMixin._sample (generation/utils.py:2743-2806) through
// Qwen2VLDecoderLayer (MQ2:597-662) and DynamicCache.update's torch.cat (cache_utils.py:102-120).
// Every kernel reads the current step from device memory so that one captured CUDA graph replays for all steps.
#include "common.cuh"
#include "spacer_b200.h"

namespace {

// x[r] = embed[token[r]]
__global__ void dec_embed_kernel(const int* __restrict__ tokens, const bf16* __restrict__ embed, bf16* __restrict__ x,
                                 int H) {
  pdl_launch_dependents();
  const bool tr_on = blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0;
  const int tr = tr_on ? sb_trace_begin(SB_TR_EMBED) : -1;
  pdl_wait();
  sb_trace_mark(tr, 1);
  const int r = blockIdx.x;
  const bf16* src = embed + (long long)tokens[r] * H;
  for (int i = threadIdx.x; i < H / 8; i += blockDim.x)
    *reinterpret_cast<uint4*>(x + (long long)r * H + i * 8) = *reinterpret_cast<const uint4*>(src + i * 8);
  sb_trace_mark(tr, 2);
}

// x[r] += bf16(sum_s parts[s][r][:]) (if parts);  xn[r] = w * bf16(x * rstd)
// one CTA of 512 threads per row, 4 consecutive elements per thread per pass (float4 partial loads).  512 threads x <= 64
// registers leave half of the SM's register file free: a CTA of 1024 threads took all of it, so the 12 SMs that ran this
// kernel could not host the next GEMV's CTA early (PDL) -- those CTAs missed the pre-wait weight prefetch, became ready
// 0.5-0.8 us after the others and were the last to finish (tools/decode_lab.py --exp skew).
constexpr int RN_THREADS = 512;
__global__ void __launch_bounds__(RN_THREADS, 2)
dec_residual_rmsnorm_kernel(bf16* __restrict__ x, const float* __restrict__ parts, int S, long long part_stride_s,
                            long long part_stride_r, const bf16* __restrict__ w, bf16* __restrict__ xn, int H,
                            float eps) {
  pdl_launch_dependents();
  const bool tr_on = blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0;
  const int tr = tr_on ? sb_trace_begin(SB_TR_RMSNORM) : -1;
  constexpr int MAXV = 4;                      // up to 4 passes of 2048 elements (H <= 8192)
  constexpr int MAXS = 8;                      // split-K partials gathered in one round trip (more: extra rounds)
  // the norm weight does not depend on the preceding kernels: fetch it before the dependency wait
  uint2 wreg[MAXV];
#pragma unroll
  for (int np = 0; np < MAXV; ++np) {
    const int i = (np * RN_THREADS + threadIdx.x) * 4;
    wreg[np] = (xn && i < H) ? ldg_nc_u2_ordered(w + i) : make_uint2(0u, 0u);
    // ptxas still sinks the load itself below the wait (it only has to be complete where the value is used); a prefetch
    // has no result to wait for and stays here, so the load after the wait is an L2 hit instead of a cold HBM read
    if (xn && i < H && (threadIdx.x & 15) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(w + i) : "memory");
  }
  pdl_wait();
  sb_trace_mark(tr, 1);
  __shared__ float red[32];
  const int r = blockIdx.x;
  float v[MAXV][4];
  float ss = 0.f;
#pragma unroll
  for (int np = 0; np < MAXV; ++np) {
    const int i = (np * RN_THREADS + threadIdx.x) * 4;
    if (i >= H) break;
    const uint2 xu = *reinterpret_cast<const uint2*>(x + (long long)r * H + i);
    float c[4];
    if (parts) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int s0 = 0; s0 < S; s0 += MAXS) {
        float4 t[MAXS];
#pragma unroll
        for (int s = 0; s < MAXS; ++s)          // every load of the round is issued before the first add
          t[s] = (s0 + s < S) ? *reinterpret_cast<const float4*>(parts + (s0 + s) * part_stride_s + r * part_stride_r + i)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int s = 0; s < MAXS; ++s) { a.x += t[s].x; a.y += t[s].y; a.z += t[s].z; a.w += t[s].w; }
      }
      const float2 x01 = unpack_bf16(xu.x), x23 = unpack_bf16(xu.y);
      c[0] = bf16_round(bf16_round(a.x) + x01.x);
      c[1] = bf16_round(bf16_round(a.y) + x01.y);
      c[2] = bf16_round(bf16_round(a.z) + x23.x);
      c[3] = bf16_round(bf16_round(a.w) + x23.y);
      uint2 o;
      o.x = pack_bf16(c[0], c[1]);
      o.y = pack_bf16(c[2], c[3]);
      *reinterpret_cast<uint2*>(x + (long long)r * H + i) = o;
    } else {
      const float2 x01 = unpack_bf16(xu.x), x23 = unpack_bf16(xu.y);
      c[0] = x01.x; c[1] = x01.y; c[2] = x23.x; c[3] = x23.y;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { v[np][j] = c[j]; ss += c[j] * c[j]; }
  }
  ss = block_sum(ss, red);
  const float rstd = rsqrtf(ss / H + eps);
  if (xn) {
#pragma unroll
    for (int np = 0; np < MAXV; ++np) {
      const int i = (np * RN_THREADS + threadIdx.x) * 4;
      if (i >= H) break;
      const float2 w01 = unpack_bf16(wreg[np].x), w23 = unpack_bf16(wreg[np].y);
      uint2 o;
      o.x = pack_bf16(w01.x * bf16_round(v[np][0] * rstd), w01.y * bf16_round(v[np][1] * rstd));
      o.y = pack_bf16(w23.x * bf16_round(v[np][2] * rstd), w23.y * bf16_round(v[np][3] * rstd));
      *reinterpret_cast<uint2*>(xn + (long long)r * H + i) = o;
    }
  }
  sb_trace_mark(tr, 2);
}

// qkv = bf16(sum parts + bias); rope(q,k) at position rope_base + step; q -> q_out, k,v -> completion cache slot.
// grid (R, nh + 2*nkv): one CTA of hd/2 threads per (row, head); blockIdx.y >= nh + nkv are the v heads (no rotation)
__global__ void __launch_bounds__(64)
dec_qkv_post_kernel(const float* __restrict__ parts, int S, long long stride_s, long long stride_r,
                    const bf16* __restrict__ bias, const int* __restrict__ step_ptr, int rope_base, float theta,
                    int nh, int nkv, int hd, bf16* __restrict__ q_out, bf16* __restrict__ k_cache,
                    bf16* __restrict__ v_cache, long long cache_stride_r, int c_max) {
  pdl_launch_dependents();
  const bool tr_on = blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0;
  const int tr = tr_on ? sb_trace_begin(SB_TR_QKVPOST) : -1;
  const int r = blockIdx.x, head = blockIdx.y;
  const int half = hd / 2;
  const int i = threadIdx.x;
  const int c0 = head * hd + i, c1 = c0 + half;
  // constants (bias, rotary frequency) do not depend on the preceding kernels: computed before the dependency wait
  float bias0 = 0.f, bias1 = 0.f, inv_freq = 0.f;
  if (i < half) {
    bias0 = __bfloat162float(bias[c0]);
    bias1 = __bfloat162float(bias[c1]);
    inv_freq = 1.0f / powf(theta, (float)(2 * i) / (float)hd);
  }
  pdl_wait();
  sb_trace_mark(tr, 1);
  const int step = *step_ptr;
  const float* pr = parts + r * stride_r;
  const int slot = min(step, c_max - 1);
  if (i >= half) return;
  float a = 0.f, b = 0.f;
  constexpr int MAXS = 8;
  for (int s0 = 0; s0 < S; s0 += MAXS) {
    float ta[MAXS], tb[MAXS];
#pragma unroll
    for (int s = 0; s < MAXS; ++s) {            // every load of the round is issued before the first add
      ta[s] = (s0 + s < S) ? pr[(s0 + s) * stride_s + c0] : 0.f;
      tb[s] = (s0 + s < S) ? pr[(s0 + s) * stride_s + c1] : 0.f;
    }
#pragma unroll
    for (int s = 0; s < MAXS; ++s) { a += ta[s]; b += tb[s]; }
  }
  if (head >= nh + nkv) {   // v: bias only
    bf16* vd = v_cache + r * cache_stride_r + (long long)slot * nkv * hd + (head - nh - nkv) * hd;
    vd[i] = __float2bfloat16_rn(a + bias0);
    vd[i + half] = __float2bfloat16_rn(b + bias1);
    return;
  }
  a = bf16_round(a + bias0);
  b = bf16_round(b + bias1);
  const float pos = (float)(rope_base + step);
  float sn, cs;
  sincosf(pos * inv_freq, &sn, &cs);
  cs = bf16_round(cs); sn = bf16_round(sn);
  const float o0 = bf16_round(bf16_round(a * cs) + bf16_round(-b * sn));
  const float o1 = bf16_round(bf16_round(b * cs) + bf16_round(a * sn));
  if (head < nh) {
    q_out[(long long)r * nh * hd + c0] = __float2bfloat16_rn(o0);
    q_out[(long long)r * nh * hd + c1] = __float2bfloat16_rn(o1);
  } else {
    bf16* kd = k_cache + r * cache_stride_r + (long long)slot * nkv * hd + (head - nh) * hd;
    kd[i] = __float2bfloat16_rn(o0);
    kd[i + half] = __float2bfloat16_rn(o1);
  }
  sb_trace_mark(tr, 2);
}

// act[r][c] = bf16( bf16(silu(g)) * u ),  g/u = bf16(sum of split-K partials) in the interleaved layout
__global__ void dec_swiglu_kernel(const float* __restrict__ parts, int S, long long stride_s, long long stride_r,
                                  bf16* __restrict__ act, int I) {
  pdl_launch_dependents();
  const bool tr_on = blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0;
  const int tr = tr_on ? sb_trace_begin(SB_TR_SWIGLU) : -1;
  pdl_wait();
  sb_trace_mark(tr, 1);
  const int r = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= I) return;
  const int gcol = (c / 64) * 128 + (c % 64);
  float g = 0.f, u = 0.f;
  for (int s = 0; s < S; ++s) {
    g += parts[s * stride_s + r * stride_r + gcol];
    u += parts[s * stride_s + r * stride_r + gcol + 64];
  }
  g = bf16_round(g);
  u = bf16_round(u);
  const float si = bf16_round(g / (1.f + __expf(-g)));
  act[(long long)r * I + c] = __float2bfloat16_rn(si * u);
  sb_trace_mark(tr, 2);
}

// L2 prefetch of weights a later GEMV will stream (see sb_dec_l2_prefetch in the header): one warp per CTA, 8 KB per
// instruction, evict_last so that the weight streams passing through L2 in between (tagged evict_first) do not push
// the lines out before they are read.  tools/labs/l2_hint_lab.cu measures what that buys.
__global__ void __launch_bounds__(32)
dec_l2_prefetch_kernel(const uint8_t* __restrict__ base, long long chunk_bytes, long long stride_bytes, int chunks) {
  constexpr long long PIECE = 8192;
  const uint64_t pol = l2_policy_evict_last();
  const long long per = (chunk_bytes + PIECE - 1) / PIECE, n = per * chunks;
  for (long long c = (long long)blockIdx.x * 32 + threadIdx.x; c < n; c += (long long)gridDim.x * 32) {
    const long long off = (c % per) * PIECE;
    const long long len = min(PIECE, chunk_bytes - off) & ~15LL;
    if (len > 0)
      asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(base + (c / per) * stride_bytes + off),
                   "r"((uint32_t)len), "l"(pol) : "memory");
  }
}

// the same request through the load/store path: one prefetch.global.L2::evict_last per 128-byte line, paced with
// nanosleep.  Unlike the bulk form it does not queue in the SM's TMA unit -- where a backlog of bulk prefetches holds up the
// weight-ring requests of the GEMV CTA that shares the SM -- and its rate can be set below what HBM drains, so that the
// latency-bound small kernels of the layer do not wait behind a saturated memory system.
__global__ void __launch_bounds__(128)
dec_l2_prefetch_lsu_kernel(const uint8_t* __restrict__ base, long long chunk_bytes, long long stride_bytes, int chunks,
                           int sleep_ns) {
  const long long per = chunk_bytes / 128, n = per * chunks;
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x) {
    const uint8_t* ptr = base + (c / per) * stride_bytes + (c % per) * 128;
    asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(ptr) : "memory");
    if (sleep_ns > 0) __nanosleep(sleep_ns);
  }
}

}  // namespace

SB_DEFINE_TRACE_SETTER(sb_trace_set_decode)

#define STREAM(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int sb_dec_embed(const int* tokens, const void* embed, void* x, int R, int H, sb_stream_t stream) {
  SB_REQUIRE(tokens && embed && x && R > 0 && H % 8 == 0, "sb_dec_embed: bad arguments");
  // The first kernel of a decode step is launched WITHOUT programmatic dependent launch: a full dependency on whatever
  // precedes it (the previous step's sampler when steps are enqueued back to back outside a graph) ends the PDL cascade at
  // the step boundary, which is what lets later kernels read the step counter and older cache rows before their own wait.
  SB_CUDA(sb_launch(dec_embed_kernel, dim3(R), dim3(256), 0, STREAM(stream), false, tokens, (const bf16*)embed,
                    (bf16*)x, H));
  return sb_check_launch("sb_dec_embed");
}


extern "C" int sb_dec_residual_rmsnorm(void* x, const float* parts, int S, long long stride_s, long long stride_r,
                                       const void* w, void* xn, int R, int H, float eps, sb_stream_t stream) {
  SB_REQUIRE(x && R > 0 && H > 0 && H % 4 == 0 && H <= 8192 && (xn == nullptr || w != nullptr),
             "sb_dec_residual_rmsnorm: bad arguments (H must be a multiple of 4, <= 8192)");
  SB_REQUIRE(parts == nullptr || (stride_s % 4 == 0 && stride_r % 4 == 0), "sb_dec_residual_rmsnorm: partial strides must be multiples of 4");
  SB_CUDA(sb_launch(dec_residual_rmsnorm_kernel, dim3(R), dim3(RN_THREADS), 0, STREAM(stream), sb_pdl_enabled(), (bf16*)x,
                    parts, S, stride_s, stride_r, (const bf16*)w, (bf16*)xn, H, eps));
  return sb_check_launch("sb_dec_residual_rmsnorm");
}

extern "C" int sb_dec_qkv_post(const float* parts, int S, long long stride_s, long long stride_r, const void* bias,
                               const int* step_ptr, int rope_base, float theta, int n_heads, int n_kv_heads,
                               int head_dim, void* q_out, void* k_cache, void* v_cache, long long cache_stride_r,
                               int c_max, int R, sb_stream_t stream) {
  SB_REQUIRE(parts && bias && step_ptr && q_out && k_cache && v_cache && R > 0 && S > 0, "sb_dec_qkv_post: bad arguments");
  SB_REQUIRE(head_dim == 128, "sb_dec_qkv_post: head_dim must be 128, got %d", head_dim);
  SB_CUDA(sb_launch(dec_qkv_post_kernel, dim3(R, n_heads + 2 * n_kv_heads), dim3(64), 0, STREAM(stream), sb_pdl_enabled(),
                    parts, S, stride_s, stride_r, (const bf16*)bias, step_ptr, rope_base, theta, n_heads, n_kv_heads,
                    head_dim, (bf16*)q_out, (bf16*)k_cache, (bf16*)v_cache, cache_stride_r, c_max));
  return sb_check_launch("sb_dec_qkv_post");
}

extern "C" int sb_dec_swiglu(const float* parts, int S, long long stride_s, long long stride_r, void* act, int R, int I,
                             sb_stream_t stream) {
  SB_REQUIRE(parts && act && R > 0 && I % 64 == 0, "sb_dec_swiglu: bad arguments");
  SB_CUDA(sb_launch(dec_swiglu_kernel, dim3((I + 255) / 256, R), dim3(256), 0, STREAM(stream), sb_pdl_enabled(), parts, S,
                    stride_s, stride_r, (bf16*)act, I));
  return sb_check_launch("sb_dec_swiglu");
}

extern "C" int sb_dec_l2_prefetch(const void* base, long long chunk_bytes, long long stride_bytes, int chunks, int ctas,
                                  int pace_ns, sb_stream_t stream) {
  SB_REQUIRE(base && (reinterpret_cast<uintptr_t>(base) & 15) == 0 && chunk_bytes > 0 && chunks >= 1 &&
                 (chunks == 1 || (stride_bytes >= chunk_bytes && stride_bytes % 16 == 0)),
             "sb_dec_l2_prefetch: bad arguments (16-byte aligned base, chunk_bytes > 0, stride >= chunk)");
  int dev = 0, sms = 0;
  SB_CUDA(cudaGetDevice(&dev));
  SB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (sms <= 0) sms = 148;
  if (pace_ns >= 0) {
    SB_REQUIRE(chunk_bytes % 128 == 0, "sb_dec_l2_prefetch: chunk_bytes must be a multiple of 128 for the paced form");
    SB_CUDA(sb_launch(dec_l2_prefetch_lsu_kernel, dim3(ctas > 0 ? ctas : sms), dim3(128), 0, STREAM(stream), false,
                      (const uint8_t*)base, chunk_bytes, stride_bytes, chunks, pace_ns));
    return sb_check_launch("sb_dec_l2_prefetch");
  }
  SB_CUDA(sb_launch(dec_l2_prefetch_kernel, dim3(ctas > 0 && ctas < sms ? ctas : sms), dim3(32), 0, STREAM(stream), false,
                    (const uint8_t*)base, chunk_bytes, stride_bytes, chunks));
  return sb_check_launch("sb_dec_l2_prefetch");
}
