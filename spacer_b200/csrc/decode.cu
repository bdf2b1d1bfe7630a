// Small kernels of the autoregressive decode step (q_len = 1, R <= 32 rows that share one or two prompts).
// The weight-streaming GEMVs are sb_gemm (swap-AB, split-K, fp32 partials [S][R][N]); the kernels here
// consume those partials: residual + RMSNorm, qkv bias + M-RoPE + KV append, split-KV attention over the
// SHARED prompt cache plus the per-row completion cache, SwiGLU.
// replaces the q_len=1 iterations of GenerationMixin._sample (generation/utils.py:2743-2806) through
// Qwen2VLDecoderLayer (MQ2:597-662) and DynamicCache.update's torch.cat (cache_utils.py:102-120).
// Every kernel reads the current step from device memory so that one captured CUDA graph replays for all steps.
#include "common.cuh"
#include "spacer_b200.h"

namespace {

// x[r] = embed[token[r]]
__global__ void dec_embed_kernel(const int* __restrict__ tokens, const bf16* __restrict__ embed, bf16* __restrict__ x,
                                 int H) {
  const int r = blockIdx.x;
  const bf16* src = embed + (long long)tokens[r] * H;
  for (int i = threadIdx.x; i < H / 8; i += blockDim.x)
    *reinterpret_cast<uint4*>(x + (long long)r * H + i * 8) = *reinterpret_cast<const uint4*>(src + i * 8);
}

// x[r] += bf16(sum_s parts[s][r][:]) (if parts);  xn[r] = w * bf16(x * rstd)
__global__ void __launch_bounds__(256)
dec_residual_rmsnorm_kernel(bf16* __restrict__ x, const float* __restrict__ parts, int S, long long part_stride_s,
                            long long part_stride_r, const bf16* __restrict__ w, bf16* __restrict__ xn, int H,
                            float eps) {
  extern __shared__ float row[];  // H floats
  __shared__ float red[32];
  const int r = blockIdx.x;
  float ss = 0.f;
  for (int i = threadIdx.x; i < H; i += 256) {
    float v = __bfloat162float(x[(long long)r * H + i]);
    if (parts) {
      float a = 0.f;
      for (int s = 0; s < S; ++s) a += parts[s * part_stride_s + r * part_stride_r + i];
      v = bf16_round(bf16_round(a) + v);
      x[(long long)r * H + i] = __float2bfloat16_rn(v);
    }
    row[i] = v;
    ss += v * v;
  }
  ss = block_sum(ss, red);
  const float rstd = rsqrtf(ss / H + eps);
  if (xn) {
    for (int i = threadIdx.x; i < H; i += 256)
      xn[(long long)r * H + i] = __float2bfloat16_rn(__bfloat162float(w[i]) * bf16_round(row[i] * rstd));
  }
}

// qkv = bf16(sum parts + bias); rope(q,k) at position rope_base + step; q -> q_out, k,v -> completion cache slot
__global__ void __launch_bounds__(256)
dec_qkv_post_kernel(const float* __restrict__ parts, int S, long long stride_s, long long stride_r,
                    const bf16* __restrict__ bias, const int* __restrict__ step_ptr, int rope_base, float theta,
                    int nh, int nkv, int hd, bf16* __restrict__ q_out, bf16* __restrict__ k_cache,
                    bf16* __restrict__ v_cache, long long cache_stride_r, int c_max) {
  const int r = blockIdx.x;
  const int step = *step_ptr;
  const int half = hd / 2;
  const float pos = (float)(rope_base + step);
  const int n_rot = (nh + nkv) * half;
  const float* pr = parts + r * stride_r;
  const int slot = min(step, c_max - 1);
  for (int idx = threadIdx.x; idx < n_rot; idx += 256) {
    const int head = idx / half, i = idx % half;
    const int c0 = head * hd + i, c1 = c0 + half;
    float a = 0.f, b = 0.f;
    for (int s = 0; s < S; ++s) { a += pr[s * stride_s + c0]; b += pr[s * stride_s + c1]; }
    a = bf16_round(a + __bfloat162float(bias[c0]));
    b = bf16_round(b + __bfloat162float(bias[c1]));
    const float inv_freq = 1.0f / powf(theta, (float)(2 * i) / (float)hd);
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
  }
  const int voff = (nh + nkv) * hd;
  for (int i = threadIdx.x; i < nkv * hd; i += 256) {
    float a = 0.f;
    for (int s = 0; s < S; ++s) a += pr[s * stride_s + voff + i];
    a += __bfloat162float(bias[voff + i]);
    v_cache[r * cache_stride_r + (long long)slot * nkv * hd + i] = __float2bfloat16_rn(a);
  }
}

// split-KV decode attention.  grid (R, nkv, n_split), 128 threads, head_dim 128, rep <= 8 q heads per kv head.
// logical key index j in [0, P + step]: j < P -> shared prompt cache of the row's group; else completion cache.
constexpr int DA_THREADS = 128;
constexpr int DA_CHUNK = 256;
constexpr int DA_MAXREP = 8;

__global__ void __launch_bounds__(DA_THREADS)
dec_attn_kernel(const bf16* __restrict__ q, const bf16* __restrict__ kp0, const bf16* __restrict__ vp0,
                const bf16* __restrict__ kp1, const bf16* __restrict__ vp1, int rows_group0, int P,
                const bf16* __restrict__ kc, const bf16* __restrict__ vc, long long cache_stride_r,
                const int* __restrict__ step_ptr, int nh, int nkv, float scale, float* __restrict__ o_part,
                float* __restrict__ ml_part) {
  constexpr int HD = 128;
  __shared__ float sq[DA_MAXREP][HD];
  __shared__ float sp[DA_MAXREP][DA_CHUNK];
  __shared__ float red[32];
  const int r = blockIdx.x, kvh = blockIdx.y, sp_i = blockIdx.z, n_split = gridDim.z;
  const int rep = nh / nkv;
  const int tid = threadIdx.x;
  const int n_ctx = P + *step_ptr + 1;
  const int per = (n_ctx + n_split - 1) / n_split;
  const int j_lo = sp_i * per, j_hi = min(j_lo + per, n_ctx);
  const bf16* kp = r < rows_group0 ? kp0 : kp1;
  const bf16* vp = r < rows_group0 ? vp0 : vp1;
  const long long kv_ld = (long long)nkv * HD;

  for (int i = tid; i < rep * HD; i += DA_THREADS)
    sq[i / HD][i % HD] = __bfloat162float(q[(long long)r * nh * HD + (kvh * rep + i / HD) * HD + i % HD]) * scale;
  __syncthreads();

  float m_run[DA_MAXREP], l_run[DA_MAXREP], o_run[DA_MAXREP];
#pragma unroll
  for (int h = 0; h < DA_MAXREP; ++h) { m_run[h] = -INFINITY; l_run[h] = 0.f; o_run[h] = 0.f; }

  for (int c0 = j_lo; c0 < j_hi; c0 += DA_CHUNK) {
    const int cn = min(DA_CHUNK, j_hi - c0);
    // phase 1: scores, one key per thread (2 rounds for 256-key chunks)
    for (int jj = tid; jj < cn; jj += DA_THREADS) {
      const int j = c0 + jj;
      const bf16* kr = j < P ? kp + (long long)j * kv_ld + kvh * HD
                             : kc + r * cache_stride_r + (long long)(j - P) * kv_ld + kvh * HD;
      float acc[DA_MAXREP];
#pragma unroll
      for (int h = 0; h < DA_MAXREP; ++h) acc[h] = 0.f;
#pragma unroll 4
      for (int d8 = 0; d8 < HD / 8; ++d8) {
        const uint4 u = *reinterpret_cast<const uint4*>(kr + d8 * 8);
        const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
        const float kv8[8] = {a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y};
#pragma unroll
        for (int h = 0; h < DA_MAXREP; ++h) {
          if (h < rep) {
            const float4 q0 = *reinterpret_cast<const float4*>(&sq[h][d8 * 8]);
            const float4 q1 = *reinterpret_cast<const float4*>(&sq[h][d8 * 8 + 4]);
            acc[h] += q0.x * kv8[0] + q0.y * kv8[1] + q0.z * kv8[2] + q0.w * kv8[3] + q1.x * kv8[4] +
                      q1.y * kv8[5] + q1.z * kv8[6] + q1.w * kv8[7];
          }
        }
      }
#pragma unroll
      for (int h = 0; h < DA_MAXREP; ++h)
        if (h < rep) sp[h][jj] = acc[h];
    }
    __syncthreads();
    // per-head chunk max / rescale / exp (all threads cooperate per head)
#pragma unroll
    for (int h = 0; h < DA_MAXREP; ++h) {
      if (h < rep) {
        float mx = -INFINITY;
        for (int jj = tid; jj < cn; jj += DA_THREADS) mx = fmaxf(mx, sp[h][jj]);
        mx = block_max(mx, red);
        const float nm = fmaxf(m_run[h], mx);
        const float corr = __expf(m_run[h] - nm);
        float ls = 0.f;
        for (int jj = tid; jj < cn; jj += DA_THREADS) {
          const float e = __expf(sp[h][jj] - nm);
          sp[h][jj] = e;
          ls += e;
        }
        ls = block_sum(ls, red);
        l_run[h] = l_run[h] * corr + ls;
        o_run[h] *= corr;
        m_run[h] = nm;
      }
    }
    __syncthreads();
    // phase 2: thread d accumulates O[h][d] over the chunk's keys
    for (int jj = 0; jj < cn; ++jj) {
      const int j = c0 + jj;
      const bf16* vr = j < P ? vp + (long long)j * kv_ld + kvh * HD
                             : vc + r * cache_stride_r + (long long)(j - P) * kv_ld + kvh * HD;
      const float vv = __bfloat162float(vr[tid]);
#pragma unroll
      for (int h = 0; h < DA_MAXREP; ++h)
        if (h < rep) o_run[h] += sp[h][jj] * vv;
    }
    __syncthreads();
  }
  const long long base = (((long long)r * nkv + kvh) * n_split + sp_i) * rep;
#pragma unroll
  for (int h = 0; h < DA_MAXREP; ++h) {
    if (h < rep) {
      o_part[(base + h) * HD + tid] = o_run[h];
      if (tid == 0) { ml_part[(base + h) * 2] = m_run[h]; ml_part[(base + h) * 2 + 1] = l_run[h]; }
    }
  }
}

// merge the splits: out[r][(kvh*rep+h)*128 + d] bf16
__global__ void dec_attn_combine_kernel(const float* __restrict__ o_part, const float* __restrict__ ml_part,
                                        int n_split, int nh, int nkv, bf16* __restrict__ out) {
  constexpr int HD = 128;
  const int r = blockIdx.x, kvh = blockIdx.y, d = threadIdx.x;
  const int rep = nh / nkv;
  for (int h = 0; h < rep; ++h) {
    float m = -INFINITY;
    for (int s = 0; s < n_split; ++s)
      m = fmaxf(m, ml_part[((((long long)r * nkv + kvh) * n_split + s) * rep + h) * 2]);
    float l = 0.f, o = 0.f;
    for (int s = 0; s < n_split; ++s) {
      const long long b = (((long long)r * nkv + kvh) * n_split + s) * rep + h;
      const float ms = ml_part[b * 2];
      if (ms == -INFINITY) continue;
      const float w = __expf(ms - m);
      l += ml_part[b * 2 + 1] * w;
      o += o_part[b * HD + d] * w;
    }
    out[(long long)r * nh * HD + (kvh * rep + h) * HD + d] = __float2bfloat16_rn(l > 0.f ? o / l : 0.f);
  }
}

// act[r][c] = bf16( bf16(silu(g)) * u ),  g/u = bf16(sum of split-K partials) in the interleaved layout
__global__ void dec_swiglu_kernel(const float* __restrict__ parts, int S, long long stride_s, long long stride_r,
                                  bf16* __restrict__ act, int I) {
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
}

}  // namespace

#define STREAM(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int sb_dec_embed(const int* tokens, const void* embed, void* x, int R, int H, sb_stream_t stream) {
  SB_REQUIRE(tokens && embed && x && R > 0 && H % 8 == 0, "sb_dec_embed: bad arguments");
  dec_embed_kernel<<<R, 256, 0, STREAM(stream)>>>(tokens, (const bf16*)embed, (bf16*)x, H);
  return sb_check_launch("sb_dec_embed");
}

extern "C" int sb_dec_residual_rmsnorm(void* x, const float* parts, int S, long long stride_s, long long stride_r,
                                       const void* w, void* xn, int R, int H, float eps, sb_stream_t stream) {
  SB_REQUIRE(x && R > 0 && H > 0 && H * 4 <= 64 * 1024 && (xn == nullptr || w != nullptr),
             "sb_dec_residual_rmsnorm: bad arguments");
  static bool done = false;
  if (!done) {
    SB_CUDA(cudaFuncSetAttribute(dec_residual_rmsnorm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    done = true;
  }
  dec_residual_rmsnorm_kernel<<<R, 256, H * sizeof(float), STREAM(stream)>>>((bf16*)x, parts, S, stride_s, stride_r,
                                                                             (const bf16*)w, (bf16*)xn, H, eps);
  return sb_check_launch("sb_dec_residual_rmsnorm");
}

extern "C" int sb_dec_qkv_post(const float* parts, int S, long long stride_s, long long stride_r, const void* bias,
                               const int* step_ptr, int rope_base, float theta, int n_heads, int n_kv_heads,
                               int head_dim, void* q_out, void* k_cache, void* v_cache, long long cache_stride_r,
                               int c_max, int R, sb_stream_t stream) {
  SB_REQUIRE(parts && bias && step_ptr && q_out && k_cache && v_cache && R > 0 && S > 0, "sb_dec_qkv_post: bad arguments");
  dec_qkv_post_kernel<<<R, 256, 0, STREAM(stream)>>>(parts, S, stride_s, stride_r, (const bf16*)bias, step_ptr,
                                                     rope_base, theta, n_heads, n_kv_heads, head_dim, (bf16*)q_out,
                                                     (bf16*)k_cache, (bf16*)v_cache, cache_stride_r, c_max);
  return sb_check_launch("sb_dec_qkv_post");
}

extern "C" int sb_dec_attn(const void* q, const void* kp0, const void* vp0, const void* kp1, const void* vp1,
                           int rows_group0, int P, const void* k_cache, const void* v_cache, long long cache_stride_r,
                           const int* step_ptr, int n_heads, int n_kv_heads, int head_dim, float scale, int n_split,
                           float* o_part, float* ml_part, void* out, int R, sb_stream_t stream) {
  SB_REQUIRE(q && kp0 && vp0 && k_cache && v_cache && step_ptr && o_part && ml_part && out, "sb_dec_attn: null pointer");
  SB_REQUIRE(head_dim == 128, "sb_dec_attn: head_dim must be 128, got %d", head_dim);
  SB_REQUIRE(n_heads % n_kv_heads == 0 && n_heads / n_kv_heads <= DA_MAXREP, "sb_dec_attn: at most %d q heads per kv head", DA_MAXREP);
  SB_REQUIRE(R > 0 && n_split > 0, "sb_dec_attn: bad sizes");
  if (!kp1) { kp1 = kp0; vp1 = vp0; }
  dec_attn_kernel<<<dim3(R, n_kv_heads, n_split), DA_THREADS, 0, STREAM(stream)>>>(
      (const bf16*)q, (const bf16*)kp0, (const bf16*)vp0, (const bf16*)kp1, (const bf16*)vp1, rows_group0, P,
      (const bf16*)k_cache, (const bf16*)v_cache, cache_stride_r, step_ptr, n_heads, n_kv_heads, scale, o_part, ml_part);
  if (sb_check_launch("sb_dec_attn")) return 1;
  dec_attn_combine_kernel<<<dim3(R, n_kv_heads), 128, 0, STREAM(stream)>>>(o_part, ml_part, n_split, n_heads,
                                                                          n_kv_heads, (bf16*)out);
  return sb_check_launch("sb_dec_attn(combine)");
}

extern "C" int sb_dec_swiglu(const float* parts, int S, long long stride_s, long long stride_r, void* act, int R, int I,
                             sb_stream_t stream) {
  SB_REQUIRE(parts && act && R > 0 && I % 64 == 0, "sb_dec_swiglu: bad arguments");
  dec_swiglu_kernel<<<dim3((I + 255) / 256, R), 256, 0, STREAM(stream)>>>(parts, S, stride_s, stride_r, (bf16*)act, I);
  return sb_check_launch("sb_dec_swiglu");
}
