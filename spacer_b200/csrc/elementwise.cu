// HBM-bound kernels of the SG-RLVR hot path (everything that is not a dense contraction):
// LayerNorm / RMSNorm fwd+bwd, 2-D ViT rotary and M-RoPE (+ KV-cache write), activation backward,
// embedding gather/scatter, row gather/scatter, column sums for bias gradients, fp32->bf16 cast.
// All are one pass over their operands with 128-bit accesses; fp32 math, bf16 storage.
#include "common.cuh"
#include "spacer_b200.h"

namespace {

SB_DEVICE void ld8f(const bf16* p, float* v) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
SB_DEVICE void st8f(bf16* p, const float* v) {
  uint4 u;
  u.x = pack_bf16(v[0], v[1]); u.y = pack_bf16(v[2], v[3]);
  u.z = pack_bf16(v[4], v[5]); u.w = pack_bf16(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

// ------------------------------------------------------------------------------------------
// cast
// ------------------------------------------------------------------------------------------
__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long n) {
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  const long long stride = (long long)gridDim.x * blockDim.x * 8;
  for (; i + 7 < n; i += stride) {
    float4 a = *reinterpret_cast<const float4*>(src + i);
    float4 b = *reinterpret_cast<const float4*>(src + i + 4);
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    st8f(dst + i, v);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (long long j = n & ~7LL; j < n; ++j) dst[j] = __float2bfloat16_rn(src[j]);
  }
}

// ------------------------------------------------------------------------------------------
// LayerNorm (ViT, eps 1e-6)        MQ2:464-465,479-486,317  (torch.nn.LayerNorm semantics)
// one CTA per row; E <= 8*blockDim*MAXV
// ------------------------------------------------------------------------------------------
constexpr int NORM_THREADS = 256;

template <bool RMS>
__global__ void __launch_bounds__(NORM_THREADS)
norm_fwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w, const bf16* __restrict__ b,
                bf16* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                int T, int E, float eps) {
  __shared__ float red[32];
  const int row = blockIdx.x;
  const bf16* xr = x + (long long)row * E;
  bf16* yr = y + (long long)row * E;
  const int nv = E / 8;
  float s = 0.f, ss = 0.f;
  for (int i = threadIdx.x; i < nv; i += NORM_THREADS) {
    float v[8];
    ld8f(xr + i * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s += v[j]; ss += v[j] * v[j]; }
  }
  float mean = 0.f, rstd;
  if (RMS) {
    ss = block_sum(ss, red);
    rstd = rsqrtf(ss / E + eps);
  } else {
    s = block_sum(s, red);
    mean = s / E;
    // second pass for the variance (row is L1/L2 resident): matches torch's two-pass numerics
    float vs = 0.f;
    for (int i = threadIdx.x; i < nv; i += NORM_THREADS) {
      float v[8];
      ld8f(xr + i * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = v[j] - mean; vs += d * d; }
    }
    vs = block_sum(vs, red);
    rstd = rsqrtf(vs / E + eps);
  }
  if (threadIdx.x == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
  for (int i = threadIdx.x; i < nv; i += NORM_THREADS) {
    float v[8], g[8], o[8];
    ld8f(xr + i * 8, v);
    ld8f(w + i * 8, g);
    if (RMS) {
      // HF: weight * (x * rstd).to(bf16)                                   MQ2:126-131
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = g[j] * bf16_round(v[j] * rstd);
    } else {
      float bb[8];
      ld8f(b + i * 8, bb);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (v[j] - mean) * rstd * g[j] + bb[j];
    }
    st8f(yr + i * 8, o);
  }
}

// backward: dx = [dres +] rstd * (g - mean(g) - xhat * mean(g*xhat))   (LayerNorm; g = dy*w)
//           dx = [dres +] rstd * (g - xhat * mean(g*xhat))             (RMSNorm)
// dw/db: every CTA writes the partial sums of its row group to a scratch row; reduce_partials_kernel then adds them to
// dw/db in CTA order (deterministic; the remaining atomics of the backward are the bf16 adds of duplicate rows in
// embed_bwd / scatter_add_rows).
template <bool RMS>
__global__ void __launch_bounds__(NORM_THREADS)
norm_bwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w, const float* __restrict__ mean_in,
                const float* __restrict__ rstd_in, const bf16* __restrict__ dy, const bf16* __restrict__ dres,
                bf16* __restrict__ dx, float* __restrict__ part_w, float* __restrict__ part_b, int T, int E,
                int rows_per_cta) {
  extern __shared__ float sm[];
  float* red = sm;            // 32
  float* dw_acc = sm + 32;    // E
  float* db_acc = dw_acc + E; // E (LayerNorm only)
  for (int i = threadIdx.x; i < E; i += NORM_THREADS) {
    dw_acc[i] = 0.f;
    if (!RMS) db_acc[i] = 0.f;
  }
  __syncthreads();
  const int nv = E / 8;
  const int r0 = blockIdx.x * rows_per_cta;
  const int r1 = min(r0 + rows_per_cta, T);
  for (int row = r0; row < r1; ++row) {
    const bf16* xr = x + (long long)row * E;
    const bf16* dyr = dy + (long long)row * E;
    const float mean = RMS ? 0.f : mean_in[row];
    const float rstd = rstd_in[row];
    float sg = 0.f, sgx = 0.f;
    for (int i = threadIdx.x; i < nv; i += NORM_THREADS) {
      float v[8], g[8], d[8];
      ld8f(xr + i * 8, v);
      ld8f(w + i * 8, g);
      ld8f(dyr + i * 8, d);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (v[j] - mean) * rstd;
        const float gg = d[j] * g[j];
        sg += gg;
        sgx += gg * xh;
        dw_acc[i * 8 + j] += d[j] * xh;   // each column is owned by one thread: no race
        if (!RMS) db_acc[i * 8 + j] += d[j];
      }
    }
    sgx = block_sum(sgx, red);
    if (!RMS) sg = block_sum(sg, red);
    const float m1 = RMS ? 0.f : sg / E;
    const float m2 = sgx / E;
    for (int i = threadIdx.x; i < nv; i += NORM_THREADS) {
      float v[8], g[8], d[8], o[8];
      ld8f(xr + i * 8, v);
      ld8f(w + i * 8, g);
      ld8f(dyr + i * 8, d);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (v[j] - mean) * rstd;
        o[j] = rstd * (d[j] * g[j] - m1 - xh * m2);
      }
      if (dres) {
        float r[8];
        ld8f(dres + (long long)row * E + i * 8, r);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += r[j];
      }
      st8f(dx + (long long)row * E + i * 8, o);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < E; i += NORM_THREADS) {
    part_w[(long long)blockIdx.x * E + i] = dw_acc[i];
    if (!RMS) part_b[(long long)blockIdx.x * E + i] = db_acc[i];
  }
}

// out[i] += sum_p part[p * stride + i], p ascending: the deterministic second stage of the column reductions
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const float* __restrict__ part, int n_parts, long long stride, float* __restrict__ out, int N) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= N) return;
  float s = 0.f;
  for (int p = 0; p < n_parts; ++p) s += part[p * stride + i];
  out[i] += s;
}

// per-device scratch for the partial rows (norm backward: up to 296 CTAs x 2 x 12000 floats; colsum: 1 M floats)
constexpr size_t COL_SCRATCH_FLOATS = (size_t)296 * 2 * 12000 + (1u << 20);
float* g_col_scratch[64] = {nullptr};
static float* col_scratch() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (g_col_scratch[dev] == nullptr && cudaMalloc(&g_col_scratch[dev], COL_SCRATCH_FLOATS * sizeof(float)) != cudaSuccess)
    g_col_scratch[dev] = nullptr;
  return g_col_scratch[dev];
}

// ------------------------------------------------------------------------------------------
// ViT 2-D rotary, in place on the q and k thirds of qkv [T, 3*heads*hd]   MQ2:257-268,725-752
// fp32 math, one rounding to bf16.  inverse=1 rotates by -angle (backward).
// token n of a (t,h,w) grid sits at frame n/(h*w), merge-block-major inside the frame.
// ------------------------------------------------------------------------------------------
// ROPE_TOK tokens per CTA iteration: their hd/2 (cos, sin) pairs are computed once into shared memory (all threads busy:
// powf / sincosf are the long-latency part), then every (token, q|k, head, 8-element chunk) is rotated with 16-byte loads
// and stores, several independent items per thread.  (One token per CTA left this kernel 8-12x off the HBM roofline:
// 160 us for 84 MB at cfg3, profiles/r01_launches_train_c3_tcbwd.txt.)
constexpr int ROPE_TOK = 8;
__global__ void __launch_bounds__(256)
rope_vit_kernel(bf16* __restrict__ qkv, int T, int heads, int hd, const int* __restrict__ grids, int n_grids, int merge,
                int inverse, const int* __restrict__ pos_hw) {
  __shared__ float s_cs[ROPE_TOK][128], s_sn[ROPE_TOK][128];
  const int half = hd / 2;      // 40: rotation pairs (i, i+half)
  const int quarter = hd / 4;   // 20: first quarter of freqs follows h, second follows w
  const int chunks = half / 8;
  const int per_tok = 2 * heads * chunks;
  for (int n0 = blockIdx.x * ROPE_TOK; n0 < T; n0 += gridDim.x * ROPE_TOK) {
    const int nt = min(ROPE_TOK, T - n0);
    for (int e = threadIdx.x; e < nt * half; e += blockDim.x) {
      const int tk = e / half, i = e % half;
      const int n = n0 + tk;
      int hpos, wpos;
      if (pos_hw) {   // explicit (h, w) per token: Qwen2.5-VL's window-reordered sequence
        hpos = pos_hw[2 * n];
        wpos = pos_hw[2 * n + 1];
      } else {
        // locate the grid this token belongs to
        int base = 0, gh = 1, gw = 1;
        for (int g = 0; g < n_grids; ++g) {
          const int gt = grids[g * 3], h_ = grids[g * 3 + 1], w_ = grids[g * 3 + 2];
          const int cnt = gt * h_ * w_;
          if (n < base + cnt || g == n_grids - 1) { gh = h_; gw = w_; break; }
          base += cnt;
        }
        const int r = (n - base) % (gh * gw);
        const int blk = r / (merge * merge), inner = r % (merge * merge);
        const int bw_n = gw / merge;
        hpos = (blk / bw_n) * merge + inner / merge;
        wpos = (blk % bw_n) * merge + inner % merge;
      }
      const int fi = i < quarter ? i : i - quarter;
      const float inv_freq = 1.0f / powf(10000.0f, (float)(2 * fi) / (float)half);
      const float ang = (float)(i < quarter ? hpos : wpos) * inv_freq;
      float sn, cs;
      sincosf(ang, &sn, &cs);
      s_cs[tk][i] = cs;
      s_sn[tk][i] = inverse ? -sn : sn;
    }
    __syncthreads();
    for (int w = threadIdx.x; w < nt * per_tok; w += blockDim.x) {
      const int tk = w / per_tok, wi = w % per_tok;
      const int c = wi % chunks;
      bf16* p = qkv + (long long)(n0 + tk) * 3 * heads * hd + (wi / chunks) * hd + c * 8;   // q heads then k heads
      float a[8], b[8], oa[8], ob[8];
      ld8f(p, a);
      ld8f(p + half, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float cs = s_cs[tk][c * 8 + j], sn = s_sn[tk][c * 8 + j];
        oa[j] = a[j] * cs - b[j] * sn;
        ob[j] = b[j] * cs + a[j] * sn;
      }
      st8f(p, oa);
      st8f(p + half, ob);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// M-RoPE in place on q,k of qkv [T, (nh+2nkv)*hd] + optional KV-cache write   MQ2:188-254
// pos: int32 [3, T].  HF numerics: cos/sin fp32 -> bf16; bf16(q*cos) + bf16(rot*sin) -> bf16.
// ------------------------------------------------------------------------------------------
// ROPE_TOK tokens per CTA iteration (see rope_vit_kernel): (cos, sin) of their hd/2 frequencies once into shared memory,
// then 16-byte chunks per (token, head)
__global__ void __launch_bounds__(256)
mrope_kernel(bf16* __restrict__ qkv, const int* __restrict__ pos, int T, int nh, int nkv, int hd, float theta,
             int sec_t, int sec_h, int inverse, bf16* __restrict__ k_out, bf16* __restrict__ v_out, long long kv_ld) {
  __shared__ float s_cs[ROPE_TOK][128], s_sn[ROPE_TOK][128];
  const int half = hd / 2;
  const int chunks = half / 8;
  const int nrot = nh + nkv;
  const int per_tok = nrot * chunks;
  const int v_chunks = nkv * hd / 8;
  const int qkv_ld = (nh + 2 * nkv) * hd;
  for (int n0 = blockIdx.x * ROPE_TOK; n0 < T; n0 += gridDim.x * ROPE_TOK) {
    const int nt = min(ROPE_TOK, T - n0);
    for (int e = threadIdx.x; e < nt * half; e += blockDim.x) {
      const int tk = e / half, i = e % half;
      const int stream = i < sec_t ? 0 : (i < sec_t + sec_h ? 1 : 2);
      const float p_ = (float)pos[(long long)stream * T + n0 + tk];
      const float inv_freq = 1.0f / powf(theta, (float)(2 * i) / (float)hd);
      float sn, cs;
      sincosf(p_ * inv_freq, &sn, &cs);
      s_cs[tk][i] = bf16_round(cs);
      sn = bf16_round(sn);
      s_sn[tk][i] = inverse ? -sn : sn;
    }
    __syncthreads();
    for (int w = threadIdx.x; w < nt * per_tok; w += blockDim.x) {
      const int tk = w / per_tok, wi = w % per_tok;
      const int c = wi % chunks, head = wi / chunks;
      const int n = n0 + tk;
      bf16* p = qkv + (long long)n * qkv_ld + head * hd + c * 8;
      float a[8], b[8], oa[8], ob[8];
      ld8f(p, a);
      ld8f(p + half, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float cs = s_cs[tk][c * 8 + j], sn = s_sn[tk][c * 8 + j];
        oa[j] = bf16_round(bf16_round(a[j] * cs) + bf16_round(-b[j] * sn));
        ob[j] = bf16_round(bf16_round(b[j] * cs) + bf16_round(a[j] * sn));
      }
      st8f(p, oa);
      st8f(p + half, ob);
      if (k_out && head >= nh) {
        bf16* kd = k_out + (long long)n * kv_ld + (head - nh) * hd + c * 8;
        st8f(kd, oa);
        st8f(kd + half, ob);
      }
    }
    if (v_out) {
      for (int w = threadIdx.x; w < nt * v_chunks; w += blockDim.x) {
        const int n = n0 + w / v_chunks, c = w % v_chunks;
        *reinterpret_cast<uint4*>(v_out + (long long)n * kv_ld + c * 8) =
            *reinterpret_cast<const uint4*>(qkv + (long long)n * qkv_ld + (nh + nkv) * hd + c * 8);
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// activations (recompute / backward)
// ------------------------------------------------------------------------------------------
SB_DEVICE float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

// mode 0: quick_gelu, 1: gelu(erf).   f = act(z);  dz = dy * act'(z)
__global__ void act_fwd_kernel(const bf16* __restrict__ z, bf16* __restrict__ f, long long n, int mode) {
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < n;
       i += (long long)gridDim.x * blockDim.x * 8) {
    float v[8], o[8];
    ld8f(z + i, v);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      o[j] = mode == 0 ? v[j] * sigmoidf_(1.702f * v[j]) : 0.5f * v[j] * (1.f + erff(v[j] * 0.70710678118654752f));
    st8f(f + i, o);
  }
}
__global__ void act_bwd_kernel(const bf16* __restrict__ z, const bf16* __restrict__ dy, bf16* __restrict__ dz,
                               long long n, int mode) {
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < n;
       i += (long long)gridDim.x * blockDim.x * 8) {
    float v[8], d[8], o[8];
    ld8f(z + i, v);
    ld8f(dy + i, d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float g;
      if (mode == 0) {
        const float s = sigmoidf_(1.702f * v[j]);
        g = s * (1.f + 1.702f * v[j] * (1.f - s));
      } else {
        const float cdf = 0.5f * (1.f + erff(v[j] * 0.70710678118654752f));
        g = cdf + v[j] * 0.3989422804014327f * __expf(-0.5f * v[j] * v[j]);
      }
      o[j] = d[j] * g;
    }
    st8f(dz + i, o);
  }
}

// SwiGLU backward on the interleaved [64 gate | 64 up] layout.
// gu [T, 2I] raw, dact [T, I] -> dgu [T, 2I]; optionally recompute act [T, I] (for the down-proj dW).
__global__ void swiglu_bwd_kernel(const bf16* __restrict__ gu, const bf16* __restrict__ dact, bf16* __restrict__ dgu,
                                  bf16* __restrict__ act, int T, int I) {
  const int chunks = I / 8;
  const long long total = (long long)T * chunks;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = idx % chunks;
    const long long t = idx / chunks;
    const int col = c * 8;                              // column in [0, I)
    const int gcol = (col / 64) * 128 + (col % 64);     // gate column in the raw layout
    float g[8], u[8], d[8], dg[8], du[8], a[8];
    ld8f(gu + t * 2 * I + gcol, g);
    ld8f(gu + t * 2 * I + gcol + 64, u);
    if (dact) ld8f(dact + t * I + col, d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float s = sigmoidf_(g[j]);
      const float si = bf16_round(g[j] * s);
      a[j] = si * u[j];
      if (dact) {
        du[j] = d[j] * si;
        dg[j] = d[j] * u[j] * s * (1.f + g[j] * (1.f - s));
      }
    }
    if (dact) {
      st8f(dgu + t * 2 * I + gcol, dg);
      st8f(dgu + t * 2 * I + gcol + 64, du);
    }
    if (act) st8f(act + t * I + col, a);
  }
}

// ------------------------------------------------------------------------------------------
// embedding gather with vision-embedding scatter                         MQ2:1255-1272
// vis_idx[t] = running index among placeholder tokens (computed by vision_index_kernel)
// ------------------------------------------------------------------------------------------
__global__ void vision_index_kernel(const int* __restrict__ ids, int* __restrict__ vis_idx, int T, int video_id,
                                    int image_id, int* __restrict__ count_out) {
  // single CTA, sequential chunks of blockDim tokens with a block scan per chunk
  __shared__ int warp_tot[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int base = 0; base < T; base += blockDim.x) {
    const int t = base + threadIdx.x;
    const int flag = (t < T && (ids[t] == video_id || ids[t] == image_id)) ? 1 : 0;
    int v = flag;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += n;
    }
    if (lane == 31) warp_tot[wid] = v;
    __syncthreads();
    int off = carry;
    for (int k = 0; k < wid; ++k) off += warp_tot[k];
    if (t < T) vis_idx[t] = flag ? off + v - 1 : -1;
    __syncthreads();
    if (threadIdx.x == 0) {
      int s = 0;
      for (int k = 0; k < nw; ++k) s += warp_tot[k];
      carry += s;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0 && count_out) *count_out = carry;
}

__global__ void embed_merge_kernel(const int* __restrict__ ids, const int* __restrict__ vis_idx,
                                   const bf16* __restrict__ embed, const bf16* __restrict__ vision,
                                   bf16* __restrict__ out, int T, int H, int n_vision) {
  const int nv = H / 8;
  const long long total = (long long)T * nv;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = idx % nv;
    const int t = idx / nv;
    const int vi = vis_idx ? vis_idx[t] : -1;
    const bf16* src = (vi >= 0 && vi < n_vision) ? vision + (long long)vi * H : embed + (long long)ids[t] * H;
    *reinterpret_cast<uint4*>(out + (long long)t * H + c * 8) = *reinterpret_cast<const uint4*>(src + c * 8);
  }
}

// backward: text rows scatter-add into d_embed (bf16x2 atomics; ids repeat), vision rows store into d_vision
__global__ void embed_bwd_kernel(const int* __restrict__ ids, const int* __restrict__ vis_idx,
                                 const bf16* __restrict__ dx, bf16* __restrict__ d_embed, bf16* __restrict__ d_vision,
                                 int T, int H, int n_vision) {
  const int nv = H / 2;
  const long long total = (long long)T * nv;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = idx % nv;
    const int t = idx / nv;
    const bf162 v = *reinterpret_cast<const bf162*>(dx + (long long)t * H + c * 2);
    const int vi = vis_idx ? vis_idx[t] : -1;
    if (vi >= 0 && vi < n_vision) {
      if (d_vision) *reinterpret_cast<bf162*>(d_vision + (long long)vi * H + c * 2) = v;
    } else if (d_embed) {
      atomicAdd(reinterpret_cast<bf162*>(d_embed + (long long)ids[t] * H + c * 2), v);
    }
  }
}

// ------------------------------------------------------------------------------------------
// row gather / scatter-add  (lm_head input rows; the last prompt row feeds all G groups)
// ------------------------------------------------------------------------------------------
__global__ void gather_rows_kernel(const bf16* __restrict__ src, const int* __restrict__ rows, bf16* __restrict__ dst,
                                   int R, int H) {
  const int nv = H / 8;
  const long long total = (long long)R * nv;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = idx % nv;
    const int r = idx / nv;
    *reinterpret_cast<uint4*>(dst + (long long)r * H + c * 8) =
        *reinterpret_cast<const uint4*>(src + (long long)rows[r] * H + c * 8);
  }
}
__global__ void scatter_add_rows_kernel(const bf16* __restrict__ src, const int* __restrict__ rows,
                                        bf16* __restrict__ dst, int R, int H) {
  const int nv = H / 2;
  const long long total = (long long)R * nv;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = idx % nv;
    const int r = idx / nv;
    atomicAdd(reinterpret_cast<bf162*>(dst + (long long)rows[r] * H + c * 2),
              *reinterpret_cast<const bf162*>(src + (long long)r * H + c * 2));
  }
}

// Deterministic scatter-add: dst[seg_dst[s]] (+)= sum over k in [seg_off[s], seg_off[s+1]) of src[order[k]], summed in
// fp32 in list order and rounded once (the atomic versions above round after every add, in arrival order).  Used for
// the embedding gradient (segments = distinct token ids) and for d_hidden of the lm_head rows (the last prompt row
// feeds every completion's first token).  One thread per (segment, 8 columns).
__global__ void __launch_bounds__(256)
segment_sum_rows_kernel(const bf16* __restrict__ src, const int* __restrict__ order, const int* __restrict__ seg_off,
                        const int* __restrict__ seg_dst, int n_seg, bf16* __restrict__ dst, int H, int accumulate) {
  const int nv = H / 8;
  const long long total = (long long)n_seg * nv;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % nv);
    const int sgm = (int)(idx / nv);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    bf16* d = dst + (long long)seg_dst[sgm] * H + c * 8;
    if (accumulate) ld8f(d, acc);
    for (int k = seg_off[sgm]; k < seg_off[sgm + 1]; ++k) {
      float v[8];
      ld8f(src + (long long)order[k] * H + c * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
    st8f(d, acc);
  }
}

// ------------------------------------------------------------------------------------------
// column sums (bias gradients): out_f32[n] += sum_t dy[t][n]
// grid (ceil(N/64), row_groups); 256 threads = 8 column-octets x 32 row lanes
// ------------------------------------------------------------------------------------------
__global__ void colsum_kernel(const bf16* __restrict__ dy, float* __restrict__ part, int T, int N, long long ld,
                              int rows_per_cta) {
  __shared__ float acc[32][65];
  const int c8 = threadIdx.x & 7;        // which 8-column group of this 64-column slab
  const int rl = threadIdx.x >> 3;       // row lane 0..31
  const int col = blockIdx.x * 64 + c8 * 8;
  const int r0 = blockIdx.y * rows_per_cta;
  const int r1 = min(r0 + rows_per_cta, T);
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (col < N) {
    for (int r = r0 + rl; r < r1; r += 32) {
      float v[8];
      ld8f(dy + (long long)r * ld + col, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] += v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[rl][c8 * 8 + j] = s[j];
  __syncthreads();
  if (threadIdx.x < 64) {
    float t = 0.f;
    for (int r = 0; r < 32; ++r) t += acc[r][threadIdx.x];
    const int cc = blockIdx.x * 64 + threadIdx.x;
    if (cc < N) part[(long long)blockIdx.y * N + cc] = t;    // row-group partial; reduce_partials_kernel sums them in order
  }
}

int grid_for(long long work_items, int threads) {
  long long b = (work_items + threads - 1) / threads;
  const long long cap = 148LL * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

#define STREAM(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int sb_cast_f32_bf16(const float* src, void* dst, long long n, sb_stream_t stream) {
  SB_REQUIRE(src && dst && n >= 0, "sb_cast_f32_bf16: bad arguments");
  if (n == 0) return 0;
  cast_f32_bf16_kernel<<<grid_for(n / 8 + 1, 256), 256, 0, STREAM(stream)>>>(src, (bf16*)dst, n);
  return sb_check_launch("sb_cast_f32_bf16");
}

extern "C" int sb_layernorm_fwd(const void* x, const void* w, const void* b, void* y, float* mean, float* rstd,
                                int T, int E, float eps, sb_stream_t stream) {
  SB_REQUIRE(x && w && b && y && T > 0 && E > 0 && E % 8 == 0, "sb_layernorm_fwd: bad arguments (E %% 8 == 0)");
  norm_fwd_kernel<false><<<T, NORM_THREADS, 0, STREAM(stream)>>>((const bf16*)x, (const bf16*)w, (const bf16*)b,
                                                                 (bf16*)y, mean, rstd, T, E, eps);
  return sb_check_launch("sb_layernorm_fwd");
}

extern "C" int sb_rmsnorm_fwd(const void* x, const void* w, void* y, float* rstd, int T, int H, float eps,
                              sb_stream_t stream) {
  SB_REQUIRE(x && w && y && T > 0 && H > 0 && H % 8 == 0, "sb_rmsnorm_fwd: bad arguments (H %% 8 == 0)");
  norm_fwd_kernel<true><<<T, NORM_THREADS, 0, STREAM(stream)>>>((const bf16*)x, (const bf16*)w, nullptr, (bf16*)y,
                                                                nullptr, rstd, T, H, eps);
  return sb_check_launch("sb_rmsnorm_fwd");
}

static int norm_bwd_launch(bool rms, const void* x, const void* w, const float* mean, const float* rstd,
                           const void* dy, const void* dres, void* dx, float* dw, float* db, int T, int E,
                           cudaStream_t st) {
  const int ctas = T < 148 * 2 ? T : 148 * 2;
  const int rows_per = (T + ctas - 1) / ctas;
  const int grid = (T + rows_per - 1) / rows_per;
  const size_t smem = (32 + 2 * (size_t)E) * sizeof(float);
  float* scratch = col_scratch();
  SB_REQUIRE(scratch != nullptr, "sb_norm_bwd: could not allocate the reduction scratch");
  SB_REQUIRE((size_t)grid * 2 * E <= COL_SCRATCH_FLOATS, "sb_norm_bwd: reduction scratch too small for %d x %d", grid, E);
  float* part_w = scratch;
  float* part_b = scratch + (size_t)grid * E;
  if (rms) {
    static bool done = false;
    if (!done) { SB_CUDA(cudaFuncSetAttribute(norm_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); done = true; }
    norm_bwd_kernel<true><<<grid, NORM_THREADS, smem, st>>>((const bf16*)x, (const bf16*)w, nullptr, rstd,
                                                            (const bf16*)dy, (const bf16*)dres, (bf16*)dx, part_w, nullptr,
                                                            T, E, rows_per);
  } else {
    static bool done = false;
    if (!done) { SB_CUDA(cudaFuncSetAttribute(norm_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); done = true; }
    norm_bwd_kernel<false><<<grid, NORM_THREADS, smem, st>>>((const bf16*)x, (const bf16*)w, mean, rstd,
                                                             (const bf16*)dy, (const bf16*)dres, (bf16*)dx, part_w, part_b,
                                                             T, E, rows_per);
  }
  if (sb_check_launch("sb_norm_bwd")) return 1;
  reduce_partials_kernel<<<(E + 255) / 256, 256, 0, st>>>(part_w, grid, E, dw, E);
  if (!rms) reduce_partials_kernel<<<(E + 255) / 256, 256, 0, st>>>(part_b, grid, E, db, E);
  return sb_check_launch("sb_norm_bwd(reduce)");
}

extern "C" int sb_layernorm_bwd(const void* x, const void* w, const float* mean, const float* rstd, const void* dy,
                                const void* dres, void* dx, float* dw, float* db, int T, int E, sb_stream_t stream) {
  SB_REQUIRE(x && w && mean && rstd && dy && dx && dw && db && T > 0 && E % 8 == 0 && E <= 12000,
             "sb_layernorm_bwd: bad arguments");
  return norm_bwd_launch(false, x, w, mean, rstd, dy, dres, dx, dw, db, T, E, STREAM(stream));
}

extern "C" int sb_rmsnorm_bwd(const void* x, const void* w, const float* rstd, const void* dy, const void* dres,
                              void* dx, float* dw, int T, int H, sb_stream_t stream) {
  SB_REQUIRE(x && w && rstd && dy && dx && dw && T > 0 && H % 8 == 0 && H <= 12000, "sb_rmsnorm_bwd: bad arguments");
  return norm_bwd_launch(true, x, w, nullptr, rstd, dy, dres, dx, dw, nullptr, T, H, STREAM(stream));
}

extern "C" int sb_rope_vit(void* qkv, int T, int heads, int head_dim, const int* grids_dev, int n_grids, int merge,
                           int inverse, sb_stream_t stream) {
  SB_REQUIRE(qkv && grids_dev && T > 0 && heads > 0 && head_dim % 16 == 0 && head_dim <= 256 && n_grids > 0 && merge > 0,
             "sb_rope_vit: bad arguments (head_dim must be a multiple of 16, <= 256)");
  rope_vit_kernel<<<(T + ROPE_TOK - 1) / ROPE_TOK, 256, 0, STREAM(stream)>>>((bf16*)qkv, T, heads, head_dim, grids_dev, n_grids,
                                                                             merge, inverse, nullptr);
  return sb_check_launch("sb_rope_vit");
}

extern "C" int sb_rope_vit_pos(void* qkv, int T, int heads, int head_dim, const int* pos_hw, int inverse,
                               sb_stream_t stream) {
  SB_REQUIRE(qkv && pos_hw && T > 0 && heads > 0 && head_dim % 16 == 0 && head_dim <= 256,
             "sb_rope_vit_pos: bad arguments (head_dim must be a multiple of 16, <= 256)");
  rope_vit_kernel<<<(T + ROPE_TOK - 1) / ROPE_TOK, 256, 0, STREAM(stream)>>>((bf16*)qkv, T, heads, head_dim, nullptr, 0, 1, inverse,
                                                                             pos_hw);
  return sb_check_launch("sb_rope_vit_pos");
}

extern "C" int sb_mrope(void* qkv, const int* pos, int T, int n_heads, int n_kv_heads, int head_dim, float theta,
                        int sec_t, int sec_h, int inverse, void* k_out, void* v_out, long long kv_ld,
                        sb_stream_t stream) {
  SB_REQUIRE(qkv && pos && T > 0 && head_dim % 16 == 0 && head_dim <= 256 && sec_t >= 0 && sec_h >= 0 &&
                 sec_t + sec_h <= head_dim / 2, "sb_mrope: bad arguments (head_dim must be a multiple of 16, <= 256)");
  SB_REQUIRE(k_out == nullptr || kv_ld % 8 == 0, "sb_mrope: kv_ld must be a multiple of 8");
  mrope_kernel<<<(T + ROPE_TOK - 1) / ROPE_TOK, 256, 0, STREAM(stream)>>>((bf16*)qkv, pos, T, n_heads, n_kv_heads, head_dim, theta,
                                                                          sec_t, sec_h, inverse, (bf16*)k_out, (bf16*)v_out, kv_ld);
  return sb_check_launch("sb_mrope");
}

extern "C" int sb_act_fwd(const void* z, void* f, long long n, int mode, sb_stream_t stream) {
  SB_REQUIRE(z && f && n > 0 && n % 8 == 0 && (mode == 0 || mode == 1), "sb_act_fwd: bad arguments");
  act_fwd_kernel<<<grid_for(n / 8, 256), 256, 0, STREAM(stream)>>>((const bf16*)z, (bf16*)f, n, mode);
  return sb_check_launch("sb_act_fwd");
}
extern "C" int sb_act_bwd(const void* z, const void* dy, void* dz, long long n, int mode, sb_stream_t stream) {
  SB_REQUIRE(z && dy && dz && n > 0 && n % 8 == 0 && (mode == 0 || mode == 1), "sb_act_bwd: bad arguments");
  act_bwd_kernel<<<grid_for(n / 8, 256), 256, 0, STREAM(stream)>>>((const bf16*)z, (const bf16*)dy, (bf16*)dz, n, mode);
  return sb_check_launch("sb_act_bwd");
}
extern "C" int sb_swiglu_bwd(const void* gu, const void* dact, void* dgu, void* act, int T, int I,
                             sb_stream_t stream) {
  SB_REQUIRE(gu && T > 0 && I % 64 == 0 && (act || (dact && dgu)) && (!dact || dgu), "sb_swiglu_bwd: bad arguments");
  swiglu_bwd_kernel<<<grid_for((long long)T * I / 8, 256), 256, 0, STREAM(stream)>>>(
      (const bf16*)gu, (const bf16*)dact, (bf16*)dgu, (bf16*)act, T, I);
  return sb_check_launch("sb_swiglu_bwd");
}

extern "C" int sb_vision_index(const int* ids, int* vis_idx, int T, int video_id, int image_id, int* count_out,
                               sb_stream_t stream) {
  SB_REQUIRE(ids && vis_idx && T > 0, "sb_vision_index: bad arguments");
  vision_index_kernel<<<1, 1024, 0, STREAM(stream)>>>(ids, vis_idx, T, video_id, image_id, count_out);
  return sb_check_launch("sb_vision_index");
}
extern "C" int sb_embed_merge(const int* ids, const int* vis_idx, const void* embed, const void* vision, void* out,
                              int T, int H, int n_vision, sb_stream_t stream) {
  SB_REQUIRE(ids && embed && out && T > 0 && H % 8 == 0, "sb_embed_merge: bad arguments");
  embed_merge_kernel<<<grid_for((long long)T * H / 8, 256), 256, 0, STREAM(stream)>>>(
      ids, vision ? vis_idx : nullptr, (const bf16*)embed, (const bf16*)vision, (bf16*)out, T, H, n_vision);
  return sb_check_launch("sb_embed_merge");
}
extern "C" int sb_embed_bwd(const int* ids, const int* vis_idx, const void* dx, void* d_embed, void* d_vision, int T,
                            int H, int n_vision, sb_stream_t stream) {
  SB_REQUIRE(ids && dx && T > 0 && H % 2 == 0, "sb_embed_bwd: bad arguments");
  embed_bwd_kernel<<<grid_for((long long)T * H / 2, 256), 256, 0, STREAM(stream)>>>(
      ids, vis_idx, (const bf16*)dx, (bf16*)d_embed, (bf16*)d_vision, T, H, n_vision);
  return sb_check_launch("sb_embed_bwd");
}
extern "C" int sb_gather_rows(const void* src, const int* rows, void* dst, int R, int H, sb_stream_t stream) {
  SB_REQUIRE(src && rows && dst && R > 0 && H % 8 == 0, "sb_gather_rows: bad arguments");
  gather_rows_kernel<<<grid_for((long long)R * H / 8, 256), 256, 0, STREAM(stream)>>>((const bf16*)src, rows,
                                                                                    (bf16*)dst, R, H);
  return sb_check_launch("sb_gather_rows");
}
extern "C" int sb_scatter_add_rows(const void* src, const int* rows, void* dst, int R, int H, sb_stream_t stream) {
  SB_REQUIRE(src && rows && dst && R > 0 && H % 2 == 0, "sb_scatter_add_rows: bad arguments");
  scatter_add_rows_kernel<<<grid_for((long long)R * H / 2, 256), 256, 0, STREAM(stream)>>>((const bf16*)src, rows,
                                                                                         (bf16*)dst, R, H);
  return sb_check_launch("sb_scatter_add_rows");
}
extern "C" int sb_segment_sum_rows(const void* src, const int* order, const int* seg_off, const int* seg_dst, int n_seg,
                                   void* dst, int H, int accumulate, sb_stream_t stream) {
  SB_REQUIRE(src && order && seg_off && seg_dst && dst && n_seg > 0 && H % 8 == 0, "sb_segment_sum_rows: bad arguments");
  segment_sum_rows_kernel<<<grid_for((long long)n_seg * H / 8, 256), 256, 0, STREAM(stream)>>>(
      (const bf16*)src, order, seg_off, seg_dst, n_seg, (bf16*)dst, H, accumulate);
  return sb_check_launch("sb_segment_sum_rows");
}
extern "C" int sb_colsum(const void* dy, float* out, int T, int N, long long ld, sb_stream_t stream) {
  SB_REQUIRE(dy && out && T > 0 && N > 0 && N % 8 == 0 && ld % 8 == 0, "sb_colsum: bad arguments");
  const int gx = (N + 63) / 64;
  int gy = (148 * 4 + gx - 1) / gx;
  if (gy > (T + 31) / 32) gy = (T + 31) / 32;
  if (gy < 1) gy = 1;
  const int rows_per = (T + gy - 1) / gy;
  gy = (T + rows_per - 1) / rows_per;
  float* scratch = col_scratch();
  SB_REQUIRE(scratch != nullptr, "sb_colsum: could not allocate the reduction scratch");
  SB_REQUIRE((size_t)gy * N <= COL_SCRATCH_FLOATS, "sb_colsum: reduction scratch too small for %d x %d", gy, N);
  colsum_kernel<<<dim3(gx, gy), 256, 0, STREAM(stream)>>>((const bf16*)dy, scratch, T, N, ld, rows_per);
  if (sb_check_launch("sb_colsum")) return 1;
  reduce_partials_kernel<<<(N + 255) / 256, 256, 0, STREAM(stream)>>>(scratch, gy, N, out, N);
  return sb_check_launch("sb_colsum(reduce)");
}
