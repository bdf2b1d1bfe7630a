// AdamW with global gradient-norm clipping over flat parameter arenas.
// replaces DeepSpeed's fused Adam + clip (run_SpaceR_SG_RLVR.sh:23-25,37: lr 1e-6 cosine, weight decay 0.01,
// max_grad_norm 5; zero3.json: bf16 params with fp32 master weights and fp32 moments).
//   sumsq pass  : total_sq += sum g^2                                   (one read of the gradients)
//   update pass : g' = g * min(1, max_norm / (sqrt(total_sq) + 1e-6)); m,v,master in fp32; bf16 param written
#include "common.cuh"
#include "spacer_b200.h"

namespace {

template <typename GT>
SB_DEVICE float ldg(const GT* p, long long i);
template <>
SB_DEVICE float ldg<float>(const float* p, long long i) { return p[i]; }
template <>
SB_DEVICE float ldg<bf16>(const bf16* p, long long i) { return __bfloat162float(p[i]); }

template <typename GT>
__global__ void __launch_bounds__(256) sumsq_kernel(const GT* __restrict__ g, long long n, float* __restrict__ partials) {
  __shared__ float red[32];
  float s = 0.f;
  if constexpr (sizeof(GT) == 2) {
    const long long nv = n / 8;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < nv; i += (long long)gridDim.x * 256) {
      const uint4 u = reinterpret_cast<const uint4*>(g)[i];
      const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
      s += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y + c.x * c.x + c.y * c.y + d.x * d.x + d.y * d.y;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
      for (long long i = nv * 8; i < n; ++i) { const float v = ldg<GT>(g, i); s += v * v; }
  } else {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
      const float v = ldg<GT>(g, i);
      s += v * v;
    }
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;     // no float atomics: the total must be bit-identical on every
}                                                     // data-parallel rank (it scales the update when clipping)

// total += sum of the per-block partials, in a fixed order (double accumulation)
__global__ void __launch_bounds__(256) sumsq_final_kernel(const float* __restrict__ partials, int n, float* __restrict__ out) {
  __shared__ double red[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += (double)partials[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = (float)((double)*out + red[0]);
}

constexpr int SUMSQ_GRID = 148 * 8;
float* g_sumsq_partials[64] = {nullptr};   // per device

struct AdamArgs {
  float lr, beta1, beta2, eps, wd, bc1, bc2, max_norm;
  float grad_scale;   // extra factor on gradients (e.g. 1/world_size after a sum all-reduce)
};

// 4 consecutive elements per thread: 16-byte loads/stores of master, m, v (fp32), 8-byte of grad/param (bf16)
template <typename T>
SB_DEVICE void ld4(const T* p, long long i, float* o);
template <>
SB_DEVICE void ld4<float>(const float* p, long long i, float* o) {
  const float4 v = *reinterpret_cast<const float4*>(p + i);
  o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
template <>
SB_DEVICE void ld4<bf16>(const bf16* p, long long i, float* o) {
  const uint2 u = *reinterpret_cast<const uint2*>(p + i);
  const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
  o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
}
SB_DEVICE void st4(float* p, long long i, const float* v) {
  *reinterpret_cast<float4*>(p + i) = make_float4(v[0], v[1], v[2], v[3]);
}
SB_DEVICE void st4(bf16* p, long long i, const float* v) {
  uint2 u;
  u.x = pack_bf16(v[0], v[1]);
  u.y = pack_bf16(v[2], v[3]);
  *reinterpret_cast<uint2*>(p + i) = u;
}

SB_DEVICE void adam_elem(float& w, float& mi, float& vi, float gr, const AdamArgs& a) {
  mi = a.beta1 * mi + (1.f - a.beta1) * gr;
  vi = a.beta2 * vi + (1.f - a.beta2) * gr * gr;
  // torch.optim.AdamW: decoupled decay, then the Adam step with bias correction
  w *= 1.f - a.lr * a.wd;
  const float denom = sqrtf(vi) / sqrtf(a.bc2) + a.eps;
  w -= (a.lr / a.bc1) * (mi / denom);
}

template <typename GT, typename MT>
__global__ void __launch_bounds__(256)
adamw_kernel(bf16* __restrict__ p, float* __restrict__ master, MT* __restrict__ m, MT* __restrict__ v,
             const GT* __restrict__ g, long long n, const float* __restrict__ total_sq, AdamArgs a) {
  float clip = 1.f;
  if (total_sq && a.max_norm > 0.f) {
    const float norm = sqrtf(*total_sq) * a.grad_scale;
    clip = fminf(1.f, a.max_norm / (norm + 1e-6f));
  }
  const float gs = clip * a.grad_scale;
  const long long n4 = n / 4;
  for (long long q = (long long)blockIdx.x * 256 + threadIdx.x; q < n4; q += (long long)gridDim.x * 256) {
    const long long i = q * 4;
    float gr[4], w[4], mi[4], vi[4];
    ld4<GT>(g, i, gr);
    ld4<float>(master, i, w);
    ld4<MT>(m, i, mi);
    ld4<MT>(v, i, vi);
#pragma unroll
    for (int j = 0; j < 4; ++j) adam_elem(w[j], mi[j], vi[j], gr[j] * gs, a);
    st4(master, i, w);
    st4(m, i, mi);
    st4(v, i, vi);
    st4(p, i, w);
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n - n4 * 4)) {   // tail (arenas are 128-element aligned: normally empty)
    const long long i = n4 * 4 + threadIdx.x;
    float w = master[i], mi, vi;
    if constexpr (sizeof(MT) == 2) { mi = __bfloat162float(m[i]); vi = __bfloat162float(v[i]); }
    else { mi = m[i]; vi = v[i]; }
    adam_elem(w, mi, vi, ldg<GT>(g, i) * gs, a);
    master[i] = w;
    if constexpr (sizeof(MT) == 2) { m[i] = __float2bfloat16_rn(mi); v[i] = __float2bfloat16_rn(vi); }
    else { m[i] = mi; v[i] = vi; }
    p[i] = __float2bfloat16_rn(w);
  }
}

__global__ void bf16_to_f32_kernel(const bf16* __restrict__ s, float* __restrict__ d, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    d[i] = __bfloat162float(s[i]);
}

}  // namespace

#define STREAM(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int sb_grad_sumsq(const void* g, long long n, int grad_is_f32, float* total_sq, sb_stream_t stream) {
  SB_REQUIRE(g && total_sq && n > 0, "sb_grad_sumsq: bad arguments");
  int dev = 0;
  SB_CUDA(cudaGetDevice(&dev));
  SB_REQUIRE(dev >= 0 && dev < 64, "sb_grad_sumsq: device ordinal %d out of range", dev);
  if (g_sumsq_partials[dev] == nullptr) SB_CUDA(cudaMalloc(&g_sumsq_partials[dev], SUMSQ_GRID * sizeof(float)));
  float* part = g_sumsq_partials[dev];
  if (grad_is_f32) sumsq_kernel<float><<<SUMSQ_GRID, 256, 0, STREAM(stream)>>>((const float*)g, n, part);
  else sumsq_kernel<bf16><<<SUMSQ_GRID, 256, 0, STREAM(stream)>>>((const bf16*)g, n, part);
  if (sb_check_launch("sb_grad_sumsq")) return 1;
  sumsq_final_kernel<<<1, 256, 0, STREAM(stream)>>>(part, SUMSQ_GRID, total_sq);
  return sb_check_launch("sb_grad_sumsq(final)");
}

extern "C" int sb_adamw_step(void* param_bf16, float* master, void* m, void* v, const void* grad, long long n,
                             int grad_is_f32, int moments_are_bf16, const float* total_sq, float lr, float beta1,
                             float beta2, float eps, float weight_decay, int step, float max_norm, float grad_scale,
                             sb_stream_t stream) {
  SB_REQUIRE(param_bf16 && master && m && v && grad && n > 0 && step >= 1, "sb_adamw_step: bad arguments");
  SB_REQUIRE(((reinterpret_cast<uintptr_t>(param_bf16) | reinterpret_cast<uintptr_t>(master) | reinterpret_cast<uintptr_t>(m) |
               reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(grad)) & 15) == 0,
             "sb_adamw_step: arenas must be 16-byte aligned");
  AdamArgs a;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.wd = weight_decay;
  a.bc1 = 1.f - powf(beta1, (float)step);
  a.bc2 = 1.f - powf(beta2, (float)step);
  a.max_norm = max_norm; a.grad_scale = grad_scale;
  const int grid = 148 * 8;
  cudaStream_t st = STREAM(stream);
  if (grad_is_f32) {
    if (moments_are_bf16) adamw_kernel<float, bf16><<<grid, 256, 0, st>>>((bf16*)param_bf16, master, (bf16*)m, (bf16*)v, (const float*)grad, n, total_sq, a);
    else adamw_kernel<float, float><<<grid, 256, 0, st>>>((bf16*)param_bf16, master, (float*)m, (float*)v, (const float*)grad, n, total_sq, a);
  } else {
    if (moments_are_bf16) adamw_kernel<bf16, bf16><<<grid, 256, 0, st>>>((bf16*)param_bf16, master, (bf16*)m, (bf16*)v, (const bf16*)grad, n, total_sq, a);
    else adamw_kernel<bf16, float><<<grid, 256, 0, st>>>((bf16*)param_bf16, master, (float*)m, (float*)v, (const bf16*)grad, n, total_sq, a);
  }
  return sb_check_launch("sb_adamw_step");
}

extern "C" int sb_bf16_to_f32(const void* src, float* dst, long long n, sb_stream_t stream) {
  SB_REQUIRE(src && dst && n > 0, "sb_bf16_to_f32: bad arguments");
  bf16_to_f32_kernel<<<148 * 8, 256, 0, STREAM(stream)>>>((const bf16*)src, dst, n);
  return sb_check_launch("sb_bf16_to_f32");
}
