// GPU video front-end: rescale + normalise + patchify (+ frame permutation, + bf16 cast) in one pass.
// replaces  Qwen2VLVideoProcessor._preprocess (transformers 5.5.0 models/qwen2_vl/video_processing_qwen2_vl.py:240-272:
//           fused rescale/normalise `(x - mean/rf) / (std/rf)` in fp32 via torchvision normalize, temporal padding by
//           repeating the last frame, view/permute (0,1,4,7,5,8,3,2,6,9), flatten),
//           the second processor pass over the frame-shuffled video of T-GRPO (SG_RLVR_trainer.py:442-458) -- here an
//           index permutation `perm` applied while reading -- and the `.to(dtype)` of PatchEmbed.forward (MQ2:306).
// HBM-bound: reads each source pixel once (1 B for uint8 frames, 4 B for float frames), writes 4 B (fp32, bit-exact
// with the HF CPU path) and/or 2 B (bf16, what the ViT consumes) per output element.
#include "common.cuh"
#include "spacer_b200.h"

namespace {

struct PatchifyParams {
  const void* frames;     // [F][C][H][W], uint8 or float32 (values 0..255)
  int is_u8;
  const int* perm;        // optional [F]: output frame f reads source frame perm[f]
  int F, C, H, W;
  int patch, t_patch, merge;
  int gt, gh, gw;         // grid (temporal padded)
  float mean[4], stdv[4]; // already divided by the rescale factor, fp32
  float* out_f32;
  bf16* out_bf16;
};

__global__ void __launch_bounds__(256) patchify_kernel(const PatchifyParams p) {
  const int row_len = p.C * p.t_patch * p.patch * p.patch;
  const long long total = (long long)p.gt * p.gh * p.gw * row_len;
  const int m = p.merge;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int e = (int)(idx % row_len);
    long long r = idx / row_len;
    // row r = (t, h/m, w/m, h%m, w%m)
    const int wi = (int)(r % m); r /= m;
    const int hi = (int)(r % m); r /= m;
    const int wb = (int)(r % (p.gw / m)); r /= (p.gw / m);
    const int hb = (int)(r % (p.gh / m)); r /= (p.gh / m);
    const int t = (int)r;
    // element e = (c, tt, ph, pw)
    const int pw = e % p.patch; e /= p.patch;
    const int ph = e % p.patch; e /= p.patch;
    const int tt = e % p.t_patch; e /= p.t_patch;
    const int c = e;
    int f = t * p.t_patch + tt;
    if (f >= p.F) f = p.F - 1;                 // temporal padding repeats the last frame
    if (p.perm) f = p.perm[f];
    const int y = (hb * m + hi) * p.patch + ph;
    const int x = (wb * m + wi) * p.patch + pw;
    const long long src = (((long long)f * p.C + c) * p.H + y) * p.W + x;
    const float v = p.is_u8 ? (float)reinterpret_cast<const uint8_t*>(p.frames)[src]
                            : reinterpret_cast<const float*>(p.frames)[src];
    const float o = __fdiv_rn(__fsub_rn(v, p.mean[c]), p.stdv[c]);   // torchvision normalize: sub_ then div_
    if (p.out_f32) p.out_f32[idx] = o;
    if (p.out_bf16) p.out_bf16[idx] = __float2bfloat16_rn(o);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Bicubic antialiased resize (separable, horizontal pass then vertical pass with an fp32 intermediate), the arithmetic of
// torchvision.transforms.functional.resize(video_uint8, [h, w], BICUBIC, antialias=True) as qwen-vl-utils calls it
// (vision_process.py:310-315): torchvision converts the uint8 tensor to float32, runs ATen's _upsample_bicubic2d_aa
// (per output index a window [xmin, xmin + xsize) and normalised cubic weights, a = -0.5, support 2 * max(scale, 1)),
// then rounds (half to even) and clamps to [0, 255].  The per-index windows and weights are computed on the host
// (vision.py:aa_weight_table, fp32 like ATen) and passed in as tables.
// ------------------------------------------------------------------------------------------------------------------
struct ResizeParams {
  const void* src; int src_is_u8;
  float* tmp;                 // [planes][H][OW]
  void* dst; int dst_is_u8;   // [planes][OH][OW]
  int planes, H, W, OH, OW;
  const float* wh; const int* xmin_h; const int* xsize_h; int taps_h;   // [OW][taps_h]
  const float* wv; const int* xmin_v; const int* xsize_v; int taps_v;   // [OH][taps_v]
  int round_u8, use_fma;
};

SB_DEVICE float acc_step(float t, float v, float w, int use_fma) {
  return use_fma ? __fmaf_rn(v, w, t) : __fadd_rn(t, __fmul_rn(v, w));
}

__global__ void __launch_bounds__(256) resize_h_kernel(const ResizeParams p) {
  const long long total = (long long)p.planes * p.H * p.OW;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(idx % p.OW);
    const long long row = idx / p.OW;                     // plane * H + y
    const int x0 = p.xmin_h[ox], n = p.xsize_h[ox];
    const float* w = p.wh + (long long)ox * p.taps_h;
    const long long base = row * p.W + x0;
    float t = 0.f;
    if (p.src_is_u8) {
      const uint8_t* s = reinterpret_cast<const uint8_t*>(p.src) + base;
      t = __fmul_rn((float)s[0], w[0]);
      for (int j = 1; j < n; ++j) t = acc_step(t, (float)s[j], w[j], p.use_fma);
    } else {
      const float* s = reinterpret_cast<const float*>(p.src) + base;
      t = __fmul_rn(s[0], w[0]);
      for (int j = 1; j < n; ++j) t = acc_step(t, s[j], w[j], p.use_fma);
    }
    p.tmp[idx] = t;
  }
}

__global__ void __launch_bounds__(256) resize_v_kernel(const ResizeParams p) {
  const long long total = (long long)p.planes * p.OH * p.OW;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(idx % p.OW);
    const int oy = (int)((idx / p.OW) % p.OH);
    const long long plane = idx / ((long long)p.OW * p.OH);
    const int y0 = p.xmin_v[oy], n = p.xsize_v[oy];
    const float* w = p.wv + (long long)oy * p.taps_v;
    const float* s = p.tmp + (plane * p.H + y0) * p.OW + ox;
    float t = __fmul_rn(s[0], w[0]);
    for (int j = 1; j < n; ++j) t = acc_step(t, s[(long long)j * p.OW], w[j], p.use_fma);
    if (p.round_u8) t = fminf(fmaxf(rintf(t), 0.f), 255.f);
    if (p.dst_is_u8) reinterpret_cast<uint8_t*>(p.dst)[idx] = (uint8_t)t;
    else reinterpret_cast<float*>(p.dst)[idx] = t;
  }
}

}  // namespace

extern "C" int sb_resize_bicubic_aa(const void* src, int src_is_u8, int planes, int H, int W, float* tmp, void* dst,
                                    int dst_is_u8, int OH, int OW, const float* wh, const int* xmin_h,
                                    const int* xsize_h, int taps_h, const float* wv, const int* xmin_v,
                                    const int* xsize_v, int taps_v, int round_u8, int use_fma, sb_stream_t stream) {
  SB_REQUIRE(src && tmp && dst && wh && xmin_h && xsize_h && wv && xmin_v && xsize_v, "sb_resize_bicubic_aa: null pointer");
  SB_REQUIRE(planes > 0 && H > 0 && W > 0 && OH > 0 && OW > 0 && taps_h > 0 && taps_v > 0, "sb_resize_bicubic_aa: bad sizes");
  SB_REQUIRE(!dst_is_u8 || round_u8, "sb_resize_bicubic_aa: a uint8 destination needs round_u8");
  ResizeParams p;
  p.src = src; p.src_is_u8 = src_is_u8; p.tmp = tmp; p.dst = dst; p.dst_is_u8 = dst_is_u8;
  p.planes = planes; p.H = H; p.W = W; p.OH = OH; p.OW = OW;
  p.wh = wh; p.xmin_h = xmin_h; p.xsize_h = xsize_h; p.taps_h = taps_h;
  p.wv = wv; p.xmin_v = xmin_v; p.xsize_v = xsize_v; p.taps_v = taps_v;
  p.round_u8 = round_u8; p.use_fma = use_fma;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  auto blocks = [](long long total) { long long b = (total + 255) / 256; return (unsigned)(b > 148 * 32 ? 148 * 32 : b); };
  resize_h_kernel<<<blocks((long long)planes * H * OW), 256, 0, st>>>(p);
  if (sb_check_launch("sb_resize_bicubic_aa(h)")) return 1;
  resize_v_kernel<<<blocks((long long)planes * OH * OW), 256, 0, st>>>(p);
  return sb_check_launch("sb_resize_bicubic_aa(v)");
}

extern "C" int sb_video_patchify(const void* frames, int frames_are_u8, int F, int C, int H, int W, const int* perm,
                                 const float* mean_host, const float* std_host, int patch, int t_patch, int merge,
                                 float* out_f32, void* out_bf16, sb_stream_t stream) {
  SB_REQUIRE(frames && mean_host && std_host && (out_f32 || out_bf16), "sb_video_patchify: null pointer");
  SB_REQUIRE(F > 0 && C > 0 && C <= 4 && patch > 0 && t_patch > 0 && merge > 0, "sb_video_patchify: bad sizes");
  SB_REQUIRE(H % (patch * merge) == 0 && W % (patch * merge) == 0,
             "sb_video_patchify: frame size %dx%d is not a multiple of patch*merge = %d (resize first)", H, W, patch * merge);
  PatchifyParams p;
  p.frames = frames; p.is_u8 = frames_are_u8; p.perm = perm;
  p.F = F; p.C = C; p.H = H; p.W = W; p.patch = patch; p.t_patch = t_patch; p.merge = merge;
  p.gt = (F + t_patch - 1) / t_patch; p.gh = H / patch; p.gw = W / patch;
  for (int c = 0; c < C; ++c) { p.mean[c] = mean_host[c]; p.stdv[c] = std_host[c]; }
  p.out_f32 = out_f32; p.out_bf16 = reinterpret_cast<bf16*>(out_bf16);
  const long long total = (long long)p.gt * p.gh * p.gw * C * t_patch * patch * patch;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  patchify_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  return sb_check_launch("sb_video_patchify");
}
