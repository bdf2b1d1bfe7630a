/* spacer_b200 -- C ABI of the B200-native SG-RLVR hot path (libspacer_b200.so).
 *
 * The reference (OuyangKun10/SpaceR) has no FFI of its own: its hot path is Python
 * (`SGRLVRTrainer.compute_loss`, SpaceR-SG-RLVR/src/r1-v/src/open_r1/trainer/SG_RLVR_trainer.py:384-686)
 * that reaches the GPU through transformers' `Qwen2VLForConditionalGeneration`
 * (transformers 5.5.0, models/qwen2_vl/modeling_qwen2_vl.py -- "MQ2" below), `generate()`
 * (generation/utils.py:2658-2830) and ATen.  This header is the boundary a maintainer binds instead:
 * every entry point names the reference call it replaces.  The ctypes binding that ships with this
 * repo is spacer_b200/_lib.py; INTEGRATION.md shows the stub on the reference side.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name ends in _host;
 *   - bf16 tensors are `void*` to 2-byte elements, row-major, 16-byte aligned;
 *   - all work is enqueued on `stream` (a cudaStream_t); nothing synchronises with the host unless the
 *     comment says so;
 *   - return 0 on success; non-zero on error with the message available from sb_last_error()
 *     (thread-local).  There is no CPU fallback anywhere in this library.
 */
#ifndef SPACER_B200_H
#define SPACER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* sb_stream_t;

#define SB_ABI_VERSION 4

const char* sb_last_error(void);
int sb_abi_version(void);
/* fills SM count and compute capability of the current device; error if it is not sm_100 */
int sb_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* number of kernels this library has launched (successfully enqueued) since the last reset; a launch recorded
 * into a CUDA graph during stream capture counts once -- replays are the caller's to add.  count_host may be NULL. */
int sb_launch_counter(long long* count_host, int reset);
/* decode-step kernels are launched with programmatic dependent launch (next kernel's prologue and weight prefetch
 * overlap the current kernel's tail); 0 turns it off (debugging / A-B timing).  Default on unless SB_NO_PDL is set. */
int sb_set_pdl(int enable);
/* decode GEMVs tag their weight stream evict_first (createpolicy cache hint): a weight is read once per step, so it
 * should not push out what is reused (activations, split-K partials, KV, lines prefetched by sb_dec_l2_prefetch);
 * 0 turns the hint off (A-B timing).  Default on unless SB_NO_L2_HINTS is set. */
int sb_set_dec_l2_hints(int enable);
/* profiling aid: with a device buffer of 1 + 4*capacity uint64 installed (buf[0] = 0), every decode-step kernel
 * appends {kind, t_entry, t_ready (dependencies satisfied), t_end} of its block 0 in globaltimer ns; buf[0] counts
 * records.  (NULL, 0) disables.  Not for production runs. */
int sb_trace_enable(unsigned long long* buf_dev, int capacity);
/* profiling aid: per-CTA timeline of the decode GEMVs.  buf: n_sms lists of 1 + 3*capacity_per_cta uint64, zeroed; CTA b
 * appends {M ^ (K << 32), t_ready, t_end} to list b (element 0 = count).  (NULL, 0) disables. */
int sb_trace_enable_gemv_ctas(unsigned long long* buf_dev, int capacity_per_cta);

/* ------------------------------------------------------------------------------------------------
 * GEMM  D[M,N] = epi(A[M,K] * B[N,K]^T)   (tcgen05 + TMEM + TMA)
 * replaces: torch.nn.functional.linear / Conv3d-as-GEMM under MQ2:304-310 (patch embed), :401-405,
 * :329-337 (ViT), :559-593 (attention projections), :502-504 (SwiGLU MLP), :1437-1438 (lm_head),
 * and autograd's backward GEMMs for the same.
 * ------------------------------------------------------------------------------------------------ */
enum {
  SB_EPI_STORE = 0,     /* D = bf16(acc + bias) [+ residual]                                        */
  SB_EPI_QUICKGELU = 1, /* z = acc + bias; aux = z (optional); D = z * sigmoid(1.702 z)   (ViT fc1) */
  SB_EPI_GELU = 2,      /* z = acc + bias; aux = z (optional); D = gelu_erf(z)       (PatchMerger)  */
  SB_EPI_SWIGLU = 3,    /* B rows (and bias, if any) interleaved [64 gate | 64 up]; D[M,N/2] = silu(g)*u; aux = raw */
  SB_EPI_F32T = 4,      /* D_f32[split][n][m] = partial acc (swap-AB decode GEMV with split-K)      */
  SB_EPI_LMHEAD = 5,    /* per-row (max, sumexp) per N tile of bf16-rounded logits + target gather  */
  SB_EPI_DLOGITS = 6,   /* D = bf16(coef[m] * (onehot(target[m]) - exp(logit - lse[m])))            */
  SB_EPI_F32T_SWIGLU = 7  /* swap-AB decode gate|up GEMV, A rows interleaved [64 gate | 64 up]:
                           D_bf16[n][m/2] = silu(gate)*up for the N <= 32 decode rows (no split-K)         */
};

typedef struct sb_gemm_args {
  int M, N, K;
  const void* A; long long lda; int a_mn; /* a_mn=0: A is [M,K] row-major; 1: A stored as [K,M]     */
  const void* B; long long ldb; int b_mn; /* b_mn=0: B is [N,K] row-major; 1: B stored as [K,N]     */
  void* D; long long ldd;
  int epilogue;
  int k_splits;                           /* >1 only with SB_EPI_F32T                               */
  int bn;                                 /* force tile N (0 = auto)                                */
  const void* bias;                       /* bf16 [N] or NULL                                       */
  const void* residual; long long ldr;    /* bf16 [M,N] or NULL (may alias D for accumulation)      */
  void* aux; long long ldaux;             /* bf16 pre-activation / raw gate-up output or NULL       */
  const int* targets;                     /* [M] LMHEAD / DLOGITS                                   */
  float* lse_part;                        /* [M][ceil(N/256)][2] LMHEAD out: (max, sumexp)          */
  float* tgt_logit;                       /* [M] LMHEAD out                                         */
  const float* lse;                       /* [M] DLOGITS in                                         */
  const float* coef;                      /* [M] DLOGITS in: dLoss/dlogprob                         */
} sb_gemm_args;

int sb_gemm(const sb_gemm_args* args, sb_stream_t stream);
/* number of non-empty K splits sb_gemm will use for (K, k_splits) */
int sb_gemm_effective_splits(int K, int k_splits);

/* ------------------------------------------------------------------------------------------------
 * Bandwidth-bound kernels (fp32 math, bf16 storage)
 * ------------------------------------------------------------------------------------------------ */
/* pixel_values fp32 -> bf16 (the `.to(dtype)` at MQ2:306) */
int sb_cast_f32_bf16(const float* src, void* dst, long long n, sb_stream_t stream);
/* torch.nn.LayerNorm(E, eps) of the ViT blocks / merger (MQ2:464-465,317); mean/rstd [T] optional outputs */
int sb_layernorm_fwd(const void* x, const void* w, const void* b, void* y, float* mean, float* rstd, int T, int E,
                     float eps, sb_stream_t stream);
/* dx = [dres +] dLN(dy); dw/db are fp32 accumulators [E] (+=) */
int sb_layernorm_bwd(const void* x, const void* w, const float* mean, const float* rstd, const void* dy,
                     const void* dres, void* dx, float* dw, float* db, int T, int E, sb_stream_t stream);
/* Qwen2VLRMSNorm (MQ2:117-131) */
int sb_rmsnorm_fwd(const void* x, const void* w, void* y, float* rstd, int T, int H, float eps, sb_stream_t stream);
int sb_rmsnorm_bwd(const void* x, const void* w, const float* rstd, const void* dy, const void* dres, void* dx,
                   float* dw, int T, int H, sb_stream_t stream);
/* apply_rotary_pos_emb_vision + rot_pos_emb (MQ2:257-268,725-752) in place on q,k of qkv [T, 3*heads*hd];
 * grids_dev: int32 [n_grids][3] (t,h,w) on the device; inverse=1 applies the transpose rotation (backward) */
int sb_rope_vit(void* qkv, int T, int heads, int head_dim, const int* grids_dev, int n_grids, int merge, int inverse,
                sb_stream_t stream);
/* same rotation with an explicit (h, w) position per token, pos_hw int32 [T][2] on the device: the window-reordered
 * sequence of Qwen2.5-VL's vision tower (modeling_qwen2_5_vl.py:382-409, 474-482) */
int sb_rope_vit_pos(void* qkv, int T, int heads, int head_dim, const int* pos_hw, int inverse, sb_stream_t stream);
/* apply_multimodal_rotary_pos_emb (MQ2:188-254) in place on q,k of qkv [T,(nh+2nkv)*hd]; pos int32 [3,T];
 * optionally copies the rotated k and v into a KV cache (row stride kv_ld elements) */
int sb_mrope(void* qkv, const int* pos, int T, int n_heads, int n_kv_heads, int head_dim, float theta, int sec_t,
             int sec_h, int inverse, void* k_out, void* v_out, long long kv_ld, sb_stream_t stream);
/* activation recompute / backward; mode 0 = quick_gelu (activations.py:117), 1 = gelu(erf) */
int sb_act_fwd(const void* z, void* f, long long n, int mode, sb_stream_t stream);
int sb_act_bwd(const void* z, const void* dy, void* dz, long long n, int mode, sb_stream_t stream);
/* SwiGLU backward on the interleaved raw [gate|up] output: dgu from dact; optional recompute of act */
int sb_swiglu_bwd(const void* gu, const void* dact, void* dgu, void* act, int T, int I, sb_stream_t stream);
/* running index of placeholder (video/image) tokens: vis_idx[t] = k for the k-th placeholder, else -1 */
int sb_vision_index(const int* ids, int* vis_idx, int T, int video_id, int image_id, int* count_out,
                    sb_stream_t stream);
/* embed_tokens lookup + masked_scatter of vision embeddings (MQ2:1255-1272) */
int sb_embed_merge(const int* ids, const int* vis_idx, const void* embed, const void* vision, void* out, int T, int H,
                   int n_vision, sb_stream_t stream);
int sb_embed_bwd(const int* ids, const int* vis_idx, const void* dx, void* d_embed, void* d_vision, int T, int H,
                 int n_vision, sb_stream_t stream);
int sb_gather_rows(const void* src, const int* rows, void* dst, int R, int H, sb_stream_t stream);
int sb_scatter_add_rows(const void* src, const int* rows, void* dst, int R, int H, sb_stream_t stream);
/* deterministic scatter-add: dst[seg_dst[s]] (+)= sum_{k in [seg_off[s], seg_off[s+1])} src[order[k]], fp32 accumulation in
 * list order, one rounding (embedding gradient by token id, d_hidden of repeated lm_head rows); int32 device arrays */
int sb_segment_sum_rows(const void* src, const int* order, const int* seg_off, const int* seg_dst, int n_seg, void* dst,
                        int H, int accumulate, sb_stream_t stream);
/* out_f32[n] += sum_t dy[t][n]   (bias gradients) */
int sb_colsum(const void* dy, float* out, int T, int N, long long ld, sb_stream_t stream);
int sb_f32_to_bf16_2d(const float* src, void* dst, int T, int W, long long ldd, sb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Video front-end: frames [F][C][H][W] (uint8, or float32 holding 0..255 as qwen_vl_utils.fetch_video returns them)
 * -> pixel_values_videos [gt*gh*gw][C*t_patch*patch*patch] with gt = ceil(F/t_patch), gh = H/patch, gw = W/patch.
 * replaces Qwen2VLVideoProcessor._preprocess (video_processing_qwen2_vl.py:240-272: fused rescale+normalise in fp32,
 * last-frame temporal padding, patch permutation), T-GRPO's second processor pass over shuffled frames
 * (SG_RLVR_trainer.py:442-458; `perm`, device int32 [F], output frame f reads source frame perm[f]) and the bf16 cast
 * at MQ2:306.  mean_host/std_host: C floats on the HOST, already divided by the rescale factor (mean*255, std*255).
 * out_f32 is bit-exact with the HF CPU path; out_bf16 is what sb_gemm consumes.  Either output may be NULL.
 * ------------------------------------------------------------------------------------------------ */
/* Bicubic antialiased resize of `planes` = F*C images [H][W] -> [OH][OW]: the arithmetic of
 * torchvision.transforms.functional.resize(uint8 video, [OH, OW], BICUBIC, antialias=True) in qwen-vl-utils' fetch_video
 * (vision_process.py:310-315; float32 separable passes, horizontal first, then round-half-even + clamp to [0, 255]).
 * wh/xmin_h/xsize_h: per output column the normalised weights [OW][taps_h], first source column and tap count (device
 * tables, built by spacer_b200/vision.py:aa_weight_table with ATen's formula); wv/...: the same per output row.
 * tmp: fp32 scratch [planes][H][OW].  dst float32 (what fetch_video returns after .float()) or uint8. */
int sb_resize_bicubic_aa(const void* src, int src_is_u8, int planes, int H, int W, float* tmp, void* dst, int dst_is_u8,
                         int OH, int OW, const float* wh, const int* xmin_h, const int* xsize_h, int taps_h,
                         const float* wv, const int* xmin_v, const int* xsize_v, int taps_v, int round_u8, int use_fma,
                         sb_stream_t stream);
int sb_video_patchify(const void* frames, int frames_are_u8, int F, int C, int H, int W, const int* perm,
                      const float* mean_host, const float* std_host, int patch, int t_patch, int merge, float* out_f32,
                      void* out_bf16, sb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Attention (flash-style, varlen/prefix mask).  replaces flash_attn_varlen_func / sdpa under MQ2:415-454
 * (ViT, head_dim 80, block-diagonal) and MQ2:575-590 (Qwen2 causal GQA, head_dim 128).
 * meta: int32 [T][4] = (prefix_len, seg_start, kv_end, 0): key j visible to query t iff
 *       j < prefix_len || seg_start <= j < kv_end.
 * ------------------------------------------------------------------------------------------------ */
typedef struct sb_attn_args {
  const void* q; const void* k; const void* v; /* bf16; head h at column h*head_dim (kv head for k, v)   */
  long long ldq, ldk, ldv;
  void* o; long long ldo;                      /* bf16 [T, n_heads*head_dim]                             */
  float* lse;                                  /* fp32 [n_heads][T] (natural log); required for backward */
  const int* meta;
  int T, Tk;                                   /* queries / keys (Tk = 0 -> T)                           */
  int n_heads, n_kv_heads, head_dim;
  float scale;
  /* backward only */
  const void* d_o; long long lddo;
  float* delta;                                /* fp32 [n_heads][T] scratch                              */
  void* dq; long long lddq;                    /* bf16 output, head h at column h*head_dim               */
  void* dk; void* dv; long long lddk, lddv;    /* bf16 outputs, kv head h at column h*head_dim           */
  void* gqa_ws;                                /* bf16 scratch, sb_attn_bwd_workspace() elements (GQA)   */
  int* tile_ws;                                /* int32 scratch, sb_attn_bwd_workspace() ints            */
} sb_attn_args;
int sb_attn_fwd(const sb_attn_args* args, sb_stream_t stream);
/* forward implementation: 0 = tcgen05/TMEM/TMA kernel (default), 1 = mma.sync kernel (A/B timing, cross-checks) */
int sb_set_attn_impl(int impl);
/* deterministic (no atomics): a key-tile-major kernel produces dK, dV, a query-tile-major kernel produces dQ */
int sb_attn_bwd(const sb_attn_args* args, sb_stream_t stream);
/* backward implementation: 0 = tcgen05/TMEM/TMA kernels (default); bit 0 = dQ on the mma.sync kernel, bit 1 = dK/dV on
 * the mma.sync kernel (A/B timing, cross-checks) */
int sb_set_attn_bwd_impl(int impl);
int sb_attn_bwd_workspace(int T, int Tk, int n_heads, int n_kv_heads, int head_dim, long long* gqa_ws_elems,
                          long long* tile_ws_ints);

/* ------------------------------------------------------------------------------------------------
 * Decode step (q_len = 1): consumers of the split-K GEMV partials parts[s][r][n]
 * (element (s,r,n) at parts + s*stride_s + r*stride_r + n).  `step_ptr` is a device int (the number of
 * completion tokens already in the cache) so that one captured CUDA graph serves every step.
 * replaces the per-token iterations of GenerationMixin._sample (generation/utils.py:2743-2806) through
 * Qwen2VLDecoderLayer (MQ2:597-662) and DynamicCache.update (cache_utils.py:102-120).
 * ------------------------------------------------------------------------------------------------ */
int sb_dec_embed(const int* tokens, const void* embed, void* x, int R, int H, sb_stream_t stream);
int sb_dec_residual_rmsnorm(void* x, const float* parts, int S, long long stride_s, long long stride_r, const void* w,
                            void* xn, int R, int H, float eps, sb_stream_t stream);
int sb_dec_qkv_post(const float* parts, int S, long long stride_s, long long stride_r, const void* bias,
                    const int* step_ptr, int rope_base, float theta, int n_heads, int n_kv_heads, int head_dim,
                    void* q_out, void* k_cache, void* v_cache, long long cache_stride_r, int c_max, int R,
                    sb_stream_t stream);
/* prefix-shared flash decoding: rows < rows_group0 share prompt cache kp0/vp0 [P][nkv*hd], the rest kp1/vp1; every
 * row also attends to its own completion cache [c_max][nkv*hd] slots 0..*step_ptr.  The shared prompt K/V is read once
 * per group (queries of all its rows batched on tensor cores).  workspace: fp32 scratch of at least
 * sb_dec_attn_workspace() floats.  out bf16 [R][n_heads*hd] */
int sb_dec_attn_workspace(int R, int rows_group0, int P, int c_max, int n_heads, int n_kv_heads, long long* floats_out);
int sb_dec_attn(const void* q, const void* kp0, const void* vp0, const void* kp1, const void* vp1, int rows_group0,
                int P, const void* k_cache, const void* v_cache, long long cache_stride_r, int c_max,
                const int* step_ptr, int n_heads, int n_kv_heads, int head_dim, float scale, float* workspace,
                long long workspace_floats, void* out, int R, sb_stream_t stream);
int sb_dec_swiglu(const float* parts, int S, long long stride_s, long long stride_r, void* act, int R, int I,
                  sb_stream_t stream);
/* asks L2 (cp.async.bulk.prefetch.L2, evict_last) for weights a later GEMV of the decode step will stream: `chunks` runs of
 * chunk_bytes each, stride_bytes apart (chunks <= 1: one run) -- e.g. the first rows of every 128-row tile of gate|up.
 * A bulk prefetch holds its issuing CTA until the memory system has taken it, so this is its own small kernel, meant for
 * a side stream / parallel graph branch: nothing on the layer's dependency chain waits for it.  ctas (0 = one per SM)
 * bounds the request rate: every issuing SM adds ~0.15 TB/s, and a request stream faster than HBM drains fills the
 * memory system's queues, which the latency-bound small kernels of the layer then wait behind.
 * pace_ns >= 0 selects the load/store-path form instead: one prefetch.global.L2::evict_last per 128-byte line from
 * `ctas` CTAs of 128 threads, each thread sleeping pace_ns between two requests (no TMA queue involved). */
int sb_dec_l2_prefetch(const void* base, long long chunk_bytes, long long stride_bytes, int chunks, int ctas, int pace_ns,
                       sb_stream_t stream);
/* top-p sampling of one token per row from fp32 logits [R][ld]; TopPLogitsWarper + multinomial semantics
 * (logits_process.py:521-533, generation/utils.py:2789-2797).  out_ids[r][*step_ptr] = token (optional).
 * seed_dev (optional device scalar) overrides `seed`, so that a captured CUDA graph can be replayed with new seeds */
int sb_sample_top_p(const float* logits, long long ld, int R, int V, float top_p, unsigned long long seed,
                    const int* step_ptr, int* finished, int* out_tokens, int* out_ids, long long out_ld,
                    float* out_logprob, int eos_id, int pad_id, int suppress_eos, const long long* seed_dev,
                    sb_stream_t stream);
/* greedy decoding: argmax of the bf16-rounded logits, lowest index on ties (do_sample=False; SpaceR-Eval's
 * `generate(..., temperature=0.01)` over a top_k = 1 generation config, SpaceR-Eval/data_utils/vsibench.py:174);
 * same EOS / pad / out_ids bookkeeping as sb_sample_top_p */
int sb_sample_greedy(const float* logits, long long ld, int R, int V, const int* step_ptr, int* finished,
                     int* out_tokens, int* out_ids, long long out_ld, int eos_id, int pad_id, int suppress_eos,
                     sb_stream_t stream);
/* One sampler entry for every decoding mode generate() is called with (training rollouts: GenerationConfig(do_sample,
 * top_p 0.95, temperature 1) -> with HF's default top_k = 50, SG_RLVR_trainer.py:277-302; evaluation: the checkpoint's
 * generation_config.json (repetition_penalty 1.05, top_k 1) with temperature 0.01, SpaceR-Eval/data_utils/vsibench.py:174).
 * Order of the transformations = HF's logits-processor list: repetition penalty (logits_process.py
 * RepetitionPenaltyLogitsProcessor: score < 0 ? score * p : score / p for every token already in prompt + completion),
 * temperature, top-k (ties at the k-th value are kept), top-p, multinomial.  A row is finished by ANY id of eos_ids.
 * Values changed by penalty / temperature are rounded to bf16 again before the radix select (HF keeps fp32). */
typedef struct sb_sample_args {
  const float* logits; long long ld; int R; int V;
  int mode;                         /* 0 = sample, 1 = greedy argmax (lowest index on ties)                        */
  float top_p;                      /* (0, 1]; 1 disables the nucleus cut                                          */
  int top_k;                        /* 0 disables                                                                  */
  float temperature;                /* > 0                                                                         */
  float repetition_penalty;         /* 1 disables; otherwise `seen` is required                                    */
  unsigned int* seen; long long seen_ld; /* bitmap [R][seen_ld] words (bit t of row r: token t occurred); the
                                       sampled token is OR-ed in.  NULL = no bookkeeping                           */
  unsigned long long seed; const long long* seed_dev;
  const int* step_ptr; int* finished; int* out_tokens; int* out_ids; long long out_ld; float* out_logprob;
  int eos_ids[4]; int n_eos; int pad_id; int suppress_eos;
} sb_sample_args;
int sb_sample(const sb_sample_args* args, sb_stream_t stream);
/* seen[r][ids[i]] = 1 for all i < n and r < rows (the prompt tokens count for the repetition penalty) */
int sb_token_bitmap_set(const int* ids, int n, unsigned int* seen, long long seen_ld, int rows, int V,
                        sb_stream_t stream);
int sb_step_advance(int* step_ptr, sb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * GRPO loss tail, forward + backward (SG_RLVR_trainer.py:353-366, 489-494, 551-552, 632-643)
 * inputs are the LMHEAD-epilogue partials of the lm_head GEMM over the G*C scoring rows.
 * out2 = {loss, mean_kl}; coef_out = dLoss/dlogprob per token (0 outside the completion mask).
 * `workspace` holds sb_grpo_loss_workspace() floats which the caller zeroes ONCE (it starts with ticket counters the
 * kernel resets itself).
 * ------------------------------------------------------------------------------------------------ */
int sb_grpo_loss_workspace(int G, int C, long long* floats_out);
int sb_grpo_loss(const float* lse_part, int n_tiles, const float* tgt_logit, const int* comp_ids, int G, int C,
                 int eos_id, const float* ref_lp, const float* adv, float beta, float* lp_out, float* lse_out,
                 float* coef_out, int* mask_out, float* row_loss, float* row_kl, int* row_len, float* out2,
                 float* workspace, sb_stream_t stream);
int sb_logprob_from_partials(const float* lse_part, int n_tiles, const float* tgt_logit, float* lp_out, long long rows,
                             sb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * AdamW + global-norm clip over flat arenas (run_SpaceR_SG_RLVR.sh:23-25,37; zero3.json:10-12)
 * ------------------------------------------------------------------------------------------------ */
int sb_grad_sumsq(const void* g, long long n, int grad_is_f32, float* total_sq, sb_stream_t stream);
int sb_adamw_step(void* param_bf16, float* master, void* m, void* v, const void* grad, long long n, int grad_is_f32,
                  int moments_are_bf16, const float* total_sq, float lr, float beta1, float beta2, float eps,
                  float weight_decay, int step, float max_norm, float grad_scale, sb_stream_t stream);
int sb_bf16_to_f32(const void* src, float* dst, long long n, sb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SPACER_B200_H */
