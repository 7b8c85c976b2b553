/*
 * lm2a_b200.h — C ABI of the B200-native LM2A reverse-diffusion sampling path.
 *
 * One shared library (liblm2a_b200.so), `extern "C"`, plain pointers and sizes.
 * Every entry point:
 *   - takes DEVICE pointers borrowed from the caller (never allocates, frees or
 *     synchronises), launches on the `stream` handle given (a cudaStream_t cast
 *     to void*), and is therefore CUDA-Graph capturable;
 *   - returns 0 on success, non-zero on an argument / shape / architecture
 *     violation *before* any launch (message via lm2a_last_error());
 *   - is sm_100a only: no fallback, no CPU path.
 *
 * Activation layout ("slab"): channels-last bf16 [R, Tp, ld] flattened to
 * M = R*Tp slots; slot (r, t) is valid for t < T, every slot t >= T is ZERO.
 * The shared zero slot between consecutive clips is the conv zero padding, so
 * a k=3 conv is three shifted 2-D TMA boxes over the flattened slab.
 *
 * Reference symbols replaced (paths relative to the reference repo):
 *   lm2a_conv1d_bf16   nn.Conv1d k1/k3/k4s2 + bias (+GroupNorm/SiLU of its input,
 *                      +FiLM, +skip conv, +residual, +GroupNorm sums of its output)
 *                      models/unet1d_ultimate.py:87-88,115,138-148,159,216-221,
 *                      255-261,295,364 and every nn.Linear / MHA in/out
 *                      projection of models/cross_attention.py:19-36,46-65
 *   lm2a_gn_silu_bf16  nn.GroupNorm + nn.SiLU, unet1d_ultimate.py:91-95,136-137,
 *                      146-147,362-363
 *   lm2a_cross_attn_bf16  softmax(q k^T) v core of nn.MultiheadAttention as used
 *                      at models/cross_attention.py:50-61
 *   lm2a_time_mlp / lm2a_time_embed / lm2a_film   models/embedding.py:19-43,
 *                      unet1d_ultimate.py:43-65; legacy models/unet1d.py:23,48-49
 *   lm2a_upsample2x_bf16  F.interpolate(linear, align_corners=True) :231-236
 *   lm2a_ingest_x      torch.cat([x, x]) + layout change, sample.py:162
 *   lm2a_cfg_posterior sample.py:167-174 (CFG blend + clamps) and
 *                      sample.py:186-210 == models/diffusion.py:71-102
 *   lm2a_cfg_step      sample.py:162-210 in one launch (blend + posterior + noise drawn
 *                      in the kernel + next step's input slab); lm2a_philox_normal
 *                      torch.randn of sample.py:136,204
 *   lm2a_cfg_ddim      models/diffusion.py:124-165 (ddim_sample)
 *   lm2a_resample_seq  datasetcode/dataset.py:49-87 (match_len 'interp')
 *   lm2a_mel_metrics   val.py:25-113 (compute_metrics)
 *   lm2a_adan_step     models/adan.py:34-114 + EMA update train.py:177-180
 */
#ifndef LM2A_B200_H
#define LM2A_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LM2A_ABI_VERSION 11

/* ---- library ---------------------------------------------------------- */
int lm2a_abi_version(void);
const char* lm2a_last_error(void);
/* 0 iff the current device is compute capability 10.x (B200). */
int lm2a_check_device(void);
/* number of kernels launched by this library since load / last reset */
int64_t lm2a_launch_count(void);
void lm2a_reset_launch_count(void);

/* ---- implicit-GEMM conv1d / linear (tcgen05 + TMEM + TMA) ------------- */
enum { LM2A_TAPS_K1 = 0, LM2A_TAPS_K3 = 1, LM2A_TAPS_K4S2 = 2 };
enum { LM2A_OUT_BF16_SLAB = 0, LM2A_OUT_F32_NCT = 1 };

typedef struct lm2a_conv_seg {
  const void* x;      /* bf16 slab base (already offset to first channel)     */
  int64_t rows;       /* slots in the slab (R*Tp of the INPUT)                */
  int32_t ld;         /* slot pitch in elements (multiple of 8)               */
  int32_t cin;        /* channels consumed (multiple of 64)                   */
  int32_t taps;       /* LM2A_TAPS_*; K4S2 needs rows == 2 * M                */
  int32_t _pad;
} lm2a_conv_seg;

typedef struct lm2a_conv_desc {
  lm2a_conv_seg seg[2];   /* seg[1].x == NULL when unused (fused skip conv /
                             second operand of a virtual concat)              */
  const void* w;          /* bf16 [n_pad, k_total], k = seg, tap, channel     */
  int32_t n_pad;          /* rows of w, multiple of 128                       */
  int32_t n_valid;        /* real output channels                             */
  int64_t m;              /* output slots = R * tp                            */
  int32_t tp;             /* output slot pitch per clip                       */
  int32_t t_valid;        /* output slots t >= t_valid are written as zero    */
  const float* bias;      /* [n_pad] fp32                                     */
  const float* film;      /* NULL or fp32 [R, film_ld]: scale at col n,
                             shift at col film_shift_off + n                  */
  int32_t film_ld;
  int32_t film_shift_off;
  const void* residual;   /* NULL or bf16 slab added in the epilogue          */
  int32_t res_ld;
  int32_t out_mode;       /* LM2A_OUT_*                                       */
  void* out;              /* bf16 slab [m, out_ld] or fp32 [R, n_valid, t_valid] */
  int32_t out_ld;
  int32_t block_n;        /* 0 = auto, else 128 or 256                        */
  /* Optional exact GroupNorm sums of the OUTPUT (bf16 slab mode only), consumed by
   * the next conv's in_gn_* transform or by lm2a_gn_apply_bf16: two int64 per
   * (clip-row r relative to this launch, group g = (stats_c0 + c) / stats_cg of
   * output channel c relative to `out`) at stats[(r * stats_pitch + g) * 2 + {0, 1}] =
   * {sum * 2^24, sum of squares * 2^20} in fixed point. Each lane converts the fp32
   * sums of its own slot and the rest is integer addition (per-lane running sums,
   * hardware warp reductions, one atomic per (clip-row, group) and warp), so the
   * totals are exact and do not
   * depend on tile shape, batch position or sharding. The buffer must be ZERO when
   * the first producer of a step runs (lm2a_ingest_x clears a region for this). */
  void* stats;
  int32_t stats_pitch;    /* groups per clip-row in the stats buffer            */
  int32_t stats_cg;       /* channels per group (multiple of 8)                 */
  int32_t cta_group;      /* 0 = auto, 1 = one CTA per 128-row tile, 2 = CTA
                             pair (cluster of 2, cta_group::2 UMMA) per 256 rows */
  int32_t stats_c0;       /* channel of the normalised tensor that `out` channel 0
                             is (two producers of one concat slab share a buffer) */
  /* Optional GroupNorm + SiLU of the INPUT of seg[0] (unet1d_ultimate.py:136-137,
   * 146-147, 362-363), applied to the operand tile in shared memory between its
   * TMA landing and the MMA: y = SiLU((x - mean) * rstd * gamma + beta) with
   * mean / rstd per (clip-row, group) from in_gn_stats (layout as `stats` above,
   * written by the kernel that produced the slab; in_gn_pitch groups per clip-row,
   * in_gn_groups groups over seg[0].cin channels). k1 / k3 segments only; seg[1]
   * (skip conv) always reads raw. cin <= 2048; clips must be long enough that a
   * 130-slot tile touches at most 512 / in_gn_groups of them.                   */
  const void* in_gn_stats;  /* NULL = raw operand                               */
  const float* in_gn_gamma; /* [cin]                                            */
  const float* in_gn_beta;  /* [cin]                                            */
  int32_t in_gn_pitch;
  int32_t in_gn_groups;
  float in_gn_eps;
  int32_t in_gn_silu;       /* must be 1 (GroupNorm without SiLU: lm2a_gn_apply_bf16) */
  /* Optional x2 linear upsampling (align_corners = True) of seg[0] fused into the
   * conv (UpSampleConv, unet1d_ultimate.py:210-239): seg[0] then describes the
   * LOW-resolution slab (rows = m / 2 slots of pitch in_up_tp, in_up_t valid per
   * clip, k3 taps) and the operand tiles are interpolated from it on the fly -
   * same arithmetic as lm2a_upsample2x_bf16, so the result is bit-identical to
   * upsampling first. tp = 2 * in_up_tp, t_valid = 2 * in_up_t. 0 = off.        */
  int32_t in_up_tp;
  int32_t in_up_t;
  /* K walk order of launches without an operand transform: 0 = taps outer, channel
   * blocks inner (default, fastest); 1 = channel blocks outer, taps inner - the order
   * the in_gn_* / in_up_* launches accumulate in, so that a plain launch over a
   * pre-normalised / pre-upsampled slab reproduces them bit for bit.              */
  int32_t k_order;
  int32_t _pad0;
} lm2a_conv_desc;

int lm2a_conv1d_bf16(void* stream, const lm2a_conv_desc* d);

/* ---- fp32 validation path -------------------------------------------------- */
/* The same launch plan on fp32 slabs and fp32 weights with plain CUDA-core kernels
 * (csrc/ref_f32.cu): selected by UNet1D_ultimate(..., precision="fp32") to check the
 * single-step eps prediction against the reference's fp32 PyTorch path at 1e-4
 * relative (BASELINE.json tolerance (i)). Not a performance path. lm2a_conv1d_f32
 * takes the SAME descriptor as lm2a_conv1d_bf16 with every slab / weight / residual
 * pointer addressing fp32 data (LM2A_OUT_BF16_SLAB then means "fp32 slab"); input
 * GroupNorm + SiLU (exact expf) and x2 upsampling are evaluated on the fly, the
 * output sums use the same fixed-point format. block_n / cta_group / k_order are
 * ignored. lm2a_cross_attn_f32 reads K | V straight from the projection output
 * kv_s fp32 [slots*lk, kv_ld] (K at channel 0, V at channel e).               */
int lm2a_conv1d_f32(void* stream, const lm2a_conv_desc* d);
int lm2a_cross_attn_f32(void* stream, const float* q, int32_t q_ld, float* o,
                        int32_t o_ld, const float* kv_motion, const float* kv_text,
                        int32_t kv_ld, const int32_t* kv_slot, int32_t slots,
                        int32_t rows, int32_t tp, int32_t t_valid, int32_t lk,
                        int32_t e, int32_t heads, int32_t n_streams);
int lm2a_bias_add_f32(void* stream, const float* x, int32_t x_ld, float* y,
                      int32_t y_ld, const float* bias, int64_t slots, int32_t tp,
                      int32_t t_valid, int32_t c, void* stats, int32_t stats_pitch,
                      int32_t stats_cg, int32_t stats_c0);
int lm2a_ingest_x_f32(void* stream, const float* x, float* slab, int32_t batch,
                      int32_t copies, int32_t c, int32_t t, int32_t tp, int32_t ld,
                      void* zero, int64_t zero_bytes);
int lm2a_ingest_seq_f32(void* stream, const float* x, float* slab, int32_t rows,
                        int32_t t, int32_t c, int32_t tp, int32_t ld);

/* ---- GroupNorm + SiLU over a slab -------------------------------------- */
/* x,y: bf16 slabs [R, tp, ld*]; stats over t < t_valid and c/groups channels */
int lm2a_gn_silu_bf16(void* stream, const void* x, int32_t x_ld, void* y,
                      int32_t y_ld, const float* gamma, const float* beta,
                      int32_t rows, int32_t tp, int32_t t_valid, int32_t c,
                      int32_t groups, float eps, int32_t apply_silu);

/* GroupNorm + SiLU as one streaming pass over a slab whose exact sums were
 * produced by the kernel that wrote it (lm2a_conv_desc.stats / lm2a_bias_add_bf16;
 * stats_pitch groups per clip-row): y = SiLU((x - mean) * rstd * gamma + beta),
 * bit-identical to the in_gn_* transform of lm2a_conv1d_bf16. c % 8 == 0,
 * (c/groups) % 8 == 0, groups <= 64.                                          */
int lm2a_gn_apply_bf16(void* stream, const void* x, int32_t x_ld, void* y,
                       int32_t y_ld, const void* stats, int32_t stats_pitch,
                       const float* gamma, const float* beta, int32_t rows,
                       int32_t tp, int32_t t_valid, int32_t c, int32_t groups,
                       float eps, int32_t apply_silu);

/* ---- cross-attention core (tcgen05 + TMEM) -------------------------------- */
/* q,o: bf16 slabs [R*tp, ld]; stream s, head h live at channel s*e + h*dh.
 * q is pre-scaled by log2(e)/sqrt(dh). k_s: bf16 [slots*lk, k_ld] (first e
 * channels used); vt_s: bf16 V^T [slots*e, vt_ld] (first lk keys of each row
 * used, vt_ld >= lk, multiple of 8). kv_slot[r] selects the cache slot of
 * clip-row r. dh = e/heads must be 32, 64, 96, 128, 192, 256 or 384.          */
int lm2a_cross_attn_bf16(void* stream, const void* q, int32_t q_ld, void* o,
                         int32_t o_ld, const void* k_motion,
                         const void* vt_motion, const void* k_text,
                         const void* vt_text, int32_t k_ld, int32_t vt_ld,
                         const int32_t* kv_slot, int32_t slots, int32_t rows,
                         int32_t tp, int32_t t_valid, int32_t lk, int32_t e,
                         int32_t heads);
/* Attention of per-head queries against the RAW condition sequence itself, for
 * blocks whose head dim equals the condition width (128): every head h of stream
 * s computes softmax(q'_h C_s^T) C_s with C_s bf16 [slots*lk, cond_ld] (first 128
 * channels), q' / o bf16 slabs with head h of stream s at channel (s*heads + h)*128.
 * The caller folds W_k,h into the query projection and W_v,h into the output
 * projection (K_h = C W_k,h^T + b_k: the bias shifts all scores of a row equally
 * and cancels in the softmax; V_h = C W_v,h^T + b_v: rows of softmax sum to 1), see
 * engine.py pack_block. Replaces the same reference lines as lm2a_cross_attn_bf16
 * (models/cross_attention.py:50-61) with 16x less key/value traffic. lk must satisfy
 * lm2a_cross_attn_cond_supported (the condition tile lives in shared memory).    */
int lm2a_cross_attn_cond_supported(int32_t lk);
int lm2a_cross_attn_cond_bf16(void* stream, const void* q, int32_t q_ld, void* o,
                              int32_t o_ld, const void* cond_motion,
                              const void* cond_text, int32_t cond_ld,
                              const int32_t* kv_slot, int32_t slots, int32_t rows,
                              int32_t tp, int32_t t_valid, int32_t lk, int32_t heads,
                              int32_t n_streams);
/* Same, for the first n_streams (1 or 2) condition streams only: n_streams = 1
 * runs the motion stream alone (q / o use their first e channels). Used when the
 * lyrics stream is constant in time — the reference's preprocessing tiles ONE
 * sentence embedding over all frames (preprocess.py:64-71), every key of that
 * stream is then identical, its softmax uniform and its attention output the
 * stream's single V row, which the caller keeps in the o slab.               */
int lm2a_cross_attn_streams_bf16(void* stream, const void* q, int32_t q_ld,
                                 void* o, int32_t o_ld, const void* k_motion,
                                 const void* vt_motion, const void* k_text,
                                 const void* vt_text, int32_t k_ld,
                                 int32_t vt_ld, const int32_t* kv_slot,
                                 int32_t slots, int32_t rows, int32_t tp,
                                 int32_t t_valid, int32_t lk, int32_t e,
                                 int32_t heads, int32_t n_streams);
/* The n_tail <= 8 query rows t0 .. t0 + n_tail - 1 of every (clip-row, stream, head)
 * on the CUDA cores: the rows that T mod 128 leaves over (4 / 2 / 1 at T = 516 / 258 /
 * 129). In the tensor-core kernels they cost a CTA slot (or, on the condition slab, a
 * serial round) of their own; the launch plan runs this kernel on a parallel graph
 * branch and gives the tensor-core launch t_valid = t0 (whole tiles only). q / o as
 * in lm2a_cross_attn_streams_bf16 (dh = e / heads in {32, 64, 128}); keys AND values
 * row-major: k_s / v_s bf16 [slots*lk, k_ld / v_ld], head h at channel h*dh
 * (shared_kv = 0: the K and V halves of the projection output) or every head at
 * channels [0, dh) (shared_kv = 1: the raw condition slab of lm2a_cross_attn_cond_bf16
 * for both, e = heads*128). Probabilities stay fp32 (no bf16 rounding of P); sums in a
 * fixed order (no atomics). Same reference lines: models/cross_attention.py:50-61.  */
int lm2a_cross_attn_tail_bf16(void* stream, const void* q, int32_t q_ld, void* o,
                              int32_t o_ld, const void* k_motion,
                              const void* v_motion, const void* k_text,
                              const void* v_text, int32_t k_ld, int32_t v_ld,
                              const int32_t* kv_slot, int32_t slots, int32_t rows,
                              int32_t tp, int32_t t0, int32_t n_tail, int32_t lk,
                              int32_t e, int32_t heads, int32_t n_streams,
                              int32_t shared_kv);
/* per-clip V cache transpose: src [slots*lk, src_ld] (c channels) ->
 * dst [slots*c, dst_ld] (lk keys); c multiple of 32.                        */
int lm2a_transpose_kv_bf16(void* stream, const void* src, int32_t src_ld,
                           void* dst, int32_t dst_ld, int32_t slots,
                           int32_t lk, int32_t c);

/* ---- timestep embedding + FiLM tables ----------------------------------- */
/* silu_temb[r,:] = SiLU(SiLU(W sinus(t[r]) + b))  (the SiLU that opens every
 * FiLM net is folded in); dim must be 256-thread friendly (<= 1024, even).  */
int lm2a_time_mlp(void* stream, const int64_t* t, const float* w,
                  const float* b, float* silu_temb, int32_t rows, int32_t dim);
/* out[r,:] = SiLU(W sinus(t[r]) + b), i.e. TimestepEmbedding.forward itself
 * (models/embedding.py:33-43), with fold_silu != 0 one more SiLU on top (what
 * lm2a_time_mlp computes). fold_silu == 0 feeds the legacy UNet1D's additive
 * time_proj (models/unet1d.py:23,48-49), which has no SiLU of its own.        */
int lm2a_time_embed(void* stream, const int64_t* t, const float* w,
                    const float* b, float* out, int32_t rows, int32_t dim,
                    int32_t fold_silu);
/* film[r, j] = sum_k silu_temb[r,k] * w[j,k] + b[j], j < cols (all FiLM nets
 * of the model concatenated).                                               */
int lm2a_film(void* stream, const float* silu_temb, const float* w,
              const float* b, float* film, int32_t rows, int32_t dim,
              int32_t cols);

/* ---- layout / resampling helpers ---------------------------------------- */
/* x fp32 [B, c, T] -> bf16 slab rows [copies*B, tp, ld] (channels >= c and
 * slots >= T zero); copy k of clip b lands in row k*B + b. As the first kernel
 * of a UNet step it also clears `zero_bytes` (multiple of 16, may be 0) at
 * `zero`: the GroupNorm statistics arena the step's epilogues accumulate into. */
int lm2a_ingest_x(void* stream, const float* x, void* slab, int32_t batch,
                  int32_t copies, int32_t c, int32_t t, int32_t tp, int32_t ld,
                  void* zero, int64_t zero_bytes);
/* fp32 [rows, t, c] -> bf16 slab [rows, tp, ld] (channels >= c zero)        */
int lm2a_ingest_seq(void* stream, const float* x, void* slab, int32_t rows,
                    int32_t t, int32_t c, int32_t tp, int32_t ld);
/* match_len(arr, t_out, mode='interp') of the reference (datasetcode/dataset.py:
 * 49-87, called at sample.py:124-125) for a padded batch: x fp32 [rows, t_in_max, c],
 * lens[r] (NULL = t_in_max) valid frames of row r. Per feature
 * np.interp(linspace(0, len-1, t_out), arange(len), x[:, d]) evaluated in fp64 with
 * numpy's operation order, cast to fp32: bit-identical to the host code. Writes
 * out_f32 [rows, t_out, c] and / or the bf16 slab [rows, tp, ld] (pad zeroed) the
 * CondProjection GEMM reads; either may be NULL.                              */
int lm2a_resample_seq(void* stream, const float* x, const int32_t* lens,
                      float* out_f32, void* out_slab, int32_t rows,
                      int32_t t_in_max, int32_t c, int32_t t_out, int32_t tp,
                      int32_t ld);
/* linear x2, align_corners=True: [R, tp_in, ld_in] (t_in valid) ->
 * [R, tp_out, ld_out] (2*t_in valid, rest zero)                             */
int lm2a_upsample2x_bf16(void* stream, const void* x, int32_t x_ld, void* y,
                         int32_t y_ld, int32_t rows, int32_t tp_in,
                         int32_t t_in, int32_t tp_out, int32_t c);

/* y[slot,:c] = x[slot,:c] + bias[:c] for slots with t < t_valid, zero otherwise
 * (identity-skip ResBlock of an all-zero-condition row: its attention output is
 * a constant vector — the classifier-free-guidance uncond shortcut). `stats`
 * (optional) accumulates the exact GroupNorm sums of y, layout as in
 * lm2a_conv_desc.                                                            */
int lm2a_bias_add_bf16(void* stream, const void* x, int32_t x_ld, void* y,
                       int32_t y_ld, const float* bias, int64_t slots,
                       int32_t tp, int32_t t_valid, int32_t c, void* stats,
                       int32_t stats_pitch, int32_t stats_cg, int32_t stats_c0);

/* ---- CFG blend + clamps + DDPM posterior update -------------------------- */
/* x [B,c,T] fp32 updated in place. eps: fp32 [2B,c,T] (uncond rows first)
 * when guided != 0, else [B,c,T]. sched: fp32 [steps,4] rows
 * {1/sqrt(alpha_t), beta_t/sqrt(1-abar_t), sqrt(beta_t), 0}; t_dev: int64[rows]
 * device timestep vector (entry b < B is clip b's t and indexes sched); noise may be NULL only if
 * every call has t == 0. If advance != 0 the kernel's last block decrements
 * all n_t entries of t_dev afterwards (graph replay); `ticket` is a zeroed
 * device uint32 used to elect that block. eps_out (optional) receives the
 * blended eps [B,c,T].                                                      */
int lm2a_cfg_posterior(void* stream, float* x, const float* eps,
                       const float* noise, const float* sched,
                       int64_t* t_dev, int32_t n_t, uint32_t* ticket,
                       int32_t batch, int64_t elems_per_clip, float guidance,
                       int32_t guided, int32_t advance, float* eps_out);

/* ---- the whole per-step update as one kernel -------------------------------- */
/* sample.py:162-210 between two UNet evaluations: CFG blend + clamps, DDPM
 * posterior update, noise injection, the next step's input slab (the
 * torch.cat([x, x]) + layout change lm2a_ingest_x performs) and the clearing of
 * the step's GroupNorm statistics arena. x fp32 [B, c, t] in place; eps / sched /
 * t_dev / ticket / guidance / guided / advance / eps_out as in lm2a_cfg_posterior.
 * Noise: `noise` fp32 [B, c, t] if given (injected draws, parity runs); else, if
 * clip_seed (uint64 [B]) is given, drawn in the kernel: z[b, ch, i] at timestep s =
 * Box-Muller of Philox4x32-10(counter = (ch * ceil(t/4) + i/4, s, 0, 0), key =
 * clip_seed[b]), component i % 4, u = (word >> 8) * 2^-24 + 2^-25 - a function of
 * the clip's seed, the element and the timestep only, so a clip's trajectory does
 * not depend on the batch (or GPU) it is sampled in; else no noise. slab (optional):
 * bf16 [copies*B, tp, ld] rows written from the NEW x (pad slots / channels zero).
 * zero / zero_bytes: region cleared by the same launch (may be NULL / 0).       */
int lm2a_cfg_step(void* stream, float* x, const float* eps, const float* noise,
                  const uint64_t* clip_seed, const float* sched, int64_t* t_dev,
                  int32_t n_t, uint32_t* ticket, int32_t batch, int32_t c,
                  int32_t t, float guidance, int32_t guided, int32_t advance,
                  void* slab, int32_t copies, int32_t tp, int32_t ld, void* zero,
                  int64_t zero_bytes, float* eps_out);
/* out fp32 [B, c, t] = the normals lm2a_cfg_step draws at counter word `step`
 * (a trajectory's x_T uses step = number of timesteps: one past the largest t;
 * reference sample.py:136 draws it with torch.randn).                          */
int lm2a_philox_normal(void* stream, float* out, const uint64_t* clip_seed,
                       int32_t batch, int32_t c, int32_t t, uint32_t step);

/* ---- CFG blend + DDIM update over a strided timestep sequence ------------- */
/* Reference models/diffusion.py:124-165 (ddim_sample; never called by the
 * reference's own loop). table: fp32 [steps, 8] rows {sqrt(1-abar_t), sqrt(abar_t),
 * sqrt(abar_prev), sqrt(1-abar_prev-sigma^2), sigma, (t_prev > 0), 0, 0} built by
 * the caller with the reference's expressions; step_idx: device int32 selecting
 * the row. x0 = clamp((x - eps*c0)/c1, -2, 2); x <- (c2*x0 + c3*eps) + sigma*z.
 * advance != 0: the last block sets every t_dev entry to t_seq[step+1] (int64
 * [steps+1], the UNet's next timestep) and increments *step_idx. x0_out
 * (optional) receives the clamped x0 prediction. eps / guidance as in
 * lm2a_cfg_posterior.                                                        */
int lm2a_cfg_ddim(void* stream, float* x, const float* eps, const float* noise,
                  const float* table, const int64_t* t_seq, int32_t* step_idx,
                  int64_t* t_dev, int32_t n_t, uint32_t* ticket, int32_t batch,
                  int64_t elems_per_clip, float guidance, int32_t guided,
                  int32_t advance, float* x0_out);

/* ---- mel evaluation metrics ------------------------------------------------ */
/* Reference val.py:25-113 (compute_metrics) for a batch: gen / real fp32
 * [B, n_mels, T]; gen is de-normalised on the fly (gen * gen_scale + gen_shift,
 * sample.py:230; pass 1, 0 for an already de-normalised mel). out: fp64 [B, 8]
 * {mse, ssim, avg_cos_sim, mean_error, std_error, snr, real_var, 0}; the host
 * rounds to 6 decimals as val.py:106-113 does. T >= 11 (SSIM window).          */
int lm2a_mel_metrics(void* stream, const float* gen, const float* real,
                     double* out, int32_t batch, int32_t n_mels, int32_t t,
                     float gen_scale, float gen_shift);

/* ---- fused Adan step (+ EMA shadow weights), all tensors in one launch ------ */
/* Reference models/adan.py:34-114 (restart_cond = None) and the EMA update of
 * train.py:177-180. `tensors`: DEVICE array of per-parameter pointers (ema may be
 * NULL); chunk_tensor / chunk_index: DEVICE int32 [n_chunks] mapping CTA b onto
 * elements [chunk_index[b] * chunk_elems, ...) of tensor chunk_tensor[b].
 * scalars: HOST float[14] = {1-b1, b1, 1-b2, b2, 1-b3, b3, correct_m, correct_v,
 * correct_n (for the step count AFTER the increment), lr, eps, 1 + wd * lr,
 * ema_decay, 1 - ema_decay}, each rounded to fp32 from the double the reference
 * computes. first_step != 0: state["step"] was 0 (moments are not updated).   */
typedef struct lm2a_adan_tensor {
  float* p;
  const float* g;
  float* prev_g;
  float* m;
  float* v;
  float* n;
  float* ema;
  int64_t numel;
} lm2a_adan_tensor;
int lm2a_adan_step(void* stream, const lm2a_adan_tensor* tensors,
                   const int32_t* chunk_tensor, const int32_t* chunk_index,
                   int32_t n_chunks, int32_t chunk_elems, int32_t first_step,
                   const float* scalars);

#ifdef __cplusplus
}
#endif
#endif /* LM2A_B200_H */
