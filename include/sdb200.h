/*
 * sdb200.h — C ABI of libsdb200.so, the sm_100a kernel library under the Stable Diffusion
 * sampling path (CLIP -> [VAE encoder] -> DDPM/UNet loop with CFG -> VAE decoder).
 *
 * The reference (dawmro/pytorch_stable_diffusion) has no FFI layer: every arithmetic call is a
 * torch.nn / torch.nn.functional call inside the sd/ modules. Each entry point below names the reference
 * call sites it replaces (paths relative to the reference checkout). Conventions:
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless marked "host";
 *   - activations are NHWC (tokens x channels) bf16 unless stated otherwise; weights are bf16,
 *     K-major ([out][in], 3x3 convs packed as [out][ky][kx][in]); biases and norm affine
 *     parameters are fp32;
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous, allocates nothing,
 *     never synchronises, and is legal under CUDA-graph stream capture;
 *   - return 0 on success, negative on error (SDB_ERR_*); sdb_last_error() gives the text.
 */
#ifndef SDB200_H_
#define SDB200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define SDB_ABI_VERSION 2   /* 2: IEEE-half operand / output flags, sdb_args_size, sdb_resample_u8, sdb_matmul_f64 */

#define SDB_OK 0
#define SDB_ERR_ARG (-1)
#define SDB_ERR_UNSUPPORTED (-2)
#define SDB_ERR_CUDA (-3)
#define SDB_ERR_DRIVER (-4)

/* ---- runtime ------------------------------------------------------------------------------ */
int sdb_abi_version(void);
const char* sdb_last_error(void);
/* Number of kernels this library has launched in this process (a launch recorded during CUDA-graph
 * capture counts once, at capture time). */
unsigned long long sdb_launch_count(void);
/* sizeof(sdb_gemm_args) (which = 0) / sizeof(sdb_attn_args) (which = 1) in this build of the library; a binding
 * checks it against its own struct layout before the first call. */
int sdb_args_size(int which);
/* Reads and clears the device watchdog word (non-zero = an mbarrier wait timed out inside a
 * kernel: (site << 8) | kind). Synchronises the device; for tests and smoke runs only. */
int sdb_read_fault(unsigned int* out_host);

/* ---- tensor-core GEMM / implicit-GEMM convolution ---------------------------------------- */
enum {
  SDB_GEMM_LINEAR = 0,            /* out[M,Cout] = A[M,C0(+C1)] . W^T   (nn.Linear, 1x1 conv)   */
  SDB_GEMM_CONV3X3_S1 = 1,        /* 3x3, stride 1, pad 1                                        */
  SDB_GEMM_CONV3X3_S2 = 2,        /* 3x3, stride 2, pad 1   (sd/diffusion.py:553,561,569)        */
  SDB_GEMM_CONV3X3_S2_PAD_RB = 3, /* 3x3, stride 2, pad right/bottom only (sd/encoder.py:120-122)*/
  SDB_GEMM_CONV2X2_UP = 4         /* one parity phase of [nearest x2 up-sampling -> 3x3 conv, pad 1] (Upsample,
                                     sd/diffusion.py:412-435): the output pixels (2y + a, 2x + b), up_phase = 2a + b, are a
                                     2x2 convolution of the LOW-resolution input a0 [NB, HI, WI, C0] with taps
                                     (a - 1 + u, b - 1 + v); w rows hold [u][v][C0] = the 3x3 taps that fall on the same
                                     input pixel summed at pack time. out / out2 are the [NB, 2HI, 2WI, Cout] tensors; four
                                     launches (phases 0..3) fill them: 4/9 of the FLOPs, no up-sampled tensor in HBM. */
};
enum { SDB_ACT_NONE = 0, SDB_ACT_QUICK_GELU = 1, SDB_ACT_SILU = 2 };

typedef struct sdb_gemm_args {
  int kind;               /* SDB_GEMM_*                                                          */
  const void* a0;         /* bf16 activations, source 0: [M, lda0] or NHWC [NB, HI, WI, C0]      */
  const void* a1;         /* optional source 1 (channel concatenation without a copy), or NULL   */
  const void* w;          /* bf16 weights [Cout, ldw], row = taps * (C0 + C1) values             */
  const float* bias;      /* fp32 [Cout] (or [M] when bias_per_row), or NULL                     */
  const void* residual;   /* [M, ldr] added after the activation, or NULL; type by res_fp32: 0 bf16, 1 fp32, 2 IEEE half */
  void* out;              /* bf16 (or fp32 when out_fp32) [M, ldo]                               */
  void* out2;             /* optional bf16 copy [M, ldo] written next to an fp32 `out`, or NULL  */
  float* workspace;       /* fp32 [nsplit, M, Cout] when nsplit > 1                              */
  int M;                  /* rows (LINEAR only)                                                  */
  int NB, HI, WI;         /* input batch / height / width (CONV only)                            */
  int C0, C1;             /* channels of source 0 / 1                                            */
  int Cout;
  long long lda0, lda1;   /* row strides in elements for LINEAR (0 = dense)                      */
  long long ldw;          /* weight row stride in elements (0 = dense)                           */
  long long ldo, ldr;     /* output / residual row strides in elements (0 = Cout)                */
  int out_fp32;
  int res_fp32;
  int bias_per_row;
  int act;                /* SDB_ACT_*                                                           */
  int block_n;            /* 0 = choose; else multiple of 16 in [16, 256]                        */
  int nsplit;             /* split-K factor, 0/1 = none                                          */
  int smem_budget;        /* bytes of shared memory for the pipeline, 0 = choose                 */
  int cta_pair;           /* 0 = choose; 1 = one CTA per tile (128 rows); 2 = CTA pairs (256 rows) */
  int out_f16;            /* the 16-bit tensor written - `out`, or `out2` next to an fp32 `out` - is IEEE half instead
                             of bf16 (saturating conversion: +-65504, never inf)                                  */
  int epi_mode;           /* epilogue: 0 = choose; 1 = per-lane global stores; 2 = TMA bulk stores whenever
                             the output geometry / alignment allows (default only for K <= 2048)          */
  float* gn_part;         /* optional: GroupNorm partial statistics of the fp32 OUTPUT, written by the epilogue:
                             fp32 [samples][K][Cout][2] = {sum, sum of squares} per 32-row slab and channel,
                             K = sdb_gemm_gn_slabs(...). Needs out_fp32, no split-K, Cout % 32 == 0. The
                             consumer reduces them with sdb_groupnorm_reduce_partials instead of reading the
                             tensor once more for its statistics (nn.GroupNorm after nn.Conv2d:
                             sd/diffusion.py:123-135,255). NULL = off.                                     */
  int gn_hw;              /* LINEAR with gn_part: rows per sample (multiple of 32, divides M)              */
  const void* ax0;        /* CONV3X3_S1 only: extra 1x1 source accumulated into the same output tile - the    */
  const void* ax1;        /* resblock's skip convolution (sd/diffusion.py:138-143,208): bf16 NHWC             */
  int Cx0, Cx1;           /* [NB, HI, WI, Cx0] (++ [.., Cx1]); w rows then hold 9*(C0+C1) + Cx0 + Cx1 values,  */
                          /* the 1x1 weights last; multiples of 64. NULL = off.                               */
  int up_phase;           /* CONV2X2_UP: 2a + b, or 4 = all four phases in one launch (w = [4][Cout][4 * C0], phase-major;
                             the tile index carries the phase). With gn_part the four phases share one partial-sum tensor
                             [samples][4 * K][Cout][2], K = sdb_gemm_gn_slabs(SDB_GEMM_CONV2X2_UP, NB, HI, WI, 0, 0)      */
  int ab_f16;             /* every 16-bit operand (a0, a1, ax0, ax1, w) is IEEE half instead of bf16: 11 instead of 8
                             significand bits at the same width and tensor-core rate (tcgen05 kind::f16 takes either),
                             fp32 accumulation unchanged. The UNet's full-resolution level runs this way: it carries
                             76 % of the squared bf16-rounding error of a UNet evaluation (tools/diag_layer_budget.py) */
} sdb_gemm_args;

/* Slabs per sample (K above) for a given problem, 0 = gn_part unsupported for this geometry. */
int sdb_gemm_gn_slabs(int kind, int NB, int HI, int WI, int M, int gn_hw);

/* Tile-chooser aid: bytes of activations one pipeline stage fetches per three k-blocks (one filter column) when a
 * stride-1 3x3 conv over [NB, HI, WI] runs with filter-column staging, 0 when it cannot (classic: 3 x 16 KiB). */
int sdb_gemm_conv_a3_bytes(int NB, int HI, int WI);

/* Replaces nn.Conv2d / nn.Linear: sd/diffusion.py:38,42,125,129,135,143,256,266,267,269,410,
 * 545-569,712; sd/attention.py:12,16,143-152; sd/decoder.py:112,121,129,235-339;
 * sd/encoder.py:56-92; sd/clip.py:117,121. */
int sdb_gemm_tc(const sdb_gemm_args* args, void* stream);

/* Debug aid: on >= 0 switches a per-tile SM-clock timeline of CTA 0 on (1) / off (0) and clears it;
 * on < 0 copies the 64 x 8 stamps to out_host (see csrc/gemm_tc.cu). Synchronises. */
int sdb_debug_gemm_trace(int on, long long* out_host);

/* ---- attention ----------------------------------------------------------------------------- */
typedef struct sdb_attn_args {
  const void* q;          /* bf16, token (n, s) head h at q + ((n*S + s) * ldq + h*d)            */
  const void* k;          /* bf16, key (n, t) head h at k + ((n*Skv_pad + t) * ldk + h*d)        */
  const void* vt;         /* bf16 V transposed: channel c of key (n, t) at vt + (c*NB + n)*vt_ld + t */
  void* out;              /* bf16 [NB*S, ldo], head h at column h*d                              */
  int NB, heads, d;       /* d = head dim (multiple of 8, <= 160)                                */
  int S;                  /* queries per sample                                                  */
  int Skv;                /* valid keys per sample                                               */
  int Skv_pad;            /* allocated key rows per sample in k (>= Skv)                         */
  int vt_ld;              /* per-sample stride of vt in elements (multiple of 8; 0 = Skv_pad)    */
  long long ldq, ldk, ldo;
  int causal;             /* key t visible to query s iff t <= s (sd/attention.py:58-62)         */
  float scale;            /* 1/sqrt(d) (sd/attention.py:66,223)                                  */
  int variant;            /* 0 = choose; 1 = force the one-tile (128 queries per CTA) kernel       */
  int sum_row;            /* vt holds R = round16(d + 1) rows per head (head h at rows [h*R, h*R + R)), row d of
                             every head is all ones: the denominator of the softmax is accumulated by the
                             P.V tensor-core product itself. No causal mask, d <= 112.                */
  int p_f16;              /* with sum_row: exponentials are taken two at a time in f16x2 and P is f16 */
  int exp_poly;           /* two-tile kernel: every fourth exponential is a degree-3 polynomial on the FMA pipe
                             (rel. error 7.5e-5, below the bf16 rounding of P) instead of a MUFU operation -
                             the kernel is bound by the 16 ex2 / clk / SM. 0 = default (on), 1 = off, 2 = on */
  int q_prescaled;        /* q already carries log2(e) * scale (folded into the query projection): the scores are
                             in log2 units and `scale` is ignored                                           */
  int qk_cols;            /* columns stored per head in q and k (head h at column h * qk_cols; all of them enter the
                             reduction); 0 = d                                                                */
  int qk_fold;            /* with sum_row + q_prescaled, qk_cols > d: column d of k is all ones, column d of q is
                             zero - the kernel writes -round(row maximum of the first key block) there (in its
                             shared-memory copy), so later score blocks arrive relative to the row's reference and
                             the softmax needs no subtraction                                                  */
} sdb_attn_args;

/* Flash-style softmax(Q K^T * scale) V with S in TMEM; replaces sd/attention.py:55-76 (self),
 * :219-234 (cross). */
int sdb_attention(const sdb_attn_args* args, void* stream);

/* ---- normalisation (HBM-bound) ------------------------------------------------------------ */
/* GroupNorm statistics over NHWC x0 (C0 channels) ++ x1 (C1 channels, may be NULL/0), each bf16 (x?_fp32 = 0),
 * fp32 (1) or IEEE half (2). Deterministic: every thread block writes its per-group partial {sum, sum of squares}
 * (fp64) into `stats`, a caller-provided buffer of sdb_groupnorm_stats_bytes(NB, groups) bytes that
 * needs no initialisation; sdb_groupnorm_apply (same NB, HW, C0, C1, groups) adds the partials in a
 * fixed order. nn.GroupNorm: sd/diffusion.py:123,133,255,708; sd/decoder.py:107,116,330;
 * sd/encoder.py:86. Wherever a normalisation / conversion entry point below takes `out_f16` (or out kind 2), the
 * 16-bit output is IEEE half instead of bf16 (the operand type of sdb_gemm_args::ab_f16 consumers). */
long long sdb_groupnorm_stats_bytes(int NB, int groups);
int sdb_groupnorm_stats(const void* x0, const void* x1, double* stats, int NB, long long HW,
                        int C0, int C1, int groups, int x0_fp32, int x1_fp32, void* stream);
/* y = (x - mean) * rstd * gamma + beta, optionally followed by SiLU (F.silu: sd/diffusion.py:176,
 * 202,738; sd/decoder.py:162,175,335); writes bf16 NHWC [NB, HW, C0 + C1]. */
int sdb_groupnorm_apply(const void* x0, const void* x1, const double* stats, const float* gamma,
                        const float* beta, void* out, int NB, long long HW, int C0, int C1,
                        int groups, float eps, int silu, int x0_fp32, int x1_fp32, int stat_chunks,
                        int out_f16, void* stream);
/* Statistics without a pass over the tensor: reduces the partial sums the producing GEMM epilogues wrote
 * (sdb_gemm_args::gn_part; part0 [NB][K0][C0][2], part1 [NB][K1][C1][2] for a channel concat or NULL) into
 * `stats` as ONE chunk - follow with sdb_groupnorm_apply(..., stat_chunks = 1). stat_chunks = 0 above means
 * "what sdb_groupnorm_stats wrote". */
int sdb_groupnorm_reduce_partials(const float* part0, const float* part1, double* stats, int NB, int K0, int K1,
                                  int C0, int C1, int groups, void* stream);
/* One-pass GroupNorm (+SiLU) for fp32 NHWC inputs: a thread-block cluster keeps (sample, slab of groups) resident
 * in shared memory, so the tensor is read from HBM once and no statistics buffer exists. Same arithmetic
 * contract as stats + apply (fp64 combination in a fixed order). sdb_groupnorm_fused_supported(): 0 = no plan
 * (use stats + apply), 1 = plan with a multi-CTA cluster (correct, but slower than stats + apply whose second
 * read hits L2), 2 = single-CTA plan (faster; the small UNet levels). */
int sdb_groupnorm_fused_supported(long long HW, int C0, int C1, int groups);
int sdb_groupnorm_fused(const float* x0, const float* x1, const float* gamma, const float* beta, void* out,
                        int NB, long long HW, int C0, int C1, int groups, float eps, int silu, int out_f16,
                        void* stream);
/* nn.LayerNorm over the last axis (sd/diffusion.py:258,261,264; sd/clip.py:105,113,225).
 * x bf16 (fp32 when in_fp32) [rows, C] -> out bf16 (out_kind 0), fp32 (1) or IEEE half (2). */
int sdb_layernorm(const void* x, const float* gamma, const float* beta, void* out, long long rows,
                  int C, float eps, int in_fp32, int out_kind, void* stream);
/* Row softmax of fp32 scores * scale -> bf16 probabilities (VAE attention, sd/attention.py:66-71). */
int sdb_softmax_rows(const float* scores, void* probs, long long rows, int cols, float scale,
                     void* stream);

/* ---- layout / elementwise ------------------------------------------------------------------ */
int sdb_fill_zero(void* ptr, long long bytes, void* stream);
/* fp32 NCHW [NB, C, H, W] -> NHWC [NB*repeat, H, W, C] in bf16 (fp32 when out_fp32), value * scale;
 * `repeat` tiles the batch (latents.repeat(2,1,1,1), sd/pipeline.py:221). */
int sdb_nchw_f32_to_nhwc(const float* x, void* out, int NB, int C, int H, int W, int repeat,
                         float scale, int out_fp32, void* stream);
/* NHWC (bf16, or fp32 when in_fp32) -> fp32 NCHW. */
int sdb_nhwc_to_nchw_f32(const void* x, float* out, int NB, int C, int H, int W, int in_fp32,
                         void* stream);
/* Nearest-neighbour x2 (F.interpolate sd/diffusion.py:430; nn.Upsample sd/decoder.py:269). */
int sdb_upsample2x_nhwc(const void* x, void* out, int NB, int H, int W, int C, void* stream);
/* Direct convolution for tiny channel counts (Cin <= 8 or Cout <= 8): k in {1, 3}, stride 1,
 * pad (k-1)/2. x NHWC bf16 (fp32 when in_fp32), w fp32 [Cout][k*k][Cin], out bf16 or fp32 NHWC; out2 =
 * optional bf16 copy next to an fp32 out (or NULL). sd/diffusion.py:545; sd/decoder.py:235,239;
 * sd/encoder.py:56,92. */
int sdb_conv_direct(const void* x, const float* w, const float* bias, void* out, void* out2, int NB,
                    int H, int W, int Cin, int Cout, int ksize, int out_fp32, int in_fp32, int out2_f16,
                    void* stream);
/* y[r, n] = act_out( sum_k act_in(x[r, k]) * W[n, k] + bias[n] ), fp32 activations, bf16 weights.
 * The time path: TimeEmbedding (sd/diffusion.py:64-76) and SiLU+linear_time (:184-187). */
int sdb_small_linear(const float* x, const void* w, const float* bias, float* out, int R, int K,
                     int N, int act_in, int act_out, void* stream);
/* Fused classifier-free guidance + DDPM ancestral step (sd/pipeline.py:228-237,
 * sd/ddpm.py:102-139). eps: fp32 NHWC (NCHW when eps_nchw) [2*NB (or NB when !do_cfg), H, W, C]; latents/noise: fp32
 * NCHW [NB, C, H, W]; coef: device fp32 [steps][5] = {sqrt(1-abar_t), sqrt(abar_t), c_x0, c_xt,
 * sigma_t}; writes latents in place and the next UNet input (bf16 NHWC, batch tiled x2 when
 * do_cfg; fp32 when next_fp32) to next_in when non-NULL. */
int sdb_cfg_ddpm_step(float* latents, const float* eps, const float* noise, const float* coef,
                      int step, float cfg_scale, int do_cfg, void* next_in, int NB, int C, int H,
                      int W, int eps_nchw, int next_fp32, void* stream);
/* VAE_AttentionBlock tail as the reference computes it (sd/decoder.py:62-71): the (n, hw, c)
 * attention output is re-viewed raw as (n, c, h, w) and added to the residual.
 * y bf16, res fp32, out fp32 (+ optional bf16 copy out2), all NHWC [NB, HW, C].
 * out[n, p, c] = y_flat[n][c*HW + p] + res[n, p, c]. */
int sdb_vae_attn_scramble_add(const void* y, const float* res, float* out, void* out2, int NB,
                              long long HW, int C, void* stream);
/* fp32 -> 16-bit copy (bf16, or IEEE half when f16: the 16-bit shadow of an fp32 residual-stream tensor). */
int sdb_f32_to_bf16(const float* x, void* out, long long n, int f16, void* stream);
/* VAE encoder tail (sd/encoder.py:127-152): moments fp32 NHWC [NB, H, W, 8] + noise fp32 NCHW
 * [NB, 4, H, W] -> latents fp32 NCHW: (mean + exp(clamp(logvar,-30,20))^0.5 * noise) * 0.18215. */
int sdb_vae_encode_tail(const float* moments, const float* noise, float* out, int NB, int H, int W,
                        void* stream);
/* DDPMSampler.add_noise (sd/ddpm.py:143-186): out = sa * x + sb * noise (fp32, any shape). */
int sdb_axpby(const float* x, const float* y, float* out, float a, float b, long long n,
              void* stream);
/* Post-processing (sd/pipeline.py:253-259): fp32 NHWC in [-1,1] -> uint8 NHWC, rescale to
 * [0,255], clamp, truncating cast. */
int sdb_image_to_uint8(const float* x, unsigned char* out, long long n, void* stream);
/* Pre-processing (sd/pipeline.py:162-173): uint8 HWC -> NHWC in [-1,1], bf16 or (out_fp32) fp32. */
int sdb_uint8_to_image(const unsigned char* x, void* out, long long n, int out_fp32, void* stream);
/* One pass of Pillow's 8-bit separable resampler over uint8 NHWC [NB, H, W, C] along W (axis 0) or H (axis 1) -> out_size
 * samples: PIL.Image.resize as the reference calls it on the input image (sd/pipeline.py:156; bicubic with antialiasing).
 * bounds int32 [out_size][2] = {first source index, tap count}, coef int32 [out_size][ksize] = taps with 22 fractional
 * bits (host-computed, pytorch_stable_diffusion_b200/imageio.py). img (optional fp32, same shape as dst) receives
 * dst * (2/255) - 1, the reference's rescale (sd/pipeline.py:162-173). Byte-exact against Pillow. */
int sdb_resample_u8(const unsigned char* src, unsigned char* dst, float* img, int NB, int H, int W, int C, int out_size,
                    int axis, const int* bounds, const int* coef, int ksize, void* stream);
/* CLIPEmbedding (sd/clip.py:58-63): out[b, t, :] = table[tokens[b, t]] + pos[t]; rows
 * t >= T (up to T_pad) are zero. tokens int64 [NB, T]; table/pos fp32; out fp32 [NB, T_pad, D]. */
int sdb_clip_embed(const long long* tokens, const float* table, const float* pos, void* out, int NB,
                   int T, int T_pad, int D, int vocab, void* stream);

/* Stream-ordered device-to-device copy (a memcpy node under graph capture). Duplicates the part of a UNet evaluation
 * that the two members of a classifier-free-guidance pair share - everything before the first cross-attention, since
 * both see the same latent and time step (sd/pipeline.py:221: latents.repeat(2, 1, 1, 1)). */
int sdb_copy_bytes(void* dst, const void* src, long long bytes, void* stream);

/* Pack-time helper (model load, not the sampling loop): C[M, N] = A[M, K] . B[K, N], row-major, A / B fp32 (or fp64
 * when a_f64 / b_f64), C fp64, accumulated in fp64 on the CUDA cores. Composes the reference's feed-forward
 * linear_geglu_2 . linear_geglu_1[:4C] (no non-linearity in between: the GEGLU gate is dead, sd/diffusion.py:359-363)
 * and conv_output (:371-381) into one matrix per attention block. */
int sdb_matmul_f64(const void* A, int a_f64, const void* B, int b_f64, double* C, int M, int N, int K, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SDB200_H_ */
