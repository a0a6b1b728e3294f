/*
 * v3d_b200 — C ABI of the B200 (sm_100a) kernels behind V3D's denoising hot path.
 *
 * The reference (heheyas/V3D) has no FFI of its own: every GPU op on the path is a PyTorch library
 * call.  Each entry point below therefore names the reference call site(s) it replaces
 * (file:line under the reference tree).  Conventions, identical for every function:
 *   - plain pointers to DEVICE memory + sizes; no allocation inside; caller owns all buffers;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - activations are bf16 NHWC / token-major ([rows, channels], channels contiguous) unless said
 *     otherwise; norm statistics, biases, sampler state are fp32;
 *   - return value 0 = ok; non-zero = error code (V3D_ERR_*), message via v3d_last_error().
 *   - launches are asynchronous on `stream`; nothing here synchronises the device.
 */
#ifndef V3D_B200_H
#define V3D_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define V3D_ABI_VERSION 1

#define V3D_ERR_OK 0
#define V3D_ERR_BAD_ARG 1
#define V3D_ERR_CUDA 2
#define V3D_ERR_NO_DRIVER 3
#define V3D_ERR_UNSUPPORTED 4

#define V3D_ACT_NONE 0
#define V3D_ACT_SILU 1
#define V3D_ACT_GEGLU 2 /* out[j] = (v[j]) * gelu_erf(g[j]); weight rows packed per N-tile, see v3d_geglu_pack_rows */

int v3d_abi_version(void);
const char* v3d_last_error(void);
/* number of kernels launched through this library since load (bench.py's gpu_launches) */
int64_t v3d_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Tensor-core GEMM family (tcgen05.mma, TMEM accumulators, TMA-staged operands).
 *
 *   D[row, n] = s0 * act( sum_k A[row, k] * B[n, k] + bias[n] + fbias[row / rows_per_frame, n] )
 *               + s1 * R1[row, n] + s2 * R2[row, n]
 *
 * One kernel, three operand-gather modes:
 *   linear   (conv_w == 0, ntaps == 1): A is [batch][rows_per_batch][K], row stride lda.
 *            replaces nn.Linear / 1x1 Conv2d: sgm/modules/attention.py:95,277-283,680,704;
 *            sgm/modules/diffusionmodules/openaimodel.py:324; model.py:167-178 (AttnBlock q,k,v,proj).
 *   temporal (conv_w == 0, ntaps == 3): K-loop runs over 3 taps; tap t reads rows shifted by
 *            (t-1)*tap_shift inside the batch item, zero outside -> Conv3d k=(3,1,1) pad (1,0,0) on a
 *            frame-major [b][t*hw][C] tensor.  replaces sgm/modules/diffusionmodules/video_model.py:42-55
 *            (time_stack ResBlock convs) and sgm/modules/autoencoding/temporal_ae.py:31-44,94-99.
 *   conv3x3  (conv_w > 0): A is an NHWC image batch [conv_n][conv_h][conv_w][K] (pixel stride lda); the
 *            K-loop runs over the 9 taps with implicit zero padding (TMA out-of-bounds fill).
 *            replaces nn.Conv2d 3x3 s1 p1: openaimodel.py:145-147,267-315; video_model.py:189,439;
 *            model.py:65-66,111-120,651-653.
 * B is always [N][ntaps*K] (K-major, tap-major then channel), optionally batched.
 * All of K, lda, ldb, ldd, ldr* are in elements; K % 64 == 0; N % 16 == 0.
 * ------------------------------------------------------------------------------------------ */
typedef struct v3d_gemm_args {
  const void* A;      /* bf16 */
  const void* B;      /* bf16 [b_batch][N][ntaps*K] */
  void* D;            /* bf16, or fp32 when out_fp32 */
  const float* bias;  /* [N] or NULL */
  const float* fbias; /* [nframes][N] or NULL: per-frame (per-sample) additive row vector */
  const void* R1;     /* bf16 [rows][ldr1] or NULL */
  const void* R2;     /* bf16 [rows][ldr2] or NULL */
  int64_t lda, ldb, ldd, ldr1, ldr2;
  int64_t ldfb; /* row stride of fbias in floats; 0 -> N */
  int64_t a_batch_stride, b_batch_stride; /* elements; b_batch_stride == 0 -> B shared by all batch items */
  int32_t batch, rows_per_batch;
  int32_t N, K;
  int32_t ntaps, tap_shift;
  int32_t rows_per_frame; /* rows sharing one fbias vector (>=1) */
  int32_t act;            /* V3D_ACT_* */
  int32_t out_fp32;
  int32_t conv_n, conv_h, conv_w; /* conv3x3 mode when conv_w > 0 (then batch/rows_per_batch ignored) */
  int32_t block_n;                /* 0 = auto; else force the N tile (16,32,64,128,160,256) */
  /* small-M mode (the M<=64 embedding linears run as W[N_w,K] x X[64,K]^T): D is fp32 [valid_cols][ldd] written
   * transposed (D[col][row] = acc + bias[row]), optionally accumulated into; columns >= valid_cols are dropped */
  int32_t out_transposed, valid_cols, accumulate;
  float s0, s1, s2;
  /* halo'd A operand (linear / temporal modes; both 0 = none): every batch item of A holds a_rows >= a_row0 +
   * rows_per_batch rows and output row r reads A row a_row0 + r (+ the tap shifts); rows outside [0, a_rows) read as
   * zero.  Frame-sharded temporal convs pass a buffer with one halo frame on each side (a_row0 = tap_shift,
   * a_rows = rows_per_batch + 2 * tap_shift): the neighbours' boundary frames of video_model.py:42-55 /
   * temporal_ae.py:94-99 when the T view-frames are split across GPUs. */
  int32_t a_rows, a_row0;
  /* fused GEMM -> all-gather over peer memory (plain linear GEMMs with bf16 output; kv_n = 0: off).  Output columns
   * >= kv_col0 are ALSO stored - by the same epilogue, with the same TMA sub-tile stores - into kv_n destination matrices
   * kv_dst[i] (row stride kv_ld elements, column c - kv_col0, row = the output row), which may be IPC-mapped memory of
   * other GPUs (v3d_peer_import): the frame-sharded temporal attention's K|V all-gather (video_attention.py:114-125
   * around attention.py:337-341) leaves the projection GEMM tile by tile over NVLink instead of through a packing copy
   * and a separate collective.  kv_col0 must be a multiple of 32. */
  int32_t kv_col0, kv_n;
  int64_t kv_ld;
  void* kv_dst[8];
} v3d_gemm_args;

int v3d_gemm_bf16(const v3d_gemm_args* args, void* stream);
/* sizeof(v3d_gemm_args) as this library was compiled: bindings check their mirror of the struct against it. */
int v3d_gemm_args_size(void);

/* Permutation used to pack GEGLU projection rows so that each N-tile of width block_n holds
 * block_n/2 "value" rows followed by the matching block_n/2 "gate" rows
 * (reference split: attention.py:97-99, value = first half, gate = second half).
 * perm[i] = source row of packed row i; n_out = inner dim (proj has 2*n_out rows). Host function. */
int v3d_geglu_pack_rows(int32_t n_out, int32_t block_n, int32_t* perm);
/* N-tile width the GEMM will choose for a given N (so packers and callers agree). Host function. */
int v3d_gemm_pick_block_n(int32_t N, int32_t act);
/* diagnostics only: device buffer (>= 3072 int64) that CTA 0 of subsequent v3d_gemm_bf16 launches fills with per-role
 * clock64() timelines; NULL disables. */
int v3d_debug_set_trace(void* buf);

/* ------------------------------------------------------------------------------------------
 * Normalisation (norm.cu). GroupNorm is split into a statistics pass and an apply(+SiLU) pass; both
 * work on NHWC bf16 with fp32/fp64 statistics, like GroupNorm32 (diffusionmodules/util.py:274-276).
 * ------------------------------------------------------------------------------------------ */
/* stats: double [nsamples][groups][2] = {sum, sumsq} over rows_per_sample x (C/groups). rows_per_sample is
 * H*W for 2-D norms (util.py:259-276; attention.py:130-133; model.py:52-55) and T*H*W for the 3-D
 * time_stack ResBlock norms (openaimodel.py:267-271 with dims=3). The kernel accumulates with atomics: stats is
 * cleared first unless pre_zeroed != 0 (callers that carve many stats slices out of one zeroed pool). */
int v3d_groupnorm_stats(const void* x, void* stats, int64_t rows_per_sample, int32_t nsamples, int32_t C,
                        int32_t ldx, int32_t groups, int32_t pre_zeroed, void* stream);
/* y (dense, ld = C) = act(GN(x)); silu != 0 fuses the nn.SiLU / nonlinearity that always follows
 * (openaimodel.py:267-271,300-303; model.py:131-143). */
int v3d_groupnorm_apply(const void* x, void* y, const void* stats, const void* gamma, const void* beta,
                        int64_t rows_per_sample, int32_t nsamples, int32_t C, int32_t ldx, int32_t groups,
                        float eps, int32_t silu, void* stream);
/* The same normalisation in ONE launch (statistics -> grid-wide barrier -> apply): what the unsharded path calls;
 * the pair above remains for the frame-sharded path, where the statistics are all-reduced between the two passes.
 * Deterministic (per-CTA fp32 partials folded by warp-shuffle trees, summed in fp64 in a fixed order; no atomics on
 * data).  `workspace`: >= v3d_groupnorm_workspace_bytes() bytes of device memory, zeroed once by the caller and then
 * passed unchanged to every call issued on the same stream (it holds the barrier state and the partials). */
int64_t v3d_groupnorm_workspace_bytes(void);
int v3d_groupnorm(const void* x, void* y, const void* gamma, const void* beta, int64_t rows_per_sample,
                  int32_t nsamples, int32_t C, int32_t ldx, int32_t groups, float eps, int32_t silu,
                  void* workspace, int64_t workspace_bytes, void* stream);
/* LayerNorm over C (attention.py:525-527; video_attention.py:51,79,93-94). Optional fused
 * z = x + add[row / rows_per_frame] (fp32 vectors; video_attention.py:286-287), z stored to ysum. */
int v3d_layernorm(const void* x, const void* add, void* ysum, void* y, const void* gamma, const void* beta,
                  int64_t rows, int32_t C, int32_t rows_per_frame, float eps, void* stream);
/* in-place softmax(scale * x) over rows of n bf16 scores (decoder AttnBlock, model.py:190-192). */
int v3d_softmax_rows(void* x, int64_t rows, int32_t n, float scale, void* stream);
/* same, fp32 scores in, bf16 probabilities out (separate buffers). */
int v3d_softmax_rows_f32(const void* x, void* y, int64_t rows, int32_t n, float scale, void* stream);

/* ------------------------------------------------------------------------------------------
 * Attention (attention.cu), head dim 64. q/k/v are column slices of one packed projection output
 * (row stride ld_qkv elements); heads are 64-column groups.
 * ------------------------------------------------------------------------------------------ */
/* F.scaled_dot_product_attention / xformers FMHA, attention.py:337-341,432-444: per (sample, head)
 * softmax(q k^T scale) v over ntok tokens; token rows are sample-major. tcgen05 MMAs (S = QK^T and O = PV with
 * TMEM accumulators), TMA-staged Q/K/V tiles, warp-specialised fp32 softmax. */
int v3d_attention_spatial(const void* q, const void* k, const void* v, void* o, int64_t ld_qkv, int64_t ld_o,
                          int32_t nbatch, int32_t ntok, int32_t nheads, float scale, void* stream);
/* validation twin of v3d_attention_spatial on mma.sync (same contract); not used by the product path. */
int v3d_attention_spatial_mma(const void* q, const void* k, const void* v, void* o, int64_t ld_qkv, int64_t ld_o,
                              int32_t nbatch, int32_t ntok, int32_t nheads, float scale, void* stream);
/* the same attention across the T view-frames of each pixel (video_attention.py:114,125), reading the
 * frame-major token matrix in place: row(b,t,s) = (b*T + t)*S + s, T <= 32. */
int v3d_attention_temporal(const void* q, const void* k, const void* v, void* o, int64_t ld_qkv, int64_t ld_o,
                           int32_t nb, int32_t T, int32_t S, int32_t nheads, float scale, void* stream);
/* frame-sharded form (the "temporal-attention KV all-gather" of BASELINE.json): q / o hold this rank's
 * Tq frames (row(b,t,s) = (b*Tq + t)*S + s); k / v point into the all-gathered K|V buffer (row stride ld_kv) where key
 * frame f of CFG half b, pixel s is row kv_row[f] + b*kv_bstride[f] + s (host int32[Tk] arrays, rank-major layout).
 * Tq <= Tk <= 32. */
int v3d_attention_temporal_kv(const void* q, const void* k, const void* v, void* o, int64_t ld_q, int64_t ld_kv,
                              int64_t ld_o, int32_t nb, int32_t Tq, int32_t Tk, int32_t S, int32_t nheads,
                              const int32_t* kv_row, const int32_t* kv_bstride, float scale, void* stream);

/* ------------------------------------------------------------------------------------------
 * Data movement / small matrices (elementwise.cu)
 * ------------------------------------------------------------------------------------------ */
/* F.interpolate(scale_factor=2, mode="nearest") on NHWC bf16 (openaimodel.py:164; model.py:68). */
int v3d_upsample_nearest2x(const void* x, void* y, int32_t N, int32_t H, int32_t W, int32_t C, void* stream);
/* [rows][ncols] bf16 block copy between row-strided buffers (th.cat skip concat, video_model.py:483). */
int v3d_copy_channels(const void* src, int64_t ld_src, void* dst, int64_t ld_dst, int64_t rows, int32_t ncols,
                      void* stream);
/* explicit im2row for the 3x3 convs the TMA gather does not cover (Cin % 64 != 0: video_model.py:189,
 * model.py:651; stride 2: openaimodel.py:202-209, model.py:82-90). y[n,oh,ow][tap*C + c], zero-padded to Kpad. */
int v3d_im2col3x3(const void* x, void* y, int32_t N, int32_t H, int32_t W, int32_t C, int32_t stride, int32_t pad,
                  int32_t Hout, int32_t Wout, int32_t Kpad, void* stream);
/* sgm boundary layout conversion: NCHW fp32 (wrappers.py:27; video_diffusion.py:184) <-> NHWC. */
int v3d_nchw_f32_to_nhwc_bf16(const void* x, void* y, int32_t N, int32_t C, int32_t H, int32_t W, float scale,
                              void* stream);
int v3d_nhwc_to_nchw_f32(const void* x, void* y, int32_t N, int32_t C, int32_t HW, int64_t ldx, int32_t src_fp32,
                         float scale, void* stream);
/* y[M,N] fp32 (+)= act_out(act_in(x[M,K] fp32) W[N,K]^T (bf16) + bias), M <= 64, row strides ldx/ldy
 * in floats (0 -> dense): time_embed/label_emb
 * (video_model.py:151-182,456-461), emb_layers (openaimodel.py:291-297), time_pos_embed
 * (video_attention.py:220-224,275), single-token cross-attention to_out(to_v(ctx)) (attention.py:277-283). */
int v3d_small_linear(const void* x, const void* W, const void* bias, void* y, int32_t M, int32_t K, int32_t N,
                     int32_t act_in, int32_t act_out, int32_t accumulate, int64_t ldx, int64_t ldy, void* stream);
/* operand prep for the tensor-core small-M path: y bf16 [64][K] = act_in(x fp32 [M][ldx]), rows >= M zero. */
int v3d_prep_small_x(const void* x, int64_t ldx, void* y, int32_t M, int32_t K, int32_t act_in, void* stream);
/* timestep_embedding (diffusionmodules/util.py:207-231): out[n][dim] = cos | sin, fp32. */
int v3d_timestep_embedding(const void* t, void* out, int32_t n, int32_t dim, float max_period, void* stream);
int v3d_add_rows(const void* a, const void* b, void* out, int32_t rows, int32_t cols, void* stream);
/* AE3DConv.time_mix_conv (temporal_ae.py:94-107): Conv3d(C,C,(3,1,1)) over frames, C <= 4; x NHWC fp32 (row
 * stride ldx), w [C][C][3] fp32, y NCHW fp32 [nb*T][C][HW]. */
int v3d_time_mix_conv(const void* x, int64_t ldx, const void* w, const void* bias, void* y, int32_t nb, int32_t T,
                      int64_t HW, int32_t C, void* stream);

/* ------------------------------------------------------------------------------------------
 * EDM sampler arithmetic, fp32 (sampler.cu)
 * ------------------------------------------------------------------------------------------ */
/* y = x * c_in(sigma), c_noise = 0.25 log sigma (denoiser.py:31-39; denoiser_scaling.py:51-59). */
int v3d_edm_scale_input(const void* x, const void* sigma, void* y, void* c_noise, int32_t nsamples,
                        int64_t per_sample, void* stream);
/* out = net * c_out + x * c_skip (denoiser.py:36-39). */
int v3d_edm_denoise_combine(const void* net, const void* x, const void* sigma, void* out, int32_t nsamples,
                            int64_t per_sample, void* stream);
/* LinearPredictionGuider.__call__ (guiders.py:78-86), batch order [uc; c]. */
int v3d_cfg_combine(const void* den, const void* scale, void* out, int32_t B, int32_t T, int64_t per_sample,
                    void* stream);
/* to_d + euler_step (sampling_utils.py:34-35; sampling.py:81-82,103-106); out may alias x. */
int v3d_euler_step(const void* x, const void* den, const void* sigma_hat, const void* sigma_next, void* out,
                   int32_t nsamples, int64_t per_sample, void* stream);
/* HeunEDMSampler.possible_correction_step (sampling.py:221-237): x + dt * (d + d_new) / 2 where sigma_next > 0,
 * else the Euler proposal; x_euler = euler step of x, den2 = denoised(x_euler, sigma_next). out may alias x. */
int v3d_heun_step(const void* x, const void* den, const void* x_euler, const void* den2, const void* sigma_hat,
                  const void* sigma_next, void* out, int32_t nsamples, int64_t per_sample, void* stream);
/* clamp((x+1)/2,0,1)*255 -> uint8 THWC from the decoder's NHWC output (scripts/pub/V3D_512.py:286-303). */
int v3d_decode_to_u8(const void* x, int64_t ldx, int32_t src_fp32, void* y, int64_t npix, void* stream);
/* same, from NCHW fp32 frames [T][3][HW] (decode_first_stage's return layout) to uint8 [T][HW][3]. */
int v3d_frames_nchw_to_u8(const void* x, void* y, int32_t T, int64_t HW, void* stream);

/* ------------------------------------------------------------------------------------------
 * One-sided exchanges over NVLink peer memory (peer.cu): the transport of the frame-sharded path - ONE image over
 * several GPUs of a box (BASELINE.json north_star).  They replace, on the data path, the three
 * torch.distributed exchanges that the reference's single-GPU modules imply once the T frames are split over ranks:
 * the temporal-attention K|V all-gather (video_attention.py:114-125 around attention.py:337-341), the 1-frame halos
 * of the (3,1,1) temporal convolutions (video_model.py:42-55, temporal_ae.py:94-99) and the (sum, sumsq) all-reduce
 * of the 3-D GroupNorm (openaimodel.py:267-271 with dims=3).  Every rank owns an arena (v3d_peer_alloc) that the
 * other ranks map through a 64-byte IPC handle; kernels then store straight into the mapped arenas and raise /
 * wait on epoch-valued flag words, so a whole sharded forward is stream-ordered and CUDA-graph capturable.
 * ------------------------------------------------------------------------------------------ */
#define V3D_PEER_MAX_SEG 16
#define V3D_PEER_MAX_FLAG 16
#define V3D_PEER_MAX_RANKS 8
/* zeroed device memory that can be exported (plain cudaMalloc, not the framework's caching allocator) */
int v3d_peer_alloc(int64_t bytes, void** out);
int v3d_peer_free(void* p);
/* 64-byte handle of an arena (cudaIpcGetMemHandle) / mapping of another rank's arena (cudaIpcOpenMemHandle) */
int v3d_peer_export(const void* p, void* handle64);
int v3d_peer_import(const void* handle64, void** out);
int v3d_peer_close(void* p);
/* *epoch += 1 on the stream: once at the start of every sharded forward (flags carry the epoch) */
int v3d_peer_epoch_bump(void* epoch, void* stream);
/* copy nseg contiguous segments src[i] -> dst[i] (bytes[i], multiples of 16; dst may be mapped peer memory), then -
 * once all stores are fenced system-wide - store *epoch into the nflag flag words (usually in the receivers' arenas).
 * done_counter: one zeroed u32 of local scratch per stream. */
int v3d_peer_put(int32_t nseg, const void* const* src, void* const* dst, const int64_t* bytes, int32_t nflag,
                 void* const* flags, const void* epoch, void* done_counter, void* stream);
/* block the stream until every flag word >= *epoch (bounded: after 30 s the site id is recorded in *status and every later wait
 * of the transport returns at once) */
int v3d_peer_wait(int32_t nflag, const void* const* flags, const void* epoch, void* status, int32_t site,
                  void* stream);
/* stats[n] (fp64) := scale * sum over ranks, rank-ordered (bit-identical on every rank): slot_base[r] / flag_base[r]
 * = rank r's slot area [world][n] doubles / flag area [world] words of this exchange site. */
int v3d_peer_allreduce_f64(void* stats, int32_t n, double scale, int32_t world, int32_t rank,
                           void* const* slot_base, void* const* flag_base, const void* epoch, void* status,
                           int32_t site, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* V3D_B200_H */
