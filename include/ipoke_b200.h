/*
 * ipoke_b200 -- C ABI of the B200-native iPOKE sampling hot path (libipoke_b200.so).
 *
 * Plain pointers and sizes only; no torch types.  Every entry point returns 0 on success or a negative
 * ipk_status; the message of the last failure on the calling thread is available from ipk_last_error().
 * All device pointers are CUDA device pointers on the current device; `stream` is a cudaStream_t passed
 * as void* (NULL = legacy default stream).  Calls are asynchronous on `stream` unless stated otherwise.
 *
 * Tensor layouts at this boundary are the reference's own (fp32, NCHW / N-T-C-H-W); the library keeps
 * NHWC / NDHWC and packed operand planes internally.
 *
 * Reference interfaces replaced (paths relative to the CompVis/ipoke checkout):
 *   ipk_flow_*   <- SupervisedMacowTransformer.forward/reverse     models/modules/INN/INN.py:469-481
 *                   (MultiScaleInternal.forward                    models/modules/INN/macow2.py:873-920)
 *   ipk_fs_*     <- PokeMotionModel.decode_first_stage             models/second_stage_video.py:361-382
 *                   (ConvGRU.forward                               models/modules/motion_models/rnn.py:104-133,
 *                    SpadeCondConvDecoder.forward                  models/modules/autoencoders/fully_conv_models.py:166-177)
 *   ipk_sample_* <- PokeMotionModel.forward_sample loop body       models/second_stage_video.py:333-341
 *   ipk_cenc_*   <- ConvEncoder.forward (poke embedder / conditioner)  models/modules/autoencoders/fully_conv_models.py:74-88
 *                   as called by make_flow_input                    models/second_stage_video.py:268-287
 *   ipk_enc_*    <- PokeMotionModel.encode_first_stage             models/second_stage_video.py:352-359
 *                   (ResNetMotionEncoder.forward                   models/modules/motion_models/motion_encoder.py:224-241)
 *   ipk_*_set_tensor takes the reference's state-dict key names unchanged.
 */
#ifndef IPOKE_B200_H_
#define IPOKE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IPK_VERSION 100 /* 0.1.0 */

typedef enum ipk_status {
  IPK_OK = 0,
  IPK_ERR_INVALID = -1,   /* bad argument / config                          */
  IPK_ERR_MISSING = -2,   /* a required state-dict tensor was not provided  */
  IPK_ERR_SHAPE = -3,     /* tensor numel / dtype mismatch                  */
  IPK_ERR_CUDA = -4,      /* CUDA runtime / driver failure                  */
  IPK_ERR_STATE = -5,     /* call order violated (e.g. run before finalize) */
  IPK_ERR_UNSUPPORTED = -6
} ipk_status;

/* Arithmetic of the GEMM/conv contractions.  State, norms, affine transforms and log-det are fp32 always. */
typedef enum ipk_precision {
  IPK_PREC_FP32_SIMT = 0,   /* fp32 FFMA kernels everywhere (validation engine, small layers)              */
  IPK_PREC_FP32_SPLIT = 1,  /* tcgen05 bf16x3 error-compensated MMA (a_hi*b_hi + a_lo*b_hi + a_hi*b_lo)      */
  IPK_PREC_BF16 = 2         /* tcgen05 bf16 operands, fp32 accumulate                                       */
} ipk_precision;

typedef enum ipk_dtype { IPK_F32 = 0, IPK_I64 = 1, IPK_U8 = 2 } ipk_dtype;

#define IPK_MAX_LEVELS 32
#define IPK_MAX_DEC 8

/* Mirrors the keys SupervisedMacowTransformer reads from config['architecture'] (INN.py:448-467). */
typedef struct ipk_flow_config {
  int32_t flow_in_channels;           /* C0                                   */
  int32_t flow_mid_channels;          /* hidden width of the NICE couplings   */
  int32_t h_channels;                 /* conditioning channels                */
  int32_t n_levels;                   /* len(num_steps)                       */
  int32_t num_steps[IPK_MAX_LEVELS];
  int32_t factor;
  int32_t kernel_h, kernel_w;         /* kernel_size (2,3)                    */
  int32_t precision;                  /* ipk_precision                        */
  int32_t max_batch;                  /* workspace is sized for this batch    */
} ipk_flow_config;

/* Mirrors config['architecture'] of the first stage (config/first_stage.yaml:50-63). */
typedef struct ipk_fs_config {
  int32_t z_dim;
  int32_t spatial;                    /* output H = W (64 or 128)             */
  int32_t n_gru_layers;
  int32_t n_dec;                      /* len(dec_channels)                    */
  int32_t dec_channels[IPK_MAX_DEC];
  int32_t precision;                  /* ipk_precision                        */
  int32_t max_batch;
  int32_t max_frames;                 /* max T per decode call                */
  int32_t chunk_videos;               /* videos decoded per pass (0 = auto)   */
} ipk_fs_config;

/* Mirrors the keys ResNetMotionEncoder.__init__ reads (motion_encoder.py:152-160). */
typedef struct ipk_enc_config {
  int32_t z_dim;
  int32_t img_size;                   /* H = W of the input frames (64 or 128)           */
  int32_t max_frames;                 /* data.max_frames; the encoder sees max_frames + 1 */
  int32_t full_seq;
  int32_t n_channels;                 /* len(ENC_M_channels)                              */
  int32_t channels[IPK_MAX_DEC];
  int32_t min_spatial_size;
  int32_t max_batch;
  int32_t precision;                  /* ipk_precision: tensor-core modes run every Conv3d with >= 64 input channels on tcgen05 */
} ipk_enc_config;

/* ConvEncoder(nf_in, nf_max, n_stages, variational=False) as wired by FirstStageWrapper (fully_conv_models.py:9-22). */
typedef struct ipk_cenc_config {
  int32_t nf_in;                      /* 2 (poke map), 3 (image) or 5 (poke map + image, embed_poke_and_image)   */
  int32_t nf_max;
  int32_t spatial;                    /* input H = W                                       */
  int32_t min_spatial_size;
  int32_t n_stages;                   /* log2(spatial / min_spatial_size)                  */
  int32_t max_batch;
} ipk_cenc_config;

/* I3D(num_classes, 'rgb') of the in-repo Frechet video distance (utils/metrics.py:999-1105). */
typedef struct ipk_i3d_config {
  int32_t num_classes;                /* 400; a multiple of 8                                                 */
  int32_t max_batch;
  int32_t max_frames;                 /* >= 9: the (2,7,7) average pool needs two time steps after 3 halvings */
  int32_t precision;                  /* IPK_PREC_FP32_SPLIT (bf16x3 tcgen05, default) or IPK_PREC_BF16        */
} ipk_i3d_config;

/* Devices: a plan belongs to the CUDA device that was current when it was created (its packed weights and workspaces live there);
 * call every entry point of that plan with the same device current.  One process may hold plans on several devices: one-time setup
 * (kernel attributes, SM counts, the staging buffers of ipk_sample_host, graph-capture streams) is kept per device ordinal. */
typedef struct ipk_flow ipk_flow;
typedef struct ipk_fs ipk_fs;
typedef struct ipk_enc ipk_enc;
typedef struct ipk_cenc ipk_cenc;
typedef struct ipk_i3d ipk_i3d;

int ipk_version(void);
const char* ipk_last_error(void);
/* number of kernels this library has launched on the calling thread since the last reset */
int64_t ipk_launch_count(void);
void ipk_launch_count_reset(void);
/* optional per-phase device timing: when enabled every phase of a plan run is bracketed by CUDA events on the launch
 * stream; ipk_prof_report synchronises the device and writes "tag count total_ms\n" lines into buf (returns bytes needed) */
void ipk_prof_enable(int on);
int ipk_prof_report(char* buf, int cap);

/* ---- conditional MaCow flow ---- */
int ipk_flow_create(const ipk_flow_config* cfg, ipk_flow** out);
/* name: state-dict key without the module prefix, e.g. "flow.layers.0.0.actnorm1.log_scale". The data is
 * read during ipk_flow_finalize and must stay valid until then. */
int ipk_flow_set_tensor(ipk_flow* f, const char* name, const void* dev_ptr, int64_t numel, int dtype);
/* folds weight-norm, packs operands for the selected precision; synchronises `stream`. */
int ipk_flow_finalize(ipk_flow* f, void* stream);
/* sampling direction: z[B,C0,8,8], cond[B,h,8,8] -> out[B,C0,8,8] */
int ipk_flow_reverse(ipk_flow* f, const float* z, const float* cond, float* out, int32_t B, void* stream);
/* density direction: x -> z[B,C0,8,8], logdet[B] */
int ipk_flow_forward(ipk_flow* f, const float* x, const float* cond, float* z, float* logdet, int32_t B, void* stream);
int ipk_flow_destroy(ipk_flow* f);
/* data-dependent initialisation of a freshly constructed flow (every `initialized` buffer 0): ActNorm2dFlow.init
 * (models/modules/INN/macow2.py:503-505,526-539: per-channel mean / UNBIASED std of the ActNorm's input over (B,H,W), +1e-6) and
 * Conv2dWeightNorm.init (models/modules/INN/macow_utils.py:231-250; zero_init=True at :281,:423 -> weight_g = 0, bias = 0).
 * Call between ipk_flow_set_tensor and ipk_flow_finalize: the registered log_scale / bias / weight_g tensors are OVERWRITTEN in place
 * (they are the caller's parameters); the caller then sets its `initialized` buffers to 1.  x[B,C0,8,8], B > 1. */
int ipk_flow_data_init(ipk_flow* f, const float* x, int32_t B, void* stream);

/* ---- first-stage decoder: latent ConvGRU + SPADE decoder ---- */
int ipk_fs_create(const ipk_fs_config* cfg, ipk_fs** out);
int ipk_fs_set_tensor(ipk_fs* d, const char* name, const void* dev_ptr, int64_t numel, int dtype);
int ipk_fs_finalize(ipk_fs* d, void* stream);
/* motion[B,z,8,8], x0[B,3,S,S] -> frames[B,T,3,S,S] */
int ipk_fs_decode(ipk_fs* d, const float* motion, const float* x0, float* frames, int32_t B, int32_t T, void* stream);
/* one ConvGRU step (ConvGRU.forward): x[B,z,8,8], hidden[L][B,z,8,8] -> new_hidden[L][B,z,8,8] */
int ipk_fs_gru_step(ipk_fs* d, const float* x, const float* hidden, float* new_hidden, int32_t B, void* stream);
/* one decoder pass (SpadeCondConvDecoder.forward): h[B,z,8,8], x0[B,3,S,S] -> frame[B,3,S,S] */
int ipk_fs_gen(ipk_fs* d, const float* h, const float* x0, float* frame, int32_t B, void* stream);
int ipk_fs_destroy(ipk_fs* d);

/* ---- first-stage 3-D conv video encoder (second-stage training path; no gradients) ---- */
int ipk_enc_create(const ipk_enc_config* cfg, ipk_enc** out);
int ipk_enc_set_tensor(ipk_enc* e, const char* name, const void* dev_ptr, int64_t numel, int dtype);
int ipk_enc_finalize(ipk_enc* e, void* stream);
/* X[B,3,T,S,S] (NCDHW), eps[B,z,8,8] (reparameterisation noise, drawn by the host so RNG stays in Python)
 * -> z = eps * exp(logvar / 2) + mu, mu, logvar, each [B,z,8,8] */
int ipk_enc_forward(ipk_enc* e, const float* X, const float* eps, float* z, float* mu, float* logvar, int32_t B, int32_t T, void* stream);
int ipk_enc_destroy(ipk_enc* e);

/* ---- conditioning encoders (frozen ConvEncoder of the poke embedder / image conditioner) ---- */
int ipk_cenc_create(const ipk_cenc_config* cfg, ipk_cenc** out);
int ipk_cenc_set_tensor(ipk_cenc* e, const char* name, const void* dev_ptr, int64_t numel, int dtype);
int ipk_cenc_finalize(ipk_cenc* e, void* stream);
/* x[B,nf_in,S,S] -> out[B,nf_max,8,8] (after the bottleneck) and, if non-null, mean[B,nf_max,8,8] (before it) */
int ipk_cenc_forward(ipk_cenc* e, const float* x, float* out, float* mean, int32_t B, void* stream);
int ipk_cenc_destroy(ipk_cenc* e);

/* ---- I3D feature extractor of the in-repo FVD: I3D.forward, utils/metrics.py:1079-1105 (Unit3Dpy :857-937 with
 *      TF-SAME padding :814-842, MaxPool3dTFPadding :940-960, Mixed :963-997); tensors by the reference's state-dict names
 *      ("conv3d_1a_7x7.conv3d.weight", "mixed_3b.branch_1.1.batch3d.running_var", ...).  BatchNorm is folded at finalize (eval mode). ---- */
int ipk_i3d_create(const ipk_i3d_config* cfg, ipk_i3d** out);
int ipk_i3d_set_tensor(ipk_i3d* m, const char* name, const void* dev_ptr, int64_t numel, int dtype);
int ipk_i3d_finalize(ipk_i3d* m, void* stream);
/* x[B,3,T,224,224] fp32 (the layout get_activations feeds, utils/metrics.py:726) -> logits[B,num_classes] (the second output of I3D.forward) */
int ipk_i3d_forward(ipk_i3d* m, const float* x, float* logits, int32_t B, int32_t T, void* stream);
int ipk_i3d_destroy(ipk_i3d* m);
/* preprocess (utils/metrics.py:786-802) of one set of frames: videos[n_frames,3,S,S] -> out[n_frames,3,224,224], bilinear with
 * align_corners; the (x + 1) / 2 map is applied when any resized value of the set is negative (decided on the device) */
int ipk_i3d_preprocess(const float* videos, float* out, int64_t n_frames, int32_t S, void* stream);

/* ---- second-stage training step of the flow (BASELINE configs[3]): forward_density + FlowLoss + backward
 *      models/second_stage_video.py:345-350, models/modules/INN/loss.py:13-31; optimizer: second_stage_video.py:633-660 ----
 * Tensors are registered by their checkpoint names with a gradient destination each (fp32 tensors; integer buffers pass NULL).
 * ipk_flowtrain_step re-packs the weights from the current parameter values, runs the density direction keeping every op's input on a
 * tape, writes loss = mean_B(0.5 sum z^2) - mean_B(logdet) to *loss_out (device) and OVERWRITES every gradient buffer with dloss/dparam.
 * precision: IPK_PREC_FP32_SPLIT (tcgen05 bf16x3 contractions) or IPK_PREC_FP32_SIMT. */
typedef struct ipk_flowtrain ipk_flowtrain;
int ipk_flowtrain_create(const ipk_flow_config* cfg, ipk_flowtrain** out);
int ipk_flowtrain_set_tensor(ipk_flowtrain* f, const char* name, const void* param, float* grad, int64_t numel, int dtype);
int ipk_flowtrain_finalize(ipk_flowtrain* f, void* stream);
int ipk_flowtrain_step(ipk_flowtrain* f, const float* x, const float* cond, float* loss_out, float* z_out, float* logdet_out,
                       int32_t B, void* stream);
/* the same step split at the loss, for a caller that computes its own loss between the two halves (torch.autograd.Function around
 * `out, logdet = self.flow(x, cond)` ... `loss.backward()`, models/second_stage_video.py:409-415):
 *   forward : re-pack + density direction with the tape kept -> z[B,C0,8,8], logdet[B]
 *   backward: upstream dz[B,C0,8,8], dlogdet[B] (NULL = zeros) -> every registered gradient buffer is OVERWRITTEN with dL/dparam;
 *             dx_out (optional) receives dL/dx[B,C0,8,8].
 *             Must follow a forward with the same B on the same plan. */
int ipk_flowtrain_forward(ipk_flowtrain* f, const float* x, const float* cond, float* z_out, float* logdet_out, int32_t B, void* stream);
int ipk_flowtrain_backward(ipk_flowtrain* f, const float* dz, const float* dlogdet, float* dx_out, int32_t B, void* stream);
int ipk_flowtrain_destroy(ipk_flowtrain* f);
/* torch.optim.Adam (amsgrad when max_exp_avg_sq != NULL) in place on a contiguous fp32 shard; step counts from 1; the gradient is
 * multiplied by grad_scale first (1 / world_size after a summing reduce-scatter).  Hyper-parameters are doubles: the bias corrections
 * 1 - beta^step are evaluated in double precision like torch.optim.Adam does (models/second_stage_video.py:647-648 uses beta2 = 0.999). */
int ipk_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* max_exp_avg_sq, int64_t n, double lr,
                  double beta1, double beta2, double eps, double weight_decay, int32_t step, float grad_scale, void* stream);

/* ---- whole sampling step with DEVICE buffers: flow inverse -> GRU + decoder ---- */
int ipk_sample(ipk_flow* f, ipk_fs* d, const float* z, const float* cond, const float* x0, float* frames,
               int32_t B, int32_t T, void* stream);
/* ---- same with HOST buffers (pinned or pageable): H2D, compute, D2H, stream-synchronised on return ---- */
int ipk_sample_host(ipk_flow* f, ipk_fs* d, const float* z_host, const float* cond_host, const float* x0_host,
                    float* frames_host, int32_t B, int32_t T, void* stream);

/* ---- sample post-processing: what the callers of forward_sample do with every sample on the host,
 * `((x + 1.) * 127.5).permute(0, 1, 3, 4, 2).numpy().astype(np.uint8)` (models/second_stage_video.py:673-675; the input of
 * utils/logging.py:797 save_video), done on the device so that 1 byte per value crosses PCIe instead of 4.
 * frames: device fp32 [n_frames][3][S][S] in [-1, 1] -> out: device uint8 [n_frames][S][S][3] (truncation, clamped). */
int ipk_frames_to_u8(const float* frames, uint8_t* out, int64_t n_frames, int32_t spatial, void* stream);
/* ipk_sample_host delivering uint8 NTHWC frames [B][T][S][S][3] to the host, chunk by chunk behind the decoder */
int ipk_sample_host_u8(ipk_flow* f, ipk_fs* d, const float* z_host, const float* cond_host, const float* x0_host,
                       uint8_t* frames_u8_host, int32_t B, int32_t T, void* stream);

/* ---- kernel-level test hooks (used by tests/ to check single kernels against the oracle) ---- */
/* out[M,N] = A[M,K] * W[N,K]^T through the engine selected by `precision` */
int ipk_test_gemm(const float* A, const float* W, float* out, int32_t M, int32_t N, int32_t K, int32_t precision, void* stream);
/* 3x3 stride-1 pad-1 conv, NHWC in[F,H,W,Cin], OIHW weights w[Cout,Cin,3,3] -> NHWC out[F,H,W,Cout] */
int ipk_test_conv3x3(const float* in, const float* w, const float* bias, float* out, int32_t F, int32_t H, int32_t W,
                     int32_t Cin, int32_t Cout, int32_t precision, void* stream);
/* ConvTranspose2d(3, stride 2, pad 1, output_pad 1), NHWC in[F,H,W,Cin], IOHW weights -> NHWC out[F,2H,2W,Cout] */
int ipk_test_convT3x3(const float* in, const float* w, const float* bias, float* out, int32_t F, int32_t H, int32_t W,
                      int32_t Cin, int32_t Cout, int32_t precision, void* stream);
/* Conv3d on the tcgen05 engine: NDHWC in[B,T,H,W,Cin], OIDHW w[Cout,Cin,kt,ky,kx] -> NDHWC out[B,To,Ho,Wo,Cout]; optional
 * stats[B,Cout,2] (fp64 sum / sum of squares per sample and channel, the fused GroupNorm statistics).
 * dims = {B,T,H,W,Cin,Cout, kt,ky,kx, st,sy,sx, pt,py,px}; precision IPK_PREC_FP32_SPLIT or IPK_PREC_BF16 */
int ipk_test_conv3d(const float* in, const float* w, float* out, double* stats, const int32_t* dims, int32_t precision, void* stream);

/* Host-side tile planning of the tcgen05 conv engine (no device work: callable without a GPU): for a layer with Npad output columns,
 * tiles_m 128-row M tiles, total_iters (tap, 64-wide k-block) iterations, nsub sub-convolutions, nsplit split-K slices, a fused
 * (residual / statistics) epilogue or not, on a device with `sms` SMs.  out[0..4] = kernel N capacity BN, columns per tile bn, CTA-pair
 * factor CG (1 | 2), N tiles, number of uneven tiles; out[5 + 2 i], out[6 + 2 i] = first column and width of uneven tile i. */
int ipk_test_tc_plan(int32_t Npad, int32_t tiles_m, int32_t total_iters, int32_t nsub, int32_t nsplit, int32_t fused, int32_t sms, int32_t* out);

/* Timeline probe of the tcgen05 conv engine (profiles/tc_trace_probe.py): after ipk_tc_trace_enable(N, Kpad) every conv_tc launch whose
 * packed weights have that N and padded K records 32 SM-clock stamps per CTA (entry, prologue, first TMA, per-tile MMA / epilogue
 * milestones; slot 30 = global timer at entry); ipk_tc_trace_read copies [n_ctas][32] int64 of the LAST such launch to the host. */
int ipk_tc_trace_enable(int32_t N, int32_t Kpad);
int ipk_tc_trace_read(long long* host, int32_t n_ctas);

#ifdef __cplusplus
}
#endif
#endif /* IPOKE_B200_H_ */
