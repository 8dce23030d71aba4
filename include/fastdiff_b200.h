/*
 * fastdiff_b200 -- C ABI of the B200-native FastDiff reverse-diffusion sampling path.
 *
 * The reference (Rongjiehuang/FastDiff) has no FFI layer: its boundary for this path is two Python
 * call signatures.  Each entry point below names the reference interface it stands behind
 * (file:line under /root/reference); fastdiff_b200/{model,sampler}.py bind them with ctypes and
 * re-expose the reference's own Python signatures on top.
 *
 * Conventions
 *   - return 0 on success, negative fd_status on failure; fd_last_error() gives the message.
 *   - every pointer named *_dev is a DEVICE pointer owned by the caller (PyTorch); the library
 *     borrows it for the duration of the call, launches on the caller's stream (plus one internal
 *     side stream for the DiffusionDBlock chain, forked from and joined back into the caller's
 *     stream with events inside the same call, so the caller sees ordinary stream semantics) and never
 *     synchronises the host.  Nothing is allocated after fd_load_weights(); scratch memory is the
 *     caller-provided workspace (fd_workspace_bytes).
 *   - layouts are the reference's: audio (B,1,L) fp32, mel (B,80,T') fp32 NCL, L = 256*T',
 *     diffusion steps (B) fp32 (fractional steps).
 *   - one handle per (device, stream); a handle is not thread-safe, distinct handles are independent.
 *   - architecture: the one network the reference ships (modules/FastDiff/config/base.yaml:21-33);
 *     fd_create() rejects any other fd_config with FD_ERR_UNSUPPORTED.
 */
#ifndef FASTDIFF_B200_H
#define FASTDIFF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fd_handle fd_handle;

typedef enum fd_status {
    FD_OK = 0,
    FD_ERR_INVALID = -1,      /* bad argument (shape, null pointer, blob mismatch) */
    FD_ERR_UNSUPPORTED = -2,  /* architecture / device the kernels are not built for */
    FD_ERR_CUDA = -3,         /* a CUDA runtime call or launch failed */
    FD_ERR_STATE = -4         /* call order (e.g. denoise before load_weights) */
} fd_status;

/* Mirrors the ctor kwargs of FastDiff.__init__ (modules/FastDiff/module/FastDiff_model.py:13-26)
 * = the hparams read by FastDiffTask.build_model (modules/FastDiff/task/FastDiff.py:17-29). */
typedef struct fd_config {
    int32_t audio_channels;               /* 1   */
    int32_t inner_channels;               /* 32  */
    int32_t cond_channels;                /* 80  */
    int32_t n_upsample;                   /* 3   */
    int32_t upsample_ratios[4];           /* 8,8,4 */
    int32_t lvc_layers_each_block;        /* 4   */
    int32_t lvc_kernel_size;              /* 3   */
    int32_t kpnet_hidden_channels;        /* 64  */
    int32_t kpnet_conv_size;              /* 3   */
    int32_t diffusion_step_embed_dim_in;  /* 128 */
    int32_t diffusion_step_embed_dim_mid; /* 512 */
    int32_t diffusion_step_embed_dim_out; /* 512 */
} fd_config;

/* Per-reverse-step scalars, computed ON THE HOST exactly as the reference computes them
 * (modules/FastDiff/module/util.py:187-204 for t; :219-229 for the update), one entry per executed
 * step in execution order (n = N-1 ... 0).
 *   ddim == 0:  x = (x - coef_eps*eps) / div;  if (add_noise) x = x + sigma*z      (util.py:226-229)
 *   ddim != 0:  x = c1*x + c2*eps + c3*eps                                          (util.py:219-224)
 */
typedef struct fd_step {
    float t;         /* steps_infer[n]: fractional diffusion step fed to the embedding */
    float coef_eps;  /* beta_n / sqrt(1 - alpha_n^2) */
    float div;       /* sqrt(1 - beta_n) */
    float sigma;     /* sigma_n */
    float c1, c2, c3;
    int32_t add_noise; /* n > 0 */
} fd_step;

/* Arithmetic mode of the heavy contractions (kernel_conv GEMM, location-variable conv, dilated convs). */
typedef enum fd_mode {
    FD_MODE_FP32_SIMT = 0,   /* fp32 FFMA everywhere (strict; the on-device cross-check path) */
    FD_MODE_TC_3XTF32 = 1,   /* tcgen05 kind::tf32 with hi/lo error compensation (fp32-level) */
    FD_MODE_TC_TF32 = 2,     /* tcgen05 kind::tf32 single pass (fast mode; error reported separately) */
    FD_MODE_TC_3XF16 = 3     /* tcgen05 kind::f16 on fp16 hi/lo pieces of power-of-two prescaled operands, 3 passes, fp32
                                accumulation: the same 22 significant bits per operand as 3xTF32 at twice the MMA rate
                                (fp32-level; operands saturate at |v*S| = 65504).  Default after load. */
} fd_mode;

/* Build a sampler/denoiser for `device`.  Stands behind FastDiff.__init__
 * (modules/FastDiff/module/FastDiff_model.py:13-72). */
int fd_create(const fd_config* cfg, int device, fd_handle** out);

/* Copy the packed weight blob (host memory, produced by fastdiff_b200.weights.pack_state_dict from
 * a reference state_dict: weight-norm folded, layouts permuted) to the device.  Stands behind
 * nn.Module.load_state_dict / Trainer.restore_weights (utils/trainer.py:348-369).  The blob is
 * borrowed during the call only. */
int fd_load_weights(fd_handle* h, const void* blob_host, size_t bytes);

/* Same, but the blob is already in device memory (e.g. received by one NCCL broadcast in
 * batch-shard mode); the library keeps its own device copy. */
int fd_load_weights_dev(fd_handle* h, const void* blob_dev, size_t bytes, void* stream);

/* Scratch bytes needed for a (B, T') problem. */
int fd_workspace_bytes(fd_handle* h, int B, int Tm, size_t* out);

/* Select arithmetic mode (default: the best mode whose kernels are built in). */
int fd_set_mode(fd_handle* h, int mode);
int fd_get_mode(fd_handle* h);

/* Integer options.  "stop_after": run fd_denoise only up to a stage (1 = kernel predictor, 2 = DBlocks,
 * 3/4/5 = LVC block 0/1/2, >= 6 = everything; default) so fd_debug_read can inspect stage outputs.
 * Tuning / cross-check switches (defaults in brackets): "overlap" [1] DBlock chain on the internal side stream; "tc_kp" [1],
 * "tc_dblock" [1], "tc_upsample" [1] tensor-core versions of the kernel-predictor stack / DBlock 0 / upsampling; "kc_2cta" [1];
 * "lvc_groups" [2]; "lvc_swizzle"; "kc_exp", "lvc_exp" [0] timing experiments that produce WRONG results;
 * "tc_b0" [0] EXPERIMENTAL tensor-core kernel for LVC block 0 in mode tc_3xf16 (validated on the CPU model only):
 * 1 = fed by the GEMM writing block 0 as fp16 pieces, 2 = fed by an in-place converter pass; "b2_skipbuf" [0] EXPERIMENTAL: LVC block 2
 * reads first_conv(audio) as rows written once per evaluation instead of recomputing it in every layer (same bits); "kc_stage" [0]
 * EXPERIMENTAL: kernel_conv GEMM epilogue through shared memory + cp.async.bulk stores (same bits); "lvc_pipe" [0] EXPERIMENTAL: LVC
 * block 2 with software-pipelined tiles (same bits). */
int fd_set_option(fd_handle* h, const char* key, int64_t value);

/* eps = FastDiff.forward((x_t, mel, t))   (modules/FastDiff/module/FastDiff_model.py:74-102).
 * x_dev (B,1,256*Tm), mel_dev (B,80,Tm), t_dev (B), eps_dev (B,1,256*Tm). */
int fd_denoise(fd_handle* h, const float* x_dev, const float* mel_dev, const float* t_dev,
               float* eps_dev, int B, int Tm, void* workspace_dev, size_t workspace_bytes, void* stream);

/* The reverse loop of sampling_given_noise_schedule (modules/FastDiff/module/util.py:216-234).
 * x_dev: x_T on entry (unless fill_xT), x_0 on return.
 * steps: HOST array of n_steps entries in execution order.
 * noise_dev: parity mode -- device (n_noise,B,1,L) Gaussian draws in the reference's order
 *            (one per step with add_noise); NULL -> on-device Philox4x32-10 keyed by `seed`.
 * fill_xT: draw x_T on the device from the same Philox stream (perf mode).
 * seq_dev: optional (n_steps+1,B,1,L) -- x after every step incl. x_T (return_sequence=True). */
int fd_sample(fd_handle* h, float* x_dev, const float* mel_dev, const fd_step* steps, int n_steps,
              const float* noise_dev, int n_noise, uint64_t seed, int fill_xT, int ddim,
              float* seq_dev, int B, int Tm, void* workspace_dev, size_t workspace_bytes, void* stream);

/* One reverse-step update by itself, in place on x_dev -- the two update rules of sampling_given_noise_schedule
 * (modules/FastDiff/module/util.py:219-229) for denoisers other than FastDiff that share the sampler
 * (modules/FastDiff/module/WaveNet.py:156, modules/parallel_wavegan/models/parallel_wavegan.py:23): the caller evaluates
 * eps = net((x, c, t)) itself and passes it here.  `step` is one HOST fd_step (see above).  z_dev: the Gaussian draw of this step
 * (count floats) or NULL -> Philox4x32-10 keyed by (element, draw, seed) when step->add_noise; ignored when ddim.
 * seq_dev: optional second destination (return_sequence).  x, eps, z, seq: `count` contiguous floats.  Needs no weights. */
int fd_reverse_update(fd_handle* h, float* x_dev, const float* eps_dev, const float* z_dev, const fd_step* step, int ddim,
                      uint64_t seed, uint32_t draw, float* seq_dev, size_t count, void* stream);

/* The step before the path: waveform -> log10-mel on the device, the reference's process_utterance
 * (data_gen/tts/data_gen_utils.py:93-147: librosa 0.8.0 stft(1024, hop 256, periodic Hann, zero centre padding), magnitude,
 * librosa.filters.mel(22050, 1024, 80, 80, 7600), log10(max(1e-6, .))).  wav_dev (B, n_samples) fp32 -> mel_dev
 * (B, 80, 1 + n_samples/256) fp32.  fb_dev (80, 513) fp32 / range_dev (80, 2) int32: the mel filter table and the [lo, hi)
 * range of non-zero FFT bins of every filter (built by fastdiff_b200/mel.py).  Needs no weights. */
int fd_mel_frontend(fd_handle* h, const float* wav_dev, int B, int n_samples, const float* fb_dev, const int32_t* range_dev,
                    float* mel_dev, void* stream);

/* The step after the path: per-utterance peak normalisation and int16 encode on the device --
 * wav_pred / wav_pred.abs().max() (modules/FastDiff/task/FastDiff.py:110) followed by utils/audio.py:11-16
 * (wav *= 32767; astype(int16)).  x_dev (B,1,L) fp32 -> out_dev (B,L) int16, bit-identical to the reference's float ops.
 * Uses the first B words of the workspace as scratch. */
int fd_wav_int16(fd_handle* h, const float* x_dev, int16_t* out_dev, int B, int L, void* workspace_dev, void* stream);

/* Stage helpers exposed for the parity tests (each is one stage of FastDiff.forward; same
 * conventions).  fd_debug_read copies a named internal tensor of the LAST fd_denoise call out of
 * the workspace, converted to the reference's NCL layout:
 *   "embed" (B,512)  FastDiff_model.py:85-87     "down0|1|2" (B,32,L/4|L/32|L/256)  modules.py:127-138
 *   "kp_hidden0|1|2" (B,64,T')  modules.py:328-329
 *   "kernels0|1|2" (B,4,32,64,3,T') and "kbias0|1|2" (B,4,64,T')  modules.py:330-342
 *   "lvc0|1|2" (B,32,8T'|64T'|256T')  modules.py:190-218
 * out_dev must hold *count floats; if out_dev is NULL only *count is written. */
int fd_debug_read(fd_handle* h, const char* name, float* out_dev, size_t* count, int B, int Tm,
                  void* workspace_dev, void* stream);

/* Number of kernels launched by this handle since creation (bench.py's gpu_launches). */
/* Device-noise mode (noise_dev == NULL): Philox element index of local sample (b, t) = b * total_samples + offset + t.  The default
 * (0, 0) indexes the tensor itself.  A rank that holds the time window [offset, offset + L) of utterances of total_samples samples
 * (time-shard mode, fastdiff_b200/timeshard.py) draws exactly the numbers the unsharded call draws for those samples. */
int fd_set_noise_window(fd_handle* h, int64_t total_samples, int64_t offset);

uint64_t fd_launch_count(fd_handle* h);

/* Range guard of the default arithmetic mode (tc_3xf16): operands of the tensor-core kernels are fp16 pieces of 16 x activation
 * (64 x predicted kernel) and SATURATE at |piece| = 65504 instead of overflowing.  Every kernel that forms pieces raises a sticky
 * device flag when a value left the range (|activation| >= 4094); this call synchronises `stream`, returns the flag in *flag (0 / 1)
 * and clears it when `reset` != 0.  A raised flag means the results of the calls since the last reset are NOT fp32-level: re-run in
 * FD_MODE_TC_3XTF32 (fp32 range).  The reference has no counterpart (it computes in fp32, FastDiff_model.py:74-102). */
int fd_check_saturation(fd_handle* h, int* flag, int reset, void* stream);

/* Optional device timing per kernel class (CUDA events around every launch on the caller's stream).
 * After the caller has synchronised the stream, fd_timing_report writes one JSON object
 * {"kc_gemm": {"ms": total, "n": launches}, ...} for the launches since the previous report. */
int fd_timing_enable(fd_handle* h, int on);
int fd_timing_report(fd_handle* h, char* buf, size_t buf_bytes);

const char* fd_last_error(fd_handle* h);   /* valid until the next call on h; h may be NULL for create errors */
void fd_destroy(fd_handle* h);
const char* fd_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FASTDIFF_B200_H */
