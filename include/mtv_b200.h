/*
 * mtv_b200.h — C ABI of libmtv_b200.so: the B200 (sm_100a) implementation of the
 * MoDiTalker MToV denoising hot path.
 *
 * The reference (cvlab-kaist/MoDiTalker) has no native code and no FFI for this
 * path: the "interface" it exposes is the Python nn.Module contract
 *     DiffusionWrapper(UNetModel(**unet_config)).forward(x, cond, image_cond, t)
 *         MToV/models/ddpm/unet.py:41-61, 995-1117
 *     DDPM.model_predictions / ddim_sample / ddim_sample_noised_start / q_sample
 *         MToV/losses/ddpm.py:338-360, 363-404, 407-454, 486-491
 * Each entry point below names the reference function it replaces.  The binding a
 * maintainer adds on the reference side is the ctypes stub in INTEGRATION.md
 * (and, as shipped, moditalker_b200/_lib.py).
 *
 * Conventions
 *   - every tensor argument is a raw DEVICE pointer to contiguous fp32 (int64 for
 *     timesteps) owned by the caller (PyTorch); the library owns only its repacked
 *     weights and workspace, released by mtv_destroy
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered.  The FIRST
 *     mtv_unet_forward for a new batch size builds that batch size's launch plan (workspace
 *     allocation, one cudaDeviceSynchronize; must not happen while the caller is capturing a
 *     graph); at most four plans are cached, least recently used dropped.  Every later call
 *     is free of host synchronisation
 *   - caller state is left alone: every entry point restores the caller's current CUDA
 *     device; the caller's stream never receives attributes (the L2 access-policy window is
 *     set on a private capture stream only; the device-wide persisting-L2 limit is restored
 *     by mtv_destroy)
 *   - results are bit-reproducible for a given (weights, inputs, batch size): GroupNorm sums
 *     and split-K reductions use fixed orders, no floating-point atomics
 *   - every function returns 0 on success, non-zero on error; mtv_last_error()
 *     returns a message for the calling thread's last failure.  No C++ exception
 *     crosses this boundary.
 *   - one handle per device, not thread-safe.
 */
#ifndef MTV_B200_H
#define MTV_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MTV_MAX_LEVELS 8
#define MTV_ABI_VERSION 1

typedef struct MtvHandle_t* MtvHandle;

/* Constructor arguments of UNetModel that reach the hot path
 * (unet.py:631-659; values from configs/latent-diffusion/base.yaml:28-39). */
typedef struct MtvConfig {
  int32_t abi_version;            /* MTV_ABI_VERSION */
  int32_t image_size;             /* 32: xy plane is image_size^2, yt/xt are (image_size/2) x image_size */
  int32_t in_channels;            /* 4  (latent channels of x) */
  int32_t out_channels;           /* 4 */
  int32_t model_channels;         /* 128 (base) / 256 (longvid) */
  int32_t num_res_blocks;         /* 2 */
  int32_t num_heads;              /* 8 */
  int32_t num_levels;             /* len(channel_mult) */
  int32_t channel_mult[MTV_MAX_LEVELS];
  int32_t attn_at_level[MTV_MAX_LEVELS]; /* 1 if (1<<level) is in attention_resolutions */
  int32_t device;                 /* CUDA device ordinal */
  int32_t kernel_path;            /* 0 = default: tcgen05 tensor-core kernels (split-bf16, ~1e-5 vs fp32) for every
                                         tap-GEMM (stem and head included) and every attention with a tile shape,
                                     1 = fp32 CUDA-core kernels everywhere (cross-check path of the tests, ~1e-6) */
} MtvConfig;

int  mtv_abi_version(void);
const char* mtv_last_error(void);

/* UNetModel.__init__ (unet.py:631-975): builds the launch plan, allocates weight storage. */
int  mtv_create(const MtvConfig* cfg, MtvHandle* out);
int  mtv_destroy(MtvHandle h);

/* nn.Module.load_state_dict (sample.py:227-231).  `name` is a state-dict key with or
 * without the "diffusion_model." prefix; `data` is a DEVICE pointer to contiguous fp32 in
 * the PyTorch layout of that key ([Cout,Cin,3,3], [Cout,Cin,1], [out,in], [C]).  The
 * tensor is copied and repacked on `stream`.  Keys the forward never reads
 * (output_bg_*, zeros) are accepted and ignored: returns 0 and *used = 0. */
int  mtv_load_weight(MtvHandle h, const char* name, const float* data,
                     const int64_t* shape, int32_t ndim, int32_t* used, void* stream);
/* Number of tensors the forward needs / how many are still missing; fails (non-zero)
 * when any is missing and writes the first missing key into mtv_last_error(). */
int  mtv_weights_ready(MtvHandle h, int32_t* needed, int32_t* missing);
/* Number of weight names the plan reads, and the i-th name (for tests / bindings). */
int  mtv_num_weight_names(MtvHandle h);
const char* mtv_weight_name(MtvHandle h, int32_t i);

/* DiffusionWrapper.forward -> UNetModel.forward (unet.py:41-44, 995-1117).
 *   x          [B, in_channels, 2048]
 *   cond       [B, 2*in_channels, 2048]
 *   image_cond [B, in_channels, image_cond_len]   (only [:, :, :1024] is read, unet.py:1024)
 *   t          [B] int64
 *   out        [B, out_channels, 2048]            (epsilon prediction) */
int  mtv_unet_forward(MtvHandle h, const float* x, const float* cond, const float* image_cond,
                      int64_t image_cond_len, const int64_t* t, int32_t B, float* out, void* stream);

/* One DDIM update, DDPM.model_predictions + the body of ddim_sample's loop
 * (ddpm.py:346-351, 386-398), in place on img [n]:
 *   x0  = clamp(sqrt_recip_ac * img - sqrt_recipm1_ac * eps, -1, 1)
 *   img = x0                                              if last
 *   img = x0 * sqrt_alpha_next + c * eps + sigma * noise  otherwise
 * The five scalars are computed by the host exactly as the reference computes them
 * (fp32).  `noise` may be NULL when last != 0. */
int  mtv_ddim_step(MtvHandle h, float* img, const float* eps, const float* noise, int64_t n,
                   float sqrt_recip_ac, float sqrt_recipm1_ac, float sqrt_alpha_next,
                   float c, float sigma, int32_t last, void* stream);

/* DDPM.q_sample (ddpm.py:486-491): out = a * x_start + b * noise, elementwise over n. */
int  mtv_q_sample(MtvHandle h, const float* x_start, const float* noise, int64_t n,
                  float a, float b, float* out, void* stream);

/* ---- chunk I/O around the loop: the per-chunk host work of MToV/sample.py as stream-ordered device
 * kernels.  Results are bit-exact against the reference's numpy / cv2 / torch-CPU sequence (the bilinear resize reproduces
 * torch's CPU kernel operation for operation, including its two rounding regimes: outputs up to 64 pixels wide / wider).  No handle: `device` is the CUDA ordinal the pointers live on.  All pointers
 * are DEVICE pointers. ---- */

/* EvalDataset._load_img_from_path + _crop_lower_half + resize_crop (tools/dataloader_sample.py:130-146, tools/data_utils.py:73-98)
 * followed by `x / 127.5 - 1` and "b t c h w -> b c t h w" (sample.py:322-325), for ONE clip:
 *   frames   uint8 [T, H, W, 3]   decoded RGB frames as PIL yields them
 *   mask_row int32 [T] or NULL    when given, frame t keeps rows < mask_row[t], the rest is zeroed and the kept pixels are
 *                                 truncated to integers as (img * mask).astype(np.uint8) does; the host passes the row
 *                                 numpy's `mask[int(landmarks[33][1]):, :] = 0` resolves to (negative values count from H)
 *   out      fp32  [3, T, R, R]   centre crop to min(H, W), bilinear (align_corners = False) to R x R, normalised to [-1, 1] */
int  mtv_io_prep_frames(int32_t device, const uint8_t* frames, int32_t T, int32_t H, int32_t W, const int32_t* mask_row,
                        int32_t R, float* out, void* stream);

/* Same, with flags.  MTV_IO_LOADER_WORKER: reproduce the resize as torch computes it inside a DataLoader WORKER process, which
 * is how the shipped script runs it (get_loaders: num_workers = 4, tools/dataloader_sample.py:288-294).  A worker has one torch
 * thread, and with one thread torch resizes 3-channel images with its "vectorized" CPU kernel at every output size; in a
 * multi-threaded process (what mtv_io_prep_frames reproduces) it does so only while out_h + out_w <= 128 and uses its generic
 * kernel above.  The two kernels differ in rounding (at most one ulp of the 0..255 value; identical whenever the weights are
 * exactly representable, e.g. even source sizes at R = 256). */
#define MTV_IO_LOADER_WORKER 1
int  mtv_io_prep_frames_ex(int32_t device, const uint8_t* frames, int32_t T, int32_t H, int32_t W, const int32_t* mask_row,
                           int32_t R, int32_t flags, float* out, void* stream);

/* EvalDataset._change_np_img_size (tools/dataloader_sample.py:153-180) + sample.py:324: landmark clip -> key-point video.
 *   landmarks fp32 or fp64 (is_f64) [T, N, dims]; dims == 3: normalised coordinates, pixel = int(v * WH / 2 + WH / 2) in the
 *             array's precision; dims == 2: pixel = int(v).  Then centre = int(pixel / WH * 256.0), and the filled radius-3
 *             disc cv2.circle draws, clipped to the 256 x 256 canvas; flip != 0 mirrors rows (cv2.flip(img, 0))
 *   out       fp32 [3, T, 256, 256] in {-1, +1} (white discs on black, already normalised) */
int  mtv_io_rasterize_landmarks(int32_t device, const void* landmarks, int32_t is_f64, int32_t T, int32_t N, int32_t dims, int32_t WH,
                                int32_t flip, float* out, void* stream);

/* sample.py:380-399 and 344-358: decoded frames -> what the script writes and what the next chunk reads back.
 *   dec       fp32  [B*T, 3, H, W]       ViTAutoencoder.decode_from_sample output (clamped to [-1, 1] here)
 *   frames_u8 uint8 [B, T, H, W, 3]      (1 + dec) * 127.5 truncated (fake.type(torch.uint8)); may be NULL
 *   last_u8   uint8 [B, H, W, 3]         frame T-1, clip(rint(.), 0, 255): the pixels of the PNG the script saves (RGB); may be NULL
 *   next_ref  fp32  [B, 3, Trep, H, W]   that PNG read back: (u8 / 255) * 2 - 1, repeated Trep times along time; may be NULL */
int  mtv_io_frames_out(int32_t device, const float* dec, int32_t B, int32_t T, int32_t H, int32_t W, uint8_t* frames_u8,
                       uint8_t* last_u8, float* next_ref, int32_t Trep, void* stream);

/* Introspection used by bench.py / tests: kernels launched by one forward at batch B
 * (graph nodes included), workspace bytes, algorithmic weight bytes read per forward. */
int  mtv_plan_info(MtvHandle h, int32_t B, int64_t* n_launches, int64_t* workspace_bytes,
                   int64_t* weight_bytes);

/* Debug taps (tests only): copy the token-major activation [B, L, C] produced by stage
 * `tag` ("in0".."in11", "mid", "out0".."out11") of the most recent forward into `dst`
 * as channel-major [B, C, L] fp32.  dst_elems guards the size. */
int  mtv_debug_read(MtvHandle h, const char* tag, float* dst, int64_t dst_elems, void* stream);

/* Diagnostics: arm (records != NULL) or disarm (NULL) in-kernel phase timing of the tensor-core
 * tap-GEMM.  While armed every CTA appends one 16 x int64 record to the DEVICE buffer `records`
 * (capacity `cap` records).  *count receives the number of records written since the previous call. */
int  mtv_debug_tc_timing(MtvHandle h, int64_t* records, int32_t cap, int32_t* count);

/* Per-kernel timing of one forward (CUDA events around every launch, serialised):
 * fills up to `cap` entries of (name, microseconds); returns the count in *n. */
typedef struct MtvKernelTime { char name[48]; float us; float flops; float bytes; } MtvKernelTime;
int  mtv_profile_forward(MtvHandle h, const float* x, const float* cond, const float* image_cond,
                         int64_t image_cond_len, const int64_t* t, int32_t B, float* out,
                         MtvKernelTime* entries, int32_t cap, int32_t* n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MTV_B200_H */
