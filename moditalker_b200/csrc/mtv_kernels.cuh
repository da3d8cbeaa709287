// mtv_kernels.cuh — parameter blocks and launchers of the CUDA kernels behind
// libmtv_b200.so (sm_100a only).  Host-side plan code (mtv_plan.cu) includes this;
// kernels live in kernels_simt.cu (fp32 CUDA-core kernels) and kernels_tc.cu
// (tcgen05 / TMA tensor-core kernels).
//
// Data layout in HBM: every activation is TOKEN-MAJOR fp32
// [B][L][C]; the token axis of one sample is the tri-plane concatenation
// xy | yt | xt, each plane row-major (h, w) — the order the reference itself uses
// when it flattens and concatenates planes for the cross-plane attention
// (MToV/models/ddpm/unet.py:1039-1043).  Level l has res = 32>>l, t = 16>>l,
// L = res*(res+2t) tokens.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <utility>

// Programmatic dependent launch: every kernel of the forward is launched with the
// programmatic-stream-serialization attribute, triggers its dependents at entry and executes
// griddepcontrol.wait before it first touches memory written by its predecessor, so launch latency
// and kernel prologues (barrier init, TMEM allocation, weight TMA requests) overlap the tail of the
// previous kernel.  Both instructions are no-ops when the attribute is absent.
#define MTV_PDL_TRIGGER() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")
#define MTV_PDL_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")

namespace mtv {

// MTV_PDL is a bit mask of kernel classes launched with programmatic stream serialisation (default 5):
//   1 tensor-core tap-GEMM (follows a short apply kernel that triggers at entry: setup, TMEM allocation and the
//     weight TMA requests overlap that kernel, nothing is pre-launched more than one kernel deep)
//   2 tensor-core attention, 4 apply kernels, 8 split-K reductions, 16 everything else
extern int g_mtv_use_pdl;
enum { PDL_CLASS_CONV_TC = 1, PDL_CLASS_ATTN_TC = 2, PDL_CLASS_APPLY = 4, PDL_CLASS_REDUCE = 8, PDL_CLASS_OTHER = 16 };
inline bool mtv_pdl_enabled(int cls) { return (g_mtv_use_pdl & cls) != 0; }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kc(int cls, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = mtv_pdl_enabled(cls) ? 1 : 0;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  return launch_kc(PDL_CLASS_OTHER, kern, grid, block, smem, s, std::forward<Args>(args)...);
}

struct Geo {   // tri-plane token geometry of one pyramid level
  int res;     // xy plane: res x res
  int t;       // yt, xt planes: t rows x res cols
  int L;       // res*res + 2*t*res
};
__host__ __device__ inline Geo make_geo(int res, int t) { return Geo{res, t, res * res + 2 * t * res}; }
__host__ __device__ inline Geo geo_up(const Geo& g) { return make_geo(g.res * 2, g.t * 2); }     // finer
__host__ __device__ inline Geo geo_down(const Geo& g) { return make_geo(g.res / 2, g.t / 2); }   // coarser

enum { RS_NONE = 0, RS_UP2 = 1, RS_DOWN2 = 2 };

// One K-segment of a tap-GEMM: a (possibly two-source, channel-concatenated)
// activation, optionally passed through y = silu?(x*a + d) with per-(sample,
// norm-segment, channel) affine tables produced by the GroupNorm-statistics
// kernel, optionally resampled AFTER that transform (ResBlock h_upd, unet.py:181-182),
// then contracted with W[tap][cin][cout] over 1 or 3x3 taps.
struct KSeg {
  const float* src0; const float* src1;   // [B][Lsrc][C0], [B][Lsrc][C1]  (src1 == nullptr when C1 == 0)
  int C0, C1;
  const float* nrm_a; const float* nrm_d; // [B][nrm_nseg][C0+C1] or nullptr (raw)
  int nrm_nseg;                           // 3 = per plane, 1 = joint
  int silu;
  int resample;                           // RS_*: source geometry relative to the OUTPUT geometry
  int taps;                               // 1 or 9
  const float* w;                         // [taps][C0+C1][Cout]
};

struct ConvParams {
  KSeg seg[2]; int nsegs;
  int B; Geo geo;                         // output geometry
  int Cout;
  const float* bias;                      // [Cout] or nullptr
  const float* resid; int resid_mode;     // [B][Lr][Cout]; RS_NONE same geometry, RS_UP2 / RS_DOWN2 like KSeg.resample
  float* out; int out_chmajor;            // token-major [B][L][Cout], or channel-major [B][Cout][L]
  float* partial; int ksplit;             // split-K workspace [ksplit][B*L][Cout] when ksplit > 1
};

struct GnParams {
  const float* src0; const float* src1; int C0, C1;
  int B, L;
  int nseg; int seg_off[4];               // segment s = tokens [seg_off[s], seg_off[s+1])
  const float* gamma; const float* beta;  // [C]
  const float* film; int film_stride;     // row b = film + b*film_stride: (scale[C] | shift[C]); nullptr: none
  float* nrm_a; float* nrm_d;             // [B][nseg][C]
  double* sums;                           // [B][nseg][32][2], zero on entry, zero on exit
  unsigned int* counter;                  // [B][nseg], zero on entry, zero on exit
  int chunk_tokens;
};

struct AttnParams {
  const float* qkv;                       // [B][L][3C], channel order per head: q(D) k(D) v(D)  (unet.py:321)
  float* out;                             // [B][L][C]
  int B, L, C, heads;
  int nseg; int seg_off[4];
};

struct EmbParams {
  const int64_t* t; int B; int mc; int ted;        // timesteps, model_channels, time_embed_dim
  const float* freqs;                               // [mc/2]
  const float* w1; const float* b1;                 // [ted][mc], [ted]
  const float* w2; const float* b2;                 // [ted][ted], [ted]
  const float* wall; const float* ball; int J;      // all ResBlock emb_layers concatenated: [J][ted], [J]
  float* temb; float* h1; float* semb; float* film; // [B][mc], [B][ted], [B][ted], [B][J]
};

struct PackParams {
  const float* x; const float* cond; const float* image_cond; int64_t ic_len;
  int B; int cx, cc, ci;                  // channel counts (4, 8, 4)
  float* out;                             // [B][2048][cx+cc+ci]
  // tensor-core stem: instead of fp32 `out`, the split-bf16 operand [B][2048][cpad] (channels >= cx+cc+ci are zero)
  void* hi; void* lo; int cpad;
};

// ---- launchers (kernels_simt.cu) ----
cudaError_t launch_conv_simt(const ConvParams& P, cudaStream_t s);
int         conv_simt_pick_ksplit(const ConvParams& P, int num_sms);
cudaError_t launch_gn_stats(const GnParams& P, cudaStream_t s);
cudaError_t launch_attn_simt(const AttnParams& P, cudaStream_t s);
cudaError_t launch_emb(const EmbParams& P, cudaStream_t s);          // 3 kernels
cudaError_t launch_pack_in(const PackParams& P, cudaStream_t s);
cudaError_t launch_repack_conv(const float* src, float* dst, int Cout, int Cin, int taps, cudaStream_t s);
cudaError_t launch_add_vec(const float* a, const float* b, float* dst, int n, cudaStream_t s);
cudaError_t launch_ddim_step(float* img, const float* eps, const float* noise, int64_t n,
                             float sr, float srm1, float san, float c, float sigma, int last, cudaStream_t s);
cudaError_t launch_q_sample(const float* x0, const float* noise, int64_t n, float a, float b, float* out, cudaStream_t s);

// chunk I/O around the loop (kernels_chunkio.cu)
cudaError_t launch_io_prep_frames(const uint8_t* frames, int T, int H, int W, const int32_t* mask_row, int R, float* out, cudaStream_t s,
                                  bool vectorized_always = false);
cudaError_t launch_io_rasterize(const void* lm, int lm_is_f64, int T, int N, int dims, int WH, int flip, float* out, cudaStream_t s);
cudaError_t launch_io_frames_out(const float* dec, int B, int T, int H, int W, uint8_t* frames_u8, uint8_t* last_u8, float* next_ref, int Trep,
                                 cudaStream_t s);
cudaError_t launch_tok2ch(const float* src, float* dst, int B, int L, int C, cudaStream_t s);

}  // namespace mtv
