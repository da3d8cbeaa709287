// mtv_tc.cuh — parameter blocks of the tcgen05 / TMA kernels (kernels_tc.cu).
#pragma once
#include <cuda.h>          // CUtensorMap (type only; the encode entry point is fetched at run time)
#include <cuda_runtime.h>
#include "mtv_kernels.cuh"

namespace mtv {

// y = silu?(x*a + d) [+ resample] -> split-bf16 (hi, lo), token-major [B][L][C0+C1]
struct ApplyParams {
  const float* src0; const float* src1; int C0, C1;   // sources at the SOURCE geometry (see resample)
  const float* nrm_a; const float* nrm_d; int nrm_nseg;
  int silu; int resample;                              // RS_*: source geometry relative to `geo`
  // fused-GroupNorm mode (csum0 != nullptr): statistics come from the producers' per-channel sums
  const double* csum0; const double* csum1;            // channel-sum slots of the sources (csum_at), SOURCE geometry
  const float* gamma; const float* beta;               // [C0+C1]
  const float* film; int film_stride;                  // FiLM row b = film + b*stride: scale[C] | shift[C]; nullptr: none
  int joint;                                           // statistics over all three planes (AttentionBlock1D.norm)
  int chunk_tokens;                                    // tokens of one plane per CTA
  int B; Geo geo;                                      // OUTPUT geometry
  void* hi; void* lo;                                  // __nv_bfloat16 [B][L][C0+C1]
  void* raw_hi; void* raw_lo;                          // optional: split of the UN-normalised (but resampled) input as well
  // L2 prefetch of the consumer GEMM's (cold, HBM-resident) split weights while this short kernel runs
  const void* pf0; const void* pf1; unsigned long long pf_bytes;
};
// each CTA prefetches one 16-byte-aligned slice of [p, p+bytes) into L2 (fire and forget)
__device__ __forceinline__ void mtv_prefetch_slice(const void* p0, const void* p1, unsigned long long bytes, unsigned int cta,
                                                   unsigned int nctas) {
  if (!p0 || threadIdx.x != 0) return;
  unsigned long long per = ((bytes + nctas - 1) / nctas + 15ull) & ~15ull;
  const unsigned long long off = (unsigned long long)cta * per;
  if (off >= bytes) return;
  if (off + per > bytes) per = (bytes - off) & ~15ull;
  if (!per) return;
  const unsigned int sz = (unsigned int)per;
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"((const char*)p0 + off), "r"(sz) : "memory");
  if (p1) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"((const char*)p1 + off), "r"(sz) : "memory");
}

// Per-channel GroupNorm sums without atomics (order-deterministic): a producing tap-GEMM CTA leaves the (sum, sum of
// squares) of its rows of channel c in ONE slot per (sample, plane, row block), consumers add the slots in index order.
//   L > 128 : slot = 128-token tile within the plane (xy: res*res/128 slots; yt, xt: t*res/128)
//   L <= 128: one slot per (sample, plane)
// Layout: double [B][3][CSUM_NS(geo)][C][2].
__host__ __device__ inline int csum_ns(const Geo& g) { return g.L > 128 ? (g.res * g.res) / 128 : 1; }
__host__ __device__ inline int csum_nslots(const Geo& g, int p) { return g.L > 128 ? (p == 0 ? g.res * g.res : g.t * g.res) / 128 : 1; }
__host__ __device__ inline size_t csum_at(const Geo& g, int C, int b, int p, int slot, int c) {
  return ((((size_t)b * 3 + p) * csum_ns(g) + slot) * C + c) * 2;
}
__host__ __device__ inline size_t csum_elems(const Geo& g, int C, int B) { return (size_t)B * 3 * csum_ns(g) * C * 2; }

// D[B*L][Cout] = sum_tap A_tap[B*L][Cin] * W[tap][Cout][Cin]^T   (+bias, +residual)
struct TcConvParams {
  // L > 128 : taps==9: [0] xy plane (C,W,H,B), [1] yt|xt planes (C,W,H,2,B), 128-token boxes; taps==1: [0] = (C, B*L)
  // L <= 128: whole-plane boxes with a batch extent of 128/L samples (see tc_tile):
  //           taps==9: [0] (C,W,H,B), [1] (C,W,H,2,B); taps==1: [0],[1] = (C, L, B) with xy- / plane-sized boxes
  CUtensorMap tmA_hi[2], tmA_lo[2];
  CUtensorMap tmW_hi, tmW_lo;         // (Cin, taps*Cout), box (64, BN)
  // optional second K-segment: a 1x1 conv of another activation accumulated into the same tile
  // (the ResBlock skip_connection, unet.py:167,207); Cin2 == 0 when absent
  CUtensorMap tmA2_hi[2], tmA2_lo[2]; // like tmA_* with taps == 1
  CUtensorMap tmW2_hi, tmW2_lo;       // (Cin2, Cout)
  int Cin2; int bn;
  int taps, Cin, Cout, B; Geo geo;
  const float* bias; const float* resid; int resid_mode;
  float* out; float* partial; int ksplit;
  int out_cvalid;                     // > 0 (head conv): `out` is channel-major [B][out_cvalid][L] and only the first out_cvalid of the Cout = 64 columns are stored
  double* csum;                       // optional channel-sum slots of `out` (csum_at) for the next GroupNorm
  // split-K tickets (64-bit generation counters, never reset, zero once at allocation): word [tile * 32] belongs to output
  // tile `tile` = blockIdx.x * gridDim.y + blockIdx.y.  Required by ksplit > 1, which the host only selects for grids of at
  // most #SMs CTAs (every CTA of the launch is co-resident, so the in-kernel wait cannot deadlock).
  unsigned long long* sync;
  // qkv mode (qkv_heads > 0): instead of fp32 `out`, the epilogue writes the attention operands directly —
  // split-bf16 Q (pre-scaled by log2(e)/sqrt(D)) and K as [B*H][L][D], V^T as [B*H][D][L]  (see k_qkv_split)
  int qkv_heads;
  void* q_hi; void* q_lo; void* k_hi; void* k_lo; void* vt_hi; void* vt_lo;
  // L2 prefetch of the NEXT tap-GEMM's (HBM-cold) split weights, issued at kernel entry (one slice per CTA)
  const void* pf0; const void* pf1; unsigned long long pf_bytes;
  // bulk (TMA) store of the epilogue tile: map of `out` [B*L][Cout] (or of `partial` [ksplit*B*L][Cout]), fp32, box 32 x 32
  CUtensorMap tmOut; int tma_store;
  int dbg_skip;   // diagnostics only (MTV_TC_DBG_SKIP; results are garbage): 1 no A-tile fetch, 2 no W-tile fetch, 4 no channel sums, 8 no output stores
};

// qkv fp32 [B][L][3C] -> split-bf16 Q (pre-scaled), K [B*H][L][D] and V^T [B*H][D][L]
struct QkvSplitParams {
  const float* qkv; int B, L, C, heads;
  void* q_hi; void* q_lo; void* k_hi; void* k_lo; void* vt_hi; void* vt_lo;
};
// softmax(Q K^T) V per (sample, head, segment) on tcgen05
struct AttnTcParams {
  CUtensorMap tmQ_hi, tmQ_lo;     // (D, B*H*L) box (D, 128)
  CUtensorMap tmK_hi, tmK_lo;     // (D, B*H*L) box (D, 64)
  CUtensorMap tmV_hi, tmV_lo;     // (L, B*H*D) box (64, D)
  float* out;                     // [B][L][C] fp32, or nullptr when out_hi / out_lo are given
  void* out_hi; void* out_lo;     // optional split-bf16 [B][L][C]: the A operand of the proj_out GEMM
  const void* pf0; const void* pf1; unsigned long long pf_bytes;   // L2 prefetch of the proj_out weights
  int B, L, C, heads;
  int nseg; int seg_off[4];
  int dbg_skip;   // diagnostics only (MTV_ATTN_DBG_SKIP; results are garbage): 1 no exp2, 2 no P stores, 4 no row-max exchange, 8 no O fold
  int kv_split;   // > 1: that many CTAs (one cluster) share a query tile, each streaming 1/kv_split of the key blocks (small batches)
};
cudaError_t launch_qkv_split(const QkvSplitParams& P, cudaStream_t s);
cudaError_t launch_attn_tc(const AttnTcParams& P, cudaStream_t s);

cudaError_t launch_apply_split(const ApplyParams& P, cudaStream_t s);
cudaError_t launch_repack_split_w(const float* src, void* hi, void* lo, int Cout, int Cin, int taps, cudaStream_t s);
cudaError_t launch_repack_split_w_pad(const float* src, void* hi, void* lo, int Cout, int Cin, int CoutPad, int CinPad, int taps, cudaStream_t s);
cudaError_t launch_conv_tc(const TcConvParams& P, cudaStream_t s);
cudaError_t tc_debug_arm(long long* buf, unsigned int cap);
cudaError_t tc_debug_count(unsigned int* n);

}  // namespace mtv
