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

// Channel-sum slots of one GroupNorm group, spread over the lanes that own the group: item i = (channel ci = i % cpg of the
// group, slot k = i / cpg of the planes involved, in plane order).  All loads of a trip are issued before any is consumed (the
// slots live in L2: a serial loop over up to 16 slots per channel was a chain of L2 round trips); the assignment of items to
// lanes is fixed, so the summation order is deterministic.
struct CsumSrc { const double* cs0; const double* cs1; int C0, C1; };
__device__ __forceinline__ void csum_group_sum(const CsumSrc& S, const Geo& gs, int b, int p, bool joint, int grp, int cpg,
                                               int l, int nl /* lanes per group */, double& s, double& ss) {
  const int n0 = csum_nslots(gs, 0), n1 = csum_nslots(gs, 1);
  const int nk = joint ? n0 + 2 * n1 : csum_nslots(gs, p);
  const int nitems = cpg * nk;
  s = 0.0; ss = 0.0;
  for (int i0 = l; i0 < nitems; i0 += 4 * nl) {
    double2 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * nl;
      v[u] = make_double2(0.0, 0.0);
      if (i < nitems) {
        const int k = i / cpg, c = grp * cpg + (i - k * cpg);
        int pp = p, sl = k;
        if (joint) { pp = k < n0 ? 0 : (k < n0 + n1 ? 1 : 2); sl = k - (pp == 0 ? 0 : (pp == 1 ? n0 : n0 + n1)); }
        const double* cs; int Cs, cc;
        if (c < S.C0) { cs = S.cs0; Cs = S.C0; cc = c; } else { cs = S.cs1; Cs = S.C1; cc = c - S.C0; }
        v[u] = __ldcg(reinterpret_cast<const double2*>(cs + csum_at(gs, Cs, b, pp, sl, cc)));
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) { s += v[u].x; ss += v[u].y; }
  }
}

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
  // Fused front end (fe_x != nullptr; AttentionBlock*.norm + qkv, unet.py:234,251 / 281,297): every CTA first computes
  // GroupNorm(x) (statistics from the producer's channel-sum slots) and the qkv 1x1 projection of ITS <= 128 tokens for ITS
  // head on the tensor cores and writes Q / K / V^T (same split-bf16 layouts as the qkv GEMM epilogue); the CTAs of one
  // (sample, head) are one thread-block cluster and meet at a cluster barrier before the attention proper streams K / V.
  // No apply launch, no qkv GEMM launch.
  const float* fe_x;                 // [B][L][C] fp32 block input
  const double* fe_csum;             // its channel-sum slots (csum_at, geometry fe_geo)
  const float* fe_gamma; const float* fe_beta;   // norm affine [C]
  const float* fe_bias;              // qkv bias [3C], head-major q|k|v rows
  int fe_joint; Geo fe_geo;
  CUtensorMap tmWq_hi, tmWq_lo;      // qkv weight (C, 3C) split bf16, box (64, 3D)
  void* q_hi; void* q_lo; void* k_hi; void* k_lo; void* vt_hi; void* vt_lo;
};
cudaError_t launch_qkv_split(const QkvSplitParams& P, cudaStream_t s);
cudaError_t launch_attn_tc(const AttnTcParams& P, cudaStream_t s);
int         attn_fused_max_clusters(int D, int nqb);   // diagnostics

cudaError_t launch_apply_split(const ApplyParams& P, cudaStream_t s);
cudaError_t launch_repack_split_w(const float* src, void* hi, void* lo, int Cout, int Cin, int taps, cudaStream_t s);
cudaError_t launch_conv_tc(const TcConvParams& P, cudaStream_t s);
cudaError_t tc_debug_arm(long long* buf, unsigned int cap);
cudaError_t tc_debug_count(unsigned int* n);

}  // namespace mtv
