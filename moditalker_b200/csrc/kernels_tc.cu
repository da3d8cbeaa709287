// kernels_tc.cu — tcgen05 / TMA tensor-core kernels (sm_100a only).
//
//   k_apply_split : y = silu?(x*a + d) (GroupNorm affine [+FiLM] produced by k_gn_stats),
//                   optional nearest-x2 / avgpool-2x2 resample, optional channel concat of
//                   two sources, written as a SPLIT-BF16 pair (hi = bf16(y), lo = bf16(y - hi))
//                   token-major [B][L][C] — the A operand of the tap-GEMM below.
//   k_conv_tc     : implicit-GEMM 3x3 / 1x1 convolution on the 5th-gen tensor cores.
//                   D[128 tokens x BN couts] (fp32, TMEM) += A_tap[128 x 64] * W_tap[BN x 64]^T
//                   for every tap and 64-channel chunk.  A tiles are TMA *spatial boxes* of the
//                   token-major activation (one shifted box per tap; the conv's zero padding is
//                   TMA out-of-bounds fill, so planes never bleed into each other), W tiles are
//                   TMA boxes of the pre-split weights.  fp32-class accuracy from three bf16
//                   products per product:  A_hi*W_hi + A_lo*W_hi + A_hi*W_lo  (error ~2^-16 per
//                   product instead of bf16's 2^-8; the MToV parity bar of 1e-3 rules out plain
//                   bf16 and leaves single-pass TF32 no margin over a 50-step trajectory,
//                   SURVEY.md §7), issued as two MMAs per k-step (stacked N: [W_hi; W_lo]).
//                   Warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc),
//                   warps 2-5 = epilogue (tcgen05.ld -> +bias +residual [+GroupNorm sums | qkv
//                   operand split] -> HBM), specialised at compile time (EPI).
//
// Reference semantics: ResBlock._forward conv3x3s (unet.py:134,159,178-207), the 1x1
// qkv / proj_out convs of AttentionBlock* (unet.py:234,242,251-254,297-300).
#include "mtv_kernels.cuh"
#include "mtv_tc.cuh"

#include <cuda_bf16.h>

namespace mtv {

// ------------------------------------------------------------------ apply + split
// x * sigmoid(x) with a fast (2-ulp) division: the IEEE division's slow-path call bloats the unrolled producers
__device__ __forceinline__ float silu_tc(float v) { return __fdividef(v, 1.0f + __expf(-v)); }

__device__ __forceinline__ void tc_decode_tok(const Geo& g, int tok, int& p, int& y, int& x) {
  const int nxy = g.res * g.res;
  if (tok < nxy) { p = 0; y = tok / g.res; x = tok - y * g.res; }
  else { int r = tok - nxy; const int np = g.t * g.res; p = 1; if (r >= np) { p = 2; r -= np; } y = r / g.res; x = r - y * g.res; }
}
__device__ __forceinline__ int tc_plane_off(const Geo& g, int p) { return p == 0 ? 0 : g.res * g.res + (p - 1) * g.t * g.res; }

__device__ __forceinline__ void tc_store_split(const float4& v, void* hi_base, void* lo_base, size_t o) {
  const float f[4] = {v.x, v.y, v.z, v.w};
  __nv_bfloat16 hi[4], lo[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    hi[i] = __float2bfloat16_rn(f[i]);
    lo[i] = __float2bfloat16_rn(f[i] - __bfloat162float(hi[i]));
  }
  *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(hi_base) + o) = *reinterpret_cast<const uint2*>(hi);
  *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(lo_base) + o) = *reinterpret_cast<const uint2*>(lo);
}

__global__ void __launch_bounds__(256) k_apply_split(const __grid_constant__ ApplyParams P) {
  MTV_PDL_TRIGGER();
  mtv_prefetch_slice(P.pf0, P.pf1, P.pf_bytes, blockIdx.x, gridDim.x);
  MTV_PDL_WAIT();
  const int C = P.C0 + P.C1;
  const int cq = C >> 2;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)P.B * P.geo.L * cq;
  if (idx >= total) return;
  const int c = (int)(idx % cq) * 4;
  const size_t m = idx / cq;
  const int b = (int)(m / P.geo.L), tok = (int)(m - (size_t)b * P.geo.L);
  int p, y, x; tc_decode_tok(P.geo, tok, p, y, x);
  const float* src; int Cs, cc;
  if (c < P.C0) { src = P.src0; Cs = P.C0; cc = c; } else { src = P.src1; Cs = P.C1; cc = c - P.C0; }
  float4 na = make_float4(1.f, 1.f, 1.f, 1.f), nd = make_float4(0.f, 0.f, 0.f, 0.f);
  if (P.nrm_a) {
    const size_t ni = ((size_t)b * P.nrm_nseg + (P.nrm_nseg == 3 ? p : 0)) * C + c;
    na = __ldg(reinterpret_cast<const float4*>(P.nrm_a + ni));
    nd = __ldg(reinterpret_cast<const float4*>(P.nrm_d + ni));
  }
  auto xf = [&](float4 v) {
    if (P.nrm_a) { v.x = fmaf(v.x, na.x, nd.x); v.y = fmaf(v.y, na.y, nd.y); v.z = fmaf(v.z, na.z, nd.z); v.w = fmaf(v.w, na.w, nd.w); }
    if (P.silu) { v.x = silu_tc(v.x); v.y = silu_tc(v.y); v.z = silu_tc(v.z); v.w = silu_tc(v.w); }
    return v;
  };
  float4 v, rw;
  if (P.resample == RS_NONE) {
    rw = __ldg(reinterpret_cast<const float4*>(src + ((size_t)b * P.geo.L + tok) * Cs + cc));
    v = xf(rw);
  } else if (P.resample == RS_UP2) {
    const Geo gs = geo_down(P.geo);
    const int ts = tc_plane_off(gs, p) + (y >> 1) * gs.res + (x >> 1);
    rw = __ldg(reinterpret_cast<const float4*>(src + ((size_t)b * gs.L + ts) * Cs + cc));
    v = xf(rw);
  } else {
    const Geo gs = geo_up(P.geo);
    const int ts = tc_plane_off(gs, p) + (2 * y) * gs.res + 2 * x;
    const float* q = src + ((size_t)b * gs.L + ts) * Cs + cc;
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(q)), w1 = __ldg(reinterpret_cast<const float4*>(q + Cs));
    const float4 w2 = __ldg(reinterpret_cast<const float4*>(q + (size_t)gs.res * Cs));
    const float4 w3 = __ldg(reinterpret_cast<const float4*>(q + (size_t)(gs.res + 1) * Cs));
    const float4 v0 = xf(w0), v1 = xf(w1), v2 = xf(w2), v3 = xf(w3);
    v.x = 0.25f * ((v0.x + v1.x) + (v2.x + v3.x)); v.y = 0.25f * ((v0.y + v1.y) + (v2.y + v3.y));
    v.z = 0.25f * ((v0.z + v1.z) + (v2.z + v3.z)); v.w = 0.25f * ((v0.w + v1.w) + (v2.w + v3.w));
    rw.x = 0.25f * ((w0.x + w1.x) + (w2.x + w3.x)); rw.y = 0.25f * ((w0.y + w1.y) + (w2.y + w3.y));
    rw.z = 0.25f * ((w0.z + w1.z) + (w2.z + w3.z)); rw.w = 0.25f * ((w0.w + w1.w) + (w2.w + w3.w));
  }
  const size_t o = m * C + c;
  tc_store_split(v, P.hi, P.lo, o);
  if (P.raw_hi) tc_store_split(rw, P.raw_hi, P.raw_lo, o);
}

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

// One work unit of the fused GroupNorm-finalise + apply: `chunk_tokens` tokens starting at chunk `chunk_idx` of plane p of
// sample b, by a 256-thread CTA.  s_aff: 2*C floats, s_mean / s_rstd: 32 doubles each (all CTA-shared scratch).
__device__ __forceinline__ void apply_norm_unit(const ApplyParams& P, float* s_aff, double* s_mean, double* s_rstd,
                                                int chunk_idx, int p, int b) {
  const int C = P.C0 + P.C1, cpg = C / 32;
  const Geo g = P.geo;
  const int plane_tokens = p == 0 ? g.res * g.res : g.t * g.res;
  const int t0 = chunk_idx * P.chunk_tokens;
  if (t0 >= plane_tokens) return;
  const int t1 = min(plane_tokens, t0 + P.chunk_tokens);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const Geo gs = P.resample == RS_NONE ? g : (P.resample == RS_UP2 ? geo_down(g) : geo_up(g));
  const double cnt = (double)cpg * (P.joint ? (double)gs.L : (double)(p == 0 ? gs.res * gs.res : gs.t * gs.res));
  const int cq = C >> 2;
  const int poff = tc_plane_off(g, p);
  const int total = (t1 - t0) * cq;
  // item = 4 channels of one token.  Two items per trip with both loads issued before either is
  // consumed: the loop is a chain of L2 round trips, not arithmetic.
  struct Item { const float* q; size_t o; int c, Cs; bool ok; };
  auto locate = [&](int idx) {
    Item it; it.ok = idx < total;
    const int id = it.ok ? idx : 0;
    const int tl = t0 + id / cq; it.c = (id % cq) * 4;
    const int y = tl / g.res, x = tl - y * g.res;
    const float* src; int cc;
    if (it.c < P.C0) { src = P.src0; it.Cs = P.C0; cc = it.c; } else { src = P.src1; it.Cs = P.C1; cc = it.c - P.C0; }
    int ts = poff + tl;
    if (P.resample == RS_UP2) ts = tc_plane_off(gs, p) + (y >> 1) * gs.res + (x >> 1);
    else if (P.resample == RS_DOWN2) ts = tc_plane_off(gs, p) + (2 * y) * gs.res + 2 * x;
    it.q = src + ((size_t)b * gs.L + ts) * it.Cs + cc;
    it.o = ((size_t)b * g.L + poff + tl) * C + it.c;
    return it;
  };
  auto fetch = [&](const Item& it, float4& w0, float4& w1, float4& w2, float4& w3) {
    if (!it.ok) return;
    w0 = __ldg(reinterpret_cast<const float4*>(it.q));
    if (P.resample == RS_DOWN2) {
      w1 = __ldg(reinterpret_cast<const float4*>(it.q + it.Cs));
      w2 = __ldg(reinterpret_cast<const float4*>(it.q + (size_t)gs.res * it.Cs));
      w3 = __ldg(reinterpret_cast<const float4*>(it.q + (size_t)(gs.res + 1) * it.Cs));
    }
  };
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  Item i0 = locate(tid), i1 = locate(tid + 256);
  float4 a0 = z4, a1 = z4, a2 = z4, a3 = z4, b0 = z4, b1 = z4, b2 = z4, b3 = z4;
  fetch(i0, a0, a1, a2, a3); fetch(i1, b0, b1, b2, b3);   // first trip's activations: in flight under the statistics pass
  // gamma / beta / FiLM rows are read once per step and have left L2 by then: issue their (DRAM-latency)
  // loads first so they overlap the statistics pass instead of following it
  float pg[8], pb[8], psc[8], psh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = tid + 256 * k;
    pg[k] = pb[k] = psc[k] = psh[k] = 0.f;
    if (c < C) {
      pg[k] = __ldg(P.gamma + c); pb[k] = __ldg(P.beta + c);
      if (P.film) { const float* f = P.film + (size_t)b * P.film_stride; psc[k] = __ldg(f + c); psh[k] = __ldg(f + C + c); }
    }
  }
  {
    // 32 groups in one pass: warp w owns groups 4w..4w+3, 8 lanes per group (one load latency, not four)
    const int grp = warp * 4 + (lane >> 3);
    double s, ss;
    const CsumSrc CS{P.csum0, P.csum1, P.C0, P.C1};
    csum_group_sum(CS, gs, b, p, P.joint != 0, grp, cpg, lane & 7, 8, s, ss);     // statistics live at the SOURCE geometry
#pragma unroll
    for (int off = 4; off > 0; off >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, off); ss += __shfl_xor_sync(0xffffffffu, ss, off); }
    if ((lane & 7) == 0) {
      const double mean = s / cnt;
      double var = ss / cnt - mean * mean; var = var < 0.0 ? 0.0 : var;
      s_mean[grp] = mean; s_rstd[grp] = rsqrt(var + 1e-5);
    }
  }
  __syncthreads();
  float* sa = s_aff; float* sd = s_aff + C;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = tid + 256 * k;
    if (c < C) {
      const int grp = c / cpg;
      double a = s_rstd[grp] * (double)pg[k];
      double d = (double)pb[k] - s_mean[grp] * a;
      if (P.film) { const double sc = 1.0 + (double)psc[k]; a *= sc; d = d * sc + (double)psh[k]; }
      sa[c] = (float)a; sd[c] = (float)d;
    }
  }
  auto finish = [&](const Item& it, const float4 w0, const float4 w1, const float4 w2, const float4 w3) {
    const float4 na = *reinterpret_cast<const float4*>(sa + it.c), nd = *reinterpret_cast<const float4*>(sd + it.c);
    auto xf = [&](float4 v) {
      v.x = fmaf(v.x, na.x, nd.x); v.y = fmaf(v.y, na.y, nd.y); v.z = fmaf(v.z, na.z, nd.z); v.w = fmaf(v.w, na.w, nd.w);
      if (P.silu) { v.x = silu_tc(v.x); v.y = silu_tc(v.y); v.z = silu_tc(v.z); v.w = silu_tc(v.w); }
      return v;
    };
    float4 v, rw;
    if (P.resample != RS_DOWN2) {
      rw = w0; v = xf(w0);
    } else {
      const float4 v0 = xf(w0), v1 = xf(w1), v2 = xf(w2), v3 = xf(w3);
      v.x = 0.25f * ((v0.x + v1.x) + (v2.x + v3.x)); v.y = 0.25f * ((v0.y + v1.y) + (v2.y + v3.y));
      v.z = 0.25f * ((v0.z + v1.z) + (v2.z + v3.z)); v.w = 0.25f * ((v0.w + v1.w) + (v2.w + v3.w));
      rw.x = 0.25f * ((w0.x + w1.x) + (w2.x + w3.x)); rw.y = 0.25f * ((w0.y + w1.y) + (w2.y + w3.y));
      rw.z = 0.25f * ((w0.z + w1.z) + (w2.z + w3.z)); rw.w = 0.25f * ((w0.w + w1.w) + (w2.w + w3.w));
    }
    tc_store_split(v, P.hi, P.lo, it.o);
    if (P.raw_hi) tc_store_split(rw, P.raw_hi, P.raw_lo, it.o);
  };
  __syncthreads();
  for (int idx = tid; idx < total; idx += 512) {
    const Item c0 = i0, c1 = i1;
    const float4 x0 = a0, x1 = a1, x2 = a2, x3 = a3, y0 = b0, y1 = b1, y2 = b2, y3 = b3;
    if (idx + 512 < total) {                       // next trip's loads go out before this trip's math
      i0 = locate(idx + 512); i1 = locate(idx + 768);
      fetch(i0, a0, a1, a2, a3); fetch(i1, b0, b1, b2, b3);
    }
    finish(c0, x0, x1, x2, x3);
    if (c1.ok) finish(c1, y0, y1, y2, y3);
  }
}

// GroupNorm32 finalise + apply in ONE kernel: group statistics come from the per-channel sums the
// producing tap-GEMM accumulated in its epilogue (csum[b][plane][c][2], fp64), so no separate
// statistics pass reads the activation again.  CTA = (token chunk, plane, sample); prologue turns
// the sums of that (sample, plane | all planes) into the per-channel affine in shared memory
// (FiLM folded in), then the body is k_apply_split's.
__global__ void __launch_bounds__(256) k_apply_norm_split(const __grid_constant__ ApplyParams P) {
  MTV_PDL_TRIGGER();
  mtv_prefetch_slice(P.pf0, P.pf1, P.pf_bytes, blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z),
                     gridDim.x * gridDim.y * gridDim.z);
  MTV_PDL_WAIT();
  extern __shared__ float s_aff[];                 // a[C] | d[C]
  __shared__ double s_mean[32], s_rstd[32];
  apply_norm_unit(P, s_aff, s_mean, s_rstd, (int)blockIdx.x, (int)blockIdx.y, (int)blockIdx.z);
}

cudaError_t launch_apply_split(const ApplyParams& P, cudaStream_t s) {
  if (P.csum0) {
    const int C = P.C0 + P.C1;
    const int maxp = P.geo.res * P.geo.res;
    dim3 grid((maxp + P.chunk_tokens - 1) / P.chunk_tokens, 3, P.B);
    const size_t smem = (size_t)C * 2 * sizeof(float);
    { cudaError_t le_ = launch_kc(PDL_CLASS_APPLY, k_apply_norm_split, dim3(grid), dim3(256), (size_t)(smem), s, P); if (le_ != cudaSuccess) return le_; }
    return cudaGetLastError();
  }
  const size_t total = (size_t)P.B * P.geo.L * ((P.C0 + P.C1) / 4);
  { cudaError_t le_ = launch_kc(PDL_CLASS_APPLY, k_apply_split, dim3((unsigned)((total + 255) / 256)), dim3(256), (size_t)(0), s, P); if (le_ != cudaSuccess) return le_; }
  return cudaGetLastError();
}

// fp32 [rows][cols] -> split bf16 pair (weights, once at load time)
__global__ void k_split_bf16(const float* __restrict__ src, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = src[i];
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  hi[i] = h; lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
}
// PyTorch conv weight [Cout][Cin][taps] -> K-major split-bf16 [taps][Cout][Cin]
__global__ void k_repack_split_w(const float* __restrict__ src, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                 int Cout, int Cin, int taps) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)Cout * Cin * taps) return;
  const int ci = (int)(i % Cin);
  const int co = (int)((i / Cin) % Cout);
  const int tp = (int)(i / ((size_t)Cin * Cout));
  const float v = src[((size_t)co * Cin + ci) * taps + tp];
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  hi[i] = h; lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
}
// same, zero-padded to [taps][CoutPad][CinPad] (stem conv: 16 input channels -> one 64-channel K chunk; head conv: 4 output
// channels -> one 64-column tile)
__global__ void k_repack_split_w_pad(const float* __restrict__ src, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                     int Cout, int Cin, int CoutPad, int CinPad, int taps) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)CoutPad * CinPad * taps) return;
  const int ci = (int)(i % CinPad);
  const int co = (int)((i / CinPad) % CoutPad);
  const int tp = (int)(i / ((size_t)CinPad * CoutPad));
  const float v = (ci < Cin && co < Cout) ? src[((size_t)co * Cin + ci) * taps + tp] : 0.0f;
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  hi[i] = h; lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
}
cudaError_t launch_repack_split_w_pad(const float* src, void* hi, void* lo, int Cout, int Cin, int CoutPad, int CinPad, int taps, cudaStream_t s) {
  const size_t n = (size_t)CoutPad * CinPad * taps;
  k_repack_split_w_pad<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, Cout, Cin, CoutPad, CinPad, taps);
  return cudaGetLastError();
}
cudaError_t launch_repack_split_w(const float* src, void* hi, void* lo, int Cout, int Cin, int taps, cudaStream_t s) {
  const size_t n = (size_t)Cout * Cin * taps;
  k_repack_split_w<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, Cout, Cin, taps);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bare retry loop for the hot waits (MMA issuer, epilogue): extra instructions in it are not free (kernels_attn_tc.cu).  The hang
// guard lives in the TMA producer warp (mbar_wait_guard): every deadlock of this kernel also blocks that warp (it waits for ring
// slots and, at the end, for the accumulator), whose bounded wait then traps — a pipeline bug surfaces as a launch failure, never
// as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void mbar_wait_guard(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok, spins = 0;
  long long t0 = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    if (!ok && (++spins & 0x3ffu) == 0) {
      const long long t = clock64();
      if (t0 == 0) t0 = t; else if (t - t0 > (1ll << 32)) __trap();     // ~2 s
    }
  } while (!ok);
}
__device__ __forceinline__ void mbar_inval(uint64_t* bar) {
  asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
               ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, 128-byte-swizzled operand tile: rows of 64 bf16 (128 B), 8-row groups 1024 B apart.
// Bit layout: cute/arch/mma_sm100_desc.hpp SmemDescriptor (start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout SWIZZLE_128B=2 [61,64)).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;                  // LBO (ignored for swizzled K-major; canonical value 1)
  d |= (uint64_t)(1024 >> 4) << 32;        // SBO = 8 rows * 128 B
  d |= (uint64_t)1 << 46;                  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                  // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major (InstrDescriptor bit-fields).
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// One lane of a converged warp.  The MMA issuer runs its loop in ALL lanes (uniform control flow) and elects only around the
// tcgen05 instructions: descriptors computed in warp-uniform code live in uniform registers, whereas inside an `if (lane == 0)`
// region ptxas wraps every tcgen05.mma in an ELECT / R2UR.BROADCAST waterfall (~115 cycles per MMA measured: the main loop was
// issue-bound at ~920 cycles per K-iteration with no operand loads at all).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ------------------------------------------------------------------ optional in-kernel timing (diagnostics)
// When armed through mtv_debug_tc_timing(), every k_conv_tc CTA appends one 16-word record of
// clock64() stamps of its pipeline phases; scripts/tc_timing.py turns them into a breakdown.
__device__ long long* g_tc_dbg = nullptr;
__device__ unsigned int g_tc_dbg_cap = 0;
__device__ unsigned int g_tc_dbg_count = 0;
cudaError_t tc_debug_arm(long long* buf, unsigned int cap) {
  unsigned int zero = 0;
  cudaError_t e = cudaMemcpyToSymbol(g_tc_dbg, &buf, sizeof(buf));
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_tc_dbg_cap, &cap, sizeof(cap));
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_tc_dbg_count, &zero, sizeof(zero));
  return e;
}
cudaError_t tc_debug_count(unsigned int* n) { return cudaMemcpyFromSymbol(n, g_tc_dbg_count, sizeof(*n)); }
__device__ __forceinline__ long long gtime_ns() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256): a lane that owns a 128-byte row segment moves it in
// four full 32-byte sectors instead of eight half sectors — the row-per-lane epilogue is LSU-transaction bound
__device__ __forceinline__ void st_global_v8(float* p, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
__device__ __forceinline__ void ld_global_nc_v8(const float* p, float* v) {
  asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
}

// ------------------------------------------------------------------ the tap-GEMM
// warps: 0 TMA, 1 MMA, 2-9 epilogue (two warps per TMEM lane quarter, each taking every other 32-column chunk: the
// row-per-lane epilogue is instruction-issue bound and was the longest phase of the small-batch GEMMs with four warps)
constexpr int TC_BM = 128, TC_BK = 64, TC_THREADS = 320;
__host__ __device__ constexpr int tc_stage_bytes(int BN) { return 2 * TC_BM * 128 + 2 * BN * 128; }
__host__ __device__ constexpr int tc_stages(int BN) { return BN == 64 ? 4 : 3; }
__host__ __device__ constexpr int tc_smem_bytes(int BN) { return tc_stages(BN) * tc_stage_bytes(BN) + 1024; }

// Tile geometry.  Levels with >= 128 tokens per sample (L = 2048, 512): a tile is 128
// consecutive tokens of one plane of one sample.  Small levels (L = 128, 32): a tile is
// spt = 128/L whole samples, rows ordered plane-major [xy of the spt samples | yt | xt] so each
// plane is ONE TMA box with a batch extent (out-of-range samples are zero-filled).
struct TcTile {
  bool small; int spt; int b0; int tok0;
  int nxy, npl;
  int sh_xy, sh_pl, sh_res;     // log2 of nxy, npl, res (all powers of two: res = 32 >> level, t = 16 >> level)
};
__device__ __forceinline__ TcTile tc_tile(const Geo& g, int tile) {
  TcTile t;
  t.small = g.L <= TC_BM; t.nxy = g.res * g.res; t.npl = g.t * g.res;
  t.sh_xy = 31 - __clz(t.nxy); t.sh_pl = 31 - __clz(t.npl); t.sh_res = 31 - __clz(g.res);
  if (t.small) { t.spt = TC_BM / g.L; t.b0 = tile * t.spt; t.tok0 = 0; }
  else { const int tps = g.L / TC_BM; t.spt = 1; t.b0 = tile / tps; t.tok0 = (tile - t.b0 * tps) * TC_BM; }
  return t;
}
// tile row -> (sample, token within the sample); division-free (hot in the epilogue)
__device__ __forceinline__ void tc_row_map(const TcTile& t, int row, int& b, int& tok) {
  if (!t.small) { b = t.b0; tok = t.tok0 + row; return; }
  const int e1 = t.spt << t.sh_xy, e2 = e1 + (t.spt << t.sh_pl);
  if (row < e1) { const int s = row >> t.sh_xy; b = t.b0 + s; tok = row & (t.nxy - 1); }
  else if (row < e2) { const int r = row - e1; b = t.b0 + (r >> t.sh_pl); tok = t.nxy + (r & (t.npl - 1)); }
  else { const int r = row - e2; b = t.b0 + (r >> t.sh_pl); tok = t.nxy + t.npl + (r & (t.npl - 1)); }
}
// token -> (plane, y, x) with shifts
__device__ __forceinline__ void tc_decode_fast(const TcTile& t, int tok, int& p, int& y, int& x) {
  int r = tok; p = 0;
  if (tok >= t.nxy) { r = tok - t.nxy; p = 1; if (r >= t.npl) { r -= t.npl; p = 2; } }
  y = r >> t.sh_res; x = r & ((1 << t.sh_res) - 1);
}

// one 128-row A operand tile (hi and lo) for K-chunk c0 of tap `tap`
__device__ __forceinline__ void tc_load_A(const Geo& g, const TcTile& t, const CUtensorMap* mh, const CUtensorMap* ml, int taps,
                                          int tap, int c0, uint32_t sA_hi, uint32_t sA_lo, uint32_t fb) {
  if (!t.small) {
    if (taps == 1) {
      const int row = t.b0 * g.L + t.tok0;
      tma_load_2d(sA_hi, &mh[0], fb, c0, row);
      tma_load_2d(sA_lo, &ml[0], fb, c0, row);
      return;
    }
    const int dy = tap / 3 - 1, dx = tap - (tap / 3) * 3 - 1;
    if (t.tok0 < t.nxy) {
      const int y0 = t.tok0 / g.res;
      tma_load_4d(sA_hi, &mh[0], fb, c0, dx, y0 + dy, t.b0);
      tma_load_4d(sA_lo, &ml[0], fb, c0, dx, y0 + dy, t.b0);
    } else {
      const int r = t.tok0 - t.nxy;
      const int pl = r / t.npl, y0 = (r - pl * t.npl) / g.res;
      tma_load_5d(sA_hi, &mh[1], fb, c0, dx, y0 + dy, pl, t.b0);
      tma_load_5d(sA_lo, &ml[1], fb, c0, dx, y0 + dy, pl, t.b0);
    }
    return;
  }
  const uint32_t o1 = (uint32_t)(t.spt * t.nxy) * 128u, o2 = o1 + (uint32_t)(t.spt * t.npl) * 128u;
  if (taps == 1) {   // (C, L, B) maps: [0] box = (64, nxy, spt), [1] box = (64, npl, spt)
    tma_load_3d(sA_hi, &mh[0], fb, c0, 0, t.b0);
    tma_load_3d(sA_lo, &ml[0], fb, c0, 0, t.b0);
    tma_load_3d(sA_hi + o1, &mh[1], fb, c0, t.nxy, t.b0);
    tma_load_3d(sA_lo + o1, &ml[1], fb, c0, t.nxy, t.b0);
    tma_load_3d(sA_hi + o2, &mh[1], fb, c0, t.nxy + t.npl, t.b0);
    tma_load_3d(sA_lo + o2, &ml[1], fb, c0, t.nxy + t.npl, t.b0);
  } else {           // [0] (C,W,H,B) box (64,res,res,spt); [1] (C,W,H,2,B) box (64,res,t,1,spt)
    const int dy = tap / 3 - 1, dx = tap - (tap / 3) * 3 - 1;
    tma_load_4d(sA_hi, &mh[0], fb, c0, dx, dy, t.b0);
    tma_load_4d(sA_lo, &ml[0], fb, c0, dx, dy, t.b0);
    tma_load_5d(sA_hi + o1, &mh[1], fb, c0, dx, dy, 0, t.b0);
    tma_load_5d(sA_lo + o1, &ml[1], fb, c0, dx, dy, 0, t.b0);
    tma_load_5d(sA_hi + o2, &mh[1], fb, c0, dx, dy, 1, t.b0);
    tma_load_5d(sA_lo + o2, &ml[1], fb, c0, dx, dy, 1, t.b0);
  }
}


// ------------------------------------------------------------------ TMA store of an epilogue sub-tile
// Each epilogue warp stages its 32 rows x 32 fp32 columns in (idle) operand-ring memory, 128B-swizzled, and ONE elected lane
// writes them with a bulk tensor store: whole 128-byte lines per request instead of 32 scattered 32-byte row segments per
// st.global instruction (which cost ~1800 cycles per chunk, profiles/r01_s2_mainloop_skip.md).
__device__ __forceinline__ void tma_store_2d(const void* tmap, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(tmap), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// lane's row of 32 floats -> row `lane` of the warp's staging tile at `sbase` (1024-byte aligned)
__device__ __forceinline__ void tc_stage_row(uint32_t sbase, int lane, const float (&v)[32]) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint32_t a = sbase + (uint32_t)lane * 128u + (uint32_t)((c ^ (lane & 7)) << 4);
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v[4 * c]), "f"(v[4 * c + 1]), "f"(v[4 * c + 2]), "f"(v[4 * c + 3]) : "memory");
  }
}

__device__ __forceinline__ float4 ld_shared_v4f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

// ------------------------------------------------------------------ co-resident grid synchronisation
// Split-K and the fused consumer apply need CTAs of ONE launch to wait for each other.  That is safe because the host only
// enables them for grids of at most #SMs CTAs at one CTA per SM (every CTA is resident once its predecessors in the stream
// have drained; a programmatically launched successor cannot start before every CTA here has triggered).  Counters are 64-bit
// generation counters that are never reset: the n participants of one launch all arrive before any participant of the next
// launch of the same op (kernel boundary), so the generation is old / n.  Bounded spin: a bug traps instead of hanging.
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// Generation of the launch in flight: read by ONE thread after the grid dependency resolved (every earlier launch of this op
// has completed) and BEFORE this CTA arrives, so the counter lies in [gen*n, (gen+1)*n): the wait target is (gen+1)*n.
__device__ __forceinline__ unsigned long long tc_gen_target(const unsigned long long* ctr, unsigned int n) {
  return (ld_acquire_u64(ctr) / n + 1ull) * n;
}
// ONE thread per CTA, after a CTA-level barrier behind which all of the CTA's global writes were issued.  red.release orders
// those writes (cumulatively) before the arrival without a separate fence or a returned value; waiters poll with acquire
// loads and back off, so the arrivals of the late CTAs are not queued behind a storm of polls on one L2 slice.
__device__ __forceinline__ void tc_gen_barrier(unsigned long long* ctr, unsigned long long target) {
  asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(ctr), "l"(1ull) : "memory");
  unsigned int spins = 0;
  long long t0 = 0;
  while (ld_acquire_u64(ctr) < target) {
    __nanosleep(40);
    if ((++spins & 0xffu) == 0) {
      const long long t = clock64();
      if (t0 == 0) t0 = t; else if (t - t0 > (1ll << 32)) __trap();     // ~2 s
    }
  }
}
constexpr int TC_SYNC_STRIDE = 32;     // 64-bit words between two counters (256 B: distinct L2 lines / slices)
#define TC_EPI_BAR() asm volatile("bar.sync 1, 256;" ::: "memory")      // the eight epilogue warps

// Shared-memory carve-up of the (idle once the accumulator is complete) operand ring during the epilogue
constexpr uint32_t TC_EPI_STAGE_OFF = 0;            // per (warp, chunk) staging tiles of 32 rows x 32 fp32, 128B-swizzled: <= 64 KB
constexpr uint32_t TC_EPI_CS_OFF = 65536;           // float [16 row groups][BN][2]: channel sums per aligned group of 8 tile rows
static_assert(TC_EPI_CS_OFF + 16384 <= 3 * tc_stage_bytes(128) && TC_EPI_CS_OFF + 16384 <= 4 * tc_stage_bytes(64), "epilogue scratch fits the ring");

// (sum, sum of squares) of a 32-column chunk over each aligned group of 8 tile rows -> dst[row group][column][2] (floats;
// `ld` floats between row groups).  8 rows never straddle a (sample, plane) boundary at any level (plane sizes are multiples of
// 8 tokens).  Butterfly transpose-reduce: after the three exchange steps lane l holds columns ((l & 7) << 2) + {0..3} of the
// warp's row group l >> 3.  Rows beyond the batch contribute zeros.
__device__ __forceinline__ void tc_csum_chunk(float (&v)[32], bool live, int lane, float* dst, int ld) {
  float q[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) { if (!live) v[i] = 0.f; q[i] = v[i] * v[i]; }
#pragma unroll
  for (int step = 0; step < 3; ++step) {
    const int off = 4 >> step, n = 32 >> step, hn = n >> 1;
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (i < hn) {
        const float sv = up ? v[i] : v[i + hn], sq = up ? q[i] : q[i + hn];
        const float rv = __shfl_xor_sync(0xffffffffu, sv, off), rq = __shfl_xor_sync(0xffffffffu, sq, off);
        v[i] = (up ? v[i + hn] : v[i]) + rv;
        q[i] = (up ? q[i + hn] : q[i]) + rq;
      }
    }
  }
  float* d = dst + (size_t)(lane >> 3) * ld + ((lane & 7) << 3);
  *reinterpret_cast<float4*>(d) = make_float4(v[0], q[0], v[1], q[1]);
  *reinterpret_cast<float4*>(d + 4) = make_float4(v[2], q[2], v[3], q[3]);
}

// plane of a large-level tile (one sample, one plane) and its 128-token block index within that plane
__device__ __forceinline__ void tc_tile_plane(const TcTile& T, int& p, int& blk) {
  if (T.tok0 < T.nxy) { p = 0; blk = T.tok0 >> 7; }
  else { const int r = T.tok0 - T.nxy; p = 1 + (r >= T.npl ? 1 : 0); blk = (r - (p - 1) * T.npl) >> 7; }
}
// first tile row and row count of (sample s of the tile, plane p) at a small level
__device__ __forceinline__ void tc_small_rows(const TcTile& T, int s, int p, int& r0, int& nr) {
  if (p == 0) { r0 = s << T.sh_xy; nr = T.nxy; }
  else { r0 = (T.spt << T.sh_xy) + (p - 1) * (T.spt << T.sh_pl) + (s << T.sh_pl); nr = T.npl; }
}

// s_cs (per 8-row group) -> one slot per (sample, plane[, 128-token block]) for tile columns [c_lo, c_hi); fixed order
template <int BN>
__device__ __forceinline__ void tc_write_slots(const TcConvParams& P, const Geo& g, const TcTile& T, int n0, const float* s_cs,
                                               int c_lo, int c_hi, int et) {
  const int ncol = c_hi - c_lo;
  if (!T.small) {
    int p, blk; tc_tile_plane(T, p, blk);
    for (int i = et; i < ncol; i += 256) {
      const int c = c_lo + i;
      double s = 0.0, q = 0.0;
#pragma unroll
      for (int rg = 0; rg < 16; ++rg) { const float2 v = *reinterpret_cast<const float2*>(s_cs + ((size_t)rg * BN + c) * 2); s += (double)v.x; q += (double)v.y; }
      *reinterpret_cast<double2*>(P.csum + csum_at(g, P.Cout, T.b0, p, blk, n0 + c)) = make_double2(s, q);
    }
  } else {
    const int nsamp = min(T.spt, P.B - T.b0);
    for (int i = et; i < nsamp * 3 * ncol; i += 256) {
      const int sp = i / ncol, c = c_lo + (i - sp * ncol);
      const int sm = sp / 3, p = sp - sm * 3;
      int r0, nr; tc_small_rows(T, sm, p, r0, nr);
      double s = 0.0, q = 0.0;
      for (int rg = r0 >> 3; rg < ((r0 + nr) >> 3); ++rg) { const float2 v = *reinterpret_cast<const float2*>(s_cs + ((size_t)rg * BN + c) * 2); s += (double)v.x; q += (double)v.y; }
      *reinterpret_cast<double2*>(P.csum + csum_at(g, P.Cout, T.b0 + sm, p, 0, n0 + c)) = make_double2(s, q);
    }
  }
}

__device__ __forceinline__ uint32_t tc_cvt_bf16x2(float lo_elem, float hi_elem) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
  return r;
}
// two values -> packed split-bf16 words (element 0 in the low half)
__device__ __forceinline__ void tc_split2(float y0, float y1, uint32_t& hi, uint32_t& lo) {
  hi = tc_cvt_bf16x2(y0, y1);
  lo = tc_cvt_bf16x2(y0 - __uint_as_float(hi << 16), y1 - __uint_as_float(hi & 0xffff0000u));
}

// bias + residual (any of the three geometries) for 4 channels of one output row
__device__ __forceinline__ float4 tc_bias_resid4(const TcConvParams& P, const Geo& g, size_t m, int n, int b, int p, int y, int x) {
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (P.bias) s = __ldg(reinterpret_cast<const float4*>(P.bias + n));
  if (P.resid) {
    if (P.resid_mode == RS_NONE) {
      const float4 rv = __ldg(reinterpret_cast<const float4*>(P.resid + m * P.Cout + n));
      s.x += rv.x; s.y += rv.y; s.z += rv.z; s.w += rv.w;
    } else if (P.resid_mode == RS_UP2) {
      const Geo gs = geo_down(g);
      const int ts = tc_plane_off(gs, p) + (y >> 1) * gs.res + (x >> 1);
      const float4 rv = __ldg(reinterpret_cast<const float4*>(P.resid + ((size_t)b * gs.L + ts) * P.Cout + n));
      s.x += rv.x; s.y += rv.y; s.z += rv.z; s.w += rv.w;
    } else {
      const Geo gs = geo_up(g);
      const int t0 = tc_plane_off(gs, p) + (2 * y) * gs.res + 2 * x;
      const float* rp = P.resid + ((size_t)b * gs.L + t0) * P.Cout + n;
      const float4 r0 = __ldg(reinterpret_cast<const float4*>(rp));
      const float4 r1 = __ldg(reinterpret_cast<const float4*>(rp + P.Cout));
      const float4 r2 = __ldg(reinterpret_cast<const float4*>(rp + (size_t)gs.res * P.Cout));
      const float4 r3 = __ldg(reinterpret_cast<const float4*>(rp + (size_t)(gs.res + 1) * P.Cout));
      s.x += 0.25f * (r0.x + r1.x + r2.x + r3.x); s.y += 0.25f * (r0.y + r1.y + r2.y + r3.y);
      s.z += 0.25f * (r0.z + r1.z + r2.z + r3.z); s.w += 0.25f * (r0.w + r1.w + r2.w + r3.w);
    }
  }
  return s;
}

// The tap-GEMM epilogue (warps 2-9 of the CTA; two warps per TMEM lane quarter, warp group (warp-2)/4 takes 32-column
// chunks group, group+2, ...).
//   EPI 1 / 3: TMEM -> +bias +residual -> fp32 `out` (bulk tensor stores of smem-staged sub-tiles) + channel-sum slots
//   EPI 2    : qkv operand split (Q / K / V^T split-bf16 for the attention kernel)
//   EPI 0    : split-K.  Partial tiles go to `partial` in a lane-coalesced scratch layout, the ksplit CTAs of an output tile
//              meet at the tile's ticket, then EACH of them reduces a contiguous share of the tile's columns in fixed order
//              (+bias +residual -> `out`, channel-sum slots) — no second launch, no atomics.
// (Also normalising the consumer's operand here, behind a grid-wide barrier, was built and measured slower than the stand-alone
// apply launch that overlaps its prologue through PDL: profiles/r02_fused_apply_experiment.md.)
template <int BN, int EPI>
__device__ __forceinline__ void tc_epilogue(const TcConvParams& P, const Geo& g, const TcTile& T, int n0, int zidx,
                                            uint32_t tmem_base, float* s_bias, uint64_t* bar_acc, long long* stamp,
                                            uint8_t* ring, uint32_t ring_u32) {
  constexpr int NCH = BN / 64;                  // 32-column chunks per warp
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int et = threadIdx.x - 64;              // epilogue thread index 0..255
  const int wg = (warp - 2) >> 2;               // column half taken by this warp
  const int q = warp & 3;                       // TMEM lane quarter this warp may read
  const int row = q * 32 + lane;                // row of the tile
  int b, tok; tc_row_map(T, row, b, tok);
  const bool live = b < P.B;
  const size_t m = (size_t)b * g.L + tok;
  float* s_cs = reinterpret_cast<float*>(ring + TC_EPI_CS_OFF);
  unsigned long long* sync = P.sync;
  MTV_PDL_WAIT();                               // residual / statistics buffers belong to earlier kernels
  unsigned long long tgt_tile = 0;
  if (EPI == 0 && et == 0)                      // this launch's ticket generation (see tc_gen_target): off the critical path
    tgt_tile = tc_gen_target(sync + (size_t)(blockIdx.x * gridDim.y + blockIdx.y) * TC_SYNC_STRIDE, (unsigned int)P.ksplit);
  // Bias: small, touched once per step and evicted from L2 by the weight stream in between, i.e. a DRAM
  // miss (~2000 cycles) if loaded on demand per chunk — so it is staged in smem during the main loop.
  if (EPI != 0) {
    if (et < BN) s_bias[et] = P.bias ? __ldg(P.bias + n0 + et) : 0.0f;
    TC_EPI_BAR();
  }
  int p = 0, y = 0, x = 0;
  tc_decode_fast(T, tok, p, y, x);
  
  if constexpr (EPI == 0) {
    // ---------------------------------------------------------------- split-K
    const int ks = P.ksplit;
    const int tile_id = blockIdx.x * gridDim.y + blockIdx.y;
    constexpr int U = BN / 4;                   // reduction units of a tile: 4 columns x 128 rows (one warp per row quarter)
    const int u0 = (zidx * U) / ks, u1 = ((zidx + 1) * U) / ks;     // this CTA's contiguous share
    constexpr int MAXPASS = U / 4;              // ks >= 2 -> <= U/2 units per CTA, two units (warp groups) per pass
    float4 add0 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live && u0 + wg < u1) add0 = tc_bias_resid4(P, g, m, n0 + 4 * (u0 + wg), b, p, y, x);   // first pass: fetched under the main loop
    mbar_wait(bar_acc, 0);
    MTV_PDL_TRIGGER();
    if (stamp && threadIdx.x == 64) stamp[4] = clock64();            // accumulator complete
    tc_fence_after();
    // partial tile -> scratch [tile][z][chunk = q*(BN/32) + c0/32][j][lane][4]: every store instruction writes 512 contiguous bytes
    float* part = P.partial + ((size_t)tile_id * ks + zidx) * (size_t)(TC_BM * BN);
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      const int c0 = (wg + 2 * k) * 32;
      uint32_t r[32], r2[32];
      tmem_ld32_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
      tmem_ld32_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(BN + c0), r2);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      float* dst = part + ((size_t)(q * (BN / 32) + (c0 >> 5)) * 8) * 128 + lane * 4;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 v = make_float4(__uint_as_float(r[4 * j]) + __uint_as_float(r2[4 * j]), __uint_as_float(r[4 * j + 1]) + __uint_as_float(r2[4 * j + 1]),
                                     __uint_as_float(r[4 * j + 2]) + __uint_as_float(r2[4 * j + 2]), __uint_as_float(r[4 * j + 3]) + __uint_as_float(r2[4 * j + 3]));
        __stcg(reinterpret_cast<float4*>(dst + (size_t)j * 128), v);
      }
    }
    TC_EPI_BAR();
    if (et == 0) tc_gen_barrier(sync + (size_t)tile_id * TC_SYNC_STRIDE, tgt_tile);
    TC_EPI_BAR();
    if (stamp && threadIdx.x == 64) stamp[5] = clock64();            // every partial of the tile is visible
    float4 val[MAXPASS];
    const float* pbase = P.partial + (size_t)tile_id * ks * (size_t)(TC_BM * BN);
#pragma unroll
    for (int ps = 0; ps < MAXPASS; ++ps) {
      const int u = u0 + 2 * ps + wg;
      val[ps] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (u0 + 2 * ps < u1) {                    // warp-uniform per pass pair; the shuffles below need whole warps
        const bool mine = u < u1;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        if (mine) {
          const float* src = pbase + ((size_t)(q * (BN / 32) + (u >> 3)) * 8 + (u & 7)) * 128 + lane * 4;
          // the reduction is a chain of L2 round trips (~600 cycles each), not arithmetic: 16 independent loads per trip, summed
          // in a fixed order (z ascending) whatever the trip size
          int z = 0;
          for (; z + 16 <= ks; z += 16) {
            float4 v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __ldcg(reinterpret_cast<const float4*>(src + (size_t)(z + i) * (TC_BM * BN)));
#pragma unroll
            for (int i = 0; i < 16; ++i) { s.x += v[i].x; s.y += v[i].y; s.z += v[i].z; s.w += v[i].w; }
          }
          if (z + 8 <= ks) {
            float4 v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __ldcg(reinterpret_cast<const float4*>(src + (size_t)(z + i) * (TC_BM * BN)));
#pragma unroll
            for (int i = 0; i < 8; ++i) { s.x += v[i].x; s.y += v[i].y; s.z += v[i].z; s.w += v[i].w; }
            z += 8;
          }
          {
            float4 v[8];                        // the last < 8 partials in one trip as well
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = (z + i < ks) ? __ldcg(reinterpret_cast<const float4*>(src + (size_t)(z + i) * (TC_BM * BN))) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 8; ++i) if (z + i < ks) { s.x += v[i].x; s.y += v[i].y; s.z += v[i].z; s.w += v[i].w; }
          }
          if (live) {
            const float4 ad = ps == 0 ? add0 : tc_bias_resid4(P, g, m, n0 + 4 * u, b, p, y, x);
            s.x += ad.x; s.y += ad.y; s.z += ad.z; s.w += ad.w;
            if (!(P.dbg_skip & 8)) *reinterpret_cast<float4*>(P.out + m * P.Cout + n0 + 4 * u) = s;
          } else {
            s = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          val[ps] = s;
        }
        if (P.csum && mine) {
          float v[4] = {s.x, s.y, s.z, s.w}, sq[4] = {s.x * s.x, s.y * s.y, s.z * s.z, s.w * s.w};
#pragma unroll
          for (int off = 1; off < 8; off <<= 1)
#pragma unroll
            for (int i = 0; i < 4; ++i) { v[i] += __shfl_xor_sync(0xffffffffu, v[i], off); sq[i] += __shfl_xor_sync(0xffffffffu, sq[i], off); }
          if ((lane & 7) == 0) {
            float* d = s_cs + ((size_t)(q * 4 + (lane >> 3)) * BN + 4 * u) * 2;
            *reinterpret_cast<float4*>(d) = make_float4(v[0], sq[0], v[1], sq[1]);
            *reinterpret_cast<float4*>(d + 4) = make_float4(v[2], sq[2], v[3], sq[3]);
          }
        }
      }
    }
    if (P.csum) {
      TC_EPI_BAR();
      if (u1 > u0) tc_write_slots<BN>(P, g, T, n0, s_cs, 4 * u0, 4 * u1, et);
    }
    if (stamp && threadIdx.x == 64) stamp[7] = clock64();
    return;
  } else {
    // ---------------------------------------------------------------- one CTA owns the whole K range
    // the residual of the first 32-column chunk is fetched while the MMAs still run
    const bool pre_res = EPI == 1 && live && P.resid;
    float rpre[32];
    if (pre_res) {
#pragma unroll
      for (int j = 0; j < 4; ++j) ld_global_nc_v8(P.resid + m * P.Cout + n0 + wg * 32 + 8 * j, rpre + 8 * j);
    }
    mbar_wait(bar_acc, 0);
    MTV_PDL_TRIGGER();
    if (stamp && threadIdx.x == 64) stamp[4] = clock64();            // accumulator complete
    tc_fence_after();
    // bulk-store path: the tile's rows are consecutive rows of the output (always at the large levels; at the small ones
    // only when a tile is exactly one sample) and the op has an output map
    const bool ts = EPI != 2 && EPI != 4 && P.tma_store && (!T.small || T.spt == 1);
    const int ts_row = (int)((size_t)T.b0 * g.L + T.tok0) + q * 32;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      const int c0 = (wg + 2 * k) * 32;
      const uint32_t sbase = ring_u32 + TC_EPI_STAGE_OFF + (uint32_t)((warp - 2) * NCH + k) * 4096u;
      uint32_t r[32];
      {
        uint32_t r2[32];
        tmem_ld32_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
        tmem_ld32_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(BN + c0), r2);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
      }
      const int n = n0 + c0;
      float rnext[32];
      const bool have_next = pre_res && (k + 1 < NCH);
      if (have_next) {
#pragma unroll
        for (int j = 0; j < 4; ++j) ld_global_nc_v8(P.resid + m * P.Cout + n + 64 + 8 * j, rnext + 8 * j);
      }
      float fv[32];
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        {
          const float4 bv = *reinterpret_cast<const float4*>(&s_bias[c0 + j]);
          v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
        }
        if constexpr (EPI == 1) {
          if (pre_res) { v.x += rpre[j]; v.y += rpre[j + 1]; v.z += rpre[j + 2]; v.w += rpre[j + 3]; }
        }
        if constexpr (EPI == 3) {
          if (live) {
            if (P.resid_mode == RS_UP2) {
              const Geo gs = geo_down(g);
              const int tsrc = tc_plane_off(gs, p) + (y >> 1) * gs.res + (x >> 1);
              const float4 rv = __ldg(reinterpret_cast<const float4*>(P.resid + ((size_t)b * gs.L + tsrc) * P.Cout + n + j));
              v.x += rv.x; v.y += rv.y; v.z += rv.z; v.w += rv.w;
            } else {
              const Geo gs = geo_up(g);
              const int t0 = tc_plane_off(gs, p) + (2 * y) * gs.res + 2 * x;
              const float* rp = P.resid + ((size_t)b * gs.L + t0) * P.Cout + n + j;
              const float4 r0 = __ldg(reinterpret_cast<const float4*>(rp));
              const float4 r1 = __ldg(reinterpret_cast<const float4*>(rp + P.Cout));
              const float4 r2 = __ldg(reinterpret_cast<const float4*>(rp + (size_t)gs.res * P.Cout));
              const float4 r3 = __ldg(reinterpret_cast<const float4*>(rp + (size_t)(gs.res + 1) * P.Cout));
              v.x += 0.25f * (r0.x + r1.x + r2.x + r3.x); v.y += 0.25f * (r0.y + r1.y + r2.y + r3.y);
              v.z += 0.25f * (r0.z + r1.z + r2.z + r3.z); v.w += 0.25f * (r0.w + r1.w + r2.w + r3.w);
            }
          }
        }
        fv[j] = v.x; fv[j + 1] = v.y; fv[j + 2] = v.z; fv[j + 3] = v.w;
      }
      if constexpr (EPI == 4) {
        // head conv (unet.py:971-975): only the first out_cvalid columns are real; eps is channel-major [B][cvalid][L] — for a fixed
        // channel the 32 lanes of a warp write 32 consecutive tokens
        if (c0 == 0 && live) {
#pragma unroll
          for (int nn = 0; nn < 32; ++nn)
            if (nn < P.out_cvalid) P.out[((size_t)b * P.out_cvalid + nn) * g.L + tok] = fv[nn];
        }
      }
      if constexpr (EPI != 2 && EPI != 4) {
        if (ts) tc_stage_row(sbase, lane, fv);
        if (ts) {      // warp-uniform: every row of a bulk-stored tile is live
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0 && !(P.dbg_skip & 8)) { tma_store_2d(&P.tmOut, sbase, n, ts_row); tma_store_commit(); }
        } else if (live && !(P.dbg_skip & 8)) {
          float* dst = P.out + m * P.Cout + n;
#pragma unroll
          for (int j = 0; j < 32; j += 8) st_global_v8(dst + j, fv + j);
        }
        if (P.csum && !(P.dbg_skip & 4)) {
          __syncwarp();
          tc_csum_chunk(fv, live, lane, s_cs + ((size_t)(q * 4) * BN + c0) * 2, BN * 2);
        }
      }
      if constexpr (EPI == 2) {
        if (live) {
        // channels are head-major [h: q(D) k(D) v(D)] (unet.py:321); D >= 16, so every aligned run of 16
        // channels is one of q / k / v of one head
        const int Dh = P.Cout / (3 * P.qkv_heads);
        const float qs = 1.4426950408889634f * rsqrtf((float)Dh);
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int nn = n + hf * 16;
          const int hd = nn / (3 * Dh), rr = nn - hd * 3 * Dh;
          const int kind = rr / Dh, d0 = rr - kind * Dh;
          const size_t bh = (size_t)b * P.qkv_heads + hd;
          __align__(16) __nv_bfloat16 hh[16], ll[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float val = kind == 0 ? fv[hf * 16 + i] * qs : fv[hf * 16 + i];
            hh[i] = __float2bfloat16_rn(val);
            ll[i] = __float2bfloat16_rn(val - __bfloat162float(hh[i]));
          }
          if (kind < 2) {
            __nv_bfloat16* ph = reinterpret_cast<__nv_bfloat16*>(kind == 0 ? P.q_hi : P.k_hi) + (bh * g.L + tok) * Dh + d0;
            __nv_bfloat16* pw = reinterpret_cast<__nv_bfloat16*>(kind == 0 ? P.q_lo : P.k_lo) + (bh * g.L + tok) * Dh + d0;
            reinterpret_cast<uint4*>(ph)[0] = reinterpret_cast<const uint4*>(hh)[0];
            reinterpret_cast<uint4*>(ph)[1] = reinterpret_cast<const uint4*>(hh)[1];
            reinterpret_cast<uint4*>(pw)[0] = reinterpret_cast<const uint4*>(ll)[0];
            reinterpret_cast<uint4*>(pw)[1] = reinterpret_cast<const uint4*>(ll)[1];
          } else {
            __nv_bfloat16* ph = reinterpret_cast<__nv_bfloat16*>(P.vt_hi) + (bh * Dh + d0) * g.L + tok;
            __nv_bfloat16* pw = reinterpret_cast<__nv_bfloat16*>(P.vt_lo) + (bh * Dh + d0) * g.L + tok;
#pragma unroll
            for (int i = 0; i < 16; ++i) { ph[(size_t)i * g.L] = hh[i]; pw[(size_t)i * g.L] = ll[i]; }
          }
        }
        }
      }
      if (have_next) {
#pragma unroll
        for (int j = 0; j < 32; ++j) rpre[j] = rnext[j];
      }
    }
    if constexpr (EPI != 2 && EPI != 4) {
      if (P.csum && !(P.dbg_skip & 4)) {
        TC_EPI_BAR();
        tc_write_slots<BN>(P, g, T, n0, s_cs, 0, BN, et);
      }
      if (ts && lane == 0) tma_store_wait_read();     // the staging tiles must outlive the bulk stores' reads
    }
    if (stamp && threadIdx.x == 64) stamp[7] = clock64();
  }
}

// EPI selects the epilogue at compile time (the row-per-lane epilogue is instruction-issue bound, so the
// paths a launch cannot take must not even be predicated off):
//   0 split-K (partials + in-kernel reduction), 1 bias [+ same-geometry residual] [+ GroupNorm sums], 2 qkv operand split,
//   3 residual through nearest-up / avg-pool geometry (up / down ResBlocks with identity skip),
//   4 head conv: channel-major output of the first out_cvalid columns (BN = 64 only)
template <int BN, int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1) k_conv_tc(const __grid_constant__ TcConvParams P) {
  constexpr int NS = tc_stages(BN);
  constexpr int STAGE = tc_stage_bytes(BN);
  // Stacked-N: W_hi and W_lo tiles are adjacent in smem, so ONE MMA with N = 2*BN computes
  // [A_hi*W_hi | A_hi*W_lo] into TMEM columns [0,BN) | [BN,2BN) and a second one adds A_lo*W_hi into [0,BN):
  // two MMAs and 14 KB of operand reads per k-step instead of three and 18 KB; the epilogue sums the two column groups.
  constexpr uint32_t IDESC = umma_idesc_bf16(TC_BM, BN);
  constexpr uint32_t IDESC2 = umma_idesc_bf16(TC_BM, 2 * BN);
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[NS], bar_empty[NS], bar_acc;
  __shared__ uint32_t tmem_base_s;
  __shared__ long long s_stamp[8];
  __shared__ __align__(16) float s_bias[128];      // this CTA's BN bias values, fetched while the main loop runs
  const bool dbg = g_tc_dbg != nullptr;
  long long g_t0 = 0;
  if (dbg && threadIdx.x == 0) { for (int i = 1; i < 8; ++i) s_stamp[i] = 0; s_stamp[0] = clock64(); g_t0 = gtime_ns(); }
  mtv_prefetch_slice(P.pf0, P.pf1, P.pf_bytes, blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z),
                     gridDim.x * gridDim.y * gridDim.z);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B tiles need 1024-B alignment
  uint8_t* ring = smem_raw + (smem0 - smem_u32(smem_raw));

  const Geo g = P.geo;
  const int n0 = blockIdx.y * BN;
  const TcTile T = tc_tile(g, blockIdx.x);

  // K range of this CTA (split-K over the flattened (tap, 64-channel chunk) space)
  const int kch = P.Cin / TC_BK;
  const int it_main = P.taps * kch;
  const int it_total = it_main + P.Cin2 / TC_BK;
  int it0 = 0, it1 = it_total;
  if (P.ksplit > 1) {
    const int per = (it_total + P.ksplit - 1) / P.ksplit;
    it0 = blockIdx.z * per; it1 = min(it_total, it0 + per);
  }

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
    mbar_init(&bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    prefetch_tmap(&P.tmA_hi[0]); prefetch_tmap(&P.tmA_lo[0]);
    prefetch_tmap(&P.tmW_hi); prefetch_tmap(&P.tmW_lo);
  }
  if (warp == 1) {   // TMEM: 2*BN fp32 accumulator columns x 128 lanes
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)(2 * BN)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (dbg && threadIdx.x == 0) s_stamp[1] = clock64();
  // bytes TMA delivers per stage (diagnostic skips: operand halves that are not fetched are not expected either)
  const uint32_t TX_BYTES = (uint32_t)(((P.dbg_skip & 1) ? 0 : 2 * TC_BM * 128) + ((P.dbg_skip & 2) ? 0 : 2 * BN * 128));
  const int npre = min(NS, it1 - it0);

  auto load_W = [&](int it, int stage) {
    if (P.dbg_skip & 2) return;
    const uint32_t sW_hi = smem0 + stage * STAGE + 2 * TC_BM * 128, sW_lo = sW_hi + BN * 128;
    const uint32_t fb = smem_u32(&bar_full[stage]);
    if (it >= it_main) {
      const int c2 = (it - it_main) * TC_BK;
      tma_load_2d(sW_hi, &P.tmW2_hi, fb, c2, n0);
      tma_load_2d(sW_lo, &P.tmW2_lo, fb, c2, n0);
    } else {
      const int tap = it / kch, c0 = (it - tap * kch) * TC_BK;
      tma_load_2d(sW_hi, &P.tmW_hi, fb, c0, tap * P.Cout + n0);
      tma_load_2d(sW_lo, &P.tmW_lo, fb, c0, tap * P.Cout + n0);
    }
  };

  if (it1 <= it0) {
    // (cannot happen: the host never creates empty K ranges) — keep the barriers consistent anyway
  }
  if (warp == 0) {
    // =============================== TMA producer ===============================
    // whole warp in the loop (uniform coordinates / descriptors stay in uniform registers), one elected lane issues
    {
      auto load_A = [&](int it, int stage) {
        const uint32_t sA_hi = smem0 + stage * STAGE, sA_lo = sA_hi + TC_BM * 128;
        const uint32_t fb = smem_u32(&bar_full[stage]);
        if (it >= it_main) {           // second K-segment: 1x1 conv of the skip operand
          tc_load_A(g, T, P.tmA2_hi, P.tmA2_lo, 1, 0, (it - it_main) * TC_BK, sA_hi, sA_lo, fb);
        } else {
          const int tap = it / kch, c0 = (it - tap * kch) * TC_BK;
          tc_load_A(g, T, P.tmA_hi, P.tmA_lo, P.taps, tap, c0, sA_hi, sA_lo, fb);
        }
      };
      // Weights do not depend on the previous kernel: fill the ring's W halves BEFORE the grid
      // dependency resolves (overlaps their HBM latency with the predecessor's tail) ...
      if (elect_one()) {
        for (int i = 0; i < npre; ++i) {
          mbar_expect_tx(&bar_full[i], TX_BYTES);
          load_W(it0 + i, i);
        }
      }
      __syncwarp();
      MTV_PDL_WAIT();                // ... the activation operand does
      int stage = 0; uint32_t phase = 0;
      for (int it = it0; it < it1; ++it) {
        if (it - it0 >= npre) mbar_wait_guard(&bar_empty[stage], phase ^ 1u);
        if (elect_one()) {
          if (it - it0 >= npre) {
            mbar_expect_tx(&bar_full[stage], TX_BYTES);
            load_W(it, stage);
          }
          if (!(P.dbg_skip & 1)) load_A(it, stage);
        }
        __syncwarp();
        if (++stage == NS) { stage = 0; phase ^= 1u; }
      }
    }
    // Dependents may be scheduled once the accumulator is complete: only this CTA's epilogue remains, so
    // the next kernel's launch latency and prologue overlap it without piling up many kernels deep.  Every
    // thread of the CTA triggers at that same point (the instruction's per-CTA semantics are not relied on).
    mbar_wait_guard(&bar_acc, 0);
    MTV_PDL_TRIGGER();
  } else if (warp == 1) {
    // =============================== MMA issuer =================================
    // whole warp in the loop, one elected lane issues (see elect_one)
    {
      const uint64_t d0 = umma_desc_sw128(smem0);                   // descriptor of the ring's first byte; tiles are +offset/16
      int stage = 0; uint32_t phase = 0;
      for (int it = it0; it < it1; ++it) {
        mbar_wait(&bar_full[stage], phase);
        if (dbg && lane == 0 && it == it0) s_stamp[2] = clock64();       // first operands landed
        if (dbg && lane == 0 && it == it1 - 1) s_stamp[3] = clock64();   // last operands landed
        tc_fence_after();
        const uint64_t dA = d0 + (uint64_t)((uint32_t)(stage * STAGE) >> 4);
        const uint32_t acc0 = it > it0 ? 1u : 0u;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            const uint64_t a_hi = dA + (uint64_t)(k * 2), a_lo = a_hi + (uint64_t)((TC_BM * 128) >> 4);
            const uint64_t w_hi = a_hi + (uint64_t)((2 * TC_BM * 128) >> 4);    // rows [0,BN) = W_hi, [BN,2BN) = W_lo
            umma_bf16(tmem_base, a_hi, w_hi, IDESC2, k > 0 ? 1u : acc0);
            umma_bf16(tmem_base, a_lo, w_hi, IDESC, 1u);
          }
          umma_commit(&bar_empty[stage]);        // frees the smem slot once these MMAs have read it
        }
        __syncwarp();
        if (++stage == NS) { stage = 0; phase ^= 1u; }
      }
      if (elect_one()) umma_commit(&bar_acc);     // accumulator complete
      __syncwarp();
    }
    mbar_wait(&bar_acc, 0);
    MTV_PDL_TRIGGER();
  } else {
    tc_epilogue<BN, EPI>(P, g, T, n0, (int)blockIdx.z, tmem_base, s_bias, &bar_acc, dbg ? s_stamp : nullptr, ring, smem0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * BN)) : "memory");
  }
  if (dbg && threadIdx.x == 0) {
    const unsigned int slot = atomicAdd(&g_tc_dbg_count, 1u);
    if (slot < g_tc_dbg_cap) {
      long long* rec = g_tc_dbg + (size_t)slot * 16;
      rec[0] = (long long)gridDim.x | ((long long)gridDim.y << 16) | ((long long)gridDim.z << 32);
      rec[1] = (long long)(it1 - it0) | ((long long)P.taps << 16) | ((long long)P.Cin << 24) | ((long long)P.Cout << 40);
      for (int i = 0; i < 8; ++i) rec[2 + i] = s_stamp[i];
      rec[10] = g_t0; rec[11] = gtime_ns();
      unsigned int smid; asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
      rec[12] = smid; rec[13] = (long long)blockIdx.x | ((long long)blockIdx.y << 16) | ((long long)blockIdx.z << 32);
      rec[14] = clock64(); rec[15] = (P.csum ? 2 : 0);
    }
  }
}

cudaError_t launch_conv_tc(const TcConvParams& P, cudaStream_t s) {
  const int BN = P.bn;
  if ((BN != 64 && BN != 128) || P.Cout % BN || P.Cin % TC_BK || P.Cin2 % TC_BK) return cudaErrorInvalidValue;
  if (P.geo.L > TC_BM ? (P.geo.L % TC_BM != 0) : (TC_BM % P.geo.L != 0)) return cudaErrorInvalidValue;
  const int M = P.B * P.geo.L;
  const int tiles = P.geo.L > TC_BM ? M / TC_BM : (P.B + (TC_BM / P.geo.L) - 1) / (TC_BM / P.geo.L);
  const int ks = P.ksplit > 1 ? P.ksplit : 1;
  dim3 grid(tiles, P.Cout / BN, ks);
  if (ks > 1 && !P.sync) return cudaErrorInvalidValue;       // in-kernel waits need the co-residency guarantee
  if (ks > 1 && (!P.partial || P.qkv_heads)) return cudaErrorInvalidValue;
  cudaError_t e = cudaSuccess;
  const int epi = ks > 1 ? 0 : (P.qkv_heads ? 2 : (P.out_cvalid ? 4 : ((P.resid && P.resid_mode != RS_NONE) ? 3 : 1)));
  if (epi == 4 && (BN != 64 || P.Cout != 64 || P.out_cvalid > 32)) return cudaErrorInvalidValue;
#define MTV_TC_LAUNCH(BN_, EPI_)                                                                                   \
  do {                                                                                                             \
    e = cudaFuncSetAttribute(k_conv_tc<BN_, EPI_>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes(BN_)); \
    if (e != cudaSuccess) return e;                                                                                \
    e = launch_kc(PDL_CLASS_CONV_TC, k_conv_tc<BN_, EPI_>, grid, dim3(TC_THREADS), (size_t)tc_smem_bytes(BN_), s, P); \
    if (e != cudaSuccess) return e;                                                                                \
  } while (0)
  if (BN == 64) {
    switch (epi) { case 0: MTV_TC_LAUNCH(64, 0); break; case 1: MTV_TC_LAUNCH(64, 1); break;
                   case 2: MTV_TC_LAUNCH(64, 2); break; case 4: MTV_TC_LAUNCH(64, 4); break; default: MTV_TC_LAUNCH(64, 3); break; }
  } else {
    switch (epi) { case 0: MTV_TC_LAUNCH(128, 0); break; case 1: MTV_TC_LAUNCH(128, 1); break;
                   case 2: MTV_TC_LAUNCH(128, 2); break; default: MTV_TC_LAUNCH(128, 3); break; }
  }
#undef MTV_TC_LAUNCH
  return cudaGetLastError();
}

}  // namespace mtv
