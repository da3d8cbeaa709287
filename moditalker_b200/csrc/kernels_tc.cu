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
    double s = 0.0, ss = 0.0;
    for (int ci = lane & 7; ci < cpg; ci += 8) {
      const int c = grp * cpg + ci;
      const double* cs; int Cs, cc;
      if (c < P.C0) { cs = P.csum0; Cs = P.C0; cc = c; } else { cs = P.csum1; Cs = P.C1; cc = c - P.C0; }
      if (P.joint) {
        const double2 q0 = *reinterpret_cast<const double2*>(cs + (((size_t)b * 3 + 0) * Cs + cc) * 2);
        const double2 q1 = *reinterpret_cast<const double2*>(cs + (((size_t)b * 3 + 1) * Cs + cc) * 2);
        const double2 q2 = *reinterpret_cast<const double2*>(cs + (((size_t)b * 3 + 2) * Cs + cc) * 2);
        s += q0.x + q1.x + q2.x; ss += q0.y + q1.y + q2.y;
      } else {
        const double2 q0 = *reinterpret_cast<const double2*>(cs + (((size_t)b * 3 + p) * Cs + cc) * 2);
        s += q0.x; ss += q0.y;
      }
    }
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
// Bounded: a pipeline bug must surface as a launch failure (trap), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
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
// direct mode: 8 producer warps (two groups alternating K-iterations) + the affine table behind the operand ring
constexpr int TC_THREADS_DIRECT = 320;
__host__ __device__ constexpr int tc_smem_bytes_direct(int BN) { return tc_smem_bytes(BN) + TC_TABLE_ENTRIES * 8; }

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


// Per-channel (sum, sum of squares) of a 32-column chunk over each aligned group of 8 tile rows,
// accumulated into csum[b][plane][channel][2] (fp64 atomics).  8 rows never straddle a
// (sample, plane) boundary at any level (plane sizes are multiples of 8 tokens).  Butterfly
// transpose-reduce: after the three exchange steps lane l holds columns ((l & 7) << 2) + {0..3}.
__device__ __forceinline__ void tc_csum_chunk(float (&v)[32], bool live, int lane, double* csum_bp /* &csum[b][plane][n] */) {
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
  if (live) {
    const int cbase = (lane & 7) << 2;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      atomicAdd(csum_bp + 2 * (cbase + i), (double)v[i]);
      atomicAdd(csum_bp + 2 * (cbase + i) + 1, (double)q[i]);
    }
  }
}

// Same, over all 32 rows of the warp (valid when they share one (sample, plane)): after five exchange
// steps lane l holds column l -> 2 atomics per lane instead of 8, and 4x fewer same-address atomics.
__device__ __forceinline__ void tc_csum_chunk32(float (&v)[32], bool live, int lane, double* csum_bp) {
  float q[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) { if (!live) v[i] = 0.f; q[i] = v[i] * v[i]; }
#pragma unroll
  for (int step = 0; step < 5; ++step) {
    const int off = 16 >> step, hn = 16 >> step;
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
  // the whole warp shares liveness here (uniform (sample, plane)); lane l owns column l
  if (__any_sync(0xffffffffu, live)) {
    atomicAdd(csum_bp + 2 * lane, (double)v[0]);
    atomicAdd(csum_bp + 2 * lane + 1, (double)q[0]);
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

// ------------------------------------------------------------------ direct A operand (no apply pass)
// sample-plane slot of a tile row in the affine table
__device__ __forceinline__ int tc_sp_index_local(const TcTile& T, bool per_plane, int sample_in_tile, int p) {
  return T.small ? (per_plane ? sample_in_tile * 3 + p : sample_in_tile) : 0;
}

// Affine table of one K-segment for this CTA's tile: tbl[sp][c] = (a, d) with y = x*a + d the GroupNorm (+FiLM) of channel c
// for the (sample, plane) pair sp.  Same arithmetic (fp64 statistics, fp32 result) as apply_norm_unit.  `scratch` holds
// 2 doubles per (sp, group).  Called by all `nthreads` threads of the CTA (nthreads % 32 == 0).
__device__ __forceinline__ void tc_build_table(const DirectSeg& S, const Geo& g, const TcTile& T, int B, float2* tbl, double* scratch,
                                               int nthreads) {
  const int C = S.C0 + S.C1, cpg = C / 32;
  const Geo gs = S.resample == RS_NONE ? g : (S.resample == RS_UP2 ? geo_down(g) : geo_up(g));
  const bool per_plane = S.mode == DS_NORM_CSUM ? !S.joint : S.nrm_nseg == 3;
  const int nsamp = T.small ? min(T.spt, B - T.b0) : 1;
  const int p_tile = T.small ? 0 : (T.tok0 < T.nxy ? 0 : 1 + (T.tok0 - T.nxy) / T.npl);
  const int nsp = T.small ? nsamp * (per_plane ? 3 : 1) : 1;
  const int tid = threadIdx.x;
  if (S.mode == DS_NORM_CSUM) {
    for (int idx = tid; idx < nsp * 256; idx += nthreads) {      // 8 lanes per (sp, group); whole warps in or out
      const int l8 = idx & 7, sgi = idx >> 3, sp = sgi >> 5, grp = sgi & 31;
      const int sl = T.small ? (per_plane ? sp / 3 : sp) : 0;
      const int p = T.small ? (per_plane ? sp - sl * 3 : 0) : p_tile;
      const int b = T.b0 + sl;
      double sm = 0.0, ss = 0.0;
      for (int ci = l8; ci < cpg; ci += 8) {
        const int c = grp * cpg + ci;
        const double* cs; int Cs, cc;
        if (c < S.C0) { cs = S.csum0; Cs = S.C0; cc = c; } else { cs = S.csum1; Cs = S.C1; cc = c - S.C0; }
        if (S.joint) {
          const double2 q0 = *reinterpret_cast<const double2*>(cs + (((size_t)b * 3 + 0) * Cs + cc) * 2);
          const double2 q1 = *reinterpret_cast<const double2*>(cs + (((size_t)b * 3 + 1) * Cs + cc) * 2);
          const double2 q2 = *reinterpret_cast<const double2*>(cs + (((size_t)b * 3 + 2) * Cs + cc) * 2);
          sm += q0.x + q1.x + q2.x; ss += q0.y + q1.y + q2.y;
        } else {
          const double2 q0 = *reinterpret_cast<const double2*>(cs + (((size_t)b * 3 + p) * Cs + cc) * 2);
          sm += q0.x; ss += q0.y;
        }
      }
#pragma unroll
      for (int off = 4; off > 0; off >>= 1) { sm += __shfl_xor_sync(0xffffffffu, sm, off); ss += __shfl_xor_sync(0xffffffffu, ss, off); }
      if (l8 == 0) {
        const double cnt = (double)cpg * (S.joint ? (double)gs.L : (double)(p == 0 ? gs.res * gs.res : gs.t * gs.res));
        const double mean = sm / cnt;
        double var = ss / cnt - mean * mean; var = var < 0.0 ? 0.0 : var;
        scratch[2 * sgi] = mean; scratch[2 * sgi + 1] = rsqrt(var + 1e-5);
      }
    }
    __syncthreads();
    for (int e = tid; e < nsp * C; e += nthreads) {
      const int sp = e / C, c = e - sp * C;
      const int sl = T.small ? (per_plane ? sp / 3 : sp) : 0;
      const int b = T.b0 + sl;
      const int grp = c / cpg;
      double a = scratch[2 * (sp * 32 + grp) + 1] * (double)__ldg(S.gamma + c);
      double d = (double)__ldg(S.beta + c) - scratch[2 * (sp * 32 + grp)] * a;
      if (S.film) {
        const float* f = S.film + (size_t)b * S.film_stride;
        const double sc = 1.0 + (double)__ldg(f + c);
        a *= sc; d = d * sc + (double)__ldg(f + C + c);
      }
      tbl[e] = make_float2((float)a, (float)d);
    }
  } else {   // DS_NORM_TABLE
    for (int e = tid; e < nsp * C; e += nthreads) {
      const int sp = e / C, c = e - sp * C;
      const int sl = T.small ? (per_plane ? sp / 3 : sp) : 0;
      const int p = T.small ? (per_plane ? sp - sl * 3 : 0) : p_tile;
      const size_t ni = ((size_t)(T.b0 + sl) * S.nrm_nseg + (per_plane ? p : 0)) * C + c;
      tbl[e] = make_float2(__ldg(S.nrm_a + ni), __ldg(S.nrm_d + ni));
    }
  }
}

// 8 transformed channels -> split bf16 (hi, lo), packed for one 16-byte operand chunk each
__device__ __forceinline__ uint32_t tc_cvt_bf16x2(float lo_elem, float hi_elem) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
  return r;
}
__device__ __forceinline__ void tc_pack_split8(const float (&y)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = tc_cvt_bf16x2(y[2 * i], y[2 * i + 1]);
    l[i] = tc_cvt_bf16x2(y[2 * i] - __uint_as_float(h[i] << 16), y[2 * i + 1] - __uint_as_float(h[i] & 0xffff0000u));
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]); lo = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ float4 ld_shared_v4f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

// One fp32 A tile (128 rows x 64 channels, 256-byte rows, unswizzled) of K-chunk c0 of tap `tap` from ONE source tensor —
// same boxes / coordinates as tc_load_A, fp32 tensor maps: m[0] = xy plane (or the 2-D / 3-D map when taps == 1), m[1] = planes.
__device__ __forceinline__ void tc_load_A32(const Geo& g, const TcTile& t, const CUtensorMap* m, int taps, int tap, int c0,
                                            uint32_t sA, uint32_t fb) {
  if (!t.small) {
    if (taps == 1) { tma_load_2d(sA, &m[0], fb, c0, t.b0 * g.L + t.tok0); return; }
    const int dy = tap / 3 - 1, dx = tap - (tap / 3) * 3 - 1;
    if (t.tok0 < t.nxy) {
      tma_load_4d(sA, &m[0], fb, c0, dx, t.tok0 / g.res + dy, t.b0);
    } else {
      const int r = t.tok0 - t.nxy;
      const int pl = r / t.npl, y0 = (r - pl * t.npl) / g.res;
      tma_load_5d(sA, &m[1], fb, c0, dx, y0 + dy, pl, t.b0);
    }
    return;
  }
  const uint32_t o1 = (uint32_t)(t.spt * t.nxy) * 256u, o2 = o1 + (uint32_t)(t.spt * t.npl) * 256u;
  if (taps == 1) {
    tma_load_3d(sA, &m[0], fb, c0, 0, t.b0);
    tma_load_3d(sA + o1, &m[1], fb, c0, t.nxy, t.b0);
    tma_load_3d(sA + o2, &m[1], fb, c0, t.nxy + t.npl, t.b0);
  } else {
    const int dy = tap / 3 - 1, dx = tap - (tap / 3) * 3 - 1;
    tma_load_4d(sA, &m[0], fb, c0, dx, dy, t.b0);
    tma_load_5d(sA + o1, &m[1], fb, c0, dx, dy, 0, t.b0);
    tma_load_5d(sA + o2, &m[1], fb, c0, dx, dy, 1, t.b0);
  }
}

// A-operand producer of the direct mode.  TMA lands the RAW fp32 tile (shifted box of this tap, zero fill outside the
// plane) in the stage's A region; a group of four warps then converts it IN PLACE into the split-bf16 operand pair:
// read own 8 channels of 8 rows into registers -> group barrier (every read done) -> y = silu?(x*a + d) -> hi / lo chunks
// written 128B-swizzled (chunk c of row r at c ^ (r & 7)) over the same 32 KB.  Two groups alternate K-iterations.
// Rows whose tap falls outside the plane must stay ZERO after the transform (the conv pads the normalised activation),
// hence the (row, tap) validity test from rowinfo[r] = (sample in tile << 16) | (plane << 12) | (y << 6) | x, or -1.
template <int BN>
__device__ __forceinline__ void tc_produce_A(const TcConvParams& P, const Geo& g, const TcTile& T, int it0, int it1, int it_main, int kch,
                                             uint32_t smem0, uint64_t* bar_raw, uint64_t* bar_full, const float2* tbl,
                                             const int* rowinfo, int group, int tg, long long* pstamp) {
  constexpr int NS = tc_stages(BN);
  constexpr int STAGE = tc_stage_bytes(BN);
  const int lane = threadIdx.x & 31;
  const int sub = tg & 7, rbase = tg >> 3;
  int stage = 0; uint32_t phase = 0;
#pragma unroll 1
  for (int it = it0; it < it1; ++it) {
    if (((it - it0) & 1) == group) {
      const bool seg1 = it >= it_main;
      const DirectSeg& S = P.dseg[seg1 ? 1 : 0];
      const int tap = seg1 ? 0 : it / kch;
      const int c0 = seg1 ? (it - it_main) * TC_BK : (it - tap * kch) * TC_BK;
      const int ntaps = seg1 ? 1 : P.taps;
      const int dy = ntaps == 1 ? 0 : tap / 3 - 1, dx = ntaps == 1 ? 0 : tap - (tap / 3) * 3 - 1;
      const int C = S.C0 + S.C1;
      const bool normed = S.mode >= DS_NORM_CSUM;
      const bool per_plane = S.mode == DS_NORM_CSUM ? !S.joint : S.nrm_nseg == 3;
      const bool silu = S.silu != 0;
      const uint32_t sA = smem0 + stage * STAGE;
      const float2* tcol = tbl + c0 + sub * 8;
      const int pj = (it - it0) >> 1;
      const bool stampit = pstamp != nullptr && group == 0 && tg == 0 && pj < 4;
      if (stampit) pstamp[4 * pj] = clock64();
      mbar_wait(&bar_raw[stage], phase);                   // the raw fp32 tile has landed
      if (stampit) pstamp[4 * pj + 1] = clock64();
      float4 w[8][2];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t a = sA + (uint32_t)(rbase + 16 * k) * 256u + (uint32_t)sub * 32u;
        w[k][0] = ld_shared_v4f(a); w[k][1] = ld_shared_v4f(a + 16u);
      }
      asm volatile("bar.sync %0, 128;" ::"r"(2 + group) : "memory");   // every thread of the group holds its part: overwrite in place
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int r = rbase + 16 * k;
        const int ri = rowinfo[r];
        const int p = (ri >> 12) & 3, yy = ((ri >> 6) & 63) + dy, xx = (ri & 63) + dx;
        const bool ok = ri >= 0 && yy >= 0 && yy < (p == 0 ? g.res : g.t) && xx >= 0 && xx < g.res;
        uint4 hi = make_uint4(0u, 0u, 0u, 0u), lo = hi;
        if (ok) {
          float y[8] = {w[k][0].x, w[k][0].y, w[k][0].z, w[k][0].w, w[k][1].x, w[k][1].y, w[k][1].z, w[k][1].w};
          if (normed) {
            const float2* te = tcol + (size_t)tc_sp_index_local(T, per_plane, ri >> 16, p) * C;
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
              const float4 ad = *reinterpret_cast<const float4*>(te + i);     // a_i, d_i, a_{i+1}, d_{i+1}
              y[i] = fmaf(y[i], ad.x, ad.y); y[i + 1] = fmaf(y[i + 1], ad.z, ad.w);
            }
          }
          if (silu) {
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] = silu_tc(y[i]);
          }
          tc_pack_split8(y, hi, lo);
        }
        const uint32_t off = (uint32_t)r * 128u + (uint32_t)((sub ^ (r & 7)) << 4);
        st_shared_v4(sA + off, hi); st_shared_v4(sA + (uint32_t)(TC_BM * 128) + off, lo);
      }
      if (stampit) pstamp[4 * pj + 2] = clock64();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the MMA (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_full[stage]);
      if (stampit) pstamp[4 * pj + 3] = clock64();
    }
    if (++stage == NS) { stage = 0; phase ^= 1u; }
  }
}

// The tap-GEMM epilogue (warps 2-5 of the CTA): TMEM -> registers -> (+bias, +residual, GroupNorm sums | qkv operand
// split | split-K partial) -> HBM.  Shared by the one-tile-per-CTA kernel (PDL = true: it owns the grid-dependency
// wait / trigger) and the persistent chain kernel (PDL = false: several tiles per CTA, s_bias is reused).
// EW = number of epilogue warps (4: warps 2-5; 8: warps 2-9, warp group (warp-2)/4 takes chunks group, group+2, ...).
template <int BN, int EPI, bool PDL, int EW>
__device__ __forceinline__ void tc_epilogue(const TcConvParams& P, const Geo& g, const TcTile& T, int n0, int zidx,
                                            uint32_t tmem_base, float* s_bias, uint64_t* bar_acc, uint32_t acc_parity,
                                            long long* stamp, uint32_t stage_smem = 0u) {
  constexpr int CSTEP = 32 * (EW / 4);          // column stride between the chunks of one warp
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cfirst = ((warp - 2) >> 2) * 32;    // first chunk of this warp (0 when EW == 4)
  const int q = warp & 3;                       // TMEM lane quarter this warp may read
  const int row = q * 32 + lane;                // row of the tile
  int b, tok; tc_row_map(T, row, b, tok);
  const bool live = b < P.B;
  const size_t m = (size_t)b * g.L + tok;
  if (PDL) MTV_PDL_WAIT();                        // residual / statistics buffers belong to earlier kernels
  // Bias: small, touched once per step and evicted from L2 by the weight stream in between, i.e. a DRAM
  // miss (~2000 cycles) if loaded on demand per chunk — so it is staged in smem during the main loop.
  {
    const int te = threadIdx.x - 64;
    if (!PDL) asm volatile("bar.sync 1, %0;" ::"n"(32 * EW) : "memory");   // persistent caller: the previous tile's readers of s_bias are done
    if (te < BN) s_bias[te] = P.bias ? __ldg(P.bias + n0 + te) : 0.0f;
    asm volatile("bar.sync 1, %0;" ::"n"(32 * EW) : "memory");
  }
  // the residual of the first 32-column chunk is fetched while the MMAs still run
  const bool pre_res = EPI == 1 && live && P.resid;
  float rpre[32];
  if (pre_res) {
#pragma unroll
    for (int j = 0; j < 4; ++j) ld_global_nc_v8(P.resid + m * P.Cout + n0 + cfirst + 8 * j, rpre + 8 * j);
  }
  mbar_wait(bar_acc, acc_parity);
  if (PDL) MTV_PDL_TRIGGER();
  if (stamp && threadIdx.x == 64) stamp[5] = clock64();            // accumulator complete
  tc_fence_after();
  int p = 0, y = 0, x = 0;
  if (EPI == 3 || ((EPI == 1) && P.csum)) tc_decode_fast(T, tok, p, y, x);
  const int pl_stat = p;
  // bulk-store path: the tile's rows are consecutive rows of the output (always at the large levels; at the small ones
  // only when a tile is exactly one sample), the operand ring is idle (accumulator complete) and the op has an output map
  const bool ts = EPI != 2 && stage_smem != 0u && P.tma_store && (!T.small || T.spt == 1);
  const uint32_t sbase = stage_smem + (uint32_t)(warp - 2) * 4096u;
  const int ts_row = (int)((size_t)(T.small ? T.b0 : T.b0) * g.L + T.tok0) + q * 32 + (EPI == 0 ? zidx * P.B * g.L : 0);
  bool ts_pending = false;
#pragma unroll 1
  for (int c0 = cfirst; c0 < BN; c0 += CSTEP) {
    uint32_t r[32];
    __syncwarp();
    {
      uint32_t r2[32];
      tmem_ld32_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
      tmem_ld32_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(BN + c0), r2);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
    }
    const int n = n0 + c0;
    // the chunk's bias is fetched in one go: loads inside the store loop would serialise behind the
    // (possibly aliasing) stores and cost ~500 cycles each
    float4 bpre[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) bpre[j] = *reinterpret_cast<const float4*>(&s_bias[c0 + 4 * j]);
    float rnext[32];
    const bool have_next = pre_res && (c0 + CSTEP < BN);
    if (have_next) {
#pragma unroll
      for (int j = 0; j < 4; ++j) ld_global_nc_v8(P.resid + m * P.Cout + n + CSTEP + 8 * j, rnext + 8 * j);
    }
    if (!live) {
      // rows of samples beyond the batch (partial last tile of a small level): nothing to store
    } else if constexpr (EPI == 0) {
      float* dst = P.partial + ((size_t)zidx * ((size_t)P.B * g.L) + m) * P.Cout + n;
      float pv[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) pv[j] = __uint_as_float(r[j]);
      if (!ts) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) st_global_v8(dst + j, pv + j);
      } else {
        if (ts_pending) { if (lane == 0) tma_store_wait_read(); __syncwarp(); }
        tc_stage_row(sbase, lane, pv);
      }
    } else {
      float* dst = P.out + m * P.Cout + n;
      float fv[32];
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        {
          const float4 bv = bpre[j >> 2];
          v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
        }
        if constexpr (EPI == 1) {
          if (P.resid) { v.x += rpre[j]; v.y += rpre[j + 1]; v.z += rpre[j + 2]; v.w += rpre[j + 3]; }
        }
        if constexpr (EPI == 3) {
          if (P.resid_mode == RS_UP2) {
            const Geo gs = geo_down(g);
            const int ts = tc_plane_off(gs, p) + (y >> 1) * gs.res + (x >> 1);
            const float4 rv = __ldg(reinterpret_cast<const float4*>(P.resid + ((size_t)b * gs.L + ts) * P.Cout + n + j));
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
        fv[j] = v.x; fv[j + 1] = v.y; fv[j + 2] = v.z; fv[j + 3] = v.w;
      }
      if constexpr (EPI != 2) {
        if (ts) {
          if (ts_pending) { if (lane == 0) tma_store_wait_read(); __syncwarp(); }
          tc_stage_row(sbase, lane, fv);
        } else if (!(P.dbg_skip & 8)) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) st_global_v8(dst + j, fv + j);
        }
      }
      if constexpr (EPI == 2) {
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
      if ((EPI == 1 || EPI == 3) && P.csum) {   // uniform: statistics of the tensor just written, for the next GroupNorm
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(fv[j]);
      }
    }
    if (ts) {      // warp-uniform: every row of a bulk-stored tile is live
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) { tma_store_2d(&P.tmOut, sbase, n, ts_row); tma_store_commit(); }
      ts_pending = true;
    }
    if ((EPI == 1 || EPI == 3) && P.csum && !(P.dbg_skip & 4)) {
      float fv[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) fv[j] = __uint_as_float(r[j]);
      __syncwarp();
      if (!T.small || T.spt == 1) tc_csum_chunk32(fv, live, lane, P.csum + (((size_t)(live ? b : 0) * 3 + pl_stat) * P.Cout + n) * 2);
      else                        tc_csum_chunk(fv, live, lane, P.csum + (((size_t)(live ? b : 0) * 3 + pl_stat) * P.Cout + n) * 2);
    }
    if (have_next) {
#pragma unroll
      for (int j = 0; j < 32; ++j) rpre[j] = rnext[j];
    }
  }
  if (ts_pending && lane == 0) tma_store_wait_read();     // the staging tile must outlive the bulk store's reads
}

// EPI selects the epilogue at compile time (the row-per-lane epilogue is instruction-issue bound, so the
// paths a launch cannot take must not even be predicated off):
//   0 split-K partial tile, 1 bias [+ same-geometry residual] [+ GroupNorm sums], 2 qkv operand split,
//   3 residual through nearest-up / avg-pool geometry (up / down ResBlocks with identity skip)
//
// DIRECT: the A operand is produced in-kernel from the fp32 activation (tc_produce_A: GroupNorm affine + FiLM + SiLU +
// resample + concat, split to bf16) by warps 2-9 instead of being fetched by TMA from a pre-split copy; warp 0 then only
// streams the weights, and the four epilogue warps double as producers while the main loop runs.
template <int BN, int EPI, bool DIRECT>
__global__ void __launch_bounds__(DIRECT ? TC_THREADS_DIRECT : TC_THREADS, 1) k_conv_tc(const __grid_constant__ TcConvParams P) {
  constexpr int NS = tc_stages(BN);
  constexpr int STAGE = tc_stage_bytes(BN);
  // Stacked-N: W_hi and W_lo tiles are adjacent in smem, so ONE MMA with N = 2*BN computes
  // [A_hi*W_hi | A_hi*W_lo] into TMEM columns [0,BN) | [BN,2BN) and a second one adds A_lo*W_hi into [0,BN):
  // two MMAs and 14 KB of operand reads per k-step instead of three and 18 KB (the main loop is
  // shared-memory-bandwidth bound); the epilogue sums the two column groups.
  constexpr uint32_t IDESC = umma_idesc_bf16(TC_BM, BN);
  constexpr uint32_t IDESC2 = umma_idesc_bf16(TC_BM, 2 * BN);
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[NS], bar_empty[NS], bar_acc;
  __shared__ __align__(8) uint64_t bar_raw[DIRECT ? NS : 1];   // direct mode: the raw fp32 A tile of a stage has landed
  __shared__ uint32_t tmem_base_s;
  __shared__ long long s_stamp[8];
  __shared__ long long s_pstamp[DIRECT ? 16 : 1];  // diagnostics: producer-phase stamps of the first iterations
  __shared__ __align__(16) float s_bias[128];      // this CTA's BN bias values, fetched while the main loop runs
  __shared__ int s_rowinfo[DIRECT ? TC_BM : 1];    // direct mode: (sample, plane, y, x) of every tile row
  const bool dbg = g_tc_dbg != nullptr;
  long long g_t0 = 0;
  if (dbg && threadIdx.x == 0) { s_stamp[0] = clock64(); g_t0 = gtime_ns(); if (DIRECT) for (int i = 0; i < 16; ++i) s_pstamp[i] = 0; }
  mtv_prefetch_slice(P.pf0, P.pf1, P.pf_bytes, blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z),
                     gridDim.x * gridDim.y * gridDim.z);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B tiles need 1024-B alignment

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
    // full barrier: the TMA thread's expect_tx arrival (+ one arrival per producer warp of the owning group when DIRECT)
    for (int s = 0; s < NS; ++s) { mbar_init(&bar_full[s], DIRECT ? 5 : 1); mbar_init(&bar_empty[s], 1); }
    if (DIRECT) for (int s = 0; s < NS; ++s) mbar_init(&bar_raw[s], 1);
    mbar_init(&bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    prefetch_tmap(&P.tmA_hi[0]); if (!DIRECT) prefetch_tmap(&P.tmA_lo[0]);
    prefetch_tmap(&P.tmW_hi); prefetch_tmap(&P.tmW_lo);
  }
  if (warp == 1) {   // TMEM: BN fp32 accumulator columns x 128 lanes
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)(2 * BN)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (dbg && threadIdx.x == 0) s_stamp[1] = clock64();
  // bytes TMA delivers per stage (diagnostic skips: operand halves that are not fetched are not expected either)
  const uint32_t TX_BYTES = DIRECT ? (uint32_t)(2 * BN * 128)
                                   : (uint32_t)(((P.dbg_skip & 1) ? 0 : 2 * TC_BM * 128) + ((P.dbg_skip & 2) ? 0 : 2 * BN * 128));
  const int npre = min(NS, it1 - it0);
  const float2* tbl = reinterpret_cast<const float2*>(smem_raw + (smem0 - smem_u32(smem_raw)) + NS * STAGE);

  auto load_W = [&](int it, int stage) {
    if (!DIRECT && (P.dbg_skip & 2)) return;
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
  if constexpr (DIRECT) {
    // weights first (they do not depend on the previous kernel), then the whole CTA turns the producers' channel sums
    // into this tile's affine table while those loads are in flight
    if (threadIdx.x == 0) {
      for (int i = 0; i < npre; ++i) { mbar_expect_tx(&bar_full[i], TX_BYTES); load_W(it0 + i, i); }
    }
    if (threadIdx.x < TC_BM) {
      int b, tok; tc_row_map(T, (int)threadIdx.x, b, tok);
      int p, y, x; tc_decode_fast(T, tok, p, y, x);
      s_rowinfo[threadIdx.x] = b < P.B ? (((b - T.b0) << 16) | (p << 12) | (y << 6) | x) : -1;
    }
    MTV_PDL_WAIT();
    if (P.dseg[0].mode >= DS_NORM_CSUM)
      tc_build_table(P.dseg[0], g, T, P.B, const_cast<float2*>(tbl), reinterpret_cast<double*>(smem_raw + (smem0 - smem_u32(smem_raw))),
                     TC_THREADS_DIRECT);
    __syncthreads();
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
      if constexpr (!DIRECT) {
        if (elect_one()) {
          for (int i = 0; i < npre; ++i) {
            mbar_expect_tx(&bar_full[i], TX_BYTES);
            load_W(it0 + i, i);
          }
        }
        __syncwarp();
        MTV_PDL_WAIT();                // ... the activation operand does
      }
      int stage = 0; uint32_t phase = 0;
      for (int it = it0; it < it1; ++it) {
        if (it - it0 >= npre) mbar_wait(&bar_empty[stage], phase ^ 1u);
        if (dbg && lane == 0 && it == it1 - 1) s_stamp[2] = clock64();          // last stage request issued
        if (elect_one()) {
        if (it - it0 >= npre) {
          mbar_expect_tx(&bar_full[stage], TX_BYTES);
          load_W(it, stage);
        }
        if constexpr (!DIRECT) {
          if (!(P.dbg_skip & 1)) load_A(it, stage);
        } else {
          // raw fp32 tile of the source that holds this 64-channel chunk (channel concat = two tensors, two map sets)
          const uint32_t rb = smem_u32(&bar_raw[stage]);
          mbar_expect_tx(&bar_raw[stage], (uint32_t)(TC_BM * 256));
          const bool seg1 = it >= it_main;
          const DirectSeg& S = P.dseg[seg1 ? 1 : 0];
          const int tap = seg1 ? 0 : it / kch;
          const int c0 = seg1 ? (it - it_main) * TC_BK : (it - tap * kch) * TC_BK;
          const CUtensorMap* m0 = seg1 ? P.tmA2_hi : P.tmA_hi;
          const CUtensorMap* m1 = seg1 ? P.tmA2_lo : P.tmA_lo;
          if (c0 < S.C0) tc_load_A32(g, T, m0, seg1 ? 1 : P.taps, tap, c0, smem0 + stage * STAGE, rb);
          else           tc_load_A32(g, T, m1, seg1 ? 1 : P.taps, tap, c0 - S.C0, smem0 + stage * STAGE, rb);
        }
        }
        __syncwarp();
        if (++stage == NS) { stage = 0; phase ^= 1u; }
      }
    }
    // Dependents may be scheduled once the accumulator is complete: only this CTA's epilogue remains, so
    // the next kernel's launch latency and prologue overlap it without piling up many kernels deep.  Every
    // thread of the CTA triggers at that same point (the instruction's per-CTA semantics are not relied on).
    mbar_wait(&bar_acc, 0);
    MTV_PDL_TRIGGER();
  } else if (warp == 1) {
    // =============================== MMA issuer =================================
    // whole warp in the loop, one elected lane issues (see elect_one)
    {
      const uint64_t d0 = umma_desc_sw128(smem0);                   // descriptor of the ring's first byte; tiles are +offset/16
      int stage = 0; uint32_t phase = 0;
      for (int it = it0; it < it1; ++it) {
        mbar_wait(&bar_full[stage], phase);
        if (dbg && lane == 0 && it == it0) s_stamp[3] = clock64();       // first operands landed
        if (dbg && lane == 0 && it == it1 - 1) s_stamp[4] = clock64();   // last operands landed
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
    if constexpr (DIRECT) {
      const int pw = warp - 2;                    // producer warp 0..7: group = pw >> 2
      tc_produce_A<BN>(P, g, T, it0, it1, it_main, kch, smem0, bar_raw, bar_full, tbl, s_rowinfo, pw >> 2, (pw & 3) * 32 + lane, dbg ? s_pstamp : nullptr);
    }
    tc_epilogue<BN, EPI, true, 8>(P, g, T, n0, (int)blockIdx.z, tmem_base, s_bias, &bar_acc, 0u, dbg ? s_stamp : nullptr, smem0);
  }
  if (dbg && threadIdx.x == 64) s_stamp[6] = clock64();              // epilogue stores issued
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
      for (int i = 0; i < 7; ++i) rec[2 + i] = s_stamp[i];
      rec[9] = clock64(); rec[10] = g_t0; rec[11] = gtime_ns();
      unsigned int smid; asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
      rec[12] = smid; rec[13] = (long long)blockIdx.x | ((long long)blockIdx.y << 16) | ((long long)blockIdx.z << 32);
      if (DIRECT) {
        const unsigned int slot2 = atomicAdd(&g_tc_dbg_count, 1u);
        if (slot2 < g_tc_dbg_cap) {
          long long* r2 = g_tc_dbg + (size_t)slot2 * 16;
          r2[0] = (1ll << 61) | rec[0]; r2[1] = rec[1]; r2[2] = s_stamp[0];
          for (int i = 0; i < 12; ++i) r2[3 + i] = s_pstamp[i];
        }
      }
    }
  }
}

// split-K epilogue for the tensor-core path (fixed summation order).  A CTA owns 32 rows x 32
// channels: warp w = channel quad, lane = row, so the per-channel statistics reduce with three
// shuffles over aligned 8-row groups (never straddling a (sample, plane) boundary).
// sum of the split-K partial tiles (fixed order) + bias + residual for 4 channels of one output row
__device__ __forceinline__ float4 tc_splitk_sum(const TcConvParams& P, const Geo& g, size_t m, int n, int b, int p, int y, int x,
                                                const float4& bv0) {
  const size_t M = (size_t)P.B * g.L;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* pp = P.partial + m * P.Cout + n;
  const size_t zstride = M * P.Cout;
  int z = 0;
  for (; z + 8 <= P.ksplit; z += 8) {           // 8 independent loads in flight, fixed summation order
    float4 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __ldcs(reinterpret_cast<const float4*>(pp + (size_t)(z + i) * zstride));
#pragma unroll
    for (int i = 0; i < 8; ++i) { s.x += v[i].x; s.y += v[i].y; s.z += v[i].z; s.w += v[i].w; }
  }
  for (; z < P.ksplit; ++z) {
    const float4 v = __ldcs(reinterpret_cast<const float4*>(pp + (size_t)z * zstride));
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  s.x += bv0.x; s.y += bv0.y; s.z += bv0.z; s.w += bv0.w;
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

__device__ __forceinline__ void tc_splitk_unit(const TcConvParams& P, int mblk, int nblk) {
  const Geo g = P.geo;
  const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
  const size_t m = (size_t)mblk * 32 + lane;
  const int n = nblk * 32 + wq * 4;
  float4 bv0 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (P.bias) bv0 = __ldg(reinterpret_cast<const float4*>(P.bias + n));     // cold line: issue first
  const int b = (int)(m / g.L), tok = (int)(m - (size_t)b * g.L);
  int p = 0, y = 0, x = 0;
  tc_decode_tok(g, tok, p, y, x);
  const float4 s = tc_splitk_sum(P, g, m, n, b, p, y, x, bv0);
  *reinterpret_cast<float4*>(P.out + m * P.Cout + n) = s;
  if (P.csum) {
    float v[4] = {s.x, s.y, s.z, s.w}, q[4] = {s.x * s.x, s.y * s.y, s.z * s.z, s.w * s.w};
#pragma unroll
    for (int off = 1; off < 8; off <<= 1)
#pragma unroll
      for (int i = 0; i < 4; ++i) { v[i] += __shfl_xor_sync(0xffffffffu, v[i], off); q[i] += __shfl_xor_sync(0xffffffffu, q[i], off); }
    if ((lane & 7) == 0) {
      double* dst = P.csum + (((size_t)b * 3 + p) * P.Cout + n) * 2;
#pragma unroll
      for (int i = 0; i < 4; ++i) { atomicAdd(dst + 2 * i, (double)v[i]); atomicAdd(dst + 2 * i + 1, (double)q[i]); }
    }
  }
}

// Split-K reduction FUSED with the consumer's GroupNorm + apply, for levels where one sample has <= 128 tokens: a CTA owns
// every token of sample b for CU channels (whole GroupNorm groups), so the statistics of the reduced tensor are CTA-local
// and the next op's operand  y = silu?(GN(x) [FiLM])  -> split bf16  can be written in the same pass — no channel-sum
// atomics, no stand-alone apply launch.  Still writes the fp32 tensor (residual / skip consumers) and its per-channel sums.
// Thread t: channel quad t % (CU/4), rows t / (CU/4) + k * (256 / (CU/4)).
template <int CU>
__global__ void __launch_bounds__(256) k_tc_splitk_reduce_apply(const __grid_constant__ TcConvParams P) {
  constexpr int QL = CU / 4;                 // float4 lanes per row
  constexpr int RPP = 256 / QL;              // rows per pass (64 or 32)
  constexpr int RPW = 32 / QL;               // rows per warp and pass (8 or 4): never straddles a plane
  constexpr int MAXPASS = 128 / RPP;
  MTV_PDL_TRIGGER();
  mtv_prefetch_slice(P.fa.pf0, P.fa.pf1, P.fa.pf_bytes, blockIdx.x + gridDim.x * blockIdx.y, gridDim.x * gridDim.y);
  const Geo g = P.geo;
  const int L = g.L, C = P.Cout;
  const int b = blockIdx.y, n = blockIdx.x * CU + (threadIdx.x % QL) * 4;
  const int r0 = threadIdx.x / QL, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int npass = (L + RPP - 1) / RPP;
  __shared__ float s_part[128 / RPW][CU][2];       // per aligned row group: channel sums, sums of squares
  __shared__ float s_a[3][CU], s_d[3][CU];
  // static per-channel parameters do not depend on the producer: fetch before the dependency resolves
  float4 bv0 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (P.bias) bv0 = __ldg(reinterpret_cast<const float4*>(P.bias + n));
  float pg = 0.f, pb = 0.f, psc = 0.f, psh = 0.f;
  if (threadIdx.x < CU) {
    const int c = blockIdx.x * CU + threadIdx.x;
    pg = __ldg(P.fa.gamma + c); pb = __ldg(P.fa.beta + c);
  }
  MTV_PDL_WAIT();
  if (threadIdx.x < CU && P.fa.film) {
    const int c = blockIdx.x * CU + threadIdx.x;
    const float* f = P.fa.film + (size_t)b * P.fa.film_stride;
    psc = __ldg(f + c); psh = __ldg(f + C + c);
  }
  float4 val[MAXPASS];
#pragma unroll
  for (int k = 0; k < MAXPASS; ++k) {
    const int r = r0 + k * RPP;
    if (k < npass && r < L) {
      int p, y, x; tc_decode_tok(g, r, p, y, x);
      const size_t m = (size_t)b * L + r;
      const float4 s = tc_splitk_sum(P, g, m, n, b, p, y, x, bv0);
      *reinterpret_cast<float4*>(P.out + m * C + n) = s;
      val[k] = s;
      float v[4] = {s.x, s.y, s.z, s.w}, q[4] = {s.x * s.x, s.y * s.y, s.z * s.z, s.w * s.w};
#pragma unroll
      for (int off = QL; off < 32; off <<= 1)
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[i] += __shfl_xor_sync(0xffffffffu, v[i], off); q[i] += __shfl_xor_sync(0xffffffffu, q[i], off); }
      if (lane < QL) {
        const int grp = (k * RPP) / RPW + warp;      // aligned row group index = first row / RPW
#pragma unroll
        for (int i = 0; i < 4; ++i) { s_part[grp][lane * 4 + i][0] = v[i]; s_part[grp][lane * 4 + i][1] = q[i]; }
      }
    }
  }
  __syncthreads();
  // per (plane, channel) sums in fixed order -> csum (complete: plain stores) -> group statistics -> affine (a, d)
  const int nxy = g.res * g.res, npl = g.t * g.res;
  __shared__ double s_cs[3][CU][2];
  if (threadIdx.x < 3 * CU) {
    const int p = threadIdx.x / CU, c = threadIdx.x - p * CU;
    const int g0 = (p == 0 ? 0 : nxy + (p - 1) * npl) / RPW, g1 = (p == 0 ? nxy : nxy + p * npl) / RPW;
    double sm = 0.0, sq = 0.0;
    for (int gi = g0; gi < g1; ++gi) { sm += (double)s_part[gi][c][0]; sq += (double)s_part[gi][c][1]; }
    s_cs[p][c][0] = sm; s_cs[p][c][1] = sq;
    if (P.csum) {
      double* dst = P.csum + (((size_t)b * 3 + p) * C + blockIdx.x * CU + c) * 2;
      dst[0] = sm; dst[1] = sq;
    }
  }
  __syncthreads();
  __shared__ double s_st[3][CU][2];                // (rstd, mean) of the group of channel c in plane p
  if (threadIdx.x < 3 * CU) {
    const int p = threadIdx.x / CU, c = threadIdx.x - p * CU;
    const int cpg = C / 32;
    const int cg0 = (c / cpg) * cpg;               // CU is a multiple of the group width: the group lies inside this CTA
    double sm = 0.0, sq = 0.0;
    for (int ci = 0; ci < cpg; ++ci) {
      if (P.fa.joint) { for (int pp = 0; pp < 3; ++pp) { sm += s_cs[pp][cg0 + ci][0]; sq += s_cs[pp][cg0 + ci][1]; } }
      else { sm += s_cs[p][cg0 + ci][0]; sq += s_cs[p][cg0 + ci][1]; }
    }
    const double cnt = (double)cpg * (P.fa.joint ? (double)L : (double)(p == 0 ? nxy : npl));
    const double mean = sm / cnt;
    double var = sq / cnt - mean * mean; var = var < 0.0 ? 0.0 : var;
    s_st[p][c][0] = rsqrt(var + 1e-5); s_st[p][c][1] = mean;
  }
  __syncthreads();
  if (threadIdx.x < CU) {                          // gamma / beta / FiLM of channel c live in thread c's registers
    const int c = threadIdx.x;
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      double a = s_st[p][c][0] * (double)pg;
      double d = (double)pb - s_st[p][c][1] * a;
      if (P.fa.film) { const double sc = 1.0 + (double)psc; a *= sc; d = d * sc + (double)psh; }
      s_a[p][c] = (float)a; s_d[p][c] = (float)d;
    }
  }
  __syncthreads();
  const int cl = (threadIdx.x % QL) * 4;
#pragma unroll
  for (int k = 0; k < MAXPASS; ++k) {
    const int r = r0 + k * RPP;
    if (k < npass && r < L) {
      const int p = r < nxy ? 0 : (r < nxy + npl ? 1 : 2);
      const float4 na = *reinterpret_cast<const float4*>(&s_a[p][cl]), nd = *reinterpret_cast<const float4*>(&s_d[p][cl]);
      float4 v = val[k];
      v.x = fmaf(v.x, na.x, nd.x); v.y = fmaf(v.y, na.y, nd.y); v.z = fmaf(v.z, na.z, nd.z); v.w = fmaf(v.w, na.w, nd.w);
      if (P.fa.silu) { v.x = silu_tc(v.x); v.y = silu_tc(v.y); v.z = silu_tc(v.z); v.w = silu_tc(v.w); }
      tc_store_split(v, P.fa.hi, P.fa.lo, ((size_t)b * L + r) * C + n);
    }
  }
}
__global__ void __launch_bounds__(256) k_tc_splitk_epilogue(const __grid_constant__ TcConvParams P) {
  MTV_PDL_TRIGGER();
  MTV_PDL_WAIT();
  tc_splitk_unit(P, (int)blockIdx.x, (int)blockIdx.y);
}

// ------------------------------------------------------------------ persistent chain kernel
// k_chain executes a run of {apply, tap-GEMM, split-K reduction} sub-ops that the launch plan would otherwise issue as
// separate kernels.  Grid = min(#SMs, work units) CTAs of 256 threads, one per SM (208 KB of shared memory), all
// co-resident, so a sense-free counting barrier in global memory can stand in for each kernel boundary:
//   * the operand ring, mbarriers and the 256-column TMEM allocation are set up once per chain, not once per op;
//   * every sub-op is a persistent loop `unit = blockIdx.x; unit < units; unit += gridDim.x`;
//   * weights do not depend on earlier sub-ops: the W halves of the next GEMM's first stages are requested BEFORE the
//     barrier, so their HBM latency overlaps the barrier and the sub-ops in between.
// GEMM roles: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = epilogue (tc_epilogue), warps 6-7 idle.
constexpr int CH_THREADS = 256;
constexpr int CH_PIPE_BYTES = 4 * tc_stage_bytes(64);
static_assert(tc_stages(64) * tc_stage_bytes(64) == CH_PIPE_BYTES && tc_stages(128) * tc_stage_bytes(128) == CH_PIPE_BYTES, "ring size");
constexpr int CH_AFF_FLOATS = 4096;                      // a[C] | d[C] of the apply sub-op, C <= 2048
constexpr int CH_SMEM_BYTES = 1024 + CH_PIPE_BYTES + CH_AFF_FLOATS * 4;
constexpr int CH_TMEM_COLS = 256;

struct ChainGemmGeo { int kch, it_main, it_total, ks, per, mt, nt, ntiles; };
__host__ __device__ inline ChainGemmGeo chain_gemm_geo(const TcConvParams& P, int BN) {
  ChainGemmGeo G;
  G.kch = P.Cin / TC_BK; G.it_main = P.taps * G.kch; G.it_total = G.it_main + P.Cin2 / TC_BK;
  G.ks = P.ksplit > 1 ? P.ksplit : 1; G.per = (G.it_total + G.ks - 1) / G.ks;
  const int M = P.B * P.geo.L;
  G.mt = P.geo.L > TC_BM ? M / TC_BM : (P.B + (TC_BM / P.geo.L) - 1) / (TC_BM / P.geo.L);
  G.nt = P.Cout / BN; G.ntiles = G.mt * G.nt * G.ks;
  return G;
}
__host__ __device__ inline int chain_apply_units(const ApplyParams& P) {
  const int nxy = (P.geo.res * P.geo.res + P.chunk_tokens - 1) / P.chunk_tokens;
  const int npl = (P.geo.t * P.geo.res + P.chunk_tokens - 1) / P.chunk_tokens;
  return (nxy + 2 * npl) * P.B;
}
int chain_max_grid_units(const ChainOp& op) {
  if (op.type == CH_APPLY) return chain_apply_units(op.apply);
  if (op.type == CH_GEMM) return chain_gemm_geo(op.conv, op.conv.bn).ntiles;
  return (op.conv.B * op.conv.geo.L / 32) * (op.conv.Cout / 32);
}

// W halves (hi | lo, adjacent: the stacked-N operand) of K-iteration `it` of the tile at output channel n0
__device__ __forceinline__ void chain_load_W(const TcConvParams& PG, int BN, const ChainGemmGeo& G, int Cout, int it, int n0,
                                             uint32_t sW_hi, uint32_t fb) {
  const uint32_t sW_lo = sW_hi + (uint32_t)BN * 128u;
  if (it >= G.it_main) {
    const int c2 = (it - G.it_main) * TC_BK;
    tma_load_2d(sW_hi, &PG.tmW2_hi, fb, c2, n0);
    tma_load_2d(sW_lo, &PG.tmW2_lo, fb, c2, n0);
  } else {
    const int tap = it / G.kch, c0 = (it - tap * G.kch) * TC_BK;
    tma_load_2d(sW_hi, &PG.tmW_hi, fb, c0, tap * Cout + n0);
    tma_load_2d(sW_lo, &PG.tmW_lo, fb, c0, tap * Cout + n0);
  }
}

// Thread 0, pipeline idle and barriers freshly initialised: request the W halves of the first stages of this CTA's first
// tile of GEMM sub-op `PG` (global-memory descriptor).  Returns the number of stages requested.
__device__ __forceinline__ int chain_prefetch_W(const TcConvParams& PG, uint32_t smem0, uint64_t* bar_full) {
  const int BN = PG.bn;
  const ChainGemmGeo G = chain_gemm_geo(PG, BN);
  const int tile = blockIdx.x;
  if (tile >= G.ntiles) return 0;
  const int r = tile / G.mt, n0 = (r % G.nt) * BN, z = r / G.nt;
  const int it0 = z * G.per, it1 = min(G.it_total, it0 + G.per);
  const int NS = tc_stages(BN), STAGE = tc_stage_bytes(BN);
  const int npre = min(NS, it1 - it0);
  const int Cout = PG.Cout;
  prefetch_tmap(&PG.tmA_hi[0]); prefetch_tmap(&PG.tmA_lo[0]);
  for (int i = 0; i < npre; ++i) {
    mbar_expect_tx(&bar_full[i], (uint32_t)STAGE);
    chain_load_W(PG, BN, G, Cout, it0 + i, n0, smem0 + i * STAGE + 2 * TC_BM * 128, smem_u32(&bar_full[i]));
  }
  return npre;
}

// PG: the op in global memory (tensor maps); P: its scalar fields staged in shared memory.
template <int BN>
__device__ __noinline__ void chain_gemm(const TcConvParams& PG, const TcConvParams& P, uint32_t smem0, uint64_t* bar_full, uint64_t* bar_empty,
                           uint64_t* bar_acc, uint64_t* bar_accfree, uint32_t tmem_base, float* s_bias, int npre) {
  constexpr int NS = tc_stages(BN);
  constexpr int STAGE = tc_stage_bytes(BN);
  constexpr uint32_t IDESC = umma_idesc_bf16(TC_BM, BN);
  constexpr uint32_t IDESC2 = umma_idesc_bf16(TC_BM, 2 * BN);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= 6) return;
  const Geo g = P.geo;
  const ChainGemmGeo G = chain_gemm_geo(P, BN);
  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      const int taps = P.taps, Cout = P.Cout;
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < G.ntiles; tile += gridDim.x) {
        const int mx = tile % G.mt, r = tile / G.mt, n0 = (r % G.nt) * BN, z = r / G.nt;
        const TcTile T = tc_tile(g, mx);
        const int it0 = z * G.per, it1 = min(G.it_total, it0 + G.per);
        for (int it = it0; it < it1; ++it) {
          const uint32_t sA_hi = smem0 + stage * STAGE, sA_lo = sA_hi + TC_BM * 128, sW_hi = sA_lo + TC_BM * 128;
          const uint32_t fb = smem_u32(&bar_full[stage]);
          if (npre > 0) {
            --npre;                                   // W of this stage was requested before the grid barrier
          } else {
            mbar_wait(&bar_empty[stage], phase ^ 1u);
            mbar_expect_tx(&bar_full[stage], (uint32_t)STAGE);
            chain_load_W(PG, BN, G, Cout, it, n0, sW_hi, fb);
          }
          if (it >= G.it_main) {                      // second K-segment: 1x1 conv of the skip operand
            tc_load_A(g, T, PG.tmA2_hi, PG.tmA2_lo, 1, 0, (it - G.it_main) * TC_BK, sA_hi, sA_lo, fb);
          } else {
            const int tap = it / G.kch, c0 = (it - tap * G.kch) * TC_BK;
            tc_load_A(g, T, PG.tmA_hi, PG.tmA_lo, taps, tap, c0, sA_hi, sA_lo, fb);
          }
          if (++stage == NS) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer =================================
    // whole warp in the loop, one elected lane issues (see elect_one)
    {
      const uint64_t d0 = umma_desc_sw128(smem0);
      int stage = 0; uint32_t phase = 0, tl = 0;
      for (int tile = blockIdx.x; tile < G.ntiles; tile += gridDim.x) {
        const int z = (tile / G.mt) / G.nt;
        const int it0 = z * G.per, it1 = min(G.it_total, it0 + G.per);
        mbar_wait(bar_accfree, (tl & 1u) ^ 1u);       // the epilogue has read the previous tile out of TMEM
        tc_fence_after();
        for (int it = it0; it < it1; ++it) {
          mbar_wait(&bar_full[stage], phase);
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
            umma_commit(&bar_empty[stage]);
          }
          __syncwarp();
          if (++stage == NS) { stage = 0; phase ^= 1u; }
        }
        if (elect_one()) umma_commit(bar_acc);
        __syncwarp();
        ++tl;
      }
    }
  } else {
    // =============================== epilogue ===================================
    const int epi = P.ksplit > 1 ? 0 : (P.qkv_heads ? 2 : ((P.resid && P.resid_mode != RS_NONE) ? 3 : 1));
    uint32_t tl = 0;
    for (int tile = blockIdx.x; tile < G.ntiles; tile += gridDim.x) {
      const int mx = tile % G.mt, r = tile / G.mt, n0 = (r % G.nt) * BN, z = r / G.nt;
      const TcTile T = tc_tile(g, mx);
      switch (epi) {
        case 0: tc_epilogue<BN, 0, false, 4>(P, g, T, n0, z, tmem_base, s_bias, bar_acc, tl & 1u, nullptr); break;
        case 1: tc_epilogue<BN, 1, false, 4>(P, g, T, n0, z, tmem_base, s_bias, bar_acc, tl & 1u, nullptr); break;
        case 2: tc_epilogue<BN, 2, false, 4>(P, g, T, n0, z, tmem_base, s_bias, bar_acc, tl & 1u, nullptr); break;
        default: tc_epilogue<BN, 3, false, 4>(P, g, T, n0, z, tmem_base, s_bias, bar_acc, tl & 1u, nullptr); break;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_accfree);
      ++tl;
    }
  }
}

__device__ __noinline__ void chain_apply(const ApplyParams& P, float* s_aff, double* s_mean, double* s_rstd) {
  mtv_prefetch_slice(P.pf0, P.pf1, P.pf_bytes, blockIdx.x, gridDim.x);
  const Geo g = P.geo;
  const int nxy = (g.res * g.res + P.chunk_tokens - 1) / P.chunk_tokens, npl = (g.t * g.res + P.chunk_tokens - 1) / P.chunk_tokens;
  const int per_b = nxy + 2 * npl, units = per_b * P.B;
  bool first = true;
  for (int u = blockIdx.x; u < units; u += gridDim.x) {
    if (!first) __syncthreads();                     // the previous unit's readers of the affine table are done
    first = false;
    const int b = u / per_b;
    int r = u - b * per_b, p = 0;
    if (r >= nxy) { r -= nxy; p = 1; if (r >= npl) { r -= npl; p = 2; } }
    apply_norm_unit(P, s_aff, s_mean, s_rstd, r, p, b);
  }
}

__device__ __noinline__ void chain_reduce(const TcConvParams& P) {
  const int mb = (P.B * P.geo.L) / 32, nb = P.Cout / 32;
  for (int u = blockIdx.x; u < mb * nb; u += gridDim.x) tc_splitk_unit(P, u % mb, u / mb);
}

// Thread 0 of every CTA, between two __syncthreads: all `gridDim.x` CTAs are co-resident (host guarantees grid <= #SMs at one
// CTA per SM), so spinning is safe; bounded anyway (trap, not hang).
__device__ __forceinline__ void chain_grid_barrier(unsigned int* ctr, unsigned int target) {
  __threadfence();
  atomicAdd(ctr, 1u);
  unsigned int v, spins = 0;
  long long t0 = 0;
  for (;;) {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    if (v >= target) break;
    if ((++spins & 0xffu) == 0) {
      const long long t = clock64();
      if (t0 == 0) t0 = t; else if (t - t0 > (1ll << 32)) __trap();
    }
  }
  __threadfence();
}

template <typename S>
__device__ __forceinline__ void chain_copy_words(S* dst, const S* src, size_t from_byte) {
  const uint32_t* s = reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(src) + from_byte);
  uint32_t* d = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(dst) + from_byte);
  const int n = (int)((sizeof(S) - from_byte) / 4);
  for (int i = threadIdx.x; i < n; i += blockDim.x) d[i] = __ldg(s + i);
}

__global__ void __launch_bounds__(CH_THREADS, 1) k_chain(const ChainOp* __restrict__ ops, int nops, unsigned int* counters) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[4], bar_empty[4], bar_acc, bar_accfree;
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_bias[128];
  __shared__ double s_mean[32], s_rstd[32];
  __shared__ TcConvParams s_cp;                    // scalar fields only (the tensor maps are used from global memory)
  __shared__ ApplyParams s_ap;
  static_assert(sizeof(TcConvParams) % 4 == 0 && sizeof(ApplyParams) % 4 == 0 && offsetof(TcConvParams, Cin2) % 4 == 0, "word copies");

  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  float* s_aff = reinterpret_cast<float*>(smem_raw + (smem0 - smem_u32(smem_raw)) + CH_PIPE_BYTES);
  const bool dbg = g_tc_dbg != nullptr;

  auto init_bars = [&]() {
    for (int s = 0; s < 4; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
    mbar_init(&bar_acc, 1); mbar_init(&bar_accfree, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  };
  auto next_gemm_prefetch = [&](int from) -> int {
    for (int j = from; j < nops; ++j)
      if (ops[j].type == CH_GEMM) return chain_prefetch_W(ops[j].conv, smem0, bar_full);
    return 0;
  };
  if (tid == 0) init_bars();
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)CH_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  int npre = 0;                                    // thread 0: stages of the next GEMM whose W halves are already in flight
  if (tid == 0) npre = next_gemm_prefetch(0);      // weights never depend on the previous kernel
  MTV_PDL_WAIT();

  for (int i = 0; i < nops; ++i) {
    const ChainOp& op = ops[i];
    const int type = op.type;
    long long t_start = 0;
    if (dbg && tid == 0) t_start = gtime_ns();
    if (i == nops - 1) MTV_PDL_TRIGGER();          // only this CTA's last sub-op remains
    if (type == CH_APPLY) chain_copy_words(&s_ap, &op.apply, 0);
    else                  chain_copy_words(&s_cp, &op.conv, offsetof(TcConvParams, Cin2));
    __syncthreads();
    if (type == CH_APPLY) {
      chain_apply(s_ap, s_aff, s_mean, s_rstd);
    } else if (type == CH_GEMM) {
      if (s_cp.bn == 64) chain_gemm<64>(op.conv, s_cp, smem0, bar_full, bar_empty, &bar_acc, &bar_accfree, tmem_base, s_bias, npre);
      else               chain_gemm<128>(op.conv, s_cp, smem0, bar_full, bar_empty, &bar_acc, &bar_accfree, tmem_base, s_bias, npre);
      npre = 0;
    } else {
      chain_reduce(s_cp);
    }
    if (i == nops - 1) {
      if (dbg && tid == 0) {
        const unsigned int slot = atomicAdd(&g_tc_dbg_count, 1u);
        if (slot < g_tc_dbg_cap) {
          long long* rec = g_tc_dbg + (size_t)slot * 16;
          rec[0] = (1ll << 62) | ((long long)type << 48) | ((long long)i << 32) | (long long)blockIdx.x;
          rec[1] = t_start; rec[2] = gtime_ns(); rec[3] = rec[2]; rec[4] = (long long)ops; rec[5] = nops;
        }
      }
      break;
    }
    asm volatile("fence.proxy.async;" ::: "memory");   // this thread's global stores vs. later TMA (async-proxy) reads by other CTAs
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      long long t_work = 0;
      if (dbg) t_work = gtime_ns();
      if (type == CH_GEMM) {                       // pipeline drained: fresh barriers for the next GEMM, its first W requests go out now
        tc_fence_after();
        for (int s = 0; s < 4; ++s) { mbar_inval(&bar_full[s]); mbar_inval(&bar_empty[s]); }
        mbar_inval(&bar_acc); mbar_inval(&bar_accfree);
        init_bars();
        npre = next_gemm_prefetch(i + 1);
      }
      chain_grid_barrier(counters, (unsigned int)(i + 1) * gridDim.x);
      asm volatile("fence.proxy.async;" ::: "memory");
      if (dbg) {
        const unsigned int slot = atomicAdd(&g_tc_dbg_count, 1u);
        if (slot < g_tc_dbg_cap) {
          long long* rec = g_tc_dbg + (size_t)slot * 16;
          rec[0] = (1ll << 62) | ((long long)type << 48) | ((long long)i << 32) | (long long)blockIdx.x;
          rec[1] = t_start; rec[2] = t_work; rec[3] = gtime_ns(); rec[4] = (long long)ops; rec[5] = nops;
        }
      }
    }
    __syncthreads();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)CH_TMEM_COLS) : "memory");
  }
  if (tid == 0 && nops > 1) {                      // the last CTA out re-arms the barrier for the next launch of this chain
    __threadfence();
    const unsigned int prev = atomicAdd(counters + 1, 1u);
    if (prev == gridDim.x - 1) { atomicExch(counters, 0u); atomicExch(counters + 1, 0u); }
  }
}

cudaError_t launch_chain(const ChainLaunch& L, cudaStream_t s) {
  if (L.nops < 1 || L.grid < 1) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(k_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  e = launch_kc(PDL_CLASS_CONV_TC, k_chain, dim3(L.grid), dim3(CH_THREADS), (size_t)CH_SMEM_BYTES, s, L.ops, L.nops, L.counters);
  if (e != cudaSuccess) return e;
  return cudaGetLastError();
}

cudaError_t launch_conv_tc(const TcConvParams& P, cudaStream_t s) {
  const int BN = P.bn;
  if ((BN != 64 && BN != 128) || P.Cout % BN || P.Cin % TC_BK || P.Cin2 % TC_BK) return cudaErrorInvalidValue;
  if (P.geo.L > TC_BM ? (P.geo.L % TC_BM != 0) : (TC_BM % P.geo.L != 0)) return cudaErrorInvalidValue;
  const int M = P.B * P.geo.L;
  const int tiles = P.geo.L > TC_BM ? M / TC_BM : (P.B + (TC_BM / P.geo.L) - 1) / (TC_BM / P.geo.L);
  dim3 grid(tiles, P.Cout / BN, P.ksplit > 1 ? P.ksplit : 1);
  cudaError_t e = cudaSuccess;
  const int epi = P.ksplit > 1 ? 0 : (P.qkv_heads ? 2 : ((P.resid && P.resid_mode != RS_NONE) ? 3 : 1));
#define MTV_TC_LAUNCH(BN_, EPI_)                                                                                   \
  do {                                                                                                             \
    if (P.direct) {                                                                                                \
      e = cudaFuncSetAttribute(k_conv_tc<BN_, EPI_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes_direct(BN_)); \
      if (e != cudaSuccess) return e;                                                                              \
      e = launch_kc(PDL_CLASS_CONV_TC, k_conv_tc<BN_, EPI_, true>, grid, dim3(TC_THREADS_DIRECT), (size_t)tc_smem_bytes_direct(BN_), s, P); \
    } else {                                                                                                       \
      e = cudaFuncSetAttribute(k_conv_tc<BN_, EPI_, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes(BN_)); \
      if (e != cudaSuccess) return e;                                                                              \
      e = launch_kc(PDL_CLASS_CONV_TC, k_conv_tc<BN_, EPI_, false>, grid, dim3(TC_THREADS), (size_t)tc_smem_bytes(BN_), s, P); \
    }                                                                                                              \
    if (e != cudaSuccess) return e;                                                                                \
  } while (0)
  if (BN == 64) {
    switch (epi) { case 0: MTV_TC_LAUNCH(64, 0); break; case 1: MTV_TC_LAUNCH(64, 1); break;
                   case 2: MTV_TC_LAUNCH(64, 2); break; default: MTV_TC_LAUNCH(64, 3); break; }
  } else {
    switch (epi) { case 0: MTV_TC_LAUNCH(128, 0); break; case 1: MTV_TC_LAUNCH(128, 1); break;
                   case 2: MTV_TC_LAUNCH(128, 2); break; default: MTV_TC_LAUNCH(128, 3); break; }
  }
#undef MTV_TC_LAUNCH
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (P.ksplit > 1 && P.fa.hi) {
    const int cu = (P.Cout / 32) > 16 ? 32 : 16;     // whole GroupNorm groups per CTA (host checked divisibility)
    dim3 rgrid(P.Cout / cu, P.B);
    cudaError_t le_ = cu == 16 ? launch_kc(PDL_CLASS_REDUCE, k_tc_splitk_reduce_apply<16>, rgrid, dim3(256), (size_t)0, s, P)
                               : launch_kc(PDL_CLASS_REDUCE, k_tc_splitk_reduce_apply<32>, rgrid, dim3(256), (size_t)0, s, P);
    if (le_ != cudaSuccess) return le_;
    e = cudaGetLastError();
  } else if (P.ksplit > 1) {
    dim3 rgrid(M / 32, P.Cout / 32);
    { cudaError_t le_ = launch_kc(PDL_CLASS_REDUCE, k_tc_splitk_epilogue, dim3(rgrid), dim3(256), (size_t)(0), s, P); if (le_ != cudaSuccess) return le_; }
    e = cudaGetLastError();
  }
  return e;
}

}  // namespace mtv
