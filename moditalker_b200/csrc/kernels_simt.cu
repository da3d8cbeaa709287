// kernels_simt.cu — fp32 CUDA-core kernels of the MToV UNet hot path (sm_100a).
//
// These are the exact-arithmetic (fp32 FMA, fp32/fp64 reductions) kernels: the
// GroupNorm statistics, the timestep-embedding MLPs, the fused
// "norm-affine + SiLU + resample -> 3x3/1x1 tap-GEMM -> bias + residual" kernel for
// shapes with too few tokens for a tensor-core tile (pyramid levels 2-3, where the
// op is a weight stream), flash-style attention, the DDIM update, layout packers.
// The tensor-core (tcgen05) versions of the two big contractions live in
// kernels_tc.cu and are cross-checked against these in tests/test_kernels_gpu.py.
#include "mtv_kernels.cuh"
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>

namespace mtv {

int g_mtv_use_pdl = [] { const char* e = getenv("MTV_PDL"); return e ? atoi(e) : 7; }();

// ------------------------------------------------------------------ token geometry
__device__ __forceinline__ void decode_tok(const Geo& g, int tok, int& p, int& y, int& x) {
  const int nxy = g.res * g.res;
  if (tok < nxy) {
    p = 0; y = tok / g.res; x = tok - y * g.res;
  } else {
    int r = tok - nxy; const int np = g.t * g.res;
    p = 1; if (r >= np) { p = 2; r -= np; }
    y = r / g.res; x = r - y * g.res;
  }
}
__device__ __forceinline__ int plane_off(const Geo& g, int p) {
  return p == 0 ? 0 : g.res * g.res + (p - 1) * g.t * g.res;
}
__device__ __forceinline__ int plane_h(const Geo& g, int p) { return p == 0 ? g.res : g.t; }

__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + __expf(-v)); }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// ------------------------------------------------------------------ tap-GEMM (conv3x3 / 1x1)
// out[m][n] = bias[n] + resid[m][n] + sum_seg sum_tap sum_c  T_seg(src)[nbr(m,tap)][c] * W_seg[tap][c][n]
// Reference semantics: ResBlock._forward (unet.py:178-207), AttentionBlock qkv /
// proj_out 1x1 convs (unet.py:251-254, 297-300), stem (unet.py:714), head (unet.py:971-975).
__device__ __forceinline__ float epilogue_add(const ConvParams& P, int b, int p, int y, int x, int tok, int n) {
  float v = P.bias ? __ldg(P.bias + n) : 0.0f;
  if (P.resid) {
    if (P.resid_mode == RS_NONE) {
      v += __ldg(P.resid + ((size_t)b * P.geo.L + tok) * P.Cout + n);
    } else if (P.resid_mode == RS_UP2) {          // x_upd = nearest x2 (unet.py:183, 549-554)
      const Geo gs = geo_down(P.geo);
      const int ts = plane_off(gs, p) + (y >> 1) * gs.res + (x >> 1);
      v += __ldg(P.resid + ((size_t)b * gs.L + ts) * P.Cout + n);
    } else {                                       // x_upd = AvgPool2d(2,2) (unet.py:183, 589-594)
      const Geo gs = geo_up(P.geo);
      const int t0 = plane_off(gs, p) + (2 * y) * gs.res + 2 * x;
      const float* r = P.resid + ((size_t)b * gs.L) * P.Cout + n;
      const float s = __ldg(r + (size_t)t0 * P.Cout) + __ldg(r + (size_t)(t0 + 1) * P.Cout) +
                      __ldg(r + (size_t)(t0 + gs.res) * P.Cout) + __ldg(r + (size_t)(t0 + gs.res + 1) * P.Cout);
      v += 0.25f * s;
    }
  }
  return v;
}

__device__ __forceinline__ void store_out(const ConvParams& P, int b, int tok, int n, float v) {
  if (P.out_chmajor) P.out[((size_t)b * P.Cout + n) * P.geo.L + tok] = v;
  else               P.out[((size_t)b * P.geo.L + tok) * P.Cout + n] = v;
}

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
k_conv_simt(const __grid_constant__ ConvParams P) {
  MTV_PDL_WAIT();
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int BK = 16;
  constexpr int LDA = BM + 4;
  static_assert(BM * 4 <= NT && BK * BN / 4 <= NT, "one float4 per thread per tile");
  __shared__ __align__(16) float As[BK][LDA];
  __shared__ __align__(16) float Bs[BK][BN];
  __shared__ int s_b[BM], s_p[BM], s_y[BM], s_x[BM];

  const int tid = threadIdx.x;
  const Geo g = P.geo;
  const int M = P.B * g.L;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

  for (int r = tid; r < BM; r += NT) {
    const int m = m0 + r;
    int b = -1, p = 0, y = 0, x = 0;
    if (m < M) { b = m / g.L; decode_tok(g, m - b * g.L, p, y, x); }
    s_b[r] = b; s_p[r] = p; s_y[r] = y; s_x[r] = x;
  }
  __syncthreads();

  // flattened K iteration space: (segment, tap, 16-channel chunk)
  const int Ct0 = P.seg[0].C0 + P.seg[0].C1;
  const int it_seg0 = P.seg[0].taps * (Ct0 / BK);
  int it_total = it_seg0;
  if (P.nsegs > 1) it_total += P.seg[1].taps * ((P.seg[1].C0 + P.seg[1].C1) / BK);
  int it_begin = 0, it_end = it_total;
  if (P.ksplit > 1) {
    const int per = (it_total + P.ksplit - 1) / P.ksplit;
    it_begin = blockIdx.z * per;
    it_end = min(it_total, it_begin + per);
  }

  const bool a_thr = tid < BM * 4;
  const int a_r = tid >> 2, a_kq = tid & 3;
  const bool b_thr = tid < BK * BN / 4;
  const int b_k = tid / (BN / 4), b_nq = tid % (BN / 4);
  const int ty = tid / (BN / TN), tx = tid % (BN / TN);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

  // prefetch registers
  float4 araw[4]; float4 na, nd, braw;
  bool a_valid = false; int a_mode = 0;   // bit0: has norm affine, bit1: silu, bits 2-3: resample

  auto prefetch = [&](int it) {
    int s = 0, itl = it;
    if (it >= it_seg0) { s = 1; itl = it - it_seg0; }
    const KSeg& S = P.seg[s];
    const int Ct = S.C0 + S.C1;
    const int kch = Ct / BK;
    const int tap = itl / kch, kc = itl - tap * kch;
    a_valid = false;
    if (a_thr) {
      const int b = s_b[a_r];
      if (b >= 0) {
        const int p = s_p[a_r];
        int yy = s_y[a_r], xx = s_x[a_r];
        if (S.taps == 9) { const int dy = tap / 3; yy += dy - 1; xx += (tap - dy * 3) - 1; }
        if (yy >= 0 && yy < plane_h(g, p) && xx >= 0 && xx < g.res) {
          a_valid = true;
          const int c = kc * BK + a_kq * 4;
          const float* src; int C, cc;
          if (c < S.C0) { src = S.src0; C = S.C0; cc = c; } else { src = S.src1; C = S.C1; cc = c - S.C0; }
          a_mode = (S.nrm_a ? 1 : 0) | (S.silu ? 2 : 0) | (S.resample << 2);
          if (S.resample == RS_NONE) {
            const int ts = plane_off(g, p) + yy * g.res + xx;
            araw[0] = ldg4(src + ((size_t)b * g.L + ts) * C + cc);
          } else if (S.resample == RS_UP2) {
            const Geo gs = geo_down(g);
            const int ts = plane_off(gs, p) + (yy >> 1) * gs.res + (xx >> 1);
            araw[0] = ldg4(src + ((size_t)b * gs.L + ts) * C + cc);
          } else {
            const Geo gs = geo_up(g);
            const int ts = plane_off(gs, p) + (2 * yy) * gs.res + 2 * xx;
            const float* q = src + ((size_t)b * gs.L + ts) * C + cc;
            araw[0] = ldg4(q);
            araw[1] = ldg4(q + C);
            araw[2] = ldg4(q + (size_t)gs.res * C);
            araw[3] = ldg4(q + (size_t)(gs.res + 1) * C);
          }
          if (S.nrm_a) {
            const size_t idx = ((size_t)b * S.nrm_nseg + (S.nrm_nseg == 3 ? p : 0)) * Ct + c;
            na = ldg4(S.nrm_a + idx);
            nd = ldg4(S.nrm_d + idx);
          }
        }
      }
    }
    if (b_thr) {
      const int k = kc * BK + b_k, n = n0 + b_nq * 4;
      braw = (n < P.Cout) ? ldg4(S.w + ((size_t)tap * Ct + k) * P.Cout + n) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };

  auto xform = [&](float4 v) -> float4 {
    if (a_mode & 1) {
      v.x = fmaf(v.x, na.x, nd.x); v.y = fmaf(v.y, na.y, nd.y);
      v.z = fmaf(v.z, na.z, nd.z); v.w = fmaf(v.w, na.w, nd.w);
    }
    if (a_mode & 2) { v.x = silu_f(v.x); v.y = silu_f(v.y); v.z = silu_f(v.z); v.w = silu_f(v.w); }
    return v;
  };

  auto stage = [&]() {
    if (a_thr) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a_valid) {
        if ((a_mode >> 2) == RS_DOWN2) {   // transform first, then average (h_upd after SiLU, unet.py:181-182)
          const float4 v0 = xform(araw[0]), v1 = xform(araw[1]), v2 = xform(araw[2]), v3 = xform(araw[3]);
          v.x = 0.25f * ((v0.x + v1.x) + (v2.x + v3.x)); v.y = 0.25f * ((v0.y + v1.y) + (v2.y + v3.y));
          v.z = 0.25f * ((v0.z + v1.z) + (v2.z + v3.z)); v.w = 0.25f * ((v0.w + v1.w) + (v2.w + v3.w));
        } else {
          v = xform(araw[0]);
        }
      }
      As[a_kq * 4 + 0][a_r] = v.x; As[a_kq * 4 + 1][a_r] = v.y;
      As[a_kq * 4 + 2][a_r] = v.z; As[a_kq * 4 + 3][a_r] = v.w;
    }
    if (b_thr) *reinterpret_cast<float4*>(&Bs[b_k][b_nq * 4]) = braw;
  };

  if (it_begin < it_end) prefetch(it_begin);
  for (int it = it_begin; it < it_end; ++it) {
    stage();
    __syncthreads();
    if (it + 1 < it_end) prefetch(it + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], bb[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[k][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) bb[j] = Bs[k][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }

  MTV_PDL_TRIGGER();
  // epilogue
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int r = ty * TM + i;
    const int b = s_b[r];
    if (b < 0) continue;
    const int m = m0 + r;
    const int tok = m - b * g.L;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= P.Cout) continue;
      if (P.ksplit > 1) {
        P.partial[((size_t)blockIdx.z * M + m) * P.Cout + n] = acc[i][j];
      } else {
        store_out(P, b, tok, n, acc[i][j] + epilogue_add(P, b, s_p[r], s_y[r], s_x[r], tok, n));
      }
    }
  }
}

// Deterministic split-K reduction + epilogue (fixed summation order over splits).
__global__ void k_splitk_epilogue(const __grid_constant__ ConvParams P) {
  MTV_PDL_TRIGGER();
  MTV_PDL_WAIT();
  const Geo g = P.geo;
  const size_t M = (size_t)P.B * g.L;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * P.Cout) return;
  const int m = (int)(idx / P.Cout), n = (int)(idx - (size_t)m * P.Cout);
  float s = 0.0f;
  for (int z = 0; z < P.ksplit; ++z) s += P.partial[((size_t)z * M + m) * P.Cout + n];
  const int b = m / g.L, tok = m - b * g.L;
  int p, y, x; decode_tok(g, tok, p, y, x);
  store_out(P, b, tok, n, s + epilogue_add(P, b, p, y, x, tok, n));
}

static inline int conv_iters(const ConvParams& P) {
  int it = 0;
  for (int s = 0; s < P.nsegs; ++s) it += P.seg[s].taps * ((P.seg[s].C0 + P.seg[s].C1) / 16);
  return it;
}
static inline int conv_bm(const ConvParams& P, int num_sms) {
  if (P.Cout <= 16) return 64;
  const int M = P.B * P.geo.L;
  const int nt = (P.Cout + 63) / 64;
  if (((M + 63) / 64) * nt >= num_sms) return 64;
  return 32;
}
int conv_simt_pick_ksplit(const ConvParams& P, int num_sms) {
  const int M = P.B * P.geo.L;
  const int bm = conv_bm(P, num_sms);
  const int bn = P.Cout <= 16 ? 16 : 64;
  const int base = ((M + bm - 1) / bm) * ((P.Cout + bn - 1) / bn);
  if (base >= num_sms) return 1;
  const int iters = conv_iters(P);
  int ks = (4 * num_sms + base - 1) / base;
  ks = ks > 32 ? 32 : ks;
  const int max_by_iters = iters / 4 > 0 ? iters / 4 : 1;
  ks = ks > max_by_iters ? max_by_iters : ks;
  if (ks < 2) return 1;
  // no empty splits: per = ceil(iters/ks); need (ks-1)*per < iters
  while (ks > 1 && (ks - 1) * ((iters + ks - 1) / ks) >= iters) --ks;
  return ks;
}

cudaError_t launch_conv_simt(const ConvParams& P, cudaStream_t s) {
  const int M = P.B * P.geo.L;
  static int sms = 0;
  if (!sms) {
    int dev = 0; cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  const int ks = P.ksplit > 1 ? P.ksplit : 1;
  if (P.Cout <= 16) {
    dim3 grid((M + 63) / 64, (P.Cout + 15) / 16, ks);
    { cudaError_t le_ = launch_k(k_conv_simt<64, 16, 4, 1>, dim3(grid), dim3(256), (size_t)(0), s, P); if (le_ != cudaSuccess) return le_; }
  } else if (conv_bm(P, sms) == 64) {
    dim3 grid((M + 63) / 64, (P.Cout + 63) / 64, ks);
    { cudaError_t le_ = launch_k(k_conv_simt<64, 64, 4, 4>, dim3(grid), dim3(256), (size_t)(0), s, P); if (le_ != cudaSuccess) return le_; }
  } else {
    dim3 grid((M + 31) / 32, (P.Cout + 63) / 64, ks);
    { cudaError_t le_ = launch_k(k_conv_simt<32, 64, 2, 4>, dim3(grid), dim3(256), (size_t)(0), s, P); if (le_ != cudaSuccess) return le_; }
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (ks > 1) {
    const size_t tot = (size_t)M * P.Cout;
    { cudaError_t le_ = launch_k(k_splitk_epilogue, dim3((unsigned)((tot + 255) / 256)), dim3(256), (size_t)(0), s, P); if (le_ != cudaSuccess) return le_; }
    e = cudaGetLastError();
  }
  return e;
}

// ------------------------------------------------------------------ GroupNorm statistics
// GroupNorm32 (diffusionmodules.py:171-173): 32 groups, eps 1e-5, statistics over
// (channels of the group) x (tokens of the segment).  Produces the per-channel affine
// y = x*a + d that consumers apply on load, with the ResBlock FiLM
// h*(1+scale)+shift (unet.py:201-202) folded in.  Sums are fp32 per thread over <= chunk
// tokens, fp64 across threads / CTAs (atomics), finalised by the last CTA of a segment.
__global__ void __launch_bounds__(256) k_gn_stats(const __grid_constant__ GnParams P) {
  MTV_PDL_TRIGGER();
  MTV_PDL_WAIT();
  __shared__ double s_sum[32][2];
  __shared__ bool s_last;
  const int tid = threadIdx.x;
  const int bs = blockIdx.y;                 // b * nseg + seg
  const int b = bs / P.nseg, sg = bs - b * P.nseg;
  const int t_lo = P.seg_off[sg], t_hi = P.seg_off[sg + 1];
  const int nchunks = (t_hi - t_lo + P.chunk_tokens - 1) / P.chunk_tokens;
  if ((int)blockIdx.x >= nchunks) return;
  const int C = P.C0 + P.C1, cpg = C / 32;
  if (tid < 64) s_sum[tid >> 1][tid & 1] = 0.0;
  __syncthreads();
  const int c_lo = t_lo + blockIdx.x * P.chunk_tokens;
  const int c_hi = min(t_hi, c_lo + P.chunk_tokens);
  for (int c = tid; c < C; c += 256) {
    const float* src; int Cs, cc;
    if (c < P.C0) { src = P.src0; Cs = P.C0; cc = c; } else { src = P.src1; Cs = P.C1; cc = c - P.C0; }
    const float* q = src + ((size_t)b * P.L) * Cs + cc;
    float s = 0.f, ss = 0.f;
    for (int tk = c_lo; tk < c_hi; ++tk) { const float v = __ldg(q + (size_t)tk * Cs); s += v; ss = fmaf(v, v, ss); }
    const int gidx = c / cpg;
    atomicAdd(&s_sum[gidx][0], (double)s);
    atomicAdd(&s_sum[gidx][1], (double)ss);
  }
  __syncthreads();
  double* gs = P.sums + (size_t)bs * 64;
  if (tid < 64) atomicAdd(gs + tid, s_sum[tid >> 1][tid & 1]);
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned prev = atomicAdd(P.counter + bs, 1u);
    s_last = (prev == (unsigned)nchunks - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (tid < 64) s_sum[tid >> 1][tid & 1] = __ldcg(gs + tid);
  __syncthreads();
  const double cnt = (double)cpg * (double)(t_hi - t_lo);
  for (int c = tid; c < C; c += 256) {
    const int gidx = c / cpg;
    const double mean = s_sum[gidx][0] / cnt;
    double var = s_sum[gidx][1] / cnt - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    const double rstd = rsqrt(var + 1e-5);
    double a = rstd * (double)__ldg(P.gamma + c);
    double d = (double)__ldg(P.beta + c) - mean * a;
    if (P.film) {
      const float* f = P.film + (size_t)b * P.film_stride;
      const double sc = 1.0 + (double)__ldg(f + c);
      a *= sc; d = d * sc + (double)__ldg(f + C + c);
    }
    P.nrm_a[(size_t)bs * C + c] = (float)a;
    P.nrm_d[(size_t)bs * C + c] = (float)d;
  }
  __syncthreads();
  if (tid < 64) gs[tid] = 0.0;          // leave the scratch zeroed for the next forward
  if (tid == 0) P.counter[bs] = 0u;
}

cudaError_t launch_gn_stats(const GnParams& P, cudaStream_t s) {
  int maxlen = 0;
  for (int i = 0; i < P.nseg; ++i) maxlen = max(maxlen, P.seg_off[i + 1] - P.seg_off[i]);
  dim3 grid((maxlen + P.chunk_tokens - 1) / P.chunk_tokens, P.B * P.nseg);
  { cudaError_t le_ = launch_k(k_gn_stats, dim3(grid), dim3(256), (size_t)(0), s, P); if (le_ != cudaSuccess) return le_; }
  return cudaGetLastError();
}

// ------------------------------------------------------------------ attention (fp32, flash-style)
// QKVAttentionLegacy.forward (unet.py:312-326): per (sample, head) softmax(q k^T / sqrt(D)) v
// over the tokens of one segment (one plane for AttentionBlock, all planes for
// AttentionBlock1D).  q and k each carry D^-1/4 in the reference; the product of the
// two scales is applied to the score here.  The L x L score matrix is never
// materialised: 64-query x 64-key tiles with an online softmax.
template <int D>
__global__ void __launch_bounds__(256) k_attn_simt(const __grid_constant__ AttnParams P) {
  MTV_PDL_TRIGGER();
  MTV_PDL_WAIT();
  constexpr int BQ = 64, BKV = 64, LD = 68;
  constexpr int CD = D / 16;                      // output columns per thread
  extern __shared__ __align__(16) float smem[];
  float* Qs = smem;                               // [D][LD]  (d, q)
  float* Ks = Qs + D * LD;                        // [D][LD]  (d, j)
  float* Vs = Ks + D * LD;                        // [BKV][D] (j, d)
  float* Ps = Vs + BKV * D;                       // [BKV][LD] (j, q)

  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  // blockIdx.x enumerates (segment, query block); blockIdx.y = b*heads + head
  int sg = 0, qb = blockIdx.x;
  for (; sg < P.nseg; ++sg) {
    const int nb = (P.seg_off[sg + 1] - P.seg_off[sg] + BQ - 1) / BQ;
    if (qb < nb) break;
    qb -= nb;
  }
  const int b = blockIdx.y / P.heads, hd = blockIdx.y - b * P.heads;
  const int t_lo = P.seg_off[sg], len = P.seg_off[sg + 1] - t_lo;
  const int q0 = qb * BQ;
  const int C3 = 3 * P.C;
  const float* base = P.qkv + ((size_t)b * P.L + t_lo) * C3 + hd * 3 * D;
  const float scale = rsqrtf((float)D);

  // Q tile (transposed)
  for (int i = tid; i < BQ * (D / 4); i += 256) {
    const int r = i / (D / 4), dq = i - r * (D / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + r < len) v = ldg4(base + (size_t)(q0 + r) * C3 + dq * 4);
    Qs[(dq * 4 + 0) * LD + r] = v.x; Qs[(dq * 4 + 1) * LD + r] = v.y;
    Qs[(dq * 4 + 2) * LD + r] = v.z; Qs[(dq * 4 + 3) * LD + r] = v.w;
  }

  float m_i[4], l_i[4], o[4][CD];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m_i[i] = -INFINITY; l_i[i] = 0.f;
#pragma unroll
    for (int e = 0; e < CD; ++e) o[i][e] = 0.f;
  }

  for (int j0 = 0; j0 < len; j0 += BKV) {
    __syncthreads();   // previous tile's Ks/Vs/Ps fully consumed (also orders the Q stores)
    for (int i = tid; i < BKV * (D / 4); i += 256) {
      const int r = i / (D / 4), dq = i - r * (D / 4);
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (j0 + r < len) {
        const float* row = base + (size_t)(j0 + r) * C3;
        kv = ldg4(row + D + dq * 4);
        vv = ldg4(row + 2 * D + dq * 4);
      }
      Ks[(dq * 4 + 0) * LD + r] = kv.x; Ks[(dq * 4 + 1) * LD + r] = kv.y;
      Ks[(dq * 4 + 2) * LD + r] = kv.z; Ks[(dq * 4 + 3) * LD + r] = kv.w;
      *reinterpret_cast<float4*>(&Vs[r * D + dq * 4]) = vv;
    }
    __syncthreads();

    float sc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) sc[i][j] = 0.f;
#pragma unroll 8
    for (int d = 0; d < D; ++d) {
      const float4 a = *reinterpret_cast<const float4*>(&Qs[d * LD + ty * 4]);
      const float4 k = *reinterpret_cast<const float4*>(&Ks[d * LD + tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, kv[4] = {k.x, k.y, k.z, k.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) sc[i][j] = fmaf(av[i], kv[j], sc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        sc[i][j] = (j0 + tx * 4 + j < len) ? sc[i][j] * scale : -INFINITY;
        mx = fmaxf(mx, sc[i][j]);
      }
#pragma unroll
      for (int off = 1; off < 16; off <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      const float m_new = fmaxf(m_i[i], mx);       // finite: every key tile holds >= 1 valid key
      const float corr = __expf(m_i[i] - m_new);   // exp(-inf) = 0 on the first tile
      float rs = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) { sc[i][j] = __expf(sc[i][j] - m_new); rs += sc[i][j]; }
#pragma unroll
      for (int off = 1; off < 16; off <<= 1) rs += __shfl_xor_sync(0xffffffffu, rs, off);
      l_i[i] = l_i[i] * corr + rs;
      m_i[i] = m_new;
#pragma unroll
      for (int e = 0; e < CD; ++e) o[i][e] *= corr;
#pragma unroll
      for (int j = 0; j < 4; ++j) Ps[(tx * 4 + j) * LD + ty * 4 + i] = sc[i][j];
    }
    __syncthreads();
#pragma unroll 8
    for (int j = 0; j < BKV; ++j) {
      const float4 p4 = *reinterpret_cast<const float4*>(&Ps[j * LD + ty * 4]);
      const float pv[4] = {p4.x, p4.y, p4.z, p4.w};
      float vv[CD];
#pragma unroll
      for (int e = 0; e < CD; ++e) vv[e] = Vs[j * D + tx * CD + e];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int e = 0; e < CD; ++e) o[i][e] = fmaf(pv[i], vv[e], o[i][e]);
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = q0 + ty * 4 + i;
    if (q >= len) continue;
    const float inv = 1.0f / l_i[i];
    float* dst = P.out + ((size_t)b * P.L + t_lo + q) * P.C + hd * D + tx * CD;
#pragma unroll
    for (int e = 0; e < CD; ++e) dst[e] = o[i][e] * inv;
  }
}

template <int D>
static cudaError_t launch_attn_d(const AttnParams& P, cudaStream_t s) {
  constexpr int LD = 68;
  const size_t smem = (size_t)(2 * D * LD + 64 * D + 64 * LD) * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k_attn_simt<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  int nqb = 0;
  for (int i = 0; i < P.nseg; ++i) nqb += (P.seg_off[i + 1] - P.seg_off[i] + 63) / 64;
  dim3 grid(nqb, P.B * P.heads);
  { cudaError_t le_ = launch_k(k_attn_simt<D>, dim3(grid), dim3(256), (size_t)(smem), s, P); if (le_ != cudaSuccess) return le_; }
  return cudaGetLastError();
}

cudaError_t launch_attn_simt(const AttnParams& P, cudaStream_t s) {
  const int D = P.C / P.heads;
  switch (D) {
    case 16: return launch_attn_d<16>(P, s);
    case 32: return launch_attn_d<32>(P, s);
    case 64: return launch_attn_d<64>(P, s);
    case 128: return launch_attn_d<128>(P, s);
    default: return cudaErrorInvalidValue;
  }
}

// ------------------------------------------------------------------ timestep embedding + FiLM table
// timestep_embedding (diffusionmodules.py:108-128) -> time_embed Linear/SiLU/Linear
// (unet.py:701-705, 1011-1012) -> every ResBlock's emb_layers = Linear(SiLU(emb))
// (unet.py:148-154, 193) in one batched GEMV.  One warp per output feature.
__global__ void k_temb(const __grid_constant__ EmbParams P) {
  MTV_PDL_TRIGGER();
  MTV_PDL_WAIT();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = P.mc / 2;
  if (i >= P.B * half) return;
  const int b = i / half, k = i - b * half;
  const float arg = (float)P.t[b] * P.freqs[k];
  P.temb[(size_t)b * P.mc + k] = cosf(arg);
  P.temb[(size_t)b * P.mc + half + k] = sinf(arg);
}

// out[b][j] = act(bias[j] + W[j][:] . in[b][:]);  K % 128 == 0
__global__ void __launch_bounds__(256) k_linear_warp(const float* __restrict__ W, const float* __restrict__ bias,
                                                     const float* __restrict__ in, float* __restrict__ out,
                                                     int J, int K, int B, int silu_out) {
  MTV_PDL_TRIGGER();
  MTV_PDL_WAIT();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= J) return;
  const float* w = W + (size_t)warp * K;
  for (int b = 0; b < B; ++b) {
    const float* x = in + (size_t)b * K;
    float s = 0.f;
    for (int k = lane * 4; k < K; k += 128) {
      const float4 wv = ldg4(w + k), xv = ldg4(x + k);
      s = fmaf(wv.x, xv.x, s); s = fmaf(wv.y, xv.y, s); s = fmaf(wv.z, xv.z, s); s = fmaf(wv.w, xv.w, s);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) {
      s += __ldg(bias + warp);
      out[(size_t)b * J + warp] = silu_out ? silu_f(s) : s;
    }
  }
}

cudaError_t launch_emb(const EmbParams& P, cudaStream_t s) {
  const int n = P.B * (P.mc / 2);
  { cudaError_t le_ = launch_k(k_temb, dim3((n + 127) / 128), dim3(128), (size_t)(0), s, P); if (le_ != cudaSuccess) return le_; }
  { cudaError_t le_ = launch_k(k_linear_warp, dim3((P.ted * 32 + 255) / 256), dim3(256), (size_t)(0), s, P.w1, P.b1, P.temb, P.h1, P.ted, P.mc, P.B, 1); if (le_ != cudaSuccess) return le_; }
  // emb is only ever consumed through SiLU (every emb_layers starts with nn.SiLU) -> store silu(emb)
  { cudaError_t le_ = launch_k(k_linear_warp, dim3((P.ted * 32 + 255) / 256), dim3(256), (size_t)(0), s, P.w2, P.b2, P.h1, P.semb, P.ted, P.ted, P.B, 1); if (le_ != cudaSuccess) return le_; }
  { cudaError_t le_ = launch_k(k_linear_warp, dim3((unsigned)(((size_t)P.J * 32 + 255) / 256)), dim3(256), (size_t)(0), s, P.wall, P.ball, P.semb, P.film, P.J, P.ted, P.B, 0); if (le_ != cudaSuccess) return le_; }
  return cudaGetLastError();
}

// ------------------------------------------------------------------ layout helpers
// h = cat([x, cond, cat([image_cond[:, :, :1024], 0])], dim=1)  (unet.py:1022-1025),
// written token-major [B][2048][16].
__global__ void k_pack_in(const __grid_constant__ PackParams P) {
  const int Ct = P.cx + P.cc + P.ci;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)P.B * 2048 * Ct) return;
  const int c = (int)(i % Ct);
  const int tok = (int)((i / Ct) % 2048);
  const int b = (int)(i / ((size_t)Ct * 2048));
  float v;
  if (c < P.cx) v = P.x[((size_t)b * P.cx + c) * 2048 + tok];
  else if (c < P.cx + P.cc) v = P.cond[((size_t)b * P.cc + (c - P.cx)) * 2048 + tok];
  else v = (tok < 1024) ? P.image_cond[((size_t)b * P.ci + (c - P.cx - P.cc)) * P.ic_len + tok] : 0.0f;
  P.out[i] = v;
}
// same gather, written as the split-bf16 A operand of the tensor-core stem conv, channels zero-padded to cpad (a K chunk of 64)
__global__ void k_pack_in_split(const __grid_constant__ PackParams P) {
  const int Ct = P.cx + P.cc + P.ci;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)P.B * 2048 * P.cpad) return;
  const int c = (int)(i % P.cpad);
  const int tok = (int)((i / P.cpad) % 2048);
  const int b = (int)(i / ((size_t)P.cpad * 2048));
  float v = 0.0f;
  if (c < P.cx) v = P.x[((size_t)b * P.cx + c) * 2048 + tok];
  else if (c < P.cx + P.cc) v = P.cond[((size_t)b * P.cc + (c - P.cx)) * 2048 + tok];
  else if (c < Ct) v = (tok < 1024) ? P.image_cond[((size_t)b * P.ci + (c - P.cx - P.cc)) * P.ic_len + tok] : 0.0f;
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  reinterpret_cast<__nv_bfloat16*>(P.hi)[i] = h;
  reinterpret_cast<__nv_bfloat16*>(P.lo)[i] = __float2bfloat16_rn(v - __bfloat162float(h));
}
cudaError_t launch_pack_in(const PackParams& P, cudaStream_t s) {
  if (P.hi) {
    const size_t n = (size_t)P.B * 2048 * P.cpad;
    k_pack_in_split<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(P);
    return cudaGetLastError();
  }
  const size_t n = (size_t)P.B * 2048 * (P.cx + P.cc + P.ci);
  k_pack_in<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(P);
  return cudaGetLastError();
}

// PyTorch conv weight [Cout][Cin][taps] -> [taps][Cin][Cout]
__global__ void k_repack_conv(const float* __restrict__ src, float* __restrict__ dst, int Cout, int Cin, int taps) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)Cout * Cin * taps) return;
  const int co = (int)(i % Cout);
  const int ci = (int)((i / Cout) % Cin);
  const int tp = (int)(i / ((size_t)Cout * Cin));
  dst[i] = src[((size_t)co * Cin + ci) * taps + tp];
}
cudaError_t launch_repack_conv(const float* src, float* dst, int Cout, int Cin, int taps, cudaStream_t s) {
  const size_t n = (size_t)Cout * Cin * taps;
  k_repack_conv<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src, dst, Cout, Cin, taps);
  return cudaGetLastError();
}

__global__ void k_add_vec(const float* a, const float* b, float* d, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) d[i] = a[i] + b[i];
}
cudaError_t launch_add_vec(const float* a, const float* b, float* dst, int n, cudaStream_t s) {
  k_add_vec<<<(n + 255) / 256, 256, 0, s>>>(a, b, dst, n);
  return cudaGetLastError();
}

// token-major [B][L][C] -> channel-major [B][C][L]  (debug taps only)
__global__ void k_tok2ch(const float* __restrict__ src, float* __restrict__ dst, int B, int L, int C) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * L * C) return;
  const int l = (int)(i % L);
  const int c = (int)((i / L) % C);
  const int b = (int)(i / ((size_t)L * C));
  dst[i] = src[((size_t)b * L + l) * C + c];
}
cudaError_t launch_tok2ch(const float* src, float* dst, int B, int L, int C, cudaStream_t s) {
  const size_t n = (size_t)B * L * C;
  k_tok2ch<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src, dst, B, L, C);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ sampler elementwise
// DDPM.model_predictions + ddim_sample loop body (ddpm.py:346-351, 386-398).  The
// reference evaluates each product and sum as a separate fp32 op; __fmul_rn/__fadd_rn
// keep nvcc from contracting them into FMAs so the update is bit-identical.
__global__ void k_ddim_step(float* __restrict__ img, const float* __restrict__ eps, const float* __restrict__ noise,
                            int64_t n, float sr, float srm1, float san, float c, float sigma, int last) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float e = eps[i];
  float x0 = __fsub_rn(__fmul_rn(sr, img[i]), __fmul_rn(srm1, e));
  x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
  if (last) { img[i] = x0; return; }
  img[i] = __fadd_rn(__fadd_rn(__fmul_rn(x0, san), __fmul_rn(c, e)), __fmul_rn(sigma, noise[i]));
}
cudaError_t launch_ddim_step(float* img, const float* eps, const float* noise, int64_t n,
                             float sr, float srm1, float san, float c, float sigma, int last, cudaStream_t s) {
  k_ddim_step<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(img, eps, noise, n, sr, srm1, san, c, sigma, last);
  return cudaGetLastError();
}

// DDPM.q_sample (ddpm.py:486-491)
__global__ void k_q_sample(const float* __restrict__ x0, const float* __restrict__ noise, int64_t n, float a, float b,
                           float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __fadd_rn(__fmul_rn(a, x0[i]), __fmul_rn(b, noise[i]));
}
cudaError_t launch_q_sample(const float* x0, const float* noise, int64_t n, float a, float b, float* out, cudaStream_t s) {
  k_q_sample<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(x0, noise, n, a, b, out);
  return cudaGetLastError();
}

}  // namespace mtv
