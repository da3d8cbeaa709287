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
  int B; Geo geo;                                      // OUTPUT geometry
  void* hi; void* lo;                                  // __nv_bfloat16 [B][L][C0+C1]
};

// D[B*L][Cout] = sum_tap A_tap[B*L][Cin] * W[tap][Cout][Cin]^T   (+bias, +residual)
struct TcConvParams {
  CUtensorMap tmA_hi[2], tmA_lo[2];   // taps==9: [0] xy plane (C,W,H,B), [1] yt|xt planes (C,W,H,2,B); taps==1: [0] = (C, B*L)
  CUtensorMap tmW_hi, tmW_lo;         // (Cin, taps*Cout), box (64, BN)
  // optional second K-segment: a 1x1 conv of another activation accumulated into the same tile
  // (the ResBlock skip_connection, unet.py:167,207); Cin2 == 0 when absent
  CUtensorMap tmA2_hi, tmA2_lo;       // (Cin2, B*L)
  CUtensorMap tmW2_hi, tmW2_lo;       // (Cin2, Cout)
  int Cin2; int bn;
  int taps, Cin, Cout, B; Geo geo;
  const float* bias; const float* resid; int resid_mode;
  float* out; float* partial; int ksplit;
};

cudaError_t launch_apply_split(const ApplyParams& P, cudaStream_t s);
cudaError_t launch_repack_split_w(const float* src, void* hi, void* lo, int Cout, int Cin, int taps, cudaStream_t s);
cudaError_t launch_conv_tc(const TcConvParams& P, cudaStream_t s);

}  // namespace mtv
