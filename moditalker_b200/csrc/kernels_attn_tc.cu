// kernels_attn_tc.cu — FlashAttention-style QKVAttentionLegacy on tcgen05 (sm_100a).
//
// Reference: QKVAttentionLegacy.forward (MToV/models/ddpm/unet.py:312-326) as used by
// AttentionBlock (per plane, unet.py:248-254) and AttentionBlock1D (cross-plane,
// unet.py:295-300): softmax((q s)^T (k s)) v per (sample, head), s = D^-1/4, fp32 softmax.
//
//   k_qkv_split : qkv fp32 [B][L][3C] (head-major q|k|v channel order, unet.py:321) ->
//                 split-bf16 Q (pre-scaled by log2(e)/sqrt(D)), K as [B*H][L][D] and V^T as
//                 [B*H][D][L], so every MMA operand is K-major and TMA-loadable.
//   k_attn_tc<D>: one CTA = 128 queries of one (sample, head, segment).  Per 64-key block:
//                 S = Q K^T            tcgen05.mma, 128x64 fp32 in TMEM (3 split-bf16 MMAs / k-step)
//                 online softmax       8 warps, two threads per query row (32 of the 64 key columns and half
//                                      of the D output columns each; row max exchanged through smem):
//                                      tcgen05.ld S, exp2, P -> smem as split bf16 in the UMMA
//                                      128B-swizzled K-major layout
//                 O_blk = P V          tcgen05.mma, 128xD fp32 in TMEM, folded into registers
//                 The L x L score matrix never exists (reference: 134 MB per top-level call).
//                 S(j+1) is issued while the softmax of block j runs.
#include "mtv_kernels.cuh"
#include "mtv_tc.cuh"

#include <cuda_bf16.h>
#include <cstdio>

namespace mtv {

namespace {

__device__ __forceinline__ uint32_t a_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void a_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void a_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void a_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a_smem_u32(bar)) : "memory");
}
// The softmax warps wait with the bare retry loop: ANY extra instruction in it (a failure counter, a back-off) measured 6-7 % of
// the whole kernel on B200 (ten inlined wait sites per key block; profiles/r02_attention.md).  The hang guard lives in the two
// single-purpose warps instead (a_mbar_wait_guard): every deadlock of this kernel also blocks the TMA and the MMA warp, whose
// bounded waits then trap — a pipeline bug still surfaces as a launch failure, never as a hung GPU.
__device__ __forceinline__ void a_mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = a_smem_u32(bar);
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void a_mbar_wait_guard(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = a_smem_u32(bar);
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
__device__ __forceinline__ void a_tma_2d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void a_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void a_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major operand descriptor; swizzle span = bytes per row (32 / 64 / 128); 8-row groups are
// 8*row_bytes apart (SmemDescriptor fields as in kernels_tc.cu; layout codes: 128B=2, 64B=4, 32B=6).
__device__ __forceinline__ uint64_t a_desc(uint32_t smem_addr, int row_bytes) {
  const uint64_t layout = row_bytes == 128 ? 2 : (row_bytes == 64 ? 4 : 6);
  uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((8 * row_bytes) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= layout << 61;
  return d;
}
__host__ __device__ constexpr uint32_t a_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void a_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// one lane of a converged warp: the TMA / MMA warps run their loops in all lanes (uniform control flow, descriptors in
// uniform registers) and elect only around the issuing instructions — inside an `if (lane == 0)` region ptxas wraps every
// tcgen05.mma in an ELECT / R2UR.BROADCAST waterfall (~115 cycles per MMA; see kernels_tc.cu: elect_one)
__device__ __forceinline__ bool a_elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// A operand from tensor memory (lane = row, each 32-bit column = two consecutive K elements; K = 16 -> 8 columns per MMA)
__device__ __forceinline__ void a_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void a_tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void a_tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void a_tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void a_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(a_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void a_tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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
__device__ __forceinline__ void a_tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr));
}
__device__ __forceinline__ void a_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// two floats -> packed bf16x2 (e0 in the low half = lower address), one instruction
__device__ __forceinline__ uint32_t cvt_bf16x2(float e0, float e1) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(e1), "f"(e0));
  return d;
}
// packed fp32 pairs (sm_100 FADD2 / FMUL2): the softmax warps are instruction-issue bound, these halve the add / mul count
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void f2_unpack(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t f2_sub(uint64_t a, uint64_t b) { uint64_t r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float max3(float a, float b, float c) { float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ void a_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {   // 2^x, max rel. error 2^-22; 2^-inf = 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

}  // namespace

// ------------------------------------------------------------------ qkv -> split operands
__global__ void __launch_bounds__(128) k_qkv_split(const __grid_constant__ QkvSplitParams P) {
  MTV_PDL_TRIGGER();
  MTV_PDL_WAIT();
  const int D = P.C / P.heads;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;     // (b, head, token), token fastest
  const size_t total = (size_t)P.B * P.heads * P.L;
  if (idx >= total) return;
  const int tok = (int)(idx % P.L);
  const int bh = (int)(idx / P.L);
  const int b = bh / P.heads, h = bh - b * P.heads;
  const float* row = P.qkv + ((size_t)b * P.L + tok) * (3 * P.C) + (size_t)h * 3 * D;
  const float qs = 1.4426950408889634f * rsqrtf((float)D);    // both D^-1/4 factors and log2(e)
  __nv_bfloat16* Qh = reinterpret_cast<__nv_bfloat16*>(P.q_hi) + ((size_t)bh * P.L + tok) * D;
  __nv_bfloat16* Ql = reinterpret_cast<__nv_bfloat16*>(P.q_lo) + ((size_t)bh * P.L + tok) * D;
  __nv_bfloat16* Kh = reinterpret_cast<__nv_bfloat16*>(P.k_hi) + ((size_t)bh * P.L + tok) * D;
  __nv_bfloat16* Kl = reinterpret_cast<__nv_bfloat16*>(P.k_lo) + ((size_t)bh * P.L + tok) * D;
  __nv_bfloat16* Vh = reinterpret_cast<__nv_bfloat16*>(P.vt_hi) + (size_t)bh * D * P.L + tok;
  __nv_bfloat16* Vl = reinterpret_cast<__nv_bfloat16*>(P.vt_lo) + (size_t)bh * D * P.L + tok;
  for (int d = 0; d < D; d += 4) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(row + d));
    const float4 k = __ldg(reinterpret_cast<const float4*>(row + D + d));
    const float4 v = __ldg(reinterpret_cast<const float4*>(row + 2 * D + d));
    const float qv[4] = {q.x * qs, q.y * qs, q.z * qs, q.w * qs}, kv[4] = {k.x, k.y, k.z, k.w}, vv[4] = {v.x, v.y, v.z, v.w};
    __nv_bfloat16 qh[4], ql[4], kh[4], kl[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      qh[i] = __float2bfloat16_rn(qv[i]); ql[i] = __float2bfloat16_rn(qv[i] - __bfloat162float(qh[i]));
      kh[i] = __float2bfloat16_rn(kv[i]); kl[i] = __float2bfloat16_rn(kv[i] - __bfloat162float(kh[i]));
      const __nv_bfloat16 vh = __float2bfloat16_rn(vv[i]);
      Vh[(size_t)(d + i) * P.L] = vh;
      Vl[(size_t)(d + i) * P.L] = __float2bfloat16_rn(vv[i] - __bfloat162float(vh));
    }
    *reinterpret_cast<uint2*>(Qh + d) = *reinterpret_cast<const uint2*>(qh);
    *reinterpret_cast<uint2*>(Ql + d) = *reinterpret_cast<const uint2*>(ql);
    *reinterpret_cast<uint2*>(Kh + d) = *reinterpret_cast<const uint2*>(kh);
    *reinterpret_cast<uint2*>(Kl + d) = *reinterpret_cast<const uint2*>(kl);
  }
}
cudaError_t launch_qkv_split(const QkvSplitParams& P, cudaStream_t s) {
  const size_t total = (size_t)P.B * P.heads * P.L;
  { cudaError_t le_ = launch_k(k_qkv_split, dim3((unsigned)((total + 127) / 128)), dim3(128), (size_t)(0), s, P); if (le_ != cudaSuccess) return le_; }
  return cudaGetLastError();
}

// ------------------------------------------------------------------ attention
constexpr int AT_BQ = 128, AT_BKV = 64, AT_THREADS = 320, AT_NS = 3;   // warps: 0 TMA, 1 MMA, 2-5 / 6-9 softmax halves
// P (softmax probabilities, split bf16) is handed to the PV MMA through TENSOR MEMORY (tcgen05.st by the softmax threads, A operand
// of tcgen05.mma read from TMEM): through shared memory the kernel was smem-bandwidth bound at head dim 16 — per 64-key block 32 KB
// of P stores and 48 KB of MMA re-reads of P against 8 KB of K / V (profiles/r01_s2_attention_skip.md)

template <int D>
struct AttnSmem {
  static constexpr int ROWB = 2 * D;                       // bytes per Q/K row (swizzle span)
  static constexpr int Q_BYTES = AT_BQ * ROWB;             // one of hi/lo
  static constexpr int K_BYTES = AT_BKV * ROWB;
  static constexpr int V_BYTES = D * 128;                  // V^T tile: D rows x 64 keys bf16
  static constexpr int align_up(int v) { return (v + 1023) & ~1023; }
  static constexpr int Q_SLOT = align_up(Q_BYTES), K_SLOT = align_up(K_BYTES), V_SLOT = align_up(V_BYTES);
  static constexpr int STAGE = 2 * K_SLOT + 2 * V_SLOT;
  static constexpr int TOTAL = 2 * Q_SLOT + AT_NS * STAGE + 1024;
  static constexpr int STAGE_TX = 2 * K_BYTES + 2 * V_BYTES;
};

template <int D>
__global__ void __launch_bounds__(AT_THREADS, (D <= 32) ? 2 : 1) k_attn_tc(const __grid_constant__ AttnTcParams P) {
  using SM = AttnSmem<D>;
  mtv_prefetch_slice(P.pf0, P.pf1, P.pf_bytes, blockIdx.x + gridDim.x * blockIdx.y, gridDim.x * gridDim.y);
  constexpr uint32_t IDESC_S = a_idesc(AT_BQ, AT_BKV);
  constexpr uint32_t IDESC_O = a_idesc(AT_BQ, D);
  // D <= 32: V_hi and V_lo tiles are adjacent in smem, so ONE MMA with N = 2D computes [P_hi V_hi | P_hi V_lo] into O columns
  // [0,D) | [D,2D) and a second adds P_lo V_hi into [0,D): 8 instead of 12 tcgen05.mma per key block (the kernel is bound by the
  // NUMBER of small MMAs: ~58 tensor-pipe cycles each whatever N); the two column groups are summed once at the end
  constexpr bool STACK_V = D <= 32;
  constexpr uint32_t IDESC_O2 = a_idesc(AT_BQ, 2 * D);
  static_assert(!STACK_V || AttnSmem<D>::V_SLOT == AttnSmem<D>::V_BYTES, "stacked V needs contiguous hi / lo tiles");
  constexpr int TMEM_COLS = 256;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_q, bar_full[AT_NS], bar_empty[AT_NS], bar_s_full, bar_s_free, bar_p_full[2], bar_pv_done[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_xchg[2][2][AT_BQ];   // [block parity][half][row]: row max (and the final row sum) exchange

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem0 = (a_smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ_hi = smem0, sQ_lo = smem0 + SM::Q_SLOT;
  const uint32_t sStage0 = smem0 + 2 * SM::Q_SLOT;

  // which (segment, query tile), (sample, head)
  // Split-KV (kv_split > 1, small batches only): the kv_split CTAs of one query tile are one cluster; rank kr streams key blocks
  // [kr, kr+1) * nblk_seg / kv_split and rank 0 merges the partial (m, l, O) of the others, handed over through distributed
  // shared memory.  At B=1 a launch has only 32-128 query tiles for 148 SMs and one CTA per SM is latency-bound: twice the CTAs
  // on the same SMs run the same key blocks in about half the time.
  const int kvs = P.kv_split > 1 ? P.kv_split : 1;
  const int kr = kvs > 1 ? (int)(blockIdx.x % kvs) : 0;
  int sg = 0, qb = kvs > 1 ? (int)(blockIdx.x / kvs) : (int)blockIdx.x;
  for (; sg < P.nseg; ++sg) {
    const int nb = (P.seg_off[sg + 1] - P.seg_off[sg] + AT_BQ - 1) / AT_BQ;
    if (qb < nb) break;
    qb -= nb;
  }
  const int bh = blockIdx.y;
  const int t_lo = P.seg_off[sg], len = P.seg_off[sg + 1] - t_lo;
  const int q0 = qb * AT_BQ;
  // distributed shared memory of a peer may only be touched once that CTA is known to be executing: every thread arrives at
  // the cluster barrier here and completes the wait (long since satisfied, no stall) right before the hand-over at the end
  if (kvs > 1) asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  const int nblk_seg = (len + AT_BKV - 1) / AT_BKV;
  const int jb0 = (kr * nblk_seg) / kvs;                       // first key block of this rank
  const int nblk = ((kr + 1) * nblk_seg) / kvs - jb0;          // host guarantees >= 1

  if (threadIdx.x == 0) {
    a_mbar_init(&bar_q, 1);
    for (int s = 0; s < AT_NS; ++s) { a_mbar_init(&bar_full[s], 1); a_mbar_init(&bar_empty[s], 1); }
    a_mbar_init(&bar_s_full, 1); a_mbar_init(&bar_pv_done[0], 1); a_mbar_init(&bar_pv_done[1], 1);
    // one arrival per softmax WARP (8), not per thread: 256 same-address mbarrier arrivals per key block serialise in shared memory
    a_mbar_init(&bar_s_free, 8); a_mbar_init(&bar_p_full[0], 8); a_mbar_init(&bar_p_full[1], 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a_smem_u32(&tmem_base_s)), "r"((uint32_t)TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  a_fence_before();
  __syncthreads();
  a_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  // S: cols [0,64), O (accumulated over ALL key blocks): [64, 64+D), P double-buffered: buffer b at 128 + 64 b = P_hi (32 cols of
  // packed bf16 pairs) | P_lo (32)
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 64;
  const uint32_t tmem_P = tmem_base + 128;
  const uint32_t last_pv_par = (uint32_t)(((nblk - 1) >> 1) & 1);      // phase parity of the last PV commit on bar_pv_done[(nblk-1)&1]
  MTV_PDL_WAIT();        // Q / K / V^T are written by the preceding k_qkv_split

  if (warp == 0) {
    // =============================== TMA producer ===============================
    {
      if (a_elect_one()) {
        a_mbar_expect_tx(&bar_q, 2 * SM::Q_BYTES);
        a_tma_2d(sQ_hi, &P.tmQ_hi, a_smem_u32(&bar_q), 0, bh * P.L + t_lo + q0);
        a_tma_2d(sQ_lo, &P.tmQ_lo, a_smem_u32(&bar_q), 0, bh * P.L + t_lo + q0);
      }
      __syncwarp();
      int stage = 0; uint32_t phase = 0;
      for (int j = 0; j < nblk; ++j) {
        a_mbar_wait_guard(&bar_empty[stage], phase ^ 1u);
        if (a_elect_one()) {
          a_mbar_expect_tx(&bar_full[stage], SM::STAGE_TX);
          const uint32_t sK_hi = sStage0 + stage * SM::STAGE, sK_lo = sK_hi + SM::K_SLOT;
          const uint32_t sV_hi = sK_lo + SM::K_SLOT, sV_lo = sV_hi + SM::V_SLOT;
          const uint32_t fb = a_smem_u32(&bar_full[stage]);
          const int krow = bh * P.L + t_lo + (jb0 + j) * AT_BKV;
          a_tma_2d(sK_hi, &P.tmK_hi, fb, 0, krow);
          a_tma_2d(sK_lo, &P.tmK_lo, fb, 0, krow);
          a_tma_2d(sV_hi, &P.tmV_hi, fb, t_lo + (jb0 + j) * AT_BKV, bh * D);
          a_tma_2d(sV_lo, &P.tmV_lo, fb, t_lo + (jb0 + j) * AT_BKV, bh * D);
        }
        __syncwarp();
        if (++stage == AT_NS) { stage = 0; phase ^= 1u; }
      }
    }
    a_mbar_wait_guard(&bar_pv_done[(nblk - 1) & 1], last_pv_par);   // all threads trigger when only the epilogue remains
    MTV_PDL_TRIGGER();
    if (kvs > 1) { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); a_cluster_sync(); }
  } else if (warp == 1) {
    // =============================== MMA issuer =================================
    // whole warp in the loop, one elected lane issues; descriptors = base descriptor + (byte offset >> 4)
    {
      const uint64_t dQK = a_desc(smem0, SM::ROWB);          // Q / K tiles: swizzle span = row bytes
      const uint64_t dPV = a_desc(smem0, 128);               // P / V^T tiles: 128-byte rows
      auto off16 = [&](uint32_t addr) { return (uint64_t)((addr - smem0) >> 4); };
      auto issue_S = [&](int stage) {
        const uint32_t sK_hi = sStage0 + stage * SM::STAGE, sK_lo = sK_hi + SM::K_SLOT;
        const uint64_t qh0 = dQK + off16(sQ_hi), ql0 = dQK + off16(sQ_lo), kh0 = dQK + off16(sK_hi), kl0 = dQK + off16(sK_lo);
        if (a_elect_one()) {
#pragma unroll
          for (int k = 0; k < D / 16; ++k) {
            a_mma(tmem_S, qh0 + 2 * k, kh0 + 2 * k, IDESC_S, k > 0 ? 1u : 0u);
            a_mma(tmem_S, ql0 + 2 * k, kh0 + 2 * k, IDESC_S, 1u);
            a_mma(tmem_S, qh0 + 2 * k, kl0 + 2 * k, IDESC_S, 1u);
          }
          a_commit(&bar_s_full);
        }
        __syncwarp();
      };
      a_mbar_wait_guard(&bar_q, 0);
      a_mbar_wait_guard(&bar_full[0], 0);
      a_fence_after();
      issue_S(0);
      int stage = 0; uint32_t phase = 0;
      for (int j = 0; j < nblk; ++j) {
        int nstage = stage + 1; uint32_t nphase = phase;
        if (nstage == AT_NS) { nstage = 0; nphase ^= 1u; }
        if (j + 1 < nblk) {
          a_mbar_wait_guard(&bar_full[nstage], nphase);
          a_mbar_wait_guard(&bar_s_free, (uint32_t)(j & 1));      // softmax has read S(j) out of TMEM
          a_fence_after();
          issue_S(nstage);
        }
        a_mbar_wait_guard(&bar_p_full[j & 1], (uint32_t)((j >> 1) & 1));   // P(j) is in tensor memory (and every correction of O is done)
        a_fence_after();
        const uint32_t sV_hi = sStage0 + stage * SM::STAGE + 2 * SM::K_SLOT, sV_lo = sV_hi + SM::V_SLOT;
        const uint64_t vh0 = dPV + off16(sV_hi), vl0 = dPV + off16(sV_lo);
        (void)vl0;
        const uint32_t tP_hi = tmem_P + (uint32_t)((j & 1) * 64), tP_lo = tP_hi + 32;
        if (a_elect_one()) {
#pragma unroll
          for (int k = 0; k < AT_BKV / 16; ++k) {       // O += P(j) V(j): the accumulator lives in tensor memory across key blocks
            if constexpr (STACK_V) {
              a_mma_ts(tmem_O, tP_hi + 8 * k, vh0 + 2 * k, IDESC_O2, (j | k) ? 1u : 0u);
              a_mma_ts(tmem_O, tP_lo + 8 * k, vh0 + 2 * k, IDESC_O, 1u);
            } else {
              a_mma_ts(tmem_O, tP_hi + 8 * k, vh0 + 2 * k, IDESC_O, (j | k) ? 1u : 0u);
              a_mma_ts(tmem_O, tP_lo + 8 * k, vh0 + 2 * k, IDESC_O, 1u);
              a_mma_ts(tmem_O, tP_hi + 8 * k, vl0 + 2 * k, IDESC_O, 1u);
            }
          }
          a_commit(&bar_pv_done[j & 1]);
          a_commit(&bar_empty[stage]);
        }
        __syncwarp();
        stage = nstage; phase = nphase;
      }
    }
    a_mbar_wait_guard(&bar_pv_done[(nblk - 1) & 1], last_pv_par);
    MTV_PDL_TRIGGER();
    if (kvs > 1) { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); a_cluster_sync(); }
  } else {
    // =============================== softmax / epilogue ==========================
    // warps 2-5: key columns [0,32) of each S block and output columns [0, D/2);
    // warps 6-9: key columns [32,64) and output columns [D/2, D).  Thread == (row, half).
    constexpr int DH = D / 2;
    const int half = (warp - 2) >> 2;             // 0 or 1
    const int qq = warp & 3;                      // TMEM lane quarter this warp may read
    const int row = qq * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(qq * 32) << 16;
    // Lazy rescaling: O accumulates in tensor memory under a reference maximum m_run that is only raised when the row maximum
    // has grown by more than 2^8 (then O and l are rescaled once, exactly); otherwise p = 2^(s - m_run) <= 256 is used as is.
    // The softmax warps therefore never wait for a PV product in the steady state: the chain per key block is
    // S -> tcgen05.ld -> max -> ex2 -> P (tcgen05.st into the buffer the MMA warp is not reading) -> arrive.
    float m_run = -INFINITY, l_part = 0.f;
    float o[DH];
    // diagnostics (MTV_ATTN_DBG_SKIP bit 5): cycles per phase of the softmax loop, summed over the key blocks, thread 64 of CTA (0, 0)
    const bool prof = (P.dbg_skip & 32) && threadIdx.x == 64 && blockIdx.x == 0 && blockIdx.y == 0;
    long long tph[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tq = 0;
#define AT_STAMP(i) do { if (prof) { const long long t_ = clock64(); tph[i] += t_ - tq; tq = t_; } } while (0)
    if (prof) tq = clock64();

    for (int j = 0; j < nblk; ++j) {
      a_mbar_wait(&bar_s_full, (uint32_t)(j & 1));
      a_fence_after();
      AT_STAMP(0);
      float s[32];
      {
        uint32_t r[32];
        a_tmem_ld32(tmem_S + lane_addr + (uint32_t)(half * 32), r);
        a_tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) s[i] = __uint_as_float(r[i]);
      }
      a_fence_before();
      __syncwarp();
      if (lane == 0) a_mbar_arrive(&bar_s_free);            // S TMEM may be overwritten by S(j+1)
      AT_STAMP(1);
      const int valid = len - (jb0 + j) * AT_BKV - half * 32;       // >= 32 except in the segment's last block (may be <= 0 there)
      if (valid < 32) {
#pragma unroll
        for (int i = 0; i < 32; ++i) if (i >= valid) s[i] = -INFINITY;
      }
      float mx = fmaxf(s[0], s[1]);
#pragma unroll
      for (int i = 2; i < 32; i += 2) mx = max3(mx, s[i], s[i + 1]);
      // row max over both halves
      if (!(P.dbg_skip & 4)) {
      s_xchg[j & 1][half][row] = mx;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mx = fmaxf(mx, s_xchg[j & 1][half ^ 1][row]);
      }
      AT_STAMP(2);
      // both threads of a row see the same mx and take the same decision (finite: the lower half always holds a valid key)
      if (j == 0) {
        m_run = mx;
      } else if (mx > m_run + 8.0f) {
        // rare after the first blocks: rescale this row of O (our DH columns) and l.  PV(j-1) must have finished accumulating.
        a_mbar_wait(&bar_pv_done[(j - 1) & 1], (uint32_t)(((j - 1) >> 1) & 1));
        a_fence_after();
        const float corr = ex2_approx(m_run - mx);
#pragma unroll
        for (int g = 0; g < (STACK_V ? 2 : 1); ++g) {
#pragma unroll
          for (int c = 0; c < DH; c += 8) {
            uint32_t r[8];
            a_tmem_ld8(tmem_O + lane_addr + (uint32_t)(g * D + half * DH + c), r);
            a_tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 8; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * corr);
            a_tmem_st8(tmem_O + lane_addr + (uint32_t)(g * D + half * DH + c), r);
          }
        }
        a_tmem_wait_st();
        l_part *= corr;
        m_run = mx;
      }
      AT_STAMP(3);
      float sum0 = 0.f, sum1 = 0.f;
      if (!(P.dbg_skip & 1)) {
        const uint64_t mm = f2_pack(m_run, m_run);
        uint64_t sums = f2_pack(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float x0, x1; f2_unpack(f2_sub(f2_pack(s[i], s[i + 1]), mm), x0, x1);
          s[i] = ex2_approx(x0); s[i + 1] = ex2_approx(x1);
          sums = f2_add(sums, f2_pack(s[i], s[i + 1]));
        }
        f2_unpack(sums, sum0, sum1);
      } else {
#pragma unroll
      for (int i = 0; i < 32; i += 2) { s[i] = s[i] - m_run; s[i + 1] = s[i + 1] - m_run; sum0 += s[i]; sum1 += s[i + 1]; }
      }
      l_part += sum0 + sum1;
      // this thread's 32 keys of its row -> 16 packed columns each of P_hi / P_lo (key 2e in the low half of column e):
      // hi = bf16x2(p), lo = bf16x2(p - float(hi))
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const float p0 = s[2 * e], p1 = s[2 * e + 1];
        hi[e] = cvt_bf16x2(p0, p1);
        float l0, l1;
        f2_unpack(f2_sub(f2_pack(p0, p1), f2_pack(__uint_as_float(hi[e] << 16), __uint_as_float(hi[e] & 0xffff0000u))), l0, l1);
        lo[e] = cvt_bf16x2(l0, l1);
      }
      AT_STAMP(4);
      // P buffer (j & 1) was last read by PV(j-2)
      if (j >= 2) { a_mbar_wait(&bar_pv_done[j & 1], (uint32_t)(((j >> 1) - 1) & 1)); a_fence_after(); }
      AT_STAMP(5);
      if (!(P.dbg_skip & 2)) {
        const uint32_t tP = tmem_P + (uint32_t)((j & 1) * 64) + lane_addr + (uint32_t)(half * 16);
        a_tmem_st16(tP, hi);
        a_tmem_st16(tP + 32, lo);
        a_tmem_wait_st();
      }
      a_fence_before();
      __syncwarp();
      if (lane == 0) a_mbar_arrive(&bar_p_full[j & 1]);
      AT_STAMP(6);
      // s_xchg is double-buffered by block parity: slot (j & 1) is rewritten in block j + 2, i.e. after the
      // bar.sync of block j + 1, which every reader of block j has passed by then
    }
    a_mbar_wait(&bar_pv_done[(nblk - 1) & 1], last_pv_par);
    AT_STAMP(7);
    if (prof)
      printf("attn D=%d L=%d nblk=%d cycles/block: wait_S %lld ld_S %lld max+xchg %lld corr %lld exp+cvt %lld wait_Pbuf %lld st_P+arrive %lld | tail %lld\n", D, P.L, nblk,
             tph[0] / nblk, tph[1] / nblk, tph[2] / nblk, tph[3] / nblk, tph[4] / nblk, tph[5] / nblk, tph[6] / nblk, tph[7]);
#undef AT_STAMP
    MTV_PDL_TRIGGER();
    a_fence_after();
#pragma unroll
    for (int c = 0; c < DH; c += 8) {     // the finished accumulator: this half's columns of O
      uint32_t r[8];
      a_tmem_ld8(tmem_O + lane_addr + (uint32_t)(half * DH + c), r);
      a_tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 8; ++i) o[c + i] = __uint_as_float(r[i]);
      if constexpr (STACK_V) {
        a_tmem_ld8(tmem_O + lane_addr + (uint32_t)(D + half * DH + c), r);
        a_tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 8; ++i) o[c + i] += __uint_as_float(r[i]);
      }
    }
    // total row sum = both halves' partial sums
    asm volatile("bar.sync 1, 256;" ::: "memory");          // all max-exchange reads done before the slots are reused
    s_xchg[0][half][row] = l_part;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    float l_run = l_part + s_xchg[0][half ^ 1][row];
    if (kvs > 1) {
      // partial results of ranks > 0 -> rank 0's merge buffer [rank - 1][row][D + 2] = O[D] | m | l (distributed shared memory)
      float* mbuf = reinterpret_cast<float*>(smem_raw + (smem0 - a_smem_u32(smem_raw)) + SM::TOTAL - 1024);
      constexpr int MW = D + 2;
      asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");     // the start-of-kernel arrival: every peer CTA is running
      if (kr > 0) {
        const uint32_t local = a_smem_u32(mbuf + ((size_t)(kr - 1) * AT_BQ + row) * MW + half * DH);
        uint32_t remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(0));
#pragma unroll
        for (int d = 0; d < DH; ++d) asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote + 4u * d), "f"(o[d]) : "memory");
        if (half == 0) {
          const uint32_t ml = remote + 4u * D;
          asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ml), "f"(m_run) : "memory");
          asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ml + 4u), "f"(l_run) : "memory");
        }
      }
      a_cluster_sync();
      if (kr == 0) {
        for (int r = 1; r < kvs; ++r) {        // fixed order: deterministic
          const float* src = mbuf + ((size_t)(r - 1) * AT_BQ + row) * MW;
          const float m_r = src[D], l_r = src[D + 1];
          const float m_new = fmaxf(m_run, m_r);
          const float s0 = ex2_approx(m_run - m_new), s1 = ex2_approx(m_r - m_new);
#pragma unroll
          for (int d = 0; d < DH; ++d) o[d] = o[d] * s0 + src[half * DH + d] * s1;
          l_run = l_run * s0 + l_r * s1;
          m_run = m_new;
        }
      }
    }
    const int q = q0 + row;
    if (q < len && kr == 0) {
      const int b = bh / P.heads, h = bh - b * P.heads;
      const float inv = 1.0f / l_run;
      const size_t oidx = ((size_t)b * P.L + t_lo + q) * P.C + h * D + half * DH;
      if (P.out_hi) {       // split-bf16 A operand of the proj_out GEMM, written in place of the fp32 tensor
        __nv_bfloat16* oh = reinterpret_cast<__nv_bfloat16*>(P.out_hi) + oidx;
        __nv_bfloat16* ol = reinterpret_cast<__nv_bfloat16*>(P.out_lo) + oidx;
#pragma unroll
        for (int d = 0; d < DH; d += 8) {
          __align__(16) __nv_bfloat16 hh[8], ll[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float val = o[d + i] * inv;
            hh[i] = __float2bfloat16_rn(val); ll[i] = __float2bfloat16_rn(val - __bfloat162float(hh[i]));
          }
          *reinterpret_cast<uint4*>(oh + d) = *reinterpret_cast<const uint4*>(hh);
          *reinterpret_cast<uint4*>(ol + d) = *reinterpret_cast<const uint4*>(ll);
        }
      } else {
        float* dst = P.out + oidx;
#pragma unroll
        for (int d = 0; d < DH; d += 4)
          *reinterpret_cast<float4*>(dst + d) = make_float4(o[d] * inv, o[d + 1] * inv, o[d + 2] * inv, o[d + 3] * inv);
      }
    }
  }
  a_fence_before();
  __syncthreads();
  if (warp == 1) {
    a_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

template <int D>
static cudaError_t launch_attn_tc_d(const AttnTcParams& P, cudaStream_t s) {
  using SM = AttnSmem<D>;
  const int kvs = P.kv_split > 1 ? P.kv_split : 1;
  if (kvs > 8) return cudaErrorInvalidValue;
  int nqb = 0;
  for (int i = 0; i < P.nseg; ++i) {
    const int len = P.seg_off[i + 1] - P.seg_off[i];
    nqb += (len + AT_BQ - 1) / AT_BQ;
    if ((len + AT_BKV - 1) / AT_BKV < kvs) return cudaErrorInvalidValue;      // every rank needs at least one key block
  }
  const size_t smem = (size_t)SM::TOTAL + (kvs > 1 ? (size_t)(kvs - 1) * AT_BQ * (D + 2) * sizeof(float) : 0);
  if (smem > 227u * 1024u) return cudaErrorInvalidValue;
  // the attribute is a per-function maximum: keep it at the largest size any launch of this head dim has asked for
  static size_t smem_max = 0;
  cudaError_t e = cudaSuccess;
  if (smem > smem_max) { e = cudaFuncSetAttribute(k_attn_tc<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); smem_max = smem; }
  if (e != cudaSuccess) return e;
  dim3 grid(nqb * kvs, P.B * P.heads);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(AT_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = mtv_pdl_enabled(PDL_CLASS_ATTN_TC) ? 1 : 0;
  at[1].id = cudaLaunchAttributeClusterDimension;
  at[1].val.clusterDim.x = (unsigned)kvs; at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = kvs > 1 ? 2 : 1;
  e = cudaLaunchKernelEx(&cfg, k_attn_tc<D>, P);
  if (e != cudaSuccess) return e;
  return cudaGetLastError();
}

cudaError_t launch_attn_tc(const AttnTcParams& P, cudaStream_t s) {
  switch (P.C / P.heads) {
    case 16: return launch_attn_tc_d<16>(P, s);
    case 32: return launch_attn_tc_d<32>(P, s);
    case 64: return launch_attn_tc_d<64>(P, s);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace mtv
