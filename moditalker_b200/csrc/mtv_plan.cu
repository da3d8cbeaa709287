// mtv_plan.cu — handle, weight store, launch plan and the C ABI of libmtv_b200.so.
//
// The plan is the B200 replacement for UNetModel.forward's Python control flow
// (MToV/models/ddpm/unet.py:995-1117): the architecture is walked ONCE per batch
// size into a flat list of kernel launches over pre-allocated HBM buffers; the body
// of the list is then captured into a CUDA graph, so a denoising step is
// {copy t, pack inputs, one graph launch, copy eps} instead of ~2300 framework
// kernels.
#include "../../include/mtv_b200.h"
#include "mtv_kernels.cuh"
#include "mtv_tc.cuh"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <set>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

using namespace mtv;

namespace {

thread_local std::string g_err;

struct MtvError : std::runtime_error { using std::runtime_error::runtime_error; };
#define CK(expr)                                                                               \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      throw MtvError(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + \
                     std::to_string(__LINE__) + ")");                                          \
  } while (0)

// ------------------------------------------------------------------ architecture walk
// Mirrors moditalker_b200/arch.py (same traversal of unet.py:710-975); the Python
// test tests/test_arch.py compares the two weight-name lists.
struct ResDesc { std::string name; int cin = 0, cout = 0; int updown = RS_NONE; int film_off = 0; };
struct AttnDesc { std::string name; int C = 0; bool joint = false; };
struct Layer { bool is_res = false; ResDesc r; AttnDesc a; };
struct StageDesc {
  std::vector<Layer> layers; bool has_joint = false; AttnDesc joint;
  int skip_ch = 0, level_in = 0, level_out = 0;
};
struct Arch {
  std::vector<StageDesc> in; StageDesc mid; std::vector<StageDesc> out;
  int head_ch = 0; int J = 0;
};

Arch build_arch(const MtvConfig& c) {
  Arch A;
  const int mc = c.model_channels;
  auto res_layer = [&](const std::string& n, int cin, int cout, int ud) {
    Layer l; l.is_res = true; l.r.name = n; l.r.cin = cin; l.r.cout = cout; l.r.updown = ud;
    l.r.film_off = A.J; A.J += 2 * cout; return l;
  };
  auto attn_layer = [&](const std::string& n, int C, bool joint) {
    Layer l; l.is_res = false; l.a.name = n; l.a.C = C; l.a.joint = joint; return l;
  };
  std::vector<int> skip = {mc};
  A.in.emplace_back();   // stem
  int ch = mc, level = 0, idx = 1;
  for (int li = 0; li < c.num_levels; ++li) {
    const int mult = c.channel_mult[li];
    for (int k = 0; k < c.num_res_blocks; ++k) {
      StageDesc st; st.level_in = st.level_out = level;
      const std::string p = "input_blocks." + std::to_string(idx);
      st.layers.push_back(res_layer(p + ".0", ch, mult * mc, RS_NONE));
      ch = mult * mc;
      if (c.attn_at_level[li]) st.layers.push_back(attn_layer(p + ".1", ch, false));
      st.has_joint = true; st.joint = attn_layer("input_attns." + std::to_string(idx), ch, true).a;
      A.in.push_back(st); skip.push_back(ch); ++idx;
    }
    if (li != c.num_levels - 1) {
      StageDesc st; st.level_in = level; st.level_out = level + 1;
      st.layers.push_back(res_layer("input_blocks." + std::to_string(idx) + ".0", ch, ch, RS_DOWN2));
      st.has_joint = true; st.joint = attn_layer("input_attns." + std::to_string(idx), ch, true).a;
      A.in.push_back(st); skip.push_back(ch); ++idx; ++level;
    }
  }
  A.mid.level_in = A.mid.level_out = level;
  A.mid.layers.push_back(res_layer("middle_block.0", ch, ch, RS_NONE));
  A.mid.layers.push_back(attn_layer("middle_block.1", ch, false));
  A.mid.layers.push_back(res_layer("middle_block.2", ch, ch, RS_NONE));
  A.mid.has_joint = true; A.mid.joint = attn_layer("mid_attn", ch, true).a;
  int oidx = 0;
  for (int li = c.num_levels - 1; li >= 0; --li) {
    const int mult = c.channel_mult[li];
    for (int i = 0; i <= c.num_res_blocks; ++i) {
      const int ich = skip.back(); skip.pop_back();
      StageDesc st; st.skip_ch = ich; st.level_in = st.level_out = level;
      const std::string p = "output_blocks." + std::to_string(oidx);
      st.layers.push_back(res_layer(p + ".0", ch + ich, mc * mult, RS_NONE));
      ch = mc * mult;
      int nxt = 1;
      if (c.attn_at_level[li]) { st.layers.push_back(attn_layer(p + "." + std::to_string(nxt), ch, false)); ++nxt; }
      if (li && i == c.num_res_blocks) {
        st.layers.push_back(res_layer(p + "." + std::to_string(nxt), ch, ch, RS_UP2));
        --level; st.level_out = level;
      }
      st.has_joint = true; st.joint = attn_layer("output_attns." + std::to_string(oidx), ch, true).a;
      A.out.push_back(st); ++oidx;
    }
  }
  A.head_ch = ch;
  return A;
}

// ------------------------------------------------------------------ weights
enum { WK_PLAIN = 0, WK_CONV = 1 };
struct Weight {
  std::string name; std::vector<int64_t> shape; size_t elems = 0;
  float* dev = nullptr; bool owned = true; bool loaded = false; int kind = WK_PLAIN;
  void* hi = nullptr; void* lo = nullptr;   // split-bf16 K-major copy [taps][Cout][Cin] for the tcgen05 path
  int pad_cin = 0, pad_cout = 0;            // > 0: the split copy is zero-padded to this many input (stem) / output (head) channels
};

struct Tensor {
  float* p = nullptr; int C = 0; int level = 0; double* csum = nullptr; /* [B][3][C][2] per-channel sums, or null */
};

struct RunCtx {
  const float* x = nullptr; const float* cond = nullptr; const float* image_cond = nullptr;
  int64_t ic_len = 0; const int64_t* t = nullptr; float* out = nullptr;
};

struct Op {
  std::string name; int phase = 1;      // 0 = before graph (reads caller pointers), 1 = graph body, 2 = after
  int launches = 1; double flops = 0, bytes = 0;
  std::function<cudaError_t(cudaStream_t)> fn;
  std::shared_ptr<TcConvParams> tc;     // tensor-core tap-GEMMs keep their parameter block (late edits: fused consumer, prefetch)
};

struct Plan {
  int B = 0;
  std::vector<Op> ops;
  std::vector<void*> allocs; size_t alloc_bytes = 0;
  std::map<std::string, Tensor> taps;
  RunCtx ctx;
  cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr; int runs = 0; unsigned long long last_use = 0;
  char* csum_arena = nullptr; size_t csum_cap = 0, csum_used = 0;
  ~Plan() {
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
    for (void* p : allocs) cudaFree(p);
  }
};

}  // namespace

struct MtvHandle_t {
  MtvConfig cfg{}; Arch arch; int num_sms = 148;
  std::vector<Weight> weights; std::unordered_map<std::string, int> windex;
  std::vector<void*> allocs;
  float* emb_wall = nullptr; float* emb_ball = nullptr; float* freqs = nullptr;
  float* head_bias_pad = nullptr;    // out.2.bias zero-padded to 64 (tensor-core head conv)
  std::unordered_map<std::string, float*> bias_sum;   // ResBlock name -> conv2.bias + skip.bias
  std::unordered_map<const float*, std::pair<void*, void*>> tc_w;   // fp32 conv weight -> split-bf16 pair
  bool dirty = true; bool use_graph = true;
  std::map<int, std::unique_ptr<Plan>> plans;
  Plan* last_plan = nullptr; unsigned long long use_clock = 0;
  int64_t weight_bytes = 0;
  // feature bits (MTV_TC_MASK): 0-4 op classes on the tensor-core kernel, 5 split-K, 6 tcgen05 attention, 7 small levels,
  // 8 fused GroupNorm statistics, 9 launch fusions, 10 weight L2 prefetch, 11 L2-persisting small-tensor arena;
  // 15 BN = 128 tiles for every split-K-able op, 16 TMA stores of the GEMM epilogue tiles.  (Bits 12 / 13 / 14 were the
  // persistent chain kernel, the direct A operand and the consumer apply fused into the producer behind a grid barrier; a
  // bit 17 fused GroupNorm + qkv into the attention kernel as a cluster front end: all measured slower on B200 and removed —
  // profiles/r01_s2_chain_experiment.md, r01_s2_direct_experiment.md, r02_fused_apply_experiment.md,
  // r02_fused_attention_experiment.md.)
  // 18 stem conv on the tensor cores (input channels padded 16 -> 64; its channel sums replace two k_gn_stats launches)
  int tc_mask = 0x5cfff;
  cudaStream_t cap_stream = nullptr;
  cudaStream_t capture_stream() {
    if (!cap_stream) CK(cudaStreamCreateWithFlags(&cap_stream, cudaStreamNonBlocking));
    return cap_stream;
  }

  float* dalloc(size_t bytes) {
    void* p = nullptr; CK(cudaMalloc(&p, bytes)); allocs.push_back(p); return (float*)p;
  }
  // Small, reused-every-step tensors (biases, GroupNorm affines, the per-step FiLM table) live in one arena
  // that is marked L2-persisting for every kernel of the forward: read once per step, they would otherwise be
  // evicted by the 529 MB weight stream and cost a DRAM miss on the critical path of each short kernel.
  char* small_arena = nullptr; size_t small_cap = 0, small_used = 0;
  float* small_alloc(size_t bytes) {
    if (!small_arena) { small_cap = 8u << 20; small_arena = (char*)dalloc(small_cap); }
    const size_t b = (bytes + 255) & ~(size_t)255;
    if (small_used + b > small_cap) return nullptr;
    float* p = (float*)(small_arena + small_used); small_used += b; return p;
  }
  // The access-policy window is set ONLY on the library's private capture stream (graph kernel nodes inherit it); the
  // caller's stream is never modified.  The device-wide persisting-L2 limit is raised if needed and restored in mtv_destroy.
  bool l2_limit_changed = false; size_t l2_limit_prev = 0; size_t l2_window_bytes = 0;
  void apply_l2_window(cudaStream_t private_stream) {
    if (!small_arena || !small_used || l2_window_bytes == small_used) return;
    if (!l2_limit_changed) {
      size_t prev = 0;
      if (cudaDeviceGetLimit(&prev, cudaLimitPersistingL2CacheSize) == cudaSuccess && prev < small_cap) {
        if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, small_cap) == cudaSuccess) { l2_limit_changed = true; l2_limit_prev = prev; }
        else cudaGetLastError();
      }
    }
    cudaStreamAttrValue attr; memset(&attr, 0, sizeof(attr));
    attr.accessPolicyWindow.base_ptr = small_arena;
    attr.accessPolicyWindow.num_bytes = small_used;
    attr.accessPolicyWindow.hitRatio = 1.0f;
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if (cudaStreamSetAttribute(private_stream, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) cudaGetLastError();
    l2_window_bytes = small_used;
  }
  void release_l2_window() {
    if (l2_limit_changed) {
      cudaCtxResetPersistingL2Cache();
      cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, l2_limit_prev);
      cudaGetLastError();
      l2_limit_changed = false;
    }
  }
  int add_weight(const std::string& name, std::vector<int64_t> shape, int kind, float* view = nullptr) {
    Weight w; w.name = name; w.shape = shape; w.kind = kind; w.elems = 1;
    for (auto d : shape) w.elems *= (size_t)d;
    if (view) { w.dev = view; w.owned = false; }
    else {
      w.dev = (shape.size() == 1 && ((tc_mask >> 11) & 1)) ? small_alloc(w.elems * sizeof(float)) : nullptr;
      if (!w.dev) w.dev = dalloc(w.elems * sizeof(float));
    }
    if (kind == WK_CONV && cfg.kernel_path != 1 && shape[0] % 64 == 0 && shape[1] % 64 == 0) {
      w.hi = dalloc(w.elems * 2); w.lo = dalloc(w.elems * 2);
      tc_w[w.dev] = std::make_pair(w.hi, w.lo);
    } else if (kind == WK_CONV && cfg.kernel_path != 1 && name == "input_blocks.0.0.weight" && shape[0] % 64 == 0 && shape[1] < 64) {
      // stem conv on the tensor cores: its 4*in_channels input channels are zero-padded to one 64-channel K chunk
      const size_t padded = w.elems / (size_t)shape[1] * 64;
      w.hi = dalloc(padded * 2); w.lo = dalloc(padded * 2); w.pad_cin = 64;
      tc_w[w.dev] = std::make_pair(w.hi, w.lo);
    } else if (kind == WK_CONV && cfg.kernel_path != 1 && name == "out.2.weight" && shape[0] < 64 && shape[1] % 64 == 0) {
      // head conv on the tensor cores: its out_channels output channels are zero-padded to one 64-column tile
      const size_t padded = w.elems / (size_t)shape[0] * 64;
      w.hi = dalloc(padded * 2); w.lo = dalloc(padded * 2); w.pad_cout = 64;
      tc_w[w.dev] = std::make_pair(w.hi, w.lo);
    }
    weight_bytes += (int64_t)w.elems * 4;
    windex[name] = (int)weights.size(); weights.push_back(w);
    return (int)weights.size() - 1;
  }
  float* W(const std::string& name) const {
    auto it = windex.find(name);
    if (it == windex.end()) throw MtvError("internal: unknown weight " + name);
    return weights[it->second].dev;
  }
};

namespace {

// every ABI entry point runs on the handle's device and leaves the caller's current device as it found it
struct DeviceGuard {
  int prev = -1; bool switched = false;
  explicit DeviceGuard(int dev) {
    CK(cudaGetDevice(&prev));
    if (prev != dev) { CK(cudaSetDevice(dev)); switched = true; }
  }
  ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
  DeviceGuard(const DeviceGuard&) = delete; DeviceGuard& operator=(const DeviceGuard&) = delete;
};

Geo level_geo(const MtvConfig& c, int level) { return make_geo(c.image_size >> level, (c.image_size / 2) >> level); }

void register_weights(MtvHandle_t* h) {
  const MtvConfig& c = h->cfg; const Arch& A = h->arch;
  const int mc = c.model_channels, ted = 4 * mc;
  h->emb_wall = h->dalloc((size_t)A.J * ted * sizeof(float));
  h->emb_ball = h->dalloc((size_t)A.J * sizeof(float));
  h->add_weight("time_embed.0.weight", {ted, mc}, WK_PLAIN);
  h->add_weight("time_embed.0.bias", {ted}, WK_PLAIN);
  h->add_weight("time_embed.2.weight", {ted, ted}, WK_PLAIN);
  h->add_weight("time_embed.2.bias", {ted}, WK_PLAIN);
  h->add_weight("input_blocks.0.0.weight", {mc, 4 * c.in_channels, 3, 3}, WK_CONV);
  h->add_weight("input_blocks.0.0.bias", {mc}, WK_PLAIN);
  auto reg_res = [&](const ResDesc& r) {
    const std::string& p = r.name;
    h->add_weight(p + ".in_layers.0.weight", {r.cin}, WK_PLAIN);
    h->add_weight(p + ".in_layers.0.bias", {r.cin}, WK_PLAIN);
    h->add_weight(p + ".in_layers.2.weight", {r.cout, r.cin, 3, 3}, WK_CONV);
    h->add_weight(p + ".in_layers.2.bias", {r.cout}, WK_PLAIN);
    h->add_weight(p + ".emb_layers.1.weight", {2 * r.cout, ted}, WK_PLAIN, h->emb_wall + (size_t)r.film_off * ted);
    h->add_weight(p + ".emb_layers.1.bias", {2 * r.cout}, WK_PLAIN, h->emb_ball + r.film_off);
    h->add_weight(p + ".out_layers.0.weight", {r.cout}, WK_PLAIN);
    h->add_weight(p + ".out_layers.0.bias", {r.cout}, WK_PLAIN);
    h->add_weight(p + ".out_layers.3.weight", {r.cout, r.cout, 3, 3}, WK_CONV);
    h->add_weight(p + ".out_layers.3.bias", {r.cout}, WK_PLAIN);
    if (r.cin != r.cout) {
      h->add_weight(p + ".skip_connection.weight", {r.cout, r.cin, 1, 1}, WK_CONV);
      h->add_weight(p + ".skip_connection.bias", {r.cout}, WK_PLAIN);
      h->bias_sum[p] = h->dalloc((size_t)r.cout * sizeof(float));
    }
  };
  auto reg_attn = [&](const AttnDesc& a) {
    const std::string& p = a.name;
    h->add_weight(p + ".norm.weight", {a.C}, WK_PLAIN);
    h->add_weight(p + ".norm.bias", {a.C}, WK_PLAIN);
    h->add_weight(p + ".qkv.weight", {3 * a.C, a.C, 1}, WK_CONV);
    h->add_weight(p + ".qkv.bias", {3 * a.C}, WK_PLAIN);
    h->add_weight(p + ".proj_out.weight", {a.C, a.C, 1}, WK_CONV);
    h->add_weight(p + ".proj_out.bias", {a.C}, WK_PLAIN);
  };
  auto reg_stage = [&](const StageDesc& st) {
    for (const Layer& l : st.layers) { if (l.is_res) reg_res(l.r); else reg_attn(l.a); }
    if (st.has_joint) reg_attn(st.joint);
  };
  for (size_t i = 1; i < A.in.size(); ++i) reg_stage(A.in[i]);
  reg_stage(A.mid);
  for (const StageDesc& st : A.out) reg_stage(st);
  h->add_weight("out.0.weight", {A.head_ch}, WK_PLAIN);
  h->add_weight("out.0.bias", {A.head_ch}, WK_PLAIN);
  h->add_weight("out.2.weight", {c.out_channels, mc, 3, 3}, WK_CONV);
  h->add_weight("out.2.bias", {c.out_channels}, WK_PLAIN);
  if (c.kernel_path != 1 && c.out_channels < 64) {
    h->head_bias_pad = h->dalloc(64 * sizeof(float));
    CK(cudaMemset(h->head_bias_pad, 0, 64 * sizeof(float)));
  }

  // timestep_embedding frequencies, fp32 like the reference (diffusionmodules.py:118-121)
  const int half = mc / 2;
  std::vector<float> fr(half);
  const float neg_log = (float)(-std::log(10000.0));
  for (int i = 0; i < half; ++i) fr[i] = expf(neg_log * (float)i / (float)half);
  h->freqs = h->dalloc(half * sizeof(float));
  CK(cudaMemcpy(h->freqs, fr.data(), half * sizeof(float), cudaMemcpyHostToDevice));
}

void ensure_ready(MtvHandle_t* h, cudaStream_t s) {
  if (!h->dirty) return;
  for (const Weight& w : h->weights)
    if (!w.loaded) throw MtvError("weight not loaded: " + w.name);
  for (auto& kv : h->bias_sum) {
    const std::string& p = kv.first;
    const int n = (int)h->weights[h->windex.at(p + ".out_layers.3.bias")].elems;
    CK(launch_add_vec(h->W(p + ".out_layers.3.bias"), h->W(p + ".skip_connection.bias"), kv.second, n, s));
  }
  if (h->head_bias_pad)
    CK(cudaMemcpyAsync(h->head_bias_pad, h->W("out.2.bias"), (size_t)h->cfg.out_channels * sizeof(float), cudaMemcpyDeviceToDevice, s));
  h->dirty = false;
}

// ------------------------------------------------------------------ TMA descriptors
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (!p || q != cudaDriverEntryPointSuccess) throw MtvError("cuTensorMapEncodeTiled entry point unavailable");
    fn = (EncodeTiledFn)p;
  }
  return fn;
}
// bf16 tensor, innermost dimension contiguous, 128-byte swizzle, zero fill out of bounds
CUtensorMap make_tmap_bf16(void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                           int swizzle_bytes = 128, bool f32 = false) {
  CUtensorMap m;
  cuuint64_t gd[5]; cuuint64_t gs[4]; cuuint32_t bx[5]; cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUresult r = encode_tiled_fn()(&m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, base, gd, gs,
                                 bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 swizzle_bytes == 0 ? CU_TENSOR_MAP_SWIZZLE_NONE :
                                 swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                                      : (swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B),
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw MtvError("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
  return m;
}

// ------------------------------------------------------------------ plan builder
struct NormRef { float* a = nullptr; float* d = nullptr; int nseg = 0; };
struct SplitBuf { void* hi = nullptr; void* lo = nullptr; };
struct TcOpts {
  const SplitBuf* pre0 = nullptr;     // seg0 operand already exists (e.g. written by the attention kernel)
  const SplitBuf* pre1 = nullptr;     // seg1 (skip) operand already exists
  SplitBuf* raw_out = nullptr;        // ask seg0's apply kernel to also emit the raw split of its source
  const QkvSplitParams* qkv = nullptr;  // qkv GEMM: epilogue writes the attention operands instead of fp32
  int chmajor_valid = 0;              // head conv: channel-major output of the first chmajor_valid columns, no split-K
};

struct Builder {
  MtvHandle_t* h; Plan* pl; int B;
  Builder(MtvHandle_t* h_, Plan* p_) : h(h_), pl(p_), B(p_->B) {}

  void* dalloc(size_t bytes, bool zero = false) {
    void* p = nullptr; CK(cudaMalloc(&p, bytes));
    if (zero) CK(cudaMemset(p, 0, bytes));
    pl->allocs.push_back(p); pl->alloc_bytes += bytes; return p;
  }
  // ---- workspace reuse.  Activations, operands and split-K scratch are short-lived: a buffer goes back to the pool once the
  // last op that reads it has been emitted.  Stream order makes that safe (an op writes only after its grid dependency
  // resolved, i.e. after every earlier op finished, PDL included).  Besides bounding the workspace, reuse keeps the dirty lines
  // of dead scratch data from being written back to HBM: a buffer that is overwritten ~10 us later is still in L2, whereas
  // one private buffer per op is evicted by the 529 MB weight stream before its next use (measured: 557 MB of DRAM
  // writes per B=1 step without reuse, profiles/r02_step_traffic.md).
  std::multimap<size_t, void*> pool_free; std::unordered_map<void*, size_t> pool_size; std::set<const void*> keep;
  static size_t pool_round(size_t b) { return (b + 4095) & ~(size_t)4095; }
  void* palloc(size_t bytes) {
    const size_t need = pool_round(bytes);
    auto it = pool_free.lower_bound(need);
    if (it != pool_free.end() && it->first <= need + need / 4) {     // best fit within 25 %
      void* p = it->second; pool_free.erase(it); return p;
    }
    void* p = dalloc(need);
    pool_size[p] = need;
    return p;
  }
  void prel(const void* p) {
    if (!p || keep.count(p)) return;
    auto it = pool_size.find(const_cast<void*>(p));
    if (it == pool_size.end()) return;                               // not a pool buffer (persistent allocation)
    for (auto r = pool_free.equal_range(it->second); r.first != r.second; ++r.first)
      if (r.first->second == p) return;                              // already released
    pool_free.emplace(it->second, it->first);
  }
  Geo geo(int level) const { return level_geo(h->cfg, level); }
  Tensor T(int C, int level) {
    Tensor t; t.C = C; t.level = level;
    t.p = (float*)palloc((size_t)B * geo(level).L * C * sizeof(float)); return t;
  }
  void Trel(const Tensor& t) { prel(t.p); }
  void set_segs(int level, bool joint, int& nseg, int* off) const {
    const Geo g = geo(level);
    if (joint) { nseg = 1; off[0] = 0; off[1] = g.L; off[2] = off[3] = g.L; }
    else { nseg = 3; off[0] = 0; off[1] = g.res * g.res; off[2] = off[1] + g.t * g.res; off[3] = g.L; }
  }

  // GroupNorm requests are recorded lazily: a tensor-core consumer whose sources carry per-channel
  // sums finalises the statistics inside its apply kernel; everything else materialises the
  // stand-alone statistics kernel (k_gn_stats) once.
  struct NormSpec {
    std::string name; Tensor x0; bool has_x1 = false; Tensor x1; bool joint = false;
    const float* gamma = nullptr; const float* beta = nullptr; int film_off = -1;
    NormRef ref; bool done = false;
  };
  std::vector<NormSpec> norms;

  int gn(const std::string& name, const Tensor& x0, const Tensor* x1, bool joint,
         const float* gamma, const float* beta, int film_off) {
    NormSpec n; n.name = name; n.x0 = x0; n.has_x1 = x1 != nullptr; if (x1) n.x1 = *x1; n.joint = joint;
    n.gamma = gamma; n.beta = beta; n.film_off = film_off;
    if ((x0.C + (x1 ? x1->C : 0)) % 32) throw MtvError("GroupNorm32 needs channels % 32 == 0 at " + name);
    norms.push_back(n);
    return (int)norms.size() - 1;
  }
  NormRef materialize(int id) {
    NormSpec& n = norms[id];
    if (n.done) return n.ref;
    GnParams P{};
    P.src0 = n.x0.p; P.C0 = n.x0.C; P.src1 = n.has_x1 ? n.x1.p : nullptr; P.C1 = n.has_x1 ? n.x1.C : 0;
    const int C = P.C0 + P.C1;
    P.B = B; P.L = geo(n.x0.level).L;
    set_segs(n.x0.level, n.joint, P.nseg, P.seg_off);
    P.gamma = n.gamma; P.beta = n.beta;
    if (n.film_off >= 0) { P.film = film_buf + n.film_off; P.film_stride = h->arch.J; }
    P.nrm_a = (float*)dalloc((size_t)B * P.nseg * C * sizeof(float));
    P.nrm_d = (float*)dalloc((size_t)B * P.nseg * C * sizeof(float));
    P.sums = (double*)dalloc((size_t)B * P.nseg * 64 * sizeof(double), true);
    P.counter = (unsigned*)dalloc((size_t)B * P.nseg * sizeof(unsigned), true);
    int maxlen = 0;
    for (int i = 0; i < P.nseg; ++i) maxlen = std::max(maxlen, P.seg_off[i + 1] - P.seg_off[i]);
    int chunk = maxlen / 32; chunk = chunk < 4 ? 4 : (chunk > 64 ? 64 : chunk);
    P.chunk_tokens = chunk;
    Op op; op.name = "gn_stats:" + n.name; op.bytes = (double)B * P.L * C * 4;
    op.fn = [P](cudaStream_t s) { return launch_gn_stats(P, s); };
    pl->ops.push_back(op);
    n.ref = NormRef{P.nrm_a, P.nrm_d, P.nseg}; n.done = true;
    return n.ref;
  }
  double* alloc_csum(int C, const Geo& g) {
    const size_t bytes = csum_elems(g, C, B) * sizeof(double);
    if (pl->csum_used + bytes > pl->csum_cap) throw MtvError("internal: csum arena exhausted");
    double* p = (double*)(pl->csum_arena + pl->csum_used);
    pl->csum_used += bytes;
    return p;
  }
  bool fuse_gn() const { return h->cfg.kernel_path != 1 && ((h->tc_mask >> 8) & 1); }

  // ---- tensor-core lowering -------------------------------------------------------------
  bool tc_ok_pb(ConvParams P) const { P.B = B; return tc_ok(P); }
  bool tc_ok(const ConvParams& P) const {
    if (h->cfg.kernel_path == 1 || P.out_chmajor) return false;
    if (P.Cout % 64) return false;
    if (P.geo.L > 128 ? (P.geo.L % 128 != 0) : (128 % P.geo.L != 0)) return false;
    {   // debug bisection mask (MTV_TC_MASK): which op classes may use the tensor-core kernel
      int level = 0;
      while ((h->cfg.image_size >> level) > P.geo.res) ++level;
      if (level >= 3 && !((h->tc_mask >> 7) & 1)) return false;          // bit 7: levels with < 128 tokens per sample
      int cls = P.nsegs == 2 ? 4 : (P.seg[0].taps == 1 ? 3 : (level > 2 ? 2 : level));
      if (!((h->tc_mask >> cls) & 1)) return false;
    }
    for (int s = 0; s < P.nsegs; ++s) {
      const KSeg& S = P.seg[s];
      if ((S.C0 + S.C1) % 64) return false;
      if (h->tc_w.find(S.w) == h->tc_w.end()) return false;
      if (s == 1 && S.taps != 1) return false;
    }
    return true;
  }
  // split-bf16 operand of one K-segment (+ optionally the un-normalised split of the same source)
  SplitBuf emit_apply(const std::string& name, const KSeg& S, const Geo& g, int norm_id, SplitBuf* raw_out,
                      const float* consumer_w = nullptr, int consumer_cout = 0) {
    const int C = S.C0 + S.C1;
    const size_t bytes = (size_t)B * g.L * C * 2;
    SplitBuf out; out.hi = palloc(bytes); out.lo = palloc(bytes);
    ApplyParams A{};
    A.src0 = S.src0; A.src1 = S.src1; A.C0 = S.C0; A.C1 = S.C1;
    A.silu = S.silu; A.resample = S.resample; A.B = B; A.geo = g; A.hi = out.hi; A.lo = out.lo;
    if (raw_out) { raw_out->hi = palloc(bytes); raw_out->lo = palloc(bytes); A.raw_hi = raw_out->hi; A.raw_lo = raw_out->lo; }
    if (consumer_w && ((h->tc_mask >> 10) & 1)) {
      auto it = h->tc_w.find(consumer_w);
      if (it != h->tc_w.end()) { A.pf0 = it->second.first; A.pf1 = it->second.second; A.pf_bytes = (unsigned long long)S.taps * C * consumer_cout * 2; }
    }
    if (norm_id >= 0) {
      const NormSpec& n = norms[norm_id];
      if (fuse_gn() && n.x0.csum && (!n.has_x1 || n.x1.csum)) {
        A.csum0 = n.x0.csum; A.csum1 = n.has_x1 ? n.x1.csum : nullptr;
        A.gamma = n.gamma; A.beta = n.beta; A.joint = n.joint ? 1 : 0;
        if (n.film_off >= 0) { A.film = film_buf + n.film_off; A.film_stride = h->arch.J; }
        // <= 2 work items (4 channels of one token) per thread: the item loop is a serial chain of L2 round trips,
        // so small levels with many channels get more, smaller CTAs rather than long loops in a handful of CTAs
        {
          const int floor_chunk = std::max(1, std::min(16, 2048 / C));
          // start large (the per-CTA statistics prologue is amortised over the chunk: 64 vs 16 tokens measured -23 % apply time at
          // B=8) and shrink only while the grid would not cover the machine twice
          int chunk = 64;
          if (const char* cm = getenv("MTV_APPLY_CHUNK")) chunk = std::max(1, atoi(cm));   // experiment knob: starting chunk
          auto ctas = [&](int ch) { return ((g.res * g.res + ch - 1) / ch) * 3 * B; };
          while (chunk > floor_chunk && ctas(chunk) < 2 * h->num_sms) chunk >>= 1;   // enough CTAs already: keep the prologue amortised
          A.chunk_tokens = chunk;
        }
      } else {
        const NormRef r = materialize(norm_id);
        A.nrm_a = r.a; A.nrm_d = r.d; A.nrm_nseg = r.nseg;
      }
    }
    Op op; op.name = "apply:" + name; op.bytes = (double)B * g.L * C * (raw_out ? 12 : 8);
    op.fn = [A](cudaStream_t s) { return launch_apply_split(A, s); };
    pl->ops.push_back(op);
    return out;
  }
  void make_A_maps(const SplitBuf& buf, int C, int taps, const Geo& g, CUtensorMap* a_hi, CUtensorMap* a_lo) {
    void* ptr[2] = {buf.hi, buf.lo};
    make_A_maps_n(ptr, 2, C, taps, g, a_hi, a_lo, false);
  }
  void make_A_maps_n(void* const* ptr, int n, int C, int taps, const Geo& g, CUtensorMap* a_hi, CUtensorMap* a_lo, bool f32) {
    CUtensorMap* dst[2] = {a_hi, a_lo};
    const bool small = g.L <= 128;
    const uint32_t spt = small ? (uint32_t)(128 / g.L) : 1u;          // samples per tile (kernels_tc.cu: tc_tile)
    const uint64_t rowb = (uint64_t)C * (f32 ? 4 : 2);
    const int sw = f32 ? 0 : 128;
    for (int k = 0; k < n; ++k) {
      char* base = (char*)ptr[k];
      if (taps == 1 && !small) {
        const uint64_t dims[2] = {(uint64_t)C, (uint64_t)B * g.L}; const uint64_t str[1] = {rowb};
        const uint32_t box[2] = {64, 128};
        dst[k][0] = make_tmap_bf16(base, 2, dims, str, box, sw, f32);
      } else if (taps == 1) {
        const uint64_t dims[3] = {(uint64_t)C, (uint64_t)g.L, (uint64_t)B}; const uint64_t str[2] = {rowb, rowb * g.L};
        const uint32_t box0[3] = {64, (uint32_t)(g.res * g.res), spt}, box1[3] = {64, (uint32_t)(g.t * g.res), spt};
        dst[k][0] = make_tmap_bf16(base, 3, dims, str, box0, sw, f32);
        dst[k][1] = make_tmap_bf16(base, 3, dims, str, box1, sw, f32);
      } else {
        const uint32_t hb_xy = small ? (uint32_t)g.res : (uint32_t)(128 / g.res);
        const uint32_t hb_pl = small ? (uint32_t)g.t : (uint32_t)(128 / g.res);
        {
          const uint64_t dims[4] = {(uint64_t)C, (uint64_t)g.res, (uint64_t)g.res, (uint64_t)B};
          const uint64_t str[3] = {rowb, rowb * g.res, rowb * g.L};
          const uint32_t box[4] = {64, (uint32_t)g.res, hb_xy, spt};
          dst[k][0] = make_tmap_bf16(base, 4, dims, str, box, sw, f32);
        }
        {
          const uint64_t dims[5] = {(uint64_t)C, (uint64_t)g.res, (uint64_t)g.t, 2, (uint64_t)B};
          const uint64_t str[4] = {rowb, rowb * g.res, rowb * g.res * g.t, rowb * g.L};
          const uint32_t box[5] = {64, (uint32_t)g.res, hb_pl, 1, spt};
          dst[k][1] = make_tmap_bf16(base + rowb * g.res * g.res, 5, dims, str, box, sw, f32);
        }
      }
    }
  }
  void conv_tc(const std::string& name, const ConvParams& P, int norm0, int norm1, Tensor* out_t, const TcOpts& o) {
    TcConvParams T{};
    if (out_t && fuse_gn()) { out_t->csum = alloc_csum(P.Cout, P.geo); T.csum = out_t->csum; }
    if (o.qkv) {
      T.qkv_heads = o.qkv->heads;
      T.q_hi = o.qkv->q_hi; T.q_lo = o.qkv->q_lo; T.k_hi = o.qkv->k_hi; T.k_lo = o.qkv->k_lo;
      T.vt_hi = o.qkv->vt_hi; T.vt_lo = o.qkv->vt_lo;
    }
    const KSeg& S = P.seg[0];
    T.taps = S.taps; T.Cin = S.C0 + S.C1; T.Cout = P.Cout; T.B = B; T.geo = P.geo;
    T.bias = P.bias; T.resid = P.resid; T.resid_mode = P.resid_mode; T.out = P.out;
    const int M = B * P.geo.L;
    const int mtiles = P.geo.L > 128 ? M / 128 : (B + (128 / P.geo.L) - 1) / (128 / P.geo.L);
    // BN = 128 (stacked N = 256) runs the tensor pipe ~1.6x more efficiently per FLOP than BN = 64 (a K-iteration costs ~1200
    // vs ~1000 cycles for twice the work, profiles/r01_s2_mainloop_skip.md): take it once the SMs are covered, or whenever the
    // op is deep enough for split-K to restore the CTA count
    const int iters_all = S.taps * ((S.C0 + S.C1) / 64) + (P.nsegs == 2 ? (P.seg[1].C0 + P.seg[1].C1) / 64 : 0);
    // split-K thresholds (K-iterations of 64 channels): ops with at least ks_min_total iterations are split so that every CTA
    // keeps at least ks_min_it of them.  The reduction runs inside the GEMM kernel (no second launch), so fine splits pay:
    // measured on B200 at B=1: min_it 3 -> 2.079 ms, 2 -> 2.054 ms
    static const int ks_min_total = [] { const char* e = getenv("MTV_KS_MIN_TOTAL"); return e ? std::max(2, atoi(e)) : 16; }();
    static const int min_it = [] { const char* e = getenv("MTV_KS_MIN_ITERS"); return e ? std::max(1, atoi(e)) : 2; }();
    int bn = 64;
    if (P.Cout % 128 == 0) {
      const int base128 = mtiles * (P.Cout / 128);
      const bool splittable = iters_all >= ks_min_total && ((h->tc_mask >> 5) & 1) && !o.qkv && !o.chmajor_valid && base128 * 2 <= h->num_sms + h->num_sms / 4;
      if (base128 >= 64 || (splittable && ((h->tc_mask >> 15) & 1))) bn = 128;
    }
    T.bn = bn;
    SplitBuf own0, own1;            // operands this op's own apply launches produce: dead once the GEMM is emitted
    {
      const SplitBuf a0 = o.pre0 ? *o.pre0 : emit_apply(name, S, P.geo, norm0, o.raw_out, S.w, P.Cout);
      if (!o.pre0) own0 = a0;
      make_A_maps(a0, T.Cin, S.taps, P.geo, T.tmA_hi, T.tmA_lo);
    }
    auto wmaps = [&](const KSeg& K, CUtensorMap& whi, CUtensorMap& wlo) {
      const auto& pr = h->tc_w.at(K.w);
      const int C = K.C0 + K.C1;
      const uint64_t dims[2] = {(uint64_t)C, (uint64_t)K.taps * P.Cout}; const uint64_t str[1] = {(uint64_t)C * 2};
      const uint32_t box[2] = {64, (uint32_t)bn};
      whi = make_tmap_bf16(pr.first, 2, dims, str, box);
      wlo = make_tmap_bf16(pr.second, 2, dims, str, box);
    };
    wmaps(S, T.tmW_hi, T.tmW_lo);
    double Ktot = (double)S.taps * T.Cin;
    if (P.nsegs == 2) {
      const KSeg& X = P.seg[1];
      T.Cin2 = X.C0 + X.C1;
      const SplitBuf a1 = o.pre1 ? *o.pre1 : emit_apply(name + ".skip", X, P.geo, norm1, nullptr);
      if (!o.pre1) own1 = a1;
      make_A_maps(a1, T.Cin2, 1, P.geo, T.tmA2_hi, T.tmA2_lo);
      wmaps(X, T.tmW2_hi, T.tmW2_lo);
      Ktot += T.Cin2;
    }
    const int iters = T.taps * (T.Cin / 64) + T.Cin2 / 64;
    const int base = mtiles * (P.Cout / bn);
    int ks = 1;
    if (iters >= ks_min_total && ((h->tc_mask >> 5) & 1) && !o.qkv && !o.chmajor_valid && base * 2 <= h->num_sms + h->num_sms / 4) {
      // spread a fixed amount of shared-memory / weight traffic over (nearly) all SMs; >= 3 K-iterations per CTA.  The ksplit CTAs
      // of an output tile wait for each other inside the kernel (tile ticket), so the whole grid must be co-resident: base*ks <= #SMs
      ks = std::min(iters / min_it, std::max(1, h->num_sms / base));
      ks = std::min(ks, 32);
      while (ks > 1 && (ks - 1) * ((iters + ks - 1) / ks) >= iters) --ks;
    }
    T.ksplit = ks;
    const bool coresident = base * ks <= h->num_sms;
    if (ks > 1 && !coresident) throw MtvError("internal: split-K grid is not co-resident at " + name);
    if (ks > 1) {   // split-K scratch lives only inside its own launch: ONE buffer shared by every split-K op (<= #SMs tiles of 128 x 128 fp32)
      if (!shared_partial) shared_partial = (float*)dalloc((size_t)h->num_sms * 128 * 128 * sizeof(float));
      T.partial = shared_partial;
    }
    if (ks > 1) T.sync = (unsigned long long*)dalloc((size_t)base * 32 * sizeof(unsigned long long), true);   // kernels_tc.cu: TC_SYNC_STRIDE
    Op op; op.name = "conv_tc:" + name; op.launches = 1;
    op.flops = 2.0 * M * P.Cout * Ktot;
    op.bytes = 4.0 * Ktot * P.Cout + 4.0 * M * Ktot / S.taps + 4.0 * M * P.Cout;
    if (const char* ds = getenv("MTV_TC_DBG_SKIP")) T.dbg_skip = atoi(ds);
    T.out_cvalid = o.chmajor_valid;
    if (!o.qkv && !o.chmajor_valid && ks == 1 && ((h->tc_mask >> 16) & 1)) {      // epilogue tiles leave through TMA stores (kernels_tc.cu: tma_store_2d)
      const uint64_t dims[2] = {(uint64_t)P.Cout, (uint64_t)M}; const uint64_t str[1] = {(uint64_t)P.Cout * 4};
      const uint32_t box[2] = {32, 32};
      T.tmOut = make_tmap_bf16(T.out, 2, dims, str, box, 128, true);
      T.tma_store = 1;
    }
    auto tp = std::make_shared<TcConvParams>(T);
    op.fn = [tp](cudaStream_t s) { return launch_conv_tc(*tp, s); };
    op.tc = tp;
    pl->ops.push_back(op);
    prel(own0.hi); prel(own0.lo); prel(own1.hi); prel(own1.lo);
  }
  float* shared_partial = nullptr;

  bool fuse_launches() const { return h->cfg.kernel_path != 1 && ((h->tc_mask >> 9) & 1); }
  void conv(const std::string& name, ConvParams P, int norm0 = -1, int norm1 = -1, Tensor* out_t = nullptr,
            int phase = 1, bool out_is_ctx = false, const TcOpts& opts = TcOpts()) {
    P.B = B;
    if (phase == 1 && !out_is_ctx && tc_ok(P)) { conv_tc(name, P, norm0, norm1, out_t, opts); return; }
    if (opts.pre0 || opts.pre1 || opts.qkv) throw MtvError("internal: tensor-core-only options on a CUDA-core op: " + name);
    const int nid[2] = {norm0, norm1};
    for (int s = 0; s < P.nsegs; ++s)
      if (nid[s] >= 0) {
        const NormRef r = materialize(nid[s]);
        P.seg[s].nrm_a = r.a; P.seg[s].nrm_d = r.d; P.seg[s].nrm_nseg = r.nseg;
      }
    double K = 0;
    for (int s = 0; s < P.nsegs; ++s) {
      const int Ct = P.seg[s].C0 + P.seg[s].C1;
      if (Ct % 16 || P.seg[s].C0 % 16) throw MtvError("tap-GEMM needs channels % 16 == 0 at " + name);
      K += (double)P.seg[s].taps * Ct;
    }
    if (P.Cout % 4) throw MtvError("tap-GEMM needs Cout % 4 == 0 at " + name);
    const double M = (double)B * P.geo.L;
    P.ksplit = conv_simt_pick_ksplit(P, h->num_sms);
    if (P.ksplit > 1) P.partial = (float*)dalloc((size_t)P.ksplit * (size_t)M * P.Cout * sizeof(float));
    Op op; op.name = "conv:" + name; op.phase = phase; op.launches = P.ksplit > 1 ? 2 : 1;
    op.flops = 2.0 * M * P.Cout * K;
    op.bytes = 4.0 * (K * P.Cout + M * K / (P.seg[0].taps) + M * P.Cout);
    Plan* plan = pl;
    if (out_is_ctx) op.fn = [P, plan](cudaStream_t s) { ConvParams Q = P; Q.out = plan->ctx.out; return launch_conv_simt(Q, s); };
    else            op.fn = [P](cudaStream_t s) { return launch_conv_simt(P, s); };
    pl->ops.push_back(op);
  }

  float* film_buf = nullptr;

  Tensor res_block(const ResDesc& r, const Tensor& x0, const Tensor* x1, int level_in, int level_out) {
    const std::string& p = r.name;
    const int cin = x0.C + (x1 ? x1->C : 0);
    if (cin != r.cin) throw MtvError("internal: channel mismatch at " + p);
    const int n1 = gn(p + ".in_layers.0", x0, x1, false, h->W(p + ".in_layers.0.weight"), h->W(p + ".in_layers.0.bias"), -1);
    Tensor hmid = T(r.cout, level_out);
    SplitBuf skip_raw; bool have_raw = false;
    {
      ConvParams P{}; P.nsegs = 1; P.geo = geo(level_out); P.Cout = r.cout;
      KSeg& S = P.seg[0];
      S.src0 = x0.p; S.C0 = x0.C; S.src1 = x1 ? x1->p : nullptr; S.C1 = x1 ? x1->C : 0;
      S.silu = 1; S.resample = r.updown; S.taps = 9;
      S.w = h->W(p + ".in_layers.2.weight");
      P.bias = h->W(p + ".in_layers.2.bias"); P.out = hmid.p;
      // when both convs run on the tensor cores, conv1's apply kernel also emits the raw split of x that
      // conv2's fused 1x1 skip segment consumes (one launch instead of two)
      if (r.cin != r.cout && fuse_launches() && tc_ok_pb(P)) {
        ConvParams P2{}; P2.nsegs = 2; P2.geo = geo(level_out); P2.Cout = r.cout;
        P2.seg[0].C0 = r.cout; P2.seg[0].taps = 9; P2.seg[0].w = h->W(p + ".out_layers.3.weight");
        P2.seg[1].C0 = x0.C; P2.seg[1].C1 = x1 ? x1->C : 0; P2.seg[1].taps = 1; P2.seg[1].w = h->W(p + ".skip_connection.weight");
        if (tc_ok_pb(P2)) { TcOpts o; o.raw_out = &skip_raw; conv(p + ".in_layers.2", P, n1, -1, &hmid, 1, false, o); have_raw = true; }
      }
      if (!have_raw) conv(p + ".in_layers.2", P, n1, -1, &hmid);
    }
    const int n2 = gn(p + ".out_layers.0", hmid, nullptr, false, h->W(p + ".out_layers.0.weight"),
                      h->W(p + ".out_layers.0.bias"), r.film_off);
    Tensor out = T(r.cout, level_out);
    {
      ConvParams P{}; P.nsegs = 1; P.geo = geo(level_out); P.Cout = r.cout;
      KSeg& S = P.seg[0];
      S.src0 = hmid.p; S.C0 = r.cout; S.silu = 1;
      S.resample = RS_NONE; S.taps = 9; S.w = h->W(p + ".out_layers.3.weight");
      if (r.cin != r.cout) {   // 1x1 skip conv folded in as a second K-segment (unet.py:167, 207)
        P.nsegs = 2; KSeg& K1 = P.seg[1];
        K1.src0 = x0.p; K1.C0 = x0.C; K1.src1 = x1 ? x1->p : nullptr; K1.C1 = x1 ? x1->C : 0;
        K1.resample = r.updown; K1.taps = 1; K1.w = h->W(p + ".skip_connection.weight");
        P.bias = h->bias_sum.at(p);
      } else {
        if (x1) throw MtvError("identity skip over a concatenated input is not on the MToV path: " + p);
        P.bias = h->W(p + ".out_layers.3.bias");
        P.resid = x0.p; P.resid_mode = r.updown;
      }
      P.out = out.p;
      TcOpts o; if (have_raw) o.pre1 = &skip_raw;
      conv(p + ".out_layers.3", P, n2, -1, &out, 1, false, o);
    }
    Trel(hmid); prel(skip_raw.hi); prel(skip_raw.lo);
    return out;
  }

  Tensor attn_block(const AttnDesc& a, const Tensor& x, int level) {
    const std::string& p = a.name;
    const int C = a.C, heads = h->cfg.num_heads;
    if (x.C != C) throw MtvError("internal: channel mismatch at " + p);
    const int D = C / heads;
    if (C % heads || !(D == 16 || D == 32 || D == 64 || D == 128))
      throw MtvError("attention head dim must be 16/32/64/128 at " + p);
    const int L = geo(level).L;
    const int n = gn(p + ".norm", x, nullptr, a.joint, h->W(p + ".norm.weight"), h->W(p + ".norm.bias"), -1);

    ConvParams Pq{}; Pq.nsegs = 1; Pq.geo = geo(level); Pq.Cout = 3 * C;
    { KSeg& S = Pq.seg[0]; S.src0 = x.p; S.C0 = C; S.silu = 0; S.taps = 1; S.w = h->W(p + ".qkv.weight"); }
    Pq.bias = h->W(p + ".qkv.bias");
    ConvParams Pp{}; Pp.nsegs = 1; Pp.geo = geo(level); Pp.Cout = C;
    { KSeg& S = Pp.seg[0]; S.C0 = C; S.taps = 1; S.w = h->W(p + ".proj_out.weight"); }
    Pp.bias = h->W(p + ".proj_out.bias"); Pp.resid = x.p; Pp.resid_mode = RS_NONE;

    const bool tc_attn = h->cfg.kernel_path != 1 && ((h->tc_mask >> 6) & 1) && (D == 16 || D == 32 || D == 64);
    const bool fuse = tc_attn && fuse_launches() && tc_ok_pb(Pq) && tc_ok_pb(Pp);

    AttnParams A{}; A.B = B; A.L = L; A.C = C; A.heads = heads;
    set_segs(level, a.joint, A.nseg, A.seg_off);
    double pairs = 0;
    for (int i = 0; i < A.nseg; ++i) { const double l = A.seg_off[i + 1] - A.seg_off[i]; pairs += l * l; }

    QkvSplitParams Q{};
    if (tc_attn) {
      const size_t bytes = (size_t)B * L * C * 2;
      Q.B = B; Q.L = L; Q.C = C; Q.heads = heads;
      Q.q_hi = palloc(bytes); Q.q_lo = palloc(bytes); Q.k_hi = palloc(bytes); Q.k_lo = palloc(bytes);
      Q.vt_hi = palloc(bytes); Q.vt_lo = palloc(bytes);
    }
    // ---- qkv projection
    Tensor qkv;
    if (fuse) {             // the GEMM epilogue writes Q / K / V^T directly; no fp32 qkv tensor
      TcOpts o; o.qkv = &Q;
      conv(p + ".qkv", Pq, n, -1, nullptr, 1, false, o);
    } else {
      qkv = T(3 * C, level);
      Pq.out = qkv.p;
      conv(p + ".qkv", Pq, n);
      if (tc_attn) {
        Q.qkv = qkv.p;
        Op op; op.name = "qkv_split:" + p; op.bytes = (double)B * L * C * 24;
        op.fn = [Q](cudaStream_t s) { return launch_qkv_split(Q, s); }; pl->ops.push_back(op);
      }
    }
    // ---- attention core
    Tensor att; SplitBuf att_split;
    if (tc_attn) {
      AttnTcParams T{}; T.B = B; T.L = L; T.C = C; T.heads = heads; T.nseg = A.nseg;
      if (const char* ds = getenv("MTV_ATTN_DBG_SKIP")) T.dbg_skip = atoi(ds);
      for (int i = 0; i < 4; ++i) T.seg_off[i] = A.seg_off[i];
      if (fuse) {
        const size_t bytes = (size_t)B * L * C * 2;
        att_split.hi = palloc(bytes); att_split.lo = palloc(bytes);
        T.out_hi = att_split.hi; T.out_lo = att_split.lo;
        if ((h->tc_mask >> 10) & 1) {
          auto it = h->tc_w.find(Pp.seg[0].w);
          if (it != h->tc_w.end()) { T.pf0 = it->second.first; T.pf1 = it->second.second; T.pf_bytes = (unsigned long long)C * C * 2; }
        }
      } else {
        att = this->T(C, level); T.out = att.p;
      }
      const uint64_t rows = (uint64_t)B * heads * L;
      {
        const uint64_t dims[2] = {(uint64_t)D, rows}; const uint64_t str[1] = {(uint64_t)D * 2};
        const uint32_t bq[2] = {(uint32_t)D, 128}, bk[2] = {(uint32_t)D, 64};
        T.tmQ_hi = make_tmap_bf16(Q.q_hi, 2, dims, str, bq, 2 * D); T.tmQ_lo = make_tmap_bf16(Q.q_lo, 2, dims, str, bq, 2 * D);
        T.tmK_hi = make_tmap_bf16(Q.k_hi, 2, dims, str, bk, 2 * D); T.tmK_lo = make_tmap_bf16(Q.k_lo, 2, dims, str, bk, 2 * D);
      }
      {
        const uint64_t dims[2] = {(uint64_t)L, (uint64_t)B * heads * D}; const uint64_t str[1] = {(uint64_t)L * 2};
        const uint32_t bv[2] = {64, (uint32_t)D};
        T.tmV_hi = make_tmap_bf16(Q.vt_hi, 2, dims, str, bv, 128); T.tmV_lo = make_tmap_bf16(Q.vt_lo, 2, dims, str, bv, 128);
      }
      {
        // split-KV: when the launch has fewer CTAs than SMs (one latency-bound CTA per SM), 2 CTAs share each query tile
        int nqb = 0, min_blk = 1 << 30;
        for (int i = 0; i < A.nseg; ++i) {
          const int len = A.seg_off[i + 1] - A.seg_off[i];
          nqb += (len + 127) / 128; min_blk = std::min(min_blk, (len + 63) / 64);
        }
        static const int kvs_max = [] { const char* e = getenv("MTV_ATTN_KV_SPLIT"); return e ? std::max(1, atoi(e)) : 2; }();   // 4 measured no better than 2 (profiles/r02_attention.md)
        int kvs = 1;
        const int cap = D == 64 ? 2 : 4;       // rank 0's merge buffer (kvs - 1) x 128 x (D + 2) floats must fit next to the K / V ring
        while (kvs * 2 <= std::min(kvs_max, cap) && nqb * B * heads * kvs < h->num_sms && min_blk >= 2 * kvs * 2) kvs *= 2;   // >= 2 key blocks per rank
        T.kv_split = kvs;
      }
      Op op; op.name = "attn_tc:" + p;
      op.flops = 4.0 * B * heads * pairs * D; op.bytes = 4.0 * B * L * 4 * C;
      op.fn = [T](cudaStream_t s) { return launch_attn_tc(T, s); };
      pl->ops.push_back(op);
    } else {
      att = this->T(C, level);
      A.qkv = qkv.p; A.out = att.p;
      Op op; op.name = "attn:" + p;
      op.flops = 4.0 * B * heads * pairs * D; op.bytes = 4.0 * B * L * 4 * C;
      op.fn = [A](cudaStream_t s) { return launch_attn_simt(A, s); };
      pl->ops.push_back(op);
    }
    // ---- output projection + residual
    Tensor out = this->T(C, level);
    Pp.out = out.p;
    if (fuse) {
      TcOpts o; o.pre0 = &att_split;
      conv(p + ".proj_out", Pp, -1, -1, &out, 1, false, o);
    } else {
      Pp.seg[0].src0 = att.p;
      conv(p + ".proj_out", Pp, -1, -1, &out);
    }
    // everything between the block input and its output is dead now
    prel(Q.q_hi); prel(Q.q_lo); prel(Q.k_hi); prel(Q.k_lo); prel(Q.vt_hi); prel(Q.vt_lo);
    prel(att_split.hi); prel(att_split.lo);
    if (qkv.p) Trel(qkv);
    if (att.p) Trel(att);
    return out;
  }

  Tensor run_stage(const StageDesc& st, Tensor cur, const Tensor* skip) {
    int level = st.level_in;
    bool first = true;
    // the stage input and the skip tensor are stage outputs (kept: taps / skip stack); tensors between layers die as soon as
    // the next layer has been emitted
    for (const Layer& l : st.layers) {
      const Tensor prev = cur;
      if (l.is_res) {
        const int lo = (l.r.updown == RS_NONE) ? level : (l.r.updown == RS_DOWN2 ? level + 1 : level - 1);
        cur = res_block(l.r, cur, (first && skip) ? skip : nullptr, level, lo);
        level = lo;
      } else {
        cur = attn_block(l.a, cur, level);
      }
      if (!first) Trel(prev);
      first = false;
    }
    if (st.has_joint) { const Tensor prev = cur; cur = attn_block(st.joint, cur, level); if (!first) Trel(prev); }
    return cur;
  }

  void build() {
    const MtvConfig& c = h->cfg; const Arch& A = h->arch;
    const int mc = c.model_channels, ted = 4 * mc;
    Plan* plan = pl;
    // ---- caller inputs -> internal buffers (outside the graph)
    int64_t* t_buf = (int64_t*)dalloc((size_t)B * sizeof(int64_t));
    {
      Op op; op.name = "copy_t"; op.phase = 0; op.launches = 0;
      const int Bc = B;
      op.fn = [plan, t_buf, Bc](cudaStream_t s) {
        return cudaMemcpyAsync(t_buf, plan->ctx.t, (size_t)Bc * sizeof(int64_t), cudaMemcpyDeviceToDevice, s);
      };
      pl->ops.push_back(op);
    }
    // tensor-core stem (unet.py:714): the packed input is written directly as the split-bf16 operand, channels padded to 64
    const bool tc_stem = c.kernel_path != 1 && (h->tc_mask & 1) && ((h->tc_mask >> 18) & 1) && mc % 64 == 0 && 4 * c.in_channels < 64 &&
                         h->tc_w.count(h->W("input_blocks.0.0.weight"));
    Tensor xin; SplitBuf xin_split;
    if (tc_stem) {
      const size_t bytes = (size_t)B * geo(0).L * 64 * 2;
      xin_split.hi = dalloc(bytes); xin_split.lo = dalloc(bytes);
    } else {
      xin = T(4 * c.in_channels, 0);
    }
    {
      PackParams P{}; P.B = B; P.cx = c.in_channels; P.cc = 2 * c.in_channels; P.ci = c.in_channels; P.out = xin.p;
      P.hi = xin_split.hi; P.lo = xin_split.lo; P.cpad = 64;
      Op op; op.name = "pack_in"; op.phase = 0;
      op.fn = [plan, P](cudaStream_t s) {
        PackParams Q = P; Q.x = plan->ctx.x; Q.cond = plan->ctx.cond; Q.image_cond = plan->ctx.image_cond;
        Q.ic_len = plan->ctx.ic_len; return launch_pack_in(Q, s);
      };
      pl->ops.push_back(op);
    }
    // ---- timestep embedding + every ResBlock's FiLM scale/shift
    // per-channel statistics arena: every slot is rewritten by its producer in every forward (no zeroing, no atomics)
    pl->csum_cap = (size_t)B * (8u << 20);
    pl->csum_arena = (char*)dalloc(pl->csum_cap, true);
    film_buf = ((h->tc_mask >> 11) & 1) ? h->small_alloc((size_t)B * A.J * sizeof(float)) : nullptr;
    if (!film_buf) film_buf = (float*)dalloc((size_t)B * A.J * sizeof(float));
    {
      EmbParams P{}; P.t = t_buf; P.B = B; P.mc = mc; P.ted = ted; P.freqs = h->freqs;
      P.w1 = h->W("time_embed.0.weight"); P.b1 = h->W("time_embed.0.bias");
      P.w2 = h->W("time_embed.2.weight"); P.b2 = h->W("time_embed.2.bias");
      P.wall = h->emb_wall; P.ball = h->emb_ball; P.J = A.J;
      P.temb = (float*)dalloc((size_t)B * mc * sizeof(float));
      P.h1 = (float*)dalloc((size_t)B * ted * sizeof(float));
      P.semb = (float*)dalloc((size_t)B * ted * sizeof(float));
      P.film = film_buf;
      Op op; op.name = "emb"; op.launches = 4;
      op.flops = 2.0 * B * ((double)ted * mc + (double)ted * ted + (double)A.J * ted);
      op.bytes = 4.0 * ((double)ted * mc + (double)ted * ted + (double)A.J * ted);
      op.fn = [P](cudaStream_t s) { return launch_emb(P, s); };
      pl->ops.push_back(op);
    }
    // ---- encoder
    std::vector<Tensor> skips;
    Tensor cur;
    {
      cur = T(mc, 0);
      ConvParams P{}; P.nsegs = 1; P.geo = geo(0); P.Cout = mc;
      KSeg& S = P.seg[0]; S.src0 = xin.p; S.C0 = tc_stem ? 64 : xin.C; S.taps = 9; S.w = h->W("input_blocks.0.0.weight");
      P.bias = h->W("input_blocks.0.0.bias"); P.out = cur.p;
      if (tc_stem) { P.B = B; TcOpts o; o.pre0 = &xin_split; conv_tc("input_blocks.0.0", P, -1, -1, &cur, o); }
      else conv("input_blocks.0.0", P, -1, -1, &cur);
      pl->taps["in0"] = cur; skips.push_back(cur); keep.insert(cur.p);
    }
    for (size_t i = 1; i < A.in.size(); ++i) {
      cur = run_stage(A.in[i], cur, nullptr);
      cur.level = A.in[i].level_out;
      pl->taps["in" + std::to_string(i)] = cur; skips.push_back(cur); keep.insert(cur.p);
    }
    cur = run_stage(A.mid, cur, nullptr);
    pl->taps["mid"] = cur; keep.insert(cur.p);
    for (size_t i = 0; i < A.out.size(); ++i) {
      Tensor sk = skips.back(); skips.pop_back();
      if (sk.C != A.out[i].skip_ch) throw MtvError("internal: skip stack mismatch");
      cur = run_stage(A.out[i], cur, &sk);
      cur.level = A.out[i].level_out;
      pl->taps["out" + std::to_string(i)] = cur; keep.insert(cur.p);
    }
    // ---- head: GN -> SiLU -> conv3x3 -> channel-major eps (unet.py:971-975, 1103-1112)
    const int nh = gn("out.0", cur, nullptr, false, h->W("out.0.weight"), h->W("out.0.bias"), -1);
    float* eps_buf = (float*)dalloc((size_t)B * c.out_channels * geo(0).L * sizeof(float));
    const bool tc_head = c.kernel_path != 1 && (h->tc_mask & 1) && ((h->tc_mask >> 18) & 1) && cur.C % 64 == 0 && c.out_channels <= 32 &&
                         h->head_bias_pad && h->tc_w.count(h->W("out.2.weight")) && cur.csum;
    if (tc_head) {   // tensor-core head: Cout padded 4 -> 64, channel-major epilogue; GroupNorm + SiLU from the producer's channel sums
      ConvParams P{}; P.nsegs = 1; P.geo = geo(0); P.Cout = 64; P.B = B;
      KSeg& S = P.seg[0]; S.src0 = cur.p; S.C0 = cur.C; S.silu = 1; S.taps = 9; S.w = h->W("out.2.weight");
      P.bias = h->head_bias_pad; P.out = eps_buf;
      TcOpts o; o.chmajor_valid = c.out_channels;
      conv_tc("out.2", P, nh, -1, nullptr, o);
    } else {
      ConvParams P{}; P.nsegs = 1; P.geo = geo(0); P.Cout = c.out_channels;
      KSeg& S = P.seg[0]; S.src0 = cur.p; S.C0 = cur.C; S.silu = 1;
      S.taps = 9; S.w = h->W("out.2.weight");
      P.bias = h->W("out.2.bias"); P.out = eps_buf; P.out_chmajor = 1;
      conv("out.2", P, nh);
    }
    {
      Op op; op.name = "copy_out"; op.phase = 2; op.launches = 0;
      const size_t bytes = (size_t)B * c.out_channels * geo(0).L * sizeof(float);
      op.fn = [plan, eps_buf, bytes](cudaStream_t s) {
        return cudaMemcpyAsync(plan->ctx.out, eps_buf, bytes, cudaMemcpyDeviceToDevice, s);
      };
      pl->ops.push_back(op);
    }
  }
};

// Plans are cached per batch size, at most MTV_MAX_PLANS of them (least recently used is dropped: a caller whose last
// chunk / batch size varies does not accumulate workspace without bound).  Building a plan allocates its workspace and
// synchronises the device once (documented in include/mtv_b200.h); steady-state calls never synchronise.
constexpr size_t MTV_MAX_PLANS = 4;
Plan* get_plan(MtvHandle_t* h, int B) {
  auto it = h->plans.find(B);
  if (it != h->plans.end()) { it->second->last_use = ++h->use_clock; return it->second.get(); }
  if (h->plans.size() >= MTV_MAX_PLANS) {
    auto victim = h->plans.begin();
    for (auto jt = h->plans.begin(); jt != h->plans.end(); ++jt)
      if (jt->second->last_use < victim->second->last_use) victim = jt;
    CK(cudaDeviceSynchronize());            // its launches may still be in flight
    if (h->last_plan == victim->second.get()) h->last_plan = nullptr;
    h->plans.erase(victim);
  }
  std::unique_ptr<Plan> pl(new Plan());
  pl->last_use = ++h->use_clock;
  pl->B = B;
  Builder b(h, pl.get());
  b.build();
  CK(cudaDeviceSynchronize());   // memsets of the statistics scratch
  Plan* raw = pl.get();
  h->plans[B] = std::move(pl);
  return raw;
}

void run_ops(Plan* pl, int phase, cudaStream_t s) {
  for (Op& op : pl->ops)
    if (op.phase == phase) {
      cudaError_t e = op.fn(s);
      if (e != cudaSuccess) throw MtvError("launch failed at " + op.name + ": " + cudaGetErrorString(e));
    }
}

void forward(MtvHandle_t* h, const RunCtx& ctx, int B, cudaStream_t s) {
  DeviceGuard dg(h->cfg.device);
  ensure_ready(h, s);
  Plan* pl = get_plan(h, B);
  pl->ctx = ctx;
  run_ops(pl, 0, s);
  if (!h->use_graph) {
    run_ops(pl, 1, s);
  } else {
    if (!pl->exec) {
      if (pl->runs == 0) {
        run_ops(pl, 1, s);     // first call runs eagerly (function attributes, lazy module load)
      } else {
        // capture on a private stream: the caller's stream may be the legacy default stream,
        // which cannot be captured
        cudaStream_t cs = h->capture_stream();
        h->apply_l2_window(cs);      // kernel nodes inherit the capturing stream's access-policy window
        CK(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
        try { run_ops(pl, 1, cs); } catch (...) { cudaGraph_t g = nullptr; cudaStreamEndCapture(cs, &g); if (g) cudaGraphDestroy(g); throw; }
        CK(cudaStreamEndCapture(cs, &pl->graph));
        CK(cudaGraphInstantiate(&pl->exec, pl->graph, 0));
        CK(cudaGraphLaunch(pl->exec, s));
      }
    } else {
      CK(cudaGraphLaunch(pl->exec, s));
    }
  }
  run_ops(pl, 2, s);
  pl->runs++;
  h->last_plan = pl;
}

template <typename F>
int guarded(F&& f) {
  try { f(); return 0; }
  catch (const std::exception& e) { g_err = e.what(); return 1; }
  catch (...) { g_err = "unknown error"; return 1; }
}

std::string strip_prefix(const char* name) {
  std::string n(name);
  const std::string pre = "diffusion_model.";
  if (n.compare(0, pre.size(), pre) == 0) n = n.substr(pre.size());
  return n;
}

}  // namespace

// ====================================================================== C ABI
extern "C" {

int mtv_abi_version(void) { return MTV_ABI_VERSION; }
const char* mtv_last_error(void) { return g_err.c_str(); }

int mtv_create(const MtvConfig* cfg, MtvHandle* out) {
  return guarded([&] {
    if (!cfg || !out) throw MtvError("null argument");
    if (cfg->abi_version != MTV_ABI_VERSION) throw MtvError("MtvConfig.abi_version mismatch");
    if (cfg->num_levels < 1 || cfg->num_levels > MTV_MAX_LEVELS) throw MtvError("num_levels out of range");
    if (cfg->image_size != 32) throw MtvError("image_size must be 32 (the tri-plane latent is [B,4,2048])");
    if (cfg->model_channels % 32) throw MtvError("model_channels must be a multiple of 32");
    if ((cfg->image_size / 2) >> (cfg->num_levels - 1) < 1) throw MtvError("too many levels for a 16-row plane");
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (cfg->device < 0 || cfg->device >= ndev) throw MtvError("no such CUDA device");
    DeviceGuard dg(cfg->device);
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10)
      throw MtvError(std::string("libmtv_b200 is built for sm_100a only; device is sm_") + std::to_string(prop.major) +
                     std::to_string(prop.minor));
    std::unique_ptr<MtvHandle_t> h(new MtvHandle_t());
    h->cfg = *cfg; h->num_sms = prop.multiProcessorCount;
    h->arch = build_arch(*cfg);
    const char* ng = getenv("MTV_NO_GRAPH");
    h->use_graph = !(ng && ng[0] == '1');
    if (const char* tm = getenv("MTV_TC_MASK")) h->tc_mask = (int)strtol(tm, nullptr, 0);
    // programmatic dependent launch by kernel class (mtv_kernels.cuh): ONE process-wide mask, default 7 (tap-GEMM + attention +
    // apply; measured on B200, B=1: 5 -> 2.120 ms, 7 / 15 / 31 -> 2.078 ms).  An MTV_PDL set in the environment at create time
    // overrides it for the whole process (diagnostics / tests); handles do not carry their own mask.
    if (const char* np = getenv("MTV_PDL")) g_mtv_use_pdl = atoi(np);
    register_weights(h.get());
    *out = h.release();
  });
}

int mtv_destroy(MtvHandle h) {
  return guarded([&] {
    if (!h) return;
    DeviceGuard dg(h->cfg.device);
    cudaDeviceSynchronize();
    h->plans.clear();
    h->release_l2_window();
    if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
    for (void* p : h->allocs) cudaFree(p);
    delete h;
  });
}

int mtv_load_weight(MtvHandle h, const char* name, const float* data, const int64_t* shape, int32_t ndim,
                    int32_t* used, void* stream) {
  return guarded([&] {
    if (!h || !name || !data) throw MtvError("null argument");
    cudaStream_t s = (cudaStream_t)stream;
    DeviceGuard dg(h->cfg.device);
    const std::string n = strip_prefix(name);
    auto it = h->windex.find(n);
    if (it == h->windex.end()) {
      const bool dead = n.compare(0, 10, "output_bg_") == 0 || n == "zeros";
      if (!dead) throw MtvError("unexpected key in state dict: " + n);
      if (used) *used = 0;
      return;
    }
    Weight& w = h->weights[it->second];
    if ((int)w.shape.size() != ndim) throw MtvError("rank mismatch for " + n);
    for (int i = 0; i < ndim; ++i)
      if (w.shape[i] != shape[i]) throw MtvError("size mismatch for " + n);
    if (w.kind == WK_CONV) {
      const int Cout = (int)w.shape[0], Cin = (int)w.shape[1];
      const int taps = (int)(w.elems / ((size_t)Cout * Cin));
      CK(launch_repack_conv(data, w.dev, Cout, Cin, taps, s));
      if (w.hi && (w.pad_cin || w.pad_cout))
        CK(launch_repack_split_w_pad(data, w.hi, w.lo, Cout, Cin, w.pad_cout ? w.pad_cout : Cout, w.pad_cin ? w.pad_cin : Cin, taps, s));
      else if (w.hi) CK(launch_repack_split_w(data, w.hi, w.lo, Cout, Cin, taps, s));
    } else {
      CK(cudaMemcpyAsync(w.dev, data, w.elems * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
    w.loaded = true; h->dirty = true;
    if (used) *used = 1;
  });
}

int mtv_weights_ready(MtvHandle h, int32_t* needed, int32_t* missing) {
  return guarded([&] {
    if (!h) throw MtvError("null handle");
    int miss = 0; std::string first;
    for (const Weight& w : h->weights)
      if (!w.loaded) { if (!miss) first = w.name; ++miss; }
    if (needed) *needed = (int)h->weights.size();
    if (missing) *missing = miss;
    if (miss) throw MtvError("missing key in state dict: " + first);
  });
}

int mtv_num_weight_names(MtvHandle h) { return h ? (int)h->weights.size() : 0; }
const char* mtv_weight_name(MtvHandle h, int32_t i) {
  if (!h || i < 0 || i >= (int)h->weights.size()) return nullptr;
  return h->weights[i].name.c_str();
}

int mtv_unet_forward(MtvHandle h, const float* x, const float* cond, const float* image_cond, int64_t image_cond_len,
                     const int64_t* t, int32_t B, float* out, void* stream) {
  return guarded([&] {
    if (!h || !x || !cond || !image_cond || !t || !out) throw MtvError("null argument");
    if (B < 1) throw MtvError("batch must be >= 1");
    if (image_cond_len < 1024) throw MtvError("image_cond needs at least the 1024-token xy plane");
    RunCtx ctx; ctx.x = x; ctx.cond = cond; ctx.image_cond = image_cond; ctx.ic_len = image_cond_len; ctx.t = t; ctx.out = out;
    forward(h, ctx, B, (cudaStream_t)stream);
  });
}

int mtv_ddim_step(MtvHandle h, float* img, const float* eps, const float* noise, int64_t n, float sr, float srm1,
                  float san, float c, float sigma, int32_t last, void* stream) {
  return guarded([&] {
    if (!h || !img || !eps || (!last && !noise)) throw MtvError("null argument");
    DeviceGuard dg(h->cfg.device);
    CK(launch_ddim_step(img, eps, noise, n, sr, srm1, san, c, sigma, last, (cudaStream_t)stream));
  });
}

int mtv_q_sample(MtvHandle h, const float* x_start, const float* noise, int64_t n, float a, float b, float* out,
                 void* stream) {
  return guarded([&] {
    if (!h || !x_start || !noise || !out) throw MtvError("null argument");
    DeviceGuard dg(h->cfg.device);
    CK(launch_q_sample(x_start, noise, n, a, b, out, (cudaStream_t)stream));
  });
}

int mtv_io_prep_frames(int32_t device, const uint8_t* frames, int32_t T, int32_t H, int32_t W, const int32_t* mask_row,
                       int32_t R, float* out, void* stream) {
  return guarded([&] {
    if (!frames || !out) throw MtvError("null argument");
    if (T < 1 || H < 1 || W < 1 || R < 4 || (R & 3)) throw MtvError("mtv_io_prep_frames: need T, H, W >= 1 and a resolution that is a multiple of 4");
    DeviceGuard dg(device);
    CK(launch_io_prep_frames(frames, T, H, W, mask_row, R, out, (cudaStream_t)stream));
  });
}

int mtv_io_prep_frames_ex(int32_t device, const uint8_t* frames, int32_t T, int32_t H, int32_t W, const int32_t* mask_row,
                          int32_t R, int32_t flags, float* out, void* stream) {
  return guarded([&] {
    if (!frames || !out) throw MtvError("null argument");
    if (T < 1 || H < 1 || W < 1 || R < 4 || (R & 3)) throw MtvError("mtv_io_prep_frames_ex: need T, H, W >= 1 and a resolution that is a multiple of 4");
    if (flags & ~1) throw MtvError("mtv_io_prep_frames_ex: unknown flag bits");
    DeviceGuard dg(device);
    CK(launch_io_prep_frames(frames, T, H, W, mask_row, R, out, (cudaStream_t)stream, (flags & 1) != 0));
  });
}

int mtv_io_rasterize_landmarks(int32_t device, const void* landmarks, int32_t is_f64, int32_t T, int32_t N, int32_t dims, int32_t WH,
                               int32_t flip, float* out, void* stream) {
  return guarded([&] {
    if (!out || (!landmarks && T * N > 0)) throw MtvError("null argument");
    if (T < 1 || N < 0 || (dims != 2 && dims != 3) || WH < 1) throw MtvError("mtv_io_rasterize_landmarks: need T >= 1, dims 2 or 3, WH >= 1");
    DeviceGuard dg(device);
    CK(launch_io_rasterize(landmarks, is_f64, T, N, dims, WH, flip, out, (cudaStream_t)stream));
  });
}

int mtv_io_frames_out(int32_t device, const float* dec, int32_t B, int32_t T, int32_t H, int32_t W, uint8_t* frames_u8, uint8_t* last_u8,
                      float* next_ref, int32_t Trep, void* stream) {
  return guarded([&] {
    if (!dec) throw MtvError("null argument");
    if (B < 1 || T < 1 || H < 1 || W < 4 || (W & 3)) throw MtvError("mtv_io_frames_out: need B, T, H >= 1 and a width that is a multiple of 4");
    if (next_ref && Trep < 1) throw MtvError("mtv_io_frames_out: Trep must be >= 1 when next_ref is requested");
    DeviceGuard dg(device);
    CK(launch_io_frames_out(dec, B, T, H, W, frames_u8, last_u8, next_ref, Trep, (cudaStream_t)stream));
  });
}

int mtv_plan_info(MtvHandle h, int32_t B, int64_t* n_launches, int64_t* workspace_bytes, int64_t* weight_bytes) {
  return guarded([&] {
    if (!h) throw MtvError("null handle");
    DeviceGuard dg(h->cfg.device);
    Plan* pl = get_plan(h, B);
    int64_t n = 0;
    for (const Op& op : pl->ops) n += op.launches;
    if (n_launches) *n_launches = n;
    if (workspace_bytes) *workspace_bytes = (int64_t)pl->alloc_bytes;
    if (weight_bytes) *weight_bytes = h->weight_bytes;
  });
}

int mtv_debug_read(MtvHandle h, const char* tag, float* dst, int64_t dst_elems, void* stream) {
  return guarded([&] {
    if (!h || !tag || !dst) throw MtvError("null argument");
    DeviceGuard dg(h->cfg.device);
    Plan* pl = h->last_plan;
    if (!pl) throw MtvError("no forward has run yet");
    auto it = pl->taps.find(tag);
    if (it == pl->taps.end()) throw MtvError(std::string("unknown tap ") + tag);
    const Tensor& t = it->second;
    const Geo g = level_geo(h->cfg, t.level);
    if ((int64_t)pl->B * g.L * t.C != dst_elems) throw MtvError("tap size mismatch");
    CK(launch_tok2ch(t.p, dst, pl->B, g.L, t.C, (cudaStream_t)stream));
  });
}

int mtv_debug_tc_timing(MtvHandle h, int64_t* records, int32_t cap, int32_t* count) {
  return guarded([&] {
    if (!h) throw MtvError("null handle");
    DeviceGuard dg(h->cfg.device);
    if (count) { unsigned int n = 0; CK(tc_debug_count(&n)); *count = (int32_t)n; }
    CK(tc_debug_arm((long long*)records, records ? (unsigned int)cap : 0u));
  });
}

int mtv_profile_forward(MtvHandle h, const float* x, const float* cond, const float* image_cond, int64_t image_cond_len,
                        const int64_t* t, int32_t B, float* out, MtvKernelTime* entries, int32_t cap, int32_t* n,
                        void* stream) {
  return guarded([&] {
    if (!h || !entries || !n) throw MtvError("null argument");
    cudaStream_t s = (cudaStream_t)stream;
    DeviceGuard dg(h->cfg.device);
    ensure_ready(h, s);
    Plan* pl = get_plan(h, B);
    pl->ctx.x = x; pl->ctx.cond = cond; pl->ctx.image_cond = image_cond; pl->ctx.ic_len = image_cond_len;
    pl->ctx.t = t; pl->ctx.out = out;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    // each op is launched `reps` times back to back between the two events (ops are idempotent):
    // the per-launch average then carries the steady-state launch gap, not the event overhead
    int reps = 1;
    if (const char* rp = getenv("MTV_PROFILE_REPS")) reps = std::max(1, atoi(rp));
    // Graph-body ops are timed as a private CUDA graph holding `reps` copies of the launch, so the
    // host's per-launch cost (several microseconds, more than many of these kernels run) stays out of
    // the measurement; ops that read caller pointers are timed eagerly.
    const char* pg = getenv("MTV_PROFILE_EAGER");
    const bool use_graph = !(pg && pg[0] == '1');
    int k = 0;
    for (Op& op : pl->ops) {
      float ms = 0;
      if (use_graph && op.phase == 1) {
        cudaStream_t cs = h->capture_stream();
        cudaGraph_t g = nullptr; cudaGraphExec_t ge = nullptr;
        CK(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
        cudaError_t e = cudaSuccess;
        for (int r = 0; r < reps && e == cudaSuccess; ++r) e = op.fn(cs);
        cudaError_t e2 = cudaStreamEndCapture(cs, &g);
        if (e != cudaSuccess || e2 != cudaSuccess) { if (g) cudaGraphDestroy(g); throw MtvError("capture failed at " + op.name); }
        CK(cudaGraphInstantiate(&ge, g, 0));
        CK(cudaGraphLaunch(ge, s));               // warm
        CK(cudaEventRecord(e0, s));
        CK(cudaGraphLaunch(ge, s));
        CK(cudaEventRecord(e1, s));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
      } else {
        CK(cudaEventRecord(e0, s));
        for (int r = 0; r < reps; ++r) {
          cudaError_t e = op.fn(s);
          if (e != cudaSuccess) throw MtvError("launch failed at " + op.name + ": " + cudaGetErrorString(e));
        }
        CK(cudaEventRecord(e1, s));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
      }
      ms /= (float)reps;
      if (k < cap) {
        MtvKernelTime& kt = entries[k];
        snprintf(kt.name, sizeof(kt.name), "%s", op.name.c_str());
        kt.us = ms * 1000.f; kt.flops = (float)op.flops; kt.bytes = (float)op.bytes;
      }
      ++k;
    }
    pl->runs++; h->last_plan = pl;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *n = k < cap ? k : cap;
  });
}

}  // extern "C"
