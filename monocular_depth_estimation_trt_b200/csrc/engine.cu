// Engine (weights + description) and execution context (workspace + launch plan) behind the C ABI.
//
// Replaces what the reference gets from TensorRT: core/common.py:141-312 `get_engine` (engine) and the
// IExecutionContext driven by core/common_runtime.py:268-275 `do_inference` (context).  The forward is
// the DINOv2 ViT encoder + DPT head of Depth Anything V2; the plan is a flat list of kernel launches
// built once at context creation -- enqueue() only launches.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <map>
#include <string>
#include <vector>

#include "host_common.h"

namespace mde {

struct HostTensor {
  std::vector<int64_t> dims;
  std::vector<float> data;
  int64_t numel() const {
    int64_t n = 1;
    for (int64_t d : dims) n *= d;
    return n;
  }
};

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

static uint16_t to16(float v, int precision) {
  if (precision == MDE_BF16) {
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    return *reinterpret_cast<uint16_t*>(&h);
  }
  __half h = __float2half_rn(v);
  return *reinterpret_cast<uint16_t*>(&h);
}

struct Block {
  float *ln1_w, *ln1_b, *qkv_b, *proj_b, *ls1, *ln2_w, *ln2_b, *fc1_b, *fc2_b, *ls2;
  void *qkv_w, *proj_w, *fc1_w, *fc2_w;
};
struct Rcu {
  void *w1, *w2;
  float *b1, *b2;
};
struct Refine {
  Rcu rcu1, rcu2;
  void* out_w;
  float* out_b;
};

}  // namespace mde

using namespace mde;

struct mde_engine {
  mde_engine_desc d;
  std::map<std::string, HostTensor> raw;
  bool finalized = false;
  std::vector<void*> allocs;
  // geometry
  int gh = 0, gw = 0, T = 0, ntok = 0, kpad = 0;
  int lvl_h[4] = {0, 0, 0, 0}, lvl_w[4] = {0, 0, 0, 0};
  // packed device weights
  void* pe_w = nullptr;
  float *pe_b = nullptr, *cls = nullptr, *pos = nullptr, *norm_w = nullptr, *norm_b = nullptr, *lut = nullptr, *reg = nullptr;
  std::vector<Block> blocks;
  void* proj_w[4] = {nullptr, nullptr, nullptr, nullptr};
  float* proj_b[4] = {nullptr, nullptr, nullptr, nullptr};
  void *ct0_w = nullptr, *ct1_w = nullptr, *rs3_w = nullptr;
  float *ct0_b = nullptr, *ct1_b = nullptr, *rs3_b = nullptr;
  void* rn_w[4] = {nullptr, nullptr, nullptr, nullptr};
  Refine refine[4];   // index i = refinenet{i+1}
  void *oc1_w = nullptr, *oc2_w = nullptr, *sky2_w = nullptr;
  float *oc1_b = nullptr, *oc2_b = nullptr, *head_w = nullptr, *sky2_b = nullptr, *sky_head_w = nullptr;
  float head_b = 0.f, sky_head_b = 0.f;
  int64_t weight_bytes = 0;
};

namespace {

struct Op {
  enum Kind { PREPROC_U8, IM2COL_F32, CLS_ROW, GEMM, LAYERNORM, ATTENTION, BILINEAR, IM2COL_S2, UPCONV_HEAD, RESIZE_DEPTH, SNAPSHOT, JOIN } kind;
  GemmOp g;
  AttnOp attn;
  // generic scalar/pointer slots for the small kernels
  const void* in = nullptr;
  void* out = nullptr;
  const float *w = nullptr, *b = nullptr;
  long long rows = 0;
  int i0 = 0, i1 = 0, i2 = 0, i3 = 0, i4 = 0, i5 = 0;
  int block = -1;   // encoder block this op belongs to (SNAPSHOT ops)
  int branch = 0;   // 0: the caller's stream; 1..3: a side stream forked from it (small batches: the DPT reassemble stage's four
                    // independent project -> resize -> layer_rn chains run side by side, see enqueue_impl)
  std::string label;   // what the launch is, for per-op timing reports
  double flops = 0.0;  // algorithmic FLOPs (2*MACs, logical sizes, no padding)
  double bytes = 0.0;  // algorithmic bytes (operands read once + result written once)
};

}  // namespace

struct mde_context {
  mde_engine* e = nullptr;
  std::vector<void*> allocs;
  char* arena = nullptr;
  std::vector<Op> plan;
  std::map<std::string, std::pair<void*, int64_t>> named;   // debug buffers
  std::map<std::string, int> named_dtype;
  void* d_input = nullptr;
  void* d_output = nullptr;
  void* d_output2 = nullptr;       // MDE_HEAD_DPT_EXP_SKY: the sky map
  int src_h = 0, src_w = 0;
  int snapshot_block = -1;
  int gather_ranks = 0, gather_rank = 0;       // mde_context_set_gather
  // The launch sequence is captured into a CUDA graph the first time it runs with a given set of bindings and replayed
  // afterwards: one cudaGraphLaunch instead of ~165 kernel launches (what dominates a batch-1 forward on the host side).
  unsigned int* attn_counters = nullptr;   // [depth][2], see mde_context_create
  cudaStream_t side[3] = {nullptr, nullptr, nullptr};      // forked branches of the plan (Op::branch 1..3)
  cudaEvent_t fork_ev = nullptr, join_ev[3] = {nullptr, nullptr, nullptr};
  cudaGraphExec_t graph_exec = nullptr;
  unsigned long long graph_key[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  bool graph_failed = false;
  void* gather_dst[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  float* x_snapshot = nullptr;
  float* x = nullptr;
  int64_t x_bytes = 0;
  int64_t workspace_bytes = 0;
  std::vector<cudaEvent_t> events;
};

// MDE_PROFILE=1: profiling runs launch kernel by kernel without PDL, so that Nsight Compute lists every launch.  It does
// not change what is computed; everything that does lives in mde_engine_desc (flags, attn_poly).
static bool profile_mode() {
  static const bool v = [] { const char* e = getenv("MDE_PROFILE"); return e && atoi(e) != 0; }();
  return v;
}

namespace {
// Makes the engine's device current and installs its launch options for the calling thread; restores both on exit.
// One process may hold engines on several GPUs (mde_engine_desc.device).
struct EngineScope {
  int prev_dev = -1;
  bool switched = false;
  LaunchOpts prev_opts;
  int rc = MDE_OK;
  explicit EngineScope(const mde_engine_desc& d) {
    prev_opts = launch_opts();
    LaunchOpts o;
    o.pdl = !(d.flags & MDE_FLAG_NO_PDL) && !profile_mode();
    o.split_k = (d.flags & MDE_FLAG_SPLIT_K) != 0;
    set_launch_opts(o);
    if (cudaGetDevice(&prev_dev) != cudaSuccess) { rc = fail(MDE_ERR_CUDA, "cudaGetDevice failed (no CUDA device?)"); return; }
    if (prev_dev != d.device) {
      if (cudaSetDevice(d.device) != cudaSuccess) { rc = fail(MDE_ERR_CUDA, "cudaSetDevice(%d) failed", d.device); return; }
      switched = true;
    }
  }
  ~EngineScope() {
    set_launch_opts(prev_opts);
    if (switched) cudaSetDevice(prev_dev);
  }
};
}  // namespace

// =============================================================================================== engine
static int require(const mde_engine* e, const std::string& name, std::vector<int64_t> dims, const HostTensor** out) {
  auto it = e->raw.find(name);
  if (it == e->raw.end()) return fail(MDE_ERR_MISSING, "missing weight tensor '%s'", name.c_str());
  if (it->second.dims != dims) {
    std::string got, want;
    for (auto v : it->second.dims) got += std::to_string(v) + ",";
    for (auto v : dims) want += std::to_string(v) + ",";
    return fail(MDE_ERR_INVALID, "weight '%s' has shape [%s] but [%s] is required", name.c_str(), got.c_str(), want.c_str());
  }
  *out = &it->second;
  return MDE_OK;
}

static int upload(mde_engine* e, const void* host, size_t bytes, void** dev) {
  void* p = nullptr;
  MDE_CUDA_TRY(cudaMalloc(&p, bytes));
  e->allocs.push_back(p);
  MDE_CUDA_TRY(cudaMemcpy(p, host, bytes, cudaMemcpyHostToDevice));
  e->weight_bytes += static_cast<int64_t>(bytes);
  *dev = p;
  return MDE_OK;
}
static int upload_f32(mde_engine* e, const std::string& name, std::vector<int64_t> dims, float** dev) {
  const HostTensor* t;
  MDE_TRY(require(e, name, dims, &t));
  return upload(e, t->data.data(), t->data.size() * 4, reinterpret_cast<void**>(dev));
}
// [rows][cols] fp32 -> 16-bit [rows][ld] (zero padded)
static int upload_mat16(mde_engine* e, const float* src, int rows, int cols, int ld, void** dev) {
  std::vector<uint16_t> h(static_cast<size_t>(rows) * ld, 0);
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < cols; ++c) h[static_cast<size_t>(r) * ld + c] = to16(src[static_cast<size_t>(r) * cols + c], e->d.precision);
  return upload(e, h.data(), h.size() * 2, dev);
}
static int upload_linear(mde_engine* e, const std::string& name, int out_f, int in_f, void** dev) {
  const HostTensor* t;
  MDE_TRY(require(e, name, {out_f, in_f}, &t));
  return upload_mat16(e, t->data.data(), out_f, in_f, in_f, dev);
}
// Conv2d weight [cout][cin][3][3] -> [cout][9*cin_pad], K index = (ky*3+kx)*cin_pad + c
static int upload_conv3x3(mde_engine* e, const std::string& name, int cout, int cin, int cin_pad, void** dev) {
  const HostTensor* t;
  MDE_TRY(require(e, name, {cout, cin, 3, 3}, &t));
  std::vector<uint16_t> h(static_cast<size_t>(cout) * 9 * cin_pad, 0);
  for (int o = 0; o < cout; ++o)
    for (int c = 0; c < cin; ++c)
      for (int tap = 0; tap < 9; ++tap)
        h[(static_cast<size_t>(o) * 9 + tap) * cin_pad + c] =
            to16(t->data[(static_cast<size_t>(o) * cin + c) * 9 + tap], e->d.precision);
  return upload(e, h.data(), h.size() * 2, dev);
}
// ConvTranspose2d weight [cin][cout][s][s] (kernel == stride) -> GEMM B [s*s*cout][cin], row = (ky*s+kx)*cout + o
static int upload_convT(mde_engine* e, const std::string& name, int c, int s, void** dev) {
  const HostTensor* t;
  MDE_TRY(require(e, name, {c, c, s, s}, &t));
  std::vector<uint16_t> h(static_cast<size_t>(s) * s * c * c, 0);
  for (int i = 0; i < c; ++i)
    for (int o = 0; o < c; ++o)
      for (int q = 0; q < s * s; ++q)
        h[(static_cast<size_t>(q) * c + o) * c + i] = to16(t->data[(static_cast<size_t>(i) * c + o) * s * s + q], e->d.precision);
  return upload(e, h.data(), h.size() * 2, dev);
}
static int upload_rcu(mde_engine* e, const std::string& pre, int F, Rcu* r) {
  const int Fp = round_up(F, 64);
  MDE_TRY(upload_conv3x3(e, pre + "conv1.weight", F, F, Fp, &r->w1));
  MDE_TRY(upload_f32(e, pre + "conv1.bias", {F}, &r->b1));
  MDE_TRY(upload_conv3x3(e, pre + "conv2.weight", F, F, Fp, &r->w2));
  MDE_TRY(upload_f32(e, pre + "conv2.bias", {F}, &r->b2));
  return MDE_OK;
}

static int validate_desc(const mde_engine_desc* d) {
  if (!d) return fail(MDE_ERR_INVALID, "null description");
  if (d->struct_size != static_cast<int32_t>(sizeof(mde_engine_desc)))
    return fail(MDE_ERR_INVALID, "mde_engine_desc size mismatch: caller %d, library %d", d->struct_size,
                static_cast<int>(sizeof(mde_engine_desc)));
  if (d->precision != MDE_FP16 && d->precision != MDE_BF16) return fail(MDE_ERR_INVALID, "precision must be MDE_FP16 or MDE_BF16");
  if (d->input_mode != MDE_INPUT_F32_NCHW && d->input_mode != MDE_INPUT_U8_HWC) return fail(MDE_ERR_INVALID, "unknown input_mode %d", d->input_mode);
  if (d->embed_dim <= 0 || d->num_heads <= 0 || d->embed_dim != d->num_heads * 64)
    return fail(MDE_ERR_INVALID, "head dim must be 64 (embed_dim %d, heads %d)", d->embed_dim, d->num_heads);
  if (d->embed_dim != 384 && d->embed_dim != 768 && d->embed_dim != 1024 && d->embed_dim != 1536)
    return fail(MDE_ERR_INVALID, "unsupported embed_dim %d", d->embed_dim);
  if (d->depth <= 0 || d->depth > 64) return fail(MDE_ERR_INVALID, "bad depth %d", d->depth);
  if (d->patch_size <= 0 || d->input_h <= 0 || d->input_w <= 0 || d->input_h % d->patch_size || d->input_w % d->patch_size)
    return fail(MDE_ERR_INVALID, "input %dx%d must be a positive multiple of patch %d", d->input_h, d->input_w, d->patch_size);
  if (d->batch <= 0 || d->batch > 4096) return fail(MDE_ERR_INVALID, "bad batch %d", d->batch);
  if (d->output_mode != MDE_OUTPUT_MODEL_GRID && d->output_mode != MDE_OUTPUT_SOURCE_GRID) return fail(MDE_ERR_INVALID, "unknown output_mode %d", d->output_mode);
  if (d->output_mode == MDE_OUTPUT_SOURCE_GRID && (d->head_mode != MDE_HEAD_DPT || d->max_src_h <= 0 || d->max_src_w <= 0))
    return fail(MDE_ERR_INVALID, "MDE_OUTPUT_SOURCE_GRID needs the DPT head and max_src_h / max_src_w");
  if (d->head_mode != MDE_HEAD_DPT && d->head_mode != MDE_HEAD_ENCODER_TAPS && d->head_mode != MDE_HEAD_DPT_EXP_SKY)
    return fail(MDE_ERR_INVALID, "unknown head_mode %d", d->head_mode);
  if (d->tap_norm_mask < 0 || d->tap_norm_mask > 0xF) return fail(MDE_ERR_INVALID, "tap_norm_mask must be a 4-bit mask");
  if (d->head_mode != MDE_HEAD_ENCODER_TAPS && d->tap_norm_mask != 0xF) return fail(MDE_ERR_INVALID, "the DPT head takes all four taps through the final LayerNorm (tap_norm_mask 0xF)");
  if (d->head_mode != MDE_HEAD_ENCODER_TAPS && (d->features <= 0 || d->features % 16)) return fail(MDE_ERR_INVALID, "features must be a positive multiple of 16");
  for (int i = 0; i < 4; ++i) {
    if (d->head_mode != MDE_HEAD_ENCODER_TAPS && (d->out_channels[i] <= 0 || d->out_channels[i] % 8)) return fail(MDE_ERR_INVALID, "out_channels must be positive multiples of 8");
    if (d->taps[i] < 0 || d->taps[i] >= d->depth || (i > 0 && d->taps[i] <= d->taps[i - 1]))
      return fail(MDE_ERR_INVALID, "taps must be increasing block indices below depth");
  }
  if (d->flags & ~(MDE_FLAG_SPLIT_K | MDE_FLAG_NO_PDL | MDE_FLAG_NO_GRAPH | MDE_FLAG_SCALE_F32 | MDE_FLAG_NORMALISE_F32)) return fail(MDE_ERR_INVALID, "unknown bits in flags 0x%x", d->flags);
  if (d->num_registers < 0 || d->num_registers > 16) return fail(MDE_ERR_INVALID, "num_registers must be 0..16");
  if (d->attn_poly < -1 || d->attn_poly > 4) return fail(MDE_ERR_INVALID, "attn_poly must be -1 (default) or 0..4 eighths");
  if (d->flags & MDE_FLAG_NORMALISE_F32) {
    if (d->input_mode != MDE_INPUT_F32_NCHW) return fail(MDE_ERR_INVALID, "MDE_FLAG_NORMALISE_F32 belongs to the float32 input binding");
    for (int c = 0; c < 3; ++c)
      if (!(d->norm_std[c] > 0.0)) return fail(MDE_ERR_INVALID, "norm_std must be positive");
  }
  if (d->input_mode == MDE_INPUT_U8_HWC) {
    if (d->max_src_h <= 0 || d->max_src_w <= 0) return fail(MDE_ERR_INVALID, "max_src_h/max_src_w are required for the uint8 input");
    for (int c = 0; c < 3; ++c)
      if (!(d->norm_std[c] > 0.0)) return fail(MDE_ERR_INVALID, "norm_std must be positive");
  }
  return MDE_OK;
}

extern "C" int mde_engine_create(const mde_engine_desc* desc, mde_engine** out) {
  clear_error();
  if (!out) return fail(MDE_ERR_INVALID, "null output pointer");
  *out = nullptr;
  MDE_TRY(validate_desc(desc));
  mde_engine* e = new mde_engine();
  e->d = *desc;
  e->gh = desc->input_h / desc->patch_size;
  e->gw = desc->input_w / desc->patch_size;
  e->T = e->gh * e->gw;
  e->ntok = e->T + 1 + desc->num_registers;
  e->kpad = round_up(3 * desc->patch_size * desc->patch_size, 64);
  e->lvl_h[0] = 4 * e->gh; e->lvl_w[0] = 4 * e->gw;
  e->lvl_h[1] = 2 * e->gh; e->lvl_w[1] = 2 * e->gw;
  e->lvl_h[2] = e->gh; e->lvl_w[2] = e->gw;
  e->lvl_h[3] = (e->gh - 1) / 2 + 1; e->lvl_w[3] = (e->gw - 1) / 2 + 1;
  *out = e;
  return MDE_OK;
}

extern "C" int mde_engine_set_weight(mde_engine* e, const char* name, const float* data, int32_t ndim, const int64_t* dims) {
  clear_error();
  if (!e || !name || !data || ndim < 0 || ndim > 8 || (ndim > 0 && !dims)) return fail(MDE_ERR_INVALID, "bad argument to mde_engine_set_weight");
  if (e->finalized) return fail(MDE_ERR_STATE, "engine is already finalized");
  HostTensor t;
  t.dims.assign(dims, dims + ndim);
  const int64_t n = t.numel();
  if (n <= 0 || n > (1LL << 31)) return fail(MDE_ERR_INVALID, "weight '%s' has a bad element count", name);
  t.data.assign(data, data + n);
  e->raw[name] = std::move(t);
  return MDE_OK;
}

// .mdew: "MDEW0001" | u32 meta_len | meta JSON | u32 count | count x { u32 name_len | name | u32 ndim | i64 dims[ndim] | f32 data[] }
extern "C" int mde_engine_load_weights(mde_engine* e, const char* path) {
  clear_error();
  if (!e || !path) return fail(MDE_ERR_INVALID, "bad argument to mde_engine_load_weights");
  FILE* f = fopen(path, "rb");
  if (!f) return fail(MDE_ERR_IO, "cannot open weights file '%s'", path);
  char magic[8];
  uint32_t count = 0, meta_len = 0;
  int rc = MDE_OK;
  // the JSON description is for the host layer (it fills mde_engine_desc from it); skipped here
  if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "MDEW0001", 8) != 0 || fread(&meta_len, 4, 1, f) != 1 ||
      meta_len > (1u << 20) || fseek(f, static_cast<long>(meta_len), SEEK_CUR) != 0 || fread(&count, 4, 1, f) != 1)
    rc = fail(MDE_ERR_IO, "'%s' is not an MDEW0001 weights file", path);
  for (uint32_t i = 0; rc == MDE_OK && i < count; ++i) {
    uint32_t nlen = 0, ndim = 0;
    if (fread(&nlen, 4, 1, f) != 1 || nlen == 0 || nlen > 512) { rc = fail(MDE_ERR_IO, "corrupt tensor header %u in '%s'", i, path); break; }
    std::string name(nlen, '\0');
    if (fread(&name[0], 1, nlen, f) != nlen || fread(&ndim, 4, 1, f) != 1 || ndim > 8) { rc = fail(MDE_ERR_IO, "corrupt tensor header %u in '%s'", i, path); break; }
    int64_t dims[8];
    if (ndim && fread(dims, 8, ndim, f) != ndim) { rc = fail(MDE_ERR_IO, "corrupt dims of '%s'", name.c_str()); break; }
    int64_t n = 1;
    for (uint32_t k = 0; k < ndim; ++k) n *= dims[k];
    if (n <= 0 || n > (1LL << 31)) { rc = fail(MDE_ERR_IO, "bad element count for '%s'", name.c_str()); break; }
    std::vector<float> buf(static_cast<size_t>(n));
    if (fread(buf.data(), 4, static_cast<size_t>(n), f) != static_cast<size_t>(n)) { rc = fail(MDE_ERR_IO, "truncated data for '%s'", name.c_str()); break; }
    rc = mde_engine_set_weight(e, name.c_str(), buf.data(), static_cast<int32_t>(ndim), dims);
  }
  fclose(f);
  return rc;
}

extern "C" int mde_engine_finalize(mde_engine* e) {
  clear_error();
  if (!e) return fail(MDE_ERR_INVALID, "null engine");
  if (e->finalized) return MDE_OK;
  const mde_engine_desc& d = e->d;
  EngineScope scope(d);
  MDE_TRY(scope.rc);
  int major = 0;
  MDE_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d.device));
  if (major != 10) return fail(MDE_ERR_CUDA, "device %d has compute capability %d.x; this library only runs on sm_100a (B200)", d.device, major);
  const int D = d.embed_dim, P = d.patch_size, F = d.features;
  const int Fp = round_up(F, 64);
  const HostTensor* t;
  // ---- embeddings
  MDE_TRY(require(e, "pretrained.patch_embed.proj.weight", {D, 3, P, P}, &t));
  MDE_TRY(upload_mat16(e, t->data.data(), D, 3 * P * P, e->kpad, &e->pe_w));
  MDE_TRY(upload_f32(e, "pretrained.patch_embed.proj.bias", {D}, &e->pe_b));
  MDE_TRY(upload_f32(e, "pretrained.cls_token", {1, 1, D}, &e->cls));
  MDE_TRY(upload_f32(e, "pretrained.pos_embed", {1, e->T + 1, D}, &e->pos));   // already resized to this grid by the host
  if (d.num_registers > 0) MDE_TRY(upload_f32(e, "pretrained.register_tokens", {1, d.num_registers, D}, &e->reg));
  MDE_TRY(upload_f32(e, "pretrained.norm.weight", {D}, &e->norm_w));
  MDE_TRY(upload_f32(e, "pretrained.norm.bias", {D}, &e->norm_b));
  // ---- blocks
  e->blocks.resize(d.depth);
  for (int i = 0; i < d.depth; ++i) {
    const std::string p = "pretrained.blocks." + std::to_string(i) + ".";
    Block& b = e->blocks[i];
    MDE_TRY(upload_f32(e, p + "norm1.weight", {D}, &b.ln1_w));
    MDE_TRY(upload_f32(e, p + "norm1.bias", {D}, &b.ln1_b));
    MDE_TRY(upload_linear(e, p + "attn.qkv.weight", 3 * D, D, &b.qkv_w));
    MDE_TRY(upload_f32(e, p + "attn.qkv.bias", {3 * D}, &b.qkv_b));
    MDE_TRY(upload_linear(e, p + "attn.proj.weight", D, D, &b.proj_w));
    MDE_TRY(upload_f32(e, p + "attn.proj.bias", {D}, &b.proj_b));
    MDE_TRY(upload_f32(e, p + "ls1.gamma", {D}, &b.ls1));
    MDE_TRY(upload_f32(e, p + "norm2.weight", {D}, &b.ln2_w));
    MDE_TRY(upload_f32(e, p + "norm2.bias", {D}, &b.ln2_b));
    MDE_TRY(upload_linear(e, p + "mlp.fc1.weight", 4 * D, D, &b.fc1_w));
    MDE_TRY(upload_f32(e, p + "mlp.fc1.bias", {4 * D}, &b.fc1_b));
    MDE_TRY(upload_linear(e, p + "mlp.fc2.weight", D, 4 * D, &b.fc2_w));
    MDE_TRY(upload_f32(e, p + "mlp.fc2.bias", {D}, &b.fc2_b));
    MDE_TRY(upload_f32(e, p + "ls2.gamma", {D}, &b.ls2));
  }
  if (d.head_mode == MDE_HEAD_ENCODER_TAPS) {
    if (d.input_mode == MDE_INPUT_U8_HWC) {
      float lut[768];
      build_norm_lut(d.norm_mean, d.norm_std, lut, (d.flags & MDE_FLAG_SCALE_F32) != 0);
      MDE_TRY(upload(e, lut, sizeof(lut), reinterpret_cast<void**>(&e->lut)));
    }
    e->raw.clear();
    e->finalized = true;
    return MDE_OK;
  }
  // ---- DPT head
  const std::string h = "depth_head.";
  const int* oc = d.out_channels;
  for (int i = 0; i < 4; ++i) {
    MDE_TRY(require(e, h + "projects." + std::to_string(i) + ".weight", {oc[i], D, 1, 1}, &t));
    MDE_TRY(upload_mat16(e, t->data.data(), oc[i], D, D, &e->proj_w[i]));
    MDE_TRY(upload_f32(e, h + "projects." + std::to_string(i) + ".bias", {oc[i]}, &e->proj_b[i]));
    MDE_TRY(upload_conv3x3(e, h + "scratch.layer" + std::to_string(i + 1) + "_rn.weight", F, oc[i], round_up(oc[i], 64), &e->rn_w[i]));
    const std::string r = h + "scratch.refinenet" + std::to_string(i + 1) + ".";
    if (i != 3) MDE_TRY(upload_rcu(e, r + "resConfUnit1.", F, &e->refine[i].rcu1));   // refinenet4 never uses its RCU1
    MDE_TRY(upload_rcu(e, r + "resConfUnit2.", F, &e->refine[i].rcu2));
    MDE_TRY(require(e, r + "out_conv.weight", {F, F, 1, 1}, &t));
    MDE_TRY(upload_mat16(e, t->data.data(), F, F, F, &e->refine[i].out_w));
    MDE_TRY(upload_f32(e, r + "out_conv.bias", {F}, &e->refine[i].out_b));
  }
  MDE_TRY(upload_convT(e, h + "resize_layers.0.weight", oc[0], 4, &e->ct0_w));
  MDE_TRY(upload_f32(e, h + "resize_layers.0.bias", {oc[0]}, &e->ct0_b));
  MDE_TRY(upload_convT(e, h + "resize_layers.1.weight", oc[1], 2, &e->ct1_w));
  MDE_TRY(upload_f32(e, h + "resize_layers.1.bias", {oc[1]}, &e->ct1_b));
  MDE_TRY(upload_conv3x3(e, h + "resize_layers.3.weight", oc[3], oc[3], oc[3], &e->rs3_w));   // fed by the explicit gather: no padding
  MDE_TRY(upload_f32(e, h + "resize_layers.3.bias", {oc[3]}, &e->rs3_b));
  MDE_TRY(upload_conv3x3(e, h + "scratch.output_conv1.weight", F / 2, F, Fp, &e->oc1_w));
  MDE_TRY(upload_f32(e, h + "scratch.output_conv1.bias", {F / 2}, &e->oc1_b));
  // output_conv2[0] as nine 1x1 contractions applied BEFORE the bilinear upsampling (csrc/upconv_head.cuh):
  // GEMM B operand [384][F/2], row (ky*3+kx)*32 + o; rows 288..383 are zero padding up to a whole number of 128-wide tiles
  MDE_TRY(require(e, h + "scratch.output_conv2.0.weight", {32, F / 2, 3, 3}, &t));
  {
    const int cin = F / 2;
    std::vector<uint16_t> wz(static_cast<size_t>(384) * cin, 0);
    for (int o = 0; o < 32; ++o)
      for (int c = 0; c < cin; ++c)
        for (int tap = 0; tap < 9; ++tap)
          wz[(static_cast<size_t>(tap) * 32 + o) * cin + c] = to16(t->data[(static_cast<size_t>(o) * cin + c) * 9 + tap], d.precision);
    MDE_TRY(upload(e, wz.data(), wz.size() * 2, &e->oc2_w));
  }
  MDE_TRY(upload_f32(e, h + "scratch.output_conv2.0.bias", {32}, &e->oc2_b));
  MDE_TRY(upload_f32(e, h + "scratch.output_conv2.2.weight", {1, 32, 1, 1}, &e->head_w));
  MDE_TRY(require(e, h + "scratch.output_conv2.2.bias", {1}, &t));
  e->head_b = t->data[0];
  if (d.head_mode == MDE_HEAD_DPT_EXP_SKY) {
    // the sky branch: a second conv3x3 (F/2 -> 32) + ReLU + conv1x1 (32 -> 1) + ReLU on the same up-sampled map, packed like
    // output_conv2 (nine 1x1 contractions applied before the up-sampling)
    MDE_TRY(require(e, h + "scratch.sky_output_conv2.0.weight", {32, F / 2, 3, 3}, &t));
    const int cin = F / 2;
    std::vector<uint16_t> wz(static_cast<size_t>(384) * cin, 0);
    for (int o = 0; o < 32; ++o)
      for (int c = 0; c < cin; ++c)
        for (int tap = 0; tap < 9; ++tap)
          wz[(static_cast<size_t>(tap) * 32 + o) * cin + c] = to16(t->data[(static_cast<size_t>(o) * cin + c) * 9 + tap], d.precision);
    MDE_TRY(upload(e, wz.data(), wz.size() * 2, &e->sky2_w));
    MDE_TRY(upload_f32(e, h + "scratch.sky_output_conv2.0.bias", {32}, &e->sky2_b));
    MDE_TRY(upload_f32(e, h + "scratch.sky_output_conv2.2.weight", {1, 32, 1, 1}, &e->sky_head_w));
    MDE_TRY(require(e, h + "scratch.sky_output_conv2.2.bias", {1}, &t));
    e->sky_head_b = t->data[0];
  }
  if (d.input_mode == MDE_INPUT_U8_HWC) {
    float lut[768];
    build_norm_lut(d.norm_mean, d.norm_std, lut, (d.flags & MDE_FLAG_SCALE_F32) != 0);
    MDE_TRY(upload(e, lut, sizeof(lut), reinterpret_cast<void**>(&e->lut)));
  }
  e->raw.clear();
  e->finalized = true;
  return MDE_OK;
}

extern "C" void mde_engine_destroy(mde_engine* e) {
  if (!e) return;
  for (void* p : e->allocs) cudaFree(p);
  delete e;
}

static int io_count(const mde_engine* e) { return e->d.head_mode == MDE_HEAD_DPT_EXP_SKY ? 3 : 2; }
extern "C" int mde_engine_num_io(const mde_engine* e) { return e ? io_count(e) : 0; }
extern "C" const char* mde_engine_io_name(const mde_engine* e, int32_t i) {
  if (!e || i < 0 || i >= io_count(e)) return nullptr;
  if (e->d.head_mode == MDE_HEAD_DPT_EXP_SKY) return i == 0 ? "image" : (i == 1 ? "depth" : "sky");   // models/depth_anything_v3/spec.json
  return i == 0 ? "input" : "output";   // models/depth_anything_v2/spec.json input.name / outputs[0].name
}
extern "C" int mde_engine_io_shape(const mde_engine* e, int32_t i, int32_t* ndim, int64_t* dims) {
  clear_error();
  if (!e || !ndim || !dims || i < 0 || i >= io_count(e)) return fail(MDE_ERR_INVALID, "bad argument to mde_engine_io_shape");
  if (i == 0) {
    *ndim = 4;
    if (e->d.input_mode == MDE_INPUT_F32_NCHW) {
      dims[0] = e->d.batch; dims[1] = 3; dims[2] = e->d.input_h; dims[3] = e->d.input_w;
    } else {
      dims[0] = e->d.batch; dims[1] = e->d.max_src_h; dims[2] = e->d.max_src_w; dims[3] = 3;
    }
  } else if (e->d.head_mode == MDE_HEAD_ENCODER_TAPS) {
    *ndim = 4;
    dims[0] = 4; dims[1] = e->d.batch; dims[2] = e->T; dims[3] = e->d.embed_dim;
  } else if (e->d.output_mode == MDE_OUTPUT_SOURCE_GRID) {
    *ndim = 3;
    dims[0] = e->d.batch; dims[1] = e->d.max_src_h; dims[2] = e->d.max_src_w;
  } else {
    *ndim = 3;
    dims[0] = e->d.batch; dims[1] = e->d.input_h; dims[2] = e->d.input_w;
  }
  return MDE_OK;
}
extern "C" int mde_engine_io_dtype(const mde_engine* e, int32_t i) {
  if (!e || i < 0 || i >= io_count(e)) return -1;
  if (i == 1 && e->d.head_mode == MDE_HEAD_ENCODER_TAPS) return e->d.precision == MDE_BF16 ? MDE_DT_BF16 : MDE_DT_F16;
  return (i == 0 && e->d.input_mode == MDE_INPUT_U8_HWC) ? MDE_DT_U8 : MDE_DT_F32;
}
extern "C" int mde_engine_io_is_input(const mde_engine* e, int32_t i) {
  if (!e || i < 0 || i >= io_count(e)) return -1;
  return i == 0 ? 1 : 0;
}

// =============================================================================================== context
namespace {

struct Planner {
  mde_context* c;
  mde_engine* e;
  int prec;
  int rc = MDE_OK;
  bool dry;            // dry run: only add up the workspace bytes
  int64_t bytes = 0;
  int branch = 0;      // stamped on every op pushed while it is set

  // One arena per context, bump-allocated with LIFO scopes: a tensor lives from its alloc() to the release() of the scope it
  // was allocated in, so the encoder's temporaries, the reassemble stage, each RefineNet level and the tail share memory
  // (ViT-L, batch 64: 10.9 GiB instead of 19.8 GiB with one allocation per tensor).  The dry run makes the same calls with a
  // null arena and only records the high-water mark; named buffers (mde_context_get_buffer) live outside every scope.
  char* arena = nullptr;
  int64_t top = 0;
  void* alloc(int64_t n, const char* name = nullptr, int dtype = 1) {
    n = (n + 255) / 256 * 256;
    const int64_t at = top;
    top += n;
    if (top > bytes) bytes = top;
    if (dry || rc != MDE_OK) return nullptr;
    void* p = arena + at;
    if (name) { c->named[name] = {p, n}; c->named_dtype[name] = dtype; }
    return p;
  }
  int64_t mark() const { return top; }
  void release(int64_t m) { top = m; }
  void* alloc16(int64_t elems, const char* name = nullptr) { return alloc(elems * 2, name, 1); }

  void gemm(const char* what, const void* a, long long m, int k, int lda, const void* b, int n, int ldb,
            const mde_epilogue& ep, int k_real = 0, int n_real = 0) {
    if (dry || rc != MDE_OK) return;
    Op op; op.kind = Op::GEMM;
    rc = make_gemm_op(&op.g, prec, a, m, k, lda, b, n, ldb, &ep);
    if (k_real <= 0) k_real = k;
    if (n_real <= 0) n_real = n;      // algorithmic sizes: zero padding of K or N is not counted
    char buf[160];
    snprintf(buf, sizeof(buf), "gemm%d %s %lldx%dx%d", op.g.block_n, what, m, n_real, k_real);
    op.label = buf;
    op.flops = 2.0 * static_cast<double>(m) * n_real * k_real;
    op.bytes = 2.0 * (static_cast<double>(m) * k_real + static_cast<double>(n_real) * k_real) + epilogue_bytes(ep, static_cast<double>(m) * n_real);
    op.branch = branch;
    if (rc == MDE_OK) c->plan.push_back(op);
  }
  void conv(const char* what, const void* in, int B, int H, int W, int cin, const void* w, int cout, const mde_epilogue& ep) {
    if (dry || rc != MDE_OK) return;
    Op op; op.kind = Op::GEMM;
    rc = make_conv_op(&op.g, prec, in, B, H, W, cin, w, cout, &ep);
    char buf[160];
    snprintf(buf, sizeof(buf), "conv%d %s %dx%dx%d %d->%d", op.g.block_n, what, B, H, W, cin, cout);
    op.label = buf;
    const double px = static_cast<double>(B) * H * W;
    op.flops = 2.0 * px * cout * 9.0 * cin + (ep.d_head_w ? 2.0 * px * 32 : 0.0);
    op.bytes = 2.0 * (px * cin + 9.0 * cin * cout) + (ep.d_head_w ? 4.0 * px : epilogue_bytes(ep, px * cout));
    op.branch = branch;
    if (rc == MDE_OK) c->plan.push_back(op);
  }
  static double epilogue_bytes(const mde_epilogue& ep, double elems) {
    double b = 0.0;
    if (ep.d_x) b += (ep.accumulate_x ? 8.0 : 4.0) * elems;
    if (ep.d_out) b += 2.0 * elems;
    if (ep.d_out_relu) b += 2.0 * elems;
    if (ep.d_res1) b += 2.0 * elems;
    if (ep.d_res2) b += 2.0 * elems;
    return b;
  }
  void push(const Op& op, const char* label = "", double bytes = 0.0, double flops = 0.0) {
    if (dry || rc != MDE_OK) return;
    c->plan.push_back(op);
    c->plan.back().label = label;
    c->plan.back().bytes = bytes;
    c->plan.back().flops = flops;
    c->plan.back().branch = branch;
  }
};

mde_epilogue ep_zero() {
  mde_epilogue ep;
  memset(&ep, 0, sizeof(ep));
  return ep;
}

// Build (or only size) the whole forward.
int build_plan(mde_context* c, mde_engine* e, bool dry, int64_t* bytes_out) {
  const mde_engine_desc& d = e->d;
  Planner pl{c, e, d.precision};
  pl.dry = dry;
  pl.arena = (c && !dry) ? c->arena : nullptr;
  const int B = d.batch, D = d.embed_dim, T = e->T, NT = e->ntok, F = d.features;
  const long long rows = static_cast<long long>(B) * NT;
  const long long prow = static_cast<long long>(B) * T;
  const int* oc = d.out_channels;

  // ---- workspace
  void* cols = pl.alloc16(prow * e->kpad, "cols");
  float* x = static_cast<float*>(pl.alloc(rows * D * 4, "x", 0));
  float* xs = static_cast<float*>(pl.alloc(rows * D * 4, "x_snapshot", 0));
  const bool taps_only = d.head_mode == MDE_HEAD_ENCODER_TAPS;   // the taps go straight to the output binding / the gather buffers
  void* tap[4] = {nullptr, nullptr, nullptr, nullptr};
  const char* tap_names[4] = {"tap0", "tap1", "tap2", "tap3"};
  if (!taps_only)
    for (int i = 0; i < 4; ++i) tap[i] = pl.alloc16(prow * D, tap_names[i]);
  // head tensors that outlive their producers' scopes: the four layer_rn maps (and their ReLU copies) and the two fusion
  // outputs that alternate between the RefineNet levels (level 3 and 1 write A, level 2 and 0 write B = "path_1")
  const int F0 = d.features;
  void *r[4] = {nullptr, nullptr, nullptr, nullptr}, *r_relu[4] = {nullptr, nullptr, nullptr, nullptr};
  void *path_a = nullptr, *path_b = nullptr;
  if (!taps_only) {
    const char* r_names[4] = {"r0", "r1", "r2", "r3"};
    for (int i = 0; i < 4; ++i) {
      const long long px = static_cast<long long>(B) * e->lvl_h[i] * e->lvl_w[i];
      r[i] = pl.alloc16(px * F0, r_names[i]);
      r_relu[i] = pl.alloc16(px * F0);
    }
    auto out_px = [&](int i) { return static_cast<long long>(B) * (i > 0 ? e->lvl_h[i - 1] : 2 * e->lvl_h[0]) * (i > 0 ? e->lvl_w[i - 1] : 2 * e->lvl_w[0]); };
    path_a = pl.alloc16(std::max(out_px(3), out_px(1)) * F0);
    path_b = pl.alloc16(std::max(out_px(2), out_px(0)) * F0, "path_1");
  }
  const int64_t encoder_scope = pl.mark();
  void* ln = pl.alloc16(rows * D);
  void* qkv = pl.alloc16(rows * 3 * D);
  void* att = pl.alloc16(rows * D);
  void* hid = pl.alloc16(rows * 4 * D);
  if (!dry) { c->x = x; c->x_snapshot = xs; c->x_bytes = rows * D * 4; }

  // ---- embed
  {
    Op op;
    if (d.input_mode == MDE_INPUT_U8_HWC) { op.kind = Op::PREPROC_U8; op.out = cols; }
    else { op.kind = Op::IM2COL_F32; op.out = cols; }
    const int kreal = 3 * d.patch_size * d.patch_size;
    if (d.input_mode == MDE_INPUT_U8_HWC)   // bytes depend on the source size bound at enqueue; counted for the maximum
      pl.push(op, "preprocess_u8 resize+normalise+im2col", static_cast<double>(B) * d.max_src_h * d.max_src_w * 3 + 2.0 * prow * e->kpad);
    else
      pl.push(op, "im2col_f32", 4.0 * B * 3 * d.input_h * d.input_w + 2.0 * prow * e->kpad);
    mde_epilogue ep = ep_zero();
    ep.d_bias = e->pe_b; ep.d_x = x; ep.ld_out = D; ep.tokens = T; ep.d_pos = e->pos; ep.token_skip = 1 + d.num_registers;
    pl.gemm("patch_embed", cols, prow, e->kpad, e->kpad, e->pe_w, D, e->kpad, ep, kreal);
    Op cr; cr.kind = Op::CLS_ROW; cr.out = x;
    pl.push(cr, "cls_row", (12.0 + 8.0 * d.num_registers) * D * B);
  }
  // ---- encoder
  int next_tap = 0;
  for (int i = 0; i < d.depth; ++i) {
    const Block& b = e->blocks.empty() ? Block{} : e->blocks[i];
    Op l1; l1.kind = Op::LAYERNORM; l1.in = x; l1.out = ln; l1.w = b.ln1_w; l1.b = b.ln1_b; l1.rows = rows; l1.i0 = 0; l1.i1 = 0; l1.i2 = -1;
    pl.push(l1, "layernorm", 6.0 * rows * D);
    { mde_epilogue ep = ep_zero(); ep.d_bias = b.qkv_b; ep.d_out = qkv; ep.ld_out = 3 * D;
      pl.gemm("qkv", ln, rows, D, D, b.qkv_w, 3 * D, D, ep); }
    { Op a; a.kind = Op::ATTENTION; a.in = qkv; a.out = att;
      if (!dry && pl.rc == MDE_OK) pl.rc = make_attention_op(&a.attn, d.precision, qkv, att, B, NT, d.num_heads);
      if (d.attn_poly >= 0) a.attn.poly = d.attn_poly;
      if (!dry && c && c->attn_counters) a.attn.counters = c->attn_counters + 2 * i;   // this context's op alone uses the pair
      pl.push(a, "attention", 8.0 * rows * D, 4.0 * static_cast<double>(B) * NT * NT * D); }
    { mde_epilogue ep = ep_zero(); ep.d_bias = b.proj_b; ep.d_gamma = b.ls1; ep.d_x = x; ep.accumulate_x = 1; ep.ld_out = D;
      pl.gemm("proj+ls+res", att, rows, D, D, b.proj_w, D, D, ep); }
    Op l2 = l1; l2.w = b.ln2_w; l2.b = b.ln2_b;
    pl.push(l2, "layernorm", 6.0 * rows * D);
    { mde_epilogue ep = ep_zero(); ep.d_bias = b.fc1_b; ep.act = 1; ep.d_out = hid; ep.ld_out = 4 * D;
      pl.gemm("fc1+gelu", ln, rows, D, D, b.fc1_w, 4 * D, D, ep); }
    { mde_epilogue ep = ep_zero(); ep.d_bias = b.fc2_b; ep.d_gamma = b.ls2; ep.d_x = x; ep.accumulate_x = 1; ep.ld_out = D;
      pl.gemm("fc2+ls+res", hid, rows, 4 * D, 4 * D, b.fc2_w, D, 4 * D, ep); }
    { Op s; s.kind = Op::SNAPSHOT; s.block = i; pl.push(s, "snapshot"); }
    if (next_tap < 4 && d.taps[next_tap] == i) {
      Op t; t.kind = Op::LAYERNORM; t.in = x; t.out = tap[next_tap]; t.w = e->norm_w; t.b = e->norm_b; t.rows = rows;
      t.i0 = 1 + d.num_registers;                              // cls and the registers are dropped from the taps
      t.i1 = ((d.tap_norm_mask >> next_tap) & 1) ? 0 : 1;     // identity: raw block output
      t.i2 = taps_only ? next_tap : -1;                        // which slice of the output binding
      pl.push(t, "layernorm tap", 4.0 * rows * D + 2.0 * prow * D);
      ++next_tap;
    }
  }
  if (taps_only) {
    if (bytes_out) *bytes_out = pl.bytes;
    return pl.rc;
  }
  pl.release(encoder_scope);
  // ---- DPT reassemble
  const int gh = e->gh, gw = e->gw;
  const int64_t reassemble_scope = pl.mark();     // projections, resized maps and the stride-2 gather die with the layer_rn convs
  // Small batches: the four chains project -> resize -> layer_rn share nothing but their inputs' producer (every temporary
  // below has its own arena range until the scope is released) and each fills a fraction of the SMs, so they run side by
  // side: chain 3 (stride-2 gather + K = 9 * oc[3] GEMM, the longest) on the caller's stream, chains 0-2 on side streams.
  const bool fork = rows <= 4LL * NT;
  auto branch_of = [&](int i) { return fork && i < 3 ? i + 1 : 0; };
  void* l[4];
  for (int i = 0; i < 4; ++i) {
    pl.branch = branch_of(i);
    void* pr = pl.alloc16(prow * oc[i]);
    mde_epilogue ep = ep_zero(); ep.d_bias = e->proj_b[i]; ep.d_out = pr; ep.ld_out = oc[i];
    pl.gemm("projects", tap[i], prow, D, D, e->proj_w[i], oc[i], D, ep);
    if (i == 0 || i == 1) {
      const int s = i == 0 ? 4 : 2;
      l[i] = pl.alloc16(prow * s * s * oc[i]);
      mde_epilogue e2 = ep_zero(); e2.d_bias = i == 0 ? e->ct0_b : e->ct1_b; e2.d_out = l[i]; e2.ld_out = oc[i];
      e2.shuffle_s = s; e2.shuffle_cout = oc[i]; e2.shuffle_h = gh; e2.shuffle_w = gw;
      pl.gemm("convT+shuffle", pr, prow, oc[i], oc[i], i == 0 ? e->ct0_w : e->ct1_w, s * s * oc[i], oc[i], e2);
    } else if (i == 2) {
      l[i] = pr;
    } else {
      const long long r4 = static_cast<long long>(B) * e->lvl_h[3] * e->lvl_w[3];
      void* g = pl.alloc16(r4 * 9 * oc[3]);
      Op s2; s2.kind = Op::IM2COL_S2; s2.in = pr; s2.out = g; s2.i0 = gh; s2.i1 = gw; s2.i2 = oc[3];
      pl.push(s2, "im2col_s2", 2.0 * prow * oc[3] + 2.0 * r4 * 9 * oc[3]);
      l[i] = pl.alloc16(r4 * oc[3]);
      mde_epilogue e2 = ep_zero(); e2.d_bias = e->rs3_b; e2.d_out = l[i]; e2.ld_out = oc[3];
      pl.gemm("conv3x3s2", g, r4, 9 * oc[3], 9 * oc[3], e->rs3_w, oc[3], 9 * oc[3], e2);
    }
  }
  // ---- layer_rn: raw r_i (residual of the first RCU) and relu(r_i) (input of its first conv)
  for (int i = 0; i < 4; ++i) {
    pl.branch = branch_of(i);
    mde_epilogue ep = ep_zero(); ep.d_out = r[i]; ep.d_out_relu = r_relu[i]; ep.ld_out = F;
    pl.conv("layer_rn", l[i], B, e->lvl_h[i], e->lvl_w[i], oc[i], e->rn_w[i], F, ep);
  }
  pl.branch = 0;
  if (fork) { Op j; j.kind = Op::JOIN; pl.push(j, "join"); }
  pl.release(reassemble_scope);
  // ---- RefineNets 4 -> 1
  // Small batches: the first convolution of levels 2, 1 and 0 (RCU1.conv1 on relu(r_i)) needs nothing from the level above
  // it; the three run on the side streams beside level 3's chain and rejoin before level 2 adds `path`.  Their outputs then
  // live outside the per-level scopes (which alias one another).
  void* a_early[3] = {nullptr, nullptr, nullptr};
  if (fork) {
    for (int i = 2; i >= 0; --i) {
      const long long px = static_cast<long long>(B) * e->lvl_h[i] * e->lvl_w[i];
      a_early[i] = pl.alloc16(px * F);
      pl.branch = i + 1;
      mde_epilogue ep = ep_zero(); ep.d_bias = e->refine[i].rcu1.b1; ep.act = 2; ep.d_out = a_early[i]; ep.ld_out = F;
      pl.conv("rcu1.conv1", r_relu[i], B, e->lvl_h[i], e->lvl_w[i], F, e->refine[i].rcu1.w1, F, ep);
    }
    pl.branch = 0;
  }
  void* path = nullptr;   // output of the previous fusion block, already at this level's resolution
  for (int i = 3; i >= 0; --i) {
    const int H = e->lvl_h[i], W = e->lvl_w[i];
    const long long px = static_cast<long long>(B) * H * W;
    const Refine& rf = e->refine[i];
    const int64_t level_scope = pl.mark();
    void* a = pl.alloc16(px * F);
    const void* s_relu = r_relu[i];
    const void* s_raw = r[i];
    if (i != 3) {
      // s = path + RCU1(r_i) = conv2(relu(conv1(relu(r_i)))) + r_i + path
      void* s = pl.alloc16(px * F);
      void* sr = pl.alloc16(px * F);
      const void* a1 = a;
      if (fork) {
        a1 = a_early[i];
        if (i == 2) { Op j; j.kind = Op::JOIN; pl.push(j, "join"); }
      } else {
        mde_epilogue ep = ep_zero(); ep.d_bias = rf.rcu1.b1; ep.act = 2; ep.d_out = a; ep.ld_out = F;
        pl.conv("rcu1.conv1", r_relu[i], B, H, W, F, rf.rcu1.w1, F, ep);
      }
      { mde_epilogue ep = ep_zero(); ep.d_bias = rf.rcu1.b2; ep.d_res1 = r[i]; ep.d_res2 = path; ep.d_out = s; ep.d_out_relu = sr; ep.ld_out = F;
        pl.conv("rcu1.conv2+res", a1, B, H, W, F, rf.rcu1.w2, F, ep); }
      s_relu = sr; s_raw = s;
    }
    // u = RCU2(s)
    void* a2 = pl.alloc16(px * F);
    void* u = pl.alloc16(px * F);
    { mde_epilogue ep = ep_zero(); ep.d_bias = rf.rcu2.b1; ep.act = 2; ep.d_out = a2; ep.ld_out = F;
      pl.conv("rcu2.conv1", s_relu, B, H, W, F, rf.rcu2.w1, F, ep); }
    { mde_epilogue ep = ep_zero(); ep.d_bias = rf.rcu2.b2; ep.d_res1 = s_raw; ep.d_out = u; ep.ld_out = F;
      pl.conv("rcu2.conv2+res", a2, B, H, W, F, rf.rcu2.w2, F, ep); }
    // 1x1 out_conv at this resolution, then bilinear (align_corners=True) to the next level's size:
    // both are linear and the interpolation weights sum to 1, so they commute exactly in real arithmetic.
    void* q = pl.alloc16(px * F);
    { mde_epilogue ep = ep_zero(); ep.d_bias = rf.out_b; ep.d_out = q; ep.ld_out = F;
      pl.gemm("out_conv1x1", u, px, F, F, rf.out_w, F, F, ep); }
    const int Ho = i > 0 ? e->lvl_h[i - 1] : 2 * H, Wo = i > 0 ? e->lvl_w[i - 1] : 2 * W;
    path = (i == 3 || i == 1) ? path_a : path_b;           // the level's input `path` is the OTHER buffer
    Op bl; bl.kind = Op::BILINEAR; bl.in = q; bl.out = path; bl.i0 = H; bl.i1 = W; bl.i2 = Ho; bl.i3 = Wo; bl.i4 = F;
    pl.push(bl, "bilinear", 2.0 * F * (px + static_cast<double>(B) * Ho * Wo));
    pl.release(level_scope);
  }
  // ---- output convs + fused depth head
  {
    const int H1 = 2 * e->lvl_h[0], W1 = 2 * e->lvl_w[0];
    void* o1 = pl.alloc16(static_cast<long long>(B) * H1 * W1 * (F / 2));
    { mde_epilogue ep = ep_zero(); ep.d_bias = e->oc1_b; ep.d_out = o1; ep.ld_out = F / 2;
      pl.conv("output_conv1", path, B, H1, W1, F, e->oc1_w, F / 2, ep); }
    // output_conv2[0]'s channel contraction at THIS resolution (nine taps x 32 channels = 288 columns: two full 128-wide tiles
    // and a quarter-filled third one whose store boxes the tensor map clips),
    // then one kernel interpolates z to the input size, sums the taps, applies bias/ReLU and the 1x1 head.
    const long long px1 = static_cast<long long>(B) * H1 * W1;
    void* z = pl.alloc16(px1 * 288, "z_taps");
    { mde_epilogue ep = ep_zero(); ep.d_out = z; ep.ld_out = 288;
      pl.gemm("output_conv2 taps", o1, px1, F / 2, F / 2, e->oc2_w, 288, F / 2, ep); }
    Op uh; uh.kind = Op::UPCONV_HEAD; uh.in = z; uh.i0 = H1; uh.i1 = W1; uh.i2 = d.input_h; uh.i3 = d.input_w;
    const double opx = static_cast<double>(B) * d.input_h * d.input_w;
    if (d.output_mode == MDE_OUTPUT_SOURCE_GRID) {
      uh.out = pl.alloc(static_cast<long long>(B) * d.input_h * d.input_w * 4, "depth_model_grid", 0);   // head output; the binding gets the resized map
      pl.push(uh, "upconv_head interpolate+taps+head", 2.0 * 288 * px1 + 4.0 * opx, 2.0 * (9 * 4 * 32 + 32) * opx);
      Op rs; rs.kind = Op::RESIZE_DEPTH; rs.in = uh.out;
      pl.push(rs, "resize_depth to source size + clamp", 4.0 * opx + 4.0 * B * d.max_src_h * d.max_src_w);
    } else {
      pl.push(uh, "upconv_head interpolate+taps+head", 2.0 * 288 * px1 + 4.0 * opx, 2.0 * (9 * 4 * 32 + 32) * opx);
    }
    if (d.head_mode == MDE_HEAD_DPT_EXP_SKY) {
      // the sky branch on the same o1: its own tap contraction into the same z buffer (the depth branch has consumed it), then
      // the same interpolate + sum + ReLU + 1x1 kernel with a ReLU at the end, into the second output binding
      { mde_epilogue ep = ep_zero(); ep.d_out = z; ep.ld_out = 288;
        pl.gemm("sky_output_conv2 taps", o1, px1, F / 2, F / 2, e->sky2_w, 288, F / 2, ep); }
      Op us = uh; us.i4 = 1;      // i4 == 1: the sky variant (weights, activation, destination)
      pl.push(us, "upconv_head sky interpolate+taps+head", 2.0 * 288 * px1 + 4.0 * opx, 2.0 * (9 * 4 * 32 + 32) * opx);
    }
  }
  if (bytes_out) *bytes_out = pl.bytes;
  return pl.rc;
}

}  // namespace

extern "C" int64_t mde_engine_workspace_bytes(const mde_engine* e) {
  clear_error();
  if (!e) return -1;
  int64_t bytes = 0;
  if (build_plan(nullptr, const_cast<mde_engine*>(e), true, &bytes) != MDE_OK) return -1;
  return bytes;
}

extern "C" int mde_context_create(mde_engine* e, mde_context** out) {
  clear_error();
  if (!e || !out) return fail(MDE_ERR_INVALID, "bad argument to mde_context_create");
  *out = nullptr;
  if (!e->finalized) return fail(MDE_ERR_STATE, "mde_engine_finalize must succeed before a context is created");
  EngineScope scope(e->d);
  MDE_TRY(scope.rc);
  mde_context* c = new mde_context();
  c->e = e;
  int64_t bytes = 0;
  int rc = build_plan(nullptr, e, true, &bytes);           // sizes the arena: the same alloc / release sequence without memory
  if (rc == MDE_OK) {
    void* arena = nullptr;
    cudaError_t err = cudaMalloc(&arena, static_cast<size_t>(std::max<int64_t>(bytes, 256)));
    if (err != cudaSuccess) rc = fail(MDE_ERR_CUDA, "cudaMalloc of %lld workspace bytes failed: %s", static_cast<long long>(bytes), cudaGetErrorString(err));
    else { c->allocs.push_back(arena); c->arena = static_cast<char*>(arena); }
  }
  if (rc == MDE_OK && e->d.depth > 0) {
    // work counters of the persistent attention kernel, one {next item, CTAs done} pair per block (attention_q3.cuh)
    void* cnt = nullptr;
    const size_t cnt_bytes = static_cast<size_t>(e->d.depth) * 2 * sizeof(unsigned int);
    cudaError_t err = cudaMalloc(&cnt, cnt_bytes);
    if (err == cudaSuccess) { c->allocs.push_back(cnt); err = cudaMemset(cnt, 0, cnt_bytes); }
    if (err != cudaSuccess) rc = fail(MDE_ERR_CUDA, "attention work counters: %s", cudaGetErrorString(err));
    else c->attn_counters = static_cast<unsigned int*>(cnt);
  }
  if (rc == MDE_OK) rc = build_plan(c, e, false, &bytes);
  if (rc == MDE_OK) {
    bool any = false;
    for (const Op& op : c->plan) any = any || op.branch > 0;
    if (any) {       // side streams + events of the forked branches (non-blocking: they only ever synchronise through the events)
      cudaError_t err = cudaEventCreateWithFlags(&c->fork_ev, cudaEventDisableTiming);
      for (int b = 0; b < 3 && err == cudaSuccess; ++b) {
        err = cudaStreamCreateWithFlags(&c->side[b], cudaStreamNonBlocking);
        if (err == cudaSuccess) err = cudaEventCreateWithFlags(&c->join_ev[b], cudaEventDisableTiming);
      }
      if (err != cudaSuccess) rc = fail(MDE_ERR_CUDA, "side streams of the plan: %s", cudaGetErrorString(err));
    }
  }
  if (rc != MDE_OK) {
    std::string msg = mde_last_error();
    mde_context_destroy(c);
    return fail(rc, "%s", msg.c_str());
  }
  c->workspace_bytes = bytes;
  if (e->d.input_mode == MDE_INPUT_U8_HWC || e->d.output_mode == MDE_OUTPUT_SOURCE_GRID) { c->src_h = e->d.max_src_h; c->src_w = e->d.max_src_w; }
  *out = c;
  return MDE_OK;
}

extern "C" void mde_context_destroy(mde_context* c) {
  if (!c) return;
  if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
  for (void* p : c->allocs) cudaFree(p);
  for (cudaEvent_t ev : c->events) cudaEventDestroy(ev);
  for (int b = 0; b < 3; ++b) {
    if (c->side[b]) cudaStreamDestroy(c->side[b]);
    if (c->join_ev[b]) cudaEventDestroy(c->join_ev[b]);
  }
  if (c->fork_ev) cudaEventDestroy(c->fork_ev);
  delete c;
}

extern "C" int mde_context_set_tensor_address(mde_context* c, const char* name, void* d_ptr) {
  clear_error();
  if (!c || !name) return fail(MDE_ERR_INVALID, "bad argument to mde_context_set_tensor_address");
  if (!d_ptr) return fail(MDE_ERR_INVALID, "null device address for tensor '%s'", name);
  if (reinterpret_cast<uintptr_t>(d_ptr) & 15) return fail(MDE_ERR_INVALID, "tensor '%s' must be 16-byte aligned", name);
  const bool sky = c->e->d.head_mode == MDE_HEAD_DPT_EXP_SKY;
  if (!strcmp(name, sky ? "image" : "input")) c->d_input = d_ptr;
  else if (!strcmp(name, sky ? "depth" : "output")) c->d_output = d_ptr;
  else if (sky && !strcmp(name, "sky")) c->d_output2 = d_ptr;
  else return fail(MDE_ERR_INVALID, "engine has no tensor named '%s'", name);
  return MDE_OK;
}

extern "C" int mde_context_set_input_shape(mde_context* c, const char* name, int32_t ndim, const int64_t* dims) {
  clear_error();
  if (!c || !name || !dims) return fail(MDE_ERR_INVALID, "bad argument to mde_context_set_input_shape");
  if (strcmp(name, c->e->d.head_mode == MDE_HEAD_DPT_EXP_SKY ? "image" : "input")) return fail(MDE_ERR_INVALID, "engine has no input named '%s'", name);
  const mde_engine_desc& d = c->e->d;
  if (d.input_mode == MDE_INPUT_F32_NCHW) {
    if (ndim != 4 || dims[0] != d.batch || dims[1] != 3 || dims[2] != d.input_h || dims[3] != d.input_w)
      return fail(MDE_ERR_INVALID, "static engine: input shape must be [%d,3,%d,%d]", d.batch, d.input_h, d.input_w);
    return MDE_OK;
  }
  if (ndim != 4 || dims[0] != d.batch || dims[3] != 3) return fail(MDE_ERR_INVALID, "uint8 input shape must be [%d,src_h,src_w,3]", d.batch);
  if (dims[1] < 1 || dims[2] < 1 || dims[1] > d.max_src_h || dims[2] > d.max_src_w)
    return fail(MDE_ERR_INVALID, "source size %lldx%lld exceeds the engine's maximum %dx%d", (long long)dims[1], (long long)dims[2], d.max_src_h, d.max_src_w);
  c->src_h = static_cast<int>(dims[1]);
  c->src_w = static_cast<int>(dims[2]);
  return MDE_OK;
}

extern "C" int mde_context_snapshot_block(mde_context* c, int32_t block) {
  clear_error();
  if (!c) return fail(MDE_ERR_INVALID, "null context");
  if (block < -1 || block >= c->e->d.depth) return fail(MDE_ERR_INVALID, "block %d out of range", block);
  c->snapshot_block = block;
  return MDE_OK;
}

extern "C" int mde_context_launches_per_enqueue(const mde_context* c) {
  if (!c) return 0;
  int n = 0;
  for (const Op& op : c->plan) n += op.kind != Op::SNAPSHOT && op.kind != Op::JOIN;
  return n;
}

extern "C" int mde_context_get_buffer(mde_context* c, const char* name, void** d_ptr, int64_t* bytes, int32_t* dtype) {
  clear_error();
  if (!c || !name || !d_ptr || !bytes || !dtype) return fail(MDE_ERR_INVALID, "bad argument to mde_context_get_buffer");
  auto it = c->named.find(name);
  if (it == c->named.end()) return fail(MDE_ERR_INVALID, "context has no buffer named '%s'", name);
  *d_ptr = it->second.first;
  *bytes = it->second.second;
  *dtype = c->named_dtype[name];
  return MDE_OK;
}

extern "C" int mde_context_set_gather(mde_context* c, int32_t n_ranks, int32_t rank, void* const* d_peer_outputs) {
  clear_error();
  if (!c) return fail(MDE_ERR_INVALID, "null context");
  if (n_ranks == 0) { c->gather_ranks = 0; return MDE_OK; }
  if (c->e->d.head_mode != MDE_HEAD_ENCODER_TAPS) return fail(MDE_ERR_STATE, "the fused gather belongs to MDE_HEAD_ENCODER_TAPS engines");
  if (n_ranks < 1 || n_ranks > 8 || rank < 0 || rank >= n_ranks || !d_peer_outputs) return fail(MDE_ERR_INVALID, "gather: 1..8 ranks, 0 <= rank < n_ranks");
  for (int r = 0; r < n_ranks; ++r) {
    if (!d_peer_outputs[r] || (reinterpret_cast<uintptr_t>(d_peer_outputs[r]) & 15)) return fail(MDE_ERR_INVALID, "gather: buffer of rank %d is null or not 16-byte aligned", r);
    c->gather_dst[r] = d_peer_outputs[r];
  }
  c->gather_ranks = n_ranks;
  c->gather_rank = rank;
  return MDE_OK;
}

static int enqueue_impl(mde_context* c, cudaStream_t s, bool timed) {
  if (!c->d_input || (!c->d_output && c->gather_ranks == 0)) return fail(MDE_ERR_STATE, "set_tensor_address must be called for 'input' and 'output' before enqueue");
  if (c->e->d.head_mode == MDE_HEAD_DPT_EXP_SKY && !c->d_output2) return fail(MDE_ERR_STATE, "set_tensor_address must be called for 'sky' before enqueue");
  mde_engine* e = c->e;
  const mde_engine_desc& d = e->d;
  const int prec = d.precision;
  size_t ev = 0;
  cudaStream_t main_stream = s;
  bool forked = false;
  bool used[3] = {false, false, false};
  for (Op& op : c->plan) {
    s = main_stream;
    // per-op timing keeps everything on one stream (the events bracket single launches); so does a context without side streams
    const bool branches = !timed && c->side[0] != nullptr;
    if (op.kind == Op::JOIN) {          // the side branches rejoin the caller's stream: what follows reads their results
      if (branches && forked) {
        for (int b = 0; b < 3; ++b) {
          if (!used[b]) continue;
          MDE_CUDA_TRY(cudaEventRecord(c->join_ev[b], c->side[b]));
          MDE_CUDA_TRY(cudaStreamWaitEvent(main_stream, c->join_ev[b], 0));
          used[b] = false;
        }
        forked = false;
      }
      continue;
    }
    if (branches && op.branch > 0) {
      if (!forked) {       // everything enqueued so far happens before the branches
        MDE_CUDA_TRY(cudaEventRecord(c->fork_ev, main_stream));
        forked = true;
      }
      if (!used[op.branch - 1]) {
        MDE_CUDA_TRY(cudaStreamWaitEvent(c->side[op.branch - 1], c->fork_ev, 0));
        used[op.branch - 1] = true;
      }
      s = c->side[op.branch - 1];
    }
    if (timed && op.kind != Op::SNAPSHOT) MDE_CUDA_TRY(cudaEventRecord(c->events[ev++], s));
    switch (op.kind) {
      case Op::JOIN: break;
      case Op::PREPROC_U8:
        MDE_TRY(launch_preprocess_u8(prec, static_cast<const uint8_t*>(c->d_input), static_cast<long long>(c->src_h) * c->src_w * 3,
                                     d.batch, c->src_h, c->src_w, d.input_h, d.input_w, d.patch_size, e->kpad, d.swap_rb, e->lut,
                                     op.out, nullptr, s));
        break;
      case Op::IM2COL_F32:
        MDE_TRY(launch_im2col_f32(prec, static_cast<const float*>(c->d_input), d.batch, d.input_h, d.input_w, d.patch_size, e->kpad, op.out, s,
                                  (d.flags & MDE_FLAG_NORMALISE_F32) ? d.norm_mean : nullptr, (d.flags & MDE_FLAG_NORMALISE_F32) ? d.norm_std : nullptr));
        break;
      case Op::CLS_ROW:
        MDE_TRY(launch_cls_row(static_cast<float*>(op.out), e->cls, e->pos, e->reg, d.num_registers, d.batch, e->ntok, d.embed_dim, s));
        break;
      case Op::GEMM:
        if (op.g.p.head_w) op.g.p.head_out = static_cast<float*>(c->d_output);
        MDE_TRY(launch_gemm(op.g, s));
        break;
      case Op::LAYERNORM:
        if (op.i2 >= 0) {
          // trunk-only engine: tap op.i2 lands in slice op.i2 of the output binding, or of every rank's gather buffer
          const long long slice = static_cast<long long>(d.batch) * e->T * d.embed_dim;     // elements of one tap of one rank
          if (c->gather_ranks > 0) {
            void* dst[8];
            for (int r = 0; r < c->gather_ranks; ++r)
              dst[r] = static_cast<uint16_t*>(c->gather_dst[r]) + static_cast<long long>(op.i2) * c->gather_ranks * slice;
            MDE_TRY(launch_layernorm(prec, static_cast<const float*>(op.in), op.w, op.b, nullptr, op.rows, d.embed_dim, 1e-6f, op.i0,
                                     e->ntok, s, op.i1, c->gather_ranks, dst, static_cast<long long>(c->gather_rank) * d.batch * e->T));
          } else {
            MDE_TRY(launch_layernorm(prec, static_cast<const float*>(op.in), op.w, op.b,
                                     static_cast<uint16_t*>(c->d_output) + op.i2 * slice, op.rows, d.embed_dim, 1e-6f, op.i0, e->ntok, s, op.i1));
          }
          break;
        }
        MDE_TRY(launch_layernorm(prec, static_cast<const float*>(op.in), op.w, op.b, op.out, op.rows, d.embed_dim, 1e-6f, op.i0, e->ntok, s, op.i1));
        break;
      case Op::ATTENTION:
        MDE_TRY(launch_attention_op(op.attn, s));
        break;
      case Op::BILINEAR:
        MDE_TRY(launch_bilinear(prec, op.in, op.out, d.batch, op.i0, op.i1, op.i2, op.i3, op.i4, s));
        break;
      case Op::UPCONV_HEAD:
        if (op.i4 == 1)
          MDE_TRY(launch_upconv_head(prec, op.in, 288, d.batch, op.i0, op.i1, op.i2, op.i3, e->sky2_b, e->sky_head_w, e->sky_head_b, 0.f,
                                     static_cast<float*>(c->d_output2), s));
        else
          MDE_TRY(launch_upconv_head(prec, op.in, 288, d.batch, op.i0, op.i1, op.i2, op.i3, e->oc2_b, e->head_w, e->head_b,
                                     d.max_depth > 0.f ? d.max_depth : 0.f, static_cast<float*>(op.out ? op.out : c->d_output), s,
                                     d.head_mode == MDE_HEAD_DPT_EXP_SKY));
        break;
      case Op::RESIZE_DEPTH:
        MDE_TRY(launch_resize_depth(static_cast<const float*>(op.in), d.batch, d.input_h, d.input_w, static_cast<float*>(c->d_output),
                                    c->src_h, c->src_w, 1e-3f, 1e3f, s));
        break;
      case Op::IM2COL_S2:
        MDE_TRY(launch_im2col_s2(prec, op.in, op.out, d.batch, op.i0, op.i1, op.i2, s));
        break;
      case Op::SNAPSHOT:
        if (op.block == c->snapshot_block)
          MDE_CUDA_TRY(cudaMemcpyAsync(c->x_snapshot, c->x, static_cast<size_t>(c->x_bytes), cudaMemcpyDeviceToDevice, s));
        break;
    }
  }
  if (timed) MDE_CUDA_TRY(cudaEventRecord(c->events[ev], s));
  return MDE_OK;
}

extern "C" int mde_context_enqueue(mde_context* c, void* stream) {
  clear_error();
  if (!c) return fail(MDE_ERR_INVALID, "null context");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  EngineScope scope(c->e->d);
  MDE_TRY(scope.rc);
  const bool no_graph = (c->e->d.flags & MDE_FLAG_NO_GRAPH) != 0 || profile_mode();
  // the legacy default stream cannot be captured; a failed capture falls back to plain launches for good
  if (no_graph || c->graph_failed || s == nullptr || s == cudaStreamLegacy) return enqueue_impl(c, s, false);
  {
    // the caller is recording its own graph (depth_pro.py captures the three trunks, the decoder and both heads in one): the
    // launches go straight into that capture
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &st) == cudaSuccess && st != cudaStreamCaptureStatusNone) return enqueue_impl(c, s, false);
    cudaGetLastError();
  }
  if (!c->d_input || (!c->d_output && c->gather_ranks == 0)) return fail(MDE_ERR_STATE, "set_tensor_address must be called for 'input' and 'output' before enqueue");
  unsigned long long key[8] = {reinterpret_cast<unsigned long long>(c->d_input), reinterpret_cast<unsigned long long>(c->d_output),
                               static_cast<unsigned long long>(c->src_h), static_cast<unsigned long long>(c->src_w),
                               static_cast<unsigned long long>(c->snapshot_block + 1), static_cast<unsigned long long>(c->gather_ranks),
                               static_cast<unsigned long long>(c->gather_rank), reinterpret_cast<unsigned long long>(c->d_output2)};
  for (int r = 0; r < c->gather_ranks; ++r) key[7] = key[7] * 1000003ull + reinterpret_cast<unsigned long long>(c->gather_dst[r]);
  if (!c->graph_exec || memcmp(key, c->graph_key, sizeof(key)) != 0) {
    if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }
    if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
      cudaGetLastError();
      c->graph_failed = true;
      return enqueue_impl(c, s, false);
    }
    const int rc = enqueue_impl(c, s, false);
    cudaGraph_t graph = nullptr;
    const cudaError_t end = cudaStreamEndCapture(s, &graph);
    if (rc != MDE_OK || end != cudaSuccess || !graph) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      c->graph_failed = true;
      if (rc != MDE_OK) return rc;
      return enqueue_impl(c, s, false);
    }
    const cudaError_t inst = cudaGraphInstantiate(&c->graph_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (inst != cudaSuccess) {
      cudaGetLastError();
      c->graph_exec = nullptr;
      c->graph_failed = true;
      return enqueue_impl(c, s, false);
    }
    memcpy(c->graph_key, key, sizeof(key));
  }
  MDE_CUDA_TRY(cudaGraphLaunch(c->graph_exec, s));
  return MDE_OK;
}

// Profiling variant: CUDA events between the launches, then a stream synchronise; ms[i] is the device time
// of launch i of the plan (mde_context_launches_per_enqueue entries).
extern "C" int mde_context_enqueue_timed(mde_context* c, void* stream, float* ms, int32_t capacity) {
  clear_error();
  if (!c || !ms) return fail(MDE_ERR_INVALID, "bad argument to mde_context_enqueue_timed");
  const int n = mde_context_launches_per_enqueue(c);
  if (capacity < n) return fail(MDE_ERR_INVALID, "ms[] holds %d entries, %d needed", capacity, n);
  while (static_cast<int>(c->events.size()) < n + 1) {
    cudaEvent_t ev;
    MDE_CUDA_TRY(cudaEventCreate(&ev));
    c->events.push_back(ev);
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  EngineScope scope(c->e->d);
  MDE_TRY(scope.rc);
  {
    // per-launch times need launches that do not overlap: no programmatic dependent launch while the events are recorded
    LaunchOpts o = launch_opts();
    o.pdl = false;
    set_launch_opts(o);
  }
  MDE_TRY(enqueue_impl(c, s, true));
  MDE_CUDA_TRY(cudaStreamSynchronize(s));
  for (int i = 0; i < n; ++i) MDE_CUDA_TRY(cudaEventElapsedTime(&ms[i], c->events[i], c->events[i + 1]));
  return MDE_OK;
}

// Label, algorithmic FLOPs and algorithmic bytes of launch i of the plan.
extern "C" int mde_context_op_info(const mde_context* c, int32_t i, char* label, int32_t label_capacity, double* flops, double* bytes) {
  clear_error();
  if (!c || !label || label_capacity < 1 || !flops || !bytes) return fail(MDE_ERR_INVALID, "bad argument to mde_context_op_info");
  int k = 0;
  for (const Op& op : c->plan) {
    if (op.kind == Op::SNAPSHOT || op.kind == Op::JOIN) continue;
    if (k++ == i) {
      snprintf(label, static_cast<size_t>(label_capacity), "%s", op.label.c_str());
      *flops = op.flops;
      *bytes = op.bytes;
      return MDE_OK;
    }
  }
  return fail(MDE_ERR_INVALID, "launch index %d out of range", i);
}
