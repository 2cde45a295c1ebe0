// Host-side helpers shared by the C-ABI translation units: error reporting, the driver entry point for
// tensor-map encoding, GEMM/conv op construction and kernel launchers.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/mde_b200.h"
#include "gemm.cuh"

namespace mde {

int fail(int code, const char* fmt, ...);
void clear_error();

#define MDE_CUDA_TRY(expr)                                                                   \
  do {                                                                                       \
    cudaError_t e__ = (expr);                                                                \
    if (e__ != cudaSuccess) return ::mde::fail(MDE_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)
#define MDE_TRY(expr)            \
  do {                           \
    int r__ = (expr);            \
    if (r__ != MDE_OK) return r__; \
  } while (0)

int num_sms();   // SM count of the current device (cached per device), 0 on failure

// Per-thread launch options.  The engine sets them from mde_engine_desc.flags around plan building and enqueue; the
// single-kernel entry points (mde_k_*) run with the defaults.
struct LaunchOpts {
  bool pdl = false;       // programmatic dependent launch for the kernels that call griddep_wait()
  bool split_k = false;   // small-batch split-K of the residual GEMMs (arrival-order fp32 adds in the L2)
};
void set_launch_opts(const LaunchOpts& o);
LaunchOpts launch_opts();

// One tensor-core GEMM / implicit-GEMM conv launch, fully described at plan-build time.
struct GemmOp {
  alignas(64) CUtensorMap map_a;
  alignas(64) CUtensorMap map_b;
  alignas(64) CUtensorMap map_out;   // 16-bit output as {N, M} with 64 x 32 boxes; only encoded when p.tma_out
  alignas(64) GatherMaps gather;     // p.gather_n destinations of the fused all-gather
  GemmParams p;
  int block_n;
  int precision;
  int grid;
  int ctas;       // 2: launched as clusters of two CTAs working on 256-row tiles (tcgen05 cta_group::2)
};

// D[M,N] = A[M,K] B[N,K]^T
int make_gemm_op(GemmOp* op, int precision, const void* d_a, long long m, int k, int lda, const void* d_b, int n,
                 int ldb, const mde_epilogue* ep);
// 3x3 s1 p1 NHWC conv, weights [cout][9*cin_pad]
int make_conv_op(GemmOp* op, int precision, const void* d_in, int batch, int h, int w, int cin, const void* d_w,
                 int cout, const mde_epilogue* ep);
int launch_gemm(const GemmOp& op, cudaStream_t stream);

// tcgen05 attention launch, described at plan-build time (tensor map over the packed q|k|v rows)
struct AttnOp {
  alignas(64) CUtensorMap map_qkv;   // 128-row boxes: query tiles
  alignas(64) CUtensorMap map_kv128; // key/value source with 128-row boxes (the same tensor as map_qkv for self-attention)
  alignas(64) CUtensorMap map_kv96;  // the same source with 96-row boxes (attention_q3.cuh)
  alignas(64) CUtensorMap map_out3;  // the output as {D, queries per image, images} with 64 x 32 x 1 boxes (attention_q3.cuh's bulk stores)
  const void* qkv;
  void* out;
  int batch, ntok, heads, precision;
  int ntok_q, k_col0, v_col0;        // queries per image; first K / V column of head 0 in the key/value source
  unsigned int* counters;            // attention_q3 only: {next item, CTAs done}, zero-initialised and owned by whoever owns the op (an
                                     // engine context: one per attention op of its plan); NULL = static round-robin over the CTAs
  int kind;                          // 0: two CTAs per SM, one query tile each (attention_tc.cuh); 1: one persistent CTA per SM, three query
                                     // tiles (attention_q3.cuh); make_* picks by the number of work items
  int poly;                          // eighths of the exponentials evaluated on the FMA pipe (0..4; make_* sets the default, 2)
};
int make_attention_op(AttnOp* op, int precision, const void* d_qkv, void* d_out, int batch, int ntok, int heads);
// queries from d_q ([batch*ntok_q][ldq], q columns first), keys/values from d_kv ([batch*ntok_kv][ldkv], K at k_col0, V at v_col0)
int make_attention_op_kv(AttnOp* op, int precision, const void* d_q, int ldq, const void* d_kv, int ldkv, int k_col0, int v_col0,
                         void* d_out, int batch, int ntok_q, int ntok_kv, int heads);
int launch_attention_op(const AttnOp& op, cudaStream_t s);
// warp-level mma.sync variant (kept as an independent cross-check of the tcgen05 kernel in the tests)
int launch_attention_mma(int precision, const void* d_qkv, void* d_out, int batch, int ntok, int heads, cudaStream_t s);
int launch_layernorm(int precision, const float* d_x, const float* d_w, const float* d_b, void* d_out, long long rows,
                     int dim, float eps, int drop_cls, int ntok, cudaStream_t s, int identity = 0, int n_dst = 0,
                     void* const* dst = nullptr, long long dst_row0 = 0);
int launch_bilinear(int precision, const void* d_in, void* d_out, int batch, int hi, int wi, int ho, int wo, int c,
                    cudaStream_t s, const float* d_addend = nullptr);
int launch_im2col_s2(int precision, const void* d_in, void* d_out, int batch, int h, int w, int c, cudaStream_t s);
int launch_im2col_f32(int precision, const float* d_nchw, int batch, int h, int w, int patch, int kpad, void* d_cols,
                      cudaStream_t s, const double* mean3 = nullptr, const double* std3 = nullptr);
int launch_preprocess_u8(int precision, const uint8_t* d_src, long long src_batch_stride, int batch, int src_h,
                         int src_w, int dst_h, int dst_w, int patch, int kpad, int swap_rb, const float* d_lut,
                         void* d_cols, float* d_nchw, cudaStream_t s, int keep_ratio_pad = 0, const double* pad_rgb = nullptr);
// bilinear upsample -> 3x3 conv (tap-contracted z, see upconv_head.cuh) -> ReLU -> 1x1 -> activation
int launch_upconv_head(int precision, const void* d_z, int ldz, int batch, int hs, int ws, int ho, int wo,
                       const float* d_bias, const float* d_head_w, float head_b, float head_scale, float* d_out,
                       cudaStream_t s, int head_exp = 0);
int launch_resize_depth(const float* d_in, int batch, int hi, int wi, float* d_out, int ho, int wo, float lo, float hi_clamp,
                        cudaStream_t s);
int launch_merge_patches(int precision, const void* d_in, int per_side, int grid, int pad, int dim, void* d_out, cudaStream_t s);
int launch_cls_row(float* d_x, const float* d_cls, const float* d_pos, const float* d_reg, int n_reg, int batch, int ntok, int dim,
                   cudaStream_t s);
// (v/255 - mean)/std in double, rounded once to float32: 3 x 256 entries (core/preprocess.py:294-328,337-342)
void build_norm_lut(const double* mean3, const double* std3, float* lut768, bool scale_f32 = false);

}  // namespace mde
