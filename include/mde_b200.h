/*
 * mde_b200.h -- C ABI of libmde_b200.so, the B200 (sm_100a) replacement for the TensorRT engine
 * path of yester31/Monocular_Depth_Estimation_TRT.
 *
 * What each group replaces in the reference (paths relative to the reference checkout):
 *
 *   mde_engine_*    the tensorrt.ICudaEngine that core/common.py:141-312 `get_engine` builds or
 *                   loads, and the engine surface core/common_runtime.py:136-171 touches
 *                   (num_io_tensors, get_tensor_name/shape/dtype/mode).  "Build" here is
 *                   weights + description -> packed device weights; there is no ONNX and no
 *                   multi-minute tactic search.
 *   mde_context_*   the tensorrt.IExecutionContext used by core/common_runtime.py:268-275
 *                   `do_inference` (set_tensor_address, execute_async_v3) and
 *                   models/depth_anything_v2/onnx2trt.py:99-100 (set_input_shape).
 *   mde_k_*         single-kernel entry points, the test surface for tests/ -m gpu.
 *
 * Conventions: every function returns 0 on success and a non-zero code otherwise;
 * mde_last_error() returns a thread-local message for the last failure (the Python shim raises
 * RuntimeError with it, like core/common_runtime.py:41-56 `cuda_call`).  No exceptions cross the
 * ABI.  All pointers named d_* are device pointers owned by the caller; `stream` is a
 * cudaStream_t passed as void*.  mde_context_enqueue never allocates and never synchronises.
 * There is no CPU fallback anywhere: without a B200 every compute entry point fails.
 */
#ifndef MDE_B200_H
#define MDE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDE_ABI_VERSION 2

/* operand precision of the tensor-core path (accumulation is always fp32) */
#define MDE_FP16 0
#define MDE_BF16 1

/* input binding */
#define MDE_INPUT_F32_NCHW 0 /* the reference's contract: float32 [B,3,H,W] from core/preprocess.py */
#define MDE_INPUT_U8_HWC 1   /* uint8 [B,src_h,src_w,3] source images; resize+normalise+im2col fused on the GPU */

/* tensor dtypes reported by mde_engine_io_dtype */
#define MDE_DT_F32 0
#define MDE_DT_U8 1
#define MDE_DT_F16 2
#define MDE_DT_BF16 3

#define MDE_OUTPUT_MODEL_GRID 0
#define MDE_OUTPUT_SOURCE_GRID 1

/* what the engine computes behind the shared DINOv2 trunk */
#define MDE_HEAD_DPT 0          /* Depth Anything V2: DPT head, output float32 depth [B,H,W] */
#define MDE_HEAD_DPT_EXP_SKY 2  /* Depth Anything V3 (models/depth_anything_v3/onnx_export.py:31-55, reports/profile/depth_anything_v3.json
                                 * layers 221-227): the same DPT head ending in exp(), plus a parallel sky branch
                                 * (`sky_output_conv2`: conv3x3 -> ReLU -> conv1x1 -> ReLU) on the same up-sampled features.
                                 * Bindings: "image" in, "depth" and "sky" out, float32 [B][H][W] each */
#define MDE_HEAD_ENCODER_TAPS 1 /* trunk only (the patch-encoder stage of Depth Pro, models/depth_pro/onnx_export.py:15-22):
                                 * output = the four tapped block outputs, cls dropped, 16-bit [4][B][T][D] */

/* mde_engine_desc.flags: everything that changes which kernels run or how they are launched is part of the engine
 * description (and of the fingerprint get_engine writes) -- nothing is read from the environment except MDE_PROFILE=1,
 * which launches without PDL and without graph replay so that Nsight Compute lists every kernel (numerics unchanged). */
#define MDE_FLAG_SPLIT_K 1  /* small batches: split K of the residual GEMMs over idle SMs; the partial products meet in the
                             * L2's fp32 adds in arrival order, i.e. results are no longer bitwise reproducible run to run */
#define MDE_FLAG_NO_PDL 2   /* launch without programmatic dependent launch (default: on for GEMM / attention / LayerNorm) */
#define MDE_FLAG_NO_GRAPH 4 /* always launch kernel by kernel instead of replaying the captured CUDA graph */
#define MDE_FLAG_NORMALISE_F32 16 /* MDE_INPUT_F32_NCHW: the graph's own first op (x - norm_mean) / norm_std, in the units of the binding
                                  * (VGGT: 0..1 images, reports/profile/vggt.json layer 2; Metric3D V2: 0..255, metric3d_v2.json layers
                                  * 0-2), applied in fp32 before the 16-bit rounding of the patch rows */
#define MDE_FLAG_SCALE_F32 8 /* MDE_INPUT_U8_HWC: v / 255 evaluated in float32 before the float64 (v - mean) / std -- depth_anything_ac's
                              * input contract (core/preprocess.py:294-305, :470-476); default: float64 throughout (depth_anything_v2) */

/* error codes */
#define MDE_OK 0
#define MDE_ERR_INVALID 1  /* bad argument / unsupported configuration */
#define MDE_ERR_CUDA 2     /* a CUDA call failed (no device, launch failure, ...) */
#define MDE_ERR_STATE 3    /* call order violated (e.g. enqueue before all addresses are set) */
#define MDE_ERR_IO 4       /* weights file unreadable / malformed */
#define MDE_ERR_MISSING 5  /* a required weight tensor was not provided */

typedef struct mde_engine mde_engine;
typedef struct mde_context mde_context;

/* DINOv2 ViT encoder + DPT head (Depth Anything V2 family).  Mirrors the constructor arguments at
 * models/depth_anything_v2/infer.py:55-62 and infer_metric.py:53-65. */
typedef struct mde_engine_desc {
  int32_t struct_size;     /* sizeof(mde_engine_desc), for ABI evolution */
  int32_t embed_dim;       /* 384 / 768 / 1024 */
  int32_t depth;           /* 12 / 12 / 24 */
  int32_t num_heads;       /* 6 / 12 / 16 (head dim must be 64) */
  int32_t patch_size;      /* 14 */
  int32_t features;        /* 64 / 128 / 256 */
  int32_t out_channels[4]; /* e.g. 256,512,1024,1024 */
  int32_t taps[4];         /* block indices whose outputs feed the head, e.g. 4,11,17,23 */
  int32_t input_h, input_w;/* model input size, multiples of patch_size */
  int32_t batch;           /* images per enqueue */
  int32_t precision;       /* MDE_FP16 | MDE_BF16 */
  int32_t input_mode;      /* MDE_INPUT_* */
  int32_t max_src_h, max_src_w; /* MDE_INPUT_U8_HWC: largest source image the input binding must hold */
  int32_t swap_rb;         /* MDE_INPUT_U8_HWC: source is BGR (cv2.imread), network wants RGB */
  double norm_mean[3];     /* MDE_INPUT_U8_HWC: (v/255 - mean)/std, evaluated in double like */
  double norm_std[3];      /*   core/preprocess.py:294-328 with float64 dtypes */
  float max_depth;         /* > 0: metric head, sigmoid * max_depth;  <= 0: relative head, ReLU */
  int32_t device;          /* CUDA device ordinal */
  int32_t output_mode;     /* MDE_OUTPUT_MODEL_GRID: float32 [B][input_h][input_w] (spec.json's contract);
                            * MDE_OUTPUT_SOURCE_GRID: the scripts' post-processing fused in -- float32 [B][src_h][src_w],
                            * resized back to the source size (align_corners=True) and clamped to [1e-3, 1e3].  The source
                            * size is the one bound with set_input_shape (uint8 input) or max_src_h x max_src_w */
  int32_t head_mode;       /* MDE_HEAD_* */
  int32_t tap_norm_mask;   /* bit i: tap i goes through the trunk's final LayerNorm (0xF for Depth Anything; 0x8 for
                            * Depth Pro's hooks, which take raw block outputs and normalise only the last one) */
  int32_t num_registers;   /* DINOv2 "with registers" trunks (Metric3D V2, VGGT): learned tokens between cls and the patch tokens
                            * ("pretrained.register_tokens" [1, R, D]), dropped again from the taps; 0 for Depth Anything / Depth Pro */
  int32_t flags;           /* MDE_FLAG_* */
  int32_t attn_poly;       /* eighths of the softmax exponentials evaluated by a polynomial on the FMA pipe instead of the
                            * SFU: 0..4, or -1 for the library default (2).  Changes the last bits of the probabilities. */
} mde_engine_desc;

const char* mde_last_error(void);
int mde_abi_version(void);

/* ---- engine ------------------------------------------------------------------------------ */
int mde_engine_create(const mde_engine_desc* desc, mde_engine** out);
/* Provide one tensor of the upstream state dict (key names as in depth_anything_v2_*.pth), fp32,
 * host memory, C-contiguous.  Copied; the caller may free `data` on return. */
int mde_engine_set_weight(mde_engine* e, const char* name, const float* data, int32_t ndim, const int64_t* dims);
/* Read every tensor from a .mdew file (see monocular_depth_estimation_trt_b200/weights.py). */
int mde_engine_load_weights(mde_engine* e, const char* path);
/* Check that every required tensor is present, pack to the kernels' layouts, upload.  Needs the GPU. */
int mde_engine_finalize(mde_engine* e);
void mde_engine_destroy(mde_engine* e);

int mde_engine_num_io(const mde_engine* e);
const char* mde_engine_io_name(const mde_engine* e, int32_t i);
/* dims[] receives up to 8 extents, *ndim their count (maximum extents for the uint8 input). */
int mde_engine_io_shape(const mde_engine* e, int32_t i, int32_t* ndim, int64_t* dims);
int mde_engine_io_dtype(const mde_engine* e, int32_t i); /* MDE_DT_* or -1 */
int mde_engine_io_is_input(const mde_engine* e, int32_t i); /* 1 input, 0 output, -1 bad index */
/* bytes of device workspace a context of this engine allocates */
int64_t mde_engine_workspace_bytes(const mde_engine* e);

/* ---- execution context -------------------------------------------------------------------- */
int mde_context_create(mde_engine* e, mde_context** out);
void mde_context_destroy(mde_context* c);
int mde_context_set_tensor_address(mde_context* c, const char* name, void* d_ptr);
/* MDE_INPUT_U8_HWC only: actual source size of this batch, dims = {B, src_h, src_w, 3}. */
int mde_context_set_input_shape(mde_context* c, const char* name, int32_t ndim, const int64_t* dims);
/* Launch one forward of the whole batch on `stream`.  Asynchronous. */
int mde_context_enqueue(mde_context* c, void* stream);
/* MDE_HEAD_ENCODER_TAPS engines, one process per GPU: fuse the all-gather of the taps into the kernel that produces
 * them.  d_peer_outputs[r] is rank r's gather buffer, 16-bit [4][n_ranks * B][T][D], mapped into this process (own
 * buffer for r == rank, cudaIpcOpenMemHandle'd peers otherwise); this rank's B images are written at image offset
 * rank * B of every buffer with plain stores over NVLink.  The "output" binding is then unused.  n_ranks == 0 turns
 * it off.  The caller orders the ranks (a barrier after the stream has drained) before anyone reads a buffer. */
int mde_context_set_gather(mde_context* c, int32_t n_ranks, int32_t rank, void* const* d_peer_outputs);
/* Number of kernels one enqueue launches (bench.py's gpu_launches). */
int mde_context_launches_per_enqueue(const mde_context* c);
/* Profiling variant of enqueue: CUDA events between the launches, then a stream synchronise.
 * ms[i] = device time of launch i (capacity >= mde_context_launches_per_enqueue). */
int mde_context_enqueue_timed(mde_context* c, void* stream, float* ms, int32_t capacity);
/* Label ("gemm256 fc1+gelu 87680x4096x1024", ...), algorithmic FLOPs (2*MACs, logical sizes) and
 * algorithmic bytes (operands read once, results written once) of launch i. */
int mde_context_op_info(const mde_context* c, int32_t i, char* label, int32_t label_capacity, double* flops, double* bytes);
/* Intermediate tensors for the per-stage parity gates: "cols", "x" (residual stream after the
 * last block), "tap0".."tap3", "path_1", "r0".."r3".  Valid after enqueue + stream sync.
 * dtype: 0 = fp32, 1 = 16-bit in the engine's precision. */
int mde_context_get_buffer(mde_context* c, const char* name, void** d_ptr, int64_t* bytes, int32_t* dtype);
/* Snapshot the fp32 residual stream after block `block` into an internal buffer during enqueue
 * (-1 disables).  Read it back with mde_context_get_buffer(c, "x_snapshot", ...). */
int mde_context_snapshot_block(mde_context* c, int32_t block);

/* ---- single kernels (test surface) ---------------------------------------------------------- */

/* Kernel (1): uint8 HWC -> cv2-exact bilinear resize -> normalise -> patch im2col.
 * d_cols: [B*(dst_h/patch)*(dst_w/patch)][kpad] 16-bit (may be NULL); d_nchw: float32 [B,3,dst_h,dst_w]
 * (may be NULL).  kpad >= 3*patch*patch, multiple of 64. */
int mde_k_preprocess_u8(int32_t precision, const uint8_t* d_src, int32_t batch, int32_t src_h, int32_t src_w,
                        int32_t dst_h, int32_t dst_w, int32_t patch, int32_t kpad, int32_t swap_rb,
                        const double* mean3, const double* std3, void* d_cols, float* d_nchw, void* stream);
/* The keep-ratio + pad form (core/preprocess.py:191-219 `resize_pad`, rounding='trunc', centred: MODELS['metric3d_v2'],
 * :487-491): INTER_LINEAR resize to (int(src_h*scale), int(src_w*scale)), scale = min(dst_h/src_h, dst_w/src_w), centred on
 * a dst_h x dst_w canvas of pad_rgb3 (saturate-cast to uint8 like cv2.copyMakeBorder).  mean3 == std3 == NULL: no
 * normalisation, values stay in 0..255 (what metric3d_v2's input binding carries). */
int mde_k_preprocess_u8_pad(int32_t precision, const uint8_t* d_src, int32_t batch, int32_t src_h, int32_t src_w,
                            int32_t dst_h, int32_t dst_w, int32_t patch, int32_t kpad, int32_t swap_rb,
                            const double* pad_rgb3, const double* mean3, const double* std3, void* d_cols, float* d_nchw,
                            void* stream);
/* VGGT / StreamVGGT input contract (core/preprocess.py:222-265 `resize_square_pad`, symmetric, :493-498 `vggt`): white square
 * pad at source resolution, ONE cv2.INTER_CUBIC resize (OpenCV's own 8-bit path, bit-exact: oracle/preprocess_np.py
 * `resize_cubic_u8`), / 255 in float32.  d_src: uint8 [B][src_h][src_w][3]; d_nchw: float32 [B][3][dst_h][dst_w] (the rank-5
 * binding [1, S, 3, H, W] of models/vggt/spec.json with B = S frames). */
int mde_k_preprocess_u8_square_pad_cubic(const uint8_t* d_src, int32_t batch, int32_t src_h, int32_t src_w, int32_t dst_h, int32_t dst_w,
                                         int32_t swap_rb, int32_t pad_value, float* d_nchw, void* stream);
/* Depth-Anything-AC `native` profile (core/preprocess.py:470-476 `da_ac(h, w, stretch=False)`, models/depth_anything_ac/
 * onnx2trt.py:50-75): no uint8 resize; float32 / 255; cv2's FLOAT INTER_CUBIC (OpenCV's own path, bit-exact:
 * oracle/preprocess_np.py `resize_cubic_f32`) to dst_h x dst_w -- the keep-ratio "ceil" size, 480 x 640 -> 518 x 700 --
 * then (x - mean) / std in float64 (mean3 / std3 of the OUTPUT channels; both NULL: the resized 0..1 image).
 * d_src: uint8 [B][src_h][src_w][3]; d_nchw: float32 [B][3][dst_h][dst_w]. */
int mde_k_preprocess_u8_cubic_f32(const uint8_t* d_src, int32_t batch, int32_t src_h, int32_t src_w, int32_t dst_h, int32_t dst_w,
                                  int32_t swap_rb, const double* mean3, const double* std3, float* d_nchw, void* stream);
int mde_k_im2col_f32(int32_t precision, const float* d_nchw, int32_t batch, int32_t h, int32_t w, int32_t patch,
                     int32_t kpad, void* d_cols, void* stream);

/* Epilogue of the tensor-core GEMM / conv:  v = acc + bias; v = act(v); v *= gamma; v += pos;
 * v += res1 + res2; x = (accumulate_x ? x : 0) + v; out = v; out_relu = relu(v). */
typedef struct mde_epilogue {
  const float* d_bias;   /* [N] or NULL */
  const float* d_gamma;  /* [N] or NULL */
  int32_t act;           /* 0 none, 1 GELU(erf), 2 ReLU */
  float* d_x;            /* fp32 [rows][ld_out] or NULL */
  int32_t accumulate_x;
  const void* d_res1;    /* 16-bit [rows][ld_out] or NULL */
  const void* d_res2;
  void* d_out;           /* 16-bit [rows][ld_out] or NULL */
  void* d_out_relu;
  int32_t ld_out;
  /* token remap (patch embed): row b*T+t -> b*(T+1)+1+t, plus pos[(1+t)][:] */
  int32_t tokens;        /* 0 = off */
  const float* d_pos;
  /* pixel shuffle (ConvTranspose2d with kernel == stride == s): A rows are (b,y,x) of an HxW map,
   * N = s*s*cout, column (ky*s+kx)*cout + o goes to pixel (s*y+ky, s*x+kx), channel o */
  int32_t shuffle_s, shuffle_cout, shuffle_h, shuffle_w; /* shuffle_s == 0: off */
  /* fused depth head (N == 32): z = sum_n relu(v_n) * head_w[n] + head_b;
   * head_out[row] = head_scale > 0 ? head_scale*sigmoid(z) : relu(z) */
  const float* d_head_w;
  float head_b, head_scale;
  float* d_head_out;
  /* GEMM -> all-gather fused (plain 16-bit outputs only): columns [gather_col0, N) are written to d_gather[0..gather_n)
   * -- every rank's gathered buffer, ALREADY offset to this rank's first row, pitch gather_ld elements, column
   * n - gather_col0 -- instead of d_out.  Peer buffers are cudaIpc-mapped device pointers.  gather_n == 0: off.
   * gather_col0 must be a multiple of 64 and N a multiple of the tile width. */
  int32_t gather_n, gather_col0, gather_ld;
  void* d_gather[8];
  /* token remap with more rows in front of the patch tokens than the cls row (DINOv2 with registers): row b*T+t ->
   * b*(T+token_skip)+token_skip+t, plus pos[(1+t)][:].  0 means 1 (cls only). */
  int32_t token_skip;
  /* fused depth head: 0 = head_scale decides (sigmoid * scale / ReLU), 1 = exp(z) (VGGT's depth head, activation "exp") */
  int32_t head_act;
} mde_epilogue;

/* D[M,N] = A[M,K] * B[N,K]^T.  A: 16-bit row-major, pitch lda elements; B: 16-bit row-major, pitch ldb.
 * lda, ldb multiples of 8; N multiple of 8. */
int mde_k_gemm(int32_t precision, const void* d_a, int64_t m, int32_t k, int32_t lda, const void* d_b, int32_t n,
               int32_t ldb, const mde_epilogue* ep, void* stream);
/* Tuning probe (tools/gemm_tiling_probe.py): mde_k_gemm with the tile width (32 / 64 / 128 / 256 columns), the CTA pairing
 * (1 / 2) and the split-K count (1..16; only residual-reduction epilogues split) forced instead of picked. */
int mde_k_gemm_tiled(int32_t precision, const void* d_a, int64_t m, int32_t k, int32_t lda, const void* d_b, int32_t n,
                     int32_t ldb, const mde_epilogue* ep, int32_t block_n, int32_t ctas, int32_t splits, void* stream);
/* 3x3 / stride 1 / pad 1 convolution, NHWC.  d_in: [B][H][W][cin] 16-bit (cin multiple of 8);
 * d_w: [cout][9*cin_pad] 16-bit with cin_pad = cin rounded up to 64 and K index = (ky*3+kx)*cin_pad + c. */
int mde_k_conv3x3(int32_t precision, const void* d_in, int32_t batch, int32_t h, int32_t w, int32_t cin,
                  const void* d_w, int32_t cout, const mde_epilogue* ep, void* stream);
/* softmax(Q K^T / 8) V over [B*ntok][3*D] packed q|k|v rows, head dim 64 -> [B*ntok][D]. */
int mde_k_attention(int32_t precision, const void* d_qkv, void* d_out, int32_t batch, int32_t ntok, int32_t heads,
                    void* stream);
/* Queries and keys/values from different row sets (sequence-sharded global attention, SURVEY section 8 e row 3: the
 * queries are this rank's tokens, the keys/values all ranks' tokens gathered into one buffer).
 * d_q: [batch*ntok_q][ldq] with the q columns first; d_kv: [batch*ntok_kv][ldkv] with K of head 0 at column k_col0 and
 * V at v_col0; d_out: [batch*ntok_q][heads*64]. */
int mde_k_attention_kv(int32_t precision, const void* d_q, int32_t ldq, const void* d_kv, int32_t ldkv, int32_t k_col0,
                       int32_t v_col0, void* d_out, int32_t batch, int32_t ntok_q, int32_t ntok_kv, int32_t heads, void* stream);
/* The same op forced onto the one-query-tile-per-CTA kernel (csrc/attention_tc.cuh) with an explicit share of polynomial
 * exponentials (poly_eighths 0..4, -1: the library default; tools/attn_sweep.py sweeps it). */
int mde_k_attention_poly(int32_t precision, const void* d_qkv, void* d_out, int32_t batch, int32_t ntok, int32_t heads,
                         int32_t poly_eighths, void* stream);
/* The same op forced onto the three-query-tiles-per-CTA persistent kernel (csrc/attention_q3.cuh) whatever the problem size;
 * mde_k_attention picks between the two by the number of work items.  poly_eighths as above. */
int mde_k_attention_q3(int32_t precision, const void* d_qkv, void* d_out, int32_t batch, int32_t ntok, int32_t heads,
                       int32_t poly_eighths, void* stream);
/* The default attention kernel with clock64 stamps of the softmax warps' phases (profiling aid, tools/attn_trace.py):
 * d_trace int64 [2048 CTAs][4 warps][64 slots], zero-initialised by the caller; slot 0 kernel entry, 1 after the prologue sync,
 * then per key tile: S available, S in registers, exponentials done, previous P V done, P stored; then O available, stored. */
int mde_k_attention_trace(int32_t precision, const void* d_qkv, void* d_out, int32_t batch, int32_t ntok, int32_t heads,
                          int64_t* d_trace, void* stream);
/* The same op on warp-level mma.sync tensor-core instructions: an independent cross-check of the
 * tcgen05 kernel for the tests; the engine never launches it. */
int mde_k_attention_mma(int32_t precision, const void* d_qkv, void* d_out, int32_t batch, int32_t ntok, int32_t heads,
                        void* stream);
/* LayerNorm of fp32 rows -> 16-bit.  drop_cls = n > 0: rows are [B][ntok]; the first n tokens of each image (cls, or cls +
 * registers, or VGGT's camera + register tokens) are skipped and the output is dense [B][ntok-n].  dim: 128, 384, 768, 1024,
 * 1536 or 2048. */
int mde_k_layernorm(int32_t precision, const float* d_x, const float* d_w, const float* d_b, void* d_out,
                    int64_t rows, int32_t dim, float eps, int32_t drop_cls, int32_t ntok, void* stream);
/* bilinear, align_corners=True, NHWC 16-bit, c multiple of 8 */
int mde_k_bilinear(int32_t precision, const void* d_in, void* d_out, int32_t batch, int32_t hi, int32_t wi, int32_t ho,
                   int32_t wo, int32_t c, void* stream);
/* mde_k_bilinear with a per-pixel, per-channel addend (fp32 [ho][wo][c], the same for every image) added after the
 * interpolation: VGGT's depth head adds its position embedding to the up-sampled map (reports/profile/vggt.json layers 705-706). */
int mde_k_bilinear_add(int32_t precision, const void* d_in, void* d_out, int32_t batch, int32_t hi, int32_t wi, int32_t ho,
                       int32_t wo, int32_t c, const float* d_addend, void* stream);
/* VGGT's aggregator input (models/vggt/onnx_export.py:38-52 -> vggt `Aggregator.forward`): per frame the camera token and the
 * four register tokens (d_special: float32 [2][n_special][dim]; global frame 0 takes variant 0, every other frame variant 1)
 * followed by the trunk's normalised patch tokens (d_patch: 16-bit [frames][tokens][dim]) -> float32
 * [frames][n_special + tokens][dim].  first_frame: global index of this rank's first frame. */
int mde_k_assemble_tokens(int32_t precision, const void* d_patch, const float* d_special, int32_t frames, int32_t tokens,
                          int32_t n_special, int32_t dim, int32_t first_frame, float* d_out, void* stream);
/* Tail of the DPT head (models/depth_anything_v2: output_conv1 -> interpolate(align_corners=True) -> output_conv2)
 * with the 3x3 conv's channel contraction done before the interpolation.  d_z: [B][hs][ws][ldz] 16-bit with
 * z[.., (ky*3+kx)*32 + o] = sum_c W2[o][c][ky][kx] * o1[.., c] (no bias); d_bias: [32] conv bias; d_head_w: [32]
 * weights of the 1x1 conv, head_b its bias; head_scale > 0: head_scale*sigmoid, else ReLU.  d_out: [B][ho][wo] fp32. */
int mde_k_upconv_head(int32_t precision, const void* d_z, int32_t ldz, int32_t batch, int32_t hs, int32_t ws, int32_t ho,
                      int32_t wo, const float* d_bias, const float* d_head_w, float head_b, float head_scale,
                      float* d_out, void* stream);
/* Stream-ordered hand-shake between the ranks of one box over peer memory, to follow / precede the fused gather kernels
 * instead of a host barrier.  Every rank owns an array of n_ranks uint32 flags (zero-initialised, cudaIpc-mapped into
 * every process).  signal: publish `epoch` in slot `rank` of every rank's array (system-scope release: all earlier work
 * of this stream, including stores into peer buffers, is visible to an acquirer).  wait: block the stream until every
 * slot of our own array has reached `epoch`.  Epochs must increase; the caller keeps a producer from overwriting a buffer
 * a peer still reads (a second flag array used as acknowledgement, as monocular_depth_estimation_trt_b200/sharding.py does). */
int mde_k_peer_signal(void* const* d_flags_of_every_rank, int32_t n_ranks, int32_t rank, uint32_t epoch, void* stream);
int mde_k_peer_wait(void* d_own_flags, int32_t n_ranks, uint32_t epoch, void* stream);
/* The same hand-shake with the epoch in device memory (one uint32, zero-initialised, private to this rank), so that the launch
 * arguments are constant and a sharded forward can be captured into a CUDA graph: signal publishes *d_counter + advance
 * (storing it back when advance != 0), wait blocks until every slot of our own array has reached *d_counter. */
int mde_k_peer_signal_counter(void* const* d_flags_of_every_rank, int32_t n_ranks, int32_t rank, uint32_t* d_counter, int32_t advance,
                              void* stream);
int mde_k_peer_wait_counter(void* d_own_flags, int32_t n_ranks, const uint32_t* d_counter, void* stream);
/* Depth Pro's patch merge (depth_pro/network/encoder.py `merge`; called from the model models/depth_pro/onnx_export.py:15-29
 * builds): per_side x per_side crops of grid x grid tokens, d_tokens [per_side^2][grid^2][dim] 16-bit (a slice of the
 * trunk-only engine's output) -> NHWC map [S][S][dim], S = per_side*grid - 2*padding*(per_side-1); each crop loses
 * `padding` tokens at every edge it shares with a neighbour. */
int mde_k_merge_patches(int32_t precision, const void* d_tokens, int32_t per_side, int32_t grid, int32_t padding, int32_t dim,
                        void* d_out, void* stream);
/* VGGT's attention prologue (the frame / global blocks of the aggregator models/vggt/onnx_export.py:38-52 runs; un-vendored
 * facebookresearch/vggt `layers/attention.py`, `layers/rope.py`): per-head LayerNorm of q and k over the 64 head features
 * (weights d_qw/d_qb, d_kw/d_kb: [64] each), then the 2-D rotary embedding: features [0,32) of a head rotate with the token's y
 * position, [32,64) with x; feature i of a half pairs with i +- 16.  In place on d_qkv [rows][3*heads*64] (q | k | v packed).
 * d_pos: int32 [rows][2] (y, x) as core/export_compat.py:84-93 builds them (patch positions + 1, special tokens 0), or NULL
 * for the normalisation alone; d_cos_sin: float32 [max_pos][32] = cos(p * f_j) for j < 16, then sin(p * f_j).
 * gather_n > 0 (sequence-sharded global attention, SURVEY section 8 e row 3): the finished K row and the V row are also
 * stored at [row][0, D) and [row][D, 2D) of each d_gather[r] (pitch gather_ld elements; every rank's gathered buffer,
 * already offset to this rank's first row, cudaIpc-mapped): the K|V all-gather is this kernel's store. */
int mde_k_qknorm_rope(int32_t precision, void* d_qkv, int64_t rows, int32_t heads, const float* d_qw, const float* d_qb,
                      const float* d_kw, const float* d_kb, float eps, const int32_t* d_pos, const float* d_cos_sin,
                      int32_t max_pos, int32_t gather_n, void* const* d_gather, int32_t gather_ld, void* stream);
/* torch.nn.functional.interpolate(mode="bilinear", align_corners=False) as Depth Pro uses it -- for the input
 * (models/depth_pro/onnx2trt.py:56-74: ToTensor -> Normalize(0.5, 0.5) -> interpolate to 1536 x 1536) and for the crop
 * pyramid inside the model (models/depth_pro/onnx_export.py:15-29).  Writes n_crops windows of out_h x out_w pixels, window
 * i taken at (y0, x0) of the source resized to level_h x level_w; the resized images are never materialised.
 * d_src: planar float32 [3][src_h][src_w], or (src_is_u8_hwc) uint8 [src_h][src_w][3] which is first divided by 255 and,
 * when mean3 / std3 are given, normalised (v - mean) / std in fp32.  d_out: float32 [n_crops][3][out_h][out_w].
 * `crops` is a host array (copied into the launch), at most 64 entries. */
typedef struct mde_crop { int32_t level_h, level_w, y0, x0; } mde_crop;
int mde_k_resize_crops(const void* d_src, int32_t src_is_u8_hwc, int32_t swap_rb, int32_t src_h, int32_t src_w,
                       const mde_crop* crops, int32_t n_crops, int32_t out_h, int32_t out_w, const float* mean3,
                       const float* std3, float* d_out, void* stream);
/* Depth Pro's post-processing on the device (models/depth_pro/onnx2trt.py:118-134): f_px = 0.5 * src_w / tan(0.5 * rad(fov)),
 * inverse = canonical_inverse_depth * (src_w / f_px), interpolate(bilinear, align_corners=False) to src_h x src_w (skipped
 * when the sizes agree), depth = 1 / clamp(inverse, 1e-4, 1e4).  d_inv: float32 [h][w]; d_fov_deg: one float on the
 * device (the engine's second output); d_depth: float32 [src_h][src_w]; d_f_px: one float or NULL. */
int mde_k_depth_pro_post(const float* d_inv, const float* d_fov_deg, int32_t h, int32_t w, int32_t src_h, int32_t src_w,
                         float* d_depth, float* d_f_px, void* stream);
/* Metric3D V2's post-processing on the device (models/metric3d_v2/onnx2trt.py:148-158): the un-padded window of the canonical
 * depth map -- d_in points at its first pixel, `pitch` is the padded map's width -- is resized to out_h x out_w with
 * F.interpolate(mode="bilinear") (align_corners=False), multiplied by `mul` BEFORE the clamp (1 for the script's canonical
 * output; real_focal * resize_scale / 1000 gives metres, tools/evaluate_gt.py:162-184 "multiply first, then clamp") and
 * clamped to [clamp_lo, clamp_hi] (0, 300).  float32 in and out. */
int mde_k_resize_depth_halfpixel(const float* d_in, int32_t pitch, int32_t h, int32_t w, float* d_out, int32_t out_h, int32_t out_w,
                                 float mul, float clamp_lo, float clamp_hi, void* stream);
/* VGGT / StreamVGGT post-processing on the device (tools/evaluate_gt.py:240-262 `_square_pad_depth`, transcribed there from the
 * model scripts): the window of the network's output that the source frame occupies inside the padded square -- d_in points
 * at its first pixel, `pitch` is the output's width -- is resized to the source size with half-pixel bilinear interpolation
 * (cv2.INTER_LINEAR on a float map) and every value that is not above `floor_value` (1e-6) becomes NaN: "not a depth". */
int mde_k_resize_depth_halfpixel_nan(const float* d_in, int32_t pitch, int32_t h, int32_t w, float* d_out, int32_t out_h, int32_t out_w,
                                     float floor_value, void* stream);
/* The reference scripts' post-processing on the device (models/depth_anything_v2/onnx2trt.py:111-117):
 * F.interpolate(depth, (ho, wo), mode="bilinear", align_corners=True) then clamp(clamp_lo, clamp_hi); fp32 [B][h][w]. */
int mde_k_resize_depth(const float* d_in, int32_t batch, int32_t hi, int32_t wi, float* d_out, int32_t ho, int32_t wo,
                       float clamp_lo, float clamp_hi, void* stream);
/* 3x3 / stride 2 / pad 1 gather: NHWC -> [(b,oy,ox)][tap*c + ch] */
int mde_k_im2col_s2(int32_t precision, const void* d_in, void* d_out, int32_t batch, int32_t h, int32_t w, int32_t c,
                    void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MDE_B200_H */
