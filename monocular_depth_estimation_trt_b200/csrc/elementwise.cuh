// Memory-bound kernels of the path: fused uint8 resize + normalise + patch im2col (kernel 1 of
// north_star), float32-NCHW im2col (the reference's input contract), warp-shuffle LayerNorm,
// bilinear (align_corners=True) NHWC upsampling, the stride-2 3x3 im2col gather and the cls row.
#pragma once
#include "ptx.cuh"

namespace mde {

// ---------------------------------------------------------------------------------------------
// Kernel (1): HWC uint8 source (any size) -> cv2.resize(INTER_LINEAR) semantics on uint8 ->
// float64-derived normalisation LUT -> patch im2col rows [(b, gy, gx)][c*p*p + ky*p + kx] (16-bit),
// optionally also the float32 NCHW tensor core/preprocess.py produces (for bit-exact parity tests).
//
// Restates OpenCV's 8-bit bilinear path (11-bit fixed-point coefficients; see oracle/preprocess_np.py
// for the line-by-line statement and SURVEY section 8 a-1): coordinates in double, fractional part in
// float, coefficients rint(f * 2048) as int16, horizontal pass un-shifted, vertical pass
// (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2.
// ---------------------------------------------------------------------------------------------
struct PreprocParams {
  const uint8_t* src;       // [B][src_h][src_w][3] (batch stride src_batch_stride bytes)
  long long src_batch_stride;
  int src_h, src_w;
  int dst_h, dst_w;         // model input size (multiples of patch)
  int patch;                // 14 or 16
  int kpad;                 // im2col row pitch in elements (>= 3*patch*patch, multiple of 64)
  int swap_rb;              // source is BGR (cv2.imread) -> RGB
  const float* lut;         // [3][256] float32: (v/255 - mean_c)/std_c evaluated in float64
  void* cols;               // [B*gh*gw][kpad] 16-bit, or null
  float* nchw;              // [B][3][dst_h][dst_w] float32, or null
  int exact2x;              // cv2 switches INTER_LINEAR to INTER_AREA when both scales are exactly 2
  // keep-ratio + pad (Metric3D V2, core/preprocess.py:191-219): the resize target is inner_h x inner_w, placed at
  // (pad_top, pad_left) of the dst_h x dst_w canvas; everything else is pad_src (uint8, SOURCE channel order).
  // Stretch (Depth Anything): inner == dst, no offset.
  int inner_h, inner_w, pad_top, pad_left;
  uint8_t pad_src[4];
};

__device__ __forceinline__ void cv_linear_coeff(int d, double scale, int src_size, int& s0, short& a0, short& a1,
                                                bool clamp_coeff) {
  // *_rn intrinsics: never contracted into an FMA, so the rounding sequence is the CPU's
  float f = __double2float_rn(__dsub_rn(__dmul_rn(__dadd_rn(static_cast<double>(d), 0.5), scale), 0.5));
  int s = static_cast<int>(floorf(f));
  f = __fsub_rn(f, static_cast<float>(s));
  if (clamp_coeff) {               // x axis: coefficients collapse at the borders
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= src_size - 1) { f = 0.f; s = src_size - 1; }
  }
  s0 = s;
  a0 = static_cast<short>(__float2int_rn((1.f - f) * 2048.f));
  a1 = static_cast<short>(__float2int_rn(f * 2048.f));
}

template <typename T>
__global__ void __launch_bounds__(256) preprocess_u8_kernel(const PreprocParams p) {
  extern __shared__ __align__(16) uint8_t pp_smem[];
  const int gw = p.dst_w / p.patch;
  const int gy = blockIdx.x, b = blockIdx.y;
  // smem: xofs[dst_w] int, xa[dst_w][2] short, tile[gw][kpad] T
  int* xofs = reinterpret_cast<int*>(pp_smem);
  short* xa = reinterpret_cast<short*>(xofs + p.dst_w);
  T* tile = reinterpret_cast<T*>(pp_smem + ((p.dst_w * 8 + 15) & ~15));
  const double inv_x = static_cast<double>(p.inner_w) / p.src_w, inv_y = static_cast<double>(p.inner_h) / p.src_h;
  const double scale_x = 1.0 / inv_x, scale_y = 1.0 / inv_y;

  for (int dx = threadIdx.x; dx < p.inner_w; dx += blockDim.x) {      // indexed by the column INSIDE the resized image
    int s0; short a0, a1;
    cv_linear_coeff(dx, scale_x, p.src_w, s0, a0, a1, true);
    xofs[dx] = s0; xa[2 * dx] = a0; xa[2 * dx + 1] = a1;
  }
  // the row coefficients of this patch row's `patch` image rows, once per CTA: their float64 arithmetic per PIXEL was what
  // bound the kernel (ncu: ALU pipe 60 %, DRAM 8 %)
  __shared__ int ysy[16];
  __shared__ short yb[32];
  if (threadIdx.x < p.patch) {
    const int iy = gy * p.patch + threadIdx.x - p.pad_top;
    int sy = 0; short b0 = 0, b1 = 0;
    if (iy >= 0 && iy < p.inner_h) cv_linear_coeff(iy, scale_y, p.src_h, sy, b0, b1, false);
    ysy[threadIdx.x] = sy; yb[2 * threadIdx.x] = b0; yb[2 * threadIdx.x + 1] = b1;
  }
  if (p.cols) {
    const int k_real = 3 * p.patch * p.patch;
    const int padw = p.kpad - k_real;
    for (int i = threadIdx.x; i < gw * padw; i += blockDim.x)
      tile[(i / padw) * p.kpad + k_real + i % padw] = F16Traits<T>::from_f(0.f);
  }
  __syncthreads();

  const uint8_t* src = p.src + static_cast<long long>(b) * p.src_batch_stride;
  const long long row_bytes = static_cast<long long>(p.src_w) * 3;
  const int pp2 = p.patch * p.patch;
  for (int i = threadIdx.x; i < p.patch * p.dst_w; i += blockDim.x) {
    const int ky = i / p.dst_w, dx = i % p.dst_w;
    const int dy = gy * p.patch + ky;
    const int iy = dy - p.pad_top, ix = dx - p.pad_left;       // position inside the resized image
    int out[3];
    if (iy < 0 || iy >= p.inner_h || ix < 0 || ix >= p.inner_w) {
#pragma unroll
      for (int c = 0; c < 3; ++c) out[c] = p.pad_src[c];
    } else if (p.exact2x) {
      const uint8_t* r0 = src + static_cast<long long>(2 * iy) * row_bytes + 6 * ix;
      const uint8_t* r1 = r0 + row_bytes;
#pragma unroll
      for (int c = 0; c < 3; ++c) out[c] = (r0[c] + r0[c + 3] + r1[c] + r1[c + 3] + 2) >> 2;
    } else {
      const int sy = ysy[ky];
      const int b0 = yb[2 * ky], b1 = yb[2 * ky + 1];
      const int y0 = min(max(sy, 0), p.src_h - 1), y1 = min(max(sy + 1, 0), p.src_h - 1);
      const int sx = xofs[ix];
      const int sx1 = min(sx + 1, p.src_w - 1);    // coefficient is 0 whenever this clamp acts
      const int a0 = xa[2 * ix], a1 = xa[2 * ix + 1];
      const uint8_t* r0 = src + y0 * row_bytes;
      const uint8_t* r1 = src + y1 * row_bytes;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int h0 = r0[sx * 3 + c] * a0 + r0[sx1 * 3 + c] * a1;
        const int h1 = r1[sx * 3 + c] * a0 + r1[sx1 * 3 + c] * a1;
        out[c] = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
      }
    }
    const int gx = dx / p.patch, kx = dx % p.patch;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int cs = p.swap_rb ? 2 - c : c;               // output channel c reads source channel cs
      const float v = __ldg(p.lut + c * 256 + out[cs]);
      if (p.cols) tile[gx * p.kpad + c * pp2 + ky * p.patch + kx] = F16Traits<T>::from_f(v);
      if (p.nchw) p.nchw[((static_cast<long long>(b) * 3 + c) * p.dst_h + dy) * p.dst_w + dx] = v;
    }
  }
  if (p.cols) {
    __syncthreads();
    // one patch row = gw consecutive im2col rows = one contiguous block: 16-byte coalesced stores
    const int n16 = gw * p.kpad / 8;
    uint4* dst = reinterpret_cast<uint4*>(static_cast<T*>(p.cols) +
                                          (static_cast<long long>(b) * (p.dst_h / p.patch) + gy) * gw * p.kpad);
    const uint4* s4 = reinterpret_cast<const uint4*>(tile);
    for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = s4[i];
  }
}

// float32 NCHW [B,3,H,W] (what core/preprocess.py emits) -> the same im2col rows.
struct Im2colParams {
  const float* nchw;
  void* cols;
  int H, W, patch, kpad;
  float mean[3], inv_std[3];   // (v - mean_c) * inv_std_c before the 16-bit rounding; 0 / 1 = the tensor is used as is
};
template <typename T>
__global__ void __launch_bounds__(256) im2col_f32_kernel(const Im2colParams p) {
  extern __shared__ __align__(16) uint8_t ic_smem[];
  T* tile = reinterpret_cast<T*>(ic_smem);
  const int gw = p.W / p.patch, gh = p.H / p.patch;
  const int gy = blockIdx.x, b = blockIdx.y;
  const int pp2 = p.patch * p.patch, k_real = 3 * pp2, padw = p.kpad - k_real;
  for (int i = threadIdx.x; i < gw * padw; i += blockDim.x)
    tile[(i / padw) * p.kpad + k_real + i % padw] = F16Traits<T>::from_f(0.f);
  const int per_c = p.patch * p.W;
  for (int i = threadIdx.x; i < 3 * per_c; i += blockDim.x) {
    const int c = i / per_c, rem = i % per_c, ky = rem / p.W, x = rem % p.W;
    const float v = p.nchw[((static_cast<long long>(b) * 3 + c) * p.H + gy * p.patch + ky) * p.W + x];
    tile[(x / p.patch) * p.kpad + c * pp2 + ky * p.patch + x % p.patch] = F16Traits<T>::from_f((v - p.mean[c]) * p.inv_std[c]);
  }
  __syncthreads();
  const int n16 = gw * p.kpad / 8;
  uint4* dst = reinterpret_cast<uint4*>(static_cast<T*>(p.cols) + (static_cast<long long>(b) * gh + gy) * gw * p.kpad);
  const uint4* s4 = reinterpret_cast<const uint4*>(tile);
  for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = s4[i];
}

// ---------------------------------------------------------------------------------------------
// LayerNorm over the last dim of the fp32 residual stream, one warp per row, statistics in fp32
// (two-pass on registers), 16-bit output.  drop_cls: rows are [B][ntok] tokens; the cls row (0) of
// every image is skipped and the output is the dense [B][ntok-1] patch grid (the DPT taps).
// ---------------------------------------------------------------------------------------------
struct LayerNormParams {
  const float* x;      // [rows][D]
  const float* w;
  const float* b;
  void* out;           // 16-bit
  long long rows;
  int D;               // multiple of 128, <= 2048
  float eps;
  int drop_cls;
  int ntok;
  int identity;        // 1: no normalisation, the row is only converted (raw block output taps)
  int n_dst;           // > 0: the row is written to dst[0..n_dst) instead of out (gather buffers of the ranks, peer memory)
  long long dst_row0;  // row offset of this rank's rows inside every dst
  void* dst[8];
};
// kTap = false: the plain per-block LayerNorm (the hot one: 48 launches per forward); kTap = true adds the tap options
// (no normalisation, several destinations) without costing the plain instantiation a register or a branch.
template <typename T, int D, bool kTap>
__global__ void __launch_bounds__(256) layernorm_kernel(const LayerNormParams p) {
  constexpr int V = D / 128;   // float4 per lane
  griddep_launch_dependents();
  griddep_wait();
  // Rows are walked from the LAST one down: the GEMM that has just updated the residual stream swept it front to back, so its
  // most recent ~100 MB are still in the 126 MB L2, and the GEMM that follows starts at row 0 -- the rows written last here.
  const long long row = p.rows - 1 - (static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5));
  if (row < 0) return;
  const int lane = threadIdx.x & 31;
  long long orow = row;
  if (p.drop_cls) {       // number of leading tokens of every image that are skipped: cls (1), cls + registers (5), ...
    const long long b = row / p.ntok;
    const int t = static_cast<int>(row % p.ntok);
    if (t < p.drop_cls) return;
    orow = b * (p.ntok - p.drop_cls) + (t - p.drop_cls);
  }
  const float4* xr = reinterpret_cast<const float4*>(p.x + row * D);
  float4 v[V];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    v[i] = xr[lane + i * 32];
    sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum * (1.0f / D);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    sq += (a * a + b * b) + (c * c + d * d);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq * (1.0f / D) + p.eps);
  if (!kTap) {
    T* out = static_cast<T*>(p.out) + orow * D;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int c = (lane + i * 32) * 4;
      const float4 w = __ldg(reinterpret_cast<const float4*>(p.w + c));
      const float4 b = __ldg(reinterpret_cast<const float4*>(p.b + c));
      uint2 u;
      u.x = F16Traits<T>::pack2((v[i].x - mean) * rstd * w.x + b.x, (v[i].y - mean) * rstd * w.y + b.y);
      u.y = F16Traits<T>::pack2((v[i].z - mean) * rstd * w.z + b.z, (v[i].w - mean) * rstd * w.w + b.w);
      *reinterpret_cast<uint2*>(out + c) = u;
    }
    return;
  }
  uint2 u[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int c = (lane + i * 32) * 4;
    if (p.identity) {
      u[i].x = F16Traits<T>::pack2(v[i].x, v[i].y);
      u[i].y = F16Traits<T>::pack2(v[i].z, v[i].w);
    } else {
      const float4 w = __ldg(reinterpret_cast<const float4*>(p.w + c));
      const float4 b = __ldg(reinterpret_cast<const float4*>(p.b + c));
      u[i].x = F16Traits<T>::pack2((v[i].x - mean) * rstd * w.x + b.x, (v[i].y - mean) * rstd * w.y + b.y);
      u[i].y = F16Traits<T>::pack2((v[i].z - mean) * rstd * w.z + b.z, (v[i].w - mean) * rstd * w.w + b.w);
    }
  }
  if (p.n_dst == 0) {
    T* out = static_cast<T*>(p.out) + orow * D;
#pragma unroll
    for (int i = 0; i < V; ++i) *reinterpret_cast<uint2*>(out + (lane + i * 32) * 4) = u[i];
  } else {
    // fused all-gather: the same row goes to every rank's buffer (local HBM for our own, NVLink stores for the peers)
    for (int d = 0; d < p.n_dst; ++d) {
      T* out = static_cast<T*>(p.dst[d]) + (p.dst_row0 + orow) * D;
#pragma unroll
      for (int i = 0; i < V; ++i) *reinterpret_cast<uint2*>(out + (lane + i * 32) * 4) = u[i];
    }
  }
}

// x[b][0][:] = cls + pos[0], x[b][1 + r][:] = register r (DINOv2 with registers: no position embedding on them); the
// patch-embed GEMM epilogue writes the rows behind them
__global__ void cls_row_kernel(float* x, const float* cls, const float* pos, const float* reg, int n_reg, int ntok, int D) {
  const int b = blockIdx.x;
  float* xb = x + static_cast<long long>(b) * ntok * D;
  for (int i = threadIdx.x; i < D; i += blockDim.x) xb[i] = cls[i] + pos[i];
  for (int i = threadIdx.x; i < n_reg * D; i += blockDim.x) xb[D + i] = reg[i];
}

// VGGT's aggregator input (vggt/models/aggregator.py: camera token, four register tokens, then the trunk's normalised patch
// tokens; frame 0 of the scene takes variant 0 of the special tokens, every other frame variant 1):
// out[s][j][:] = special[variant(s)][j][:] for j < n_special, float(patch[s][j - n_special][:]) behind them.
struct AssembleTokensParams {
  const void* patch;      // [frames][tokens][D] 16-bit
  const float* special;   // [2][n_special][D]
  float* out;             // [frames][n_special + tokens][D]
  int frames, tokens, n_special, D, first_frame;   // first_frame: global index of this rank's frame 0
};
template <typename T>
__global__ void __launch_bounds__(256) assemble_tokens_kernel(const AssembleTokensParams p) {
  const int ntok = p.n_special + p.tokens;
  const long long row = blockIdx.x;                    // (frame, token)
  const int s = static_cast<int>(row / ntok), j = static_cast<int>(row % ntok);
  float* o = p.out + row * p.D;
  if (j < p.n_special) {
    const float* src = p.special + (static_cast<long long>((p.first_frame + s) == 0 ? 0 : 1) * p.n_special + j) * p.D;
    for (int i = threadIdx.x * 4; i < p.D; i += blockDim.x * 4) *reinterpret_cast<float4*>(o + i) = *reinterpret_cast<const float4*>(src + i);
    return;
  }
  const T* src = static_cast<const T*>(p.patch) + (static_cast<long long>(s) * p.tokens + (j - p.n_special)) * p.D;
  for (int i = threadIdx.x * 8; i < p.D; i += blockDim.x * 8) {
    const uint4 q = *reinterpret_cast<const uint4*>(src + i);
    const uint32_t* u = &q.x;
    float4 a, b;
    float2 v = F16Traits<T>::unpack2(u[0]); a.x = v.x; a.y = v.y;
    v = F16Traits<T>::unpack2(u[1]); a.z = v.x; a.w = v.y;
    v = F16Traits<T>::unpack2(u[2]); b.x = v.x; b.y = v.y;
    v = F16Traits<T>::unpack2(u[3]); b.z = v.x; b.w = v.y;
    *reinterpret_cast<float4*>(o + i) = a;
    *reinterpret_cast<float4*>(o + i + 4) = b;
  }
}

// ---------------------------------------------------------------------------------------------
// Bilinear resize, align_corners=True, NHWC 16-bit, 8 channels (16 bytes) per thread.
// Matches torch.nn.functional.interpolate: src = dst * (in-1)/(out-1), lerp in fp32.
// ---------------------------------------------------------------------------------------------
struct BilinearParams {
  const void* in;   // [B][Hi][Wi][C]
  void* out;        // [B][Ho][Wo][C]
  int B, Hi, Wi, Ho, Wo, C;
  float sy, sx;     // (Hi-1)/(Ho-1), (Wi-1)/(Wo-1)
  const float* addend;   // optional fp32 [Ho][Wo][C] added to every image after the interpolation (VGGT's position embedding)
};
// grid = (Ho, B): one output row per CTA, so the vertical taps and weight are CTA-uniform and the two source
// rows (Wi*C*2 bytes each) stay in L1 while the row is produced; no 64-bit index arithmetic per element.
template <typename T>
__global__ void __launch_bounds__(256) bilinear_nhwc_kernel(const BilinearParams p) {
  using Tr = F16Traits<T>;
  const int y = blockIdx.x, b = blockIdx.y;
  const int cv = p.C / 8;
  const float fy = y * p.sy;
  const int y0 = min(static_cast<int>(fy), p.Hi - 1);
  const int y1 = min(y0 + 1, p.Hi - 1);
  const float wy = fy - y0;
  const T* row0 = static_cast<const T*>(p.in) + (static_cast<long long>(b) * p.Hi + y0) * p.Wi * p.C;
  const T* row1 = static_cast<const T*>(p.in) + (static_cast<long long>(b) * p.Hi + y1) * p.Wi * p.C;
  T* orow = static_cast<T*>(p.out) + (static_cast<long long>(b) * p.Ho + y) * p.Wo * p.C;
  const int total = p.Wo * cv;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int x = i / cv, c = (i - x * cv) * 8;
    const float fx = x * p.sx;
    const int x0 = min(static_cast<int>(fx), p.Wi - 1);
    const int x1 = min(x0 + 1, p.Wi - 1);
    const float wx = fx - x0;
    const uint4 q00 = *reinterpret_cast<const uint4*>(row0 + x0 * p.C + c);
    const uint4 q01 = *reinterpret_cast<const uint4*>(row0 + x1 * p.C + c);
    const uint4 q10 = *reinterpret_cast<const uint4*>(row1 + x0 * p.C + c);
    const uint4 q11 = *reinterpret_cast<const uint4*>(row1 + x1 * p.C + c);
    const uint32_t* a = &q00.x; const uint32_t* bb = &q01.x; const uint32_t* cc = &q10.x; const uint32_t* d = &q11.x;
    uint4 r; uint32_t* ro = &r.x;
    float add[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (p.addend) {
      const float4* ad = reinterpret_cast<const float4*>(p.addend + (static_cast<long long>(y) * p.Wo + x) * p.C + c);
      const float4 a0 = __ldg(ad), a1 = __ldg(ad + 1);
      add[0] = a0.x; add[1] = a0.y; add[2] = a0.z; add[3] = a0.w; add[4] = a1.x; add[5] = a1.y; add[6] = a1.z; add[7] = a1.w;
    }
    // a + w * (b - a) on packed fp32 pairs: the same arithmetic (subtract, multiply-add; torch's own form), half the issue slots
    const f32x2 wx2 = f2_splat(wx), wy2 = f2_splat(wy), neg1 = f2_splat(-1.0f);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 v00 = Tr::unpack2(a[k]), v01 = Tr::unpack2(bb[k]), v10 = Tr::unpack2(cc[k]), v11 = Tr::unpack2(d[k]);
      const f32x2 p00 = f2_pack(v00.x, v00.y), p10 = f2_pack(v10.x, v10.y);
      const f32x2 top = f2_fma(wx2, f2_fma(p00, neg1, f2_pack(v01.x, v01.y)), p00);
      const f32x2 bot = f2_fma(wx2, f2_fma(p10, neg1, f2_pack(v11.x, v11.y)), p10);
      float r0, r1;
      f2_unpack(f2_fma(wy2, f2_fma(top, neg1, bot), top), r0, r1);
      ro[k] = Tr::pack2(r0 + add[2 * k], r1 + add[2 * k + 1]);
    }
    *reinterpret_cast<uint4*>(orow + static_cast<long long>(i) * 8) = r;
  }
}

// ---------------------------------------------------------------------------------------------
// Post-processing of the reference's inference scripts on the device (models/depth_anything_v2/onnx2trt.py:111-117):
// F.interpolate(depth, (src_h, src_w), mode="bilinear", align_corners=True) then clamp(min, max), fp32, one channel.
// ---------------------------------------------------------------------------------------------
struct ResizeDepthParams {
  const float* in;   // [B][Hi][Wi]
  float* out;        // [B][Ho][Wo]
  int B, Hi, Wi, Ho, Wo;
  float sy, sx, lo, hi;
};
__global__ void __launch_bounds__(256) resize_depth_kernel(const ResizeDepthParams p) {
  const int y = blockIdx.y, b = blockIdx.z;
  const int x = blockIdx.x * 256 + threadIdx.x;
  if (x >= p.Wo) return;
  const float fy = y * p.sy, fx = x * p.sx;
  const int y0 = min(static_cast<int>(fy), p.Hi - 1), x0 = min(static_cast<int>(fx), p.Wi - 1);
  const int y1 = min(y0 + 1, p.Hi - 1), x1 = min(x0 + 1, p.Wi - 1);
  const float wy = fy - y0, wx = fx - x0;
  const float* r0 = p.in + (static_cast<long long>(b) * p.Hi + y0) * p.Wi;
  const float* r1 = p.in + (static_cast<long long>(b) * p.Hi + y1) * p.Wi;
  const float t = __ldg(r0 + x0) + wx * (__ldg(r0 + x1) - __ldg(r0 + x0));
  const float u = __ldg(r1 + x0) + wx * (__ldg(r1 + x1) - __ldg(r1 + x0));
  p.out[(static_cast<long long>(b) * p.Ho + y) * p.Wo + x] = fminf(fmaxf(t + wy * (u - t), p.lo), p.hi);
}

// ---------------------------------------------------------------------------------------------
// Stream-ordered hand-shake between the ranks of one box over peer memory (no host barrier, no collective):
// every rank owns a small flag array, mapped into every process.  `signal` runs after the kernel that wrote into the
// peers' buffers (same stream): thread r publishes `epoch` into slot `rank` of rank r's flags with a system-scope
// release, so everything this rank stored before is visible to whoever acquires the flag.  `wait` spins (system-scope
// acquire) until every rank's slot of OUR flags has reached `epoch`; kernels enqueued behind it may read the gathered data.
// ---------------------------------------------------------------------------------------------
struct PeerFlagsParams {
  unsigned int* flags[8];   // signal: every rank's array (peer pointers); wait: flags[0] = our own array
  int n_ranks, rank;
  unsigned int epoch;
};
__global__ void peer_signal_kernel(const PeerFlagsParams p) {
  const int r = threadIdx.x;
  if (r >= p.n_ranks) return;
  __threadfence_system();
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p.flags[r] + p.rank), "r"(p.epoch) : "memory");
}
__global__ void peer_wait_kernel(const PeerFlagsParams p) {
  const int r = threadIdx.x;
  if (r >= p.n_ranks) return;
  const long long t0 = clock64();
  unsigned int v;
  do {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p.flags[0] + r) : "memory");
    if (clock64() - t0 > 20LL * MDE_WATCHDOG_CYCLES) __trap();      // ~40 s: a rank died
  } while (static_cast<int>(v - p.epoch) < 0);                      // wrap-safe "v >= epoch"
  __threadfence_system();
}

// The same hand-shake with the epoch kept in device memory, so that the launch arguments never change and a whole sharded
// forward can be captured into a CUDA graph and replayed: `signal` publishes *counter + advance (and stores it back when it
// advanced), `wait` blocks until every slot has reached *counter.
struct PeerCounterParams {
  unsigned int* flags[8];
  unsigned int* counter;
  int n_ranks, rank, advance;
};
__global__ void peer_signal_counter_kernel(const PeerCounterParams p) {
  const int r = threadIdx.x;
  const unsigned int epoch = *reinterpret_cast<volatile unsigned int*>(p.counter) + static_cast<unsigned int>(p.advance);
  __syncwarp();
  if (r == 0 && p.advance) *p.counter = epoch;
  if (r >= p.n_ranks) return;
  __threadfence_system();
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p.flags[r] + p.rank), "r"(epoch) : "memory");
}
__global__ void peer_wait_counter_kernel(const PeerCounterParams p) {
  const int r = threadIdx.x;
  if (r >= p.n_ranks) return;
  const unsigned int epoch = *reinterpret_cast<volatile unsigned int*>(p.counter);
  const long long t0 = clock64();
  unsigned int v;
  do {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p.flags[0] + r) : "memory");
    if (clock64() - t0 > 20LL * MDE_WATCHDOG_CYCLES) __trap();
  } while (static_cast<int>(v - epoch) < 0);
  __threadfence_system();
}

// ---------------------------------------------------------------------------------------------
// Depth Pro's `merge`: per_side x per_side crops of grid x grid tokens each (row-major crops, [crop][token][D] 16-bit,
// the layout the trunk-only engine writes) -> one NHWC feature map [S][S][D], S = per_side*grid - 2*pad*(per_side-1):
// every crop loses `pad` tokens at each edge it shares with a neighbour.  16 bytes per thread.
// ---------------------------------------------------------------------------------------------
struct MergeParams {
  const void* in;
  void* out;
  int per_side, grid, pad, D, S;
};
__device__ __forceinline__ void merge_axis(int o, int per_side, int grid, int pad, int& crop, int& tok) {
  const int first = per_side > 1 ? grid - pad : grid, mid = grid - 2 * pad;
  if (o < first) { crop = 0; tok = o; return; }
  const int r = o - first;
  crop = min(1 + r / mid, per_side - 1);
  tok = pad + r - (crop - 1) * mid;
}
template <typename T>
__global__ void __launch_bounds__(256) merge_patches_kernel(const MergeParams p) {
  const int cv = p.D / 8;
  const long long total = static_cast<long long>(p.S) * p.S * cv;
  for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * 256) {
    const int c = static_cast<int>(i % cv);
    const int x = static_cast<int>((i / cv) % p.S), y = static_cast<int>(i / cv / p.S);
    int cy, ty, cx, tx;
    merge_axis(y, p.per_side, p.grid, p.pad, cy, ty);
    merge_axis(x, p.per_side, p.grid, p.pad, cx, tx);
    const long long src = ((static_cast<long long>(cy) * p.per_side + cx) * p.grid * p.grid + ty * p.grid + tx) * p.D + c * 8;
    reinterpret_cast<uint4*>(static_cast<T*>(p.out))[i] = *reinterpret_cast<const uint4*>(static_cast<const T*>(p.in) + src);
  }
}

// ---------------------------------------------------------------------------------------------
// 3x3 / stride 2 / pad 1 gather: NHWC [B][H][W][C] -> rows [(b, oy, ox)][tap*C + c] for a plain GEMM
// (resize_layers[3]: 37x37 -> 19x19; 0.5 % of the FLOPs, not worth a strided tensor map).
// ---------------------------------------------------------------------------------------------
struct Im2colS2Params {
  const void* in;
  void* out;
  int B, H, W, C, Ho, Wo;
};
template <typename T>
__global__ void __launch_bounds__(256) im2col_s2_kernel(const Im2colS2Params p) {
  const int cv = p.C / 8;
  const long long total = static_cast<long long>(p.B) * p.Ho * p.Wo * 9 * cv;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % cv);
    long long r = i / cv;
    const int tap = static_cast<int>(r % 9); r /= 9;
    const int ox = static_cast<int>(r % p.Wo); r /= p.Wo;
    const int oy = static_cast<int>(r % p.Ho);
    const int b = static_cast<int>(r / p.Ho);
    const int y = 2 * oy + tap / 3 - 1, x = 2 * ox + tap % 3 - 1;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (y >= 0 && y < p.H && x >= 0 && x < p.W)
      v = *reinterpret_cast<const uint4*>(static_cast<const T*>(p.in) +
                                          ((static_cast<long long>(b) * p.H + y) * p.W + x) * p.C + c * 8);
    *reinterpret_cast<uint4*>(static_cast<T*>(p.out) +
                              (((static_cast<long long>(b) * p.Ho + oy) * p.Wo + ox) * 9 + tap) * p.C + c * 8) = v;
  }
}

// ---------------------------------------------------------------------------------------------
// torch.nn.functional.interpolate(mode="bilinear", align_corners=False), the resize Depth Pro uses everywhere
// (models/depth_pro/onnx2trt.py:56-74 for the input, the crop pyramid inside the model, :124-128 for the output):
// src = fma(scale, dst + 0.5, -0.5) clamped at 0, i0 = trunc(src), i1 = i0 + (i0 < size - 1), l1 = src - i0, l0 = 1 - l1,
// out = h0 * (w0 * a + w1 * b) + h1 * (w0 * c + w1 * d), every step a separately rounded fp32 operation.
// One launch writes any number of crop windows of (virtually) resized copies of ONE source image: the 25 + 9 + 1 crops
// of the pyramid never exist as resized images in memory.  The source is planar fp32 or uint8 HWC; the latter goes
// through ToTensor (/ 255) and Normalize ((v - mean) / std) in fp32 first, as the script's transform does.
// ---------------------------------------------------------------------------------------------
struct HalfPixelCrop { int level_h, level_w, y0, x0; };
struct ResizeCropsParams {
  const void* src;
  float* out;               // [n_crops][3][out_h][out_w]
  int src_u8, swap_rb, normalise, src_h, src_w, out_h, out_w, n_crops;
  float mean[3], std[3];
  HalfPixelCrop crops[64];
};
__device__ __forceinline__ void halfpixel_coord(int dst, float scale, int size, int* i0, int* i1, float* l0, float* l1) {
  float s = fmaf(scale, __fadd_rn(static_cast<float>(dst), 0.5f), -0.5f);   // fused, as torch's kernels compile it (x86 -mfma, nvcc)
  s = s < 0.f ? 0.f : s;
  const int i = min(static_cast<int>(s), size - 1);
  *i0 = i;
  *i1 = i + (i < size - 1 ? 1 : 0);
  const float l = fminf(fmaxf(__fsub_rn(s, static_cast<float>(i)), 0.f), 1.f);
  *l1 = l;
  *l0 = __fsub_rn(1.f, l);
}
__global__ void __launch_bounds__(256) resize_crops_kernel(const ResizeCropsParams p) {
  const int ox = blockIdx.x * 256 + threadIdx.x, oy = blockIdx.y;
  const int crop = blockIdx.z / 3, ch = blockIdx.z % 3;
  if (ox >= p.out_w) return;
  const HalfPixelCrop c = p.crops[crop];
  int ya, yb, xa, xb;
  float h0, h1, w0, w1;
  halfpixel_coord(c.y0 + oy, __fdiv_rn(static_cast<float>(p.src_h), static_cast<float>(c.level_h)), p.src_h, &ya, &yb, &h0, &h1);
  halfpixel_coord(c.x0 + ox, __fdiv_rn(static_cast<float>(p.src_w), static_cast<float>(c.level_w)), p.src_w, &xa, &xb, &w0, &w1);
  float a, b, cc, d;
  if (p.src_u8) {
    const unsigned char* s = static_cast<const unsigned char*>(p.src);
    const int sc = p.swap_rb ? 2 - ch : ch;
    const float m = p.mean[ch], sd = p.std[ch];
    auto px = [&](int y, int x) {
      float v = __fdiv_rn(static_cast<float>(s[(static_cast<long long>(y) * p.src_w + x) * 3 + sc]), 255.f);
      return p.normalise ? __fdiv_rn(__fsub_rn(v, m), sd) : v;
    };
    a = px(ya, xa); b = px(ya, xb); cc = px(yb, xa); d = px(yb, xb);
  } else {
    const float* s = static_cast<const float*>(p.src) + static_cast<long long>(ch) * p.src_h * p.src_w;
    a = __ldg(s + static_cast<long long>(ya) * p.src_w + xa); b = __ldg(s + static_cast<long long>(ya) * p.src_w + xb);
    cc = __ldg(s + static_cast<long long>(yb) * p.src_w + xa); d = __ldg(s + static_cast<long long>(yb) * p.src_w + xb);
  }
  const float top = __fadd_rn(__fmul_rn(w0, a), __fmul_rn(w1, b));
  const float bot = __fadd_rn(__fmul_rn(w0, cc), __fmul_rn(w1, d));
  p.out[((static_cast<long long>(crop) * 3 + ch) * p.out_h + oy) * p.out_w + ox] = __fadd_rn(__fmul_rn(h0, top), __fmul_rn(h1, bot));
}

// ---------------------------------------------------------------------------------------------
// Depth Pro's post-processing (models/depth_pro/onnx2trt.py:118-134) on the device: f_px = 0.5 W / tan(0.5 rad(fov)),
// inverse depth * (W / f_px), resized back to the source size (the interpolation above), depth = 1 / clamp(., 1e-4, 1e4).
// ---------------------------------------------------------------------------------------------
struct DepthProPostParams {
  const float* inv;         // [h][pitch] map (Depth Pro: canonical inverse depth; Metric3D: the un-padded window of the canonical depth)
  const float* fov_deg;     // [1] on the device, or NULL: `mul` is used instead of src_w / f_px
  float* depth;             // [src_h][src_w]
  float* f_px;              // [1] or NULL
  int h, w, pitch, src_h, src_w;
  float mul, lo, hi;
  int reciprocal;           // 1: out = 1 / clamp(v, lo, hi);  0: out = clamp(v, lo, hi)
  int nan_below;            // 1: out = v > lo ? min(v, hi) : NaN   (VGGT: "not a depth", tools/evaluate_gt.py:258-262)
};
__global__ void __launch_bounds__(256) depth_pro_post_kernel(const DepthProPostParams p) {
  const int ox = blockIdx.x * 256 + threadIdx.x, oy = blockIdx.y;
  float s = p.mul;
  if (p.fov_deg) {
    const float half_rad = __fmul_rn(0.5f, __fmul_rn(__ldg(p.fov_deg), 0.017453292519943295f));
    const float f_px = __fdiv_rn(__fmul_rn(0.5f, static_cast<float>(p.src_w)), tanf(half_rad));
    if (p.f_px && oy == 0 && ox == 0) *p.f_px = f_px;
    s = __fdiv_rn(static_cast<float>(p.src_w), f_px);
  }
  if (ox >= p.src_w) return;
  int ya, yb, xa, xb;
  float h0, h1, w0, w1;
  halfpixel_coord(oy, __fdiv_rn(static_cast<float>(p.h), static_cast<float>(p.src_h)), p.h, &ya, &yb, &h0, &h1);
  halfpixel_coord(ox, __fdiv_rn(static_cast<float>(p.w), static_cast<float>(p.src_w)), p.w, &xa, &xb, &w0, &w1);
  const float a = __fmul_rn(__ldg(p.inv + static_cast<long long>(ya) * p.pitch + xa), s), b = __fmul_rn(__ldg(p.inv + static_cast<long long>(ya) * p.pitch + xb), s);
  const float c = __fmul_rn(__ldg(p.inv + static_cast<long long>(yb) * p.pitch + xa), s), d = __fmul_rn(__ldg(p.inv + static_cast<long long>(yb) * p.pitch + xb), s);
  float v;
  if (p.h == p.src_h && p.w == p.src_w) v = a;       // the script skips the interpolation when the sizes agree
  else v = __fadd_rn(__fmul_rn(h0, __fadd_rn(__fmul_rn(w0, a), __fmul_rn(w1, b))), __fmul_rn(h1, __fadd_rn(__fmul_rn(w0, c), __fmul_rn(w1, d))));
  if (p.nan_below) {
    p.depth[static_cast<long long>(oy) * p.src_w + ox] = v > p.lo ? fminf(v, p.hi) : __int_as_float(0x7fc00000);
    return;
  }
  v = fminf(fmaxf(v, p.lo), p.hi);
  p.depth[static_cast<long long>(oy) * p.src_w + ox] = p.reciprocal ? __fdiv_rn(1.f, v) : v;
}

// ---------------------------------------------------------------------------------------------
// VGGT's attention prologue (vggt/layers/attention.py `Attention.forward`, the blocks the aggregator of
// models/vggt/onnx_export.py:38-52 alternates): per-head LayerNorm of q and k over the 64 head features (`qk_norm`), then
// the 2-D rotary embedding -- features [0,32) rotate with the token's y position, [32,64) with x; inside each half
// feature i pairs with i +- 16 and frequency index i % 16 (positions as core/export_compat.py:84-93 builds them).
// In place on the packed q|k|v rows the QKV GEMM wrote; cos/sin come from a small fp32 table [position][cos 16 | sin 16]
// built on the host with upstream's formula.  One warp per token row, a lane owns whole head vectors (no shuffles).
// Sequence-sharded global attention: the finished K row and the V row are ALSO stored into every rank's gathered
// [tokens, 2D] buffer (peer memory over NVLink) -- the all-gather of the exchange step is this kernel's store.
// ---------------------------------------------------------------------------------------------
struct QkNormRopeParams {
  void* qkv;
  const float *qw, *qb, *kw, *kb;
  const int* pos;            // [rows][2] (y, x), or NULL: normalisation only
  const float* cos_sin;      // [max_pos][32]
  long long rows;
  int heads, max_pos, gather_n, gather_ld;
  float eps;
  void* gather[8];           // each already offset to this rank's first row
};
constexpr int kRopeSmemBytes = 8 * 2 * 32 * 128;      // per warp: two buffers of up to 32 head vectors of 128 bytes
// One head-vector pass over the swizzled warp buffer `wb`: lane v < nv owns vector v0 + v (see the kernel's comment).
template <typename T>
__device__ __forceinline__ void qknorm_rope_vectors(const QkNormRopeParams& p, uint4* wb, int v0, int nv, int lane, const float* cy, const float* cx) {
  using Tr = F16Traits<T>;
  if (lane < nv) {
      const int gv = v0 + lane;
      const bool is_k = gv >= p.heads;
      const float* wp = is_k ? p.kw : p.qw;
      const float* bp = is_k ? p.kb : p.qb;
      uint4* mine = wb + lane * 8;
      const int sw = lane & 7;
      float x[64];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint4 u = mine[i ^ sw];
        const uint32_t* w = &u.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = Tr::unpack2(w[k]);
          x[i * 8 + 2 * k] = f.x; x[i * 8 + 2 * k + 1] = f.y;
        }
      }
      float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 64; ++i) s4[i & 3] += x[i];
      const float mean = ((s4[0] + s4[1]) + (s4[2] + s4[3])) * (1.0f / 64.0f);
      float q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 64; ++i) { x[i] -= mean; q4[i & 3] = fmaf(x[i], x[i], q4[i & 3]); }
      const float rstd = __frsqrt_rn(((q4[0] + q4[1]) + (q4[2] + q4[3])) * (1.0f / 64.0f) + p.eps);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(wp) + i), b = __ldg(reinterpret_cast<const float4*>(bp) + i);
        x[4 * i] = x[4 * i] * rstd * w.x + b.x; x[4 * i + 1] = x[4 * i + 1] * rstd * w.y + b.y;
        x[4 * i + 2] = x[4 * i + 2] * rstd * w.z + b.z; x[4 * i + 3] = x[4 * i + 3] * rstd * w.w + b.w;
      }
      if (p.pos) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const float* t = half ? cx : cy;
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 c = __ldg(reinterpret_cast<const float4*>(t) + j4), sn = __ldg(reinterpret_cast<const float4*>(t + 16) + j4);
            const float cc[4] = {c.x, c.y, c.z, c.w}, ss[4] = {sn.x, sn.y, sn.z, sn.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int j = j4 * 4 + k;
              const float lo = x[half * 32 + j], hi = x[half * 32 + 16 + j];
              x[half * 32 + j] = lo * cc[k] - hi * ss[k];
              x[half * 32 + 16 + j] = hi * cc[k] + lo * ss[k];
            }
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        uint4 u;
        u.x = Tr::pack2(x[i * 8], x[i * 8 + 1]); u.y = Tr::pack2(x[i * 8 + 2], x[i * 8 + 3]);
        u.z = Tr::pack2(x[i * 8 + 4], x[i * 8 + 5]); u.w = Tr::pack2(x[i * 8 + 6], x[i * 8 + 7]);
        mine[i ^ sw] = u;
      }
    }
}
template <typename T>
__global__ void __launch_bounds__(256, 2) qknorm_rope_kernel(const QkNormRopeParams p) {
  extern __shared__ __align__(16) uint8_t rope_smem[];
  griddep_launch_dependents();
  griddep_wait();
  const int lane = threadIdx.x & 31;
  const int D = p.heads * 64;
  const int nvec = 2 * p.heads;
  uint4* wbuf = reinterpret_cast<uint4*>(rope_smem) + (threadIdx.x >> 5) * 512;
  // The q | k part of a row is one contiguous run of 2 * heads head vectors (128 bytes each).  The warp moves it between
  // global and shared memory in fully coalesced 16-byte pieces (512 contiguous bytes per instruction); in between, lane v owns
  // head vector v -- statistics and rotation partners stay inside the thread, no shuffles.  Chunk i of vector v lives at
  // v * 8 + (i ^ (v & 7)): both the coalesced side (4 vectors x 8 chunks per instruction) and the per-vector side (32 vectors,
  // one chunk index) touch every bank group the same number of times.  Warps are persistent (grid = 2 CTAs per SM): while a
  // row is normalised and rotated, the next row of the warp is already on its way into the other buffer (cp.async).
  const long long warps = static_cast<long long>(gridDim.x) * 8;
  long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  auto slot = [](int c) { return (c & ~7) + ((c & 7) ^ ((c >> 3) & 7)); };
  auto fetch = [&](long long r, uint4* wb) {
    if (r < p.rows) {
      const uint4* gsrc = reinterpret_cast<const uint4*>(static_cast<const T*>(p.qkv) + r * 3 * D);
      const int chunks = min(nvec, 32) * 8;
      for (int c = lane; c < chunks; c += 32) cp_async_16(wb + slot(c), gsrc + c, true);
    }
    cp_async_commit();
  };
  int buf = 0;
  fetch(row, wbuf);
  // the token's (y, x) position is needed before the table rows can be addressed: loaded one row ahead as well
  int2 pos_next = make_int2(0, 0);
  if (p.pos && row < p.rows) pos_next = __ldg(reinterpret_cast<const int2*>(p.pos) + row);
  for (; row < p.rows; row += warps, buf ^= 1) {
    uint4* wb = wbuf + buf * 256;
    fetch(row + warps, wbuf + (buf ^ 1) * 256);
    const int2 pos_cur = pos_next;
    if (p.pos && row + warps < p.rows) pos_next = __ldg(reinterpret_cast<const int2*>(p.pos) + row + warps);
    const float* cy = nullptr;
    const float* cx = nullptr;
    if (p.pos) {
      cy = p.cos_sin + static_cast<long long>(min(max(pos_cur.x, 0), p.max_pos - 1)) * 32;
      cx = p.cos_sin + static_cast<long long>(min(max(pos_cur.y, 0), p.max_pos - 1)) * 32;
    }
    T* base = static_cast<T*>(p.qkv) + row * 3 * D;
    cp_async_wait<1>();
    __syncwarp();
    for (int v0 = 0; v0 < nvec; v0 += 32) {
      const int nv = min(32, nvec - v0);
      if (v0 > 0) {                                  // more than 32 vectors per row: the later chunks are loaded in place
        const uint4* gsrc = reinterpret_cast<const uint4*>(base) + v0 * 8;
        for (int c = lane; c < nv * 8; c += 32) wb[slot(c)] = gsrc[c];
        __syncwarp();
      }
      qknorm_rope_vectors<T>(p, wb, v0, nv, lane, cy, cx);
      __syncwarp();
      uint4* gdst = reinterpret_cast<uint4*>(base) + v0 * 8;
      for (int c = lane; c < nv * 8; c += 32) {
        const uint4 u = wb[slot(c)];
        gdst[c] = u;
        const int kc = v0 * 8 + c - p.heads * 8;           // 16-byte chunk index inside the K part, if this chunk is K
        if (kc >= 0)
          for (int r = 0; r < p.gather_n; ++r)
            *(reinterpret_cast<uint4*>(static_cast<T*>(p.gather[r]) + row * p.gather_ld) + kc) = u;
      }
      __syncwarp();
    }
    if (p.gather_n > 0) {
      const uint4* v = reinterpret_cast<const uint4*>(base + 2 * D);
      for (int i = lane; i < D / 8; i += 32) {
        const uint4 u = v[i];
        for (int r = 0; r < p.gather_n; ++r)
          *reinterpret_cast<uint4*>(static_cast<T*>(p.gather[r]) + row * p.gather_ld + D + i * 8) = u;
      }
    }
  }
  cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------
// VGGT / StreamVGGT input contract (core/preprocess.py:222-265, 493-498): the RGB frame is padded to a square with white at
// source resolution (the same count on both sides), resized ONCE with cv2.INTER_CUBIC, divided by 255 in float32.  The
// padded image is never built: a tap outside the frame reads the pad value.  The resize is OpenCV's own 8-bit cubic
// (oracle/preprocess_np.py `resize_cubic_u8`): Keys kernel (A = -0.75) in fp32 in `interpolateCubic`'s order, coefficients
// rounded to 11 bits, integer horizontal pass with replicated borders, vertical pass in fp32 exactly as the vector unit does
// it (int row * (beta * 2^-22), summed right to left, every step separately rounded, round half to even), and the scalar
// integer tail for the last (W * 3) % 8 interleaved elements of a row.
// ---------------------------------------------------------------------------------------------
struct CubicPadParams {
  const unsigned char* src;   // [B][src_h][src_w][3]
  float* out;                 // [B][3][dst_h][dst_w]
  int B, src_h, src_w, dst_h, dst_w, top, left, pad_h, pad_w, swap_rb, pad_value, tail;
  double scale_y, scale_x;
};
__device__ __forceinline__ void cubic_taps(int d, double scale, int* s, int (&co)[4]) {
  const float fx = static_cast<float>((d + 0.5) * scale - 0.5);
  const float fl = floorf(fx);
  *s = static_cast<int>(fl);
  const float x = __fsub_rn(fx, fl), A = -0.75f;
  const float x1 = __fadd_rn(x, 1.f), xm = __fsub_rn(1.f, x);
  const float c0 = __fsub_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fsub_rn(__fmul_rn(A, x1), -3.75f), x1), -6.f), x1), -3.f);
  const float c1 = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(1.25f, x), 2.25f), x), x), 1.f);
  const float c2 = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(1.25f, xm), 2.25f), xm), xm), 1.f);
  const float c3 = __fsub_rn(__fsub_rn(__fsub_rn(1.f, c0), c1), c2);
  co[0] = __float2int_rn(__fmul_rn(c0, 2048.f)); co[1] = __float2int_rn(__fmul_rn(c1, 2048.f));
  co[2] = __float2int_rn(__fmul_rn(c2, 2048.f)); co[3] = __float2int_rn(__fmul_rn(c3, 2048.f));
}
__global__ void __launch_bounds__(256) preprocess_cubic_pad_kernel(const CubicPadParams p) {
  const int dx = blockIdx.x * 256 + threadIdx.x, dy = blockIdx.y;
  const int b = blockIdx.z / 3, c = blockIdx.z % 3;
  if (dx >= p.dst_w) return;
  int sy, sx, ya[4], xa[4];
  cubic_taps(dy, p.scale_y, &sy, ya);
  cubic_taps(dx, p.scale_x, &sx, xa);
  const unsigned char* img = p.src + static_cast<long long>(b) * p.src_h * p.src_w * 3;
  const int sc = p.swap_rb ? 2 - c : c;
  int row[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int y = min(max(sy + k - 1, 0), p.pad_h - 1) - p.top;          // replicated border of the PADDED image, then into the frame
    int acc = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int x = min(max(sx + j - 1, 0), p.pad_w - 1) - p.left;
      const int v = (y >= 0 && y < p.src_h && x >= 0 && x < p.src_w) ? img[(static_cast<long long>(y) * p.src_w + x) * 3 + sc] : p.pad_value;
      acc += v * xa[j];
    }
    row[k] = acc;
  }
  int level;
  if (dx * 3 + c >= p.dst_w * 3 - p.tail) {
    const long long sum = static_cast<long long>(row[0]) * ya[0] + static_cast<long long>(row[1]) * ya[1] +
                          static_cast<long long>(row[2]) * ya[2] + static_cast<long long>(row[3]) * ya[3];
    level = static_cast<int>((sum + (1LL << 21)) >> 22);
  } else {
    const float sc22 = 1.0f / 4194304.0f;
    const float t3 = __fmul_rn(static_cast<float>(row[3]), __fmul_rn(static_cast<float>(ya[3]), sc22));
    const float t2 = __fadd_rn(__fmul_rn(static_cast<float>(row[2]), __fmul_rn(static_cast<float>(ya[2]), sc22)), t3);
    const float t1 = __fadd_rn(__fmul_rn(static_cast<float>(row[1]), __fmul_rn(static_cast<float>(ya[1]), sc22)), t2);
    const float t0 = __fadd_rn(__fmul_rn(static_cast<float>(row[0]), __fmul_rn(static_cast<float>(ya[0]), sc22)), t1);
    level = __float2int_rn(t0);
  }
  level = min(max(level, 0), 255);
  p.out[((static_cast<long long>(b) * 3 + c) * p.dst_h + dy) * p.dst_w + dx] = __fdiv_rn(static_cast<float>(level), 255.f);
}

// ---------------------------------------------------------------------------------------------
// Depth-Anything-AC `native` profile (core/preprocess.py:470-476 `da_ac(h, w, stretch=False)`, models/depth_anything_ac/
// onnx2trt.py:50-75): the uint8 frame is NOT resized; it becomes float32 / 255 (RGB), is resized with cv2's FLOAT
// INTER_CUBIC to the keep-ratio size and standardised in float64.  OpenCV's own float cubic (oracle/preprocess_np.py
// `resize_cubic_f32`, bit-exact with cv2 when IPP is off): Keys coefficients (A = -0.75) in fp32 in `interpolateCubic`'s order,
// horizontal taps with replicated borders added left to right, vertical taps added right to left in the 4-lane vector loop
// and left to right for the last (W * 3) % 4 interleaved elements of a row, every product and sum rounded separately.
// ---------------------------------------------------------------------------------------------
struct CubicF32Params {
  const unsigned char* src;   // [B][src_h][src_w][3]
  float* out;                 // [B][3][dst_h][dst_w]
  int B, src_h, src_w, dst_h, dst_w, swap_rb, tail;
  double scale_y, scale_x;
  double mean[3], std[3];     // of the OUTPUT channels (RGB)
};
__device__ __forceinline__ void cubic_taps_f32(int d, double scale, int* s, float (&co)[4]) {
  const float fx = static_cast<float>((d + 0.5) * scale - 0.5);
  const float fl = floorf(fx);
  *s = static_cast<int>(fl);
  const float x = __fsub_rn(fx, fl), A = -0.75f;
  const float x1 = __fadd_rn(x, 1.f), xm = __fsub_rn(1.f, x);
  co[0] = __fsub_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fsub_rn(__fmul_rn(A, x1), -3.75f), x1), -6.f), x1), -3.f);
  co[1] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(1.25f, x), 2.25f), x), x), 1.f);
  co[2] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(1.25f, xm), 2.25f), xm), xm), 1.f);
  co[3] = __fsub_rn(__fsub_rn(__fsub_rn(1.f, co[0]), co[1]), co[2]);
}
__global__ void __launch_bounds__(256) preprocess_cubic_f32_kernel(const CubicF32Params p) {
  const int dx = blockIdx.x * 256 + threadIdx.x, dy = blockIdx.y;
  const int b = blockIdx.z / 3, c = blockIdx.z % 3;
  if (dx >= p.dst_w) return;
  int sy, sx;
  float ya[4], xa[4];
  cubic_taps_f32(dy, p.scale_y, &sy, ya);
  cubic_taps_f32(dx, p.scale_x, &sx, xa);
  const unsigned char* img = p.src + static_cast<long long>(b) * p.src_h * p.src_w * 3;
  const int sc = p.swap_rb ? 2 - c : c;
  float r[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int y = min(max(sy + k - 1, 0), p.src_h - 1);
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int x = min(max(sx + j - 1, 0), p.src_w - 1);
      const float v = __fdiv_rn(static_cast<float>(img[(static_cast<long long>(y) * p.src_w + x) * 3 + sc]), 255.f);
      const float t = __fmul_rn(v, xa[j]);
      acc = j == 0 ? t : __fadd_rn(acc, t);
    }
    r[k] = __fmul_rn(acc, ya[k]);
  }
  float v;
  if (dx * 3 + c >= p.dst_w * 3 - p.tail) v = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), r[2]), r[3]);
  else v = __fadd_rn(r[0], __fadd_rn(r[1], __fadd_rn(r[2], r[3])));
  const double z = __ddiv_rn(__dsub_rn(static_cast<double>(v), p.mean[c]), p.std[c]);
  p.out[((static_cast<long long>(b) * 3 + c) * p.dst_h + dy) * p.dst_w + dx] = __double2float_rn(z);
}

}  // namespace mde
