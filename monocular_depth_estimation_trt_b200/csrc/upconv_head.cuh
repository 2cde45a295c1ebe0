// Tail of the DPT head: bilinear upsample (align_corners=True) -> 3x3 conv 128->32 (+bias, ReLU) -> 1x1 conv 32->1
// -> activation, without ever materialising the upsampled map.
//
// Both the interpolation and the convolution are linear and nothing non-linear sits between output_conv1 and
// output_conv2[0], so the conv's channel contraction is moved in front of the interpolation:
//
//   z[b, y, x, t*32 + o] = sum_c W2[o, c, ky, kx] * o1[b, y, x, c]        t = ky*3 + kx   (a tensor-core GEMM at
//                                                                          the LOW resolution, N = 9*32 = 288)
//   acc[o](p)            = b2[o] + sum_t [p + d_t inside the map] * bilinear(z[.., t*32 + o], p + d_t)
//   depth(p)             = act(b3 + sum_o w3[o] * relu(acc[o](p)))
//
// which is the same real-arithmetic function as conv(upsample(o1)) with zero padding applied in the upsampled
// domain.  Compared with an implicit-GEMM conv at the full resolution (N = 32: the tensor core then spends its
// time re-reading a 128-wide A operand) this needs 9x less tensor work, no 518 x 518 x 128 intermediate in HBM
// (4.4 GB written and 4.4 GB read at batch 64) and leaves 1152 fp32 MACs per output pixel for the CUDA cores.
//
// One CTA = a 16 x 16 tile of output pixels of one image, one thread per pixel.  The z footprint of the tile
// (plus the one-pixel ring the 3x3 taps reach into) is staged in shared memory with cp.async; a pixel occupies
// 592 bytes (576 used) so that the 16-byte reads of eight neighbouring threads fall into distinct bank groups.
#pragma once
#include "ptx.cuh"

namespace mde {

constexpr int kUpTile = 16;
constexpr int kUpTaps = 9;
constexpr int kUpZc = kUpTaps * 32;        // 288 live channels of z
constexpr int kUpPixBytes = 592;           // smem pitch of one staged z pixel

struct UpconvHeadParams {
  const void* z;        // [B][Hs][Ws][ldz] 16-bit, first 288 channels live
  float* out;           // [B][Ho][Wo] fp32
  const float* bias;    // [32]  output_conv2[0].bias
  const float* head_w;  // [32]  output_conv2[2].weight
  float head_b;
  float head_scale;     // > 0: head_scale * sigmoid(v);  < 0: relu(v)
  int head_exp;         // 1: exp(v) (Depth Anything V3's depth branch) instead of the two above
  int B, Hs, Ws, Ho, Wo, ldz;
  float sy, sx;         // (Hs-1)/(Ho-1), (Ws-1)/(Wo-1)
  int fh, fw;           // rows / columns of z staged per tile (upper bound computed on the host)
};

template <typename T>
__global__ void __launch_bounds__(256, 2) upconv_head_kernel(const UpconvHeadParams p) {
  using Tr = F16Traits<T>;
  extern __shared__ __align__(16) uint8_t up_smem[];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int ox0 = blockIdx.x * kUpTile, oy0 = blockIdx.y * kUpTile, b = blockIdx.z;

  // ---- footprint of the tile in z: source cells touched by upsampled rows oy0-1 .. oy0+16 (clamped to the map)
  const int uy_lo = max(oy0 - 1, 0), uy_hi = min(oy0 + kUpTile, p.Ho - 1);
  const int ux_lo = max(ox0 - 1, 0), ux_hi = min(ox0 + kUpTile, p.Wo - 1);
  const int sy_lo = min(static_cast<int>(uy_lo * p.sy), p.Hs - 1);
  const int sx_lo = min(static_cast<int>(ux_lo * p.sx), p.Ws - 1);
  const int sy_hi = min(min(static_cast<int>(uy_hi * p.sy), p.Hs - 1) + 1, p.Hs - 1);
  const int sx_hi = min(min(static_cast<int>(ux_hi * p.sx), p.Ws - 1) + 1, p.Ws - 1);
  const int nrow = sy_hi - sy_lo + 1, ncol = sx_hi - sx_lo + 1;      // <= fh, fw
  {
    const T* zb = static_cast<const T*>(p.z) + static_cast<long long>(b) * p.Hs * p.Ws * p.ldz;
    const int chunks = nrow * ncol * (kUpZc / 8);
    for (int i = threadIdx.x; i < chunks; i += 256) {
      const int pix = i / (kUpZc / 8), ck = i - pix * (kUpZc / 8);
      const int py = pix / ncol, px = pix - py * ncol;
      cp_async_16(up_smem + (py * p.fw + px) * kUpPixBytes + ck * 16,
                  zb + (static_cast<long long>(sy_lo + py) * p.Ws + (sx_lo + px)) * p.ldz + ck * 8, true);
    }
    cp_async_commit();
  }

  // ---- per-thread geometry of the three tap rows / columns (while the copies are in flight)
  const int oy = oy0 + ty, ox = ox0 + tx;
  int yo0[3], yo1[3], xo0[3], xo1[3];      // smem offsets of the two source rows / columns
  float wy[3], wx[3];
  bool vy[3], vx[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int uy = oy + k - 1, ux = ox + k - 1;
    vy[k] = uy >= 0 && uy < p.Ho;
    vx[k] = ux >= 0 && ux < p.Wo;
    const float fy = max(uy, 0) * p.sy, fx = max(ux, 0) * p.sx;
    const int y0 = min(static_cast<int>(fy), p.Hs - 1), x0 = min(static_cast<int>(fx), p.Ws - 1);
    const int y1 = min(y0 + 1, p.Hs - 1), x1 = min(x0 + 1, p.Ws - 1);
    wy[k] = fy - y0;
    wx[k] = fx - x0;
    // clamp into the staged window: only reached by pixels of a partial tile, whose results are discarded
    yo0[k] = min(max(y0 - sy_lo, 0), nrow - 1) * p.fw * kUpPixBytes;
    yo1[k] = min(max(y1 - sy_lo, 0), nrow - 1) * p.fw * kUpPixBytes;
    xo0[k] = min(max(x0 - sx_lo, 0), ncol - 1) * kUpPixBytes;
    xo1[k] = min(max(x1 - sx_lo, 0), ncol - 1) * kUpPixBytes;
  }
  f32x2 acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = f2_pack(__ldg(p.bias + 2 * j), __ldg(p.bias + 2 * j + 1));

  cp_async_wait<0>();
  __syncthreads();

#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      if (!(vy[ky] && vx[kx])) continue;                 // zero padding of the conv, in the upsampled domain
      const int tap_off = (ky * 3 + kx) * 64;
      const float w11 = wy[ky] * wx[kx], w10 = wy[ky] - w11, w01 = wx[kx] - w11, w00 = 1.0f - wy[ky] - w01;
      const f32x2 ww[4] = {f2_splat(w00), f2_splat(w01), f2_splat(w10), f2_splat(w11)};
      const uint8_t* src[4] = {up_smem + yo0[ky] + xo0[kx] + tap_off, up_smem + yo0[ky] + xo1[kx] + tap_off,
                               up_smem + yo1[ky] + xo0[kx] + tap_off, up_smem + yo1[ky] + xo1[kx] + tap_off};
#pragma unroll
      for (int g = 0; g < 4; ++g) {
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          const uint4 q = *reinterpret_cast<const uint4*>(src[n] + g * 16);
          const uint32_t* u = &q.x;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 v = Tr::unpack2(u[k]);
            acc[g * 4 + k] = f2_fma(f2_pack(v.x, v.y), ww[n], acc[g * 4 + k]);
          }
        }
      }
    }
  }
  // ---- ReLU, 1x1 conv 32 -> 1, activation
  float v = p.head_b;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float a0, a1;
    f2_unpack(acc[j], a0, a1);
    v = fmaf(fmaxf(a0, 0.f), __ldg(p.head_w + 2 * j), v);
    v = fmaf(fmaxf(a1, 0.f), __ldg(p.head_w + 2 * j + 1), v);
  }
  if (oy < p.Ho && ox < p.Wo)
    p.out[(static_cast<long long>(b) * p.Ho + oy) * p.Wo + ox] = p.head_exp ? expf(v) : (p.head_scale < 0.f ? fmaxf(v, 0.f) : p.head_scale / (1.0f + __expf(-v)));
}

}  // namespace mde
