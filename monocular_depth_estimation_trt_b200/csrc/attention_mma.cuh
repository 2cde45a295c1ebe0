// Fused softmax(Q K^T * scale) V for the ViT encoder, head dim 64, non-causal, no mask other than the
// ragged last tile (N = 1370 tokens at 518^2; any N works).  Reads Q/K/V straight out of the QKV GEMM's
// [rows, 3D] output and writes [rows, D] in the layout the projection GEMM consumes, so there is no
// transpose anywhere.
//
// This variant keeps scores and the output accumulator in registers and uses warp-level mma.sync
// (m16n8k16, fp32 accumulate) with an online softmax; K/V tiles are double-buffered with cp.async.
// One CTA = 128 query rows of one (image, head); 8 warps x 16 rows.
#pragma once
#include "ptx.cuh"

namespace mde {

struct AttnParams {
  const void* qkv;   // [B*ntok, 3*D], 16-bit
  void* out;         // [B*ntok, D], 16-bit
  int ntok;          // keys per image (== tokens per image for self-attention)
  int heads;
  int D;             // heads * 64
  float scale_log2;  // head_dim^-0.5 * log2(e)
  // tcgen05 kernels only: queries and keys/values may come from different row sets (sequence-sharded global attention:
  // the queries are this rank's tokens, the keys/values every rank's, gathered into one [ntok, 2D] k|v buffer)
  int ntok_q;        // queries per image
  int k_col0, v_col0;// first column of K / V of head 0 in the key/value tensor (D and 2D in a packed q|k|v tensor)
  int batch;         // images (the persistent tcgen05 kernel decodes its work items itself)
  long long* trace;  // trace instantiation only (tools/attn_trace.py): clock64 stamps of the softmax warps' phases, else NULL
};

constexpr int kAttnBlockQ = 128;
constexpr int kAttnBlockKV = 64;
constexpr int kAttnSmemBytes = (kAttnBlockQ * 64 + 4 * kAttnBlockKV * 64) * 2;

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// element offset of (row, 16-byte chunk) in a [rows][64] 16-bit tile with an XOR-8 swizzle
__device__ __forceinline__ int swz(int row, int chunk) { return row * 64 + ((chunk ^ (row & 7)) << 3); }

template <typename T>
__global__ void __launch_bounds__(256, 2) attention_mma_kernel(const AttnParams p) {
  using Tr = F16Traits<T>;
  extern __shared__ __align__(128) uint8_t attn_smem[];
  T* sQ = reinterpret_cast<T*>(attn_smem);
  T* sK = sQ + kAttnBlockQ * 64;            // [2][64][64]
  T* sV = sK + 2 * kAttnBlockKV * 64;       // [2][64][64]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * kAttnBlockQ;
  const int head = blockIdx.y;
  const int img = blockIdx.z;
  const long long row_base = static_cast<long long>(img) * p.ntok;
  const int ld = 3 * p.D;
  const T* gQ = static_cast<const T*>(p.qkv) + row_base * ld + head * 64;
  const T* gK = gQ + p.D;
  const T* gV = gQ + 2 * p.D;

  // ---- async loads
  {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256, r = idx >> 3, c = idx & 7;
      const int n = q0 + r;
      const bool ok = n < p.ntok;
      cp_async_16(sQ + swz(r, c), gQ + static_cast<long long>(ok ? n : 0) * ld + c * 8, ok);
    }
  }
  auto load_kv = [&](int tile, int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + i * 256, r = idx >> 3, c = idx & 7;
      const int n = tile * kAttnBlockKV + r;
      const bool ok = n < p.ntok;
      const long long off = static_cast<long long>(ok ? n : 0) * ld + c * 8;
      cp_async_16(sK + buf * kAttnBlockKV * 64 + swz(r, c), gK + off, ok);
      cp_async_16(sV + buf * kAttnBlockKV * 64 + swz(r, c), gV + off, ok);
    }
  };
  const int nkv = (p.ntok + kAttnBlockKV - 1) / kAttnBlockKV;
  load_kv(0, 0);
  cp_async_commit();

  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};
  uint32_t qf[4][4];

  for (int j = 0; j < nkv; ++j) {
    const int buf = j & 1;
    if (j + 1 < nkv) {
      load_kv(j + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (j == 0) {
      // Q fragments stay in registers for the whole KV loop
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        ldmatrix_x4(qf[ks], smem_u32(sQ + swz(warp * 16 + (lane & 15), ks * 2 + (lane >> 4))));
    }
    const T* bK = sK + buf * kAttnBlockKV * 64;
    const T* bV = sV + buf * kAttnBlockKV * 64;

    // ---- S = Q K^T  (16 x 64 per warp)
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t kb[4];
        const int key = np * 16 + ((lane >> 4) << 3) + (lane & 7);
        ldmatrix_x4(kb, smem_u32(bK + swz(key, ks * 2 + ((lane >> 3) & 1))));
        mma_16816<T>(s[np * 2], qf[ks], kb[0], kb[1]);
        mma_16816<T>(s[np * 2 + 1], qf[ks], kb[2], kb[3]);
      }
    }
    // ---- mask the ragged tail
    const int kv0 = j * kAttnBlockKV;
    if (kv0 + kAttnBlockKV > p.ntok) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int key = kv0 + nt * 8 + (lane & 3) * 2;
        if (key >= p.ntok) s[nt][0] = s[nt][2] = -INFINITY;
        if (key + 1 >= p.ntok) s[nt][1] = s[nt][3] = -INFINITY;
      }
    }
    // ---- online softmax (rows g and g+8 of this warp's 16)
    float mx[2] = {m_run[0], m_run[1]};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
    }
    float corr[2], msc[2], rs[2] = {0.f, 0.f};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      corr[h] = fast_exp2((m_run[h] - mx[h]) * p.scale_log2);   // exp2(-inf) = 0 on the first tile
      msc[h] = mx[h] * p.scale_log2;
      m_run[h] = mx[h];
    }
    uint32_t pf[8][2];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = fast_exp2(fmaf(s[nt][0], p.scale_log2, -msc[0]));
      const float p1 = fast_exp2(fmaf(s[nt][1], p.scale_log2, -msc[0]));
      const float p2 = fast_exp2(fmaf(s[nt][2], p.scale_log2, -msc[1]));
      const float p3 = fast_exp2(fmaf(s[nt][3], p.scale_log2, -msc[1]));
      rs[0] += p0 + p1;
      rs[1] += p2 + p3;
      pf[nt][0] = Tr::pack2(p0, p1);
      pf[nt][1] = Tr::pack2(p2, p3);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) l_run[h] = l_run[h] * corr[h] + rs[h];
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      o[dt][0] *= corr[0]; o[dt][1] *= corr[0];
      o[dt][2] *= corr[1]; o[dt][3] *= corr[1];
    }
    // ---- O += P V
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const uint32_t a[4] = {pf[2 * ks][0], pf[2 * ks][1], pf[2 * ks + 1][0], pf[2 * ks + 1][1]};
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t vb[4];
        const int key = ks * 16 + (((lane >> 3) & 1) << 3) + (lane & 7);
        ldmatrix_x4_trans(vb, smem_u32(bV + swz(key, dp * 2 + (lane >> 4))));
        mma_16816<T>(o[dp * 2], a, vb[0], vb[1]);
        mma_16816<T>(o[dp * 2 + 1], a, vb[2], vb[3]);
      }
    }
    __syncthreads();
  }

  // ---- normalise, stage through this warp's slice of sQ, store coalesced
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 1);
    l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 2);
  }
  const float inv0 = 1.0f / l_run[0], inv1 = 1.0f / l_run[1];
  const int g = lane >> 2, t = lane & 3;
  T* sO = sQ + warp * 16 * 64;
#pragma unroll
  for (int dt = 0; dt < 8; ++dt) {
    const int e0 = (g * 64) + (((dt) ^ (g & 7)) << 3) + t * 2;
    const int e1 = ((g + 8) * 64) + (((dt) ^ ((g + 8) & 7)) << 3) + t * 2;
    *reinterpret_cast<uint32_t*>(sO + e0) = Tr::pack2(o[dt][0] * inv0, o[dt][1] * inv0);
    *reinterpret_cast<uint32_t*>(sO + e1) = Tr::pack2(o[dt][2] * inv1, o[dt][3] * inv1);
  }
  __syncwarp();
  T* gO = static_cast<T*>(p.out) + row_base * p.D + head * 64;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = lane + i * 32, r = idx >> 3, c = idx & 7;
    const int n = q0 + warp * 16 + r;
    if (n < p.ntok) {
      const uint4 v = *reinterpret_cast<const uint4*>(sO + r * 64 + ((c ^ (r & 7)) << 3));
      *reinterpret_cast<uint4*>(gO + static_cast<long long>(n) * p.D + c * 8) = v;
    }
  }
}

}  // namespace mde
