// Persistent, warp-specialised tcgen05 GEMM / implicit-GEMM 3x3 convolution for sm_100a.
//
//   D[M,N] = A[M,K] * B[N,K]^T      A, B 16-bit (fp16 or bf16) K-major, fp32 accumulation in TMEM.
//
// One CTA per SM loops over 128 x BLOCK_N output tiles.  Roles: warp 0 = TMA producer, warp 1 = MMA
// issuer (one thread), warp 2 = TMEM allocator, warps 4..7 = epilogue (TMEM -> registers -> fused
// epilogue -> global).  Two accumulator stages in TMEM let the epilogue of tile i overlap the MMAs of
// tile i+1.  Operand tiles are 128-byte-swizzled K-major boxes of 64 elements written by TMA.
//
// Conv mode: A is an NHWC activation addressed through a 4-D tensor map {C, W, H, B}; k-block kb maps
// to filter tap (kb / cin_blocks) and channel block (kb % cin_blocks); the box {64, tile_w, tile_h, 1}
// is fetched at (x0 + kx - 1, y0 + ky - 1) and TMA's out-of-bounds zero fill implements the padding.
#pragma once
#include "ptx.cuh"

namespace mde {

enum RowMap : int { ROW_IDENTITY = 0, ROW_TOKENS = 1, ROW_CONV = 2, ROW_SHUFFLE = 3 };
enum Act : int { ACT_NONE = 0, ACT_GELU = 1, ACT_RELU = 2 };

struct GemmParams {
  int M, N, K;          // logical problem; in conv mode M = B*H*W and K = 9 * cin_pad
  int num_k_blocks;     // ceil(K / 64)
  int m_tiles, n_tiles;
  // ---- A addressing
  int conv;             // 0: 2-D A map {K, M};  1: 4-D NHWC map, 3x3 taps
  int H, W;             // conv / shuffle: spatial size of the A-side map
  int tile_w, tile_h;   // conv: tile_w * tile_h <= 128 output pixels per M tile
  int tiles_x, tiles_y;
  int cin_blocks;       // conv: 64-channel blocks per tap
  // ---- epilogue
  int row_map;
  int tokens;           // ROW_TOKENS: T patch tokens per image (output rows per image = T + 1)
  int shuffle_s;        // ROW_SHUFFLE: ConvTranspose kernel == stride
  int shuffle_cout;     // ROW_SHUFFLE: N = s*s*cout, column = (ky*s + kx)*cout + o
  int act;
  int ld_out;           // row pitch (elements) of out / out_relu / res1 / res2 / x
  int accumulate_x;     // x = x + v (residual stream) instead of x = v
  const float* bias;    // [N] (ROW_SHUFFLE: [cout])
  const float* gamma;   // [N] LayerScale, applied after bias/activation
  const float* pos;     // ROW_TOKENS: [(T+1), ld_out] fp32, row 1+t is added
  float* x;             // fp32 output / residual stream
  const void* res1;     // 16-bit residuals added before the store
  const void* res2;
  void* out;            // 16-bit output
  void* out_relu;       // 16-bit relu(output) copy (input of the next pre-activation conv)
  // ---- fused depth head (BLOCK_N == 32 == N): z = sum_n relu(v_n) * head_w[n] + head_b
  const float* head_w;
  float head_b;
  float head_scale;     // metric: sigmoid(z) * head_scale; relative (head_scale < 0): relu(z)
  float* head_out;      // [M] fp32
};

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int kBlockM = 128;
  static constexpr int kBlockK = 64;
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kMaxStages = (226 * 1024 - 2048) / kStageBytes;
  static constexpr int kStages = kMaxStages > 8 ? 8 : kMaxStages;
  static constexpr int kTmemCols = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int kThreads = 256;
};

__device__ __forceinline__ float gelu_erf(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f)); }

template <int BLOCK_N, typename T>
__global__ void __launch_bounds__(256, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const GemmParams p) {
  using Cfg = GemmCfg<BLOCK_N>;
  using Tr = F16Traits<T>;
  constexpr int kStages = Cfg::kStages;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * Cfg::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tmem_full = bars + 2 * kStages;
  uint64_t* tmem_empty = bars + 2 * kStages + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.m_tiles * p.n_tiles;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);   // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      const uint32_t a_bytes = p.conv ? static_cast<uint32_t>(p.tile_w * p.tile_h * 128) : Cfg::kABytes;
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / p.n_tiles;
        const int n_blk = tile % p.n_tiles;
        int img = 0, y0 = 0, x0 = 0;
        if (p.conv) {
          const int per_img = p.tiles_x * p.tiles_y;
          img = m_blk / per_img;
          const int t = m_blk % per_img;
          y0 = (t / p.tiles_x) * p.tile_h;
          x0 = (t % p.tiles_x) * p.tile_w;
        }
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], a_bytes + Cfg::kBBytes);
          if (p.conv) {
            const int tap = kb / p.cin_blocks;
            const int c0 = (kb % p.cin_blocks) * 64;
            tma_load_4d(smem_a + stage * Cfg::kABytes, &map_a, &full_bar[stage], c0, x0 + tap % 3 - 1,
                        y0 + tap / 3 - 1, img);
          } else {
            tma_load_2d(smem_a + stage * Cfg::kABytes, &map_a, &full_bar[stage], kb * 64, m_blk * 128);
          }
          tma_load_2d(smem_b + stage * Cfg::kBBytes, &map_b, &full_bar[stage], kb * 64, n_blk * BLOCK_N);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(Tr::kFmt, 128, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t a_desc = umma_desc_k_sw128(smem_u32(smem_a + stage * Cfg::kABytes));
          const uint64_t b_desc = umma_desc_k_sw128(smem_u32(smem_b + stage * Cfg::kBBytes));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // +32 bytes per UMMA_K=16 step inside the 128-byte swizzle atom: +2 in the (addr >> 4) field
            tc_mma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
          }
          tc_commit(&empty_bar[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue
    const int ew = warp - 4;                 // == warp % 4: TMEM lane quarter this warp may read
    const int r = ew * 32 + lane;            // row inside the 128-row tile
    int acc = 0;
    uint32_t acc_phase = 0;
    T* out = static_cast<T*>(p.out);
    T* out_relu = static_cast<T*>(p.out_relu);
    const T* res1 = static_cast<const T*>(p.res1);
    const T* res2 = static_cast<const T*>(p.res2);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / p.n_tiles;
      const int n_blk = tile % p.n_tiles;
      // ---- where does this thread's row go?
      bool valid;
      long long orow;            // output row (ROW_SHUFFLE: row of sub-pixel (0,0))
      long long prow = 0;        // ROW_TOKENS: row of the pos_embed table
      if (p.conv) {
        const int per_img = p.tiles_x * p.tiles_y;
        const int img = m_blk / per_img;
        const int t = m_blk % per_img;
        const int y = (t / p.tiles_x) * p.tile_h + r / p.tile_w;
        const int x = (t % p.tiles_x) * p.tile_w + r % p.tile_w;
        valid = (r < p.tile_w * p.tile_h) && y < p.H && x < p.W;
        orow = (static_cast<long long>(img) * p.H + y) * p.W + x;
      } else {
        const long long m = static_cast<long long>(m_blk) * 128 + r;
        valid = m < p.M;
        orow = m;
        if (p.row_map == ROW_TOKENS) {
          const long long b = m / p.tokens;
          const long long t = m % p.tokens;
          orow = b * (p.tokens + 1) + 1 + t;
          prow = 1 + t;
        } else if (p.row_map == ROW_SHUFFLE) {
          const int hw = p.H * p.W;
          const long long b = m / hw;
          const int rem = static_cast<int>(m % hw);
          const int y = rem / p.W, x = rem % p.W;
          orow = (b * (p.H * p.shuffle_s) + static_cast<long long>(y) * p.shuffle_s) * (p.W * p.shuffle_s) +
                 static_cast<long long>(x) * p.shuffle_s;
        }
      }

      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_base = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * BLOCK_N;

#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        uint32_t raw[32];
        tmem_ld_32x32b_x32(t_base + c0, raw);
        tmem_ld_wait();
        const int n_base = n_blk * BLOCK_N + c0;
        if (p.head_w != nullptr) {
          // fused depth head: 3x3 conv (+bias, ReLU) -> 1x1 conv 32->1 -> activation
          float z = p.head_b;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float v = __uint_as_float(raw[j]) + __ldg(p.bias + n_base + j);
            z = fmaf(fmaxf(v, 0.f), __ldg(p.head_w + n_base + j), z);
          }
          if (valid) p.head_out[orow] = p.head_scale < 0.f ? fmaxf(z, 0.f) : p.head_scale / (1.0f + __expf(-z));
          continue;
        }
        if (!valid) continue;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int n = n_base + g * 8;
          if (n >= p.N) break;
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(raw[g * 8 + j]);
          long long row = orow;
          int col = n;
          if (p.row_map == ROW_SHUFFLE) {
            const int q = n / p.shuffle_cout;
            col = n % p.shuffle_cout;
            row = orow + static_cast<long long>(q / p.shuffle_s) * (p.W * p.shuffle_s) + (q % p.shuffle_s);
          }
          if (p.bias) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col + 4));
            v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
            v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
          }
          if (p.act == ACT_GELU) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = gelu_erf(v[j]);
          } else if (p.act == ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          if (p.gamma) {
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gamma + col));
            const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.gamma + col + 4));
            v[0] *= g0.x; v[1] *= g0.y; v[2] *= g0.z; v[3] *= g0.w;
            v[4] *= g1.x; v[5] *= g1.y; v[6] *= g1.z; v[7] *= g1.w;
          }
          const long long off = row * p.ld_out + col;
          if (p.pos) {
            const float4 q0 = __ldg(reinterpret_cast<const float4*>(p.pos + prow * p.ld_out + col));
            const float4 q1 = __ldg(reinterpret_cast<const float4*>(p.pos + prow * p.ld_out + col + 4));
            v[0] += q0.x; v[1] += q0.y; v[2] += q0.z; v[3] += q0.w;
            v[4] += q1.x; v[5] += q1.y; v[6] += q1.z; v[7] += q1.w;
          }
          if (res1) {
            const uint4 u = *reinterpret_cast<const uint4*>(res1 + off);
            const float2 a = Tr::unpack2(u.x), b = Tr::unpack2(u.y), c = Tr::unpack2(u.z), d = Tr::unpack2(u.w);
            v[0] += a.x; v[1] += a.y; v[2] += b.x; v[3] += b.y; v[4] += c.x; v[5] += c.y; v[6] += d.x; v[7] += d.y;
          }
          if (res2) {
            const uint4 u = *reinterpret_cast<const uint4*>(res2 + off);
            const float2 a = Tr::unpack2(u.x), b = Tr::unpack2(u.y), c = Tr::unpack2(u.z), d = Tr::unpack2(u.w);
            v[0] += a.x; v[1] += a.y; v[2] += b.x; v[3] += b.y; v[4] += c.x; v[5] += c.y; v[6] += d.x; v[7] += d.y;
          }
          if (p.x) {
            float4* xp = reinterpret_cast<float4*>(p.x + off);
            if (p.accumulate_x) {
              const float4 x0 = xp[0], x1 = xp[1];
              v[0] += x0.x; v[1] += x0.y; v[2] += x0.z; v[3] += x0.w;
              v[4] += x1.x; v[5] += x1.y; v[6] += x1.z; v[7] += x1.w;
            }
            xp[0] = make_float4(v[0], v[1], v[2], v[3]);
            xp[1] = make_float4(v[4], v[5], v[6], v[7]);
          }
          if (out) {
            uint4 u;
            u.x = Tr::pack2(v[0], v[1]); u.y = Tr::pack2(v[2], v[3]);
            u.z = Tr::pack2(v[4], v[5]); u.w = Tr::pack2(v[6], v[7]);
            *reinterpret_cast<uint4*>(out + off) = u;
          }
          if (out_relu) {
            uint4 u;
            u.x = Tr::pack2(fmaxf(v[0], 0.f), fmaxf(v[1], 0.f)); u.y = Tr::pack2(fmaxf(v[2], 0.f), fmaxf(v[3], 0.f));
            u.z = Tr::pack2(fmaxf(v[4], 0.f), fmaxf(v[5], 0.f)); u.w = Tr::pack2(fmaxf(v[6], 0.f), fmaxf(v[7], 0.f));
            *reinterpret_cast<uint4*>(out_relu + off) = u;
          }
        }
      }
      // all of this warp's TMEM reads of the stage are complete (wait::ld above): hand it back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace mde
