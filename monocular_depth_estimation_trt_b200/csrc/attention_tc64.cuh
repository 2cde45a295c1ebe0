// Fused softmax(Q K^T * scale) V on tcgen05, head dim 64 -- occupancy variant: FOUR CTAs per SM.
//
// attention_tc.cuh keeps two CTAs per SM, each overlapping its own S(j+1) MMA with the exponentials of tile j.
// Its softmax warps (two per SM sub-partition) still spend more than half of their time in the serial parts of a
// key tile (TMEM round trips, the max chain, barrier hand-offs), and neither the issue slots (47 %) nor the SFU
// (65 %) saturate.  This kernel trades the intra-CTA overlap for plain occupancy: key tiles of 64, 128 TMEM
// columns (S 64 | O 64, P overwrites S in place), 48 KB of shared memory and a 104-register softmax thread let
// four CTAs share an SM, so every sub-partition always has four softmax warps in different phases to pick from.
// Per CTA the chain S(j) -> softmax(j) -> P V(j) -> S(j+1) is strictly serial.
//   warps 0-3   softmax (thread = query row = TMEM lane)
//   warp 4      TMA producer: Q once, then K / V tiles of 64 keys through 2-stage rings
//   warp 5      MMA issuer (one thread)
// Lazy rescaling, probabilities <= 2^8, FMA-pipe exp2 for part of the elements and the trimmed last key tile are
// as in attention_tc.cuh.
#pragma once
#include <cuda/std/type_traits>

#include "attention_tc.cuh"

namespace mde {

constexpr int kA64Stages = 2;
constexpr int kA64TileBytes = 64 * 64 * 2;             // 64 keys x 64 dims, 16-bit
constexpr int kA64TmemCols = 128;                      // S / P [0,64)   O [64,128)
constexpr int kA64SmemBytes = kAtcQBytes + 2 * kA64Stages * kA64TileBytes + 256;   // 49 408 B: four fit in 227 KB

template <typename T, int kPoly>
__global__ void __launch_bounds__(kAtcThreads, 4)
attention_tc64_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv, const AttnParams p) {
  using Tr = F16Traits<T>;
  extern __shared__ __align__(1024) uint8_t a64_smem[];
  if ((smem_u32(a64_smem) & 1023u) != 0) __trap();
  uint8_t* sQ = a64_smem;
  uint8_t* sK = sQ + kAtcQBytes;
  uint8_t* sV = sK + kA64Stages * kA64TileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kA64Stages * kA64TileBytes);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = k_full + kA64Stages;
  uint64_t* v_full = k_empty + kA64Stages;
  uint64_t* v_empty = v_full + kA64Stages;
  uint64_t* s_full = v_empty + kA64Stages;   // S(j) complete (and with it P V(j-1): it was issued after o_full(j-1))
  uint64_t* p_ready = s_full + 1;            // P(j) in TMEM, O rescaled if needed (128 arrivals)
  uint64_t* o_full = p_ready + 1;            // O += P V(j) complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y, img = blockIdx.z;
  const int q0 = blockIdx.x * 128;
  const int nkv = (p.ntok + 63) / 64;
  const int last_chunks = (p.ntok - (nkv - 1) * 64 + 31) / 32;   // 1 or 2 live 32-key chunks in the last key tile
  const int row_base = img * p.ntok_q;     // first query row of this image
  const int kv_base = img * p.ntok;        // first key/value row of this image

  if (warp == 4 && lane == 0) {
    prefetch_tmap(&map_q);
    prefetch_tmap(&map_kv);
    mbar_init(q_full, 1);
    for (int i = 0; i < kA64Stages; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
    }
    mbar_init(s_full, 1); mbar_init(p_ready, 128); mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 5) {
    tmem_alloc(tmem_slot, kA64TmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ===================================================== TMA producer
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, kAtcQBytes);
      tma_load_2d(sQ, &map_q, q_full, head * 64, row_base + q0);
      for (int j = 0; j < nkv; ++j) {
        const int st = j % kA64Stages;
        const uint32_t ph = (j / kA64Stages) & 1;
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[st], kA64TileBytes);
        tma_load_2d(sK + st * kA64TileBytes, &map_kv, &k_full[st], p.k_col0 + head * 64, kv_base + j * 64);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[st], kA64TileBytes);
        tma_load_2d(sV + st * kA64TileBytes, &map_kv, &v_full[st], p.v_col0 + head * 64, kv_base + j * 64);
      }
    }
  } else if (warp == 5) {
    // ===================================================== MMA issuer
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
    if (lane == 0) {
      constexpr uint32_t idesc_o = umma_idesc_f16(Tr::kFmt, 128, 64) | (1u << 16);   // B (= V) is MN-major
      mbar_wait(q_full, 0);
      for (int j = 0; j < nkv; ++j) {
        const int st = j % kA64Stages;
        const uint32_t ph = (j / kA64Stages) & 1;
        const bool last = j == nkv - 1;
        // ---- S(j) = Q K_j^T into columns [0,64): P(j-1) lives there until P V(j-1) has completed
        mbar_wait(&k_full[st], ph);
        if (j > 0) mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after();
        {
          const uint32_t idesc_s = umma_idesc_f16(Tr::kFmt, 128, last ? last_chunks * 32 : 64);
          const uint64_t a = umma_desc_k_sw128(smem_u32(sQ));
          const uint64_t b = umma_desc_k_sw128(smem_u32(sK + st * kA64TileBytes));
#pragma unroll
          for (int k = 0; k < 4; ++k) tc_mma_f16(tmem_base, a + 2 * k, b + 2 * k, idesc_s, k != 0);
          tc_commit(s_full);
          tc_commit(&k_empty[st]);
        }
        // ---- O += P(j) V_j
        mbar_wait(&v_full[st], ph);
        mbar_wait(p_ready, j & 1);
        tc_fence_after();
        const uint64_t vb = umma_desc_mn_sw128(smem_u32(sV + st * kA64TileBytes));
        const int ksteps = last ? 2 * last_chunks : 4;
#pragma unroll
        for (int k = 0; k < 4; ++k)     // 16 keys per step: 8 packed P columns, two 8-row groups of V (2048 bytes)
          if (k < ksteps) tc_mma_f16_ts(tmem_base + 64, tmem_base + 8 * k, vb + 128 * k, idesc_o, (j | k) != 0);
        tc_commit(o_full);
        tc_commit(&v_empty[st]);
      }
    }
  } else if (warp >= 6) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
  } else {
    // ===================================================== softmax group (thread = query row)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int r = warp * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_base;
    const uint32_t o_addr = tmem_base + lane_base + 64;
    float m_ref = -INFINITY;
    float l_run = 0.f;
    const float sl = p.scale_log2;

    auto tile = [&](auto nch_tag, auto full_tag, int j) {
      constexpr bool kFull = decltype(full_tag)::value;
      constexpr int nch = decltype(nch_tag)::value;
      const int nvalid = kFull ? 64 : p.ntok - j * 64;
      uint32_t raw[2][32];
#pragma unroll
      for (int ch = 0; ch < 2; ++ch)
        if (ch < nch) tmem_ld_32x32b_x32(s_addr + ch * 32, raw[ch]);
      tmem_ld_wait();
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int ch = 0; ch < 2; ++ch)
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (ch < nch && (kFull || ch * 32 + i < nvalid)) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(raw[ch][i]));
      const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      const bool grow = (mx - m_ref) * sl > kAtcRescaleThreshold;
      const float msl_new = (grow ? mx : m_ref) * sl;
      f32x2 rs2[4] = {0ull, 0ull, 0ull, 0ull};
      uint32_t pk[32];
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        if (ch < nch) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const f32x2 xs = f2_fma(f2_pack(__uint_as_float(raw[ch][i]), __uint_as_float(raw[ch][i + 1])), f2_splat(sl), f2_splat(-msl_new));
            float p0, p1;
            if (((i >> 1) & 7) < kPoly) {
              exp2_fma2<Tr::kFmt == 1 ? 3 : 4>(xs, p0, p1);
            } else {
              float x0, x1;
              f2_unpack(xs, x0, x1);
              p0 = fast_exp2(x0);
              p1 = fast_exp2(x1);
            }
            if (!kFull) {
              if (ch * 32 + i >= nvalid) p0 = 0.f;
              if (ch * 32 + i + 1 >= nvalid) p1 = 0.f;
            }
            rs2[(i >> 1) & 3] = f2_add(rs2[(i >> 1) & 3], f2_pack(p0, p1));
            pk[ch * 16 + (i >> 1)] = Tr::pack2(p0, p1);
          }
        }
      }
      // P V(j-1) has completed (S(j) was only issued after it): O may be rescaled and S overwritten with P
      if (__any_sync(0xffffffffu, grow)) {
        const float factor = grow ? fast_exp2((m_ref - mx) * sl) : 1.0f;
        if (grow) { m_ref = mx; l_run *= factor; }
        if (j > 0) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t o[32];
            tmem_ld_32x32b_x32(o_addr + h * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
            tmem_st_32x32b_x32(o_addr + h * 32, o);
          }
        }
      }
      tmem_st_32x32b_x32(s_addr, pk);
      tmem_st_wait();
      {
        float a0, a1, b0, b1;
        f2_unpack(f2_add(rs2[0], rs2[1]), a0, a1);
        f2_unpack(f2_add(rs2[2], rs2[3]), b0, b1);
        l_run += (a0 + a1) + (b0 + b1);
      }
      tc_fence_before();
      mbar_arrive(p_ready);
    };

    for (int j = 0; j < nkv; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      using cuda::std::integral_constant;
      if (j * 64 + 64 <= p.ntok) tile(integral_constant<int, 2>{}, cuda::std::true_type{}, j);
      else if (last_chunks == 1) tile(integral_constant<int, 1>{}, cuda::std::false_type{}, j);
      else tile(integral_constant<int, 2>{}, cuda::std::false_type{}, j);
    }
    mbar_wait(o_full, (nkv - 1) & 1);
    tc_fence_after();
    const float inv = 1.0f / l_run;
    const int n = q0 + r;
    T* gout = static_cast<T*>(p.out) + (static_cast<long long>(row_base) + n) * p.D + head * 64;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint32_t o[32];
      tmem_ld_32x32b_x32(o_addr + h * 32, o);
      tmem_ld_wait();
      if (n < p.ntok_q) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 u;
          u.x = Tr::pack2(__uint_as_float(o[c * 8 + 0]) * inv, __uint_as_float(o[c * 8 + 1]) * inv);
          u.y = Tr::pack2(__uint_as_float(o[c * 8 + 2]) * inv, __uint_as_float(o[c * 8 + 3]) * inv);
          u.z = Tr::pack2(__uint_as_float(o[c * 8 + 4]) * inv, __uint_as_float(o[c * 8 + 5]) * inv);
          u.w = Tr::pack2(__uint_as_float(o[c * 8 + 6]) * inv, __uint_as_float(o[c * 8 + 7]) * inv);
          *reinterpret_cast<uint4*>(gout + h * 32 + c * 8) = u;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kA64TmemCols);
  }
}

}  // namespace mde
