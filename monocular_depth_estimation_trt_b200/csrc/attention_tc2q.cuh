// Fused softmax(Q K^T * scale) V on tcgen05, head dim 64 -- two query tiles per CTA with EXPLICIT ping-pong.
//
// attention_tc.cuh runs two independent CTAs per SM and hopes their phases interleave.  They do not: both softmax
// groups execute the same code, fall into step on the resource they share (the special-function unit, 16 ex2/clk/SM),
// and the SFU then idles while both are in their TMEM / max / barrier phases (measured: SFU 56-65 % busy, issue 47 %).
// Here ONE CTA per SM owns both 128-row query tiles and the two softmax warpgroups take turns through a pair of named
// barriers: while one is inside its exponential loop the other does everything else (pull the next scores out of
// TMEM, row maximum, wait for the previous P V, write P back, hand-shakes).  K and V tiles are loaded once for both
// query tiles.
//   warps 0-3   softmax group 0 (query rows q0 .. q0+127;     S0 / O0 / P0)
//   warps 4-7   softmax group 1 (query rows q0+128 .. q0+255; S1 / O1 / P1)
//   warp 8      TMA producer: both Q tiles once, then K and V tiles through 3-stage rings
//   warp 9      MMA issuer (one thread): S0(j+1), S1(j+1), P0 V(j), P1 V(j) in a fixed order
// TMEM (512 columns): S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384) P0 [384,448) P1 [448,512).
// Everything else (lazy rescale at 2^8, FMA-pipe exp2 share, trimmed last key tile, separate query / key row sets)
// is as in attention_tc.cuh.
#pragma once
#include <cuda/std/type_traits>

#include "attention_tc.cuh"

namespace mde {

constexpr int kA2Threads = 384;
constexpr int kA2Stages = 3;
constexpr int kA2SmemBytes = 2 * kAtcQBytes + 2 * kA2Stages * kAtcQBytes + 256;   // 131 328 B, one CTA per SM

__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

template <typename T, int kPoly>
__global__ void __launch_bounds__(kA2Threads, 1)
attention_tc2q_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv, const AttnParams p) {
  using Tr = F16Traits<T>;
  extern __shared__ __align__(1024) uint8_t a2_smem[];
  if ((smem_u32(a2_smem) & 1023u) != 0) __trap();
  uint8_t* sQ = a2_smem;                                   // two tiles
  uint8_t* sK = sQ + 2 * kAtcQBytes;
  uint8_t* sV = sK + kA2Stages * kAtcQBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kA2Stages * kAtcQBytes);
  uint64_t* q_full = bars;                     // [1]
  uint64_t* k_full = bars + 1;                 // [stages]
  uint64_t* k_empty = k_full + kA2Stages;
  uint64_t* v_full = k_empty + kA2Stages;
  uint64_t* v_empty = v_full + kA2Stages;
  uint64_t* s_full = v_empty + kA2Stages;      // [2] S_g ready in TMEM
  uint64_t* s_free = s_full + 2;               // [2] S_g copied to registers (128 arrivals)
  uint64_t* p_ready = s_free + 2;              // [2] P_g in TMEM, O_g rescaled if needed (128 arrivals)
  uint64_t* o_full = p_ready + 2;              // [2] O_g += P_g V_j complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y, img = blockIdx.z;
  const int q0 = blockIdx.x * 256;
  const int nkv = (p.ntok + 127) / 128;
  const int last_chunks = (p.ntok - (nkv - 1) * 128 + 31) / 32;
  const int row_base = img * p.ntok_q;
  const int kv_base = img * p.ntok;

  if (warp == 8 && lane == 0) {
    prefetch_tmap(&map_q);
    prefetch_tmap(&map_kv);
    mbar_init(q_full, 1);
    for (int i = 0; i < kA2Stages; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1); mbar_init(&s_free[g], 128); mbar_init(&p_ready[g], 128); mbar_init(&o_full[g], 1);
    }
    fence_mbar_init();
  }
  if (warp == 9) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    // ===================================================== TMA producer
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, 2 * kAtcQBytes);
      tma_load_2d(sQ, &map_q, q_full, head * 64, row_base + q0);
      tma_load_2d(sQ + kAtcQBytes, &map_q, q_full, head * 64, row_base + q0 + 128);
      for (int j = 0; j < nkv; ++j) {
        const int st = j % kA2Stages;
        const uint32_t ph = (j / kA2Stages) & 1;
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[st], kAtcQBytes);
        tma_load_2d(sK + st * kAtcQBytes, &map_kv, &k_full[st], p.k_col0 + head * 64, kv_base + j * 128);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[st], kAtcQBytes);
        tma_load_2d(sV + st * kAtcQBytes, &map_kv, &v_full[st], p.v_col0 + head * 64, kv_base + j * 128);
      }
    }
  } else if (warp == 9) {
    // ===================================================== MMA issuer
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (lane == 0) {
      constexpr uint32_t idesc_o = umma_idesc_f16(Tr::kFmt, 128, 64) | (1u << 16);   // B (= V) is MN-major
      auto issue_s = [&](int g, int st, int j) {
        const uint32_t idesc_s = umma_idesc_f16(Tr::kFmt, 128, j == nkv - 1 ? last_chunks * 32 : 128);
        const uint64_t a = umma_desc_k_sw128(smem_u32(sQ + g * kAtcQBytes));
        const uint64_t b = umma_desc_k_sw128(smem_u32(sK + st * kAtcQBytes));
#pragma unroll
        for (int k = 0; k < 4; ++k) tc_mma_f16(tmem_base + g * 128, a + 2 * k, b + 2 * k, idesc_s, k != 0);
        tc_commit(&s_full[g]);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_s(0, 0, 0);
      issue_s(1, 0, 0);
      tc_commit(&k_empty[0]);
      for (int j = 0; j < nkv; ++j) {
        const int st = j % kA2Stages;
        if (j + 1 < nkv) {
          const int st1 = (j + 1) % kA2Stages;
          mbar_wait(&k_full[st1], ((j + 1) / kA2Stages) & 1);
          mbar_wait(&s_free[0], j & 1);
          tc_fence_after();
          issue_s(0, st1, j + 1);
          mbar_wait(&s_free[1], j & 1);
          tc_fence_after();
          issue_s(1, st1, j + 1);
          tc_commit(&k_empty[st1]);
        }
        mbar_wait(&v_full[st], (j / kA2Stages) & 1);
        const uint64_t vb = umma_desc_mn_sw128(smem_u32(sV + st * kAtcQBytes));
        const int ksteps = j == nkv - 1 ? 2 * last_chunks : 8;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          mbar_wait(&p_ready[g], j & 1);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (k < ksteps) tc_mma_f16_ts(tmem_base + 256 + g * 64, tmem_base + 384 + g * 64 + 8 * k, vb + 128 * k, idesc_o, (j | k) != 0);
          tc_commit(&o_full[g]);
        }
        tc_commit(&v_empty[st]);
      }
    }
  } else if (warp >= 10) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  } else {
    // ===================================================== softmax groups (thread = query row)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    const int g = warp >> 2;                                  // group 0 / 1
    const int r = (warp & 3) * 32 + lane;                     // row inside the group's tile
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_base + g * 128;
    const uint32_t o_addr = tmem_base + lane_base + 256 + g * 64;
    const uint32_t p_addr = tmem_base + lane_base + 384 + g * 64;
    // turn-taking on the SFU: group g waits on barrier 1+g for its turn and grants barrier 2-g when its loop is done
    const int bar_mine = 1 + g, bar_other = 2 - g;
    if (g == 1) named_bar_arrive(1, 256);                     // group 0 goes first
    float m_ref = -INFINITY;
    float l_run = 0.f;
    const float sl = p.scale_log2;

    auto tile = [&](auto nch_tag, auto full_tag, int j) {
      constexpr bool kFull = decltype(full_tag)::value;
      constexpr int nch = decltype(nch_tag)::value;
      const int nvalid = kFull ? 128 : p.ntok - j * 128;
      uint32_t raw[4][32];
#pragma unroll
      for (int ch = 0; ch < 4; ++ch)
        if (ch < nch) tmem_ld_32x32b_x32(s_addr + ch * 32, raw[ch]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&s_free[g]);
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int ch = 0; ch < 4; ++ch)
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (ch < nch && (kFull || ch * 32 + i < nvalid)) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(raw[ch][i]));
      const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      const bool grow = (mx - m_ref) * sl > kAtcRescaleThreshold;
      const float msl_new = (grow ? mx : m_ref) * sl;
      f32x2 rs2[4] = {0ull, 0ull, 0ull, 0ull};
      uint32_t pk[2][32];
      named_bar_sync(bar_mine, 256);                          // ---- our turn on the SFU
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
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
            pk[ch >> 1][(ch & 1) * 16 + (i >> 1)] = Tr::pack2(p0, p1);
          }
        }
      }
      if (!(g == 1 && j == nkv - 1)) named_bar_arrive(bar_other, 256);   // ---- the other group's turn (no dangling arrival at the end)
      if (j > 0) {
        mbar_wait(&o_full[g], (j - 1) & 1);
        tc_fence_after();
      }
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
      tmem_st_32x32b_x32(p_addr, pk[0]);
      if (nch > 2) tmem_st_32x32b_x32(p_addr + 32, pk[1]);
      tmem_st_wait();
      {
        float a0, a1, b0, b1;
        f2_unpack(f2_add(rs2[0], rs2[1]), a0, a1);
        f2_unpack(f2_add(rs2[2], rs2[3]), b0, b1);
        l_run += (a0 + a1) + (b0 + b1);
      }
      tc_fence_before();
      mbar_arrive(&p_ready[g]);
    };

    for (int j = 0; j < nkv; ++j) {
      mbar_wait(&s_full[g], j & 1);
      tc_fence_after();
      using cuda::std::integral_constant;
      if (j * 128 + 128 <= p.ntok) tile(integral_constant<int, 4>{}, cuda::std::true_type{}, j);
      else if (last_chunks == 1) tile(integral_constant<int, 1>{}, cuda::std::false_type{}, j);
      else if (last_chunks == 2) tile(integral_constant<int, 2>{}, cuda::std::false_type{}, j);
      else if (last_chunks == 3) tile(integral_constant<int, 3>{}, cuda::std::false_type{}, j);
      else tile(integral_constant<int, 4>{}, cuda::std::false_type{}, j);
    }
    mbar_wait(&o_full[g], (nkv - 1) & 1);
    tc_fence_after();
    const float inv = 1.0f / l_run;
    const int n = q0 + g * 128 + r;
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
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace mde
